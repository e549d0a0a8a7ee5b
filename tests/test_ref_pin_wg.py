"""Pins the waveguide oracle (oracle/wg_oracle.cpp) to the REFERENCE'S OWN kernel source.

oracle/_ref/lib_ref.so is compiled by oracle/ref_recipe/build.py from the OpenCL-C raw strings
of /root/reference (src/waveguide/src/program.cpp:11-531, cl/utils.cpp:8-77, cl/filters.cpp:17-75
and the cl_representation struct strings) -- verbatim apart from three documented syntax
rewrites. These tests step that code and the oracle on identical meshes and assert BIT
IDENTITY of pressures, filter memories and error flags:

  * `float` mode  : the reference's own arithmetic types (float pressures, double filters)
  * `double` mode : the same source built with float -> double (how an fp64 build reads),
                    against the oracle mode the CUDA kernels are compared with.

CPU only; on the GPU box (no /root/reference) the prebuilt library that travelled with the
snapshot is used."""
import json
import os

import numpy as np
import pytest

from oracle import refk, wgo

pytestmark = pytest.mark.skipif(not refk.available(), reason="no /root/reference and no prebuilt oracle/_ref")

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lrs_coefficients.json")
MODES = ["float", "double"]


def golden_coeffs(i=0):
    s = json.load(open(GOLD))["sets"][i]["impedance"]
    c = np.zeros((), wgo.COEFF_DT)
    c["b"], c["a"] = s["b"], s["a"]
    return c


def same_bits(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def step_both(om, mode, steps, writes, check_every=None):
    o = wgo.Sim(om, mode)
    r = refk.Sim(om, mode)
    for node, v in writes:
        o.write(node, v)
        r.write(node, v)
    fo = fr = 0
    done = 0
    while done < steps:
        k = min(check_every or steps, steps - done)
        fo |= o.step(k)
        fr |= r.step(k)
        done += k
        assert same_bits(o.field(), r.field()), "fields differ after %d steps (%s)" % (done, mode)
    for n in (1, 2, 3):
        bo, br = o.boundary_data(n), r.boundary_data(n)
        assert same_bits(bo["mem"], br["mem"]), "filter memory %d-d (%s)" % (n, mode)
        assert np.array_equal(bo["coefficient_index"], br["coefficient_index"])
    assert fo == fr
    return o, r, fo


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("dims", [(24, 18, 14), (5, 5, 5), (33, 27, 22)])
def test_cuboid_plaster(dims, mode):
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [golden_coeffs(0)])
    src = om.index(dims[0] // 2, dims[1] // 2, dims[2] // 2)
    o, r, flag = step_both(om, mode, 80, [(src, 1.0)], check_every=20)
    assert flag == 0
    assert np.abs(o.field()).max() > 0
    assert np.abs(o.boundary_data(1)["mem"]).max() > 0 or min(dims) <= 5


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("absorption", [0.0, 0.1, 0.9])
def test_cuboid_flat_coefficients(absorption, mode):
    """to_flat_coefficients surfaces (fitted_boundary.h:72-75); absorption 0 gives a0 == 0,
    the case the `== 0 ? 0 :` guards of filter_step exist for."""
    dims = (20, 16, 12)
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [wgo.to_flat(absorption)])
    step_both(om, mode, 60, [(om.index(9, 8, 6), 1.0), (om.index(4, 4, 4), -0.25)], check_every=15)


@pytest.mark.parametrize("mode", MODES)
def test_l_shaped_room_three_surfaces(mode):
    """reentrant nodes, three surfaces, 2-d / 3-d nodes with mixed coefficient sets -- the mesh of
    tests/test_wg_gpu.py::test_l_shaped_room_three_surfaces at a CPU-friendly size."""
    dz, dy, dx = 12, 22, 40
    ins = np.zeros((dz, dy, dx), bool)
    ins[2:dz - 2, 2:dy - 2, 2:18] = True
    ins[2:dz - 2, 2:10, 2:dx - 2] = True
    zz, yy, xx = np.indices(ins.shape)
    surf = ((xx > 15).astype(np.uint32) + (yy > 8).astype(np.uint32)).ravel()
    om = wgo.mesh_from_inside(ins, [golden_coeffs(0), golden_coeffs(1), golden_coeffs(2)], surf)
    assert (om.nodes["boundary_type"] == wgo.ID_REENTRANT).any()
    assert om.b2.shape[0] and om.b3.shape[0]
    o, r, flag = step_both(om, mode, 100, [(om.index(8, 6, 6), 1.0)], check_every=25)
    assert flag == 0


@pytest.mark.parametrize("mode", MODES)
def test_error_flags(mode):
    """The meshes of test_error_flags_match_reference_semantics: a boundary node at the mesh
    edge (id_outside_mesh_error), a 1-d node whose in-plane neighbour is air
    (id_suspicious_boundary_error), and an inf / nan excitation."""
    dims = (12, 10, 8)
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [wgo.to_flat(0.2)])
    nodes = om.nodes.copy()
    # a 1-d boundary node in the outermost layer: its in-plane ports leave the mesh
    nodes["boundary_type"][om.index(0, 5, 4)] = wgo.ID_PZ
    bad = wgo.Mesh(om.dims, nodes, om.coeffs, om.b1, om.b2, om.b3)
    _, _, flag = step_both(bad, mode, 3, [(bad.index(5, 5, 4), 1.0)], check_every=1)
    assert flag & wgo.ERR_OUTSIDE_MESH

    om2 = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [wgo.to_flat(0.2)])
    nodes = om2.nodes.copy()
    # turn one face node's in-plane neighbour into an inside node: "suspicious" for its neighbours
    victim = om2.index(5, 5, 1)
    assert nodes["boundary_type"][victim] not in (wgo.ID_NONE, wgo.ID_INSIDE)
    nodes["boundary_type"][victim] = wgo.ID_INSIDE
    om2 = wgo.Mesh(om2.dims, nodes, om2.coeffs, om2.b1, om2.b2, om2.b3)
    _, _, flag = step_both(om2, mode, 3, [(om2.index(5, 5, 4), 1.0)], check_every=1)
    assert flag & wgo.ERR_SUSPICIOUS

    om3 = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [wgo.to_flat(0.2)])
    for bad, bit in ((np.inf, wgo.ERR_INF), (np.nan, wgo.ERR_NAN)):
        o = wgo.Sim(om3, mode)
        r = refk.Sim(om3, mode)
        o.write(om3.index(5, 5, 4), bad)
        r.write(om3.index(5, 5, 4), bad)
        fo, fr = o.step(1), r.step(1)
        assert fo == fr and fo & bit
        assert same_bits(o.field(), r.field())


@pytest.mark.parametrize("mode", MODES)
def test_random_field_all_node_classes(mode):
    """A dense random pressure field (every node non-zero, incl. id_none nodes, which the kernel
    overwrites with 0) exercises every port of every node class at once."""
    dims = (17, 13, 11)
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [golden_coeffs(2)])
    rng = np.random.default_rng(5)
    f = rng.uniform(-1, 1, om.num_nodes).astype(np.float32).astype(np.float64)
    o, r = wgo.Sim(om, mode), refk.Sim(om, mode)
    o.set_field(f)
    r.set_field(f)
    for s in range(30):
        assert o.step(1) == r.step(1)
        assert same_bits(o.field(), r.field()), s
    for n in (1, 2, 3):
        assert same_bits(o.boundary_data(n)["mem"], r.boundary_data(n)["mem"])


@pytest.mark.parametrize("mode", MODES)
def test_run_loop_hard_and_soft_source(mode):
    """waveguide::run with hard_source / soft_source + postprocessor::node: the receiver trace
    of the reference-kernel loop equals the oracle's run()."""
    dims = (22, 18, 16)
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [golden_coeffs(1)])
    src, rcv = om.index(8, 9, 7), [om.index(14, 9, 8), om.index(8, 9, 7)]
    sig = np.zeros(120)
    sig[0] = 1.0
    sig[3] = -0.5
    for soft in (False, True):
        so, out_o, fo = wgo.Sim(om, mode).run(src, sig, rcv, soft=soft)
        sr, out_r, fr = refk.Sim(om, mode).run(src, sig, rcv, soft=soft)
        assert (so, fo) == (sr, fr) == (120, 0)
        assert same_bits(out_o, out_r)


def test_filter_kernels_match_oracle():
    """filter_test / filter_test_2 (cl/filters.cpp:56-75) against the oracle's filter step,
    random peak-filter sets as in tests/rectangular_kernel.cpp:242-305."""
    rng = np.random.default_rng(3)
    for _ in range(4):
        biq = np.stack([wgo.peak_biquad(rng.uniform(0.1, 1), rng.uniform(0, 0.5), rng.uniform(0, 1))
                        for _ in range(3)])
        canon = wgo.convolve3(biq)
        x = rng.uniform(-0.25, 0.25, 300).astype(np.float32)
        assert same_bits(refk.filter_biquads(biq, x), wgo.filter_biquads(biq, x.copy()))
        assert same_bits(refk.filter_canonical(canon, x), wgo.filter_canonical(canon, x.copy()))
        xd = x.astype(np.float64)
        assert same_bits(refk.filter_canonical(canon, xd, "double"), wgo.filter_canonical(canon, xd.copy(), f64=True))


def test_index_helpers_round_trip():
    """to_locator / neighbor_index (cl/utils.cpp:25-69) vs the closed forms the CUDA layout relies on."""
    dims = (7, 5, 4)
    n = dims[0] * dims[1] * dims[2]
    for i in range(n):
        x, y, z = refk.to_locator(i, dims)
        assert (x, y, z) == (i % 7, (i // 7) % 5, i // 35)
        for port, (ox, oy, oz) in enumerate([(-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)]):
            nx, ny, nz = x + ox, y + oy, z + oz
            inside = 0 <= nx < 7 and 0 <= ny < 5 and 0 <= nz < 4
            want = nx + ny * 7 + nz * 35 if inside else 0xFFFFFFFF
            assert refk.neighbor_index((x, y, z), dims, port) == want
