"""Pins the CPU oracle (oracle/wg_oracle.cpp) against the reference's own
known-answer tests and golden vectors, on CPU.

reference tests mirrored here:
  src/waveguide/tests/waveguide_tests.cpp:30-41      peak coefficients, gain 0 => b == a
  src/waveguide/tests/rectangular_kernel.cpp:307-360 biquad cascade == convolved canonical (<1e-3)
  src/waveguide/tests/verify_compensation_signal.cpp:23-91  run-to-run bit determinism
  bin/boundary_test/output.soft/coefficients.txt     reflectance -> impedance golden sets
"""
import json
import os

import numpy as np
import pytest

from oracle import wgo


def load_sets(golden_dir):
    return json.load(open(os.path.join(golden_dir, "lrs_coefficients.json")))["sets"]


def as_coeffs(d):
    c = np.zeros((), wgo.COEFF_DT)
    c["b"] = d["b"]
    c["a"] = d["a"]
    return c


def test_golden_impedance_sets(golden_dir):
    # fitted_boundary.h:35-48 applied to the checked-in reflectance filters must
    # give the checked-in impedance filters (cereal prints 17 significant digits)
    sets = load_sets(golden_dir)
    assert len(sets) == 9
    for s in sets:
        got = wgo.to_impedance(as_coeffs(s["reflectance"]))
        np.testing.assert_allclose(got["b"], s["impedance"]["b"], rtol=1e-14, atol=1e-15)
        np.testing.assert_allclose(got["a"], s["impedance"]["a"], rtol=1e-14, atol=1e-15)
    # survey's spot value: plaster impedance b0
    assert abs(sets[0]["impedance"]["b"][0] - 3.2629165258739678) < 1e-15


def test_flat_coefficients():
    rigid = wgo.to_flat(0.0)
    assert rigid["b"].tolist() == [2, 0, 0, 0, 0, 0, 0]
    assert rigid["a"].tolist() == [0, 0, 0, 0, 0, 0, 0]  # a0 == 0: left un-normalised
    c = wgo.to_flat(0.1)
    r = np.sqrt(0.9)
    assert c["a"][0] == 1.0
    assert abs(c["b"][0] - (1 + r) / (1 - r)) < 1e-12
    assert not c["b"][1:].any() and not c["a"][1:].any()


def test_peak_coefficients_gain_zero():
    rng = np.random.default_rng(1)
    for _ in range(10):
        pk = wgo.peak_biquad(0.0, rng.uniform(0, 0.5), 1.414)
        assert np.array_equal(pk[:3], pk[3:])


@pytest.mark.parametrize("kind", ["impulse", "noise"])
def test_biquad_cascade_equals_canonical(kind):
    rng = np.random.default_rng(7)
    n = 200 if kind == "impulse" else 10000
    for _ in range(16):
        biq = np.stack([wgo.peak_biquad(rng.uniform(0.1, 1), rng.uniform(0, 0.5), rng.uniform(0, 1))
                        for _ in range(3)])
        canon = wgo.convolve3(biq)
        if kind == "impulse":
            x = np.zeros(n, np.float32)
            x[0] = 0.25
        else:
            x = rng.uniform(-0.25, 0.25, n).astype(np.float32)
        y1 = wgo.filter_biquads(biq, x)
        y2 = wgo.filter_canonical(canon, x)
        assert np.isfinite(y1).all() and np.isfinite(y2).all()
        assert np.abs(y1 - y2).max() < 1e-3


def test_rigid_filter_memory_stays_zero():
    # a0 == 0 makes `out` inf/nan, the ==0 guards keep memory at 0
    mem = np.zeros(6)
    wgo.filter_canonical(wgo.to_flat(0.0), np.ones(16, np.float32), mem)
    assert not mem.any()


def small_box(dims=(14, 12, 10), absorption=0.1):
    return wgo.mesh_from_inside(wgo.cuboid_inside(dims), [wgo.to_flat(absorption)])


def test_cuboid_classification_counts():
    dims = (14, 12, 10)
    m = small_box(dims)
    bt = m.nodes["boundary_type"].reshape(dims[2], dims[1], dims[0])
    n = [d - 4 for d in dims]  # inside extents
    pop = np.array([bin(int(v)).count("1") for v in bt.ravel()]).reshape(bt.shape)
    inside = bt == wgo.ID_INSIDE
    assert inside.sum() == n[0] * n[1] * n[2]
    assert ((pop == 1) & ~inside).sum() == 2 * (n[0] * n[1] + n[1] * n[2] + n[0] * n[2])
    assert (pop == 2).sum() == 4 * (n[0] + n[1] + n[2])
    assert (pop == 3).sum() == 8
    assert m.b1.shape[0] == ((pop == 1) & ~inside).sum()
    assert m.b2.shape[0] == (pop == 2).sum() and m.b3.shape[0] == 8
    # outermost layer is id_none; direction bit names the side of the inner node
    assert not bt[0].any() and not bt[:, 0].any() and not bt[:, :, 0].any()
    assert bt[5, 5, 1] == wgo.ID_PX and bt[5, 5, dims[0] - 2] == wgo.ID_NX
    assert bt[1, 5, 5] == wgo.ID_PZ and bt[5, 1, 5] == wgo.ID_PY
    assert bt[1, 1, 1] == (wgo.ID_PX | wgo.ID_PY | wgo.ID_PZ)
    # boundary_index is a running count in node order per class
    idx1 = m.nodes["boundary_index"][((pop == 1) & ~inside).ravel()]
    assert np.array_equal(idx1, np.arange(idx1.size))


def test_reentrant_nodes_in_l_shape():
    ins = np.zeros((12, 14, 14), bool)
    ins[2:10, 2:12, 2:7] = True
    ins[2:10, 2:7, 2:12] = True
    m = wgo.mesh_from_inside(ins, [wgo.to_flat(0.2)])
    bt = m.nodes["boundary_type"]
    assert (bt == wgo.ID_REENTRANT).sum() > 0
    sim = wgo.Sim(m)
    sim.write(m.index(4, 4, 5), 1.0)
    assert sim.step(50) == 0
    assert np.isfinite(sim.field()).all()


def test_free_field_first_steps_known_answer():
    # impulse 1 at the centre, nothing else: p1 = 1/3 on the 6 neighbours;
    # p2[c] = (6 * 1/3) / 3 - 1 = -1/3, p2[+2x] = 1/9, p2[+x+y] = 2/9
    m = small_box((17, 17, 17), 0.0)
    for mode, tol in (("double", 1e-15), ("float", 1e-6)):
        sim = wgo.Sim(m, mode)
        c = m.index(8, 8, 8)
        sim.write(c, 1.0)
        assert sim.step(1) == 0
        assert abs(sim.read(m.index(9, 8, 8)) - 1 / 3) <= tol
        assert sim.read(c) == 0.0
        assert sim.step(1) == 0
        assert abs(sim.read(c) + 1 / 3) <= tol
        assert abs(sim.read(m.index(10, 8, 8)) - 1 / 9) <= tol
        assert abs(sim.read(m.index(9, 9, 8)) - 2 / 9) <= tol


def test_hard_source_pins_node_and_post_sees_current():
    # SURVEY 3.4(1) + 3.2: output sample n is p(n) including the injected source
    m = small_box((13, 13, 13), 0.1)
    sim = wgo.Sim(m)
    src = m.index(6, 6, 6)
    sig = np.zeros(20)
    sig[0] = 1.0
    steps, out, flag = sim.run(src, sig, [src, m.index(7, 6, 6)])
    assert steps == 20 and flag == 0
    assert out[0, 0] == 1.0 and not out[1:, 0].any()  # pinned to the input
    assert out[0, 1] == 0.0 and abs(out[1, 1] - 1 / 3) < 1e-15


def test_bit_determinism():
    m = small_box((20, 18, 16), 0.5)
    outs = []
    for _ in range(3):
        sim = wgo.Sim(m)
        src = m.index(9, 9, 8)
        sig = np.zeros(100)
        sig[0] = 1.0
        _, out, flag = sim.run(src, sig, [m.index(5, 6, 7)])
        assert flag == 0
        outs.append((out.copy(), sim.field()))
    for o, f in outs[1:]:
        assert np.array_equal(o, outs[0][0]) and np.array_equal(f, outs[0][1])


def test_rigid_wall_equals_mirror_image():
    # a rigid LRS wall (a0 = 0) must act as a perfect mirror: the field in a
    # half-space with a rigid wall equals the field of source + image source in
    # free space, as long as no other wall is reached (1 node per step).
    steps = 9
    big = 2 * steps + 8
    half = wgo.cuboid_inside((big, big, big))
    m_free = wgo.mesh_from_inside(half, [wgo.to_flat(0.0)])
    # wall: everything with x < wx is outside
    wx = big // 2
    walled = half.copy()
    walled[:, :, :wx] = False
    m_wall = wgo.mesh_from_inside(walled, [wgo.to_flat(0.0)])
    d = 3  # source distance from the boundary-node plane x = wx - 1
    sy = sz = big // 2
    a, b = wgo.Sim(m_wall), wgo.Sim(m_free)
    a.write(m_wall.index(wx - 1 + d, sy, sz), 1.0)
    b.write(m_free.index(wx - 1 + d, sy, sz), 1.0)
    b.write(m_free.index(wx - 1 - d, sy, sz), 1.0)
    assert a.step(steps) == 0 and b.step(steps) == 0
    fa = a.field().reshape(big, big, big)[:, :, wx - 1:]
    fb = b.field().reshape(big, big, big)[:, :, wx - 1:]
    assert np.abs(fa).max() > 1e-3
    assert np.abs(fa - fb).max() < 1e-14


def test_rigid_box_is_lossless():
    # bin/solution_growth's concern. At the Courant limit 1/sqrt(3) the scheme
    # has marginally stable DC and (pi,pi,pi)-checkerboard modes that grow
    # linearly when excited (a property of the scheme, measured here with a
    # bare impulse). A soft source fed [1, 0, -1] (zeros at DC and Nyquist)
    # excites neither: with rigid LRS walls, edges and corners the L2 norm of
    # the field must then stay constant -- the boundary update is lossless.
    m = small_box((16, 14, 12), 0.0)
    sim = wgo.Sim(m)
    sig = np.zeros(2000)
    sig[:3] = [1.0, 0.0, -1.0]
    src = m.index(7, 7, 6)
    norms = []
    for k in range(20):
        steps, _, flag = sim.run(src, sig[100 * k:100 * (k + 1)], [src], soft=True)
        assert steps == 100 and flag == 0
        norms.append(np.sqrt(np.square(sim.field()).sum()))
    assert 0.85 * norms[0] < min(norms) and max(norms) < 1.15 * norms[0]


def test_absorbing_box_decays():
    coeffs = as_coeffs(json.load(open(os.path.join(os.path.dirname(__file__), "golden",
                                                    "lrs_coefficients.json")))["sets"][0]["impedance"])
    m = wgo.mesh_from_inside(wgo.cuboid_inside((16, 14, 12)), [coeffs])
    sim = wgo.Sim(m)
    src = m.index(7, 7, 6)
    sig = np.zeros(2200)
    sig[:3] = [1.0, 0.0, -1.0]  # no DC / checkerboard component (see above)
    steps, _, flag = sim.run(src, sig[:200], [src], soft=True)
    assert steps == 200 and flag == 0
    e0 = np.square(sim.field()).sum()
    steps, _, flag = sim.run(src, sig[200:], [src], soft=True)
    assert steps == 2000 and flag == 0
    e1 = np.square(sim.field()).sum()
    assert np.isfinite(e1) and e1 < 0.2 * e0
    assert np.abs(sim.boundary_data(1)["mem"]).max() > 0  # filters are exercised


def test_float_mode_tracks_double_mode():
    m = small_box((20, 18, 16), 0.3)
    f, d = wgo.Sim(m, "float"), wgo.Sim(m, "double")
    for s in (f, d):
        s.write(m.index(9, 9, 8), 1.0)
        assert s.step(300) == 0
    rel = np.sqrt(np.mean((f.field() - d.field()) ** 2)) / np.abs(d.field()).max()
    assert rel < 1e-4


def test_error_flags():
    # inf / nan detection (program.cpp:522-527)
    m = small_box((10, 10, 10), 0.1)
    sim = wgo.Sim(m)
    sim.write(m.index(5, 5, 5), np.inf)
    assert sim.step(1) & wgo.ERR_INF
    sim = wgo.Sim(m)
    sim.write(m.index(5, 5, 5), np.nan)
    assert sim.step(1) & wgo.ERR_NAN
    # suspicious boundary: a 1-d boundary node whose in-plane neighbour is air
    nodes = m.nodes.copy()
    nodes["boundary_type"][m.index(5, 5, 1)] = wgo.ID_INSIDE  # was id_pz, sits among id_pz nodes
    bad = wgo.Mesh(m.dims, nodes, m.coeffs, m.b1, m.b2, m.b3)
    assert wgo.Sim(bad).step(1) & wgo.ERR_SUSPICIOUS
    # outside mesh: a boundary node on the mesh edge whose in-plane port (-x) is off-mesh
    nodes = m.nodes.copy()
    nodes["boundary_type"][m.index(0, 5, 5)] = wgo.ID_PZ
    nodes["boundary_index"][m.index(0, 5, 5)] = 0
    bad = wgo.Mesh(m.dims, nodes, m.coeffs, m.b1, m.b2, m.b3)
    assert wgo.Sim(bad).step(1) & wgo.ERR_OUTSIDE_MESH


# ---- the reference's transparent-source identity (SURVEY 8c KAT 1) ---------------------
def _mesh_impulse_response(n):
    """write_compensation_signal (compensation_signal/cmd/main.cpp:33-47, lib/.../waveguide.h:53-115):
    a free-field mesh, float pressures, hard source {0, 1, 0, 0, ...} at the centre node (a hard
    source keeps pinning the node to its next sample, zeros included), output = that node after
    every step. The reference folds the mesh into a 1/48 wedge; a cube too large for any
    reflection to return within n steps gives the same values."""
    half = n // 2 + 3
    d = 2 * half + 5
    inside = np.zeros((d, d, d), bool)
    inside[2:-2, 2:-2, 2:-2] = True
    m = wgo.mesh_from_inside(inside, [wgo.to_flat(0.1)])
    c = m.index(d // 2, d // 2, d // 2)
    sim = wgo.Sim(m, real="float")
    out = np.zeros(n)
    for k in range(n):
        sim.write(c, 1.0 if k == 1 else 0.0)   # the writer runs before the kernel ...
        assert sim.step(1) == 0
        out[k] = sim.read(c)                   # ... and the output is read after the swap
    return out


def test_transparent_soft_source_reproduces_its_input():
    """src/waveguide/tests/waveguide_init.cpp:18-63: a soft source fed make_transparent(input)
    (make_transparent.cpp:10-33: input minus its convolution with the right-Hanning-windowed mesh
    impulse response) leaves exactly `input` at the source node -- within 1e-4 for the first
    input.size() steps, in a 2 m box of absorption 0.001 meshed at 0.04 m."""
    steps, n_ir = 100, 512
    ir = np.zeros(n_ir)
    ir[:steps + 4] = _mesh_impulse_response(steps + 4)  # later samples cannot act within `steps`
    window = 0.5 - 0.5 * np.cos(2 * np.pi * (0.5 + np.arange(n_ir) / (2 * (n_ir - 1.0))))  # right_hanning
    windowed = (window * ir).astype(np.float32)
    x = np.ones(20, np.float32)
    conv = np.convolve(x.astype(np.float64), windowed.astype(np.float64))
    transparent = -conv
    transparent[:x.size] += x
    transparent = transparent[:steps].astype(np.float32).astype(np.float64)
    # geo::box(-1, 1), spacing 0.04, two layers of padding each side like compute_adjusted_boundary
    side = int(round(2.0 / 0.04)) + 1 + 4
    inside = np.zeros((side, side, side), bool)
    inside[2:-2, 2:-2, 2:-2] = True
    m = wgo.mesh_from_inside(inside, [wgo.to_flat(0.001)])
    c = m.index(side // 2, side // 2, side // 2)
    for real in ("float", "double"):
        done, out, flag = wgo.Sim(m, real=real).run(c, transparent, [c], soft=True)
        assert done == steps and flag == 0
        assert np.abs(out[:x.size, 0] - x).max() < 1e-4, real
        # and it is the compensation that does it: the plain soft source rings
        _, plain, _ = wgo.Sim(m, real=real).run(c, np.r_[x, np.zeros(steps - x.size)], [c], soft=True)
        assert np.abs(plain[:x.size, 0] - x).max() > 0.1
