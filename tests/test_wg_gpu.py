"""GPU parity tests of the waveguide path, through the C ABI, against the CPU
oracle (Real=double mode) on identical meshes and excitation.

Tolerance: BASELINE's bar is <= 1e-10 RMS (relative to peak |p|). The CUDA code
keeps the reference's operation order with FMA contraction off, so we assert
the stronger statement first -- numerically identical values -- and the 1e-10
bound as the stated contract."""
import json
import os

import numpy as np
import pytest

import wayverb_b200 as wvb
from wayverb_b200 import _lib
from oracle import wgo

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lrs_coefficients.json")
TOL_RMS = 1e-10


def golden_coeffs(i=0):
    s = json.load(open(GOLD))["sets"][i]["impedance"]
    c = np.zeros((), _lib.COEFF_DT)
    c["b"], c["a"] = s["b"], s["a"]
    return c


def to_wvb(m: wgo.Mesh) -> wvb.Mesh:
    return wvb.Mesh(m.dims, m.nodes, m.coeffs, m.b1, m.b2, m.b3)


def rel_rms(a, b):
    peak = max(np.abs(b).max(), 1e-300)
    return np.sqrt(np.mean((a - b) ** 2)) / peak


def assert_parity(got, want, what=""):
    assert np.isfinite(got).all(), what
    r = rel_rms(got, want)
    assert r <= TOL_RMS, "%s rel RMS %.3e" % (what, r)
    assert np.array_equal(got, want), "%s: within 1e-10 (%.1e) but not identical" % (what, r)


def run_both(omesh, steps, src, kernel, impulse=1.0):
    o = wgo.Sim(omesh, "double")
    g = wvb.Waveguide(to_wvb(omesh), kernel=kernel)
    o.write(src, impulse)
    g.write(src, impulse)
    fo = o.step(steps)
    fg = g.step(steps)
    return o, g, fo, fg


KERNELS = [("direct", _lib.KERNEL_DIRECT), ("tma", _lib.KERNEL_TMA)]


def test_division_by_three_is_correctly_rounded():
    """The kernels divide by 3 with an FMA-corrected reciprocal multiply
    (csrc/wg_kernels.cuh third<true>); it must equal the IEEE division bit for
    bit, including specials, the denormal range and adversarial significands."""
    rng = np.random.default_rng(11)
    n = 1 << 24
    bits = rng.integers(0, 1 << 52, n, dtype=np.uint64)
    mode = np.arange(n) & 3
    bits = np.where(mode == 1, np.uint64((1 << 52) - 1) - (bits & np.uint64(0xffff)), bits)  # ~all ones
    bits = np.where(mode == 2, bits & np.uint64(0xfffff), bits)                             # ~power of two
    exp = rng.integers(0, 2047, n, dtype=np.uint64)       # every exponent incl. denormals / inf / nan
    exp = np.where(np.arange(n) % 5 == 0, exp, np.uint64(1023) + (exp % np.uint64(120)) - np.uint64(60))
    sign = rng.integers(0, 2, n, dtype=np.uint64) << np.uint64(63)
    x = (bits | (exp << np.uint64(52)) | sign).view(np.float64)
    x[:8] = [0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324, 1.7976931348623157e308, 3.0]
    fast, ref = np.zeros(n), np.zeros(n)
    _lib.check(_lib.lib().wvb_test_third(_lib.ptr(x), n, _lib.ptr(fast), _lib.ptr(ref)))
    with np.errstate(all="ignore"):
        host = x / 3.0
    same = (fast.view(np.uint64) == ref.view(np.uint64)) | (np.isnan(fast) & np.isnan(ref)) | \
           ((fast == 0) & (ref == 0))
    assert same.all(), "%d mismatches vs device division" % (~same).sum()
    ok_host = (ref == host) | (np.isnan(ref) & np.isnan(host))
    assert ok_host.all()


@pytest.mark.parametrize("kind", ["impulse", "noise", "quiet"])
def test_device_filter_kernels(kind):
    """tests/rectangular_kernel.cpp:242-360 on the device: `filter_test` (biquad cascade)
    and `filter_test_2` (convolved canonical filter) on 256 parallel random peak-filter
    sets: no NaN/Inf (also for +-1e-35 inputs), biquad == canonical within 1e-3, and both
    identical to the oracle's restatement."""
    rng = np.random.default_rng(17)
    streams = 256
    n = {"impulse": 200, "noise": 2000, "quiet": 4000}[kind]
    biq = np.stack([[wgo.peak_biquad(rng.uniform(0.1, 1), rng.uniform(0, 0.5), rng.uniform(0, 1))
                     for _ in range(3)] for _ in range(streams)])
    canon = np.stack([wgo.convolve3(b) for b in biq]).view(_lib.COEFF_DT).reshape(streams)
    if kind == "impulse":
        x = np.zeros((n, streams), np.float32)
        x[0] = 0.25
    elif kind == "noise":
        x = rng.uniform(-0.25, 0.25, (n, streams)).astype(np.float32)
    else:
        x = rng.uniform(-1e-35, 1e-35, (n, streams)).astype(np.float32)
    y1, y2 = np.zeros_like(x), np.zeros_like(x)
    bq = np.ascontiguousarray(biq, np.float64)
    _lib.check(_lib.lib().wvb_test_filter(_lib.ptr(bq), None, _lib.ptr(x), streams, n, _lib.ptr(y1)))
    _lib.check(_lib.lib().wvb_test_filter(None, _lib.ptr(canon), _lib.ptr(x), streams, n, _lib.ptr(y2)))
    assert np.isfinite(y1).all() and np.isfinite(y2).all()
    if kind != "quiet":
        assert np.abs(y1 - y2).max() < 1e-3
    for s in range(0, streams, 37):
        assert np.array_equal(y1[:, s], wgo.filter_biquads(biq[s], x[:, s].copy()))
        assert np.array_equal(y2[:, s], wgo.filter_canonical(canon[s], x[:, s].copy()))


@pytest.mark.parametrize("kname,kernel", KERNELS)
@pytest.mark.parametrize("dims", [(140, 24, 14), (150, 37, 19), (133, 11, 9), (260, 20, 12)])
def test_box_plaster_field_and_filters(dims, kname, kernel):
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [golden_coeffs(0)])
    src = om.index(dims[0] // 2, dims[1] // 2, dims[2] // 2)
    o, g, fo, fg = run_both(om, 60, src, kernel)
    assert fo == 0 and fg == 0
    assert g.info()["kernel_variant"] == kname
    assert_parity(g.field(), o.field(), "field")
    for n in (1, 2, 3):
        bo, bg = o.boundary_data(n), g.boundary_data(n)
        assert bo.shape == bg.shape
        assert np.array_equal(bo["coefficient_index"], bg["coefficient_index"])
        assert_parity(bg["filter_memory"].ravel(), bo["mem"].ravel(), "filter memory %d-d" % n)
    assert np.abs(o.boundary_data(1)["mem"]).max() > 0


@pytest.mark.parametrize("dims", [(37, 29, 23), (16, 14, 12), (5, 5, 5), (64, 8, 8), (33, 27, 22)])
def test_small_and_odd_boxes_direct(dims):
    # config 1 sized meshes (33x27x22 is the 500 Hz shoebox of BASELINE.md)
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [wgo.to_flat(0.1)])
    src = om.index(dims[0] // 2, dims[1] // 2, dims[2] // 2)
    o, g, fo, fg = run_both(om, 100, src, _lib.KERNEL_AUTO)
    assert fo == 0 and fg == 0
    assert_parity(g.field(), o.field())


@pytest.mark.parametrize("kname,kernel", KERNELS)
def test_rigid_box(kname, kernel):
    dims = (136, 16, 12)
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [wgo.to_flat(0.0)])
    o, g, fo, fg = run_both(om, 80, om.index(60, 8, 6), kernel)
    assert fo == 0 and fg == 0
    assert_parity(g.field(), o.field())
    assert not g.boundary_data(1)["filter_memory"].any()  # a0 == 0 guards keep memory at 0


@pytest.mark.parametrize("kname,kernel", KERNELS)
def test_l_shaped_room_three_surfaces(kname, kernel):
    # reentrant nodes, several surfaces, 2-d/3-d nodes with mixed coefficient sets
    dz, dy, dx = 14, 30, 150
    ins = np.zeros((dz, dy, dx), bool)
    ins[2:dz - 2, 2:dy - 2, 2:70] = True
    ins[2:dz - 2, 2:14, 2:dx - 2] = True
    zz, yy, xx = np.indices(ins.shape)
    surf = ((xx > 60).astype(np.uint32) + (yy > 12).astype(np.uint32)).ravel()
    om = wgo.mesh_from_inside(ins, [golden_coeffs(0), golden_coeffs(1), golden_coeffs(2)], surf)
    assert (om.nodes["boundary_type"] == wgo.ID_REENTRANT).any()
    assert len(set(om.b1.ravel().tolist())) == 3
    o, g, fo, fg = run_both(om, 120, om.index(30, 8, 7), kernel)
    assert fo == 0 and fg == 0
    assert_parity(g.field(), o.field())
    for n in (1, 2, 3):
        assert_parity(g.boundary_data(n)["filter_memory"].ravel(), o.boundary_data(n)["mem"].ravel())


def test_hard_source_run_matches_oracle_and_callbacks():
    dims = (40, 30, 20)
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [golden_coeffs(1)])
    src, rcv = om.index(12, 11, 9), [om.index(25, 17, 10), om.index(12, 11, 9)]
    sig = np.zeros(150)
    sig[0] = 1.0
    steps_o, out_o, flag_o = wgo.Sim(om).run(src, sig, rcv)
    with wvb.Waveguide(to_wvb(om)) as g:
        steps_g, out_g, flag_g = g.run_device(src, sig, rcv)
    assert (steps_o, flag_o) == (150, 0) and (steps_g, flag_g) == (150, 0)
    assert_parity(out_g, out_o, "receiver traces")
    # the callback protocol of waveguide::run gives the same trace
    trace = []
    steps = wvb.run(to_wvb(om), wvb.hard_source(src, sig[:40]), wvb.node_receiver(rcv[0], trace))
    assert steps == 40
    assert np.array_equal(np.array(trace), out_o[:40, 0])


def test_soft_source_run():
    dims = (30, 30, 30)
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [wgo.to_flat(0.3)])
    src, rcv = om.index(15, 15, 15), [om.index(15, 15, 15), om.index(20, 12, 9)]
    sig = np.zeros(80)
    sig[:3] = [1.0, 0.0, -1.0]
    _, out_o, _ = wgo.Sim(om).run(src, sig, rcv, soft=True)
    with wvb.Waveguide(to_wvb(om)) as g:
        steps, out_g, flag = g.run_device(src, sig, rcv, soft=True, check_interval=16)
    assert steps == 80 and flag == 0
    assert_parity(out_g, out_o)


def test_field_io_and_f32_view():
    dims = (20, 18, 16)
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [wgo.to_flat(0.2)])
    rng = np.random.default_rng(3)
    f0 = rng.standard_normal(om.num_nodes)
    o = wgo.Sim(om)
    o.set_field(f0)
    with wvb.Waveguide(to_wvb(om)) as g:
        g.set_field(f0)
        assert np.array_equal(g.field(), f0)
        assert np.array_equal(g.field_f32(), f0.astype(np.float32))
        assert g.read(om.index(3, 4, 5)) == f0[om.index(3, 4, 5)]
        assert o.step(7) == 0 and g.step(7) == 0
        assert_parity(g.field(), o.field())


def test_error_flags_match_reference_semantics():
    dims = (12, 12, 12)
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [wgo.to_flat(0.1)])
    for bad, bit in ((np.inf, 1), (np.nan, 2)):
        with wvb.Waveguide(to_wvb(om)) as g:
            g.write(om.index(6, 6, 6), bad)
            assert g.step(1) & bit
    # suspicious boundary / outside mesh, same constructions as the oracle KAT
    nodes = om.nodes.copy()
    nodes["boundary_type"][om.index(5, 5, 1)] = wgo.ID_INSIDE
    bad = wgo.Mesh(om.dims, nodes, om.coeffs, om.b1, om.b2, om.b3)
    with wvb.Waveguide(to_wvb(bad)) as g:
        assert g.step(1) == wgo.Sim(bad).step(1) == 16
    nodes = om.nodes.copy()
    nodes["boundary_type"][om.index(0, 5, 5)] = wgo.ID_PZ
    bad = wgo.Mesh(om.dims, nodes, om.coeffs, om.b1, om.b2, om.b3)
    with wvb.Waveguide(to_wvb(bad)) as g:
        fg = g.step(1)
    assert fg == wgo.Sim(bad).step(1) and fg & 8
    # run() rethrows like waveguide.h:100-119
    with pytest.raises(wvb.waveguide.ValueIsNan):
        wvb.run(to_wvb(om), wvb.hard_source(om.index(6, 6, 6), [np.nan, 0.0]), lambda wg, s: None)


def test_determinism_bit_identical():
    # verify_compensation_signal.cpp:23-91: same input => identical output
    dims = (140, 20, 16)
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [golden_coeffs(2)])
    fields = []
    for _ in range(3):
        with wvb.Waveguide(to_wvb(om), kernel=_lib.KERNEL_TMA) as g:
            g.write(om.index(70, 10, 8), 1.0)
            assert g.step(100) == 0
            fields.append(g.field())
    assert np.array_equal(fields[0], fields[1]) and np.array_equal(fields[0], fields[2])


def test_slab_handle_matches_whole_mesh_away_from_its_edges():
    # a handle that owns planes [6, 14) with zero ghosts: after s steps, planes
    # further than s from the slab faces cannot have seen the missing halo
    dims = (24, 20, 20)
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [wgo.to_flat(0.2)])
    o = wgo.Sim(om)
    src = om.index(12, 10, 10)
    o.write(src, 1.0)
    with wvb.Waveguide(to_wvb(om), z_range=(6, 14)) as g:
        assert g.owns(src) and not g.owns(om.index(1, 1, 2))
        g.write(src, 1.0)
        assert o.step(3) == 0 and g.step(3) == 0
        want = o.field().reshape(dims[2], dims[1], dims[0])[6:14]
        got = g.field().reshape(8, dims[1], dims[0])
        assert np.array_equal(got[3:5], want[3:5])


def test_config2_256_cube_rigid_against_oracle():
    """BASELINE config 2 (256^3, rigid walls, hard source 1.0 at (128,128,128), receiver
    (160,140,120)): the first 120 of its 10 000 steps against the oracle, full field and
    receiver trace."""
    dims = (256, 256, 256)
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [wgo.to_flat(0.0)])
    src, rcv = om.index(128, 128, 128), [om.index(160, 140, 120)]
    sig = np.zeros(120)
    sig[0] = 1.0
    o = wgo.Sim(om)
    steps_o, out_o, flag_o = o.run(src, sig, rcv)
    with wvb.Waveguide(to_wvb(om)) as g:
        steps_g, out_g, flag_g = g.run_device(src, sig, rcv, check_interval=40)
        assert (steps_o, flag_o) == (steps_g, flag_g) == (120, 0)
        assert_parity(out_g, out_o, "receiver trace")
        assert_parity(g.field(), o.field(), "field")
        assert np.abs(out_o).max() > 0


@pytest.mark.parametrize("dims,steps", [((256, 256, 64), 12)])
def test_config2_sized_slice_against_oracle(dims, steps):
    # 4.2 M nodes, rigid (config 2's boundary) and plaster: both kernels vs oracle
    for coeffs in (wgo.to_flat(0.0), golden_coeffs(0)):
        om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [coeffs])
        src = om.index(128, 128, 32)
        o = wgo.Sim(om)
        o.write(src, 1.0)
        assert o.step(steps) == 0
        want = o.field()
        for _, kernel in KERNELS:
            with wvb.Waveguide(to_wvb(om), kernel=kernel) as g:
                g.write(src, 1.0)
                assert g.step(steps) == 0
                assert_parity(g.field(), want)


def test_config3_full_size_against_oracle():
    """BASELINE config 3 at full size (512^3, plaster LRS) against the oracle: a seeded random
    pressure field over the whole mesh (so that all 134 M air nodes and all 1.55 M boundary
    nodes of every class carry data from the first step on), 6 steps, then the field and the
    three filter-memory arrays of BOTH kernels must equal the oracle's bit for bit. The oracle
    steps 512^3 in about a third of a second per step on 16 cores."""
    dims = (512, 512, 512)
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [golden_coeffs(0)])
    rng = np.random.default_rng(2026)
    f0 = rng.standard_normal(om.num_nodes)
    f0[om.nodes["boundary_type"] == 0] = 0.0          # id_none nodes hold no pressure
    steps = 6
    o = wgo.Sim(om)
    o.set_field(f0)
    assert o.step(steps) == 0
    want = o.field()
    want_mem = [o.boundary_data(n)["mem"].copy() for n in (1, 2, 3)]
    del o
    assert [m.shape[0] for m in want_mem] == [6 * 508 * 508, 12 * 508, 8]
    assert all(np.abs(m).max() > 0 for m in want_mem)
    m = to_wvb(om)
    for name, kernel in KERNELS:
        with wvb.Waveguide(m, kernel=kernel) as g:
            g.set_field(f0)
            assert g.step(steps) == 0
            assert g.info()["kernel_variant"] == name
            assert_parity(g.field(), want, "field (%s)" % name)
            for n in (1, 2, 3):
                bg = g.boundary_data(n)
                assert_parity(bg["filter_memory"].ravel(), want_mem[n - 1].ravel(),
                              "filter memory %d-d (%s)" % (n, name))


def test_full_size_512_cube_properties():
    """Config 3 over more steps than the oracle comparison above affords: the two independent
    kernels agree bit for bit after 24 steps of a soft source, and the field keeps the mirror
    symmetry about the source while the wave has not reached a wall."""
    dims = (512, 512, 512)
    m = wvb.cuboid_mesh(dims, [golden_coeffs(0)])
    src = m.index(255, 255, 255)
    fields = {}
    for name, kernel in KERNELS:
        with wvb.Waveguide(m, kernel=kernel) as g:
            sig = np.zeros(24)
            sig[:3] = [1.0, 0.0, -1.0]
            steps, _, flag = g.run_device(src, sig, [src], soft=True)
            assert steps == 24 and flag == 0
            fields[name] = g.field()
    assert np.array_equal(fields["direct"], fields["tma"])
    f = fields["tma"].reshape(512, 512, 512)
    assert np.isfinite(f).all() and np.abs(f).max() > 0
    k = 20
    a = f[255 + k, 200:311, 200:311]
    b = f[255 - k, 200:311, 200:311]
    assert np.abs(a - b).max() <= 1e-12 * np.abs(f).max()
    a = f[200:311, 200:311, 255 + k]
    b = f[200:311, 200:311, 255 - k]
    assert np.abs(a - b).max() <= 1e-12 * np.abs(f).max()


def test_config2_ten_thousand_steps():
    """BASELINE config 2: 256^3, rigid boundaries, fp64, 10 000 steps on the device in one
    wvb_wg_run call (source (128,128,128), receiver (160,140,120)).
    The oracle needs ~6 minutes for 10 000 steps of 256^3 on 16 cores, so:
      (a) full size: the first 300 receiver samples against the oracle at 256^3;
      (b) full length: all 10 000 steps against the oracle on a 64^3 twin (same walls, same
          source/receiver offsets from the corner), receiver trace and final field;
      (c) full size and full length: the run completes without an error flag, stays finite
          and bounded (a rigid box neither gains nor loses energy: the trace's RMS over the
          last 1000 steps is within a factor of 4 of the RMS over steps 1000-2000)."""
    rigid = wgo.to_flat(0.0)
    n = 10000
    sig = np.zeros(n)
    sig[0] = 1.0
    # (b) 64^3 twin, all 10 000 steps
    dims_s = (64, 64, 64)
    om_s = wgo.mesh_from_inside(wgo.cuboid_inside(dims_s), [rigid])
    src_s, rcv_s = om_s.index(32, 32, 32), [om_s.index(40, 35, 30)]
    o = wgo.Sim(om_s)
    steps_o, out_o, flag_o = o.run(src_s, sig, rcv_s)
    with wvb.Waveguide(to_wvb(om_s)) as g:
        steps_g, out_g, flag_g = g.run_device(src_s, sig, rcv_s, check_interval=500)
        assert (steps_o, flag_o) == (steps_g, flag_g) == (n, 0)
        assert_parity(out_g, out_o, "64^3 receiver trace, 10 000 steps")
        assert_parity(g.field(), o.field(), "64^3 field after 10 000 steps")
    # (a) + (c) full size
    dims = (256, 256, 256)
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [rigid])
    src, rcv = om.index(128, 128, 128), [om.index(160, 140, 120)]
    head = 300
    o = wgo.Sim(om)
    _, head_o, flag_o = o.run(src, sig[:head], rcv)
    assert flag_o == 0
    with wvb.Waveguide(to_wvb(om)) as g:
        steps_g, out_g, flag_g = g.run_device(src, sig, rcv, check_interval=1000)
        assert (steps_g, flag_g) == (n, 0)
        assert_parity(out_g[:head], head_o, "256^3 receiver trace, first %d steps" % head)
        assert np.isfinite(out_g).all() and np.abs(out_g[head:]).max() > 0
        early = np.sqrt(np.mean(out_g[1000:2000] ** 2))
        late = np.sqrt(np.mean(out_g[-1000:] ** 2))
        assert 0.25 < late / early < 4.0, (early, late)
        f = g.field()
        assert np.isfinite(f).all() and np.abs(f).max() < 1.0


# ---- temporal blocking: two steps per pass (WVB_WG_TEMPORAL2) --------------------------------------
@pytest.mark.parametrize("dims,steps", [((140, 24, 14), 60), ((150, 37, 19), 41), ((260, 20, 12), 7), ((133, 11, 9), 2)])
def test_temporal_blocking_is_bit_identical(dims, steps):
    """pairs of steps fused into one pass (wg_air_tb2 + shell + two boundary launches) against the
    oracle's single steps: field and filter memories identical, odd step counts included"""
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [golden_coeffs(0)])
    src = om.index(dims[0] // 2, dims[1] // 2, dims[2] // 2)
    o = wgo.Sim(om)
    o.write(src, 1.0)
    o.write(om.index(3, 3, 3), -0.5)          # next to a corner: walls, edges and the corner react early
    assert o.step(steps) == 0
    with wvb.Waveguide(to_wvb(om), kernel=_lib.KERNEL_TMA, flags=_lib.TEMPORAL2) as g:
        g.write(src, 1.0)
        g.write(om.index(3, 3, 3), -0.5)
        assert g.step(steps) == 0
        assert_parity(g.field(), o.field(), "field")
        for n in (1, 2, 3):
            assert_parity(g.boundary_data(n)["filter_memory"].ravel(), o.boundary_data(n)["mem"].ravel(),
                          "filter memory %d-d" % n)
        # and on: a few more calls of mixed length keep the arrays' roles straight
        for k in (1, 2, 3, 4):
            assert g.step(k) == 0 and o.step(k) == 0
        assert_parity(g.field(), o.field(), "field after mixed-length calls")
        assert g.read(src) == o.read(src)


def test_temporal_blocking_l_shaped_room_and_random_field():
    dz, dy, dx = 14, 30, 150
    ins = np.zeros((dz, dy, dx), bool)
    ins[2:dz - 2, 2:dy - 2, 2:70] = True
    ins[2:dz - 2, 2:14, 2:dx - 2] = True
    zz, yy, xx = np.indices(ins.shape)
    surf = ((xx > 60).astype(np.uint32) + (yy > 12).astype(np.uint32)).ravel()
    om = wgo.mesh_from_inside(ins, [golden_coeffs(0), golden_coeffs(1), golden_coeffs(2)], surf)
    f0 = np.random.default_rng(7).standard_normal(om.num_nodes)
    f0[om.nodes["boundary_type"] == 0] = 0.0
    o = wgo.Sim(om)
    o.set_field(f0)
    assert o.step(10) == 0
    with wvb.Waveguide(to_wvb(om), kernel=_lib.KERNEL_TMA, flags=_lib.TEMPORAL2) as g:
        g.set_field(f0)
        assert g.step(10) == 0
        assert_parity(g.field(), o.field())
        for n in (1, 2, 3):
            assert_parity(g.boundary_data(n)["filter_memory"].ravel(), o.boundary_data(n)["mem"].ravel())
    # error flags survive the fusion: an inf planted next to a wall is reported like the reference does
    # (wvb_wg_step ORs the flags of the steps of one call; the oracle reports step by step)
    o2 = wgo.Sim(om)
    o2.write(om.index(30, 8, 7), np.inf)
    fo = o2.step(1) | o2.step(1)
    with wvb.Waveguide(to_wvb(om), kernel=_lib.KERNEL_TMA, flags=_lib.TEMPORAL2) as g:
        g.write(om.index(30, 8, 7), np.inf)
        assert g.step(2) == fo and fo != 0
