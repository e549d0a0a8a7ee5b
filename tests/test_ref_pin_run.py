"""The reference's OWN `waveguide::run` template, run here: waveguide/waveguide.h:36-126 with the stock
processors -- preprocessor::hard_source / soft_source / gaussian, postprocessor::node /
directional_receiver under core::callback_accumulator -- program.cpp's class, setup.cpp and mesh's
constructor, compiled unmodified from /root/reference over a host-memory stand-in for cl.hpp
(oracle/ref_recipe/hostcl/CL/cl.hpp: buffers are host memory, the queue is synchronous, make_kernel
looks the kernel up by name) and enqueueing the reference's own condensed_waveguide kernel as compiled
for the host (oracle/_ref, ref_wg_f32.cpp). No line of the host loop is restated on that side.

Held against it, bit for bit: the oracle's loop (oracle/wg_oracle.cpp wgo_run, float mode = the
reference's own types) -- which is what the GPU tests compare `waveguide::run` of the shim with -- and
the numpy restatement of gaussian + directional_receiver that tests/test_cpp_shim.py uses."""
import json
import os

import numpy as np
import pytest

from oracle import refk, wgo

pytestmark = pytest.mark.skipif(not refk.available(), reason="no /root/reference and no prebuilt oracle/_ref")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def golden_coefficients(k):
    s = json.load(open(os.path.join(ROOT, "tests", "golden", "lrs_coefficients.json")))["sets"][k]
    c = np.zeros((), wgo.COEFF_DT)
    c["b"], c["a"] = s["impedance"]["b"], s["impedance"]["a"]
    return c


def l_shape(dims):
    inside = wgo.cuboid_inside(dims)
    inside[dims[2] // 2:, dims[1] // 2:, :] = False          # cut a quadrant away: re-entrant edges
    return inside


@pytest.mark.parametrize("soft", [False, True])
def test_cuboid_hard_and_soft_source(soft):
    dims = (30, 24, 40)
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [wgo.to_flat(0.1)])
    steps = 120
    rng = np.random.default_rng(1)
    sig = np.zeros(steps)
    # a signal, not just an impulse -- in float, which is what the reference's sources hold (canonical.h:48)
    sig[:20] = rng.standard_normal(20).astype(np.float32)
    src = om.index(15, 12, 8)
    rcv = [om.index(15, 12, z) for z in (12, 18, 24, 30)] + [src]
    done, want, flag = wgo.Sim(om, "float").run(src, sig, rcv, soft=soft)
    got_done, got, _ = refk.run_waveguide(om, src, sig, rcv, soft=soft)
    assert done == got_done == steps and flag == 0
    assert np.abs(got).max() > 0.01
    assert np.array_equal(got, want.astype(np.float32))


@pytest.mark.parametrize("soft", [False, True])
def test_l_shaped_room_with_fitted_walls(soft):
    """2-d and 3-d boundary nodes, re-entrant edges, three of the reference's own LRS coefficient sets"""
    dims = (22, 20, 18)
    om = wgo.mesh_from_inside(l_shape(dims), [golden_coefficients(k) for k in (0, 1, 2)])
    assert om.b2.shape[0] > 0 and om.b3.shape[0] > 0
    steps = 150
    sig = np.zeros(steps)
    sig[0] = 1.0
    src = om.index(6, 5, 5)
    rcv = [om.index(5, 5, 5), om.index(15, 6, 4), om.index(4, 14, 6), om.index(8, 4, 14)]
    done, want, flag = wgo.Sim(om, "float").run(src, sig, rcv, soft=soft)
    got_done, got, _ = refk.run_waveguide(om, src, sig, rcv, soft=soft)
    assert done == got_done == steps and flag == 0
    assert np.array_equal(got, want.astype(np.float32))


def test_unstable_filter_throws_where_the_oracle_flags():
    """waveguide.h:100-119: the error flag becomes an exception"""
    dims = (10, 9, 8)
    bad = np.zeros((), wgo.COEFF_DT)
    bad["b"][0], bad["a"][0], bad["a"][1] = 1.0, 1.0, -3.0    # a pole far outside the unit circle
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [bad])
    sig = np.zeros(4000)
    sig[0] = 1.0
    src, rcv = om.index(5, 4, 4), [om.index(4, 4, 4)]
    done, _, flag = wgo.Sim(om, "float").run(src, sig, rcv)
    assert flag != 0 and done < sig.size
    with pytest.raises(RuntimeError) as e:
        refk.run_waveguide(om, src, sig, rcv)
    text = str(e.value)
    # the reference tests inf first, then nan (waveguide.h:102-110)
    assert flag & (wgo.ERR_INF | wgo.ERR_NAN)
    assert ("is inf" in text) if flag & wgo.ERR_INF else ("is nan" in text)


def test_gaussian_preprocessor_and_directional_receiver():
    """the same scenario, and the same numpy restatement, as tests/test_cpp_shim.py's GPU test -- here
    with the reference's own gaussian.cpp and directional_receiver.cpp on the other side"""
    steps = 40
    c = wgo.to_flat(0.2)
    dims, spacing = (28, 26, 24), np.float32(0.05)
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [c])
    centre, sdev = (0.6, 0.65, 0.55), np.float32(0.1)
    rcv = om.index(17, 12, 10)
    rate, density = 11776.0, 1.1765
    done, pressures, directional = refk.run_waveguide(om, receivers=[rcv], gaussian=(centre, float(sdev), steps),
                                                      directional=(rcv, rate, density), spacing=float(spacing))
    assert done == steps
    # gaussian.cpp:13-16, 36-47
    z, y, x = np.indices((dims[2], dims[1], dims[0]))
    pos = [np.float32(0) + v.astype(np.float32) * spacing for v in (x, y, z)]
    d = [pos[k] - np.float32(centre[k]) for k in range(3)]
    ln = np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]).astype(np.float32)
    g = (np.exp(-np.power(ln.astype(np.float64), 2) / (2 * np.power(np.float64(sdev), 2))) /
         np.power(np.float64(sdev) * np.sqrt(2 * np.pi), 3)).astype(np.float32)
    sim = wgo.Sim(om, "float")
    sim.set_field(g.astype(np.float64).ravel())
    nb = [om.index(16, 12, 10), om.index(18, 12, 10), om.index(17, 11, 10), om.index(17, 13, 10),
          om.index(17, 12, 9), om.index(17, 12, 11)]
    vel = np.zeros(3)
    want = []
    for _ in range(steps):                                    # directional_receiver.cpp:29-69
        p = np.float32(sim.read(rcv))
        s = [np.float32((np.float32(sim.read(n)) - p) / np.float64(spacing)) for n in nb]
        m = np.array([(s[1] - s[0]) * 0.5, (s[3] - s[2]) * 0.5, (s[5] - s[4]) * 0.5], np.float64)
        vel -= m / (density * rate)
        want.append([np.float32(v * np.float64(p)) for v in vel] + [p])
        assert sim.step(1) == 0
    want = np.array(want, np.float32)
    assert np.abs(want[:, :3]).max() > 0
    assert np.array_equal(pressures[:, 0], want[:, 3])
    # exp() of the Gaussian goes through libm on one side and numpy on the other: the field may differ in
    # the last bit of a few nodes, so the derived quantities carry a float tolerance
    np.testing.assert_allclose(directional, want, rtol=2e-6, atol=1e-12)


def test_directional_receiver_next_to_a_wall_is_refused():
    om = wgo.mesh_from_inside(wgo.cuboid_inside((8, 8, 8)), [wgo.to_flat(0.1)])
    with pytest.raises(RuntimeError, match="adjacent to a boundary"):
        refk.run_waveguide(om, 0, np.zeros(2), [om.index(4, 4, 4)], directional=(om.index(0, 4, 4), 1000.0, 1.2))


def test_canonical_single_band_is_the_references():
    """canonical.h:24-110 on the scenario of tests/test_cpp_shim.py::test_canonical_single_band_matches_oracle:
    the node the source / receiver positions round to, the calibrated impulse, ceil(sample_rate * time)
    steps, the sample rate, valid_hz -- from the reference's own header -- and the receiver pressure
    against the oracle's loop."""
    dims, spacing = (28, 26, 24), np.float32(0.05)
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [wgo.to_flat(0.2)])
    source, receiver = (0.52, 0.61, 0.48), (0.86, 0.59, 0.51)
    rate = refk.hm_rates(spacing, 340.0)[0]
    surfaces = np.zeros((1, 16), np.float32)
    bands, calls = refk.canonical(om, surfaces, source, receiver, 120.2 / rate, float(spacing))
    assert len(bands) == 1 and calls == 121
    band = bands[0]
    assert band["directional"].shape == (121, 4) and band["sample_rate"] == rate and band["valid_hz"] == (0.0, 500.0)
    # what the GPU test expects of the shim, now read off the reference itself
    loc = lambda p: [int(np.round(np.float32(v) / spacing)) for v in p]  # noqa: E731
    src, rcv = om.index(*loc(source)), om.index(*loc(receiver))
    amp = np.float32(np.sqrt(400.0 / (4 * np.pi)) / (0.3405 * np.float64(spacing)))
    assert amp == np.float32(refk.lib().refk_hm_calibration_factor(float(spacing), 400.0))
    sig = np.zeros(121)
    sig[0] = float(amp)
    done, want, flag = wgo.Sim(om, "float").run(src, sig, [rcv])
    assert done == 121 and flag == 0
    assert np.array_equal(band["directional"][:, 3], want[:, 0].astype(np.float32))
    assert np.abs(band["directional"][:, :3]).max() > 0
    # a position outside the mesh (canonical.h:42-50)
    with pytest.raises(RuntimeError, match="outside"):
        refk.canonical(om, surfaces, (-5.0, 0.0, 0.0), receiver, 120.2 / rate, float(spacing))


def test_canonical_multiple_bands_is_the_references():
    """canonical.h:140-177: one run per band with that band's flat coefficients, band edges of
    hrtf_band_params_hz -- the same expectations tests/cpp/test_waveguide_shim.cpp holds the shim to"""
    dims, spacing = (20, 18, 16), np.float32(0.05)
    absorption = [0.05, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7]
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [wgo.to_flat(0.9)])      # overwritten per band
    surfaces = np.zeros((1, 16), np.float32)
    surfaces[0, :8] = absorption
    source, receiver = (0.42, 0.41, 0.38), (0.61, 0.49, 0.41)
    rate = refk.hm_rates(spacing, 340.0)[0]
    bands, calls = refk.canonical(om, surfaces, source, receiver, 60.5 / rate, float(spacing), bands=3)
    assert len(bands) == 3 and calls == 3 * 61
    loc = lambda p: [int(np.round(np.float32(v) / spacing)) for v in p]  # noqa: E731
    src, rcv = om.index(*loc(source)), om.index(*loc(receiver))
    amp = np.float32(np.sqrt(400.0 / (4 * np.pi)) / (0.3405 * np.float64(spacing)))
    sig = np.zeros(61)
    sig[0] = float(amp)
    for b, band in enumerate(bands):
        assert band["valid_hz"][0] == pytest.approx(20.0 * 1000.0 ** (b / 8.0), rel=1e-14)
        assert band["valid_hz"][1] == pytest.approx(20.0 * 1000.0 ** ((b + 1) / 8.0), rel=1e-14)
        flat = wgo.to_flat(float(np.float32(absorption[b])))                    # surface.absorption.s[band] is a float
        mb = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [flat])
        done, want, flag = wgo.Sim(mb, "float").run(src, sig, [rcv])
        assert done == 61 and flag == 0
        assert np.array_equal(band["directional"][:, 3], want[:, 0].astype(np.float32))
    assert not np.array_equal(bands[0]["directional"], bands[2]["directional"])


@pytest.mark.parametrize("cutoff", [500.0, 1000.0])
def test_baseline_config_1_end_to_end_against_the_reference_itself(cutoff):
    """BASELINE.json configs[0]: 5 x 4 x 3 m shoebox, absorption 0.1, source
    (1, 1, 1), receiver (2, 3, 1.5), single_band_parameters{cutoff, 0.6} -- the reference's own CPU-
    runnable case, run here by the reference's own code from the scene onwards: adjusted boundary around
    the receiver, depth-5 voxelisation, compute_mesh with its fitted wall filters, canonical(). The
    oracle's pipeline on the same descriptor (voxel inside test, classification, closest surface, the
    float-mode stepping loop) must return the same receiver pressures, bit for bit."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lrs as olrs
    from oracle import rto
    from wayverb_b200 import scene

    def fit(order, f, m):
        return olrs.yulewalk(order, list(f), list(m))

    b = scene.box_scene((5.0, 4.0, 3.0), subdiv=1, surfaces=[scene.make_surface(0.1, 0.1)])
    sc = scene.Scene(b.vertices, b.triangles, b.surfaces, voxeliser="octree", depth=5)
    src, rcv = (1.0, 1.0, 1.0), (2.0, 3.0, 1.5)
    fs = cutoff / (0.25 * 0.6)                                # compute_sampling_frequency, simulation_parameters.h:65-68
    ref = refk.compute_mesh(sc, fit, anchor=rcv, sample_rate=fs)
    assert ref.dims == {500.0: (33, 26, 22), 1000.0: (60, 50, 37)}[cutoff]
    bands, calls = refk.canonical(ref, np.zeros((1, 16), np.float32), src, rcv, 0.1, spacing=ref.spacing,
                                  min_corner=ref.min_corner, cutoff=cutoff, capacity_steps=8192)
    steps = int(np.ceil(bands[0]["sample_rate"] * 0.1))
    assert len(bands) == 1 and calls == steps == bands[0]["directional"].shape[0]
    assert bands[0]["valid_hz"] == (0.0, cutoff) and abs(bands[0]["sample_rate"] - fs) < 1e-3

    o = rto.Scene(sc)
    sp = np.float32(ref.spacing)
    ins = o.nodes_inside(ref.min_corner, ref.dims, sp)
    z, y, x = np.indices(ins.shape)
    pts = np.stack([ref.min_corner[0] + x.astype(np.float32) * sp, ref.min_corner[1] + y.astype(np.float32) * sp,
                    ref.min_corner[2] + z.astype(np.float32) * sp], -1).reshape(-1, 3)
    surf, _ = o.closest_surface(pts)
    om = wgo.mesh_from_inside(ins, [ref.coeffs[0]], surf)
    assert np.array_equal(om.nodes["boundary_type"], ref.nodes["boundary_type"])
    loc = lambda p: [int(np.round((np.float32(v) - m) / sp)) for v, m in zip(p, ref.min_corner)]  # noqa: E731
    sig = np.zeros(steps)
    sig[0] = float(np.float32(refk.lib().refk_hm_calibration_factor(float(sp), 400.0)))
    done, want, flag = wgo.Sim(om, "float").run(om.index(*loc(src)), sig, [om.index(*loc(rcv))])
    assert done == steps and flag == 0 and np.abs(want).max() > 0.1
    assert np.array_equal(bands[0]["directional"][:, 3], want[:, 0].astype(np.float32))
