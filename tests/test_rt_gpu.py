"""GPU parity tests of the ray path, through the C ABI, against the CPU oracle.

All ray arithmetic is fp32 in the reference's operation order with FMA
contraction off on both sides, and the random numbers / sincos are the shared
fixed definitions, so hits, visibility flags and reflection points must be
IDENTICAL (the reference's own GPU-vs-CPU bound is 1e-5 on positions,
reflector_tests.cpp:122,145). Histograms are accumulated with fp64 atomics whose
order is not fixed: tolerance 1e-9 relative to the largest bin (BASELINE's
statistical criterion is 10 % per band / 1 % total)."""
import numpy as np
import pytest

import wayverb_b200 as wvb
from wayverb_b200 import scene
from oracle import rto

pytestmark = pytest.mark.gpu

SRC = np.array([2.09, 2.12, 2.12], np.float32)
RCV = np.array([2.09, 1.08, 0.96], np.float32)
BOX = (4.0, 3.0, 6.0)


def room(scatter, subdiv=2, side=8, absorption=0.1, per_wall=False):
    surfs = [scene.make_surface(absorption, scatter)]
    if per_wall:
        surfs = [scene.make_surface(0.05 + 0.05 * k, [scatter * (1 + 0.1 * b) for b in range(8)]) for k in range(3)]
    return scene.box_scene(BOX, subdiv=subdiv, side=side, surfaces=surfs, per_wall_surfaces=per_wall)


def assert_hist_close(got, want):
    scale = np.abs(want).max()
    assert scale > 0
    assert np.abs(got - want).max() <= 1e-9 * scale


def test_generated_directions_identical():
    with wvb.RayTracer(room(0.1)) as g:
        assert np.array_equal(g.directions(1234, 100000), rto.directions(1234, 100000))
        assert np.array_equal(g.directions(1234, 1000, base=77), rto.directions(1234, 1000, base=77))


@pytest.mark.parametrize("subdiv,side", [(1, 4), (4, 8), (8, 32)])
def test_closest_hit_matches_cpu_twins(subdiv, side):
    sc = room(0.1, subdiv, side)
    o = rto.Scene(sc)
    n = 20000
    rng = np.random.default_rng(5)
    pos = (rng.uniform(0.05, 0.95, (n, 3)) * np.array(BOX)).astype(np.float32)
    d = rto.directions(7, n)
    with wvb.RayTracer(sc) as g:
        tri_g, t_g = g.closest_hit(pos, d)
    tri_v, t_v = o.closest_hit(pos, d, brute=False)
    tri_b, t_b = o.closest_hit(pos, d, brute=True)
    assert np.array_equal(tri_g, tri_v) and np.array_equal(t_g, t_v)
    assert np.array_equal(tri_g, tri_b) and np.array_equal(t_g, t_b)


@pytest.mark.parametrize("scatter", [0.0, 0.1, 0.6])
@pytest.mark.parametrize("per_wall", [False, True])
def test_trace_reflections_and_histogram(scatter, per_wall):
    sc = room(scatter, subdiv=3, side=8, per_wall=per_wall)
    o = rto.Scene(sc)
    n, depth, keep = 30000, 40, 8
    d = rto.directions(99, n)
    want_h, want_r, want_drop = o.trace(d, SRC, RCV, depth, seed=5, specular_from_step=2, keep_steps=keep)
    with wvb.RayTracer(sc) as g:
        refl, dropped, ms = g.trace(d, SRC, RCV, depth, seed=5, specular_from_step=2, keep_steps=keep)
        got_h = g.histogram()
    assert dropped == want_drop == 0
    assert got_h.shape == want_h.shape
    for f in ("triangle", "keep_going", "receiver_visible"):
        assert np.array_equal(refl[f], want_r[f]), f
    assert np.array_equal(refl["position"], want_r["position"])
    assert_hist_close(got_h, want_h)
    assert ms > 0


def test_device_generated_directions_and_segments_accumulate():
    # two calls with ray_index_base == one call (raytracer.h:246-262 segments + accumulate)
    sc = room(0.3, subdiv=2, side=8)
    o = rto.Scene(sc)
    n, depth = 20000, 30
    d = rto.directions(42, n)
    want_h, _, _ = o.trace(d, SRC, RCV, depth, seed=42)
    with wvb.RayTracer(sc) as g:
        bins = g.safe_bins(depth)
        g.trace(None, SRC, RCV, depth, n_rays=12000, total_rays=n, seed=42, n_bins=bins)
        g.trace(None, SRC, RCV, depth, n_rays=8000, total_rays=n, seed=42, ray_index_base=12000, n_bins=bins)
        got = g.histogram()
        assert_hist_close(got, want_h)
        g.reset_histogram()
        assert not g.histogram().any()


def test_directional_histogram():
    sc = room(0.2, subdiv=2, side=8)
    o = rto.Scene(sc)
    n, depth = 20000, 25
    d = rto.directions(3, n)
    want, _, _ = o.trace(d, SRC, RCV, depth, directional=True)
    with wvb.RayTracer(sc) as g:
        g.trace(d, SRC, RCV, depth, directional=True)
        got = g.histogram()
    assert got.shape == want.shape == (20, 9) + want.shape[2:]
    assert_hist_close(got.sum((0, 1)), want.sum((0, 1)))
    # device atan2f/asinf vs host libm may flip a cell for a direction exactly on a
    # boundary; allow a vanishing fraction of the energy to sit in a neighbouring cell
    moved = np.abs(got - want).sum() / want.sum()
    assert moved < 1e-3


def test_open_scene_dead_rays():
    sc = room(0.1, subdiv=1, side=4)
    keep = np.ones(sc.triangles.size, bool)
    keep[:2] = False
    open_sc = scene.Scene(sc.vertices[:, :3], sc.triangles[keep], sc.surfaces, side=4)
    d = rto.directions(8, 20000)
    want_h, want_r, _ = rto.Scene(open_sc).trace(d, SRC, RCV, 12, keep_steps=12)
    with wvb.RayTracer(open_sc) as g:
        refl, _, _ = g.trace(d, SRC, RCV, 12, keep_steps=12)
        got_h = g.histogram()
    assert not want_r["keep_going"].all()
    assert np.array_equal(refl["keep_going"], want_r["keep_going"])
    assert np.array_equal(refl["position"], want_r["position"])
    assert_hist_close(got_h, want_h)


def test_config5_sized_run_statistics():
    """1 M rays in a many-triangle room (config 5's ray count): energy bookkeeping must
    agree with an independent 256 k-ray oracle run (different seed, different directions)
    within the reference's statistical criterion (10 % per band, equal_energy.cpp:88).
    The diffuse-rain part is compared on its own too: it is a sum over ~10^8 small
    contributions and must agree much more tightly than the specular part, which is a
    few thousand full-energy impulses."""
    sc = scene.box_scene((12.0, 8.0, 20.0), subdiv=12, side=16,
                         surfaces=[scene.make_surface([0.1, 0.1, 0.12, 0.15, 0.2, 0.25, 0.3, 0.35], 0.3)])
    src, rcv = [3.0, 2.0, 4.0], [8.0, 5.0, 15.0]
    depth = 60
    with wvb.RayTracer(sc) as g:
        _, dropped, ms = g.trace(None, src, rcv, depth, n_rays=1 << 20, seed=9)
        big = g.histogram()
        assert dropped == 0 and np.isfinite(big).all() and (big >= 0).all()
        g.reset_histogram()
        g.trace(None, src, rcv, depth, n_rays=1 << 20, seed=9, specular_from_step=depth)
        big_diffuse = g.histogram()
    o = rto.Scene(sc)
    dirs = rto.directions(1234, 1 << 18)
    small, _, _ = o.trace(dirs, src, rcv, depth, seed=77)
    small_diffuse, _, _ = o.trace(dirs, src, rcv, depth, seed=77, specular_from_step=depth)
    assert np.abs(big.sum(0) / small.sum(0) - 1).max() < 0.10
    assert np.abs(big_diffuse.sum(0) / small_diffuse.sum(0) - 1).max() < 0.02


# ---- the wavefront schedule (one launch per reflection, rays re-binned in between) ------------------
from wayverb_b200 import _lib  # noqa: E402


@pytest.mark.parametrize("scatter,per_wall", [(0.0, False), (0.1, True), (0.6, False)])
def test_wavefront_schedule_matches_oracle_and_ray_life(scatter, per_wall):
    sc = room(scatter, subdiv=3, side=8, per_wall=per_wall)
    o = rto.Scene(sc)
    n, depth, keep = 30000, 40, 8
    d = rto.directions(99, n)
    want_h, want_r, want_drop = o.trace(d, SRC, RCV, depth, seed=5, specular_from_step=2, keep_steps=keep)
    with wvb.RayTracer(sc) as g:
        refl, dropped, ms = g.trace(d, SRC, RCV, depth, seed=5, specular_from_step=2, keep_steps=keep,
                                    mode=_lib.RT_MODE_WAVEFRONT)
        got_h = g.histogram()
        g.reset_histogram()
        refl1, _, _ = g.trace(d, SRC, RCV, depth, seed=5, specular_from_step=2, keep_steps=keep,
                              mode=_lib.RT_MODE_RAY_LIFE)
        one_h = g.histogram()
    assert dropped == want_drop == 0 and ms > 0
    assert np.array_equal(refl.view(np.uint8), want_r.view(np.uint8))      # every byte of every record
    assert np.array_equal(refl.view(np.uint8), refl1.view(np.uint8))
    assert_hist_close(got_h, want_h)
    assert_hist_close(got_h, one_h)


def test_wavefront_dead_rays_directional_and_segments():
    # an open box: rays escape and die at different steps; directional histogram; two segments
    sc = room(0.2, subdiv=1, side=4)
    keep = np.ones(sc.triangles.size, bool)
    keep[:2] = False
    open_sc = scene.Scene(sc.vertices[:, :3], sc.triangles[keep], sc.surfaces, side=4)
    n, depth = 40000, 14
    d = rto.directions(8, n)
    want_h, want_r, _ = rto.Scene(open_sc).trace(d, SRC, RCV, depth, keep_steps=depth, directional=True, seed=8)
    with wvb.RayTracer(open_sc) as g:
        bins = want_h.shape[2]
        a, _, _ = g.trace(d[:25000], SRC, RCV, depth, total_rays=n, keep_steps=depth, directional=True, seed=8,
                          n_bins=bins, mode=_lib.RT_MODE_WAVEFRONT)
        b, _, _ = g.trace(d[25000:], SRC, RCV, depth, total_rays=n, ray_index_base=25000, keep_steps=depth,
                          directional=True, seed=8, n_bins=bins, mode=_lib.RT_MODE_WAVEFRONT)
        got_h = g.histogram()
    got_r = np.concatenate([a, b], 1)
    assert (want_r["keep_going"] == 0).any() and (want_r["keep_going"] == 1).any()
    assert np.array_equal(got_r.view(np.uint8), want_r.view(np.uint8))
    assert_hist_close(got_h.sum((0, 1)), want_h.sum((0, 1)))
    assert np.abs(got_h - want_h).sum() / want_h.sum() < 1e-3


def test_wavefront_on_the_concert_hall_with_image_sources():
    sc, meta = scene.concert_hall()
    src, rcv = meta["source"], meta["receiver"]
    depth, order, n = wvb.reflection_depth(meta["min_absorption"]), 4, 70000
    o = rto.Scene(sc)
    d = rto.directions(0x5eed, n)
    want_h, want_r, _ = o.trace(d, src, rcv, depth, seed=0x5eed, specular_from_step=order + 1, keep_steps=order)
    want_i, _ = rto.image_source(o, rto.path_elements(want_r, order), src, rcv)
    with wvb.RayTracer(sc) as g, wvb.ImageSource(g, src, rcv, max_elements=n * order) as s:
        got_r = s.trace(d, depth=depth, order=order, seed=0x5eed, specular_from_step=order + 1, keep_steps=order,
                        mode=_lib.RT_MODE_WAVEFRONT)
        got_h = g.histogram()
        got_i, _, _ = s.results()
    assert np.array_equal(got_r.view(np.uint8), want_r.view(np.uint8))
    assert np.abs(got_h - want_h).max() <= 1e-9 * np.abs(want_h).max()
    assert np.array_equal(got_i.view(np.uint8), want_i.view(np.uint8))


def test_occluder_lying_on_a_voxel_face():
    """the visibility ray stops walking at the receiver's distance, the reference walks
    to the grid's edge; the two agree as long as voxel lists are conservative. A panel lying
    EXACTLY on a voxel face (x = 2.0 with the grid [-0.1, 4.1] / 4) between source and receiver is
    the case where a sloppy list would show: octree-voxelised (the panel is listed on both sides
    of the face) the records, visibility flags included, equal the oracle's full walk."""
    base = scene.box_scene(BOX, subdiv=1, side=4)
    v = [tuple(p[:3]) for p in base.vertices]
    t = [tuple(int(c) for c in r) for r in base.triangles.tolist()]
    n0 = len(v)
    v += [(2.0, 0.5, 1.0), (2.0, 2.5, 1.0), (2.0, 2.5, 5.0), (2.0, 0.5, 5.0)]
    t += [(0, n0, n0 + 1, n0 + 2), (0, n0, n0 + 2, n0 + 3)]
    sc = scene.Scene(np.array(v, np.float32), np.array(t, np.uint32).view(scene.TRI_DT).reshape(-1),
                     [scene.make_surface(0.1, 0.3)], pad=0.1, voxeliser="octree", depth=2)
    assert sc.side == 4 and np.isclose((sc.aabb[3] - sc.aabb[0]) / 4 * 2 + sc.aabb[0], 2.0)
    src, rcv = (0.9, 1.4, 2.8), (3.2, 1.6, 3.1)     # the panel stands between them
    n, depth = 40000, 12
    d = rto.directions(21, n)
    want_h, want_r, _ = rto.Scene(sc).trace(d, src, rcv, depth, seed=21, keep_steps=depth)
    for mode in (_lib.RT_MODE_RAY_LIFE, _lib.RT_MODE_WAVEFRONT):
        with wvb.RayTracer(sc) as g:
            got_r, _, _ = g.trace(d, src, rcv, depth, seed=21, keep_steps=depth, mode=mode)
            got_h = g.histogram()
        assert np.array_equal(got_r.view(np.uint8), want_r.view(np.uint8))
        assert_hist_close(got_h, want_h)
    vis = want_r["receiver_visible"][want_r["keep_going"] == 1]
    assert 0.1 < vis.mean() < 0.9          # the panel hides the receiver from many hits, not from all
    assert (want_r["triangle"] >= 12).any()  # and is hit itself
