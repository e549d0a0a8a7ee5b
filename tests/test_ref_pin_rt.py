"""Pins the ray oracle (oracle/rt_oracle.cpp) to the REFERENCE'S OWN kernel source.

oracle/_ref/lib_ref.so holds the reference's `reflections` and `stochastic` kernels
(src/raytracer/src/program.cpp:59-153, stochastic/program.cpp:58-152) with the geometry / voxel /
brdf sources they include (core/src/cl/geometry.cpp, voxel.cpp, raytracer/src/cl/brdf.cpp),
compiled from /root/reference by oracle/ref_recipe/build.py. The host loop around them
(raytracer.h:223-244, reflector.cpp:31-51, finder.h:48-79, stochastic_histogram.h:70-111) is driven
by oracle/refk.py in the reference's order. What OpenCL leaves to the platform -- the random
stream, sin / cos / normalize / dot rounding -- is supplied identically to both sides
(DESIGN.md "Precision").

Asserted: reflection records of EVERY step bit-identical; the energy histogram equal to 1e-12
of its peak (the two sides add the same fp32 impulses into fp64 bins in different orders).

The host binning itself is then taken from the reference too: incremental_histogram
(raytracer/histogram.h:62-81), energy_histogram_sum_functor (stochastic_histogram.h:17-39) and
vector_look_up_table<..., 20, 9>::index with core/src/az_el.cpp, compiled as they are
(refk.reference_histogram, refk.lut_index). The reference sums float bands in impulse order, the
product and the oracle sum the same floats in double: same bins hit, same directional cells, and
values within float accumulation error (2e-5 of the peak) -- the one stated precision difference of
the histogram (DESIGN.md "Precision")."""
import numpy as np
import pytest

from oracle import refk, rto
from wayverb_b200 import scene

pytestmark = pytest.mark.skipif(not refk.available(), reason="no /root/reference and no prebuilt oracle/_ref")

SRC = (1.1, 1.2, 1.3)
RCV = (3.0, 2.0, 4.5)


def refl_fields_equal(a, b):
    return (np.array_equal(a["position"][..., :3].view(np.uint32), b["position"][..., :3].view(np.uint32))
            and np.array_equal(a["triangle"], b["triangle"])
            and np.array_equal(a["keep_going"], b["keep_going"])
            and np.array_equal(a["receiver_visible"], b["receiver_visible"]))


def run_pair(sc, n, depth, seed, specular_from_step=0, radius=0.1):
    o = rto.Scene(sc)
    r = refk.RayScene(sc)
    dirs = rto.directions(seed, n)
    n_bins = 400
    want_h, want_r, want_drop = o.trace(dirs, SRC, RCV, depth, seed=seed, keep_steps=depth, n_bins=n_bins,
                                        specular_from_step=specular_from_step, receiver_radius=radius)
    energy = rto.ray_energy(n, SRC, RCV, radius)
    steps = []
    for step, refl, sto, hit in r.trace_steps(dirs, SRC, RCV, depth, lambda s: rto.step_rng(seed, n, s),
                                              receiver_radius=radius, initial_energy=energy):
        assert refl_fields_equal(refl, want_r[step]), "reflections differ at step %d" % step
        steps.append((step, refl, sto, hit))
    got_h, got_drop = refk.histogram_from_steps(steps, n_bins, specular_from_step=specular_from_step)
    assert got_drop == want_drop
    assert want_h.max() > 0
    assert np.abs(got_h - want_h).max() <= 1e-12 * want_h.max()
    # the reference's own float histogram of the same impulses (it grows as needed: rows beyond
    # n_bins are what the fixed-size side counts as dropped)
    ref_h = refk.reference_histogram(steps, RCV, specular_from_step=specular_from_step)
    rows = min(n_bins, ref_h.shape[0])
    padded = np.zeros((n_bins, 8))
    padded[:rows] = ref_h[:rows]
    assert np.array_equal(padded != 0, want_h != 0)
    assert np.abs(padded - want_h).max() <= 2e-5 * want_h.max()
    assert (want_drop == 0) == (ref_h.shape[0] <= n_bins)
    return want_r, want_h


def test_directional_histogram_cells_are_the_references():
    """directional_energy_histogram<20, 9>: every impulse lands in the reference's cell"""
    sc = scene.box_scene((4.0, 3.0, 6.0), subdiv=2, side=8, surfaces=[scene.make_surface(0.1, 0.3)])
    n, depth, seed, n_bins = 3000, 10, 17, 400
    o, r = rto.Scene(sc), refk.RayScene(sc)
    dirs = rto.directions(seed, n)
    want_h, _, _ = o.trace(dirs, SRC, RCV, depth, seed=seed, n_bins=n_bins, directional=True)
    energy = rto.ray_energy(n, SRC, RCV, 0.1)
    steps = list(r.trace_steps(dirs, SRC, RCV, depth, lambda s: rto.step_rng(seed, n, s), initial_energy=energy))
    ref_h = refk.reference_histogram(steps, RCV, directional=True)
    rows = min(n_bins, ref_h.shape[2])
    padded = np.zeros_like(want_h)
    padded[:, :, :rows] = ref_h[:, :, :rows]
    assert (want_h != 0).sum() > 5000 and (want_h.sum((2, 3)) != 0).sum() > 100      # many cells in use
    assert np.array_equal(padded != 0, want_h != 0)
    assert np.abs(padded - want_h).max() <= 2e-5 * want_h.max()


def test_lut_index_is_the_references():
    rng = np.random.default_rng(23)
    v = rng.standard_normal((200000, 3))
    v = (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)
    special = np.array([[0, 1, 0], [0, -1, 0], [1, 0, 0], [-1, 0, 0], [0, 0, 1], [0, 0, -1],
                        [0, 0.99999994, 0.0003], [1e-8, -1, 0], [0.70710677, 0, 0.70710677],
                        [-0.70710677, 0, -0.70710677], [0.15643446, 0, -0.98768836]], np.float32)
    # cell borders: azimuth multiples of 9 degrees off the 18-degree grid, elevation multiples of 9
    az = np.radians(np.arange(-180, 181, 9, dtype=np.float64))
    el = np.radians(np.arange(-81, 90, 9, dtype=np.float64))
    grid = np.array([[np.sin(-a) * np.cos(e), np.sin(e), -np.cos(-a) * np.cos(e)] for a in az for e in el], np.float32)
    for pts in (v, special, grid):
        a_r, e_r = refk.lut_index(pts)
        a_o, e_o = rto.lut_index(pts)
        assert np.array_equal(a_r, a_o) and np.array_equal(e_r, e_o)
        assert a_r.min() >= 0 and a_r.max() <= 19 and e_r.min() >= 0 and e_r.max() <= 8


@pytest.mark.parametrize("outward", [True, False])
@pytest.mark.parametrize("scatter", [0.0, 0.1, 0.7])
def test_box_reflections_and_histogram(scatter, outward):
    """outward-wound boxes take the diffuse branch on about half the bounces, inward-wound ones
    never (the scalar-signbit quirk) -- both windings must agree."""
    sc = scene.box_scene((4.0, 3.0, 6.0), subdiv=2, side=8, outward=outward,
                         surfaces=[scene.make_surface(0.1, scatter)])
    refl, _ = run_pair(sc, 3000, 12, seed=5)
    if scatter == 0.0:
        # closed box, specular only: rays keep going (reflector_tests.cpp:152) bar the few on an edge
        assert refl["keep_going"].mean() > 0.98


def test_per_wall_surfaces_and_specular_gate():
    surfaces = [scene.make_surface(0.05 + 0.1 * k, 0.05 * k + 0.02) for k in range(6)]
    sc = scene.box_scene((5.56, 3.97, 2.81), subdiv=3, side=16, surfaces=surfaces, per_wall_surfaces=True)
    run_pair(sc, 4000, 10, seed=11, specular_from_step=3, radius=0.3)


def test_open_room_rays_die():
    """Remove a wall: rays that leave get an all-zero reflection and stay dead (program.cpp:77-104)."""
    sc = scene.box_scene((4.0, 3.0, 6.0), subdiv=1, side=4, surfaces=[scene.make_surface(0.2, 0.3)])
    open_sc = scene.Scene(sc.vertices[:, :3], sc.triangles[2:], sc.surfaces, side=4)   # wall z = 0 gone
    refl, _ = run_pair(open_sc, 3000, 8, seed=3)
    dead = refl["keep_going"] == 0
    assert dead.any() and not dead.all()
    first = dead.argmax(0)
    for ray in np.nonzero(dead.any(0))[0][:200]:
        assert dead[first[ray]:, ray].all()   # once dead, always dead


@pytest.mark.parametrize("subdiv,side", [(1, 4), (3, 8)])
def test_closest_hit_voxel_and_brute(subdiv, side):
    """reflector_tests.cpp:98-154 / gpu_geometry_tests.cpp:157-291: the reference's voxel_traversal
    and ray_triangle_intersection against the oracle's, triangle and t bit for bit."""
    sc = scene.box_scene((4.0, 3.0, 6.0), subdiv=subdiv, side=side)
    o, r = rto.Scene(sc), refk.RayScene(sc)
    n = 5000
    rng = np.random.default_rng(5)
    pos = (rng.uniform(0.05, 0.95, (n, 3)) * np.array((4.0, 3.0, 6.0))).astype(np.float32)
    d = rto.directions(7, n)
    for brute in (False, True):
        tri_o, t_o = o.closest_hit(pos, d, brute=brute)
        tri_r, t_r = r.closest_hit(pos, d, brute=brute)
        assert np.array_equal(t_o.view(np.uint32), t_r.view(np.uint32))
        hit = t_r != 0
        assert np.array_equal(tri_o[hit], tri_r[hit])
