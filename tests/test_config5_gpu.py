"""BASELINE config 5 on real geometry: the reference's demo concert hall (322 triangles, and the
subdivided ~20 k-triangle variant), octree-voxelised like the engine does (depth 5, padding 0.1),
through the C ABI against the CPU oracle: closest hits (voxel walk == brute force, the
reference's core/tests/voxel_tests.cpp 'surrounded'), the stochastic ray loop's reflections and
histogram, the image-source impulses, and the mesh classification that precedes the waveguide."""
import numpy as np
import pytest

import wayverb_b200 as wvb
from wayverb_b200 import scene
from oracle import rto

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("subdiv", [0, 3])
def test_closest_hit_voxel_walk_equals_brute_force(subdiv):
    sc, meta = scene.concert_hall(subdiv)
    o = rto.Scene(sc)
    n = 20000
    rng = np.random.default_rng(11)
    src = np.array(meta["source"], np.float32)
    pos = (src + rng.uniform(-1.0, 1.0, (n, 3))).astype(np.float32)
    d = rto.directions(13, n)
    with wvb.RayTracer(sc) as g:
        tri_g, t_g = g.closest_hit(pos, d)
    tri_v, t_v = o.closest_hit(pos, d, brute=False)
    tri_b, t_b = o.closest_hit(pos, d, brute=True)
    assert np.array_equal(tri_g, tri_v) and np.array_equal(t_g, t_v)
    # every ray from inside the closed hall hits something; the voxel walk finds the brute-force hit
    assert np.array_equal(tri_v, tri_b) and np.array_equal(t_v, t_b)
    assert (tri_g != 0xffffffff).all()


@pytest.mark.parametrize("subdiv,n", [(0, 60000), (3, 20000)])
def test_ray_loop_and_image_sources_match_oracle(subdiv, n):
    sc, meta = scene.concert_hall(subdiv)
    src, rcv = meta["source"], meta["receiver"]
    depth = wvb.reflection_depth(meta["min_absorption"])
    assert depth == 49                      # ceil(-6 / log10(1 - 0.25)), optimum_reflection_number.h:38-40
    order = 4
    o = rto.Scene(sc)
    d = rto.directions(0x5eed, n)
    want_h, want_r, want_drop = o.trace(d, src, rcv, depth, seed=0x5eed, specular_from_step=order + 1,
                                        keep_steps=order)
    want_i, _ = rto.image_source(o, rto.path_elements(want_r, order), src, rcv)
    with wvb.RayTracer(sc) as g, wvb.ImageSource(g, src, rcv, max_elements=n * order) as s:
        got_r = s.trace(d, depth=depth, order=order, seed=0x5eed, specular_from_step=order + 1, keep_steps=order)
        got_h = g.histogram()
        got_i, stats, _ = s.results()
    assert np.array_equal(got_r.view(np.uint8), want_r.view(np.uint8))
    assert got_h.shape == want_h.shape and want_h.max() > 0
    assert np.abs(got_h - want_h).max() <= 1e-9 * np.abs(want_h).max()
    assert got_i.shape == want_i.shape and got_i.size > 1          # the direct path and more
    assert np.array_equal(got_i.view(np.uint8), want_i.view(np.uint8))


def test_closest_triangle_through_the_voxel_grid_equals_brute_force():
    """boundary_coefficient_finder_1d on the 20 608-triangle hall with one surface PER TRIANGLE
    (so the surface index returned names the triangle picked): the device's voxel-grid search
    against the oracle's slow_closest_triangle over all triangles, node for node"""
    from oracle import wgo
    sc, meta = scene.concert_hall(3)
    n_tri = sc.triangles.size
    tri = sc.triangles.copy()
    tri["surface"] = np.arange(n_tri, dtype=np.uint32)
    surf = np.repeat(sc.surfaces[:1], n_tri)
    sc2 = scene.Scene(sc.vertices, tri, surf, pad=0.1, voxeliser="octree", depth=5)
    spacing = 1.1
    lo, hi = sc2.aabb[:3] + 0.1, sc2.aabb[3:] - 0.1
    dims = tuple(int(np.ceil((hi[k] - lo[k]) / spacing)) + 5 for k in range(3))
    mc = (lo - 2 * spacing).astype(np.float32)
    coeffs = [wgo.to_flat(0.25)] * n_tri
    with wvb.RayTracer(sc2) as g:
        m = wvb.build_mesh(dims, mc, spacing, coeffs, scene=g)
    o = rto.Scene(sc2)
    want_inside = o.nodes_inside(mc, dims, spacing)
    z, y, x = np.indices(want_inside.shape)
    pts = np.stack([mc[0] + x.astype(np.float32) * np.float32(spacing), mc[1] + y.astype(np.float32) * np.float32(spacing),
                    mc[2] + z.astype(np.float32) * np.float32(spacing)], -1).reshape(-1, 3)
    surf_o, _ = o.closest_surface(pts)
    om = wgo.mesh_from_inside(want_inside, coeffs, surf_o)
    assert np.array_equal(m.nodes, om.nodes)
    assert m.b[0].shape[0] > 1000
    assert np.array_equal(m.b[0], om.b1) and np.array_equal(m.b[1], om.b2) and np.array_equal(m.b[2], om.b3)
    assert len(set(m.b[0].ravel().tolist())) > 500       # many different triangles were picked


def test_mesh_of_the_hall_matches_oracle_and_steps():
    """the 1 kHz-cutoff waveguide of config 5 at a coarser spacing (the full 1 kHz mesh is
    bench-sized): classification + boundary indices on the device == oracle, and the mesh runs"""
    from oracle import wgo
    sc, meta = scene.concert_hall()
    spacing = 0.75
    lo, hi = sc.aabb[:3] + 0.1, sc.aabb[3:] - 0.1
    dims = tuple(int(np.ceil((hi[k] - lo[k]) / spacing)) + 5 for k in range(3))
    mc = (lo - 2 * spacing).astype(np.float32)
    coeffs = [wgo.to_flat(0.25)]
    with wvb.RayTracer(sc) as g:
        m, inside = wvb.build_mesh(dims, mc, spacing, coeffs, scene=g, return_inside=True)
    o = rto.Scene(sc)
    want_inside = o.nodes_inside(mc, dims, spacing)
    assert np.array_equal(inside.astype(bool), want_inside)
    om = wgo.mesh_from_inside(want_inside, coeffs)
    assert np.array_equal(m.nodes["boundary_type"], om.nodes["boundary_type"])
    assert np.array_equal(m.nodes["boundary_index"], om.nodes["boundary_index"])
    src = m.index(*[int(round((meta["source"][k] - mc[k]) / spacing)) for k in range(3)])
    assert m.nodes["boundary_type"][src] == 1
    sim = wgo.Sim(om)
    sim.write(src, 1.0)
    with wvb.Waveguide(m) as w:
        w.write(src, 1.0)
        # re-entrant geometry may raise the reference's suspicious-boundary flag: same on both sides
        assert w.step(30) == sim.step(30)
        assert np.array_equal(w.field(), sim.field())
