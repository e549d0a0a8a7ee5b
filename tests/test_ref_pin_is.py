"""Pins the image-source oracle (oracle/is_oracle.inc) to the REFERENCE'S OWN host source:
raytracer/src/image_source/{tree,postprocess_branches,exact}.cpp with image_source/*.h, multitree.h,
core/recursive_vector.h, core/surfaces.h, fast_pressure_calculator.h, get_direct.h,
reflection_path_builder.h; core/src/geo/{geometric,box,triangle_vec,tri_cube_intersection}.cpp; the CPU
voxel walk of core/src/spatial_division/voxel_collection.cpp with voxelised_scene_data.h;
core/src/pressure_intensity.cpp -- compiled unmodified from /root/reference into oracle/_ref behind the
GLM stand-in (oracle/ref_recipe/hoststubs/glm/glm.hpp, the one place where this repository decides
arithmetic: componentwise operators, dot, cross, normalize, length, distance, mix, comparisons).
Impulse for impulse, in the reference's order, bit for bit."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import rto  # noqa: E402
from oracle import refk  # noqa: E402
from wayverb_b200 import scene as S  # noqa: E402

pytestmark = pytest.mark.skipif(not refk.available(), reason="no /root/reference and no prebuilt oracle/_ref")

BOX = (4.0, 3.0, 6.0)


def same_impulses(a, b):
    """every field the reference defines (the 12 tail bytes of the 64-byte record are padding)"""
    return a.shape == b.shape and all(np.array_equal(a[f], b[f]) for f in ("volume", "position", "distance"))


def octree_box(surfaces, per_wall=False):
    b = S.box_scene(BOX, subdiv=1, surfaces=surfaces, per_wall_surfaces=per_wall)
    return S.Scene(b.vertices, b.triangles, b.surfaces, voxeliser="octree", depth=5)


@pytest.mark.parametrize("subdiv,n,order", [(0, 60000, 4), (2, 20000, 5), (3, 20000, 4)])
def test_concert_hall_image_sources_are_the_references(subdiv, n, order):
    sc, meta = S.concert_hall(subdiv)
    src, rcv = meta["source"], meta["receiver"]
    o = rto.Scene(sc)
    d = rto.directions(0x5eed + subdiv, n)
    _, refl, _ = o.trace(d, src, rcv, 49, seed=0x5eed, specular_from_step=order + 1, keep_steps=order)
    want, stats = rto.image_source(o, rto.path_elements(refl, order), src, rcv)
    ref = refk.is_image_source(sc, refl, src, rcv, order)
    assert stats[0] > 50000 and ref.size > 5          # a big tree, the direct path and more
    assert same_impulses(want, ref)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_shoebox_deep_tree_three_surfaces(seed):
    """order 14 in a closed box: ~90 000 tree nodes, ~570 valid image sources, three materials
    with scattering (fast_pressure_calculator.h:33-62)"""
    sc = octree_box([S.make_surface(0.1, 0.0), S.make_surface(0.3, 0.2), S.make_surface(0.05, 0.5)], per_wall=True)
    o = rto.Scene(sc)
    rng = np.random.default_rng(seed)
    src = (rng.random(3) * np.array(BOX) * 0.8 + np.array(BOX) * 0.1).astype(np.float32)
    rcv = (rng.random(3) * np.array(BOX) * 0.8 + np.array(BOX) * 0.1).astype(np.float32)
    d = rto.directions(seed, 10000)
    _, refl, _ = o.trace(d, src, rcv, 14, seed=seed, keep_steps=14)
    want, stats = rto.image_source(o, rto.path_elements(refl, 14), src, rcv)
    ref = refk.is_image_source(sc, refl, src, rcv, 14)
    assert stats[2] == 0 and ref.size > 300
    assert same_impulses(want, ref)
    # flip_phase and no direct contribution (postprocess_branches' flag; get_direct left out)
    want2, _ = rto.image_source(o, rto.path_elements(refl, 6), src, rcv, flip_phase=True, with_direct=False)
    ref2 = refk.is_image_source(sc, refl, src, rcv, 6, flip_phase=True, with_direct=False)
    assert ref2.size > 50 and (ref2["volume"] < 0).any()
    assert same_impulses(want2, ref2)


def test_source_equal_to_receiver_and_blocked_direct_path():
    sc = octree_box([S.make_surface(0.1, 0.0)])
    o = rto.Scene(sc)
    p = np.array([1.0, 1.0, 1.0], np.float32)
    d = rto.directions(5, 3000)
    _, refl, _ = o.trace(d, p, p, 4, seed=5, keep_steps=4)
    want, _ = rto.image_source(o, rto.path_elements(refl, 4), p, p)      # get_direct: source == receiver -> none
    ref = refk.is_image_source(sc, refl, p, p, 4)
    assert ref.size > 5 and same_impulses(want, ref)
    # receiver outside the box: the direct ray is stopped by a wall (get_direct.h:30-32)
    out = np.array([1.0, 1.0, 9.0], np.float32)
    _, refl, _ = o.trace(d, p, out, 3, seed=5, keep_steps=3)
    want, _ = rto.image_source(o, rto.path_elements(refl, 3), p, out)
    ref = refk.is_image_source(sc, refl, p, out, 3)
    assert same_impulses(want, ref)


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_exact_shoebox_solution_is_the_references(seed):
    rng = np.random.default_rng(seed)
    src = (rng.random(3) * np.array(BOX)).astype(np.float32)
    rcv = (rng.random(3) * np.array(BOX)).astype(np.float32)
    for absorption, reach in ((0.1, 10.0), (0.4, 17.5)):
        want = rto.exact_shoebox((0, 0, 0), BOX, src, rcv, absorption, reach)
        ref = refk.is_exact_shoebox((0, 0, 0), BOX, src, rcv, absorption, reach)
        assert ref.size > 40 and same_impulses(want, ref)


def test_cpu_voxel_walk_finds_the_brute_force_hit():
    """the reference's HOST walk -- intersects(voxelised, ray) (voxelised_scene_data.h:80-106,
    voxel_collection.cpp:41-124) -- against the oracle's restatement of its DEVICE walk
    (core/src/cl/voxel.cpp, itself pinned in test_ref_pin_rt.py): same triangle for every ray."""
    sc, meta = S.concert_hall(0)
    d = rto.directions(9, 4000)
    rng = np.random.default_rng(9)
    origins = (np.asarray(meta["source"], np.float32) + rng.uniform(-1.0, 1.0, (d.shape[0], 3))).astype(np.float32)
    t_cpu, tri_cpu = refk.is_intersects(sc, origins, d[:, :3])
    tri_o, t_o = rto.Scene(sc).closest_hit(origins, d)
    assert (tri_cpu != 0xffffffff).all()
    assert np.array_equal(tri_cpu, tri_o.astype(np.uint32))
    np.testing.assert_allclose(t_cpu, t_o, rtol=2e-6)
