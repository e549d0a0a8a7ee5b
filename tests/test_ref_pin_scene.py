"""Pins the scene-preparation oracle (oracle/scene_oracle.cpp) AND the product's voxeliser
(wvb_voxelise) to the REFERENCE'S OWN host source: src/core/src/geo/tri_cube_intersection.cpp,
box.cpp, ndim_tree.h, voxel_collection.h/.cpp (get_flattened) and voxelised_scene_data.h
(make_voxelised_scene_data), whole files compiled unmodified from /root/reference into oracle/_ref behind a
stand-in for the GLM operations they use (oracle/ref_recipe/hoststubs/glm/glm.hpp -- the one
place where this repository, not the reference, decides arithmetic: componentwise operators,
dot, cross, normalize, min/max, written after GLM 0.9.8.1's generic code paths). Voxel for voxel,
triangle for triangle."""
import numpy as np
import pytest

from oracle import refk, sco
from wayverb_b200 import scene

pytestmark = pytest.mark.skipif(not refk.available(), reason="no /root/reference and no prebuilt oracle/_ref")


def soup(seed, n, size):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-4, 4, (n, 1, 3))
    v = (c + rng.uniform(-size, size, (n, 3, 3))).reshape(-1, 3).astype(np.float32)
    t = np.zeros(n, scene.TRI_DT)
    t["v0"], t["v1"], t["v2"] = np.arange(n) * 3, np.arange(n) * 3 + 1, np.arange(n) * 3 + 2
    v4 = np.zeros((v.shape[0], 4), np.float32)
    v4[:, :3] = v
    return v4, t


@pytest.mark.parametrize("subdiv,depth", [(0, 5), (0, 2), (2, 5), (3, 4)])
def test_concert_hall_voxelisation_is_the_references(subdiv, depth):
    sc, _ = scene.concert_hall(subdiv)
    ref_aabb, ref_idx = refk.voxelise(sc.vertices, sc.triangles, depth, 0.1)
    ora_aabb, ora_idx = sco.voxelise(sc.vertices, sc.triangles, depth, 0.1)
    got_aabb, got_idx, side = scene.voxelise(sc.vertices, sc.triangles, depth, 0.1)
    assert ref_idx.size > (1 << depth) ** 3
    assert np.array_equal(ref_aabb, ora_aabb) and np.array_equal(ref_idx, ora_idx)     # oracle == reference
    assert np.array_equal(ref_aabb, got_aabb) and np.array_equal(ref_idx, got_idx)     # product == reference


@pytest.mark.parametrize("seed,size", [(1, 0.05), (2, 0.8), (3, 3.0)])
def test_triangle_soups_and_the_overlap_predicate(seed, size):
    v, t = soup(seed, 300, size)
    for depth, pad in ((4, 0.1), (1, 0.3)):
        ref_aabb, ref_idx = refk.voxelise(v, t, depth, pad)
        _, ora_idx = sco.voxelise(v, t, depth, pad)
        _, got_idx, _ = scene.voxelise(v, t, depth, pad)
        assert np.array_equal(ref_idx, ora_idx) and np.array_equal(ref_idx, got_idx)
    # geo::overlaps(padded(box, 0.001), triangle) on boxes that graze the triangles
    rng = np.random.default_rng(seed)
    tri = v[:, :3].reshape(-1, 3, 3)
    hits = 0
    for k in range(400):
        a = tri[rng.integers(tri.shape[0])]
        centre = a[rng.integers(3)] + rng.normal(0, 0.3, 3)
        half = rng.uniform(0.05, 0.6, 3)
        box = np.concatenate([centre - half, centre + half]).astype(np.float32)
        want = refk.overlaps(box, a.reshape(9))
        assert sco.overlaps(box, a.reshape(9)) == want
        hits += want
    assert 20 < hits < 380          # both outcomes exercised
