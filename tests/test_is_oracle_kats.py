"""Pins the image-source oracle (oracle/is_oracle.inc) with the reference's own
known-answer test, src/raytracer/tests/image_source.cpp:33-115: in a 4 x 3 x 6 m
shoebox with absorption 0.1 every image source of the exact cuboid solution
within 10 m of the receiver must be found from 10 000 random rays, with volume,
position and distance within 1e-4. CPU only."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import rto  # noqa: E402
from wayverb_b200 import scene as S  # noqa: E402

BOX = (4.0, 3.0, 6.0)


def _found_impulses(source, receiver, seed, n_rays=10000, depth=14):
    sc = S.box_scene(BOX, subdiv=1, surfaces=[S.make_surface(0.1, 0.0)], side=8)
    o = rto.Scene(sc)
    dirs = rto.directions(seed, n_rays)
    _, refl, _ = o.trace(dirs, source, receiver, depth=depth, seed=seed, keep_steps=depth)
    elems = rto.path_elements(refl, depth)
    imp, stats = rto.image_source(o, elems, source, receiver)
    return imp, stats, elems


@pytest.mark.parametrize("seed", [1, 2])
def test_exact_shoebox_image_sources_are_found(seed):
    rng = np.random.default_rng(seed)
    source = (rng.random(3) * np.array(BOX) * 0.8 + np.array(BOX) * 0.1).astype(np.float32)
    receiver = (rng.random(3) * np.array(BOX) * 0.8 + np.array(BOX) * 0.1).astype(np.float32)
    exact = rto.exact_shoebox((0, 0, 0), BOX, source, receiver, 0.1, 10.0)
    found, stats, _ = _found_impulses(source, receiver, seed)
    assert stats[2] == 0
    assert found.size > 1
    # check_distances (image_source.cpp:52-59)
    for imps in (exact, found):
        d = np.linalg.norm(imps["position"][:, :3] - receiver, axis=1)
        np.testing.assert_allclose(d, imps["distance"], atol=1e-4)
    missing = 0
    for e in exact:
        near = np.abs(found["distance"] - e["distance"]) < 1e-4
        near &= (np.abs(found["position"][:, :3] - e["position"][:3]) < 1e-4).all(1)
        near &= (np.abs(found["volume"] - e["volume"][0]) < 1e-4).all(1)
        missing += not near.any()
    assert missing == 0, f"{missing} of {exact.size} exact image sources not found"


def test_tree_keeps_the_first_visibility_flag_and_prefix_order():
    """multitree insert semantics (recursive_vector.h:249-255): an element that is
    already in the set is not replaced, so a node's `visible` is the flag of the
    first ray that reached it; output is in pre-order of triangle indices."""
    sc = S.box_scene(BOX, subdiv=1, surfaces=[S.make_surface(0.1, 0.0)], side=8)
    o = rto.Scene(sc)
    source, receiver = (1.0, 1.0, 1.0), (3.0, 2.0, 4.5)
    dirs = rto.directions(7, 2000)
    _, refl, _ = o.trace(dirs, source, receiver, depth=4, seed=7, keep_steps=4)
    e = rto.path_elements(refl, 4)
    full, st_full = rto.image_source(o, e, source, receiver)
    # clearing the flags of all but the first ray of every first-order node must not change anything
    first = {}
    e2 = e.copy()
    for r in range(e.shape[1]):
        t = int(e[0, r] & 0x7FFFFFFF)
        if e[0, r] != rto.IS_NONE and t in first:
            e2[0, r] &= np.uint32(0x7FFFFFFF)
        first.setdefault(t, r)
    again, st2 = rto.image_source(o, e2, source, receiver)
    assert st_full[0] == st2[0] and np.array_equal(full.view(np.uint8), again.view(np.uint8))
    # direct impulse comes last (image_source.cpp:55-58) and has unit volume before the distance factor
    direct = full[-1]
    assert np.allclose(direct["position"][:3], source)
    p = np.sqrt(400.0 / (4 * np.pi)) / direct["distance"]
    np.testing.assert_allclose(direct["volume"], p, rtol=1e-6)
    # reversing the ray order flips which ray is "first" but not the set of tree nodes
    _, st3 = rto.image_source(o, e[:, ::-1], source, receiver)
    assert st3[0] == st_full[0]
