"""GPU parity tests of the image-source stage (wvb_is_*) against the CPU oracle.

fp32 in the host code's operation order on both sides: the impulses (volume,
image-source position, distance) and their ORDER must be bit-identical to the
oracle's restatement of tree.cpp / fast_pressure_calculator.h, whichever way
the paths reach the tree (host elements, host reflection records, or straight
from the trace kernel)."""
import numpy as np
import pytest

import wayverb_b200 as wvb
from wayverb_b200 import scene
from oracle import rto

pytestmark = pytest.mark.gpu

BOX = (4.0, 3.0, 6.0)
SRC = (1.1, 1.2, 1.3)
RCV = (3.0, 2.0, 4.5)


def room(subdiv=1, side=8, per_wall=False, scatter=0.0):
    surfs = [scene.make_surface(0.1, scatter)]
    if per_wall:
        surfs = [scene.make_surface([0.05 + 0.02 * k + 0.01 * b for b in range(8)],
                                    [0.05 * k + 0.01 * b for b in range(8)]) for k in range(3)]
    return scene.box_scene(BOX, subdiv=subdiv, side=side, surfaces=surfs, per_wall_surfaces=per_wall)


def same(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def oracle_paths(sc, n, depth, order, seed):
    o = rto.Scene(sc)
    dirs = rto.directions(seed, n)
    _, refl, _ = o.trace(dirs, SRC, RCV, depth=depth, seed=seed, keep_steps=order)
    return o, dirs, refl, rto.path_elements(refl, order)


@pytest.mark.parametrize("subdiv,side,per_wall,scatter", [(1, 8, False, 0.0), (3, 16, True, 0.0), (2, 4, True, 0.3)])
@pytest.mark.parametrize("order", [1, 4, 9])
def test_impulses_identical_to_oracle(subdiv, side, per_wall, scatter, order):
    sc = room(subdiv, side, per_wall, scatter)
    n = 6000
    o, dirs, refl, elems = oracle_paths(sc, n, depth=10, order=order, seed=3)
    want, wstats = rto.image_source(o, elems, SRC, RCV)
    assert want.size > 1
    with wvb.RayTracer(sc) as g:
        # (a) elements from the host
        with wvb.ImageSource(g, SRC, RCV, max_elements=n * order) as s:
            s.push_elements(elems)
            got, stats, _ = s.results()
            assert same(got, want)
            assert stats[0] == wstats[0] and stats[1] == wstats[1] and stats[2] == wstats[2] == 0 and stats[3] == 0
            again, _, _ = s.results()  # results() is repeatable
            assert same(again, want)
        # (b) reflection records from the host, in two segments (tree::push per group)
        with wvb.ImageSource(g, SRC, RCV, max_elements=n * order) as s:
            h = n // 3
            s.push_reflections(refl[:, :h], 0)
            s.push_reflections(refl[:, h:], h)
            got, _, _ = s.results()
            assert same(got, want)
        # (c) straight from the trace kernel
        with wvb.ImageSource(g, SRC, RCV, max_elements=n * order) as s:
            s.trace(dirs, depth=10, order=order, seed=3)
            got, _, _ = s.results()
            assert same(got, want)


def test_first_ray_decides_visibility_regardless_of_push_order():
    sc = room(2, 8)
    n, order = 4000, 3
    o, _, refl, elems = oracle_paths(sc, n, depth=6, order=order, seed=11)
    want, _ = rto.image_source(o, elems, SRC, RCV)
    with wvb.RayTracer(sc) as g, wvb.ImageSource(g, SRC, RCV, max_elements=n * order) as s:
        # second half first: ray_index_base, not arrival order, decides
        h = n // 2
        s.push_elements(elems[:, h:], h)
        s.push_elements(elems[:, :h], 0)
        got, _, _ = s.results()
        assert same(got, want)


def test_flip_phase_no_direct_and_receiver_outside_line_of_sight():
    sc = room(1, 8)
    n, order = 3000, 5
    o, _, _, elems = oracle_paths(sc, n, depth=6, order=order, seed=5)
    want, _ = rto.image_source(o, elems, SRC, RCV, acoustic_impedance=415.0, flip_phase=True, with_direct=False)
    with wvb.RayTracer(sc) as g, wvb.ImageSource(g, SRC, RCV, n * order, acoustic_impedance=415.0,
                                                 flip_phase=True, with_direct=False) as s:
        s.push_elements(elems)
        got, _, _ = s.results()
        assert same(got, want)
    assert (want["volume"][np.isclose(want["distance"], want["distance"].min())] < 0).all()


def test_exact_shoebox_known_answer_on_the_device():
    """the reference's own test (tests/image_source.cpp:33-115), device end to end"""
    sc = room(1, 8)
    n, depth = 10000, 14
    exact = rto.exact_shoebox((0, 0, 0), BOX, SRC, RCV, 0.1, 10.0)
    with wvb.RayTracer(sc) as g, wvb.ImageSource(g, SRC, RCV, max_elements=n * depth) as s:
        s.trace(None, depth=depth, order=depth, n_rays=n, seed=9)
        found, stats, _ = s.results()
    assert stats[2] == 0 and found.size > 1
    d = np.linalg.norm(found["position"][:, :3] - np.array(RCV, np.float32), axis=1)
    np.testing.assert_allclose(d, found["distance"], atol=1e-4)
    for e in exact:
        near = np.abs(found["distance"] - e["distance"]) < 1e-4
        near &= (np.abs(found["position"][:, :3] - e["position"][:3]) < 1e-4).all(1)
        near &= (np.abs(found["volume"] - e["volume"][0]) < 1e-4).all(1)
        assert near.any()


def test_capacity_and_bad_elements_are_reported():
    sc = room(1, 8)
    with wvb.RayTracer(sc) as g, wvb.ImageSource(g, SRC, RCV, max_elements=100) as s:
        with pytest.raises(wvb._lib.WvbError):
            s.push_elements(np.zeros((2, 100), np.uint32))
        bad = np.full((1, 10), 12345, np.uint32)  # triangle index out of range
        s.push_elements(bad)
        got, stats, _ = s.results()
        assert stats[3] == 10 and stats[0] == 0
        assert got.size == 1  # the direct impulse only
