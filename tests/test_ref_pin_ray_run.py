"""The reference's OWN `raytracer::run` template, run here: raytracer/raytracer.h:188-266 with
reflector.cpp, stochastic/finder.cpp and the stock reflection processors (make_image_source,
make_stochastic_histogram / make_directional_histogram, make_visual; canonical.cpp's gates), compiled
unmodified from /root/reference over the host-memory cl.hpp stand-in (oracle/ref_recipe/hostcl) and
enqueueing the reference's own `reflections` / `stochastic` kernels as compiled for the host.
No line of the host loop is restated on that side: 16384-ray segments, the reflection depth from the
scene's least absorbent used surface, one engine per reflector::run_step, the image-source tree fed
across segments, per-segment histograms summed in float.

Held against it, bit for bit: the step-by-step loop of oracle/refk.py (RayScene.trace_steps +
reference_histogram + is_image_source) -- which tests/test_ref_pin_rt.py in turn holds the oracle's
rto.trace to, and that is what the GPU tests compare the product with. The one input both sides must
share is the random stream: the reference seeds an engine from std::random_device in every step
(reflector.cpp:13-25); the build spells that name as a device that counts up from a chosen seed, and
refk.ray_direction_rng(seed + k, n) returns the k-th step's numbers through the same function.

This run also shows where the reference touches the caller's directions: `geo::ray{source, d}`
normalises d again (geometric.cpp:13-15), which moves about a quarter of already-unit float vectors by
an ulp. The step-by-step side does the same here; the product's raytracer::run hands directions on as
they come (DESIGN.md "Precision")."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from oracle import refk, rto  # noqa: E402
from wayverb_b200 import scene  # noqa: E402

pytestmark = pytest.mark.skipif(not refk.available(), reason="no /root/reference and no prebuilt oracle/_ref")

SEGMENT = 1 << 14                                            # raytracer.h:219


def glm_normalize(d):
    """glm::normalize in float: v * (1 / sqrt((x*x + y*y) + z*z))"""
    d = np.ascontiguousarray(d, np.float32)
    t = d * d
    inv = np.float32(1.0) / np.sqrt((t[:, 0] + t[:, 1]) + t[:, 2])
    return (d * inv[:, None]).astype(np.float32)


def add_padded(a, b, axis):
    """sum_vectors (stochastic/postprocessing.h:62-70): grow to the longer one, add in float"""
    n = max(a.shape[axis], b.shape[axis])
    out = []
    for x in (a, b):
        pad = [(0, 0)] * x.ndim
        pad[axis] = (0, n - x.shape[axis])
        out.append(np.pad(x, pad))
    return (out[0] + out[1]).astype(np.float32)


def step_by_step(sc, src, rcv, dirs, seed, depth, order, directional=False, radius=0.1, rate=1000.0):
    n = dirs.shape[0]
    energy = refk.hm_ray_energy(n, src, rcv, radius)         # compute_ray_energy(total_rays, ...)
    scene_ = refk.RayScene(sc)
    total, records = None, []
    for segment, lo in enumerate(range(0, n, SEGMENT)):
        d = glm_normalize(dirs[lo:lo + SEGMENT])             # geo::ray's constructor
        m = d.shape[0]
        steps = list(scene_.trace_steps(
            d, src, rcv, depth, lambda s: refk.ray_direction_rng(seed + segment * depth + s, m),
            receiver_radius=radius, initial_energy=energy))
        h = refk.reference_histogram(steps, rcv, histogram_rate=rate, specular_from_step=order + 1,
                                     directional=directional)
        total = h if total is None else add_padded(total, h, 2 if directional else 0)
        records.append(np.stack([s[1] for s in steps]))
    return total, records


def box_three_materials():
    b = scene.box_scene((4.0, 3.0, 6.0), subdiv=1, per_wall_surfaces=True,
                        surfaces=[scene.make_surface(0.1, 0.1), scene.make_surface(0.2, 0.3),
                                  scene.make_surface(0.3, 0.05)])
    return scene.Scene(b.vertices, b.triangles, b.surfaces, voxeliser="octree", depth=5)


def same_impulses(a, b):
    return a.shape == b.shape and all(np.array_equal(a[f], b[f]) for f in ("volume", "position", "distance"))


@pytest.mark.parametrize("directional", [False, True])
def test_box_two_segments(directional):
    sc = box_three_materials()
    src, rcv = (1.1, 1.2, 1.3), (3.0, 2.0, 4.5)
    n, order, seed = 20000, 3, 100                           # 16384 + 3616 rays
    dirs = rto.directions(5, n)
    run = refk.ray_run(sc, src, rcv, dirs, seed, order, directional=directional, visual_items=50)
    # compute_optimum_reflection_number: min absorption 0.1 -> ceil(-6 / log10(0.9))
    assert run["depth"] == 132 and run["segments_reported"] == 1
    hist, records = step_by_step(sc, src, rcv, dirs, seed, run["depth"], order, directional)
    assert run["histogram"].shape == hist.shape and run["histogram"].sum() > 0.1
    assert np.array_equal(run["histogram"], hist)
    first = np.concatenate([r[:order] for r in records], 1)
    assert same_impulses(run["impulses"], refk.is_image_source(sc, first, src, rcv, order))
    assert run["impulses"].size > 20
    # make_visual keeps the first `items` rays of the first segment, every step
    assert run["visual"].shape == (132, 50)
    for f in ("position", "triangle", "keep_going", "receiver_visible"):
        assert np.array_equal(run["visual"][f], records[0][f][:, :50])


def test_concert_hall_image_source_order_four():
    sc, meta = scene.concert_hall(0)
    src, rcv = meta["source"], meta["receiver"]
    n, order, seed = 17000, 4, 7
    dirs = rto.directions(11, n)
    run = refk.ray_run(sc, src, rcv, dirs, seed, order)
    assert run["depth"] == 49                                # the hall's least absorbent surface: 0.25
    hist, records = step_by_step(sc, src, rcv, dirs, seed, 49, order)
    assert np.array_equal(run["histogram"], hist) and hist.sum() > 0
    first = np.concatenate([r[:order] for r in records], 1)
    assert same_impulses(run["impulses"], refk.is_image_source(sc, first, src, rcv, order))


def test_directions_are_renormalised_by_the_reference():
    """what the docstring says about geo::ray: feeding the step-by-step side the raw directions does NOT
    reproduce the run, feeding it glm::normalize(d) does -- and most unit vectors are fixed points"""
    sc = box_three_materials()
    src, rcv = (1.1, 1.2, 1.3), (3.0, 2.0, 4.5)
    dirs = rto.directions(5, 2000)
    fixed = (glm_normalize(dirs) == dirs).all(1)
    assert 0.5 < fixed.mean() < 1.0
    run = refk.ray_run(sc, src, rcv, dirs, 3, 2, visual_items=2000)
    scene_ = refk.RayScene(sc)
    for d, expect_equal in ((glm_normalize(dirs), True), (dirs, False)):
        step0 = next(iter(scene_.trace_steps(d, src, rcv, 1, lambda s: refk.ray_direction_rng(3 + s, 2000))))
        same = np.array_equal(run["visual"][0]["position"], step0[1]["position"])
        assert same == expect_equal
    # rays whose direction is a fixed point agree either way
    step0 = next(iter(scene_.trace_steps(dirs, src, rcv, 1, lambda s: refk.ray_direction_rng(3 + s, 2000))))
    assert np.array_equal(run["visual"][0]["position"][fixed], step0[1]["position"][fixed])
