"""Pins the CPU ray oracle (oracle/rt_oracle.cpp) with the reference's own
CPU-twin style checks, on CPU:

  src/raytracer/tests/reflector_tests.cpp:98-154   voxel-accelerated hit == brute-force hit,
                                                   every ray keeps going in a closed box
  src/raytracer/tests/image_source.cpp:28-111      ray-found reflection points lie on the
                                                   analytic image-source geometry (specular case)
  src/raytracer/tests/equal_energy.cpp:72-90       histogram energy at the direct-path time ~
                                                   image-source direct energy (10 % criterion)
  src/core/include/core/vector_look_up_table.h     direction -> LUT cell mapping
"""
import numpy as np
import pytest

from wayverb_b200 import scene
from oracle import rto

SRC = np.array([2.09, 2.12, 2.12], np.float32)  # reflector_tests.cpp's source/receiver style
RCV = np.array([2.09, 1.08, 0.96], np.float32)  # inside the 4 x 3 x 6 box
BOX = (4.0, 3.0, 6.0)


def test_sincos_accuracy():
    th = np.linspace(-np.pi, np.pi, 200001).astype(np.float32)
    s, c = rto.sincos(th)
    assert np.abs(s - np.sin(th.astype(np.float64))).max() < 2e-7
    assert np.abs(c - np.cos(th.astype(np.float64))).max() < 2e-7
    assert np.abs(s * s + c * c - 1).max() < 5e-7


def test_generated_directions_are_uniform_unit_vectors():
    d = rto.directions(seed=99, n=200000)
    assert np.abs(np.linalg.norm(d.astype(np.float64), axis=1) - 1).max() < 1e-6
    assert np.abs(d.mean(0)).max() < 0.01
    # uniform on the sphere: z (here the y component) uniform in [-1, 1]
    hist, _ = np.histogram(d[:, 1], bins=20, range=(-1, 1))
    assert np.abs(hist / hist.mean() - 1).max() < 0.05
    assert np.array_equal(d[:100], rto.directions(99, 100))           # reproducible
    assert np.array_equal(d[50:100], rto.directions(99, 50, base=50))  # addressable by index


@pytest.mark.parametrize("subdiv,side", [(1, 4), (3, 8), (6, 16)])
def test_voxel_traversal_equals_brute_force(subdiv, side):
    sc = scene.box_scene(BOX, subdiv=subdiv, side=side)
    o = rto.Scene(sc)
    n = 10000
    rng = np.random.default_rng(5)
    pos = (rng.uniform(0.05, 0.95, (n, 3)) * np.array(BOX)).astype(np.float32)
    d = rto.directions(7, n)
    tri_v, t_v = o.closest_hit(pos, d, brute=False)
    tri_b, t_b = o.closest_hit(pos, d, brute=True)
    assert (tri_b != 0xFFFFFFFF).all()       # closed box: every ray hits something
    assert np.array_equal(tri_v, tri_b)
    assert np.array_equal(t_v, t_b)
    hit = pos + d * t_v[:, None]
    on_wall = np.minimum(np.abs(hit), np.abs(hit - np.array(BOX))).min(1)
    assert on_wall.max() < 1e-4


def test_all_rays_keep_going_and_hits_stay_on_walls():
    sc = scene.box_scene(BOX, subdiv=2, side=8, surfaces=[scene.make_surface(0.1, 0.0)])
    o = rto.Scene(sc)
    d = rto.directions(3, 10000)
    _, refl, _ = o.trace(d, SRC, RCV, depth=10, keep_steps=10)
    assert refl["keep_going"].all()
    p = refl["position"][..., :3]
    on_wall = np.minimum(np.abs(p), np.abs(p - np.array(BOX, np.float32))).min(-1)
    assert on_wall.max() < 1e-4


def test_specular_paths_follow_image_source_geometry():
    # scattering 0: a ray's k-th reflection point, the source's k-th order image and the
    # previous reflection point are collinear (unfolding the path gives a straight line),
    # i.e. the path length up to bounce k equals |image_k - hit_k|.
    sc = scene.box_scene(BOX, subdiv=1, side=4, surfaces=[scene.make_surface(0.1, 0.0)])
    o = rto.Scene(sc)
    n = 2000
    d = rto.directions(11, n)
    _, refl, _ = o.trace(d, SRC, RCV, depth=6, keep_steps=6)
    p = refl["position"][..., :3].astype(np.float64)
    tri = refl["triangle"]
    v = sc.vertices[:, :3].astype(np.float64)
    t = sc.triangles
    img = np.tile(SRC.astype(np.float64), (n, 1))
    prev = np.tile(SRC.astype(np.float64), (n, 1))
    length = np.zeros(n)
    for k in range(6):
        length += np.linalg.norm(p[k] - prev, axis=1)
        assert np.abs(np.linalg.norm(p[k] - img, axis=1) - length).max() < 2e-4
        # mirror the image in the plane of the triangle just hit
        a, b, c = v[t["v0"][tri[k]]], v[t["v1"][tri[k]]], v[t["v2"][tri[k]]]
        nrm = np.cross(b - a, c - a)
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        img = img - 2 * np.sum((img - a) * nrm, 1, keepdims=True) * nrm
        prev = p[k]


def test_ray_energy_formula():
    e = rto.ray_energy(1000, SRC, RCV, 0.1)
    dist = np.linalg.norm(SRC.astype(np.float64) - RCV)
    cos_y = np.sqrt(1 - (0.1 / dist) ** 2)
    assert abs(e - 2.0 / (4 * np.pi * 1000 * dist ** 2 * (1 - cos_y))) / e < 1e-4


def test_direct_path_energy_matches_image_source():
    # equal_energy.cpp: rays that pass through the receiver sphere on their first segment
    # deposit (step 0 specular output) N_hit * E_ray ~ 1 / (4 pi d^2) -- the image-source
    # direct intensity -- within the reference's 10 %.
    sc = scene.box_scene(BOX, subdiv=1, side=4, surfaces=[scene.make_surface(0.1, 0.1)])
    o = rto.Scene(sc)
    n = 1 << 19
    d = rto.directions(2024, n)
    hist, _, dropped = o.trace(d, SRC, RCV, depth=1, specular_from_step=0)
    assert dropped == 0
    dist = np.linalg.norm(SRC.astype(np.float64) - RCV)
    bin_direct = int(dist / 340.0 * 1000.0)
    got = hist[bin_direct, 0]
    want = 1.0 / (4 * np.pi * dist ** 2)
    assert abs(got - want) / want < 0.10
    # the gate: with specular_from_step = 1 that contribution is left to the image-source model
    hist2, _, _ = o.trace(d, SRC, RCV, depth=1, specular_from_step=1)
    assert hist2[bin_direct, 0] < 0.5 * got


def test_energy_decays_with_absorption_and_dead_rays_stop():
    sc = scene.box_scene(BOX, subdiv=1, side=4, surfaces=[scene.make_surface(0.3, 0.2)])
    o = rto.Scene(sc)
    d = rto.directions(8, 20000)
    hist, refl, _ = o.trace(d, SRC, RCV, depth=40, keep_steps=2)
    e = hist.sum(1)
    nz = np.nonzero(e)[0]
    assert e[nz[:len(nz) // 4]].sum() > 10 * e[nz[-len(nz) // 4:]].sum()
    assert (hist >= 0).all() and np.isfinite(hist).all()
    # open scene (one wall missing): escaping rays die and stay dead
    keep = sc.triangles["v0"] >= 0
    keep[:2] = False
    open_sc = scene.Scene(sc.vertices[:, :3], sc.triangles[keep], sc.surfaces, side=4)
    _, refl, _ = rto.Scene(open_sc).trace(d, SRC, RCV, depth=12, keep_steps=12)
    kg = refl["keep_going"].astype(bool)
    assert (~kg).any()
    assert not (kg[1:] & ~kg[:-1]).any()                     # never resurrected
    assert not refl["position"][~kg].any() and not refl["triangle"][~kg].any()


def test_lut_cells():
    # table::index, 20 x 9 cells: forward (-z) is azimuth 0, +x is -90 deg after the sign flip
    az, el = rto.lut_index(np.array([[0, 0, -1], [1, 0, 0], [0, 0, 1], [-1, 0, 0], [0, 1, 0], [0, -1, 0]],
                                    np.float32))
    assert az.tolist()[:4] == [0, 15, 10, 5]
    assert el.tolist() == [4, 4, 4, 4, 8, 0]
    v = rto.directions(4, 50000)
    az, el = rto.lut_index(v)
    assert az.min() == 0 and az.max() == 19 and el.min() == 0 and el.max() == 8


def test_directional_histogram_sums_to_plain_histogram():
    sc = scene.box_scene(BOX, subdiv=1, side=4, surfaces=[scene.make_surface(0.2, 0.3)])
    o = rto.Scene(sc)
    d = rto.directions(21, 4000)
    h, _, _ = o.trace(d, SRC, RCV, depth=15)
    hd, _, _ = o.trace(d, SRC, RCV, depth=15, directional=True)
    assert hd.shape[:2] == (20, 9)
    np.testing.assert_allclose(hd.sum((0, 1)), h, rtol=1e-12, atol=1e-18)
