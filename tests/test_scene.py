"""Scene preparation (host code): the OBJ reader and the octree voxeliser of the product
(wvb_obj_parse / wvb_voxelise, csrc/scene_host.cpp) against the oracle restatement of
make_voxelised_scene_data + get_flattened (oracle/scene_oracle.cpp), on the reference's demo
concert hall (BASELINE config 5's geometry), its subdivided variant, box rooms and random
triangle soups; plus the properties the reference's own tests rely on (core/tests/voxel_tests.cpp)."""
import numpy as np
import pytest

from wayverb_b200 import _lib, scene
from oracle import sco


def soup(seed, n, extent=4.0, size=1.0):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-extent, extent, (n, 1, 3))
    v = (c + rng.uniform(-size, size, (n, 3, 3))).reshape(-1, 3).astype(np.float32)
    t = np.zeros(n, scene.TRI_DT)
    t["v0"], t["v1"], t["v2"] = np.arange(n) * 3, np.arange(n) * 3 + 1, np.arange(n) * 3 + 2
    v4 = np.zeros((v.shape[0], 4), np.float32)
    v4[:, :3] = v
    return v4, t


def test_concert_hall_fixture_is_the_reference_model():
    sc, meta = scene.concert_hall()
    assert sc.triangles.size == 322 and sc.vertices.shape[0] == 214       # 322 triangles
    assert sc.side == 32 and sc.voxel_index.size > 32 ** 3
    # ~33 x 15 x 50 m (docs_source/evaluation.md:566-567), padded by 0.1
    ext = sc.aabb[3:] - sc.aabb[:3]
    assert np.allclose(ext, [33.04, 15.61, 50.32], atol=0.02)
    assert sc.material_names == ["FrontColor"]


@pytest.mark.parametrize("subdiv,depth", [(0, 5), (0, 3), (1, 4), (3, 5)])
def test_voxeliser_matches_oracle_on_the_concert_hall(subdiv, depth):
    sc, _ = scene.concert_hall(subdiv)
    aabb, idx, side = scene.voxelise(sc.vertices, sc.triangles, depth, 0.1)
    want_aabb, want = sco.voxelise(sc.vertices, sc.triangles, depth, 0.1)
    assert side == 1 << depth
    assert np.array_equal(aabb, want_aabb)
    assert np.array_equal(idx, want)            # offsets, counts and triangle order
    # layout invariants of get_flattened (voxel_collection.cpp:9-37)
    n = side ** 3
    offs = idx[:n]
    assert offs[0] == n and np.all(np.diff(offs.astype(np.int64)) >= 1)
    counts = idx[offs]
    assert offs[-1] + 1 + counts[-1] == idx.size
    assert np.array_equal(np.diff(offs.astype(np.int64)), counts[:-1].astype(np.int64) + 1)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_voxeliser_matches_oracle_on_triangle_soups(seed):
    v, t = soup(seed, 400, size=[0.05, 0.8, 3.0][seed - 1])
    for depth, pad in ((4, 0.1), (2, 0.3), (0, 0.0)):
        aabb, idx, _ = scene.voxelise(v, t, depth, pad)
        want_aabb, want = sco.voxelise(v, t, depth, pad)
        assert np.array_equal(aabb, want_aabb) and np.array_equal(idx, want)


def test_degenerate_triangles_are_listed_nowhere_like_the_reference():
    # zero-area triangle: normalize() gives nan and `fabs(dist) <= r` is false
    # (tri_cube_intersection.cpp:165-169), so the octree drops it at the root
    v, t = soup(5, 10)
    v[3:6] = v[3]                      # triangle 1 collapses to a point
    _, idx, side = scene.voxelise(v, t, 3, 0.1)
    _, want = sco.voxelise(v, t, 3, 0.1)
    assert np.array_equal(idx, want)
    n = side ** 3
    listed = set()
    for o in idx[:n]:
        listed.update(idx[o + 1:o + 1 + idx[o]].tolist())
    assert 1 not in listed and 0 in listed


def test_lists_are_conservative():
    """every point of a triangle lies in a voxel that lists the triangle (what voxel_tests.cpp
    'surrounded' needs): sample points on the concert hall's triangles"""
    sc, _ = scene.concert_hall()
    side, n = sc.side, sc.side ** 3
    rng = np.random.default_rng(0)
    v = sc.vertices[:, :3].astype(np.float64)
    t = sc.triangles
    lo, ext = sc.aabb[:3].astype(np.float64), (sc.aabb[3:] - sc.aabb[:3]).astype(np.float64)
    cells = [set(sc.voxel_index[o + 1:o + 1 + sc.voxel_index[o]].tolist()) for o in sc.voxel_index[:n]]
    for ti in range(t.size):
        a, b, c = v[t["v0"][ti]], v[t["v1"][ti]], v[t["v2"][ti]]
        w = rng.dirichlet([1, 1, 1], 40)
        p = w[:, :1] * a + w[:, 1:2] * b + w[:, 2:] * c
        ijk = np.clip(np.floor((p - lo) / ext * side).astype(int), 0, side - 1)
        for x, y, z in ijk:
            assert ti in cells[(x * side + y) * side + z], (ti, x, y, z)


def test_obj_reader():
    text = """# comment
mtllib x.mtl
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
vn 0 0 1
f 1 2 3
usemtl wall
f 1/1/1 2/2/1 3/3/1 4/4/1
v 0 0 1
usemtl floor
f -1 -2 -3
usemtl wall
f 1//1 3//1 5//1
"""
    v, t, names = scene.parse_obj(text)
    assert names == ["default", "wall", "floor"]
    assert v.shape == (5, 4) and np.array_equal(v[4, :3], [0, 0, 1])
    got = [tuple(int(x) for x in r) for r in t.tolist()]
    assert got == [(0, 0, 1, 2), (1, 0, 1, 2), (1, 0, 2, 3), (2, 4, 3, 2), (1, 0, 2, 4)]
    for bad in ("v 0 0\nf 1 1 1\n", "v 0 0 0\nf 1 2 3\n", "v 0 0 0\nv 1 0 0\nf 1 2\n", "# nothing\n"):
        with pytest.raises(_lib.WvbError):
            scene.parse_obj(bad)


def test_load_obj_builds_the_engine_defaults(tmp_path):
    p = tmp_path / "room.obj"
    sc0 = scene.box_scene((4.0, 3.0, 6.0))
    lines = ["v %r %r %r" % tuple(float(c) for c in v[:3]) for v in sc0.vertices]
    lines += ["f %d %d %d" % (a + 1, b + 1, c + 1) for _, a, b, c in sc0.triangles.tolist()]
    p.write_text("\n".join(lines) + "\n")
    sc = scene.load_obj(str(p))
    assert sc.side == 32 and sc.triangles.size == 12
    assert np.allclose(sc.aabb, [-0.1, -0.1, -0.1, 4.1, 3.1, 6.1])      # padded by 0.1
    _, want = sco.voxelise(sc.vertices, sc.triangles, 5, 0.1)
    assert np.array_equal(sc.voxel_index, want)


def test_obj_reader_survives_damaged_files():
    """Whatever bytes arrive, wvb_obj_parse either refuses (WvbError) or returns triangles whose vertex
    and material indices are in range -- never a crash, never an index the ray kernels would follow
    out of the arrays. Hand-written corner cases, then 600 random mutations of the concert hall."""
    import random
    rnd = random.Random(7)
    accepted = refused = 0

    def parse(text):
        nonlocal accepted, refused
        try:
            v, t, names = scene.parse_obj(text)
        except _lib.WvbError:
            refused += 1
            return
        accepted += 1
        if t.size:
            assert max(t["v0"].max(), t["v1"].max(), t["v2"].max()) < v.shape[0]
            assert t["surface"].max() < max(len(names), 1)

    tri = "v 0 0 0\nv 1 0 0\nv 0 1 0\n"
    for text in ["", "\n", "v", "v 1", "v 1 2", "f 1 2 3", tri + "f -1 -2 -3", tri + "f 1 2 99999999999999999999",
                 tri + "f 1/1/1 2//2 3/3", tri + "f 1 2", "v nan inf -inf\nv 1e999 1 1\nv 1 1 1\nf 1 2 3",
                 "usemtl\nusemtl a\nusemtl a\n" + tri + "f 1 2 3 1 2 3 1 2 3", "f 0 0 0", tri + "f 0 1 2", "f / / /",
                 "f 1/ 2/ 3/", "v 1 2 3 4 5 6 7\n", "v\t1\t2\t3\r\nv 1 2 3\r\nv 3 2 1\r\nf 1 2 3\r\n", "\x00\x00v 1 2 3",
                 "v 0 0 0\n" * 5 + "f " + " ".join(str(i % 5 + 1) for i in range(10000)), "f " + "1" * 5000,
                 "v " + "9" * 5000 + " 1 1", "usemtl " + "x" * 100000, tri + "f 4294967297 2 3",
                 tri + "f -4294967297 2 3"]:
        parse(text)
    lines = open(scene.CONCERT_OBJ).read().splitlines()
    words = ["1", "-1", "0", "99999", "1/2/3", "a", "1e400", "-", "//", "4294967296", "-4294967297"]
    for _ in range(600):
        ls = list(lines)
        for _ in range(rnd.randint(1, 8)):
            i = rnd.randrange(len(ls))
            r = rnd.random()
            if r < 0.2:
                del ls[i]
            elif r < 0.4:
                ls[i] = ls[i][:rnd.randrange(len(ls[i]) + 1)]
            elif r < 0.6:
                ls[i] = rnd.choice(["v", "f", "vt", "vn", "usemtl", "g", "#", "l"]) + " " + \
                    " ".join(rnd.choice(words) for _ in range(rnd.randint(0, 6)))
            elif r < 0.8 and ls[i]:
                b = bytearray(ls[i].encode("latin1"))
                b[rnd.randrange(len(b))] = rnd.randrange(256)
                ls[i] = b.decode("latin1")
            else:
                ls.insert(i, ls[rnd.randrange(len(ls))])
        parse("\n".join(ls).encode("latin1"))
    assert accepted > 50 and refused > 50
