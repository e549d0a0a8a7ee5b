"""Synthetic scenes in the reference's voxelised-scene layout (host, numpy).

The ray path consumes `voxelised_scene_data` exactly as the reference uploads it
(scene_buffers.h:14-38): vertices as cl_float3 (16 B), `triangle{surface,v0,v1,v2}`
(16 B), `surface{absorption[8], scattering[8]}` (64 B), the voxel grid's AABB and
side, and the flattened voxel index of voxel_collection.cpp:9-37:
`index[x*side*side + y*side + z]` = offset of a run `[count, tri, tri, ...]`.
Two ways to build that structure:
  voxeliser="octree"  make_voxelised_scene_data's octree + get_flattened
                      (voxelised_scene_data.h:27-71, voxel_collection.cpp:9-37) through
                      wvb_voxelise (csrc/scene_host.cpp): aabb = vertex bounds padded by
                      `pad`, side = 2^depth. What `load_obj` / `concert_hall` use.
  voxeliser="bbox"    (default, numpy) a fixed-side grid where a triangle is listed in every
                      voxel its bounding box touches: a superset of the overlap set; the
                      traversal accepts a hit only inside the voxel it is visiting
                      (voxel.cpp:80-90), so supersets do not change results. Used by the
                      synthetic box scenes of the tests.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

TRI_DT = np.dtype([("surface", "<u4"), ("v0", "<u4"), ("v1", "<u4"), ("v2", "<u4")])
SURF_DT = np.dtype([("absorption", "<f4", (8,)), ("scattering", "<f4", (8,))])
REFL_DT = np.dtype([("position", "<f4", (4,)), ("triangle", "<u4"), ("keep_going", "i1"),
                    ("receiver_visible", "i1"), ("pad", "i1", (10,))])
assert TRI_DT.itemsize == 16 and SURF_DT.itemsize == 64 and REFL_DT.itemsize == 32


def voxelise(vertices4, triangles, depth=5, pad=0.1):
    """wvb_voxelise: make_voxelised_scene_data(scene, depth, pad) + get_flattened
    -> (aabb[6] float32, flattened voxel index uint32, side)."""
    from ._lib import check, lib, ptr
    v = np.ascontiguousarray(vertices4, np.float32).reshape(-1, 4)
    t = np.ascontiguousarray(triangles, TRI_DT).reshape(-1)
    lo, hi = (C.c_float * 3)(), (C.c_float * 3)()
    n = C.c_uint64()
    check(lib().wvb_voxelise(ptr(v), v.shape[0], ptr(t), t.size, int(depth), float(pad), C.byref(lo), C.byref(hi),
                             None, 0, C.byref(n)))
    out = np.zeros(n.value, np.uint32)
    check(lib().wvb_voxelise(ptr(v), v.shape[0], ptr(t), t.size, int(depth), float(pad), C.byref(lo), C.byref(hi),
                             ptr(out), out.size, C.byref(n)))
    return np.array(list(lo) + list(hi), np.float32), out, 1 << int(depth)


def parse_obj(text):
    """wvb_obj_parse -> (vertices [n,4] float32, triangles TRI_DT (surface = material index),
    material names)."""
    from ._lib import check, lib, ptr
    raw = text.encode() if isinstance(text, str) else bytes(text)
    nv, nt, nn = C.c_uint64(), C.c_uint64(), C.c_uint64()
    check(lib().wvb_obj_parse(raw, len(raw), None, C.byref(nv), None, C.byref(nt), None, C.byref(nn)))
    v = np.zeros((nv.value, 4), np.float32)
    t = np.zeros(nt.value, TRI_DT)
    names = C.create_string_buffer(max(nn.value, 1))
    check(lib().wvb_obj_parse(raw, len(raw), ptr(v), C.byref(nv), ptr(t), C.byref(nt), names, C.byref(nn)))
    return v, t, names.raw[:nn.value].decode("utf-8", "replace").split("\n")[:-1]


class Scene:
    def __init__(self, vertices, triangles, surfaces, side=16, pad=0.1, voxeliser="bbox", depth=None):
        v = np.asarray(vertices, np.float32)
        v = v.reshape(-1, v.shape[-1])[:, :3]
        self.vertices = np.zeros((v.shape[0], 4), np.float32)  # cl_float3 = 16 bytes
        self.vertices[:, :3] = v
        self.triangles = np.ascontiguousarray(triangles, TRI_DT).reshape(-1)
        self.surfaces = np.ascontiguousarray(surfaces, SURF_DT).reshape(-1)
        if voxeliser == "octree":
            d = int(depth if depth is not None else round(np.log2(side)))
            self.aabb, self.voxel_index, self.side = voxelise(self.vertices, self.triangles, d, pad)
            return
        self.side = int(side)
        lo = v.min(0) - np.float32(pad)
        hi = v.max(0) + np.float32(pad)
        self.aabb = np.concatenate([lo, hi]).astype(np.float32)
        self.voxel_index = self._voxelise()

    def _voxelise(self):
        side = self.side
        lo, hi = self.aabb[:3].astype(np.float64), self.aabb[3:].astype(np.float64)
        vd = (hi - lo) / side
        v = self.vertices[:, :3].astype(np.float64)
        t = self.triangles
        tv = np.stack([v[t["v0"]], v[t["v1"]], v[t["v2"]]], 1)  # [n,3,3]
        eps = 1e-4 * vd
        i0 = np.clip(np.floor((tv.min(1) - eps - lo) / vd).astype(int), 0, side - 1)
        i1 = np.clip(np.floor((tv.max(1) + eps - lo) / vd).astype(int), 0, side - 1)
        cells = [[] for _ in range(side ** 3)]
        for ti in range(t.size):
            for x in range(i0[ti, 0], i1[ti, 0] + 1):
                for y in range(i0[ti, 1], i1[ti, 1] + 1):
                    base = x * side * side + y * side
                    for z in range(i0[ti, 2], i1[ti, 2] + 1):
                        cells[base + z].append(ti)
        out = np.zeros(side ** 3 + sum(len(c) + 1 for c in cells), np.uint32)
        pos = side ** 3
        for ci, c in enumerate(cells):
            out[ci] = pos
            out[pos] = len(c)
            out[pos + 1:pos + 1 + len(c)] = c
            pos += 1 + len(c)
        return out

    @property
    def diagonal(self):
        return float(np.linalg.norm(self.aabb[3:] - self.aabb[:3]))


def make_surface(absorption, scattering):
    s = np.zeros((), SURF_DT)
    s["absorption"] = absorption
    s["scattering"] = scattering
    return s


def box_scene(size=(5.56, 3.97, 2.81), subdiv=1, surfaces=None, side=16, outward=True, per_wall_surfaces=False):
    """Shoebox room [0,size] with each wall split into subdiv x subdiv quads (2 triangles each).
    outward=True winds triangles so that normals point out of the room (like geo::get_scene_data boxes)."""
    sx, sy, sz = size
    verts, tris = [], []
    if surfaces is None:
        surfaces = [make_surface(0.1, 0.1)]

    def wall(origin, eu, ev, surf, flip):
        base = len(verts)
        n = subdiv
        for i in range(n + 1):
            for j in range(n + 1):
                verts.append(origin + eu * (i / n) + ev * (j / n))
        for i in range(n):
            for j in range(n):
                a = base + i * (n + 1) + j
                b, c, d = a + (n + 1), a + (n + 1) + 1, a + 1
                for tri in ((a, b, c), (a, c, d)):
                    tri = tri[::-1] if flip else tri
                    tris.append((surf,) + tri)

    o = np.zeros(3)
    ex, ey, ez = np.array([sx, 0, 0.0]), np.array([0, sy, 0.0]), np.array([0, 0, sz])
    walls = [  # origin, eu, ev chosen so that eu x ev points INTO the room
        (o, ex, ey), (o + ez, ey, ex),   # z = 0 (normal +z), z = sz (normal -z)
        (o, ey, ez), (o + ex, ez, ey),   # x = 0 (+x), x = sx (-x)
        (o, ez, ex), (o + ey, ex, ez),   # y = 0 (+y), y = sy (-y)
    ]
    for k, (org, eu, ev) in enumerate(walls):
        surf = k % len(surfaces) if per_wall_surfaces else 0
        wall(org, eu, ev, surf, flip=outward)
    return Scene(np.array(verts), np.array(tris, dtype=np.uint32).view(TRI_DT).reshape(-1), surfaces, side=side)


def load_obj(path_or_text, surfaces=None, depth=5, pad=0.1) -> "Scene":
    """scene_data_loader + make_voxelised_scene_data(scene, 5, 0.1f) (threaded_engine.cpp:123):
    an OBJ file (or its text) -> voxelised Scene. `surfaces`: one SURF_DT per material, or one
    for all; default: absorption 0.1 / scattering 0.1 everywhere."""
    text = path_or_text
    if isinstance(path_or_text, str) and "\n" not in path_or_text:
        with open(path_or_text) as f:
            text = f.read()
    v, t, names = parse_obj(text)
    if surfaces is None:
        surfaces = [make_surface(0.1, 0.1)]
    surfaces = np.ascontiguousarray(surfaces, SURF_DT).reshape(-1)
    if surfaces.size == 1 and len(names) > 1:
        surfaces = np.repeat(surfaces, len(names))
    assert surfaces.size >= len(names), "one surface per material (%d) needed" % len(names)
    sc = Scene(v, t, surfaces, pad=pad, voxeliser="octree", depth=depth)
    sc.material_names = names
    return sc


def subdivide(vertices4, triangles, times=1):
    """Each triangle -> 4 (edge midpoints), `times` times: the reference tree holds only the
    322-triangle concert hall; BASELINE config 5 speaks of ~30 k triangles, which this produces
    from the same surfaces (322 * 4^3 = 20 608)."""
    v = [tuple(p) for p in np.asarray(vertices4, np.float32)[:, :3]]
    t = np.ascontiguousarray(triangles, TRI_DT).reshape(-1)
    for _ in range(times):
        mid, out = {}, []

        def m(a, b):
            k = (a, b) if a < b else (b, a)
            if k not in mid:
                pa, pb = np.float32(v[a]), np.float32(v[b])
                v.append(tuple((pa + pb) * np.float32(0.5)))
                mid[k] = len(v) - 1
            return mid[k]
        for s, a, b, c in t.tolist():
            ab, bc, ca = m(a, b), m(b, c), m(c, a)
            out += [(s, a, ab, ca), (s, ab, b, bc), (s, ca, bc, c), (s, ab, bc, ca)]
        t = np.array(out, np.uint32).view(TRI_DT).reshape(-1)
    vv = np.zeros((len(v), 4), np.float32)
    vv[:, :3] = np.array(v, np.float32)
    return vv, t


CONCERT_OBJ = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "concert_hall.obj")


def concert_hall(subdivisions=0):
    """BASELINE config 5's geometry: the reference's demo concert hall (322 triangles,
    ~33 x 15 x 50 m; tests/golden/concert_hall.obj is its geometry, see make_concert_hall.py),
    set up as the reference's own evaluation does (docs_source/evaluation.md:566-572): wall
    absorption rising from 0.25 in the lowest band to 0.67 in the highest, scattering 0.1, the
    receiver 10 m along x and z from the source. Octree depth 5, padding 0.1 like the engine."""
    absorption = [float(a) for a in np.linspace(0.25, 0.67, 8)]
    surf = make_surface(absorption, 0.1)
    with open(CONCERT_OBJ) as f:
        v, t, names = parse_obj(f.read())
    if subdivisions:
        v, t = subdivide(v, t, subdivisions)
    sc = Scene(v, t, np.repeat(np.ascontiguousarray(surf, SURF_DT).reshape(1), max(len(names), 1)),
               pad=0.1, voxeliser="octree", depth=5)
    sc.material_names = names
    meta = {"name": "concert hall (reference demo model, %d triangles)" % t.size,
            "min_absorption": min(absorption), "absorption": absorption, "scattering": 0.1,
            "source": [-5.0, 0.6, -20.0], "receiver": [5.0, -1.2, -10.0]}  # ~1.5 m above the raked floor
    return sc, meta
