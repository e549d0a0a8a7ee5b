"""Synthetic scenes in the reference's voxelised-scene layout (host, numpy).

The ray path consumes `voxelised_scene_data` exactly as the reference uploads it
(scene_buffers.h:14-38): vertices as cl_float3 (16 B), `triangle{surface,v0,v1,v2}`
(16 B), `surface{absorption[8], scattering[8]}` (64 B), the voxel grid's AABB and
side, and the flattened voxel index of voxel_collection.cpp:9-37:
`index[x*side*side + y*side + z]` = offset of a run `[count, tri, tri, ...]`.
Building that structure is the reference's host code (octree, out of scope);
this module only fabricates inputs of the same shape for tests and bench.py.
Triangles are assigned to every voxel their bounding box touches (a superset
of the exact overlap set; the traversal accepts a hit only inside the voxel it
is visiting, voxel.cpp:80-90, so supersets do not change results).
"""
from __future__ import annotations

import numpy as np

TRI_DT = np.dtype([("surface", "<u4"), ("v0", "<u4"), ("v1", "<u4"), ("v2", "<u4")])
SURF_DT = np.dtype([("absorption", "<f4", (8,)), ("scattering", "<f4", (8,))])
REFL_DT = np.dtype([("position", "<f4", (4,)), ("triangle", "<u4"), ("keep_going", "i1"),
                    ("receiver_visible", "i1"), ("pad", "i1", (10,))])
assert TRI_DT.itemsize == 16 and SURF_DT.itemsize == 64 and REFL_DT.itemsize == 32


class Scene:
    def __init__(self, vertices, triangles, surfaces, side=16, pad=0.1):
        v = np.asarray(vertices, np.float32).reshape(-1, 3)
        self.vertices = np.zeros((v.shape[0], 4), np.float32)  # cl_float3 = 16 bytes
        self.vertices[:, :3] = v
        self.triangles = np.ascontiguousarray(triangles, TRI_DT).reshape(-1)
        self.surfaces = np.ascontiguousarray(surfaces, SURF_DT).reshape(-1)
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
