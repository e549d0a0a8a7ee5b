"""ctypes front-end of the CPU scene-preparation oracle (oracle/scene_oracle.cpp).
TEST INFRASTRUCTURE ONLY (see that file's header)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libsceneoracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "scene_oracle.cpp")
        if not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
            subprocess.run(["make", "-C", _HERE, "_build/libsceneoracle.so"], check=True, stdout=subprocess.DEVNULL)
        L = C.CDLL(_LIB_PATH)
        vp, sz = C.c_void_p, C.c_size_t
        L.sco_voxelise.restype = sz
        L.sco_voxelise.argtypes = [vp, sz, vp, sz, sz, C.c_float, vp, vp, sz]
        L.sco_overlaps.restype = C.c_int
        L.sco_overlaps.argtypes = [vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def voxelise(vertices4, triangles, depth=5, padding=0.1):
    """make_voxelised_scene_data(scene, depth, padding) + get_flattened -> (aabb[6], index)."""
    v = np.ascontiguousarray(vertices4, np.float32).reshape(-1, 4)
    t = np.ascontiguousarray(triangles).view(np.uint32).reshape(-1, 4)
    aabb = np.zeros(6, np.float32)
    n = lib().sco_voxelise(_p(v), v.shape[0], _p(t), t.shape[0], int(depth), float(padding), _p(aabb), None, 0)
    out = np.zeros(n, np.uint32)
    lib().sco_voxelise(_p(v), v.shape[0], _p(t), t.shape[0], int(depth), float(padding), _p(aabb), _p(out), n)
    return aabb, out


def overlaps(box6, tri9) -> bool:
    b = np.ascontiguousarray(box6, np.float32).reshape(6)
    t = np.ascontiguousarray(tri9, np.float32).reshape(9)
    return bool(lib().sco_overlaps(_p(b), _p(t)))
