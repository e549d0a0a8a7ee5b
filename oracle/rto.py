"""ctypes front-end of the CPU ray-tracing oracle (oracle/rt_oracle.cpp).
TEST INFRASTRUCTURE ONLY (see wg_oracle.cpp's header)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "librtoracle.so")

REFL_DT = np.dtype([("position", "<f4", (4,)), ("triangle", "<u4"), ("keep_going", "i1"),
                    ("receiver_visible", "i1"), ("pad", "i1", (10,))])


class TraceParams(C.Structure):
    _fields_ = [
        ("source", C.c_float * 3), ("receiver", C.c_float * 3),
        ("receiver_radius", C.c_float), ("pad0", C.c_float),
        ("speed_of_sound", C.c_double), ("histogram_rate", C.c_double),
        ("total_rays", C.c_uint64), ("seed", C.c_uint64), ("ray_index_base", C.c_uint64),
        ("depth", C.c_uint32), ("specular_from_step", C.c_uint32), ("n_bins", C.c_uint32),
        ("directional", C.c_uint32), ("keep_steps", C.c_uint32), ("pad1", C.c_uint32),  # pad1: the product's scheduling mode, unused here
    ]


_lib = None


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("rt_oracle.cpp", "is_oracle.inc")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(map(os.path.getmtime, srcs)):
        subprocess.run(["make", "-C", _HERE, "_build/librtoracle.so"], check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, sz = C.c_void_p, C.c_size_t
        L.rto_scene_create.restype = vp
        L.rto_scene_create.argtypes = [vp, sz, vp, C.c_uint32, vp, sz, vp, sz, vp, sz]
        L.rto_scene_destroy.argtypes = [vp]
        L.rto_closest_hit.argtypes = [vp, vp, sz, C.c_int, vp, vp]
        L.rto_ray_energy.restype = C.c_float
        L.rto_ray_energy.argtypes = [C.c_uint64, vp, vp, C.c_float]
        L.rto_directions.argtypes = [C.c_uint64, C.c_uint64, sz, vp]
        L.rto_sincos.argtypes = [vp, sz, vp, vp]
        L.rto_step_rng.argtypes = [C.c_uint64, C.c_uint64, sz, C.c_uint32, vp]
        L.rto_lut_index.argtypes = [vp, sz, vp, vp]
        L.rto_nodes_inside.argtypes = [vp, vp, vp, C.c_float, vp]
        L.rto_closest_surface.argtypes = [vp, vp, sz, vp, vp]
        L.rto_trace.argtypes = [vp, vp, vp, sz, vp, vp, vp]
        L.rto_image_source.restype = sz
        L.rto_image_source.argtypes = [vp, vp, vp, vp, sz, sz, C.c_double, C.c_int, C.c_int, vp, sz, vp]
        L.rto_exact_shoebox.restype = sz
        L.rto_exact_shoebox.argtypes = [vp, vp, vp, vp, C.c_float, C.c_double, C.c_double, vp, sz]
        L.rto_num_threads.restype = C.c_int
        L.rto_params_size.restype = sz
        assert L.rto_params_size() == C.sizeof(TraceParams)
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Scene:
    def __init__(self, sc):
        """sc: a wayverb_b200.scene.Scene (or anything with the same arrays)."""
        self.sc = sc
        self._h = lib().rto_scene_create(_p(sc.voxel_index), sc.voxel_index.size, _p(sc.aabb), sc.side,
                                         _p(sc.triangles), sc.triangles.size, _p(sc.vertices),
                                         sc.vertices.shape[0], _p(sc.surfaces), sc.surfaces.size)

    def __del__(self):
        try:
            if self._h:
                lib().rto_scene_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def closest_hit(self, pos, dirs, brute=False):
        rays = np.ascontiguousarray(np.concatenate([pos, dirs], 1), np.float32)
        n = rays.shape[0]
        tri = np.zeros(n, np.uint32)
        t = np.zeros(n, np.float32)
        lib().rto_closest_hit(self._h, _p(rays), n, int(brute), _p(tri), _p(t))
        return tri, t

    def nodes_inside(self, min_corner, dims, spacing):
        """set_node_inside for a mesh descriptor -> bool array [z, y, x]"""
        mc = np.asarray(min_corner, np.float32)
        d = np.asarray(dims, np.int32)
        out = np.zeros(int(d[0]) * int(d[1]) * int(d[2]), np.uint8)
        lib().rto_nodes_inside(self._h, _p(mc), _p(d), float(spacing), _p(out))
        return out.reshape(int(d[2]), int(d[1]), int(d[0])).astype(bool)

    def closest_surface(self, points):
        """the 1d finder: surface (and triangle) of the closest triangle per point"""
        p = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        surf = np.zeros(p.shape[0], np.uint32)
        tri = np.zeros(p.shape[0], np.uint32)
        lib().rto_closest_surface(self._h, _p(p), p.shape[0], _p(surf), _p(tri))
        return surf, tri

    def trace(self, dirs, source, receiver, depth, total_rays=None, receiver_radius=0.1, speed_of_sound=340.0,
              histogram_rate=1000.0, seed=1, ray_index_base=0, specular_from_step=0, n_bins=None,
              directional=False, keep_steps=0):
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = d.shape[0]
        P = TraceParams()
        P.source[:] = [float(v) for v in source]
        P.receiver[:] = [float(v) for v in receiver]
        P.receiver_radius = receiver_radius
        P.speed_of_sound, P.histogram_rate = speed_of_sound, histogram_rate
        P.total_rays = n if total_rays is None else total_rays
        P.seed, P.ray_index_base = seed, ray_index_base
        P.depth, P.specular_from_step = depth, specular_from_step
        if n_bins is None:
            n_bins = int(np.ceil((depth + 1) * self.sc.diagonal / speed_of_sound * histogram_rate)) + 1
        P.n_bins, P.directional, P.keep_steps = n_bins, int(directional), keep_steps
        shape = (20, 9, n_bins, 8) if directional else (n_bins, 8)
        hist = np.zeros(shape)
        dropped = C.c_uint64(0)
        refl = np.zeros((keep_steps, n), REFL_DT) if keep_steps else None
        lib().rto_trace(self._h, C.byref(P), _p(d), n, _p(hist), C.byref(dropped),
                        _p(refl) if refl is not None else None)
        return hist, refl, dropped.value


IMPULSE_DT = np.dtype([("volume", np.float32, 8), ("position", np.float32, 4), ("distance", np.float32),
                       ("pad_", np.float32, 3)])  # raytracer::impulse<8>, 64 B
IS_NONE = 0xFFFFFFFF
IS_VISIBLE = 0x80000000


def path_elements(refl, order):
    """reflection records [steps][n] -> image-source path elements [order][n]
    (reflection_path_builder.h:16-26, group processor: step < max_order)."""
    k = min(order, refl.shape[0])
    tri = refl["triangle"][:k].astype(np.uint32)
    e = np.where(refl["keep_going"][:k] != 0,
                 tri | np.where(refl["receiver_visible"][:k] != 0, np.uint32(IS_VISIBLE), np.uint32(0)),
                 np.uint32(IS_NONE)).astype(np.uint32)
    return np.ascontiguousarray(e)


def image_source(scene, elems, source, receiver, acoustic_impedance=400.0, flip_phase=False, with_direct=True):
    """the reference's image-source stage on path elements [order][n_rays] -> (impulses, stats)"""
    e = np.ascontiguousarray(elems, np.uint32)
    order, n = e.shape
    s = np.asarray(source, np.float32)
    r = np.asarray(receiver, np.float32)
    stats = np.zeros(3, np.uint64)
    cap = 1 << 16
    while True:
        out = np.zeros(cap, IMPULSE_DT)
        cnt = lib().rto_image_source(scene._h, _p(s), _p(r), _p(e), n, order, float(acoustic_impedance),
                                     int(flip_phase), int(with_direct), _p(out), cap, _p(stats))
        if cnt <= cap:
            return out[:cnt], stats
        cap = cnt


def exact_shoebox(box_min, box_max, source, receiver, absorption, max_distance, acoustic_impedance=400.0):
    a = [np.asarray(v, np.float32) for v in (box_min, box_max, source, receiver)]
    cap = 1 << 16
    out = np.zeros(cap, IMPULSE_DT)
    n = lib().rto_exact_shoebox(_p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), float(absorption), float(max_distance),
                                float(acoustic_impedance), _p(out), cap)
    assert n <= cap
    return out[:n]


def ray_energy(total_rays, source, receiver, radius):
    s = np.asarray(source, np.float32)
    r = np.asarray(receiver, np.float32)
    return lib().rto_ray_energy(int(total_rays), _p(s), _p(r), float(radius))


def directions(seed, n, base=0):
    out = np.zeros((n, 3), np.float32)
    lib().rto_directions(int(seed), int(base), n, _p(out))
    return out


def step_rng(seed, n, step, base=0):
    """(z, theta) per ray for one reflection step -- the stream trace() consumes"""
    out = np.zeros((n, 2), np.float32)
    lib().rto_step_rng(int(seed), int(base), n, int(step), _p(out))
    return out


def sincos(theta):
    t = np.ascontiguousarray(theta, np.float32)
    s, c = np.zeros_like(t), np.zeros_like(t)
    lib().rto_sincos(_p(t), t.size, _p(s), _p(c))
    return s, c


def lut_index(v):
    v = np.ascontiguousarray(v, np.float32).reshape(-1, 3)
    az, el = np.zeros(v.shape[0], np.int32), np.zeros(v.shape[0], np.int32)
    lib().rto_lut_index(_p(v), v.shape[0], _p(az), _p(el))
    return az, el


def num_threads():
    return lib().rto_num_threads()
