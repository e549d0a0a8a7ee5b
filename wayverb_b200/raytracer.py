"""Host mirror of the reference's ray-tracing interface over the C ABI.

  RayTracer(scene)            core::scene_buffers + the per-run state of raytracer::run
                              (raytracer.h:188-266)
  RayTracer.trace(...)        the segment x depth loop with the stochastic histogram
                              processor (reflection_processor/stochastic_histogram.h) folded in;
                              reflections of the first `keep_steps` steps are what the
                              image-source / visual processors consume
  reflection_depth(...)       compute_optimum_reflection_number (optimum_reflection_number.h:38-40)
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import IMPULSE_DT, IsDesc, RtSceneDesc, RtTraceParams, check, lib, ptr
from .scene import REFL_DT, Scene


def reflection_depth(min_absorption: float) -> int:
    return int(lib().wvb_rt_reflection_depth(float(min_absorption)))


def ray_energy(total_rays, source, receiver, radius) -> float:
    s = np.asarray(source, np.float32)
    r = np.asarray(receiver, np.float32)
    return float(lib().wvb_rt_ray_energy(int(total_rays), ptr(s), ptr(r), float(radius)))


class RayTracer:
    def __init__(self, scene: Scene, device=0):
        self.scene = scene
        d = RtSceneDesc()
        d.voxel_index = scene.voxel_index.ctypes.data
        d.voxel_index_count = scene.voxel_index.size
        d.aabb_min[:] = [float(v) for v in scene.aabb[:3]]
        d.aabb_max[:] = [float(v) for v in scene.aabb[3:]]
        d.side = scene.side
        d.triangles = scene.triangles.ctypes.data
        d.num_triangles = scene.triangles.size
        d.vertices = scene.vertices.ctypes.data
        d.num_vertices = scene.vertices.shape[0]
        d.surfaces = scene.surfaces.ctypes.data
        d.num_surfaces = scene.surfaces.size
        d.device = int(device)
        h = C.c_void_p()
        check(lib().wvb_rt_create(C.byref(d), C.byref(h)))
        self._h = h
        self._shape = None

    def close(self):
        if getattr(self, "_h", None):
            lib().wvb_rt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def safe_bins(self, depth, speed_of_sound=340.0, rate=1000.0) -> int:
        return int(lib().wvb_rt_safe_bins(self._h, int(depth), float(speed_of_sound), float(rate)))

    def trace(self, dirs, source, receiver, depth, n_rays=None, total_rays=None, receiver_radius=0.1,
              speed_of_sound=340.0, histogram_rate=1000.0, seed=1, ray_index_base=0, specular_from_step=0,
              n_bins=None, directional=False, keep_steps=0, mode=0):
        """dirs: n x 3 float32 or None (generate n_rays directions on the device).
        mode: _lib.RT_MODE_* (0 = by batch size).
        Returns (reflections or None, dropped, device_ms); the histogram accumulates on the
        device, read it with histogram()."""
        d, n, P = self._params(dirs, source, receiver, depth, n_rays, total_rays, receiver_radius, speed_of_sound,
                               histogram_rate, seed, ray_index_base, specular_from_step, n_bins, directional,
                               keep_steps)
        P.mode = int(mode)
        refl = np.zeros((keep_steps, n), REFL_DT) if keep_steps else None
        dropped, ms = C.c_uint64(0), C.c_float(0)
        check(lib().wvb_rt_trace(self._h, C.byref(P), ptr(d) if d is not None else None, n,
                                 ptr(refl) if refl is not None else None, C.byref(dropped), C.byref(ms)))
        return refl, dropped.value, ms.value

    def _params(self, dirs, source, receiver, depth, n_rays, total_rays, receiver_radius, speed_of_sound,
                histogram_rate, seed, ray_index_base, specular_from_step, n_bins, directional, keep_steps):
        if dirs is not None:
            d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
            n = d.shape[0]
        else:
            d, n = None, int(n_rays)
        P = RtTraceParams()
        P.source[:] = [float(v) for v in source]
        P.receiver[:] = [float(v) for v in receiver]
        P.receiver_radius = receiver_radius
        P.speed_of_sound, P.histogram_sample_rate = speed_of_sound, histogram_rate
        P.total_rays = n if total_rays is None else int(total_rays)
        P.seed, P.ray_index_base = int(seed), int(ray_index_base)
        P.depth, P.specular_from_step = int(depth), int(specular_from_step)
        P.n_bins = self.safe_bins(depth, speed_of_sound, histogram_rate) if n_bins is None else int(n_bins)
        P.directional, P.keep_steps = int(bool(directional)), int(keep_steps)
        self._shape = (20, 9, P.n_bins, 8) if directional else (P.n_bins, 8)
        return d, n, P

    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        """attach an NCCL communicator (unique_id: wayverb_b200.waveguide.nccl_unique_id() of rank 0)"""
        buf = np.frombuffer(unique_id, np.uint8).copy()
        check(lib().wvb_rt_comm_init(self._h, ptr(buf), int(rank), int(nranks)))

    def allreduce_histogram(self):
        check(lib().wvb_rt_allreduce_histogram(self._h))

    def histogram(self) -> np.ndarray:
        out = np.zeros(self._shape)
        check(lib().wvb_rt_read_histogram(self._h, ptr(out)))
        return out

    def reset_histogram(self):
        check(lib().wvb_rt_reset_histogram(self._h))

    def closest_hit(self, pos, dirs):
        rays = np.ascontiguousarray(np.concatenate([pos, dirs], 1), np.float32)
        n = rays.shape[0]
        tri, t = np.zeros(n, np.uint32), np.zeros(n, np.float32)
        check(lib().wvb_rt_closest_hit(self._h, ptr(rays), n, ptr(tri), ptr(t)))
        return tri, t

    def directions(self, seed, n, base=0):
        out = np.zeros((n, 3), np.float32)
        check(lib().wvb_rt_directions(self._h, int(seed), int(base), int(n), ptr(out)))
        return out


class ImageSource:
    """reflection_processor::make_image_source on the device
    (reflection_processor/image_source.h:15-86): push the first reflections of every
    ray (from the host, or straight from a trace), then results() = get_results()."""

    def __init__(self, tracer: RayTracer, source, receiver, max_elements, acoustic_impedance=400.0,
                 flip_phase=False, with_direct=True):
        self.tracer = tracer
        self.source, self.receiver = tuple(map(float, source)), tuple(map(float, receiver))
        d = IsDesc()
        d.source[:] = self.source
        d.receiver[:] = self.receiver
        d.acoustic_impedance = float(acoustic_impedance)
        d.flip_phase, d.with_direct = int(bool(flip_phase)), int(bool(with_direct))
        d.max_elements = int(max_elements)
        h = C.c_void_p()
        check(lib().wvb_is_create(tracer._h, C.byref(d), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().wvb_is_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def push_elements(self, elems, ray_index_base=0):
        e = np.ascontiguousarray(elems, np.uint32)
        order, n = e.shape
        check(lib().wvb_is_push_elements(self._h, ptr(e), n, order, int(ray_index_base)))

    def push_reflections(self, refl, ray_index_base=0):
        r = np.ascontiguousarray(refl, REFL_DT)
        steps, n = r.shape
        check(lib().wvb_is_push_reflections(self._h, ptr(r), n, steps, int(ray_index_base)))

    def trace(self, dirs, depth, order, n_rays=None, total_rays=None, receiver_radius=0.1, speed_of_sound=340.0,
              histogram_rate=1000.0, seed=1, ray_index_base=0, specular_from_step=0, n_bins=None,
              directional=False, keep_steps=0, mode=0):
        """traces and feeds the tree on the device; returns the reflections of the first
        keep_steps steps (or None) for host-side consumers"""
        d, n, P = self.tracer._params(dirs, self.source, self.receiver, depth, n_rays, total_rays, receiver_radius,
                                      speed_of_sound, histogram_rate, seed, ray_index_base, specular_from_step,
                                      n_bins, directional, keep_steps)
        P.mode = int(mode)
        refl = np.zeros((keep_steps, n), REFL_DT) if keep_steps else None
        dropped, ms = C.c_uint64(0), C.c_float(0)
        check(lib().wvb_is_trace(self._h, C.byref(P), ptr(d) if d is not None else None, n, int(order),
                                 ptr(refl) if refl is not None else None, C.byref(dropped), C.byref(ms)))
        return refl

    def results(self):
        """-> (impulses, stats[4], validation kernel ms)"""
        count, ms = C.c_uint64(0), C.c_float(0)
        stats = (C.c_uint64 * 4)()
        check(lib().wvb_is_results(self._h, None, 0, C.byref(count), C.byref(stats), C.byref(ms)))
        out = np.zeros(count.value, IMPULSE_DT)
        check(lib().wvb_is_results(self._h, ptr(out), out.size, C.byref(count), C.byref(stats), C.byref(ms)))
        return out[:count.value], np.array(list(stats), np.uint64), ms.value
