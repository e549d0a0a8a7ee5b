"""ctypes front-end of oracle/_ref/lib_ref.so: the REFERENCE'S OWN OpenCL-C kernels, extracted
from /root/reference at build time and compiled for the host by oracle/ref_recipe/build.py.

TEST INFRASTRUCTURE ONLY (tests/ and bench.py's reference arm). It exists to pin the oracle
(oracle/wg_oracle.cpp, oracle/rt_oracle.cpp) to reference-run output; the product never loads it.

The reference's HOST code is in the same library, compiled where it lies behind stand-ins for the
dependencies this image lacks (ref_recipe/hoststubs: GLM, FFTW, IT++, libsamplerate; ref_recipe/hostcl: a
host-memory cl.hpp): the octree voxeliser, the image-source stage, histogram binning, post-processing,
the filter-design pipeline, the mesh-descriptor maths, and the waveguide::run template itself with its
stock processors (run_waveguide below, enqueueing the kernel above). Sim / RayScene.trace_steps drive the
kernels step by step from Python in the reference's order for the tests that need to look between steps
(waveguide.h:36-126, reflector.cpp:31-51, stochastic/finder.h:48-79), each step citing its line.
"""
from __future__ import annotations

import ctypes as C
import importlib.util
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("_wvb_ref_recipe", os.path.join(_HERE, "ref_recipe", "build.py"))
recipe = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(recipe)

NODE_DT = np.dtype([("boundary_type", "<i4"), ("boundary_index", "<u4")])
COEFF_DT = np.dtype([("b", "<f8", (7,)), ("a", "<f8", (7,))])
BDATA_DT = np.dtype([("mem", "<f8", (6,)), ("coefficient_index", "<u4"), ("pad", "<u4")])
REFL_DT = np.dtype([("position", "<f4", (4,)), ("triangle", "<u4"), ("keep_going", "i1"),
                    ("receiver_visible", "i1"), ("pad", "i1", (10,))])
RAY_DT = np.dtype([("position", "<f4", (4,)), ("direction", "<f4", (4,))])
IMPULSE_DT = np.dtype([("volume", "<f4", (8,)), ("position", "<f4", (4,)), ("distance", "<f4"),
                       ("pad", "<f4", (3,))])

_lib = None
# (order, points, f[points], m[points], b_out[order + 1], a_out[order + 1]) -- hoststubs/itpp/signal/filter_design.h
YULEWALK_CB = C.CFUNCTYPE(None, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                          C.POINTER(C.c_double), C.POINTER(C.c_double))


def available() -> bool:
    return recipe.build() is not None


_native_wg = None


def use_native_wg() -> bool:
    """bench.py's reference arm: recompile the generated waveguide unit (oracle/_ref/ref_wg_f32.cpp,
    the reference's kernel source behind the prelude -- it travels with oracle/_ref) with
    -O3 -march=native ON this machine and let Sim(real="float") step through it. Same source, same
    -ffp-contract=off. Returns False (portable lib_ref.so stays in use) if that is not possible."""
    global _native_wg
    if _native_wg is not None:
        return bool(_native_wg)
    import subprocess
    src = os.path.join(recipe.OUT, "ref_wg_f32.cpp")
    out_dir = os.path.join(recipe.OUT, "native")
    out = os.path.join(out_dir, "lib_ref_wg_f32.so")
    _native_wg = False
    if os.path.exists(src):
        try:
            os.makedirs(out_dir, exist_ok=True)
            flags = [f for f in recipe.FLAGS if f != "-O2"] + ["-O3", "-march=native"]
            subprocess.run([recipe.CXX] + flags + ["-shared", "-o", out, src], check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            L = C.CDLL(out)
            L.refk_f32_wg_step.restype = C.c_int
            L.refk_f32_wg_step.argtypes = [C.c_void_p] * 3 + [C.c_int] * 3 + [C.c_void_p] * 4 + [C.c_size_t]
            _native_wg = L
        except (subprocess.CalledProcessError, OSError):
            _native_wg = False
    return bool(_native_wg)


def lib():
    global _lib
    if _lib is None:
        path = recipe.build()
        if path is None:
            raise RuntimeError("oracle/_ref/lib_ref.so is absent and /root/reference is not here to build it")
        L = C.CDLL(path)
        vp, sz, i, u, f = C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_float
        for p in ("f32_", "f64_"):
            g = lambda n: getattr(L, "refk_" + p + n)  # noqa: E731
            g("wg_step").restype = i
            g("wg_step").argtypes = [vp, vp, vp, i, i, i, vp, vp, vp, vp, sz]
            g("filter_test").argtypes = [vp, vp, vp, vp, sz]
            g("filter_test_2").argtypes = [vp, vp, vp, vp, sz]
            g("to_locator").argtypes = [sz, i, i, i, vp]
            g("neighbor_index").restype = u
            g("neighbor_index").argtypes = [i, i, i, i, i, i, i]
            for n in ("sizeof_real", "sizeof_boundary_data", "sizeof_coefficients", "sizeof_node"):
                g(n).restype = sz
        assert L.refk_f32_sizeof_real() == 4 and L.refk_f64_sizeof_real() == 8
        assert L.refk_f32_sizeof_boundary_data() == BDATA_DT.itemsize == 56
        assert L.refk_f32_sizeof_coefficients() == COEFF_DT.itemsize == 112
        assert L.refk_f32_sizeof_node() == NODE_DT.itemsize == 8
        L.refk_init_reflections.argtypes = [vp, sz]
        L.refk_reflections.argtypes = [vp, vp, vp, vp, u, vp, vp, vp, vp, vp, sz]
        L.refk_triangle_vert_intersection.argtypes = [vp, vp, vp]
        L.refk_closest_hit.argtypes = [vp, sz, i, vp, vp, u, vp, u, vp, vp, vp]
        L.refk_sphere_point.argtypes = [f, f, vp]
        L.refk_init_stochastic_path_info.argtypes = [vp, f, vp, sz]
        L.refk_stochastic.argtypes = [vp, vp, f, vp, vp, vp, vp, vp, vp, sz]
        L.refk_set_node_inside.argtypes = [vp, vp, vp, f, vp, vp, u, vp, vp, sz]
        L.refk_set_node_boundary_type.argtypes = [vp, vp, vp, f, sz]
        L.refk_boundary_coefficient_finder_1d.argtypes = [vp, vp, vp, f, vp, vp, vp, u, vp, u, vp, sz, i]
        L.refk_boundary_coefficient_finder_2d.argtypes = [vp, vp, vp, f, vp, vp, sz]
        L.refk_boundary_coefficient_finder_3d.argtypes = [vp, vp, vp, f, vp, vp, sz]
        L.refk_voxelise.restype = sz
        L.refk_voxelise.argtypes = [vp, sz, vp, sz, sz, f, vp, vp, sz]
        L.refk_overlaps.restype = i
        L.refk_overlaps.argtypes = [vp, vp]
        d = C.c_double
        L.refk_is_intersects.argtypes = [vp, sz, vp, sz, sz, f, vp, vp, vp, sz, vp, vp]
        L.refk_is_mirror.argtypes = [vp, vp, vp]
        L.refk_is_image_source.restype = sz
        L.refk_is_image_source.argtypes = [vp, sz, vp, sz, vp, sz, sz, f, vp, vp, vp, sz, sz, sz, d, i, i, vp, sz]
        L.refk_run_waveguide.restype = i
        L.refk_run_waveguide.argtypes = [vp, vp, f, vp, sz, vp, sz, vp, sz, vp, sz, vp, sz, i, sz, vp, sz, vp, vp, sz, vp,
                                         sz, d, d, vp, vp, C.c_char_p, sz]
        L.refk_ray_direction_rng.argtypes = [u, sz, vp]
        L.refk_ray_run.restype = i
        L.refk_ray_run.argtypes = [vp, sz, vp, sz, vp, sz, sz, f, vp, vp, d, d, vp, sz, u, sz, sz, sz, f, f, i, sz, vp, vp,
                                   C.c_char_p, sz]
        L.refk_ray_read.argtypes = [vp, vp, vp]
        L.refk_compute_mesh.restype = i
        L.refk_compute_mesh.argtypes = [vp, sz, vp, sz, vp, sz, sz, f, vp, d, f, f, YULEWALK_CB, vp, vp, vp, vp,
                                        C.c_char_p, sz]
        L.refk_mesh_read.argtypes = [vp, vp, vp, vp, vp, vp]
        L.refk_canonical.restype = i
        L.refk_canonical.argtypes = [vp, vp, f, vp, sz, vp, sz, vp, sz, vp, sz, vp, sz, vp, sz, vp, vp, d, d, sz, d, d, d,
                                     vp, sz, vp, vp, vp, C.c_char_p, sz]
        L.refk_hm_to_impedance.argtypes = [vp, vp, vp, vp]
        L.refk_hm_to_flat.argtypes = [d, vp, vp]
        L.refk_hm_is_stable.restype = i
        L.refk_hm_is_stable.argtypes = [vp]
        L.refk_hm_peak_biquad.argtypes = [d, d, d, vp]
        L.refk_hm_convolve3.argtypes = [vp, vp, vp]
        L.refk_hm_reflectance_filter.restype = i
        L.refk_hm_reflectance_filter.argtypes = [vp, d, YULEWALK_CB, vp, vp, vp]
        L.refk_hm_reflection_number.restype = sz
        L.refk_hm_reflection_number.argtypes = [d]
        L.refk_hm_ray_energy.restype = f
        L.refk_hm_ray_energy.argtypes = [sz, vp, vp, f]
        L.refk_hm_rates.argtypes = [f, d, vp]
        L.refk_hm_calibration_factor.restype = d
        L.refk_hm_calibration_factor.argtypes = [f, d]
        L.refk_lut_index.argtypes = [vp, sz, vp, vp]
        L.refk_histogram.restype = sz
        L.refk_histogram.argtypes = [vp, vp, vp, sz, vp, d, d, i, vp, sz]
        L.refk_is_exact_shoebox.restype = sz
        L.refk_is_exact_shoebox.argtypes = [vp, vp, vp, vp, f, d, d, vp, sz]
        L.refk_pp_rate_law.argtypes = [d, d, d, vp]
        L.refk_pp_intervals.argtypes = [u, sz, vp, vp]
        L.refk_pp_dirac_sequence.restype = sz
        L.refk_pp_dirac_sequence.argtypes = [u, d, d, d, d, vp, sz]
        for n in ("weight_sequence", "postprocessing"):
            getattr(L, "refk_pp_" + n).restype = sz
            getattr(L, "refk_pp_" + n).argtypes = [vp, sz, d, vp, sz, d, d, vp, sz]
        L.refk_pp_multiband_mixdown.argtypes = [vp, sz, d, vp]
        L.refk_pp_band_params.argtypes = [d, vp, vp]
        L.refk_pp_magnitudes.argtypes = [d, d, d, d, sz, vp]
        L.refk_pp_crossover.restype = sz
        L.refk_pp_crossover.argtypes = [vp, sz, vp, sz, d, d, vp]
        L.refk_pp_left_hanning.argtypes = [sz, vp]
        L.refk_pp_fft_length.restype = sz
        L.refk_pp_fft_length.argtypes = [sz]
        for n, want in (("reflection", 32), ("ray", 32), ("surface", 64), ("triangle", 16), ("impulse", 64),
                        ("path_info", 64), ("mesh_descriptor", 48)):
            fn = getattr(L, "refk_sizeof_" + n)
            fn.restype = sz
            assert fn() == want, (n, fn())
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ---- waveguide --------------------------------------------------------------------------
class Sim:
    """waveguide::run's state and step (waveguide.h:43-124) around the reference's
    condensed_waveguide kernel. real="float": the reference's own types (float pressures,
    double filters); real="double": the same source built with float -> double."""

    def __init__(self, mesh, real="float"):
        self.mesh = mesh
        self.p = {"float": "f32_", "double": "f64_"}[real]
        self.dtype = np.float32 if real == "float" else np.float64
        n = mesh.num_nodes
        # waveguide.h:47-56 two zeroed pressure buffers
        self.prev = np.zeros(n, self.dtype)
        self.cur = np.zeros(n, self.dtype)
        self.nodes = np.ascontiguousarray(mesh.nodes, NODE_DT)
        self.coeffs = np.ascontiguousarray(mesh.coeffs, COEFF_DT)
        # setup.h:68-85 get_boundary_data<n>: zero filter memory, coefficient_index copied
        self.bd = []
        for b in (mesh.b1, mesh.b2, mesh.b3):
            a = np.zeros((max(b.shape[0], 1), b.shape[1]), BDATA_DT)
            if b.shape[0]:
                a["coefficient_index"][:b.shape[0]] = b
            self.bd.append(a)
        self._n = [mesh.b1.shape[0], mesh.b2.shape[0], mesh.b3.shape[0]]

    def write(self, node, v):
        self.cur[int(node)] = v

    def read(self, node):
        return float(self.cur[int(node)])

    def field(self):
        return self.cur.astype(np.float64)

    def set_field(self, f):
        self.cur[:] = np.asarray(f, np.float64).reshape(-1)

    def step(self, n=1) -> int:
        dx, dy, dz = self.mesh.dims
        fn = getattr(_native_wg if (_native_wg and self.p == "f32_") else lib(), "refk_" + self.p + "wg_step")
        flags = 0
        for _ in range(int(n)):
            # waveguide.h:82-97 flag reset + kernel, :123 swap
            flags |= fn(_p(self.prev), _p(self.cur), _p(self.nodes), dx, dy, dz,
                        _p(self.bd[0]), _p(self.bd[1]), _p(self.bd[2]), _p(self.coeffs), self.mesh.num_nodes)
            self.prev, self.cur = self.cur, self.prev
        return flags

    def run(self, src_node, signal, rcv_nodes, soft=False):
        """waveguide::run with hard_source / soft_source (preprocessor/*.h:17-25) and
        postprocessor::node (node.cpp:14-18). Stops at the first error flag like
        waveguide.h:100-119 throws."""
        out = np.zeros((len(signal), len(rcv_nodes)))
        for s, x in enumerate(signal):
            if soft:
                self.cur[src_node] = self.dtype(self.cur[src_node] + self.dtype(x))
            else:
                self.cur[src_node] = x
            # the reference enqueues the kernel, then `post` reads `current` (waveguide.h:85-121)
            for k, r in enumerate(rcv_nodes):
                out[s, k] = self.cur[int(r)]
            flag = self.step(1)
            if flag:
                return s, out, flag
        return len(signal), out, 0

    def boundary_data(self, n):
        return self.bd[n - 1][:self._n[n - 1]]


def filter_biquads(biquads, x_f32, real="float"):
    """filter_test (cl/filters.cpp:56-65) on one stream, sample by sample."""
    p = {"float": "f32_", "double": "f64_"}[real]
    dt = np.float32 if real == "float" else np.float64
    c = np.ascontiguousarray(biquads, np.float64).reshape(3, 6).copy()  # 3 x {b[3], a[3]}
    m = np.zeros(6)
    x = np.ascontiguousarray(x_f32, dt)
    y = np.zeros_like(x)
    fn = getattr(lib(), "refk_" + p + "filter_test")
    for i in range(x.size):
        fn(_p(x[i:i + 1]), _p(y[i:i + 1]), _p(m), _p(c), 1)
    return y


def filter_canonical(coeffs, x, real="float"):
    """filter_test_2 (cl/filters.cpp:67-75) on one stream."""
    p = {"float": "f32_", "double": "f64_"}[real]
    dt = np.float32 if real == "float" else np.float64
    c = np.ascontiguousarray(coeffs, COEFF_DT).reshape(1).copy()
    m = np.zeros(6)
    xx = np.ascontiguousarray(x, dt)
    y = np.zeros_like(xx)
    fn = getattr(lib(), "refk_" + p + "filter_test_2")
    for i in range(xx.size):
        fn(_p(xx[i:i + 1]), _p(y[i:i + 1]), _p(m), _p(c), 1)
    return y


def to_locator(index, dims):
    out = np.zeros(3, np.int32)
    lib().refk_f32_to_locator(int(index), int(dims[0]), int(dims[1]), int(dims[2]), _p(out))
    return tuple(int(v) for v in out)


def neighbor_index(loc, dims, port):
    return int(lib().refk_f32_neighbor_index(int(loc[0]), int(loc[1]), int(loc[2]),
                                             int(dims[0]), int(dims[1]), int(dims[2]), int(port)))


# ---- rays -------------------------------------------------------------------------------------
class RayScene:
    """Holds the flattened scene arrays (core/spatial_division/scene_buffers.h:14-38)."""

    def __init__(self, sc):
        self.voxel_index = np.ascontiguousarray(sc.voxel_index, np.uint32)
        self.aabb = np.ascontiguousarray(sc.aabb, np.float32).reshape(6)
        self.side = int(sc.side)
        self.triangles = np.ascontiguousarray(sc.triangles)
        v = np.asarray(sc.vertices, np.float32)
        self.vertices = np.zeros((v.shape[0], 4), np.float32)
        self.vertices[:, :v.shape[1]] = v
        self.surfaces = np.ascontiguousarray(sc.surfaces)
        assert self.triangles.dtype.itemsize == 16 and self.surfaces.dtype.itemsize == 64

    def closest_hit(self, pos, dirs, brute=False):
        rays = np.ascontiguousarray(np.concatenate([pos, dirs], 1), np.float32)
        n = rays.shape[0]
        tri = np.zeros(n, np.uint32)
        t = np.zeros(n, np.float32)
        lib().refk_closest_hit(_p(rays), n, int(brute), _p(self.voxel_index), _p(self.aabb), self.side,
                               _p(self.triangles), self.triangles.size, _p(self.vertices), _p(tri), _p(t))
        return tri, t

    def trace_steps(self, dirs, source, receiver, depth, rng_for_step, receiver_radius=0.1, initial_energy=1.0):
        """raytracer::run's inner loop for one segment (raytracer.h:223-240) with the
        stochastic finder (stochastic/finder.h:48-79): yields, per step, the reflection
        records and the two impulse arrays exactly as the kernels wrote them.
        rng_for_step(step) -> float32 [n, 2] (z, theta): the reference draws these on the host
        (reflector.cpp:13-25); the caller supplies the stream."""
        L = lib()
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = d.shape[0]
        rays = np.zeros(n, RAY_DT)
        rays["position"][:, :3] = np.asarray(source, np.float32)
        rays["direction"][:, :3] = d
        rcv = np.asarray(receiver, np.float32)
        src = np.asarray(source, np.float32)
        refl = np.zeros(n, REFL_DT)
        L.refk_init_reflections(_p(refl), n)                                   # reflector.h:44-52
        path = np.zeros(n, np.dtype([("volume", "<f4", (8,)), ("position", "<f4", (4,)), ("distance", "<f4"),
                                     ("pad", "<f4", (3,))]))
        L.refk_init_stochastic_path_info(_p(path), float(initial_energy), _p(src), n)   # finder.cpp:24-34
        for step in range(depth):
            rng = np.ascontiguousarray(rng_for_step(step), np.float32).reshape(n, 2)
            L.refk_reflections(_p(rays), _p(rcv), _p(self.voxel_index), _p(self.aabb), self.side,
                               _p(self.triangles), _p(self.vertices), _p(self.surfaces), _p(rng), _p(refl), n)
            sto = np.zeros(n, IMPULSE_DT)
            hit = np.zeros(n, IMPULSE_DT)
            L.refk_stochastic(_p(refl), _p(rcv), float(receiver_radius), _p(self.triangles), _p(self.vertices),
                              _p(self.surfaces), _p(path), _p(sto), _p(hit), n)
            yield step, refl.copy(), sto, hit


def histogram_from_steps(steps, n_bins, speed_of_sound=340.0, histogram_rate=1000.0, specular_from_step=0):
    """stochastic_group_processor::process + energy_histogram_sum
    (reflection_processor/stochastic_histogram.h:17-32,70-111): impulses with distance == 0 are
    dropped (stochastic/finder.h:65-76); bin = size_t(time * sample_rate); specular (intersected)
    output only when step >= max_image_source_order."""
    hist = np.zeros((n_bins, 8))
    dropped = 0
    for step, _refl, sto, hit in steps:
        groups = [sto] + ([hit] if step >= specular_from_step else [])
        for g in groups:
            live = g[g["distance"] != 0]
            time = live["distance"].astype(np.float64) / speed_of_sound
            b = (time * histogram_rate).astype(np.int64)
            ok = b < n_bins
            dropped += int((~ok).sum())
            np.add.at(hist, b[ok], live["volume"][ok].astype(np.float64))
    return hist, dropped


def reference_histogram(steps, receiver, speed_of_sound=340.0, histogram_rate=1000.0, specular_from_step=0,
                        directional=False):
    """The same through the reference's OWN host code -- incremental_histogram (raytracer/histogram.h:62-81)
    with energy_histogram_sum_functor (stochastic_histogram.h:17-39) and, for the directional case,
    vector_look_up_table<..., 20, 9>::index -- on the impulses in the order the group processor pushes them
    (stochastic first, then specular, step by step): float bins, summed in that order.
    -> float32 [bins, 8] or [20, 9, bins, 8]."""
    vols, poss, dists = [], [], []
    for step, _refl, sto, hit in steps:
        for g in [sto] + ([hit] if step >= specular_from_step else []):
            live = g[g["distance"] != 0]                                      # finder.h:65-76
            vols.append(live["volume"])
            poss.append(live["position"][:, :3])
            dists.append(live["distance"])
    v = np.ascontiguousarray(np.concatenate(vols), np.float32)
    p = np.ascontiguousarray(np.concatenate(poss), np.float32)
    dd = np.ascontiguousarray(np.concatenate(dists), np.float32)
    rcv = np.asarray(receiver, np.float32)
    args = (_p(v), _p(p), _p(dd), dd.size, _p(rcv), float(speed_of_sound), float(histogram_rate), int(directional))
    bins = lib().refk_histogram(*args, None, 0)
    out = np.zeros((20, 9, bins, 8) if directional else (bins, 8), np.float32)
    lib().refk_histogram(*args, _p(out), bins)
    return out


def lut_index(v):
    """vector_look_up_table<T, 20, 9>::index for unit vectors [n, 3] -> (azimuth cell, elevation cell)"""
    v = np.ascontiguousarray(v, np.float32).reshape(-1, 3)
    az, el = np.zeros(v.shape[0], np.int32), np.zeros(v.shape[0], np.int32)
    lib().refk_lut_index(_p(v), v.shape[0], _p(az), _p(el))
    return az, el


# ---- mesh setup ----------------------------------------------------------------------------------
def classify_scene(sc, min_corner, dims, spacing):
    """set_node_inside + set_node_boundary_type (mesh.cpp:73-101) -> condensed_node array with
    boundary_type set (boundary_index still 0)."""
    s = RayScene(sc)
    mc = np.asarray(min_corner, np.float32)
    d = np.asarray(dims, np.int32)
    n = int(d[0]) * int(d[1]) * int(d[2])
    nodes = np.zeros(n, NODE_DT)
    lib().refk_set_node_inside(_p(nodes), _p(mc), _p(d), float(spacing), _p(s.voxel_index), _p(s.aabb), s.side,
                               _p(s.triangles), _p(s.vertices), n)
    lib().refk_set_node_boundary_type(_p(nodes), _p(mc), _p(d), float(spacing), n)
    return nodes


def boundary_type_only(nodes, dims):
    """set_node_boundary_type alone on nodes whose inside flags are already set."""
    out = np.ascontiguousarray(nodes, NODE_DT).copy()
    mc = np.zeros(3, np.float32)
    d = np.asarray(dims, np.int32)
    lib().refk_set_node_boundary_type(_p(out), _p(mc), _p(d), 1.0, out.size)
    return out


def coefficient_indices(sc, nodes, min_corner, dims, spacing, n1, n2, n3):
    """the three finder kernels (boundary_coefficient_finder.cpp:66-124) on numbered nodes."""
    s = RayScene(sc)
    mc = np.asarray(min_corner, np.float32)
    d = np.asarray(dims, np.int32)
    nd = np.ascontiguousarray(nodes, NODE_DT)
    b1 = np.zeros((max(n1, 1), 1), np.uint32)
    b2 = np.zeros((max(n2, 1), 2), np.uint32)
    b3 = np.zeros((max(n3, 1), 3), np.uint32)
    L = lib()
    L.refk_boundary_coefficient_finder_1d(_p(nd), _p(mc), _p(d), float(spacing), _p(b1), _p(s.voxel_index),
                                          _p(s.aabb), s.side, _p(s.triangles), s.triangles.size, _p(s.vertices),
                                          nd.size, 1)
    L.refk_boundary_coefficient_finder_2d(_p(nd), _p(mc), _p(d), float(spacing), _p(b2), _p(b1), nd.size)
    L.refk_boundary_coefficient_finder_3d(_p(nd), _p(mc), _p(d), float(spacing), _p(b3), _p(b1), nd.size)
    return b1[:n1], b2[:n2], b3[:n3]


# ---- scene preparation (HOST code of the reference, compiled behind the GLM stand-in) ------------
def voxelise(vertices4, triangles, depth=5, padding=0.1):
    """make_voxelised_scene_data(scene, depth, padding) + get_flattened through the reference's own
    ndim_tree / voxel_collection / tri_cube_intersection source -> (aabb[6], flattened index)."""
    v = np.ascontiguousarray(vertices4, np.float32).reshape(-1, 4)
    t = np.ascontiguousarray(triangles).view(np.uint32).reshape(-1, 4)
    aabb = np.zeros(6, np.float32)
    n = lib().refk_voxelise(_p(v), v.shape[0], _p(t), t.shape[0], int(depth), float(padding), _p(aabb), None, 0)
    out = np.zeros(n, np.uint32)
    lib().refk_voxelise(_p(v), v.shape[0], _p(t), t.shape[0], int(depth), float(padding), _p(aabb), _p(out), n)
    return aabb, out


def overlaps(box6, tri9) -> bool:
    b = np.ascontiguousarray(box6, np.float32).reshape(6)
    t = np.ascontiguousarray(tri9, np.float32).reshape(9)
    return bool(lib().refk_overlaps(_p(b), _p(t)))


# ---- image sources (HOST code of the reference: image_source/*.cpp, geometric.cpp, the CPU voxel walk) ----
def _scene_arrays(sc):
    v = np.ascontiguousarray(sc.vertices, np.float32).reshape(-1, 4)
    t = np.ascontiguousarray(sc.triangles).view(np.uint32).reshape(-1, 4)
    s = np.ascontiguousarray(sc.surfaces).view(np.float32).reshape(-1, 16)
    return v, t, s


def is_intersects(sc, origins, directions, ignore=None, depth=5, padding=0.1):
    """intersects(voxelised, ray, to_ignore) (voxelised_scene_data.h:80-106) -> (t, triangle or 0xffffffff)"""
    v, t, _ = _scene_arrays(sc)
    o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
    dirs = np.ascontiguousarray(directions, np.float32).reshape(-1, 3)
    ig = None if ignore is None else np.ascontiguousarray(ignore, np.uint32)
    ts, idx = np.zeros(o.shape[0], np.float32), np.zeros(o.shape[0], np.uint32)
    lib().refk_is_intersects(_p(v), v.shape[0], _p(t), t.shape[0], int(depth), float(padding), _p(o), _p(dirs),
                             None if ig is None else _p(ig), o.shape[0], _p(ts), _p(idx))
    return ts, idx


def is_mirror(tri9, point3):
    """geo::mirror(point, triangle), geo::normal(triangle)"""
    out = np.zeros(6, np.float32)
    lib().refk_is_mirror(_p(np.ascontiguousarray(tri9, np.float32)), _p(np.ascontiguousarray(point3, np.float32)), _p(out))
    return out[:3], out[3:]


def is_image_source(sc, reflections, source, receiver, max_order, acoustic_impedance=400.0, flip_phase=False,
                    with_direct=True, depth=5, padding=0.1):
    """the image-source stage (reflection_processor/image_source.cpp:36-68) on reflection records
    [steps][n_rays] -> impulses, in the reference's order"""
    v, t, s = _scene_arrays(sc)
    r = np.ascontiguousarray(reflections, REFL_DT)
    steps, n = r.shape
    src, rcv = np.asarray(source, np.float32), np.asarray(receiver, np.float32)
    cap = 1 << 16
    while True:
        out = np.zeros(cap, IMPULSE_DT)
        cnt = lib().refk_is_image_source(_p(v), v.shape[0], _p(t), t.shape[0], _p(s), s.shape[0], int(depth),
                                         float(padding), _p(src), _p(rcv), _p(r), steps, n, int(max_order),
                                         float(acoustic_impedance), int(flip_phase), int(with_direct), _p(out), cap)
        if cnt <= cap:
            return out[:cnt]
        cap = cnt


def is_exact_shoebox(box_min, box_max, source, receiver, absorption, max_distance, acoustic_impedance=400.0):
    a = [np.asarray(x, np.float32) for x in (box_min, box_max, source, receiver)]
    cap = 1 << 16
    out = np.zeros(cap, IMPULSE_DT)
    n = lib().refk_is_exact_shoebox(_p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), float(absorption), float(max_distance),
                                    float(acoustic_impedance), _p(out), cap)
    assert n <= cap
    return out[:n]


# ---- post-processing (HOST code of the reference: stochastic/postprocessing.cpp, frequency_domain) ----
def pp_rate_law(speed_of_sound, room_volume, t):
    """-> (constant_mean_event_occurrence, mean_event_occurrence(constant, t), t0(constant))"""
    out = np.zeros(3, np.float64)
    lib().refk_pp_rate_law(float(speed_of_sound), float(room_volume), float(t), _p(out))
    return tuple(out)


def pp_intervals(seed, n):
    """interval_size(engine, 1.0) n times for std::default_random_engine{seed} -> (log(1 / x), x)"""
    iv, xs = np.zeros(n, np.float64), np.zeros(n, np.float64)
    lib().refk_pp_intervals(int(seed), n, _p(iv), _p(xs))
    return iv, xs


def pp_dirac_sequence(seed, speed_of_sound, room_volume, sample_rate, max_time):
    """generate_dirac_sequence with its engine seeded by `seed` instead of std::random_device"""
    args = (int(seed), float(speed_of_sound), float(room_volume), float(sample_rate), float(max_time))
    n = lib().refk_pp_dirac_sequence(*args, None, 0)
    out = np.zeros(n, np.float32)
    lib().refk_pp_dirac_sequence(*args, _p(out), n)
    return out


def _pp_hist_seq(histogram, sequence):
    h = np.ascontiguousarray(np.asarray(histogram, np.float64).reshape(-1, 8).astype(np.float32))
    return h, np.ascontiguousarray(sequence, np.float32)


def pp_weight_sequence(histogram, hist_rate, sequence, seq_rate, acoustic_impedance):
    h, s = _pp_hist_seq(histogram, sequence)
    out = np.zeros((s.size, 8), np.float32)
    n = lib().refk_pp_weight_sequence(_p(h), h.shape[0], float(hist_rate), _p(s), s.size, float(seq_rate),
                                      float(acoustic_impedance), _p(out), s.size)
    return out[:n]


def pp_postprocessing(histogram, hist_rate, sequence, seq_rate, acoustic_impedance):
    h, s = _pp_hist_seq(histogram, sequence)
    out = np.zeros(s.size, np.float32)
    n = lib().refk_pp_postprocessing(_p(h), h.shape[0], float(hist_rate), _p(s), s.size, float(seq_rate),
                                     float(acoustic_impedance), _p(out), s.size)
    return out[:n]


def pp_multiband_mixdown(multiband, sample_rate):
    m = np.ascontiguousarray(multiband, np.float32).reshape(-1, 8)
    out = np.zeros(m.shape[0], np.float32)
    lib().refk_pp_multiband_mixdown(_p(m), m.shape[0], float(sample_rate), _p(out))
    return out


def pp_band_params(sample_rate):
    edges, wf = np.zeros(9, np.float64), np.zeros(1, np.float64)
    lib().refk_pp_band_params(float(sample_rate), _p(edges), _p(wf))
    return edges, float(wf[0])


def pp_magnitudes(frequency, edge, edge_hi, width_factor, l=0):
    """-> (lopass at edge, hipass at edge, bandpass over [edge, edge_hi])"""
    out = np.zeros(3, np.float64)
    lib().refk_pp_magnitudes(float(frequency), float(edge), float(edge_hi), float(width_factor), int(l), _p(out))
    return tuple(out)


def pp_crossover(lo, hi, cutoff, width):
    a, b = np.ascontiguousarray(lo, np.float32), np.ascontiguousarray(hi, np.float32)
    out = np.zeros(max(a.size, b.size, 1), np.float32)
    n = lib().refk_pp_crossover(_p(a), a.size, _p(b), b.size, float(cutoff), float(width), _p(out))
    return out[:n]


def pp_left_hanning(length):
    out = np.zeros(length, np.float32)
    lib().refk_pp_left_hanning(length, _p(out))
    return out


def pp_fft_length(n):
    return int(lib().refk_pp_fft_length(n))


# ---- closed-form host functions (waveguide/fitted_boundary.h, filters.cpp, stable.h, ...) --------------
def hm_to_impedance(b, a):
    b, a = np.ascontiguousarray(b, np.float64), np.ascontiguousarray(a, np.float64)
    ob, oa = np.zeros(7), np.zeros(7)
    lib().refk_hm_to_impedance(_p(b), _p(a), _p(ob), _p(oa))
    return ob, oa


def hm_to_flat(absorption):
    ob, oa = np.zeros(7), np.zeros(7)
    lib().refk_hm_to_flat(float(absorption), _p(ob), _p(oa))
    return ob, oa


def hm_is_stable(a) -> bool:
    return bool(lib().refk_hm_is_stable(_p(np.ascontiguousarray(a, np.float64))))


def hm_peak_biquad(gain, centre, q):
    out = np.zeros(6)
    lib().refk_hm_peak_biquad(float(gain), float(centre), float(q), _p(out))
    return out[:3], out[3:]


def hm_convolve3(biquads):
    c = np.ascontiguousarray(biquads, np.float64).reshape(18)
    ob, oa = np.zeros(7), np.zeros(7)
    lib().refk_hm_convolve3(_p(c), _p(ob), _p(oa))
    return ob, oa


def hm_reflectance_filter(absorption, sample_rate, fit):
    """compute_reflectance_filter_coefficients (fitted_boundary.h:79-104) as the reference wrote it, the
    Yule-Walker fit itself supplied by `fit(order, f, m) -> (b, a)` (IT++ is not in the image).
    -> (b, a, f256, m256, threw)"""
    def cb(order, n, f, m, b_out, a_out):
        b, a = fit(order, np.array(f[:n]), np.array(m[:n]))
        for k in range(order + 1):
            b_out[k], a_out[k] = float(b[k]), float(a[k])
    ab = np.ascontiguousarray(absorption, np.float64)
    ob, oa, grid = np.zeros(7), np.zeros(7), np.zeros(512)
    threw = lib().refk_hm_reflectance_filter(_p(ab), float(sample_rate), YULEWALK_CB(cb), _p(ob), _p(oa), _p(grid))
    return ob, oa, grid[:256], grid[256:], bool(threw)


def hm_reflection_number(min_absorption) -> int:
    return int(lib().refk_hm_reflection_number(float(min_absorption)))


def hm_ray_energy(total_rays, source, receiver, radius) -> float:
    s, r = np.asarray(source, np.float32), np.asarray(receiver, np.float32)
    return float(lib().refk_hm_ray_energy(int(total_rays), _p(s), _p(r), float(radius)))


def hm_rates(spacing, speed_of_sound):
    """-> (compute_sample_rate, config::time_step, config::grid_spacing(c, dt), config::speed_of_sound(dt, h))"""
    out = np.zeros(4)
    lib().refk_hm_rates(float(spacing), float(speed_of_sound), _p(out))
    return tuple(out)


# ---- the reference's own waveguide::run template, compiled for the host -----------------------------
def run_waveguide(mesh, source_node=0, signal=(), receivers=(), soft=False, gaussian=None, directional=None,
                  min_corner=(0.0, 0.0, 0.0), spacing=1.0):
    """waveguide::run (waveguide.h:36-126) with the reference's stock processors, all compiled unmodified
    over a host-memory cl.hpp stand-in and enqueueing the reference's kernel as compiled for the host.
      source: hard_source (default) / soft_source (soft=True) fed `signal` at `source_node`, or
              gaussian=(centre xyz, sdev, steps): preprocessor::gaussian
      receivers: node indices read by postprocessor::node under callback_accumulator
      directional=(node, sample_rate, ambient_density): a postprocessor::directional_receiver as well
    -> (steps completed, pressures float32 [steps, receivers], directional float32 [steps, 4] or None)
    raises RuntimeError with the reference's exception text."""
    mc = np.asarray(min_corner, np.float32)
    dims = np.asarray(mesh.dims, np.int32)
    nodes = np.ascontiguousarray(mesh.nodes, NODE_DT)
    coeffs = np.ascontiguousarray(mesh.coeffs, COEFF_DT)
    b1 = np.ascontiguousarray(mesh.b1, np.uint32).reshape(-1, 1)
    b2 = np.ascontiguousarray(mesh.b2, np.uint32).reshape(-1, 2)
    b3 = np.ascontiguousarray(mesh.b3, np.uint32).reshape(-1, 3)
    sig = np.ascontiguousarray(signal, np.float32)
    steps = sig.size
    g4 = np.zeros(4, np.float32)
    kind = 1 if soft else 0
    if gaussian is not None:
        centre, sdev, steps = gaussian
        g4[:3], g4[3], kind = centre, sdev, 2
    rcv = np.ascontiguousarray(receivers, np.uint64)
    out = np.zeros((steps, rcv.size), np.float32)
    dnode, rate, density = (directional if directional is not None else (0xFFFFFFFFFFFFFFFF, 0.0, 0.0))
    dout = np.zeros((steps, 4), np.float32)
    done = C.c_size_t(0)
    err = C.create_string_buffer(512)
    status = lib().refk_run_waveguide(_p(mc), _p(dims), float(spacing), _p(nodes), nodes.size, _p(coeffs), coeffs.size,
                                      _p(b1), b1.shape[0], _p(b2), b2.shape[0], _p(b3), b3.shape[0], kind,
                                      int(source_node), _p(sig), int(steps), _p(g4), _p(rcv), rcv.size, _p(out),
                                      int(dnode), float(rate), float(density), _p(dout), C.byref(done), err, 512)
    if status:
        raise RuntimeError(err.value.decode())
    return done.value, out[:done.value], (dout[:done.value] if directional is not None else None)


def canonical(mesh, surfaces, source, receiver, simulation_time, spacing, min_corner=(0.0, 0.0, 0.0), bands=0,
              cutoff=500.0, usable_portion=0.6, speed_of_sound=340.0, acoustic_impedance=400.0, capacity_steps=4096):
    """waveguide::canonical (canonical.h:97-177) as the reference wrote it: bands == 0 is the
    single_band_parameters overload, otherwise the multiple-band one (flat coefficients per band from
    `surfaces`, SURF_DT-like [n, 16] floats; the mesh must hold one coefficient set per surface).
    -> None when the reference returned nullopt, else a list of bands
       {"directional": float32 [steps, 4] (intensity xyz, pressure), "sample_rate", "valid_hz": (min, max)},
       and the number of times the pressure callback ran. Raises RuntimeError with the reference's text."""
    mc = np.asarray(min_corner, np.float32)
    dims = np.asarray(mesh.dims, np.int32)
    nodes = np.ascontiguousarray(mesh.nodes, NODE_DT)
    coeffs = np.ascontiguousarray(mesh.coeffs, COEFF_DT)
    b1 = np.ascontiguousarray(mesh.b1, np.uint32).reshape(-1, 1)
    b2 = np.ascontiguousarray(mesh.b2, np.uint32).reshape(-1, 2)
    b3 = np.ascontiguousarray(mesh.b3, np.uint32).reshape(-1, 3)
    surf = np.ascontiguousarray(surfaces, np.float32).reshape(-1, 16)
    src, rcv = np.asarray(source, np.float32), np.asarray(receiver, np.float32)
    n_out = max(int(bands), 1)
    out = np.zeros((n_out, capacity_steps, 4), np.float32)
    band3 = np.zeros((n_out, 3), np.float64)
    steps, calls = C.c_size_t(0), C.c_size_t(0)
    err = C.create_string_buffer(512)
    n = lib().refk_canonical(_p(mc), _p(dims), float(spacing), _p(nodes), nodes.size, _p(coeffs), coeffs.size,
                             _p(b1), b1.shape[0], _p(b2), b2.shape[0], _p(b3), b3.shape[0], _p(surf), surf.shape[0],
                             _p(src), _p(rcv), float(speed_of_sound), float(acoustic_impedance), int(bands),
                             float(cutoff), float(usable_portion), float(simulation_time), _p(out), capacity_steps,
                             _p(band3), C.byref(steps), C.byref(calls), err, 512)
    if n < 0:
        raise RuntimeError(err.value.decode())
    if n == 0:
        return None, calls.value
    assert steps.value <= capacity_steps
    return [{"directional": out[b, :steps.value].copy(), "sample_rate": band3[b, 0],
             "valid_hz": (band3[b, 1], band3[b, 2])} for b in range(n)], calls.value


class RefMesh:
    """what waveguide::mesh holds (mesh.h:13-31): descriptor + vectors"""

    def __init__(self, min_corner, dims, spacing, nodes, coeffs, b1, b2, b3, voxel_aabb):
        self.min_corner, self.dims, self.spacing = min_corner, tuple(int(v) for v in dims), spacing
        self.nodes, self.coeffs, self.b1, self.b2, self.b3 = nodes, coeffs, b1, b2, b3
        self.voxel_aabb = voxel_aabb


def compute_mesh(sc, fit, mesh_spacing=None, speed_of_sound=340.0, depth=5, padding=0.1, anchor=None, sample_rate=None):
    """The reference's own mesh construction, run on the host (mesh.cpp, boundary_coefficient_finder.cpp,
    boundary_adjust.cpp with their programs' kernels as compiled for the host):
      anchor=None : compute_mesh(cc, make_voxelised_scene_data(scene, depth, padding), mesh_spacing, c)
      anchor=xyz  : compute_voxels_and_mesh(cc, scene, anchor, sample_rate, c) -- what the engine calls
    `fit(order, f, m) -> (b, a)` is the Yule-Walker fit the IT++ stand-in forwards to. -> RefMesh"""
    v, t, s = _scene_arrays(sc)

    def cb(order, n, f, m, b_out, a_out):
        b, a = fit(order, np.array(f[:n]), np.array(m[:n]))
        for k in range(order + 1):
            b_out[k], a_out[k] = float(b[k]), float(a[k])
    mc, dims, sp = np.zeros(3, np.float32), np.zeros(3, np.int32), np.zeros(1, np.float32)
    counts = np.zeros(5, np.uint64)
    err = C.create_string_buffer(512)
    an = None if anchor is None else np.asarray(anchor, np.float32)
    status = lib().refk_compute_mesh(_p(v), v.shape[0], _p(t), t.shape[0], _p(s), s.shape[0], int(depth), float(padding),
                                     None if an is None else _p(an), float(sample_rate or 0.0),
                                     float(mesh_spacing or 0.0), float(speed_of_sound), YULEWALK_CB(cb), _p(mc), _p(dims),
                                     _p(sp), _p(counts), err, 512)
    if status:
        raise RuntimeError(err.value.decode())
    n, nc, n1, n2, n3 = (int(c) for c in counts)
    nodes, coeffs = np.zeros(n, NODE_DT), np.zeros(nc, COEFF_DT)
    b1, b2, b3 = np.zeros((n1, 1), np.uint32), np.zeros((n2, 2), np.uint32), np.zeros((n3, 3), np.uint32)
    aabb = np.zeros(6, np.float32)
    lib().refk_mesh_read(_p(nodes), _p(coeffs), _p(b1), _p(b2), _p(b3), _p(aabb))
    return RefMesh(mc, dims, float(sp[0]), nodes, coeffs, b1, b2, b3, aabb)


# ---- the reference's own raytracer::run template, compiled for the host -------------------------------
def ray_direction_rng(seed, n):
    """get_direction_rng (reflector.cpp:13-25) with its engine seeded by `seed` -> float32 [n, 2] (z, theta)"""
    out = np.zeros((n, 2), np.float32)
    lib().refk_ray_direction_rng(int(seed), int(n), _p(out))
    return out


def ray_run(sc, source, receiver, directions, seed, image_source_order, specular_from_step=None, total_rays=None,
            receiver_radius=0.1, histogram_rate=1000.0, directional=False, visual_items=0, speed_of_sound=340.0,
            acoustic_impedance=400.0, depth=5, padding=0.1):
    """raytracer::run (raytracer.h:188-266) as the reference wrote it -- reflector.cpp, stochastic/finder.cpp,
    the image-source / histogram / visual processors -- over the host-memory cl.hpp stand-in, enqueueing
    the reference's kernels as compiled for the host. The engine behind each step's (z, theta) draws is
    seeded seed, seed + 1, ... (one per reflector::run_step, in order): ray_direction_rng(seed + k, n)
    reproduces the k-th.
    -> dict(impulses, histogram float32 [bins, 8] or [20, 9, bins, 8], visual REFL_DT [steps, items],
            depth, segments_reported)"""
    v, t, s = _scene_arrays(sc)
    src, rcv = np.asarray(source, np.float32), np.asarray(receiver, np.float32)
    dirs = np.ascontiguousarray(directions, np.float32).reshape(-1, 3)
    counts = np.zeros(4, np.uint64)
    dep = C.c_size_t(0)
    err = C.create_string_buffer(512)
    order = int(image_source_order)
    status = lib().refk_ray_run(_p(v), v.shape[0], _p(t), t.shape[0], _p(s), s.shape[0], int(depth), float(padding),
                                _p(src), _p(rcv), float(speed_of_sound), float(acoustic_impedance), _p(dirs),
                                dirs.shape[0], int(seed), order,
                                int(dirs.shape[0] if total_rays is None else total_rays),
                                int(order + 1 if specular_from_step is None else specular_from_step),
                                float(receiver_radius), float(histogram_rate), int(directional), int(visual_items),
                                _p(counts), C.byref(dep), err, 512)
    if status:
        raise RuntimeError(err.value.decode())
    n_imp, bins, vsteps, calls = (int(c) for c in counts)
    imp = np.zeros(n_imp, IMPULSE_DT)
    hist = np.zeros((20, 9, bins, 8) if directional else (bins, 8), np.float32)
    vis = np.zeros((vsteps, visual_items), REFL_DT)
    lib().refk_ray_read(_p(imp) if n_imp else None, _p(hist) if hist.size else None, _p(vis) if vis.size else None)
    return {"impulses": imp, "histogram": hist, "visual": vis, "depth": dep.value, "segments_reported": calls}
