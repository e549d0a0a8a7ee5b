"""ctypes binding of libwvb200.so (the C ABI declared in include/wvb200.h).

There is deliberately no fallback of any kind: if the shared library is missing
or cannot be loaded the import of the product path fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WVB_LIB") or os.path.join(HERE, "libwvb200.so")  # WVB_LIB: experiment builds

# numpy views of the reference PODs (see include/wvb200.h for file:line)
NODE_DT = np.dtype([("boundary_type", "<i4"), ("boundary_index", "<u4")])
COEFF_DT = np.dtype([("b", "<f8", (7,)), ("a", "<f8", (7,))])
BDATA_DT = np.dtype([("filter_memory", "<f8", (6,)), ("coefficient_index", "<u4"), ("pad", "<u4")])

WVB_OK, WVB_ERR_INVALID, WVB_ERR_CUDA, WVB_ERR_NO_DEVICE, WVB_ERR_NCCL, WVB_ERR_UNSUPPORTED, WVB_ERR_SIM = range(7)
KERNEL_AUTO, KERNEL_DIRECT, KERNEL_TMA = 0, 1, 2
HALO_AUTO, HALO_NCCL, HALO_P2P, HALO_OVERLAP = 0, 1 << 28, 2 << 28, 1 << 30  # wvb_wg_desc.flags
TEMPORAL2 = 1 << 31
RT_MODE_AUTO, RT_MODE_RAY_LIFE, RT_MODE_WAVEFRONT = 0, 1, 2  # wvb_rt_trace_params.mode


class WgDesc(C.Structure):
    _fields_ = [
        ("dim", C.c_int32 * 3),
        ("z_begin", C.c_int32), ("z_end", C.c_int32),
        ("nodes", C.c_void_p),
        ("nodes_z0", C.c_int32), ("nodes_nz", C.c_int32),
        ("coefficients", C.c_void_p),
        ("num_coefficients", C.c_uint32),
        ("boundary_index", C.c_void_p * 3),
        ("boundary_count", C.c_uint64 * 3),
        ("index_base", C.c_uint32 * 3),
        ("device", C.c_int32),
        ("rank", C.c_int32), ("nranks", C.c_int32),
        ("nccl_unique_id", C.c_void_p),
        ("flags", C.c_uint32),
    ]


class WgRunParams(C.Structure):
    _fields_ = [
        ("source_node", C.c_uint64),
        ("signal", C.c_void_p),
        ("n_steps", C.c_uint32),
        ("soft", C.c_int32),
        ("receiver_nodes", C.c_void_p),
        ("n_receivers", C.c_uint32),
        ("out", C.c_void_p),
        ("check_interval", C.c_uint32),
        ("keep_going", C.c_void_p),
        ("keep_going_user", C.c_void_p),
    ]


class WgInfo(C.Structure):
    _fields_ = [
        ("local_nodes", C.c_uint64), ("air_nodes", C.c_uint64),
        ("boundary_nodes", C.c_uint64 * 3),
        ("device_bytes", C.c_uint64), ("kernel_launches", C.c_uint64),
        ("kernel_variant", C.c_int32), ("tile", C.c_int32 * 3),
        ("sm_count", C.c_int32), ("halo", C.c_int32),
    ]


class PpParams(C.Structure):
    _fields_ = [
        ("speed_of_sound", C.c_double), ("acoustic_impedance", C.c_double), ("room_volume", C.c_double),
        ("histogram_sample_rate", C.c_double), ("output_sample_rate", C.c_double), ("max_time", C.c_double),
        ("seed", C.c_uint64), ("device", C.c_int32), ("pad", C.c_int32),
    ]


class RtSceneDesc(C.Structure):
    _fields_ = [
        ("voxel_index", C.c_void_p), ("voxel_index_count", C.c_uint64),
        ("aabb_min", C.c_float * 3), ("aabb_max", C.c_float * 3), ("side", C.c_uint32),
        ("triangles", C.c_void_p), ("num_triangles", C.c_uint32),
        ("vertices", C.c_void_p), ("num_vertices", C.c_uint32),
        ("surfaces", C.c_void_p), ("num_surfaces", C.c_uint32),
        ("device", C.c_int32),
    ]


class RtTraceParams(C.Structure):
    _fields_ = [
        ("source", C.c_float * 3), ("receiver", C.c_float * 3),
        ("receiver_radius", C.c_float), ("pad0", C.c_float),
        ("speed_of_sound", C.c_double), ("histogram_sample_rate", C.c_double),
        ("total_rays", C.c_uint64), ("seed", C.c_uint64), ("ray_index_base", C.c_uint64),
        ("depth", C.c_uint32), ("specular_from_step", C.c_uint32), ("n_bins", C.c_uint32),
        ("directional", C.c_uint32), ("keep_steps", C.c_uint32), ("mode", C.c_uint32),
    ]


class IsDesc(C.Structure):
    _fields_ = [
        ("source", C.c_float * 3), ("receiver", C.c_float * 3),
        ("acoustic_impedance", C.c_double), ("flip_phase", C.c_int32), ("with_direct", C.c_int32),
        ("max_elements", C.c_uint64),
    ]


IMPULSE_DT = np.dtype([("volume", "<f4", (8,)), ("position", "<f4", (4,)), ("distance", "<f4"),
                       ("pad_", "<f4", (3,))])  # raytracer::impulse<8>
assert IMPULSE_DT.itemsize == 64

# every symbol include/wvb200.h declares (tests check the .so exports them all)
WG_SYMBOLS = [
    "wvb_wg_create", "wvb_wg_destroy", "wvb_wg_write_f64", "wvb_wg_read_f64", "wvb_wg_read_field",
    "wvb_wg_read_field_f32", "wvb_wg_write_field", "wvb_wg_step", "wvb_wg_launch", "wvb_wg_swap",
    "wvb_wg_run",
    "wvb_wg_boundary_count", "wvb_wg_read_boundary_data", "wvb_wg_time_steps", "wvb_wg_time_kernels",
    "wvb_wg_get_info", "wvb_nccl_unique_id", "wvb_test_third", "wvb_test_filter",
    "wvb_mesh_cuboid", "wvb_version", "wvb_device_count", "wvb_last_error",
]

MESH_SYMBOLS = ["wvb_mesh_create", "wvb_mesh_destroy", "wvb_mesh_counts", "wvb_mesh_read"]

LRS_SYMBOLS = ["wvb_lrs_arbitrary_magnitude_filter", "wvb_lrs_reflectance_filter", "wvb_lrs_to_impedance",
               "wvb_lrs_flat", "wvb_lrs_is_stable"]

IS_SYMBOLS = ["wvb_is_create", "wvb_is_destroy", "wvb_is_push_elements", "wvb_is_push_reflections",
              "wvb_is_trace", "wvb_is_results"]

RT_SYMBOLS = [
    "wvb_rt_create", "wvb_rt_destroy", "wvb_rt_trace", "wvb_rt_read_histogram", "wvb_rt_reset_histogram",
    "wvb_rt_reflection_depth", "wvb_rt_ray_energy", "wvb_rt_safe_bins", "wvb_rt_closest_hit",
    "wvb_rt_directions", "wvb_rt_comm_init", "wvb_rt_allreduce_histogram",
]

SCENE_SYMBOLS = ["wvb_voxelise", "wvb_obj_parse"]
PP_SYMBOLS = ["wvb_pp_dirac_sequence", "wvb_pp_stochastic", "wvb_pp_multiband_mixdown", "wvb_pp_crossover"]

_lib = None


class WvbError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("wvb status %d: %s" % (status, message))
        self.status = status


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s is missing: build it with `python -m wayverb_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32
    L.wvb_version.restype = C.c_int
    L.wvb_device_count.restype = C.c_int
    L.wvb_last_error.restype = C.c_char_p
    L.wvb_wg_create.argtypes = [C.POINTER(WgDesc), C.POINTER(vp)]
    L.wvb_wg_destroy.argtypes = [vp]
    L.wvb_wg_destroy.restype = None
    L.wvb_wg_write_f64.argtypes = [vp, u64, C.c_double]
    L.wvb_wg_read_f64.argtypes = [vp, u64, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.wvb_wg_read_field.argtypes = [vp, vp]
    L.wvb_wg_read_field_f32.argtypes = [vp, vp]
    L.wvb_wg_write_field.argtypes = [vp, vp]
    L.wvb_wg_step.argtypes = [vp, u32, C.POINTER(i32)]
    L.wvb_wg_launch.argtypes = [vp, C.POINTER(i32)]
    L.wvb_wg_swap.argtypes = [vp]
    L.wvb_wg_run.argtypes = [vp, C.POINTER(WgRunParams), C.POINTER(u32), C.POINTER(i32)]
    L.wvb_wg_boundary_count.argtypes = [vp, C.c_int, C.POINTER(u64)]
    L.wvb_wg_read_boundary_data.argtypes = [vp, C.c_int, vp]
    L.wvb_wg_time_steps.argtypes = [vp, u32, C.POINTER(C.c_float), C.POINTER(i32)]
    L.wvb_wg_time_kernels.argtypes = [vp, u32, C.POINTER(C.c_float * 2)]
    L.wvb_test_third.argtypes = [vp, C.c_size_t, vp, vp]
    L.wvb_test_filter.argtypes = [vp, vp, vp, u32, u32, vp]
    L.wvb_nccl_unique_id.argtypes = [vp, C.c_size_t]
    L.wvb_wg_get_info.argtypes = [vp, C.POINTER(WgInfo)]
    L.wvb_mesh_cuboid.argtypes = [C.POINTER(i32 * 3), i32, i32, vp, C.POINTER(u64 * 3)]
    L.wvb_mesh_create.argtypes = [vp, C.POINTER(C.c_float * 3), C.POINTER(i32 * 3), C.c_float, vp, vp, i32,
                                  C.POINTER(vp)]
    L.wvb_mesh_destroy.argtypes = [vp]
    L.wvb_mesh_destroy.restype = None
    L.wvb_mesh_counts.argtypes = [vp, C.POINTER(u64 * 3)]
    L.wvb_mesh_read.argtypes = [vp, vp, vp, vp, vp, vp]
    L.wvb_rt_create.argtypes = [C.POINTER(RtSceneDesc), C.POINTER(vp)]
    L.wvb_rt_destroy.argtypes = [vp]
    L.wvb_rt_destroy.restype = None
    L.wvb_rt_trace.argtypes = [vp, C.POINTER(RtTraceParams), vp, u64, vp, C.POINTER(u64), C.POINTER(C.c_float)]
    L.wvb_rt_read_histogram.argtypes = [vp, vp]
    L.wvb_rt_reset_histogram.argtypes = [vp]
    L.wvb_rt_reflection_depth.restype = u32
    L.wvb_rt_reflection_depth.argtypes = [C.c_double]
    L.wvb_rt_ray_energy.restype = C.c_float
    L.wvb_rt_ray_energy.argtypes = [u64, vp, vp, C.c_float]
    L.wvb_rt_safe_bins.restype = u32
    L.wvb_rt_safe_bins.argtypes = [vp, u32, C.c_double, C.c_double]
    L.wvb_rt_closest_hit.argtypes = [vp, vp, u64, vp, vp]
    L.wvb_rt_directions.argtypes = [vp, u64, u64, u64, vp]
    L.wvb_rt_comm_init.argtypes = [vp, vp, i32, i32]
    L.wvb_rt_allreduce_histogram.argtypes = [vp]
    L.wvb_lrs_arbitrary_magnitude_filter.argtypes = [vp, vp, u32, vp]
    L.wvb_lrs_reflectance_filter.argtypes = [vp, C.c_double, vp]
    L.wvb_lrs_to_impedance.argtypes = [vp, vp]
    L.wvb_lrs_to_impedance.restype = None
    L.wvb_lrs_flat.argtypes = [C.c_double, vp]
    L.wvb_lrs_flat.restype = None
    L.wvb_lrs_is_stable.argtypes = [vp, u32]
    L.wvb_lrs_is_stable.restype = C.c_int
    L.wvb_is_create.argtypes = [vp, C.POINTER(IsDesc), C.POINTER(vp)]
    L.wvb_is_destroy.argtypes = [vp]
    L.wvb_is_destroy.restype = None
    L.wvb_is_push_elements.argtypes = [vp, vp, u64, u32, u64]
    L.wvb_is_push_reflections.argtypes = [vp, vp, u64, u32, u64]
    L.wvb_is_trace.argtypes = [vp, C.POINTER(RtTraceParams), vp, u64, u32, vp, C.POINTER(u64),
                               C.POINTER(C.c_float)]
    L.wvb_is_results.argtypes = [vp, vp, u64, C.POINTER(u64), C.POINTER(u64 * 4), C.POINTER(C.c_float)]
    L.wvb_voxelise.argtypes = [vp, u32, vp, u32, u32, C.c_float, C.POINTER(C.c_float * 3), C.POINTER(C.c_float * 3),
                               vp, u64, C.POINTER(u64)]
    L.wvb_pp_dirac_sequence.argtypes = [C.POINTER(PpParams), C.c_double, C.c_double, vp, u64, C.POINTER(u64),
                                        C.POINTER(u32)]
    L.wvb_pp_stochastic.argtypes = [vp, u32, C.POINTER(PpParams), vp, u64, C.POINTER(u64), vp]
    L.wvb_pp_multiband_mixdown.argtypes = [vp, u64, C.c_double, i32, vp]
    L.wvb_pp_crossover.argtypes = [vp, u64, vp, u64, C.c_double, C.c_double, u64, i32, vp, u64]
    L.wvb_obj_parse.argtypes = [C.c_char_p, u64, vp, C.POINTER(u64), vp, C.POINTER(u64), vp, C.POINTER(u64)]
    _lib = L
    return L


def check(status, allow_sim=False):
    if status == WVB_OK or (allow_sim and status == WVB_ERR_SIM):
        return status
    raise WvbError(status, lib().wvb_last_error().decode())


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)
