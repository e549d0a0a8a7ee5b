"""Host mirror of the reference's waveguide interface over the C ABI.

Names and argument meaning follow the reference so that the parity tests read
like its own tests:

  Mesh                       waveguide::mesh / vectors   (mesh.h:12-26, setup.h:27-48)
  run(mesh, pre, post, ...)  waveguide::run              (waveguide.h:36-126)
  hard_source / soft_source  preprocessor::*             (preprocessor/hard_source.h, soft_source.h)
  node_receiver              postprocessor::node         (postprocessor/node.cpp:14-18)

`Waveguide` is the handle (`wvb_wg`); its read/write are core::read_value /
write_value on the `current` buffer (core/cl/common.h:42-57).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import numpy as np

from . import _lib
from ._lib import BDATA_DT, COEFF_DT, NODE_DT, WgDesc, WgInfo, WgRunParams, check, lib, ptr

ERROR_MESSAGES = {  # waveguide.h:100-119
    1: "Pressure value is inf, check filter coefficients.",
    2: "Pressure value is nan, check filter coefficients.",
    8: "Tried to read non-existant node.",
    16: "Suspicious boundary read.",
}


class ValueIsInf(RuntimeError):  # core::exceptions::value_is_inf
    pass


class ValueIsNan(RuntimeError):  # core::exceptions::value_is_nan
    pass


def raise_for_flags(flags: int):
    """Same precedence and messages as waveguide.h:100-119."""
    if flags & 1:
        raise ValueIsInf(ERROR_MESSAGES[1])
    if flags & 2:
        raise ValueIsNan(ERROR_MESSAGES[2])
    if flags & 8:
        raise RuntimeError(ERROR_MESSAGES[8])
    if flags & 16:
        raise RuntimeError(ERROR_MESSAGES[16])


def slab_range(dz: int, rank: int, nranks: int):
    """z-planes owned by `rank`: contiguous, balanced, covering [0, dz)."""
    return (dz * rank) // nranks, (dz * (rank + 1)) // nranks


class Mesh:
    """dims (x fastest) + condensed nodes + coefficients + boundary_index_array_{1,2,3}.

    `nodes` may hold only planes [nodes_z0, nodes_z0 + nodes_nz) of the mesh, and
    the index arrays only the entries from index_base on (slab-local pieces)."""

    def __init__(self, dims, nodes, coeffs, b1, b2, b3, nodes_z0=0, index_base=(0, 0, 0)):
        self.dims = tuple(int(d) for d in dims)
        self.nodes = np.ascontiguousarray(nodes, NODE_DT).reshape(-1)
        self.coeffs = np.ascontiguousarray(coeffs, COEFF_DT).reshape(-1)
        self.b = [np.ascontiguousarray(b1, np.uint32).reshape(-1, 1),
                  np.ascontiguousarray(b2, np.uint32).reshape(-1, 2),
                  np.ascontiguousarray(b3, np.uint32).reshape(-1, 3)]
        self.nodes_z0 = int(nodes_z0)
        plane = self.dims[0] * self.dims[1]
        assert self.nodes.size % plane == 0
        self.nodes_nz = self.nodes.size // plane
        self.index_base = tuple(int(i) for i in index_base)

    @property
    def num_nodes(self):
        return self.dims[0] * self.dims[1] * self.dims[2]

    def index(self, x, y, z):  # compute_index, mesh_descriptor.cpp:7-10
        dx, dy, _ = self.dims
        return int(x) + int(y) * dx + int(z) * dx * dy


def cuboid_nodes(dims, z0=0, nz=None):
    """wvb_mesh_cuboid: nodes of planes [z0, z0+nz) + global class counts."""
    dx, dy, dz = (int(d) for d in dims)
    nz = dz - z0 if nz is None else int(nz)
    nodes = np.zeros(dx * dy * nz, NODE_DT)
    counts = (C.c_uint64 * 3)()
    d3 = (C.c_int32 * 3)(dx, dy, dz)
    check(lib().wvb_mesh_cuboid(C.byref(d3), int(z0), nz, ptr(nodes), C.byref(counts)))
    return nodes, tuple(int(c) for c in counts)


def cuboid_mesh(dims, coeffs, z0=0, nz=None) -> Mesh:
    """Synthetic box room (BASELINE configs 2-4) with one surface (index 0)."""
    nodes, counts = cuboid_nodes(dims, z0, nz)
    c = np.ascontiguousarray(coeffs, COEFF_DT).reshape(-1)
    return Mesh(dims, nodes, c, np.zeros((counts[0], 1), np.uint32), np.zeros((counts[1], 2), np.uint32),
                np.zeros((counts[2], 3), np.uint32), nodes_z0=z0)


def nccl_unique_id() -> bytes:
    """128 bytes of a fresh ncclUniqueId (rank 0 creates, everybody gets a copy)."""
    buf = C.create_string_buffer(128)
    check(lib().wvb_nccl_unique_id(buf, 128))
    return buf.raw


def build_mesh(dims, min_corner, spacing, coeffs, scene=None, inside=None, surface_1d=None, device=0,
               return_inside=False):
    """compute_mesh's node part (mesh.cpp:53-141) on the device -> Mesh.
    scene: a RayTracer (the voxelised scene) or None when `inside` and `surface_1d`
    (arrays indexed like the mesh, x fastest) are given."""
    dx, dy, dz = (int(d) for d in dims)
    mc = (C.c_float * 3)(*[float(v) for v in min_corner])
    d3 = (C.c_int32 * 3)(dx, dy, dz)
    ins = None if inside is None else np.ascontiguousarray(inside, np.uint8).reshape(-1)
    s1 = None if surface_1d is None else np.ascontiguousarray(surface_1d, np.uint32).reshape(-1)
    h = C.c_void_p()
    check(lib().wvb_mesh_create(scene._h if scene is not None else None, C.byref(mc), C.byref(d3),
                                float(spacing), ptr(ins) if ins is not None else None,
                                ptr(s1) if s1 is not None else None, int(device), C.byref(h)))
    try:
        counts = (C.c_uint64 * 3)()
        check(lib().wvb_mesh_counts(h, C.byref(counts)))
        n1, n2, n3 = (int(c) for c in counts)
        nodes = np.zeros(dx * dy * dz, NODE_DT)
        b1, b2, b3 = np.zeros(max(n1, 1), np.uint32), np.zeros(max(n2, 1) * 2, np.uint32), \
            np.zeros(max(n3, 1) * 3, np.uint32)
        mask = np.zeros(dx * dy * dz, np.uint8)
        check(lib().wvb_mesh_read(h, ptr(nodes), ptr(b1), ptr(b2), ptr(b3), ptr(mask)))
    finally:
        lib().wvb_mesh_destroy(h)
    m = Mesh(dims, nodes, coeffs, b1[:n1], b2[:n2 * 2], b3[:n3 * 3])
    return (m, mask.reshape(dz, dy, dx)) if return_inside else m


class Waveguide:
    """One `wvb_wg` handle: a z-slab of the mesh on one GPU."""

    def __init__(self, mesh: Mesh, device=0, z_range=None, rank=0, nranks=1, nccl_unique_id: Optional[bytes] = None,
                 kernel=_lib.KERNEL_AUTO, flags=0):
        self.mesh = mesh
        dz = mesh.dims[2]
        z0, z1 = (0, dz) if z_range is None else z_range
        d = WgDesc()
        d.dim[:] = mesh.dims
        d.z_begin, d.z_end = int(z0), int(z1)
        d.nodes = mesh.nodes.ctypes.data
        d.nodes_z0, d.nodes_nz = mesh.nodes_z0, mesh.nodes_nz
        d.coefficients = mesh.coeffs.ctypes.data
        d.num_coefficients = mesh.coeffs.size
        for k in range(3):
            d.boundary_index[k] = mesh.b[k].ctypes.data if mesh.b[k].size else None
            d.boundary_count[k] = mesh.b[k].shape[0]
            d.index_base[k] = mesh.index_base[k]
        d.device = int(device)
        d.rank, d.nranks = int(rank), int(nranks)
        self._uid = C.create_string_buffer(nccl_unique_id, 128) if nccl_unique_id else None
        d.nccl_unique_id = C.cast(self._uid, C.c_void_p) if self._uid else None
        d.flags = int(kernel) | int(flags)
        self.z_range = (int(z0), int(z1))
        h = C.c_void_p()
        check(lib().wvb_wg_create(C.byref(d), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().wvb_wg_destroy(self._h)
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

    # core::write_value / read_value on `current`
    def write(self, node, value):
        check(lib().wvb_wg_write_f64(self._h, int(node), float(value)))

    def read(self, node) -> float:
        v, o = C.c_double(), C.c_int()
        check(lib().wvb_wg_read_f64(self._h, int(node), C.byref(v), C.byref(o)))
        return v.value

    def owns(self, node) -> bool:
        v, o = C.c_double(), C.c_int()
        check(lib().wvb_wg_read_f64(self._h, int(node), C.byref(v), C.byref(o)))
        return bool(o.value)

    def field(self) -> np.ndarray:
        dx, dy, _ = self.mesh.dims
        out = np.zeros(dx * dy * (self.z_range[1] - self.z_range[0]))
        check(lib().wvb_wg_read_field(self._h, ptr(out)))
        return out

    def field_f32(self) -> np.ndarray:
        dx, dy, _ = self.mesh.dims
        out = np.zeros(dx * dy * (self.z_range[1] - self.z_range[0]), np.float32)
        check(lib().wvb_wg_read_field_f32(self._h, ptr(out)))
        return out

    def set_field(self, f):
        a = np.ascontiguousarray(f, np.float64).reshape(-1)
        dx, dy, _ = self.mesh.dims
        assert a.size == dx * dy * (self.z_range[1] - self.z_range[0])
        check(lib().wvb_wg_write_field(self._h, ptr(a)))

    def step(self, n=1) -> int:
        """n kernel launches + swaps; returns the error_code bits (0 = ok)."""
        f = C.c_int32()
        check(lib().wvb_wg_step(self._h, int(n), C.byref(f)), allow_sim=True)
        return f.value

    def launch(self) -> int:
        """kernel launch + flag readback, no swap (waveguide.h:82-100)."""
        f = C.c_int32()
        check(lib().wvb_wg_launch(self._h, C.byref(f)), allow_sim=True)
        return f.value

    def swap(self):
        check(lib().wvb_wg_swap(self._h))

    def time_steps(self, n):
        ms, f = C.c_float(), C.c_int32()
        check(lib().wvb_wg_time_steps(self._h, int(n), C.byref(ms), C.byref(f)), allow_sim=True)
        return ms.value, f.value

    def time_kernels(self, n):
        """(air kernel ms, boundary kernels ms) for n launches each; invalidates the field."""
        ms = (C.c_float * 2)()
        check(lib().wvb_wg_time_kernels(self._h, int(n), C.byref(ms)))
        return float(ms[0]), float(ms[1])

    def run_device(self, src_node, signal, rcv_nodes, soft=False, check_interval=0):
        """wvb_wg_run: the stock hard/soft source + node receivers on the device."""
        sig = np.ascontiguousarray(signal, np.float64)
        rcv = np.ascontiguousarray(rcv_nodes, np.uint64)
        out = np.zeros((sig.size, rcv.size))
        p = WgRunParams()
        p.source_node = int(src_node)
        p.signal = sig.ctypes.data
        p.n_steps = sig.size
        p.soft = int(bool(soft))
        p.receiver_nodes = rcv.ctypes.data if rcv.size else None
        p.n_receivers = rcv.size
        p.out = out.ctypes.data if rcv.size else None
        p.check_interval = int(check_interval)
        done, f = C.c_uint32(), C.c_int32()
        check(lib().wvb_wg_run(self._h, C.byref(p), C.byref(done), C.byref(f)), allow_sim=True)
        return done.value, out, f.value

    def boundary_data(self, n) -> np.ndarray:
        cnt = C.c_uint64()
        check(lib().wvb_wg_boundary_count(self._h, n, C.byref(cnt)))
        out = np.zeros((cnt.value, n), BDATA_DT)
        if cnt.value:
            check(lib().wvb_wg_read_boundary_data(self._h, n, ptr(out)))
        return out

    def info(self) -> dict:
        i = WgInfo()
        check(lib().wvb_wg_get_info(self._h, C.byref(i)))
        return {"local_nodes": i.local_nodes, "air_nodes": i.air_nodes,
                "boundary_nodes": tuple(i.boundary_nodes), "device_bytes": i.device_bytes,
                "kernel_launches": i.kernel_launches,
                "kernel_variant": {1: "direct", 2: "tma"}.get(i.kernel_variant, "?"),
                "tile": tuple(i.tile), "sm_count": i.sm_count,
                "halo": {0: "none", 1: "nccl", 2: "p2p"}.get(i.halo & 3, "?") + ("+overlap" if i.halo & 4 else "")}


# ---- stock processors ------------------------------------------------------------
def hard_source(node, signal):
    """preprocessor::hard_source: overwrite the node every step; False when exhausted."""
    it = iter(signal)

    def pre(wg: Waveguide, step: int) -> bool:
        try:
            v = next(it)
        except StopIteration:
            return False
        wg.write(node, v)
        return True
    return pre


def soft_source(node, signal):
    """preprocessor::soft_source: add to the node every step."""
    it = iter(signal)

    def pre(wg: Waveguide, step: int) -> bool:
        try:
            v = next(it)
        except StopIteration:
            return False
        wg.write(node, wg.read(node) + v)
        return True
    return pre


def node_receiver(node, out: list):
    """postprocessor::node: append current[node] each step."""
    def post(wg: Waveguide, step: int):
        out.append(wg.read(node))
    return post


def run(mesh: Mesh, pre: Callable, post: Callable, keep_going: Callable[[], bool] = lambda: True,
        device=0, **kw) -> int:
    """waveguide::run (waveguide.h:36-126): pre -> kernel -> flag check -> post -> swap.

    pre(wg, step) -> bool (False ends the run); post(wg, step) sees `current`,
    i.e. p(step) including what pre injected. Returns the steps completed; raises
    like the reference on error flags."""
    with Waveguide(mesh, device=device, **kw) as wg:
        step = 0
        while pre(wg, step) and keep_going():
            flags = wg.launch()
            if flags:
                raise_for_flags(flags)
            post(wg, step)
            wg.swap()
            step += 1
        return step
