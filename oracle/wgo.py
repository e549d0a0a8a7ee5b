"""ctypes front-end of the CPU waveguide oracle (oracle/wg_oracle.cpp).

TEST INFRASTRUCTURE ONLY. Imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py. Never imported by the
wayverb_b200 package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libwgoracle.so")

# numpy views of the reference PODs (layouts: see wg_oracle.cpp header)
NODE_DT = np.dtype([("boundary_type", "<i4"), ("boundary_index", "<u4")])
COEFF_DT = np.dtype([("b", "<f8", (7,)), ("a", "<f8", (7,))])
BDATA_DT = np.dtype([("mem", "<f8", (6,)), ("coefficient_index", "<u4"), ("pad", "<u4")])
assert NODE_DT.itemsize == 8 and COEFF_DT.itemsize == 112 and BDATA_DT.itemsize == 56

ID_NONE, ID_INSIDE = 0, 1
ID_NX, ID_PX, ID_NY, ID_PY, ID_NZ, ID_PZ, ID_REENTRANT = 2, 4, 8, 16, 32, 64, 128
ERR_INF, ERR_NAN, ERR_OUTSIDE_RANGE, ERR_OUTSIDE_MESH, ERR_SUSPICIOUS = 1, 2, 4, 8, 16


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only, seconds)."""
    src = os.path.join(_HERE, "wg_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "_build/libwgoracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None
_native = False


def use_native() -> bool:
    """bench.py's CPU legs: load the -O3 -march=native build of the same source (Makefile target
    `native`, compiled ON the machine that runs it -- a -march=native binary must never travel).
    Must be called before the first lib(); falls back to the portable build when g++ fails."""
    global _LIB_PATH, _native
    if _lib is not None:
        return _native
    path = os.path.join(_HERE, "_build", "native", "libwgoracle.so")
    try:
        subprocess.run(["make", "-B", "-C", _HERE, "_build/native/libwgoracle.so"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        _LIB_PATH, _native = path, True
    except (subprocess.CalledProcessError, OSError):
        _native = False
    return _native


def is_native() -> bool:
    return _native


def lib():
    global _lib
    if _lib is None:
        if not _native:
            build()
        L = C.CDLL(_LIB_PATH)
        vp, sz, u32p, dp = C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_double)
        L.wgo_create.restype = vp
        L.wgo_create.argtypes = [C.c_int, C.c_int, C.c_int, vp, vp, C.c_int,
                                 vp, sz, vp, sz, vp, sz, C.c_int]
        L.wgo_destroy.argtypes = [vp]
        L.wgo_write.argtypes = [vp, sz, C.c_double]
        L.wgo_read.restype = C.c_double
        L.wgo_read.argtypes = [vp, sz]
        L.wgo_field.argtypes = [vp, vp]
        L.wgo_set_field.argtypes = [vp, vp]
        L.wgo_step.restype = C.c_int
        L.wgo_step.argtypes = [vp, C.c_int]
        L.wgo_run.restype = sz
        L.wgo_run.argtypes = [vp, sz, vp, sz, C.c_int, vp, sz, vp, C.POINTER(C.c_int)]
        L.wgo_boundary_count.restype = sz
        L.wgo_boundary_count.argtypes = [vp, C.c_int]
        L.wgo_boundary_data.argtypes = [vp, C.c_int, vp]
        L.wgo_to_impedance.argtypes = [vp, vp]
        L.wgo_to_flat.argtypes = [C.c_double, vp]
        L.wgo_peak_biquad.argtypes = [C.c_double, C.c_double, C.c_double, vp]
        L.wgo_convolve3.argtypes = [vp, vp]
        L.wgo_filter_biquads.argtypes = [vp, vp, vp, vp, sz]
        L.wgo_filter_canonical.argtypes = [vp, vp, vp, vp, sz]
        L.wgo_filter_canonical_f64.argtypes = [vp, vp, vp, vp, sz]
        L.wgo_classify.argtypes = [C.c_int, C.c_int, C.c_int, vp, vp]
        L.wgo_count_boundaries.argtypes = [sz, vp, vp]
        L.wgo_boundary_indices.restype = C.c_int
        L.wgo_boundary_indices.argtypes = [C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp]
        L.wgo_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ---- coefficient helpers -----------------------------------------------------
def to_impedance(reflectance: np.ndarray) -> np.ndarray:
    out = np.zeros((), COEFF_DT)
    r = np.ascontiguousarray(reflectance, COEFF_DT)
    lib().wgo_to_impedance(_p(r), _p(out))
    return out


def to_flat(absorption: float) -> np.ndarray:
    out = np.zeros((), COEFF_DT)
    lib().wgo_to_flat(float(absorption), _p(out))
    return out


def peak_biquad(gain_db, centre, Q) -> np.ndarray:
    out = np.zeros(6)
    lib().wgo_peak_biquad(gain_db, centre, Q, _p(out))
    return out


def convolve3(biquads: np.ndarray) -> np.ndarray:
    b = np.ascontiguousarray(biquads, np.float64).reshape(18)
    out = np.zeros((), COEFF_DT)
    lib().wgo_convolve3(_p(b), _p(out))
    return out


def filter_biquads(biquads, x_f32, mem=None):
    b = np.ascontiguousarray(biquads, np.float64).reshape(18)
    m = np.zeros(6) if mem is None else mem
    x = np.ascontiguousarray(x_f32, np.float32)
    y = np.zeros_like(x)
    lib().wgo_filter_biquads(_p(b), _p(m), _p(x), _p(y), x.size)
    return y


def filter_canonical(coeffs, x, mem=None, f64=False):
    c = np.ascontiguousarray(coeffs, COEFF_DT)
    m = np.zeros(6) if mem is None else mem
    if f64:
        xx = np.ascontiguousarray(x, np.float64)
        y = np.zeros_like(xx)
        lib().wgo_filter_canonical_f64(_p(c), _p(m), _p(xx), _p(y), xx.size)
    else:
        xx = np.ascontiguousarray(x, np.float32)
        y = np.zeros_like(xx)
        lib().wgo_filter_canonical(_p(c), _p(m), _p(xx), _p(y), xx.size)
    return y


# ---- mesh construction ---------------------------------------------------------
class Mesh:
    """What `waveguide::mesh` holds (mesh.h:12-26): dims + nodes + coefficients +
    boundary_index_array_{1,2,3}."""

    def __init__(self, dims, nodes, coeffs, b1, b2, b3):
        self.dims = tuple(int(d) for d in dims)
        self.nodes = nodes
        self.coeffs = np.ascontiguousarray(coeffs, COEFF_DT).reshape(-1)
        self.b1 = np.ascontiguousarray(b1, np.uint32).reshape(-1, 1)
        self.b2 = np.ascontiguousarray(b2, np.uint32).reshape(-1, 2)
        self.b3 = np.ascontiguousarray(b3, np.uint32).reshape(-1, 3)

    @property
    def num_nodes(self):
        return self.nodes.size

    def index(self, x, y, z):
        dx, dy, _ = self.dims
        return int(x) + int(y) * dx + int(z) * dx * dy


def classify(inside: np.ndarray) -> np.ndarray:
    """inside: bool array indexed [z, y, x] (x fastest, like the mesh)."""
    dz, dy, dx = inside.shape
    ins = np.ascontiguousarray(inside, np.uint8)
    nodes = np.zeros(ins.size, NODE_DT)
    lib().wgo_classify(dx, dy, dz, _p(ins), _p(nodes))
    return nodes


def mesh_from_inside(inside: np.ndarray, coeffs, surface_1d=None) -> Mesh:
    """Restates compute_mesh's node/boundary part (mesh.cpp:53-141) for a given
    inside mask. surface_1d: uint32 per node (flattened [z,y,x]); default 0."""
    dz, dy, dx = inside.shape
    nodes = classify(inside)
    counts = np.zeros(3, np.uint64)
    lib().wgo_count_boundaries(nodes.size, _p(nodes), _p(counts))
    n1, n2, n3 = (int(c) for c in counts)
    s1 = np.zeros(nodes.size, np.uint32) if surface_1d is None else \
        np.ascontiguousarray(surface_1d, np.uint32).reshape(-1)
    b1 = np.zeros(max(n1, 1), np.uint32)
    b2 = np.zeros(max(n2, 1) * 2, np.uint32)
    b3 = np.zeros(max(n3, 1) * 3, np.uint32)
    lib().wgo_boundary_indices(dx, dy, dz, _p(nodes), _p(s1), _p(b1), _p(b2), _p(b3))
    return Mesh((dx, dy, dz), nodes, coeffs, b1[:n1], b2[:n2 * 2], b3[:n3 * 3])


def cuboid_inside(dims, pad=2) -> np.ndarray:
    """Inside mask of a box room: `pad` layers of non-inside nodes all round
    (pad=2 -> outermost id_none layer, then the boundary shell)."""
    dx, dy, dz = dims
    ins = np.zeros((dz, dy, dx), bool)
    ins[pad:dz - pad, pad:dy - pad, pad:dx - pad] = True
    return ins


# ---- simulation ---------------------------------------------------------------
class Sim:
    def __init__(self, mesh: Mesh, real="double"):
        self.mesh = mesh
        self.mode = {"float": 0, "double": 1}[real]
        dx, dy, dz = mesh.dims
        self._h = lib().wgo_create(dx, dy, dz, _p(mesh.nodes), _p(mesh.coeffs), mesh.coeffs.size,
                                   _p(mesh.b1), mesh.b1.shape[0], _p(mesh.b2), mesh.b2.shape[0],
                                   _p(mesh.b3), mesh.b3.shape[0], self.mode)

    def close(self):
        if self._h:
            lib().wgo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def write(self, node, v):
        lib().wgo_write(self._h, int(node), float(v))

    def read(self, node):
        return lib().wgo_read(self._h, int(node))

    def step(self, n=1) -> int:
        return lib().wgo_step(self._h, int(n))

    def field(self) -> np.ndarray:
        out = np.zeros(self.mesh.num_nodes)
        lib().wgo_field(self._h, _p(out))
        return out

    def set_field(self, f):
        a = np.ascontiguousarray(f, np.float64).reshape(-1)
        assert a.size == self.mesh.num_nodes
        lib().wgo_set_field(self._h, _p(a))

    def run(self, src_node, signal, rcv_nodes, soft=False):
        sig = np.ascontiguousarray(signal, np.float64)
        rcv = np.ascontiguousarray(rcv_nodes, np.uint64)
        out = np.zeros((sig.size, rcv.size))
        flag = C.c_int(0)
        steps = lib().wgo_run(self._h, int(src_node), _p(sig), sig.size, int(bool(soft)),
                              _p(rcv), rcv.size, _p(out), C.byref(flag))
        return int(steps), out, flag.value

    def boundary_data(self, n) -> np.ndarray:
        cnt = lib().wgo_boundary_count(self._h, n)
        out = np.zeros((cnt, n), BDATA_DT)
        if cnt:
            lib().wgo_boundary_data(self._h, n, _p(out))
        return out


def num_threads() -> int:
    return lib().wgo_num_threads()
