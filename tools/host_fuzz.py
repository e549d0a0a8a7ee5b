"""Fuzzes the HOST-only entry points of the library (wvb_obj_parse, wvb_voxelise, wvb_lrs_*) built with
AddressSanitizer + UndefinedBehaviorSanitizer by tools/host_sanitize.sh: damaged OBJ files, degenerate /
non-finite triangle soups at every octree depth, out-of-range absorptions and envelopes. Any report of
either sanitizer goes to stderr; the run prints three progress lines when nothing was found."""
import os
import ctypes as C, random, sys, numpy as np
L = C.CDLL(os.environ.get("WVB_HOST_ASAN_LIB", "gpurun_out/asan/libhost_asan.so"))
u64 = C.c_uint64
L.wvb_obj_parse.argtypes = [C.c_char_p, u64, C.c_void_p, C.POINTER(u64), C.c_void_p, C.POINTER(u64), C.c_void_p, C.POINTER(u64)]
L.wvb_voxelise.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, u64, C.POINTER(u64)]
L.wvb_lrs_reflectance_filter.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
L.wvb_lrs_arbitrary_magnitude_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
def p(a): return a.ctypes.data_as(C.c_void_p)
def parse(raw):
    nv, nt, nn = u64(), u64(), u64()
    s = L.wvb_obj_parse(raw, len(raw), None, C.byref(nv), None, C.byref(nt), None, C.byref(nn))
    if s: return None
    v = np.zeros((nv.value, 4), np.float32); t = np.zeros((nt.value, 4), np.uint32); names = C.create_string_buffer(max(nn.value, 1))
    s = L.wvb_obj_parse(raw, len(raw), p(v), C.byref(nv), p(t), C.byref(nt), names, C.byref(nn))
    return (v, t) if s == 0 else None
def vox(v, t, depth, pad):
    lo = np.zeros(3, np.float32); hi = np.zeros(3, np.float32); cnt = u64()
    s = L.wvb_voxelise(p(v), v.shape[0], p(t), t.shape[0], depth, pad, p(lo), p(hi), None, 0, C.byref(cnt))
    if s: return None
    out = np.zeros(cnt.value, np.uint32)
    s = L.wvb_voxelise(p(v), v.shape[0], p(t), t.shape[0], depth, pad, p(lo), p(hi), p(out), cnt.value, C.byref(cnt))
    return out
rnd = random.Random(3)
lines = open("tests/golden/concert_hall.obj").read().splitlines()
words = ["1", "-1", "0", "99999", "1/2/3", "a", "1e400", "-", "//", "4294967296", "-4294967297", "nan", "inf"]
ok = 0
for it in range(1500):
    ls = list(lines)
    for _ in range(rnd.randint(0, 8)):
        i = rnd.randrange(len(ls)); r = rnd.random()
        if r < 0.2: del ls[i]
        elif r < 0.4: ls[i] = ls[i][:rnd.randrange(len(ls[i]) + 1)]
        elif r < 0.6: ls[i] = rnd.choice(["v", "f", "vt", "vn", "usemtl", "g", "#", "l"]) + " " + " ".join(rnd.choice(words) for _ in range(rnd.randint(0, 6)))
        elif r < 0.8 and ls[i]:
            b = bytearray(ls[i].encode("latin1")); b[rnd.randrange(len(b))] = rnd.randrange(256); ls[i] = b.decode("latin1")
        else: ls.insert(i, ls[rnd.randrange(len(ls))])
    got = parse("\n".join(ls).encode("latin1"))
    if got is not None:
        ok += 1
        if it % 10 == 0:
            vox(got[0], got[1], rnd.choice([0, 1, 3, 5]), rnd.choice([0.1, 0.0, -1.0, float("nan")]))
print("parsed", ok)
# voxeliser on soups incl. degenerate
rng = np.random.default_rng(0)
for it in range(60):
    n = int(rng.integers(1, 400))
    v = np.zeros((3 * n, 4), np.float32); v[:, :3] = rng.uniform(-4, 4, (3 * n, 3)) * rng.choice([1.0, 1e-3, 1e6])
    if it % 7 == 0: v[rng.integers(0, 3 * n), rng.integers(0, 3)] = rng.choice([np.nan, np.inf, -np.inf])
    if it % 5 == 0: v[3:6] = v[0:3]
    t = np.zeros((n, 4), np.uint32); t[:, 1] = np.arange(n) * 3; t[:, 2] = t[:, 1] + 1; t[:, 3] = t[:, 1] + 2
    if it % 9 == 0: t[0, 3] = t[0, 1]
    vox(v, t, int(rng.integers(0, 7)), float(rng.choice([0.1, 0.0, 1.0])))
print("voxelised")
for it in range(400):
    a = rng.uniform(-0.2, 1.2, 8); fs = float(rng.choice([100.0, 8000.0, 44100.0, 1e6, 0.0, -1.0]))
    out = np.zeros(14); L.wvb_lrs_reflectance_filter(p(a), fs, p(out))
    n = int(rng.integers(0, 40)); f = np.sort(rng.uniform(-0.5, 1.5, n)); m = rng.uniform(-1, 2, n)
    if it % 11 == 0 and n: m[rng.integers(0, n)] = np.nan
    L.wvb_lrs_arbitrary_magnitude_filter(p(f), p(m), n, p(out))
print("designed")
