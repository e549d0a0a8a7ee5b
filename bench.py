#!/usr/bin/env python
"""bench.py -- the headline measurement: fp64 waveguide node-updates/s on a
512^3-per-GPU cuboid mesh with 6th-order LRS (plaster) boundaries.

    python bench.py --gpus N --steps K --warmup W            (N=1)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...    the reference's own kernel source on the host CPU cores

A "step" is one pass of the hot path (wayverb's condensed_waveguide launch +
swap, reference waveguide.h:85-123) over the whole mesh. Prints ONE JSON line
on rank 0. See DESIGN.md "Measurement" for every field.

Beside the headline (`value`: weak scaling, 512^3 nodes per GPU) the same line carries
  multi_gpu_parity  N>1: a reduced slab stack stepped by N ranks and by ONE handle on rank 0,
                    compared bit for bit BEFORE anything is timed (the line is not printed if
                    they differ)
  config4           BASELINE config 4's mesh, 512x512x2048, split over the N ranks
  strong            the 512^3 mesh split over the N ranks (strong scaling)
  slab256           N=1: the 512x512x256 slab one GPU of config 4 owns (its efficiency baseline)
  ray               the stochastic ray loop on the concert-hall scene (BASELINE config 5's
                    geometry): ray-reflections/s, sharded over the N ranks + one all-reduce
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# Algorithmic bytes per node-update of the air kernel: previous R+W 16 B,
# current read once 8 B, plus the 4-bit class code this implementation streams instead of the
# reference's 8-byte condensed_node. The layout-faithful 32-B figure of SURVEY 8(d) is reported
# next to it as frac_layout_32B.
BYTES_PER_NODE = 24.5
BYTES_PER_NODE_LAYOUT = 32


def boundary_bytes(counts):
    """Algorithmic bytes of one boundary-kernel launch: per N-d node 96 N B filter state R+W,
    16 B own pressure (previous R + W), 8 B per neighbour read (N inner + 4 / 2 / 0 surrounding),
    8 + 4 N B list entry."""
    total = 0
    for n_dims, cnt in zip((1, 2, 3), counts):
        surround = {1: 4, 2: 2, 3: 0}[n_dims]
        total += cnt * (96 * n_dims + 16 + 8 * (n_dims + surround) + 8 + 4 * n_dims)
    return total


def plaster_coeffs(dtype):
    s = json.load(open(os.path.join(ROOT, "tests", "golden", "lrs_coefficients.json")))["sets"][0]["impedance"]
    c = np.zeros((), dtype)
    c["b"], c["a"] = s["b"], s["a"]
    return c


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def use_all_host_threads() -> int:
    """torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the CPU arms are meant to
    use every host core. Must run before the OpenMP libraries are loaded, and also tells an
    already-loaded libgomp."""
    n = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        C.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except OSError:
        pass
    return n


class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed regions, through NVML
    (nvidia-ml-py) from a polling thread: a 20-step region lasts ~11 ms, too short for an
    `nvidia-smi -lms` child to land a sample in. mark(t0, t1) registers a timed region
    (perf_counter seconds); summary() keeps the samples that fell inside any of them."""

    def __init__(self, index):
        self.index, self.rows, self.windows = index, [], []
        self._stop = threading.Event()
        self.t = None
        self.err = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.err = "NVML unavailable: %s" % e
            return
        self.t = threading.Thread(target=self._poll, daemon=True)
        self.t.start()

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.rows.append((time.perf_counter(), sm, pw, rs))
            except Exception as e:  # noqa: BLE001
                self.err = str(e)
                return
            time.sleep(0.0005)

    def mark(self, t0, t1):
        self.windows.append((t0, t1))

    def summary(self):
        self._stop.set()
        if self.t:
            self.t.join(timeout=2)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": [self.err or "no samples"]}
        nv = self.nv
        names = {
            "hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
            "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
            "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
            "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap,
        }
        inside = [r for r in self.rows if any(a <= r[0] <= b for a, b in self.windows)]
        reasons = sorted(n for n, bit in names.items() if any(r[3] & bit for r in inside))
        return {"sm_mhz": float(np.median([r[1] for r in inside])) if inside else None,
                "sm_max_mhz": self.max_sm,
                "power_w_max": max(r[2] for r in inside) if inside else None,
                "samples": len(inside), "reasons": reasons,
                "how": "NVML polled from a thread; samples inside the timed regions (device-timed steps, "
                       "end-to-end loops, kernels alone)"}


# ---------------------------------------------------------------------------------------------
# CPU legs
# ---------------------------------------------------------------------------------------------
def cpu_baseline(seconds_target=12.0):
    """The oracle (a port of the reference's algorithm, fp64 like the GPU path) on the host
    cores, on a bounded sample of the same workload: a 512x512x32 plaster-walled slice."""
    use_all_host_threads()
    from oracle import wgo
    wgo.use_native()
    dims = (512, 512, 32)
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [plaster_coeffs(wgo.COEFF_DT)])
    sim = wgo.Sim(om, "double")
    sim.write(om.index(256, 256, 16), 1.0)
    sim.step(2)
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < seconds_target:
        sim.step(4)
        n += 4
    dt = time.perf_counter() - t0
    nodes = dims[0] * dims[1] * dims[2]
    return {"value": nodes * n / dt / 1e6, "unit": "Mnode-updates/s", "cores": wgo.num_threads(),
            "kind": "port",
            "sample": "%dx%dx%d plaster-walled slice of the 512^3 mesh, %d fp64 steps, oracle/wg_oracle.cpp "
                      "(-O3 -march=native where it could be built on this host: %s) with OpenMP over all host "
                      "threads" % (dims + (n, wgo.is_native()))}


def run_reference(args, rank, world):
    """--impl reference: the reference's OpenCL host code cannot be built here (no OpenCL
    runtime, no glm / assimp / IT++), but its kernel source can: this arm
    times oracle/_ref -- the reference's own `condensed_waveguide` OpenCL-C source compiled
    for the host, one work-item per node under an OpenMP loop -- in the reference's arithmetic
    (float pressures, double filters) with every host thread. Falls back to the oracle port
    when oracle/_ref is absent. Each step = one kernel pass over a 512x512xNZ slab sample of
    the 512^3 plaster mesh (a rate, so comparable with the full-mesh GPU number)."""
    if rank != 0:
        return
    cores = use_all_host_threads()
    from oracle import refk, wgo

    kind = "reference" if refk.available() else "port"
    if kind == "reference":
        native = refk.use_native_wg()
        what = ("oracle/_ref: the reference's condensed_waveguide kernel source (program.cpp:11-531) compiled "
                "for the host (%s), float pressures + double filters, OpenMP over work-items"
                % ("-O3 -march=native on this host" if native else "portable -O2 build"))
    else:
        wgo.use_native()
        what = "oracle port of condensed_waveguide, float pressures + double filters, OpenMP"

    def make(nz):
        dims = (512, 512, nz)
        om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [plaster_coeffs(wgo.COEFF_DT)])
        sim = refk.Sim(om, "float") if kind == "reference" else wgo.Sim(om, "float")
        sim.write(om.index(256, 256, nz // 2), 1.0)
        return dims, sim

    # size the per-step sample so that K + W steps finish in about 90 s on this host
    dims, sim = make(8)
    sim.step(1)
    t0 = time.perf_counter()
    sim.step(2)
    per_plane = (time.perf_counter() - t0) / 2 / dims[2]
    nz = int(max(8, min(64, 90.0 / max(per_plane * (args.steps + args.warmup), 1e-9))))
    dims, sim = make(nz)
    sim.step(args.warmup)
    t0 = time.perf_counter()
    sim.step(args.steps)
    dt = time.perf_counter() - t0
    nodes = dims[0] * dims[1] * dims[2]
    v = nodes * args.steps / dt / 1e6
    threads = wgo.num_threads()
    sample = "512x512x%d slab sample of the 512^3 plaster mesh per step; %s" % (nz, what)
    line = {
        "impl": "reference", "metric": "Mnode-updates/s (fp64)", "value": v, "unit": "Mnode-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64 filters",
        "data": "synthetic",
        "config": {"workload": "512^3 cuboid, plaster 6th-order LRS walls (sampled: 512x512x%d slab per step; "
                               "a rate, comparable with the full-mesh number)" % nz,
                   "host_threads": threads, "host_cores_visible": cores},
        "cpu_baseline": {"value": v, "unit": "Mnode-updates/s", "cores": threads, "kind": kind,
                         "sample": sample},
        "e2e": {"value": v, "unit": "Mnode-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU legs
# ---------------------------------------------------------------------------------------------
class Ctx:
    """rank / world / collectives of this process"""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None

    def init(self):
        import torch
        self.torch = torch
        assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
        torch.cuda.set_device(self.local)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, x):
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x):
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def fresh_uid(self):
        """an ncclUniqueId serves one communicator: a fresh one per handle, made by rank 0"""
        if not self.dist:
            return None
        import wayverb_b200 as wvb
        box = [wvb.waveguide.nccl_unique_id() if self.rank == 0 else None]
        self.dist.broadcast_object_list(box, src=0)
        return box[0]


HALO_FLAGS = 0  # set from --halo


def slab_handle(ctx, gdims, coeffs, kernel=None):
    """this rank's z-slab of the cuboid mesh `gdims` (balanced contiguous planes)"""
    import wayverb_b200 as wvb
    from wayverb_b200 import _lib
    z0, z1 = wvb.slab_range(gdims[2], ctx.rank, ctx.world)
    lo, hi = max(z0 - 1, 0), min(z1 + 1, gdims[2])
    mesh = wvb.cuboid_mesh(gdims, [coeffs], z0=lo, nz=hi - lo)
    wg = wvb.Waveguide(mesh, device=ctx.local, z_range=(z0, z1), rank=ctx.rank, nranks=ctx.world,
                       nccl_unique_id=ctx.fresh_uid(), kernel=_lib.KERNEL_AUTO if kernel is None else kernel,
                       flags=HALO_FLAGS)
    return mesh, wg, (z0, z1)


def multi_gpu_parity(ctx, coeffs):
    """N ranks step a 512x512x(24 N) plaster mesh (soft source on a seam plane, 12 steps + the
    mesh is only 24 planes thick per rank, so the wave crosses seams); rank 0 then steps the
    SAME mesh with one handle and compares every plane of the gathered field bit for bit."""
    import wayverb_b200 as wvb
    torch, dist = ctx.torch, ctx.dist
    per, steps = 24, 16
    gdims = (512, 512, per * ctx.world)
    mesh, wg, (z0, z1) = slab_handle(ctx, gdims, coeffs)
    src = mesh.index(256, 256, per - 1)            # last plane of rank 0's slab: on a seam
    rcv = [mesh.index(250, 260, per), mesh.index(5, 6, gdims[2] - 4)]
    sig = np.zeros(steps)
    sig[:3] = [1.0, 0.0, -1.0]
    done, out, flag = wg.run_device(src, sig, rcv, soft=True)
    field = torch.from_numpy(wg.field()).cuda()
    wg.close()
    assert done == steps and flag == 0, (done, flag)
    out_t = torch.from_numpy(out).cuda()
    dist.all_reduce(out_t)                          # receivers a rank does not own read 0
    parts = [torch.empty_like(field) for _ in range(ctx.world)] if ctx.rank == 0 else None
    dist.gather(field, parts, dst=0)
    res = {"identical": True}
    if ctx.rank == 0:
        full = wvb.cuboid_mesh(gdims, [coeffs])
        with wvb.Waveguide(full, device=ctx.local) as one:
            d1, out1, f1 = one.run_device(src, sig, rcv, soft=True)
            want = one.field()
        got = torch.cat(parts).cpu().numpy()
        plane = gdims[0] * gdims[1]
        seams = [z for r in range(1, ctx.world) for z in (r * per - 1, r * per)]
        seam_same = all(np.array_equal(got[z * plane:(z + 1) * plane], want[z * plane:(z + 1) * plane])
                        for z in seams)
        same_f = bool(np.array_equal(got, want))
        same_o = bool(np.array_equal(out_t.cpu().numpy(), out1))
        nz_seam = int(sum(np.count_nonzero(want[z * plane:(z + 1) * plane]) for z in seams))
        res = {"identical": same_f and same_o and seam_same and d1 == steps and f1 == 0 and nz_seam > 0,
               "field_identical": same_f, "seam_planes_identical": bool(seam_same),
               "receiver_traces_identical": same_o, "nonzero_seam_values": nz_seam,
               "max_abs_diff": float(np.abs(got - want).max()),
               "mesh": "%dx%dx%d plaster, %d steps, soft source on the rank-0/1 seam" % (gdims + (steps,)),
               "against": "one wvb_wg handle holding the whole mesh on rank 0's GPU (itself bit-compared "
                          "with the CPU oracle by tests/test_wg_gpu.py)"}
    flag_t = torch.tensor([1 if res["identical"] else 0], device="cuda")
    dist.broadcast(flag_t, 0)
    res["identical"] = bool(int(flag_t.item()))
    return res


def time_mesh(ctx, gdims, coeffs, steps, warmup, sampler=None):
    """K device-timed steps of the cuboid mesh `gdims` split over the ranks; max over ranks"""
    mesh, wg, (z0, z1) = slab_handle(ctx, gdims, coeffs)
    src = mesh.index(gdims[0] // 2, gdims[1] // 2, 2)
    wg.write(src, 1.0)
    wg.time_steps(warmup)
    ctx.barrier()
    t0 = time.perf_counter()
    ms, flags = wg.time_steps(steps)
    ctx.barrier()
    if sampler:
        sampler.mark(t0, time.perf_counter())
    assert flags == 0
    info = wg.info()
    wg.close()
    ms = ctx.max(ms) / steps
    nodes = gdims[0] * gdims[1] * gdims[2]
    return {"value": nodes / ms / 1e3, "unit": "Mnode-updates/s", "ms_per_step": ms,
            "mesh": "%dx%dx%d" % gdims, "planes_per_gpu": z1 - z0, "steps": steps,
            "kernel": info["kernel_variant"], "tile": list(info["tile"]), "halo": info["halo"]}


def probe_lib():
    from wayverb_b200 import build as b
    path = b.PROBE
    if not os.path.exists(path):
        path = b.build_probe()
    L = C.CDLL(path)
    L.wvb_probe_waveguide_run.restype = C.c_int
    L.wvb_probe_waveguide_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_uint, C.c_uint,
                                          C.c_size_t, C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    return L


def ray_row(ctx, sampler, with_cpu):
    """The stochastic ray loop (raytracer::run's segment x depth loop with the histogram
    processor) on the concert-hall geometry of BASELINE config 5: 1 M rays in total, split
    over the ranks, scene replicated, histograms summed with one all-reduce."""
    import wayverb_b200 as wvb
    from wayverb_b200 import scene
    from wayverb_b200.slab import ray_range
    sc, meta = scene.concert_hall()
    total = 1 << 20
    depth = wvb.reflection_depth(meta["min_absorption"])
    src, rcv = meta["source"], meta["receiver"]
    b, e = ray_range(total, ctx.rank, ctx.world)
    with wvb.RayTracer(sc, device=ctx.local) as g:
        if ctx.world > 1:
            g.comm_init(ctx.fresh_uid(), ctx.rank, ctx.world)
        # warm-up with the batch size of the timed call (the schedule and its buffers depend on it)
        g.trace(None, src, rcv, depth, n_rays=e - b, total_rays=total, seed=1, ray_index_base=b)
        if ctx.world > 1:
            g.allreduce_histogram()   # the communicator's first collective sets up its connections
        g.reset_histogram()
        ctx.barrier()
        t0 = time.perf_counter()
        _, dropped, ms = g.trace(None, src, rcv, depth, n_rays=e - b, total_rays=total, seed=0x5eed,
                                 ray_index_base=b)
        ctx.barrier()
        t1 = time.perf_counter()
        ar_ms = 0.0
        if ctx.world > 1:
            g.allreduce_histogram()
            ctx.barrier()
            ar_ms = (time.perf_counter() - t1) * 1e3
        h = g.histogram()
        t2 = time.perf_counter()
        sampler.mark(t0, t1)
        # end to end with HOST directions (the iterator range raytracer::run receives): H2D of this
        # rank's 12 B/ray directions, trace, all-reduce, D2H of the histogram
        d = g.directions(0x5eed, e - b, base=b)
        g.reset_histogram()
        ctx.barrier()
        t3 = time.perf_counter()
        g.trace(d, src, rcv, depth, total_rays=total, seed=0x5eed, ray_index_base=b)
        if ctx.world > 1:
            g.allreduce_histogram()
        h2 = g.histogram()
        ctx.barrier()
        t4 = time.perf_counter()
    ms = ctx.max(ms)
    e2e_s = ctx.max(t4 - t3)
    row = {"metric": "ray-reflections/s", "reflections_per_s": total * depth / (ms * 1e-3),
           "e2e": {"value": total * depth / e2e_s, "unit": "ray-reflections/s",
                   "h2d_bytes": int(d.nbytes), "d2h_bytes": int(h2.nbytes),
                   "what": "wvb_rt_trace with host directions + wvb_rt_allreduce_histogram + "
                           "wvb_rt_read_histogram"},
           "kernel_ms": ms, "allreduce_ms": ar_ms, "rays": total, "depth": depth,
           "scene": meta["name"], "triangles": int(sc.triangles.size), "voxel_side": int(sc.side),
           "histogram_bins": int(h.shape[0]), "histogram_energy": float(h.sum()),
           "dropped": int(ctx.sum(dropped)), "sharding": "rays split over %d ranks, scene replicated" % ctx.world}
    if with_cpu and ctx.rank == 0:
        use_all_host_threads()
        from oracle import rto
        n = 1 << 14
        o = rto.Scene(sc)
        dd = rto.directions(5, n)
        t0 = time.perf_counter()
        o.trace(dd, src, rcv, depth, total_rays=total, seed=2)
        dt = time.perf_counter() - t0
        row["cpu_baseline"] = {"value": n * depth / dt, "unit": "ray-reflections/s", "cores": rto.num_threads(),
                               "kind": "port", "sample": "%d rays x %d steps, oracle/rt_oracle.cpp" % (n, depth)}
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dims", default="512,512,512", help="per-GPU slab (x,y,z planes per GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--halo", default="auto", choices=["auto", "nccl", "p2p", "p2p+overlap", "nccl+overlap"],
                    help="ghost-plane transport (auto: peer-to-peer stores when mappable, else NCCL)")
    ap.add_argument("--no-extras", action="store_true",
                    help="headline only: skip config4 / strong / slab256 / ray rows")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    ctx = Ctx()
    if args.impl == "reference":
        run_reference(args, ctx.rank, ctx.world)
        return

    ctx.init()
    import wayverb_b200 as wvb
    from wayverb_b200 import _lib
    rank, world = ctx.rank, ctx.world
    coeffs = plaster_coeffs(_lib.COEFF_DT)
    global HALO_FLAGS
    HALO_FLAGS = {"auto": _lib.HALO_AUTO, "nccl": _lib.HALO_NCCL, "p2p": _lib.HALO_P2P}[args.halo.split("+")[0]] | \
        (_lib.HALO_OVERLAP if args.halo.endswith("+overlap") else 0)

    parity = multi_gpu_parity(ctx, coeffs) if world > 1 else None
    if parity is not None and not parity["identical"]:
        if rank == 0:
            print("multi-GPU parity check FAILED, no bench line: %s" % json.dumps(parity), file=sys.stderr, flush=True)
        sys.exit(1)

    sx, sy, sz = (int(v) for v in args.dims.split(","))
    gdims = (sx, sy, sz * world)
    t_setup = time.perf_counter()
    mesh, wg, (z0, z1) = slab_handle(ctx, gdims, coeffs)
    setup_s = time.perf_counter() - t_setup
    info0 = wg.info()
    nodes_local = sx * sy * sz
    nodes_total = nodes_local * world

    src = mesh.index(sx // 2, sy // 2, sz // 2)  # in rank 0's slab
    wg.write(src, 1.0)

    # ---- kernel-resident timing: K steps, inputs already in HBM -----------------
    sampler = ClockSampler(ctx.local)
    sampler.start()
    wg.time_steps(args.warmup)
    ctx.barrier()
    l0 = wg.info()["kernel_launches"]
    t0 = time.perf_counter()
    ms, flags = wg.time_steps(args.steps)
    ctx.barrier()
    t1 = time.perf_counter()
    sampler.mark(t0, t1)
    wall_ms = (t1 - t0) * 1e3
    launches = wg.info()["kernel_launches"] - l0
    assert flags == 0, "simulation raised error flags 0x%x" % flags

    # ---- end to end through the step-wise C ABI with host buffers ------------------
    # what waveguide::run does per step with hard_source + postprocessor::node:
    # H2D 8 B (source sample), kernel, D2H 4 B flag + 8 B receiver sample.
    rcv = mesh.index(sx // 2 + 5, sy // 2 + 3, sz // 2 - 2)
    e2e_steps = args.steps
    ctx.barrier()
    t0 = time.perf_counter()
    acc = 0.0
    for i in range(e2e_steps):
        wg.write(src, 0.0)
        f = wg.launch()
        acc += wg.read(rcv)
        wg.swap()
        assert f == 0
    ctx.barrier()
    t1 = time.perf_counter()
    sampler.mark(t0, t1)
    e2e_ms = (t1 - t0) * 1e3

    # ---- the whole run in ONE C-ABI call (wvb_wg_run: stock source + receiver on the device):
    # H2D K x 8 B signal, K steps, D2H K x 8 B receiver trace + flag
    sig = np.zeros(args.steps)
    sig[0] = 1.0
    ctx.barrier()
    t0 = time.perf_counter()
    done, trace, f = wg.run_device(src, sig, [rcv])
    ctx.barrier()
    t1 = time.perf_counter()
    sampler.mark(t0, t1)
    assert done == args.steps and f == 0
    run_ms = (t1 - t0) * 1e3

    # ---- dominant kernel alone (roofline) ---------------------------------------------
    ctx.barrier()
    t0 = time.perf_counter()
    k_ms = wg.time_kernels(args.steps)  # [air, boundary]
    ctx.barrier()
    sampler.mark(t0, time.perf_counter())
    wg.close()

    ms = ctx.max(ms)
    wall_ms = ctx.max(wall_ms)
    e2e_ms = ctx.max(e2e_ms)
    run_ms = ctx.max(run_ms)
    air_ms = ctx.max(k_ms[0]) / args.steps
    bnd_ms = ctx.max(k_ms[1]) / args.steps

    # ---- end to end through the reference-facing C++ entry point (N = 1) ---------------
    # waveguide::run<hard_source, callback_accumulator<postprocessor::node>> of the shim, host
    # vectors in, per-step host callbacks: csrc/e2e_probe.cpp
    cpp = None
    if world == 1:
        L = probe_lib()
        out_ms = (C.c_double * 2)()
        chk = C.c_double()
        t0 = time.perf_counter()
        rc = L.wvb_probe_waveguide_run(sx, sy, sz, coeffs.ctypes.data, ctx.local, args.warmup, args.steps,
                                       src, rcv, out_ms, C.byref(chk))
        sampler.mark(t0, time.perf_counter())
        assert rc == 0, "wvb_probe_waveguide_run failed (%d)" % rc
        cpp = {"ms_per_step": out_ms[0] / args.steps, "whole_call_ms": out_ms[1], "checksum": chk.value}

    # ---- BASELINE's other multi-GPU shapes + the ray loop ----------------------------------
    extras = {}

    def extra(name, fn):
        """one optional row; on a single GPU a failure is recorded instead of losing the headline
        (with several ranks an exception on one of them cannot be contained: it propagates)"""
        if world > 1:
            extras[name] = fn()
            return
        try:
            extras[name] = fn()
        except Exception as e:  # noqa: BLE001
            extras[name] = {"error": "%s: %s" % (type(e).__name__, e)}

    if not args.no_extras:
        k = min(args.steps, 200)
        if (sx, sy, sz) == (512, 512, 512):
            def config4():
                r = time_mesh(ctx, (512, 512, 2048), coeffs, k, args.warmup, sampler)
                r["what"] = "BASELINE config 4's mesh split over the %d rank(s)" % world
                return r

            def strong():
                r = time_mesh(ctx, (512, 512, 512), coeffs, k, args.warmup, sampler)
                r["what"] = "the 512^3 mesh split over the %d ranks (strong scaling)" % world
                return r

            def slab256():
                r = time_mesh(ctx, (512, 512, 256), coeffs, k, args.warmup, sampler)
                r["what"] = ("the 256-plane slab one GPU of config 4 owns, alone on one GPU: "
                             "8x this rate is config 4's ideal")
                return r

            extra("config4", config4)
            if world > 1:
                extra("strong", strong)
            else:
                extra("slab256", slab256)
        extra("ray", lambda: ray_row(ctx, sampler, with_cpu=(world == 1 and not args.no_cpu_baseline)))

    if rank == 0:
        peak, peak_src = measured_peak()
        ms_step = ms / args.steps
        value = nodes_total / ms_step / 1e3  # Mnode-updates/s
        air_bytes = nodes_local * BYTES_PER_NODE
        bnd_bytes = boundary_bytes(info0["boundary_nodes"])
        achieved = air_bytes / (air_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp) and (sx, sy, sz) == (512, 512, 512):
            try:
                tj = json.load(open(tp))
                traffic = tj.get("dram_bytes_per_launch")
                traffic_src = tj.get("source", "profiles/traffic.json (ncu --set full capture of this kernel, "
                                               "committed; not re-measured in this run)")
            except Exception:  # noqa: BLE001
                traffic = None
        e2e_step_ms = cpp["ms_per_step"] if cpp else e2e_ms / e2e_steps
        line = {
            "metric": "Mnode-updates/s (fp64)", "value": value, "unit": "Mnode-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": "%dx%dx%d cuboid mesh (%dx%dx%d z-slab per GPU), plaster 6th-order LRS walls "
                            "(BASELINE config 3 at N=1), centred impulse" % (gdims + (sx, sy, sz)),
                "parallelism": "z-slabs x%d, one ghost-plane exchange per face per step (transport: %s)"
                               % (world, info0["halo"]),
                "l2": "no flush needed: the two fp64 pressure arrays are %.2f GB per GPU, far larger than "
                      "the 126 MB L2" % (2 * nodes_local * 8 / 1e9),
                "kernel": info0["kernel_variant"], "tile": list(info0["tile"]),
                "timing": "CUDA events on the library's launch stream, max over ranks",
            },
            "wall_ms_per_step": wall_ms / args.steps,
            "clocks": sampler.summary(),
            "e2e": {"value": nodes_total / e2e_step_ms / 1e3, "unit": "Mnode-updates/s",
                    "h2d_bytes_per_step": 8, "d2h_bytes_per_step": 12,
                    "what": ("the shim's C++ template waveguide::run<hard_source, callback_accumulator<"
                             "postprocessor::node>> (csrc/e2e_probe.cpp): per step write_value 8 B H2D, kernel, "
                             "4 B flag + 8 B read_value D2H, host callbacks on the caller thread") if cpp else
                            ("wvb_wg_write_f64 + wvb_wg_launch + wvb_wg_read_f64 + wvb_wg_swap per step on every "
                             "rank (the calls waveguide::run makes with hard_source + postprocessor::node)"),
                    "c_abi_per_step_calls": nodes_total / (e2e_ms / e2e_steps) / 1e3,
                    "cpp_whole_call_ms": cpp["whole_call_ms"] if cpp else None},
            "e2e_run": {"value": nodes_total / (run_ms / args.steps) / 1e3, "unit": "Mnode-updates/s",
                        "h2d_bytes_per_step": 8, "d2h_bytes_per_step": 8,
                        "what": "one wvb_wg_run call: signal uploaded, stock hard source + node receiver on the "
                                "device, receiver trace downloaded (what canonical() uses when there is no "
                                "per-step host callback)"},
            "gpu_launches": int(launches),
            "roofline": {
                "bound": "hbm", "kernel": "wg_air_%s" % info0["kernel_variant"],
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src,
                "bytes_per_node": BYTES_PER_NODE,
                "bytes_per_node_what": "previous R+W 16 B + current 8 B + 0.5 B class nibble (this "
                                       "implementation's stream instead of the 8-B condensed_node)",
                "frac_layout_32B": nodes_local * BYTES_PER_NODE_LAYOUT / (air_ms * 1e-3) / 1e9 / peak,
                "kernel_ms": air_ms,
                "kernel_share_of_step": air_ms / ms_step,
                "boundary": {"kernel": "wg_boundary_all", "bytes": bnd_bytes, "ms": bnd_ms,
                             "achieved": bnd_bytes / (bnd_ms * 1e-3) / 1e9 if bnd_ms > 0 else None,
                             "frac": bnd_bytes / (bnd_ms * 1e-3) / 1e9 / peak if bnd_ms > 0 else None,
                             "nodes": list(info0["boundary_nodes"])},
                "step_achieved": (air_bytes + bnd_bytes) / (ms_step * 1e-3) / 1e9,
                "step_frac": (air_bytes + bnd_bytes) / (ms_step * 1e-3) / 1e9 / peak,
                "traffic": traffic, "traffic_source": traffic_src,
            },
            "setup_s": setup_s,
            "device_bytes": info0["device_bytes"],
        }
        if parity is not None:
            line["multi_gpu_parity"] = parity
        line.update(extras)
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline()
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"error": "%s: %s" % (type(e).__name__, e)}
        print(json.dumps(line), flush=True)
    if ctx.dist:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
