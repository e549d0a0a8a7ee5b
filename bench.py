#!/usr/bin/env python
"""bench.py -- the headline measurement: fp64 waveguide node-updates/s on a
512^3-per-GPU cuboid mesh with 6th-order LRS (plaster) boundaries.

    python bench.py --gpus N --steps K --warmup W            (N=1)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...    the reference's algorithm on the host CPU cores

A "step" is one pass of the hot path (wayverb's condensed_waveguide launch +
swap, reference waveguide.h:85-123) over the whole mesh. Prints ONE JSON line
on rank 0. See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

BYTES_PER_NODE = 32  # prev R+W 16 B, current 8 B, condensed_node 8 B
BYTES_PER_NODE_STENCIL = 24


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


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.
    nvidia-smi needs ~1 s before its first line, so it is started before the warm-up and
    every line is stamped with the host clock; stop(t0, t1) keeps the samples that fell
    inside the timed region [t0, t1] (perf_counter seconds)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def wait_first_sample(self, timeout=5.0):
        t_end = time.perf_counter() + timeout
        while not self.rows and time.perf_counter() < t_end and self.proc:
            time.sleep(0.02)

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            if len(r) < 7 or not (t0 <= ts <= t1):
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                pw.append(float(r[2]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(seconds_target=12.0):
    """The oracle (a port of the reference's algorithm) on the host cores, on a
    bounded sample of the same workload: a 512x512x32 plaster-walled slice."""
    from oracle import wgo
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
                      "with OpenMP over all host threads" % (dims + (n,))}


def run_reference(args, rank, world):
    """--impl reference: the reference's OpenCL path cannot run here (no OpenCL CPU
    runtime), so this arm times the oracle port of its algorithm in
    the reference's own arithmetic types (float pressures, double filters) with
    every host thread. Each step = one kernel pass over a 512x512x64 slab sample."""
    if rank != 0:
        return
    from oracle import wgo

    def make(nz):
        dims = (512, 512, nz)
        om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [plaster_coeffs(wgo.COEFF_DT)])
        sim = wgo.Sim(om, "float")
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
    sample = ("512x512x%d slab sample of the 512^3 plaster mesh per step; oracle port of "
              "condensed_waveguide, float pressures + double filters, OpenMP" % nz)
    line = {
        "impl": "reference", "metric": "Mnode-updates/s (fp64)", "value": v, "unit": "Mnode-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64 filters",
        "data": "synthetic",
        "config": {"workload": "512^3 cuboid, plaster 6th-order LRS walls (sampled: 512x512x%d slab per step)" % nz},
        "cpu_baseline": {"value": v, "unit": "Mnode-updates/s", "cores": wgo.num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": v, "unit": "Mnode-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dims", default="512,512,512", help="per-GPU slab (x,y,z planes per GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import wayverb_b200 as wvb
    from wayverb_b200 import _lib

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    uid = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        box = [wvb.waveguide.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sx, sy, sz = (int(v) for v in args.dims.split(","))
    gdims = (sx, sy, sz * world)
    z0, z1 = rank * sz, (rank + 1) * sz
    lo, hi = max(z0 - 1, 0), min(z1 + 1, gdims[2])
    t_setup = time.perf_counter()
    mesh = wvb.cuboid_mesh(gdims, [plaster_coeffs(_lib.COEFF_DT)], z0=lo, nz=hi - lo)
    wg = wvb.Waveguide(mesh, device=local_rank, z_range=(z0, z1), rank=rank, nranks=world, nccl_unique_id=uid)
    setup_s = time.perf_counter() - t_setup
    info0 = wg.info()
    nodes_local = sx * sy * sz
    nodes_total = nodes_local * world

    src = mesh.index(sx // 2, sy // 2, sz // 2)  # in rank 0's slab
    wg.write(src, 1.0)

    # ---- kernel-resident timing: K steps, inputs already in HBM -----------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    wg.time_steps(args.warmup)
    sampler.wait_first_sample()
    barrier()
    l0 = wg.info()["kernel_launches"]
    t0 = time.perf_counter()
    ms, flags = wg.time_steps(args.steps)
    barrier()
    t1 = time.perf_counter()
    wall_ms = (t1 - t0) * 1e3
    launches = wg.info()["kernel_launches"] - l0
    clocks = sampler.stop(t0, t1)
    assert flags == 0, "simulation raised error flags 0x%x" % flags

    # ---- end to end through the step-wise C ABI with host buffers ------------------
    # what waveguide::run does per step with hard_source + postprocessor::node:
    # H2D 8 B (source sample), kernel, D2H 4 B flag + 8 B receiver sample.
    rcv = mesh.index(sx // 2 + 5, sy // 2 + 3, sz // 2 - 2)
    e2e_steps = args.steps
    barrier()
    t0 = time.perf_counter()
    acc = 0.0
    for i in range(e2e_steps):
        wg.write(src, 0.0)
        f = wg.launch()
        acc += wg.read(rcv)
        wg.swap()
        assert f == 0
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3

    # ---- dominant kernel alone (roofline) ---------------------------------------------
    k_ms = wg.time_kernels(args.steps)  # [air, boundary]
    barrier()

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms = reduce_max(ms)
    wall_ms = reduce_max(wall_ms)
    e2e_ms = reduce_max(e2e_ms)
    air_ms = reduce_max(k_ms[0]) / args.steps
    bnd_ms = reduce_max(k_ms[1]) / args.steps

    if rank == 0:
        peak, peak_src = measured_peak()
        ms_step = ms / args.steps
        value = nodes_total / ms_step / 1e3  # Mnode-updates/s
        achieved = nodes_local * BYTES_PER_NODE / (air_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:  # noqa: BLE001
                traffic = None
        line = {
            "metric": "Mnode-updates/s (fp64)", "value": value, "unit": "Mnode-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": "%dx%dx%d cuboid mesh (%dx%dx%d z-slab per GPU), plaster 6th-order LRS walls "
                            "(BASELINE config 3 at N=1), centred impulse" % (gdims + (sx, sy, sz)),
                "parallelism": "z-slabs x%d, one NCCL ghost-plane send/recv per face per step" % world,
                "l2": "no flush needed: the two fp64 pressure arrays are %.2f GB per GPU, far larger than "
                      "the 126 MB L2" % (2 * nodes_local * 8 / 1e9),
                "kernel": info0["kernel_variant"], "tile": list(info0["tile"]),
                "timing": "CUDA events on the library's launch stream, max over ranks",
            },
            "wall_ms_per_step": wall_ms / args.steps,
            "clocks": clocks,
            "e2e": {"value": nodes_total / (e2e_ms / e2e_steps) / 1e3, "unit": "Mnode-updates/s",
                    "h2d_bytes_per_step": 8, "d2h_bytes_per_step": 12,
                    "what": "wvb_wg_write_f64 + wvb_wg_launch + wvb_wg_read_f64 + wvb_wg_swap per step "
                            "(the calls waveguide::run makes with hard_source + postprocessor::node)"},
            "gpu_launches": int(launches),
            "roofline": {
                "bound": "hbm", "kernel": "wg_air_%s" % info0["kernel_variant"],
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src,
                "bytes_per_node": BYTES_PER_NODE,
                "achieved_24B": nodes_local * BYTES_PER_NODE_STENCIL / (air_ms * 1e-3) / 1e9,
                "kernel_ms": air_ms, "boundary_kernels_ms": bnd_ms,
                "kernel_share_of_step": air_ms / ms_step,
                "step_achieved": nodes_local * BYTES_PER_NODE / (ms_step * 1e-3) / 1e9,
                "step_frac": nodes_local * BYTES_PER_NODE / (ms_step * 1e-3) / 1e9 / peak,
                "traffic": traffic,
            },
            "setup_s": setup_s,
            "device_bytes": info0["device_bytes"],
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    wg.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
