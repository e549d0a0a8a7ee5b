"""Turns gpurun_out/*.ncu-rep and the launch list CSV into the small text/JSON
summaries committed under profiles/ (run in the build container, no GPU needed).

usage: summarise_ncu.py <round-tag> <air.ncu-rep> <boundary.ncu-rep> <launches.csv>
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "sm__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in KEEP:
                d[h] = (r[i], units[i])
        res.append(d)
    return res


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(v) * mult


def to_s(v, unit):
    mult = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9}[unit]
    return float(v) * mult


def main():
    tag, air, bnd, launches = sys.argv[1:5]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    lines = []
    traffic = {}
    for name, rep in (("air", air), ("boundary", bnd)):
        for d in raw(rep):
            lines.append("== %s: %s" % (name, d["kernel"]))
            for k in KEEP:
                if k in d:
                    lines.append("   %-86s %s %s" % (k, d[k][0], d[k][1]))
            rb = to_bytes(*d["dram__bytes_read.sum"])
            wb = to_bytes(*d["dram__bytes_write.sum"])
            t = to_s(*d["gpu__time_duration.sum"])
            lines.append("   -> DRAM traffic %.3f GB in %.1f us = %.0f GB/s" % ((rb + wb) / 1e9, t * 1e6, (rb + wb) / t / 1e9))
            if name == "air":
                traffic = {"kernel": d["kernel"].split("(")[0], "dram_bytes_per_launch": rb + wb,
                           "dram_read_bytes": rb, "dram_write_bytes": wb, "ncu_time_us": t * 1e6,
                           "workload": "512^3 fp64, plaster LRS", "algorithmic_bytes_32B": 512 ** 3 * 32,
                           "source": os.path.basename(rep)}
    open(os.path.join(ROOT, "profiles", tag + "_ncu_summary.txt"), "w").write("\n".join(lines) + "\n")
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    # launch list: per-kernel average duration and share of the step
    rows = [r for r in csv.reader(open(launches)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        agg.setdefault(r[4].split("(")[0].replace("void ", ""), []).append(float(r[-1]))
    unit = rows[0][-2]
    tot = sum(sum(v) for v in agg.values())
    out = ["launch list: %d launches captured (ncu --metrics gpu__time_duration.sum --clock-control none); "
           "cold-cache serialised times: compare SHARES" % len(rows)]
    for k, v in agg.items():
        out.append("%-48s n=%3d avg=%10.1f %s  share=%.3f" % (k, len(v), sum(v) / len(v), unit, sum(v) / tot))
    open(os.path.join(ROOT, "profiles", tag + "_launches.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(lines[-12:]))
    print("\n".join(out))


if __name__ == "__main__":
    main()
