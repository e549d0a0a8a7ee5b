"""Builds libwvb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

The built .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("WVB_LIB_OUT") or os.path.join(HERE, "libwvb200.so")
SOURCES = ["wg_host.cu", "rt_host.cu", "mesh_host.cu", "is_host.cu", "lrs_design.cpp", "scene_host.cpp", "pp_host.cu"]
HEADERS = ["common.h", "nccl_dyn.h", "wg_kernels.cuh", "rt_kernels.cuh", "mesh_kernels.cuh", "is_kernels.cuh", "pp_kernels.cuh",
           os.path.join("..", "..", "include", "wvb200.h")]

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    # keep the reference's operation order: no FMA contraction, IEEE div/sqrt
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-shared",
]


def _nvcc() -> str:
    for c in (os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc"), shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [_nvcc()]
    if os.path.exists("/usr/bin/g++"):
        cmd += ["-ccbin", "/usr/bin/g++"]
    extra = os.environ.get("WVB_NVCC_DEFS", "").split()  # e.g. "-DRT_MIN_BLOCKS=8" for experiments
    cmd += NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs + ["-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


PROBE = os.path.join(HERE, "libwvb200_probe.so")


def build_probe(force: bool = False) -> str:
    """bench.py's end-to-end leg through the C++ `waveguide::run` template (csrc/e2e_probe.cpp):
    host-only C++14, linked against libwvb200.so (found next to it at run time)."""
    src = os.path.join(CSRC, "e2e_probe.cpp")
    deps = [src, os.path.join(HERE, "..", "include", "wayverb_b200", "waveguide.hpp"), LIB]
    if not force and os.path.exists(PROBE) and all(os.path.getmtime(d) <= os.path.getmtime(PROBE) for d in deps):
        return PROBE
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [gxx, "-std=c++14", "-O2", "-fPIC", "-shared", "-o", PROBE, src, "-L" + HERE, "-l:libwvb200.so",
           "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return PROBE


if __name__ == "__main__":
    import sys
    print(build_lib(force=True, verbose="-v" in sys.argv))
    print(build_probe(force=True))
