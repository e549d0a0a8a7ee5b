"""The C++ drop-in shim (include/wayverb_b200/*.hpp): it must compile as C++14
against the C ABI everywhere, and on a GPU `waveguide::run` driven exactly like
the reference's own run_waveguide test must reproduce the oracle."""
import os
import subprocess

import numpy as np
import pytest

from oracle import wgo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def build(tmp_path, name):
    exe = str(tmp_path / name)
    lib_dir = os.path.join(ROOT, "wayverb_b200")
    cmd = [GXX, "-std=c++14", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
           "-o", exe, os.path.join(ROOT, "tests", "cpp", name + ".cpp"),
           "-L" + lib_dir, "-lwvb200", "-Wl,-rpath," + lib_dir]
    subprocess.run(cmd, check=True)
    return exe


def test_shim_compiles_as_cpp14(tmp_path):
    build(tmp_path, "test_waveguide_shim")
    build(tmp_path, "test_raytracer_shim")


@pytest.mark.gpu
def test_waveguide_run_template_matches_oracle(tmp_path):
    exe = build(tmp_path, "test_waveguide_shim")
    steps = 90
    c = wgo.to_flat(0.1)
    r = subprocess.run([exe, str(steps), "%.17g" % c["b"][0]], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stderr)
    got = np.array([[float(v) for v in line.split()] for line in r.stdout.strip().splitlines()])
    dims = (30, 24, 40)
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [c])
    sig = np.zeros(steps)
    sig[0] = 1.0
    rcv = [om.index(15, 12, z) for z in (12, 18, 24, 30)]
    done, want, flag = wgo.Sim(om).run(om.index(15, 12, 8), sig, rcv, soft=True)
    assert done == steps and flag == 0
    assert got.shape == want.shape
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_raytracer_run_template(tmp_path):
    exe = build(tmp_path, "test_raytracer_shim")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "RT_SHIM_OK" in r.stdout
