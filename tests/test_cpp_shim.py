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


def test_boundary_filter_design_through_the_shim(tmp_path):
    """host code: runs here, no GPU (fitted_boundary.h call sequence vs the golden file)"""
    import json
    exe = build(tmp_path, "test_lrs_shim")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "LRS_SHIM_OK" in r.stdout, (r.returncode, r.stdout, r.stderr)
    rows = [np.array([float(v) for v in line.split()]) for line in r.stdout.splitlines()[:3]]
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "lrs_coefficients.json")))["sets"]
    for row, s in zip(rows, golden[:3]):  # plaster, wood, concrete
        want = np.concatenate([s["reflectance"]["b"], s["reflectance"]["a"], s["impedance"]["b"], s["impedance"]["a"]])
        assert np.abs(row - want).max() < 1e-10


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
def test_gaussian_preprocessor_and_directional_receiver(tmp_path):
    """preprocessor::gaussian (gaussian.cpp:26-53) + postprocessor::directional_receiver
    (directional_receiver.cpp:29-69) through the shim vs a numpy restatement on the
    oracle's field."""
    exe = build(tmp_path, "test_waveguide_shim")
    steps = 40
    c = wgo.to_flat(0.2)
    r = subprocess.run([exe, str(steps), "%.17g" % c["b"][0], "gaussian"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, (r.returncode, r.stderr)
    got = np.array([[float(v) for v in line.split()] for line in r.stdout.strip().splitlines()])
    dims, spacing = (28, 26, 24), np.float32(0.05)
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [c])
    z, y, x = np.indices((dims[2], dims[1], dims[0]))
    pos = [np.float32(0) + v.astype(np.float32) * spacing for v in (x, y, z)]
    d = [pos[0] - np.float32(0.6), pos[1] - np.float32(0.65), pos[2] - np.float32(0.55)]
    ln = np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]).astype(np.float32)
    sdev = np.float32(0.1)
    g = (np.exp(-np.power(ln.astype(np.float64), 2) / (2 * np.power(np.float64(sdev), 2))) /
         np.power(np.float64(sdev) * np.sqrt(2 * np.pi), 3)).astype(np.float32)
    sim = wgo.Sim(om)
    sim.set_field(g.astype(np.float64).ravel())
    rcv = om.index(17, 12, 10)
    nb = [om.index(16, 12, 10), om.index(18, 12, 10), om.index(17, 11, 10), om.index(17, 13, 10),
          om.index(17, 12, 9), om.index(17, 12, 11)]
    vel = np.zeros(3)
    want = []
    for _ in range(steps):
        p = np.float32(sim.read(rcv))
        s = [np.float32((np.float32(sim.read(n)) - p) / np.float64(spacing)) for n in nb]
        m = np.array([(s[1] - s[0]) * 0.5, (s[3] - s[2]) * 0.5, (s[5] - s[4]) * 0.5], np.float64)
        vel -= m / (1.1765 * 11776.0)
        want.append([p] + [np.float32(v * np.float64(p)) for v in vel])
        assert sim.step(1) == 0
    want = np.array(want, np.float64)
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-12)


@pytest.mark.gpu
def test_canonical_single_band_matches_oracle(tmp_path):
    """waveguide::canonical (canonical.h:24-110) through the shim -- per-step callback path and
    the one-call device path, which the executable checks are identical -- against the oracle:
    calibrated unit impulse as a hard source, pressure at the receiver node read as float."""
    exe = build(tmp_path, "test_waveguide_shim")
    c = wgo.to_flat(0.2)
    r = subprocess.run([exe, "0", "%.17g" % c["b"][0], "canonical"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout[-300:], r.stderr)
    lines = r.stdout.strip().splitlines()
    got = np.array([[float(v) for v in line.split()] for line in lines if not line.startswith("#")])
    meta = [line for line in lines if line.startswith("#")][0].split()
    src, rcv, amp = int(meta[2]), int(meta[4]), float(meta[6])
    dims = (28, 26, 24)
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [c])
    # compute_index(descriptor, position): round((p - min_corner) / spacing) in float
    sp = np.float32(0.05)
    loc = lambda p: [int(np.round(np.float32(v) / sp)) for v in p]  # noqa: E731
    assert src == om.index(*loc((0.52, 0.61, 0.48))) and rcv == om.index(*loc((0.86, 0.59, 0.51)))
    # rectilinear_calibration_factor (calibration.h:21-31), stored as float (canonical.h:48-56)
    want_amp = np.float32(np.sqrt(400.0 / (4 * np.pi)) / (0.3405 * np.float64(sp)))
    assert amp == pytest.approx(float(want_amp), rel=1e-7)
    steps = got.shape[0]
    assert steps == 121
    sig = np.zeros(steps)
    sig[0] = float(want_amp)
    done, want, flag = wgo.Sim(om).run(src, sig, [rcv], soft=False)
    assert done == steps and flag == 0
    # printed with 9 significant digits from a float: compare as float32
    assert np.array_equal(got[:, 0].astype(np.float32), want[:, 0].astype(np.float32))
    assert np.abs(got[:, 1:]).max() > 0


@pytest.mark.gpu
def test_raytracer_run_template(tmp_path):
    exe = build(tmp_path, "test_raytracer_shim")
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "RT_SHIM_OK" in r.stdout
    # make_image_source through raytracer::run == the oracle's image-source stage fed by the
    # oracle's own trace of the same directions (same seed, same segments)
    from oracle import rto
    from wayverb_b200 import scene as S
    dirs = np.fromfile(os.path.join(tmp_path, "dirs.f32"), np.float32).reshape(-1, 3)
    got = np.fromfile(os.path.join(tmp_path, "impulses.bin"), rto.IMPULSE_DT)
    sx, sy, sz = 5.56, 3.97, 2.81
    X, Y, Z = (0, sx), (0, sy), (0, sz)
    verts = [(X[i & 1], Y[(i >> 1) & 1], Z[(i >> 2) & 1]) for i in range(8)]
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    tris = []
    for q in quads:
        tris += [(0, q[0], q[1], q[2]), (0, q[0], q[2], q[3])]
    sc = S.Scene(np.array(verts, np.float32), np.array(tris, np.uint32).view(S.TRI_DT).reshape(-1),
                 [S.make_surface(0.1, 0.1)], side=4)
    # the C++ test lists every triangle in every voxel and pads the box by 0.1
    sc.aabb = np.array([-0.1, -0.1, -0.1, sx + 0.1, sy + 0.1, sz + 0.1], np.float32)
    cells = 4 ** 3
    flat = list(range(cells))
    for c in range(cells):
        flat[c] = len(flat)
        flat += [len(tris)] + list(range(len(tris)))
    sc.voxel_index = np.array(flat, np.uint32)
    o = rto.Scene(sc)
    src, rcv = (1, 1, 1), (2, 3, 1.5)
    n = dirs.shape[0]
    elems = np.zeros((4, n), np.uint32)
    seg = 1 << 14
    for b in range(0, n, seg):  # raytracer.h:219-244: segments of 16384 rays, global ray index base
        e = min(n, b + seg)
        # only the first four steps matter here; a ray's path does not depend on the depth limit
        _, refl, _ = o.trace(dirs[b:e], src, rcv, depth=4, total_rays=n, seed=1234, ray_index_base=b,
                             keep_steps=4, specular_from_step=5, n_bins=8)
        elems[:, b:e] = rto.path_elements(refl, 4)
    want, _ = rto.image_source(o, elems, src, rcv)
    assert got.shape == want.shape and np.array_equal(got.view(np.uint8), want.view(np.uint8))


# ---- the shim as an overlay on the reference tree ---------------------------------------------
REFERENCE = "/root/reference/src"
OVERLAY_EXE = os.path.join(ROOT, "tests", "cpp", "_build", "test_overlay")


def build_overlay():
    """tests/cpp/test_overlay.cpp + the reference's UNMODIFIED hard_source.h / soft_source.h /
    postprocessor/node.h / node.cpp, compiled against include/compat. Needs /root/reference,
    so it is built here (build container) and the executable travels to the GPU box
    (tests/cpp/_build is git-ignored, not gpurun-ignored)."""
    src = os.path.join(ROOT, "tests", "cpp", "test_overlay.cpp")
    os.makedirs(os.path.dirname(OVERLAY_EXE), exist_ok=True)
    cmd = [GXX, "-std=c++14", "-O1", "-Wall", "-DWVB_WITH_REFERENCE_HEADERS",
           "-I", os.path.join(ROOT, "include", "compat"), "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(REFERENCE, "waveguide", "include"), "-I", os.path.join(REFERENCE, "utilities", "include"),
           src, os.path.join(REFERENCE, "waveguide", "src", "postprocessor", "node.cpp"),
           "-L", os.path.join(ROOT, "wayverb_b200"), "-l:libwvb200.so", "-Wl,-rpath," + os.path.join(ROOT, "wayverb_b200"),
           "-Wl,-rpath,$ORIGIN/../../../wayverb_b200", "-o", OVERLAY_EXE]
    subprocess.run(cmd, check=True)
    return OVERLAY_EXE


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="/root/reference is only present in the build container")
def test_overlay_compiles_against_unmodified_reference_headers():
    """the reference's own processor headers + node.cpp compile and link against
    include/compat + libwvb200.so; the shim's re-declarations are switched off
    (WVB_WITH_REFERENCE_HEADERS), so there is no ODR clash"""
    from wayverb_b200 import build as b
    b.build_lib()
    exe = build_overlay()
    assert os.path.exists(exe)
    # and without the macro the same translation unit must NOT compile: the static_assert
    # that node::return_type is float (the reference's) guards against silently testing the shim's copy
    r = subprocess.run([GXX, "-std=c++14", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "cpp", "test_overlay.cpp")], capture_output=True, text=True)
    assert r.returncode != 0 and "not the reference's" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("soft", [False, True])
def test_overlay_reference_processors_drive_waveguide_run(soft):
    if os.path.isdir(REFERENCE):
        build_overlay()
    if not os.path.exists(OVERLAY_EXE):
        pytest.skip("tests/cpp/_build/test_overlay was not built (needs /root/reference at build time)")
    steps, c = 60, wgo.to_flat(0.2)
    r = subprocess.run([OVERLAY_EXE, str(steps), "%.17g" % c["b"][0]] + (["soft"] if soft else []),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout[-300:], r.stderr)
    lines = r.stdout.strip().splitlines()
    meta = lines[0].split()
    src, rcv = int(meta[2]), int(meta[4])
    got = np.array([float(v) for v in lines[1:]], np.float32)
    dims = (30, 24, 20)
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [c])
    assert src == om.index(14, 11, 9) and rcv == om.index(19, 13, 8)
    sig = np.zeros(steps, np.float32)
    sig[0] = 1.0
    if soft:
        sig[2] = -0.5
    # what the reference's headers do, step by step, on the oracle: hard_source writes the sample;
    # soft_source reads the node AS cl_float (soft_source.h:21-22), adds, writes; node reads cl_float
    sim = wgo.Sim(om)
    want = []
    for s in range(steps):
        if soft:
            # cl_float + float sample: a float addition (soft_source.h:21-23)
            sim.write(src, float(np.float32(sim.read(src)) + np.float32(sig[s])))
        else:
            sim.write(src, float(sig[s]))
        want.append(np.float32(sim.read(rcv)))
        assert sim.step(1) == 0
    assert np.abs(got).max() > 0
    assert np.array_equal(got, np.array(want, np.float32))


@pytest.mark.gpu
def test_overlay_raytracer_run_with_reference_parameter_types():
    if os.path.isdir(REFERENCE):
        build_overlay()
    if not os.path.exists(OVERLAY_EXE):
        pytest.skip("tests/cpp/_build/test_overlay was not built (needs /root/reference at build time)")
    r = subprocess.run([OVERLAY_EXE, "rt"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout[-300:], r.stderr)
    assert "OVERLAY_RT_OK" in r.stdout
