"""Pins the closed-form HOST functions -- in the shim (include/wayverb_b200/*.hpp), behind the C ABI
(csrc/lrs_design.cpp, wvb_rt_ray_energy, wvb_rt_reflection_depth) and in the oracles (oracle/lrs.py,
oracle/wg_oracle.cpp's coefficient helpers) -- to the REFERENCE'S OWN source for them:
waveguide/src/mesh_descriptor.cpp, config.cpp, calibration.h, fitted_boundary.h with
arbitrary_magnitude_filter.h / frequency_domain_envelope.cpp / stable.h, filters.cpp,
raytracer/optimum_reflection_number.h, compute_ray_energy of stochastic/finder.{h,cpp} -- compiled
unmodified from /root/reference into oracle/_ref (oracle/ref_recipe/build.py, hostmath_driver.inc).
All host code: runs without a GPU.

IT++ (itpp::yulewalk) is not in the image. The stand-in of hoststubs/itpp/ does not fit anything: it
records the grid the reference hands to the fit and forwards to a callback, here the oracle's
restatement of the published method. So the reference's design pipeline runs as written AROUND the
fit, and the fit itself stays pinned where it already was: the reference's nine checked-in
coefficient sets (tests/test_lrs_design.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lrs as olrs  # noqa: E402
import wgo  # noqa: E402
from oracle import refk, rto  # noqa: E402
from wayverb_b200 import lrs  # noqa: E402

pytestmark = pytest.mark.skipif(not refk.available(), reason="no /root/reference and no prebuilt oracle/_ref")

GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
MATERIALS = [[0.08, 0.08, 0.2, 0.5, 0.4, 0.4, 0.36, 0.0], [0.15, 0.15, 0.11, 0.1, 0.07, 0.06, 0.06, 0.0],
             [0.02, 0.02, 0.03, 0.03, 0.03, 0.04, 0.07, 0.0]]


def test_shim_and_c_abi_host_functions_equal_the_references(tmp_path):
    """tests/cpp/test_hostmath_pin.cpp: compute_index / locator / position / neighbors / sample_rate,
    rectilinear_calibration_factor, compute_optimum_reflection_number (value and scene forms),
    wvb_rt_ray_energy, to_impedance / to_flat coefficients, is_stable -- exact equality on tens of
    thousands of random inputs."""
    exe = str(tmp_path / "test_hostmath_pin")
    lib_dir, ref_dir = os.path.join(ROOT, "wayverb_b200"), os.path.join(ROOT, "oracle", "_ref")
    subprocess.run([GXX, "-std=c++14", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(ROOT, "tests", "cpp", "test_hostmath_pin.cpp"), "-L" + lib_dir, "-lwvb200",
                    "-Wl,-rpath," + lib_dir, "-L" + ref_dir, "-l:lib_ref.so", "-Wl,-rpath," + ref_dir], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "HOSTMATH_PIN_OK" in r.stdout, (r.returncode, r.stdout, r.stderr)


def fit(order, f, m):
    return olrs.yulewalk(order, list(f), list(m))


@pytest.mark.parametrize("fs", [8000.0, 16000.0, 44100.0])
def test_reflectance_filter_pipeline_around_the_fit(fs):
    rng = np.random.default_rng(int(fs))
    cases = MATERIALS + [list(rng.uniform(0.01, 0.95, 8)) for _ in range(12)]
    for absorption in cases:
        rb, ra, f256, m256, threw = refk.hm_reflectance_filter(absorption, fs, fit)
        # the envelope the reference hands to yulewalk == the oracle's, bit for bit
        centres = olrs.band_centres_hz() / fs * 2
        of, om = olrs.envelope_grid(centres, np.sqrt(1 - np.asarray(absorption, float)))
        assert np.array_equal(f256, of) and np.array_equal(m256, om)
        try:
            ob, oa = olrs.reflectance_filter(absorption, fs)
        except RuntimeError:
            assert threw
            continue
        assert not threw
        # the same fit in the reference's pipeline and in the oracle's -> identical coefficients
        assert np.array_equal(rb, ob) and np.array_equal(ra, oa)
        # and the product's own pipeline + fit within its stated 1e-9
        c = lrs.compute_reflectance_filter_coefficients(absorption, fs)
        assert np.abs(c["b"] - rb).max() < 1e-9 and np.abs(c["a"] - ra).max() < 1e-9
        # impedance conversion: reference == oracle (numpy) == oracle (C) == product
        ib, ia = refk.hm_to_impedance(rb, ra)
        nb, na = olrs.to_impedance(rb, ra)
        assert np.array_equal(ib, nb) and np.array_equal(ia, na)
        refl = np.zeros((), wgo.COEFF_DT)
        refl["b"], refl["a"] = rb, ra
        w = wgo.to_impedance(refl)
        p = lrs.to_impedance_coefficients(refl)
        assert np.array_equal(w["b"], ib) and np.array_equal(w["a"], ia)
        assert np.array_equal(p["b"], ib) and np.array_equal(p["a"], ia)


def test_unstable_fit_makes_the_reference_throw():
    def bad_fit(order, f, m):
        a = np.zeros(order + 1)
        a[0], a[-1] = 1.0, 1.5                         # |reflection coefficient| >= 1 (stable.h:44-47)
        return np.ones(order + 1), a
    *_, threw = refk.hm_reflectance_filter(MATERIALS[0], 8000.0, bad_fit)
    assert threw


def test_flat_coefficients_and_stability():
    rng = np.random.default_rng(2)
    for absorption in [0.0, 0.1, 0.5, 0.999] + list(rng.uniform(0, 0.999, 200)):
        fb, fa = refk.hm_to_flat(absorption)
        w = wgo.to_flat(absorption)
        p = lrs.to_flat_coefficients(absorption)
        assert np.array_equal(w["b"], fb) and np.array_equal(w["a"], fa)
        assert np.array_equal(p["b"], fb) and np.array_equal(p["a"], fa)
    for k in range(3000):
        a = rng.uniform(-1, 1, 7) * (0.2 if k % 3 == 0 else 1.0)
        a[0] = 1.0
        assert olrs.is_stable(a) == refk.hm_is_stable(a)


def test_peak_biquads_and_their_convolution():
    """filters.cpp:10-33 against the waveguide oracle's helpers (used by the biquad == canonical KAT,
    boundary filters of the reference's tests/waveguide.cpp)"""
    rng = np.random.default_rng(3)
    for _ in range(500):
        gain, centre, q = rng.uniform(-60, 0), rng.uniform(0.001, 0.49), rng.uniform(0.1, 5.0)
        b, a = refk.hm_peak_biquad(gain, centre, q)
        got = wgo.peak_biquad(gain, centre, q)
        assert np.array_equal(got[:3], b) and np.array_equal(got[3:], a)
    for _ in range(200):
        bq = np.concatenate([np.concatenate(refk.hm_peak_biquad(rng.uniform(-60, 0), rng.uniform(0.001, 0.49),
                                                                rng.uniform(0.1, 5.0))) for _ in range(3)])
        cb, ca = refk.hm_convolve3(bq)
        got = wgo.convolve3(bq)
        assert np.array_equal(got["b"], cb) and np.array_equal(got["a"], ca)


def test_ray_energy_reflection_number_and_rates_of_the_oracle():
    rng = np.random.default_rng(4)
    for _ in range(2000):
        s, r = rng.uniform(0, 10, 3).astype(np.float32), rng.uniform(0, 10, 3).astype(np.float32)
        radius, rays = float(np.float32(rng.uniform(0.05, 0.5))), int(rng.integers(1000, 3_000_000))
        assert np.float32(rto.ray_energy(rays, s, r, radius)) == np.float32(refk.hm_ray_energy(rays, s, r, radius))
    from wayverb_b200 import raytracer
    for a in list(rng.uniform(0.005, 0.98, 2000)) + [0.25, 0.1, 0.5]:
        assert raytracer.reflection_depth(a) == refk.hm_reflection_number(a)
    for h, c in ((0.05, 340.0), (0.0123, 343.2)):
        rate, dt, h_back, c_back = refk.hm_rates(h, c)
        assert rate == 1 / (float(np.float32(h)) / (c * np.sqrt(3.0)))
        assert abs(h_back - np.float32(h)) < 1e-12 and abs(c_back - c) < 1e-9
