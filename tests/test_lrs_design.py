"""Boundary-filter design: the library's own Yule-Walker fit
(csrc/lrs_design.cpp, host code -- no GPU involved) and its numpy oracle
(oracle/lrs.py) against

  (1) the nine coefficient sets the reference checked in, produced by the reference
      binary with IT++'s yulewalk (bin/boundary_test/output.soft/coefficients.txt ->
      tests/golden/lrs_coefficients.json; inputs: boundary_test.cpp:302-334);
  (2) each other on random absorptions;
  (3) the reference's own tests: tests/arbitrary_magnitude_filter.cpp:10-44 (every fit
      of a random envelope is stable) and tests/fitted_boundary.cpp:31-40.

Floating point, different least-squares / root-finding routines on each side:
tolerance 1e-10 against the golden values (the reference's own comparison helper
uses 1e-8, tests/fitted_boundary.cpp:27-29) and 1e-9 between product and oracle."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lrs as olrs  # noqa: E402
import wgo  # noqa: E402
from wayverb_b200 import lrs  # noqa: E402

GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "lrs_coefficients.json")))
# boundary_test.cpp:321-334: seven values into an eight-band array (the last band is 0)
MATERIALS = {
    "plaster": [0.08, 0.08, 0.2, 0.5, 0.4, 0.4, 0.36, 0.0],
    "wood": [0.15, 0.15, 0.11, 0.1, 0.07, 0.06, 0.06, 0.0],
    "concrete": [0.02, 0.02, 0.03, 0.03, 0.03, 0.04, 0.07, 0.0],
}
FS = 8000.0  # boundary_test.cpp:302


@pytest.mark.parametrize("k", range(9))
def test_golden_sets_reproduced_by_oracle_and_product(k):
    s = GOLDEN["sets"][k]
    absorption = MATERIALS[s["material"]]
    ob, oa = olrs.reflectance_filter(absorption, FS)
    c = lrs.compute_reflectance_filter_coefficients(absorption, FS)
    for b, a in ((ob, oa), (c["b"], c["a"])):
        assert np.abs(b - s["reflectance"]["b"]).max() < 1e-10
        assert np.abs(a - s["reflectance"]["a"]).max() < 1e-10
    ib, ia = olrs.to_impedance(ob, oa)
    imp = lrs.to_impedance_coefficients(c)
    for b, a in ((ib, ia), (imp["b"], imp["a"])):
        assert np.abs(b - s["impedance"]["b"]).max() < 1e-10
        assert np.abs(a - s["impedance"]["a"]).max() < 1e-10


def test_product_matches_oracle_on_random_absorptions():
    rng = np.random.default_rng(4)
    for _ in range(200):
        absorption = rng.uniform(0.01, 0.95, 8)
        fs = float(rng.choice([4000.0, 8000.0, 16000.0, 44100.0]))
        ob, oa = olrs.reflectance_filter(absorption, fs)
        c = lrs.compute_reflectance_filter_coefficients(absorption, fs)
        scale = max(np.abs(ob).max(), np.abs(oa).max())
        assert np.abs(c["b"] - ob).max() < 1e-9 * scale
        assert np.abs(c["a"] - oa).max() < 1e-9 * scale
        assert lrs.is_stable(c["a"]) and olrs.is_stable(oa)


def test_random_envelopes_give_stable_filters():
    """tests/arbitrary_magnitude_filter.cpp:10-44, with its fixed envelopes first"""
    fixed = [([], []), ([0.0], [0.0]), ([0.0, 0.5], [0.0, 1.0]), ([0.0, 0.5, 0.49], [0.0, 1.0, 0.0]),
             ([0.0, 0.5, 0.49, 0.51], [0.0, 1.0, 0.0, 0.0])]
    for f, a in fixed:
        c = lrs.arbitrary_magnitude_filter(f, a)
        assert lrs.is_stable(c["a"])
    rng = np.random.default_rng(9)
    for i in range(300):
        f = rng.random(100).astype(np.float32).astype(np.float64)
        a = rng.random(100).astype(np.float32).astype(np.float64)
        c = lrs.arbitrary_magnitude_filter(f, a)
        assert lrs.is_stable(c["a"])
        assert np.isfinite(c["b"]).all() and np.isfinite(c["a"]).all()
        if i < 40:
            ob, oa = olrs.arbitrary_magnitude_filter(f, a)
            assert np.abs(c["b"] - ob).max() < 1e-8 and np.abs(c["a"] - oa).max() < 1e-8


def test_fitted_boundary_example_and_magnitude_fit():
    """tests/fitted_boundary.cpp:31-40 only prints; check that the fit follows the envelope"""
    centres = [0.2, 0.4, 0.6, 0.8, 1.0]
    amps = [0.0, 1.0, 0.5, 1.0, 0.0]
    c = lrs.arbitrary_magnitude_filter(centres, amps)
    ob, oa = olrs.arbitrary_magnitude_filter(centres, amps)
    assert np.abs(c["b"] - ob).max() < 1e-9 and np.abs(c["a"] - oa).max() < 1e-9
    w = np.pi * np.array([0.4, 0.6, 0.8])
    z = np.exp(-1j * np.outer(w, np.arange(7)))
    mag = np.abs((z @ c["b"]) / (z @ c["a"]))
    assert np.abs(mag - [1.0, 0.5, 1.0]).max() < 0.2


def test_is_stable_and_flat_coefficients():
    assert lrs.is_stable([1.0]) and lrs.is_stable([1.0, 0.5]) and not lrs.is_stable([1.0, 1.0])
    assert not lrs.is_stable([1.0, -2.5, 1.0])        # roots 2 and 0.5
    assert lrs.is_stable([1.0, -1.0, 0.25])           # double root 0.5
    for a in ([1.0, -0.3, 0.2, 0.1], [1.0, 1.8, 0.9], [1.0, 0.2, -1.3]):
        assert lrs.is_stable(a) == olrs.is_stable(a) == bool((np.abs(np.roots(a)) < 1).all())
    for absorption in (0.0, 0.1, 0.5, 0.99):
        got = lrs.to_flat_coefficients(absorption)
        want = wgo.to_flat(absorption)
        assert np.array_equal(got["b"], want["b"]) and np.array_equal(got["a"], want["a"])


def test_designed_filter_runs_in_the_oracle_waveguide():
    """the designed impedance filter is what wvb_wg_create consumes: a short lossy-box run
    with it must stay finite and decay (CPU oracle only; the GPU parity tests use the golden set)"""
    c = lrs.to_impedance_coefficients(lrs.compute_reflectance_filter_coefficients(MATERIALS["plaster"], FS))
    dims = (16, 14, 12)
    m = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [c])
    sim = wgo.Sim(m)
    sig = np.zeros(400)
    sig[:3] = [1.0, 0.0, -1.0]
    done, trace, flag = sim.run(m.index(8, 7, 6), sig, [m.index(5, 5, 5)], soft=True)
    assert done == 400 and flag == 0 and np.isfinite(trace).all()
    assert np.abs(trace[300:]).max() < np.abs(trace[:100]).max()
