"""The check on the stand-ins: the REFERENCE'S OWN unit tests, compiled unmodified on top of them.

oracle/ref_recipe/ compiles the reference's sources for the host behind stand-ins for what this image
lacks (GLM, FFTW, IT++, cl.hpp, googletest), and the test_ref_pin_*.py files hold the oracles against
that build. Whether the stand-ins themselves behave is answered by the reference: its own test files for
the code on and around the path (src/core/tests, src/raytracer/tests, src/waveguide/tests,
src/frequency_domain/tests -- geometry, tri/cube intersection, indexing, recursive_vector, the reflector
against its CPU twin, image sources against the exact shoebox solution, multitree, histograms, BRDF,
"does the program build", the biquad cascade against the canonical filter through the reference's filter
kernels, the filter bank on noise, convolution, and the tests that run on the reference's
own models -- voxel walk / flatten / surrounded / compare on the vault, the mesh fixtures on the tunnel
and the bedroom, bad reflections in the vault; assimp's loader stood in for by the library's OBJ reader)
are built by
oracle/ref_recipe/build_tests.py and must pass (its six-minute nan_in_waveguide case -- a 56-million-node
fitted-wall room stepped 432 times -- runs on request only; profiles/r02_reference_own_tests.txt). What is left out, and the two cases the
reference contradicts itself on, are listed in that file's docstring."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("_wvb_ref_tests", os.path.join(ROOT, "oracle", "ref_recipe", "build_tests.py"))
bt = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(bt)

BUILT = bt.build()
pytestmark = pytest.mark.skipif(BUILT is None, reason="no /root/reference and no prebuilt oracle/_ref/reftest_*")

EXPECTED_CASES = {"core": 34, "raytracer": 14, "waveguide": 6, "frequency_domain": 2, "models": 9}


@pytest.mark.parametrize("group", sorted(EXPECTED_CASES))
def test_the_references_own_tests_pass_on_the_stand_ins(group):
    results = bt.run(group)
    assert len(results) == EXPECTED_CASES[group]
    # Most of these tests draw their inputs from std::random_device; the ones known to fail now and then
    # are listed (with their rates) in build_tests.py. Whatever fails is run again: a case that the code
    # does not satisfy fails every time, a statistical one passes within a few attempts.
    for attempt in range(14):
        retry = [c for c, ok in results.items() if not ok and c not in bt.KNOWN_STALE]
        if not retry:
            break
        again = bt.run(group)
        for c in retry:
            results[c] = again[c]
    failed = {c for c, ok in results.items() if not ok}
    assert failed == (bt.KNOWN_STALE & set(results)), failed
