"""Pins the post-processing oracle (oracle/ppo.py) to the REFERENCE'S OWN host source:
raytracer/src/stochastic/postprocessing.cpp (whole file) with core/cl/traits.h, core/mixdown.h,
core/pressure_intensity.h; the frequency_domain library (envelope.cpp, filter.cpp, plan.cpp,
buffer.cpp, traits.cpp, multiband_filter.h); hrtf/multiband.h; core/sinc.h; crossover_filter of
combined/postprocess.h:33-60 -- compiled unmodified from /root/reference into oracle/_ref by
oracle/ref_recipe/build.py. Two stand-ins decide arithmetic and both are stated:
  * FFTW is not in the image: hoststubs/fftw3.h evaluates r2c / c2r in double and rounds once, so
    everything that passes through a transform is compared at float rounding (5e-7 of the peak);
  * generate_dirac_sequence seeds its engine from std::random_device: the build spells that as a
    device returning a chosen seed, and the oracle is run in its reference-arithmetic mode on the
    numbers that engine produces (libm log and pow instead of the product's Philox + fixed series).
Everything else -- the event-rate law, the event loop, weight_sequence's float / double mix, the
band edges, the crossover magnitudes, the half window -- is asserted bit for bit."""
import math

import numpy as np
import pytest

from oracle import ppo, refk

pytestmark = pytest.mark.skipif(not refk.available(), reason="no /root/reference and no prebuilt oracle/_ref")

ROOMS = [(340.0, 1000.0, 44100.0, 0.5), (343.0, 30.0, 48000.0, 1.0), (340.0, 20000.0, 16000.0, 2.0),
         (331.0, 5.0, 8000.0, 0.3)]


@pytest.mark.parametrize("c,volume,rate,max_time", ROOMS)
def test_event_rate_law(c, volume, rate, max_time):
    k = ppo.constant_mean_event_occurrence(c, volume)
    for t in (1e-4, 0.01, 0.1, max_time, 10.0):
        ref = refk.pp_rate_law(c, volume, t)
        assert ref[0] == k and ref[2] == ppo.t0(k)
        assert ref[1] == min(k * math.pow(t, 2.0), 10000.0)


def test_interval_draws_are_the_logarithm_of_the_engines_uniforms():
    iv, xs = refk.pp_intervals(11, 4096)
    assert np.all((xs > 0) & (xs <= 1))
    assert np.array_equal(iv, np.array([math.log(1.0 / x) for x in xs]))   # libm's log, as std::log


@pytest.mark.parametrize("seed,room", list(enumerate(ROOMS, 3)))
def test_dirac_sequence_loop_is_the_references(seed, room):
    c, volume, rate, max_time = room
    ref = refk.pp_dirac_sequence(seed, c, volume, rate, max_time)
    intervals, _ = refk.pp_intervals(seed, int(10000 * max_time * 1.5) + 4096)
    got, events = ppo.dirac_sequence(c, volume, rate, max_time, intervals=intervals)
    assert events > 100 and np.count_nonzero(ref) > 100
    assert np.array_equal(ref, got)
    # the product-mode sequence (Philox, fixed series) obeys the same law: same length, same
    # first event (t0 is deterministic), event count within 5 sigma of the reference draw's
    own, own_events = ppo.dirac_sequence(c, volume, rate, max_time, seed=seed)
    assert own.size == ref.size
    assert np.flatnonzero(own)[0] == np.flatnonzero(ref)[0] and own[np.flatnonzero(own)[0]] == ref[np.flatnonzero(ref)[0]]
    assert abs(own_events - events) < 5 * np.sqrt(2 * events)


@pytest.mark.parametrize("hist_rate", [1000.0, 997.0, 2000.0, 44100.0])
def test_weight_sequence_bit_for_bit(hist_rate):
    rng = np.random.default_rng(int(hist_rate))
    h = rng.uniform(0, 1e-3, (int(0.2 * hist_rate) + 3, 8))
    h[5] = 0
    h[7, 3] = -1e-4                      # copysign branch of intensity_to_pressure
    h[9] = 1e-30
    seq, _ = ppo.dirac_sequence(340.0, 1000.0, 44100.0, 0.25)
    ref = refk.pp_weight_sequence(h, hist_rate, seq, 44100.0, 400.0)
    got = ppo.weight_sequence(h, hist_rate, seq, 44100.0, 400.0)
    assert ref.shape == got.shape and np.count_nonzero(ref) > 1000
    assert np.array_equal(ref, got)


def test_weight_sequence_truncates_to_the_histogram():
    h = np.full((10, 8), 1e-4)
    seq = np.ones(5000, np.float32)
    ref = refk.pp_weight_sequence(h, 100.0, seq, 44100.0, 400.0)
    got = ppo.weight_sequence(h, 100.0, seq, 44100.0, 400.0)
    assert ref.shape == got.shape == (4410, 8) and np.array_equal(ref, got)
    short = np.ones(100, np.float32)       # sequence shorter than the histogram covers
    assert np.array_equal(refk.pp_weight_sequence(h, 100.0, short, 44100.0, 400.0),
                          ppo.weight_sequence(h, 100.0, short, 44100.0, 400.0))


@pytest.mark.parametrize("rate", [8000.0, 44100.0, 96000.0])
def test_band_parameters_and_magnitudes_bit_for_bit(rate):
    edges, wf = refk.pp_band_params(rate)
    assert np.array_equal(edges, ppo.band_edges(rate)) and wf == ppo.width_factor()
    f = np.concatenate([np.linspace(0, 0.5, 1201), edges, edges * (1 - wf), edges * (1 + wf)])
    for b in range(8):
        lo_o, hi_o = ppo.lopass(f, edges[b + 1], wf), ppo.hipass(f, edges[b], wf)
        for i, x in enumerate(f):
            lo, _, _ = refk.pp_magnitudes(x, edges[b + 1], edges[b + 1], wf)
            _, hi, _ = refk.pp_magnitudes(x, edges[b], edges[b + 1], wf)
            assert lo == lo_o[i] and hi == hi_o[i]
            assert refk.pp_magnitudes(x, edges[b], edges[b + 1], wf)[2] == lo_o[i] * hi_o[i]
    # the width-0 step functions of envelope.cpp:30-32,40-42
    assert refk.pp_magnitudes(0.1, 0.1, 0.2, 0.0)[:2] == (0.0, 1.0)
    assert ppo.lopass(np.array([0.1]), 0.1, 0.0)[0] == 0.0 and ppo.hipass(np.array([0.1]), 0.1, 0.0)[0] == 1.0


@pytest.mark.parametrize("n", [1, 2, 300, 1000, 4096, 4097])
def test_fft_length(n):
    assert refk.pp_fft_length(n) == ppo.fft_length(n)


@pytest.mark.parametrize("n,rate", [(300, 16000.0), (1000, 44100.0), (4097, 48000.0)])
def test_filter_bank_and_mixdown(n, rate):
    m = np.random.default_rng(n).standard_normal((n, 8)).astype(np.float32)
    ref = refk.pp_multiband_mixdown(m, rate)
    got = ppo.multiband_mixdown(m, rate)
    assert np.abs(ref).max() > 1
    assert np.abs(ref - got).max() <= 5e-7 * np.abs(ref).max()


@pytest.mark.parametrize("n_lo,n_hi", [(1000, 1500), (2048, 2048), (700, 0), (0, 300)])
def test_crossover_filter(n_lo, n_hi):
    rng = np.random.default_rng(n_lo + n_hi)
    lo, hi = rng.standard_normal(n_lo).astype(np.float32), rng.standard_normal(n_hi).astype(np.float32)
    ref = refk.pp_crossover(lo, hi, 0.1, 0.2)
    got = ppo.crossover(lo, hi, 0.1, 0.2)
    assert ref.shape == got.shape == (max(n_lo, n_hi),)
    assert np.abs(ref - got).max() <= 5e-7 * max(np.abs(ref).max(), 1.0)


@pytest.mark.parametrize("length", [2, 3, 17, 1000])
def test_left_hanning_bit_for_bit(length):
    assert np.array_equal(refk.pp_left_hanning(length), ppo.left_hanning(length))


def test_whole_stochastic_postprocessing():
    """postprocessing() (postprocessing.cpp:99-111) = weight_sequence -> filter bank -> mixdown"""
    rng = np.random.default_rng(5)
    h = rng.uniform(0, 1e-3, (200, 8)) * np.exp(-np.arange(200) / 40.0)[:, None]
    seq, _ = ppo.dirac_sequence(340.0, 1000.0, 44100.0, 0.2)
    ref = refk.pp_postprocessing(h, 1000.0, seq, 44100.0, 400.0)
    got, _ = ppo.stochastic(h, 1000.0, 44100.0, 1000.0, 340.0, 400.0, 0.2)
    assert ref.shape == got.shape and np.abs(ref).max() > 0.01
    assert np.abs(ref - got).max() <= 5e-7 * np.abs(ref).max()
