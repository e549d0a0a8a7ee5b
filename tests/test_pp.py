"""Post-processing of the ray path's histogram (second half).

CPU part: the oracle (oracle/ppo.py) against closed-form properties of the reference's algorithm
(the reference holds no golden vectors for it): the event-rate law of the dirac sequence, energy
bookkeeping of weight_sequence, and the filter bank's partition of unity.
GPU part: wvb_pp_* against the oracle -- dirac sequence and weighted sequence bit-identical, the
FFT-filtered signals within 2e-6 of the peak (float FFT of our own against numpy's double one;
the reference's is FFTW in float)."""
import numpy as np
import pytest

from oracle import ppo

TOL_FFT = 2e-6   # relative to the peak of the expected signal


def histogram(n_bins, seed=3):
    rng = np.random.default_rng(seed)
    t = np.arange(n_bins)[:, None]
    h = np.exp(-t / (n_bins / 6.0)) * (1 + 0.2 * rng.standard_normal((n_bins, 8))) * np.linspace(1.0, 0.3, 8)
    h[:5] = 0
    return np.abs(h) * 1e-6


def test_neg_log_fixed_is_a_logarithm():
    xs = np.concatenate([np.random.default_rng(0).uniform(1e-300, 1.0, 2000), [1.0, 0.5, 2.0 ** -52, 0.70710678118654757]])
    got = np.array([ppo.neg_log_fixed(float(x)) for x in xs])
    want = -np.log(xs)
    assert np.all(np.abs(got - want) <= 4e-16 * np.maximum(1.0, np.abs(want)))
    assert ppo.neg_log_fixed(1.0) == 0.0


def test_dirac_sequence_follows_the_event_rate_law():
    c, vol, rate, T = 340.0, 4.0 * 3.0 * 6.0, 44100.0, 1.0
    seq, events = ppo.dirac_sequence(c, vol, rate, T, seed=11)
    assert seq.size == int(np.ceil(T * rate)) and set(np.unique(seq)) <= {-1.0, 0.0, 1.0}
    k = ppo.constant_mean_event_occurrence(c, vol)
    first = np.flatnonzero(seq)[0] / rate
    assert abs(first - ppo.t0(k)) < 1.0 / rate                       # starts at t0 (postprocessing.cpp:25-27)
    # expected count: integral of min(k t^2, 10000) from t0 to T
    tc = np.sqrt(10000.0 / k)
    expect = k * (min(tc, T) ** 3 - ppo.t0(k) ** 3) / 3 + 10000.0 * max(T - tc, 0)
    assert abs(events - expect) < 5 * np.sqrt(expect)
    # both signs occur about equally often
    nz = seq[seq != 0]
    assert abs((nz > 0).mean() - 0.5) < 0.05
    # another seed, another sequence; same seed, same sequence
    assert not np.array_equal(seq, ppo.dirac_sequence(c, vol, rate, T, seed=12)[0])
    assert np.array_equal(seq, ppo.dirac_sequence(c, vol, rate, T, seed=11)[0])


def test_weight_sequence_carries_each_bins_energy():
    c, vol, rate, hr, Z = 340.0, 72.0, 16000.0, 1000.0, 400.0
    h = histogram(300)
    seq, _ = ppo.dirac_sequence(c, vol, rate, 0.3, seed=5)
    w = ppo.weight_sequence(h, hr, seq, rate, Z)
    assert w.shape == (4800, 8)
    per = int(rate / hr)
    for i in (10, 50, 299):
        seg = w[i * per:(i + 1) * per].astype(np.float64)
        n_ev = np.count_nonzero(seq[i * per:(i + 1) * per])
        if n_ev:
            # sum of p^2 / Z over the bin == the bin's energy (pressure_to_intensity, pressure_intensity.h:9-13)
            assert np.allclose((seg ** 2).sum(0) / Z, h[i], rtol=1e-5)
        else:
            assert not seg.any()


def test_filter_bank_partitions_unity_inside_the_audible_range():
    sr = 44100.0
    e, wf = ppo.band_edges(sr), ppo.width_factor()
    f = np.linspace(e[0] * (1 + wf), e[-1] * (1 - wf), 4000)   # between the outermost transitions
    total = sum(ppo.lopass(f, e[b + 1], wf) * ppo.hipass(f, e[b], wf) for b in range(8))
    assert np.allclose(total, 1.0, atol=1e-12)                       # cos^2 + sin^2 across every crossover
    lo, hi = ppo.lopass(f, 0.1, 0.2), ppo.hipass(f, 0.1, 0.2)
    assert np.allclose(lo + hi, 1.0, atol=1e-12)


def test_crossover_passes_each_signal_in_its_own_band():
    n, sr = 6000, 44100.0
    t = np.arange(n) / sr
    low, high = np.sin(2 * np.pi * 200 * t).astype(np.float32), np.sin(2 * np.pi * 8000 * t).astype(np.float32)
    out = ppo.crossover(low + high, low + high, cutoff=2000 / sr, width=0.2)
    mid = slice(500, n - 500)
    assert np.abs(out[mid] - (low + high)[mid]).max() < 2e-3          # lo + hi == identity
    only = ppo.crossover(low, np.zeros(n, np.float32), 2000 / sr, 0.2)
    assert np.abs(only[mid] - low[mid]).max() < 2e-3
    gone = ppo.crossover(high, np.zeros(n, np.float32), 2000 / sr, 0.2)
    assert np.abs(gone[mid]).max() < 2e-3
    w = ppo.crossover(low, high, 2000 / sr, 0.2, window_length=300)
    assert w[0] == 0 and np.allclose(w[300:], ppo.crossover(low, high, 2000 / sr, 0.2)[300:])


# ---- GPU parity ---------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("vol,rate,T,seed", [(72.0, 44100.0, 1.0, 11), (33 * 15 * 50.0, 16000.0, 2.5, 7), (8.0, 8000.0, 0.2, 1)])
def test_device_dirac_sequence_is_the_oracles(vol, rate, T, seed):
    from wayverb_b200 import postprocess as pp
    got, ev = pp.dirac_sequence(340.0, vol, rate, T, seed=seed)
    want, ev_o = ppo.dirac_sequence(340.0, vol, rate, T, seed=seed)
    assert ev == ev_o and ev > 0
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_device_stochastic_postprocessing_matches_oracle():
    from wayverb_b200 import postprocess as pp
    h = histogram(700)
    for rate, vol in ((44100.0, 72.0), (16000.0, 24750.0)):
        got, got_w = pp.stochastic(h, 1000.0, rate, vol, seed=9, return_weighted=True)
        want, want_w = ppo.stochastic(h, 1000.0, rate, vol, seed=9)
        assert got_w.shape == want_w.shape and np.array_equal(got_w, want_w)      # exact: no FFT involved yet
        assert got.shape == want.shape == (int(700 * rate / 1000.0),)
        assert np.abs(want).max() > 0
        assert np.abs(got - want).max() <= TOL_FFT * np.abs(want).max()


@pytest.mark.gpu
def test_device_multiband_and_crossover_match_oracle():
    from wayverb_b200 import postprocess as pp
    rng = np.random.default_rng(4)
    m = rng.standard_normal((5000, 8)).astype(np.float32)
    want = ppo.multiband_mixdown(m, 44100.0)
    got = pp.multiband_mixdown(m, 44100.0)
    assert np.abs(got - want).max() <= TOL_FFT * np.abs(want).max()
    lo, hi = rng.standard_normal(7000).astype(np.float32), rng.standard_normal(4100).astype(np.float32)
    for cutoff, width, win in ((0.05, 0.2, 0), (0.11, 0.2, 333), (0.02, 0.0, 10)):
        want = ppo.crossover(lo, hi, cutoff, width, win)
        got = pp.crossover(lo, hi, cutoff, width, win)
        assert got.shape == want.shape == (7000,)
        assert np.abs(got - want).max() <= TOL_FFT * np.abs(want).max()
    assert pp.crossover(np.zeros(0, np.float32), hi, 0.05).shape == (4100,)


@pytest.mark.gpu
def test_histogram_of_a_real_trace_becomes_a_signal():
    """ray loop -> histogram -> post-processing, end to end on the device, against the oracle chain"""
    import wayverb_b200 as wvb
    from wayverb_b200 import postprocess as pp, scene
    from oracle import rto
    sc = scene.box_scene((4.0, 3.0, 6.0), subdiv=2, side=8, surfaces=[scene.make_surface(0.1, 0.1)])
    src, rcv, n, depth = (1.1, 1.2, 1.3), (3.0, 2.0, 4.5), 20000, 30
    d = rto.directions(5, n)
    with wvb.RayTracer(sc) as g:
        g.trace(d, src, rcv, depth, seed=5, n_bins=400)
        h = g.histogram()
    want_h, _, _ = rto.Scene(sc).trace(d, src, rcv, depth, seed=5, n_bins=400)
    got = pp.stochastic(h, 1000.0, 22050.0, 72.0, seed=3)
    want, _ = ppo.stochastic(want_h, 1000.0, 22050.0, 72.0, seed=3)
    assert np.abs(want).max() > 0
    # the histograms agree to 1e-9 of the largest bin (fp64 atomics); the signal inherits that
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
