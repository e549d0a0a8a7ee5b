"""CPU restatement (numpy) of the reference's post-processing of the ray path's histogram.
TEST INFRASTRUCTURE ONLY: imported by tests/ alone.

  generate_dirac_sequence   src/raytracer/src/stochastic/postprocessing.cpp:16-50
                            (+ interval_size, include/raytracer/stochastic/postprocessing.h:38-44)
  weight_sequence           postprocessing.cpp:57-97, core/pressure_intensity.h:15-21
  multiband_filter          src/frequency_domain/include/frequency_domain/multiband_filter.h:49-93,
                            src/frequency_domain/src/filter.cpp:22-47 (FFTW r2c -> callback -> c2r, / n),
                            src/frequency_domain/src/envelope.cpp:5-110, src/hrtf/lib/include/hrtf/multiband.h:10-44
  mixdown                   src/core/include/core/mixdown.h:12-24
  crossover_filter + window src/combined/include/combined/postprocess.h:33-60,104-134, core/sinc.h:62-82

Where the reference is not reproducible the same stated replacements as in the product are used:
  * the engine seeded from std::random_device becomes Philox4x32-10(seed; k, 0, 7, 0) -> 53-bit
    uniform u_k, x = 1 - u_k in (0, 1] (the reference's uniform_real_distribution{1.0, 0.0});
  * log(1 / x) is evaluated by a fixed series (neg_log_fixed) in plain double arithmetic, and
    pow(t, 2.0) as t * t, so that the event times do not depend on a libm.
Parity pinning: PINNED to reference-run output. oracle/ref_recipe/build.py compiles the reference's own
postprocessing.cpp, frequency_domain library, hrtf/multiband.h, core/sinc.h and crossover_filter,
unmodified, behind an FFTW stand-in (hoststubs/fftw3.h) and with the engine's seed chosen;
tests/test_ref_pin_pp.py asserts this file bit-identical to that build for the rate law, the event
loop (dirac_sequence(intervals=...): the reference's arithmetic on the reference engine's numbers),
weight_sequence, band edges, magnitudes and the window, and within float rounding (5e-7 of the peak)
for everything that passes through a transform. On top: the reference's own property tests restated
in tests/test_pp.py (band magnitudes sum to 1 across the crossover, energy splits between bands).
"""
from __future__ import annotations

import math

import numpy as np

M32 = np.uint64(0xffffffff)


def philox(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 on uint64-held 32-bit lanes (vectorised over c0); returns (o0, o1)."""
    c0 = np.asarray(c0, np.uint64)
    c1 = np.full_like(c0, c1)
    c2 = np.full_like(c0, c2)
    c3 = np.full_like(c0, c3)
    k0, k1 = np.uint64(k0), np.uint64(k1)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c0
        p1 = np.uint64(0xCD9E8D57) * c2
        h0, l0 = p0 >> np.uint64(32), p0 & M32
        h1, l1 = p1 >> np.uint64(32), p1 & M32
        n0 = h1 ^ c1 ^ k0
        n2 = h0 ^ c3 ^ k1
        c0, c1, c2, c3 = n0, l1, n2, l0
        k0 = (k0 + np.uint64(0x9E3779B9)) & M32
        k1 = (k1 + np.uint64(0xBB67AE85)) & M32
    return c0, c1


def neg_log_fixed(x: float) -> float:
    m, e = math.frexp(x)
    if m < 0.70710678118654752440:
        m = m * 2.0
        e -= 1
    s = (m - 1.0) / (m + 1.0)
    z = s * s
    p = 1.0 / 25.0
    for k in range(23, 0, -2):
        p = p * z + 1.0 / float(k)
    lm = 2.0 * s * p
    le = float(e) * 0.693147180369123816490 + float(e) * 1.90821492927058770002e-10
    return -(le + lm)


def exponentials(seed: int, n: int) -> np.ndarray:
    o0, o1 = philox(np.arange(n, dtype=np.uint64), 0, 7, 0, seed & 0xffffffff, (seed >> 32) & 0xffffffff)
    u = ((o0 >> np.uint64(5)).astype(np.float64) * 67108864.0 + (o1 >> np.uint64(6)).astype(np.float64)) \
        * (1.0 / 9007199254740992.0)
    return np.array([neg_log_fixed(float(1.0 - v)) for v in u])


def constant_mean_event_occurrence(speed_of_sound, room_volume):
    return 4 * math.pi * math.pow(speed_of_sound, 3.0) / room_volume


def t0(constant):
    return math.pow(2.0 * math.log(2.0) / constant, 1.0 / 3.0)


def dirac_sequence(speed_of_sound, room_volume, sample_rate, max_time, seed=1, intervals=None):
    """-> (float32 sequence, events drawn)

    intervals=None: the stated replacements (Philox uniforms, fixed -log series, t * t).
    intervals=array: the unit-rate draws log(1 / x) are taken from the caller and pow(t, 2.0) from
    libm -- the reference's own arithmetic, so that tests/test_ref_pin_pp.py can hold this loop
    against the reference's generate_dirac_sequence run with the same engine."""
    c = constant_mean_event_occurrence(speed_of_sound, room_volume)
    ret = np.zeros(int(math.ceil(max_time * sample_rate)), np.float32)
    ref_arith = intervals is not None
    exps = intervals if ref_arith else exponentials(seed, int(min(1.5 * 10000.0 * max_time + 4096.0, 2.0e9)))
    t, k = t0(c), 0
    while t < max_time:
        sample_index = t * sample_rate
        twice = int(2 * sample_index)
        ret[int(sample_index)] = -1.0 if twice % 2 else 1.0
        mean = min(c * (math.pow(t, 2.0) if ref_arith else t * t), 10000.0)
        t += exps[k] / mean
        k += 1
    return ret, k


def weight_sequence(histogram, hist_rate, sequence, seq_rate, acoustic_impedance):
    h = np.asarray(histogram, np.float64).reshape(-1, 8).astype(np.float32)   # stored as float bands
    seq = np.asarray(sequence, np.float32)

    def convert(ind):
        return int(ind * seq_rate / hist_rate)
    ideal = convert(h.shape[0])
    n = min(seq.size, ideal)
    ret = np.repeat(seq[:n, None], 8, 1).astype(np.float32)
    for i in range(h.shape[0]):
        beg, end = min(convert(i), n), min(convert(i + 1), n)
        ss = np.float32(np.sum(seq[beg:end].astype(np.float32) ** 2, dtype=np.float32))
        if ss != 0:
            q = (h[i] / ss).astype(np.float32)
            scale = np.copysign(np.sqrt(np.abs(q.astype(np.float64) * acoustic_impedance)), q.astype(np.float64))
        else:
            scale = np.zeros(8)
        ret[beg:end] = (ret[beg:end].astype(np.float64) * scale).astype(np.float32)
    return ret


# ---- envelopes (envelope.cpp) ---------------------------------------------------------------------
def band_edges(sample_rate, bands=8, lo=20.0, hi=20000.0):
    return np.array([lo * math.pow(hi / lo, i / bands) for i in range(bands + 1)]) / sample_rate


def width_factor(bands=8, lo=20.0, hi=20000.0, overlap=1.0):
    base = math.pow(hi / lo, 1.0 / bands)
    return (base - 1) / (base + 1) * overlap


def lopass(f, edge, wf):
    f = np.asarray(f, np.float64)
    aw = edge * wf
    out = np.zeros_like(f)
    out[f < edge - aw] = 1.0
    mid = (f >= edge - aw) & (f < edge + aw)
    if aw > 0:
        out[mid] = np.cos(np.pi * (((f[mid] - edge) / aw) + 1) / 2 / 2) ** 2
    else:
        out[mid] = ((f[mid] - edge) < 0).astype(np.float64)
    return out


def hipass(f, edge, wf):
    f = np.asarray(f, np.float64)
    aw = edge * wf
    out = np.ones_like(f)
    out[f < edge - aw] = 0.0
    mid = (f >= edge - aw) & (f < edge + aw)
    if aw > 0:
        out[mid] = np.sin(np.pi * (((f[mid] - edge) / aw) + 1) / 2 / 2) ** 2
    else:
        out[mid] = (0 <= (f[mid] - edge)).astype(np.float64)
    return out


def fft_length(n):
    return int(math.pow(2, math.ceil(math.log2(n)))) << 2


def _filter(x, mag):
    """frequency_domain::filter::run: zero-pad, r2c, scale bin i by mag(i / n), c2r, / n, truncate"""
    x = np.asarray(x, np.float32)
    n = fft_length(x.size) if not isinstance(mag, tuple) else mag[1]
    f = (np.arange(n // 2 + 1, dtype=np.float32) / np.float32(n)).astype(np.float64)
    m = (mag[0] if isinstance(mag, tuple) else mag)(f)
    spec = np.fft.rfft(np.concatenate([x.astype(np.float64), np.zeros(n - x.size)]))
    return np.fft.irfft(spec * m.astype(np.float32), n)[:x.size].astype(np.float32)


def multiband_mixdown(multiband, sample_rate):
    m = np.asarray(multiband, np.float32).reshape(-1, 8)
    e, wf = band_edges(sample_rate), width_factor()
    out = np.zeros(m.shape[0], np.float32)
    for b in range(8):
        out = (out + _filter(m[:, b], lambda f, b=b: lopass(f, e[b + 1], wf) * hipass(f, e[b], wf))).astype(np.float32)
    return out


def stochastic(histogram, hist_rate, output_rate, room_volume, speed_of_sound=340.0, acoustic_impedance=400.0,
               max_time=0.0, seed=1):
    h = np.asarray(histogram, np.float64).reshape(-1, 8)
    mt = max_time if max_time > 0 else h.shape[0] / hist_rate
    seq, _ = dirac_sequence(speed_of_sound, room_volume, output_rate, mt, seed)
    w = weight_sequence(h, hist_rate, seq, output_rate, acoustic_impedance)
    return multiband_mixdown(w, output_rate), w


def left_hanning(length):
    i = np.arange(length, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2 * np.pi * (i / (2 * (length - 1.0))))).astype(np.float32)


def crossover(lo, hi, cutoff, width=0.2, window_length=0):
    lo = np.asarray(lo, np.float32)
    hi = np.asarray(hi, np.float32)
    n = fft_length(max(lo.size, hi.size))
    a = _filter(lo, (lambda f: lopass(f, cutoff, width), n)) if lo.size else lo
    b = _filter(hi, (lambda f: hipass(f, cutoff, width), n)) if hi.size else hi
    out = np.zeros(max(a.size, b.size), np.float32)
    out[:a.size] += a
    out[:b.size] = (out[:b.size] + b).astype(np.float32)
    w = min(window_length, out.size)
    if w:
        out[:w] = left_hanning(w) * out[:w]
    return out
