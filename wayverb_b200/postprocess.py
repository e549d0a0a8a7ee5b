"""Host mirror of the reference's post-processing of the ray path over the C ABI (wvb_pp_*):

  dirac_sequence(...)        raytracer::stochastic::generate_dirac_sequence (postprocessing.cpp:29-50)
  stochastic(...)            raytracer::stochastic::postprocessing          (postprocessing.cpp:57-112)
  multiband_mixdown(...)     core::multiband_filter_and_mixdown             (core/mixdown.h:17-24)
  crossover(...)             combined::crossover_filter + left_hanning      (combined/postprocess.h:33-60,104-134)
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import PpParams, check, lib, ptr


def _params(speed_of_sound, acoustic_impedance, room_volume, histogram_rate, output_rate, max_time, seed, device):
    p = PpParams()
    p.speed_of_sound, p.acoustic_impedance, p.room_volume = float(speed_of_sound), float(acoustic_impedance), float(room_volume)
    p.histogram_sample_rate, p.output_sample_rate, p.max_time = float(histogram_rate), float(output_rate), float(max_time)
    p.seed, p.device = int(seed), int(device)
    return p


def dirac_sequence(speed_of_sound, room_volume, sample_rate, max_time, seed=1, device=0):
    """-> (sequence float32[ceil(max_time * sample_rate)], events drawn)"""
    p = _params(speed_of_sound, 400.0, room_volume, 1000.0, sample_rate, max_time, seed, device)
    n, ev = C.c_uint64(), C.c_uint32()
    check(lib().wvb_pp_dirac_sequence(C.byref(p), float(sample_rate), float(max_time), None, 0, C.byref(n), C.byref(ev)))
    out = np.zeros(n.value, np.float32)
    check(lib().wvb_pp_dirac_sequence(C.byref(p), float(sample_rate), float(max_time), ptr(out), out.size, C.byref(n),
                                      C.byref(ev)))
    return out, ev.value


def stochastic(histogram, histogram_rate, output_rate, room_volume, speed_of_sound=340.0, acoustic_impedance=400.0,
               max_time=0.0, seed=1, device=0, return_weighted=False):
    h = np.ascontiguousarray(histogram, np.float64).reshape(-1, 8)
    p = _params(speed_of_sound, acoustic_impedance, room_volume, histogram_rate, output_rate, max_time, seed, device)
    n = C.c_uint64()
    check(lib().wvb_pp_stochastic(ptr(h), h.shape[0], C.byref(p), None, 0, C.byref(n), None))
    out = np.zeros(n.value, np.float32)
    w = np.zeros((n.value, 8), np.float32) if return_weighted else None
    check(lib().wvb_pp_stochastic(ptr(h), h.shape[0], C.byref(p), ptr(out), out.size, C.byref(n),
                                  ptr(w) if w is not None else None))
    return (out, w) if return_weighted else out


def multiband_mixdown(multiband, sample_rate, device=0):
    m = np.ascontiguousarray(multiband, np.float32).reshape(-1, 8)
    out = np.zeros(m.shape[0], np.float32)
    check(lib().wvb_pp_multiband_mixdown(ptr(m), m.shape[0], float(sample_rate), int(device), ptr(out)))
    return out


def crossover(lo, hi, cutoff, width=0.2, window_length=0, device=0):
    a = np.ascontiguousarray(lo, np.float32).reshape(-1)
    b = np.ascontiguousarray(hi, np.float32).reshape(-1)
    out = np.zeros(max(a.size, b.size), np.float32)
    check(lib().wvb_pp_crossover(ptr(a) if a.size else None, a.size, ptr(b) if b.size else None, b.size, float(cutoff),
                                 float(width), int(window_length), int(device), ptr(out), out.size))
    return out
