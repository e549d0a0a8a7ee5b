"""Host mirror of the reference's boundary-filter design helpers over the C ABI
(fitted_boundary.h:20-104, arbitrary_magnitude_filter.h:63-95, stable.h:11-50)."""
from __future__ import annotations

import numpy as np

from ._lib import COEFF_DT, check, lib, ptr


def arbitrary_magnitude_filter(frequency, amplitude) -> np.ndarray:
    f = np.ascontiguousarray(frequency, np.float64)
    a = np.ascontiguousarray(amplitude, np.float64)
    out = np.zeros((), COEFF_DT)
    check(lib().wvb_lrs_arbitrary_magnitude_filter(ptr(f), ptr(a), f.size, ptr(out)))
    return out


def compute_reflectance_filter_coefficients(absorption, sample_rate) -> np.ndarray:
    a = np.ascontiguousarray(absorption, np.float64)
    assert a.size == 8
    out = np.zeros((), COEFF_DT)
    check(lib().wvb_lrs_reflectance_filter(ptr(a), float(sample_rate), ptr(out)))
    return out


def to_impedance_coefficients(c) -> np.ndarray:
    c = np.ascontiguousarray(c, COEFF_DT)
    out = np.zeros((), COEFF_DT)
    lib().wvb_lrs_to_impedance(ptr(c), ptr(out))
    return out


def to_flat_coefficients(absorption) -> np.ndarray:
    out = np.zeros((), COEFF_DT)
    lib().wvb_lrs_flat(float(absorption), ptr(out))
    return out


def is_stable(a) -> bool:
    a = np.ascontiguousarray(a, np.float64)
    return bool(lib().wvb_lrs_is_stable(ptr(a), a.size))
