"""oracle/lrs.py -- TEST INFRASTRUCTURE ONLY (see wg_oracle.cpp's header).

numpy restatement of the reference's locally-reacting-surface filter DESIGN
(SURVEY 8f rank 4), the step that turns a surface's 8-band absorption into the
`coefficients_canonical` the waveguide step consumes. Paths relative to
/root/reference:

  compute_reflectance_filter_coefficients   src/waveguide/include/waveguide/fitted_boundary.h:79-104
  arbitrary_magnitude_filter<N>             src/waveguide/include/waveguide/arbitrary_magnitude_filter.h:63-95
  frequency_domain_envelope                 src/waveguide/src/frequency_domain_envelope.cpp:27-62
  interp / linear_interp                    src/core/include/core/cosine_interp.h:17-76
  band centres                              src/hrtf/lib/include/hrtf/multiband.h:11-20,
                                            src/frequency_domain/src/envelope.cpp:49-56
  is_stable                                 src/waveguide/include/waveguide/stable.h:11-50
  to_impedance_coefficients                 src/waveguide/include/waveguide/fitted_boundary.h:20-50

The fit itself is `itpp::yulewalk` (arbitrary_magnitude_filter.h:84-92): IT++ is
fetched at configure time (unpinned HEAD, config/dependencies.cmake:136-140) and is
absent from /root/reference. Its yulewalk is the modified Yule-Walker design of
Friedlander & Porat as published in the MATLAB signal toolbox (`yulewalk.m`:
512-point magnitude grid, Hamming-tapered 4N autocorrelation lags, least-squares
denominator, polystab, additive decomposition, cepstral minimum-phase numerator).
That published algorithm is restated below and PINNED: it reproduces all nine
coefficient sets the reference checked in
(bin/boundary_test/output.soft/coefficients.txt, produced by the reference binary
with IT++) to 3e-13 (tests/test_lrs_design.py).
"""
import numpy as np

ORDER = 6
AUDIBLE = (20.0, 20000.0)


def band_centres_hz(bands=8, lo=AUDIBLE[0], hi=AUDIBLE[1]):
    # band_centre_frequency(band, bands, r) = band_edge_frequency(2 band + 1, 2 bands, r)
    return np.array([lo * (hi / lo) ** ((2 * b + 1) / (2.0 * bands)) for b in range(bands)])


def is_stable(a):
    """Schur-Cohn recursion on the denominator, as stable.h:11-50 writes it"""
    a = [float(v) for v in a]
    while len(a) > 1:
        rci = a[-1]
        if 1 <= abs(rci):
            return False
        size = len(a) - 1
        a = [(a[i] - a[size - i] * rci) / (1 - rci * rci) for i in range(size)]
    return True


def _polystab(a):
    v = np.roots(a).astype(complex)
    out = np.abs(v) > 1
    v[out] = 1 / np.conj(v[out])
    lead = a[np.nonzero(a)[0][0]]
    return (lead * np.poly(v)).real


def _impulse(a, n):
    # filter(1, a, [1 0 0 ...])
    h = np.zeros(n)
    for i in range(n):
        acc = 1.0 if i == 0 else 0.0
        for k in range(1, min(i, len(a) - 1) + 1):
            acc -= a[k] * h[i - k]
        h[i] = acc / a[0]
    return h


def _numf(h, a, nb):
    nh = len(h)
    impr = _impulse(a, nh)
    T = np.zeros((nh, nb + 1))
    for c in range(nb + 1):
        T[c:, c] = impr[:nh - c]
    return np.linalg.lstsq(T, h, rcond=None)[0]


def _denf(R, na):
    nr = len(R)
    rows = nr - 1 - na
    Rm = np.zeros((rows, na))
    for i in range(rows):
        for j in range(na):
            Rm[i, j] = R[abs(na + i - j)]
    x = np.linalg.lstsq(Rm, -R[na + 1:nr], rcond=None)[0]
    return np.r_[1.0, x]


def yulewalk(na, ff, aa, npt=512):
    """MATLAB/IT++ yulewalk(N, f, m): returns (b, a)"""
    ff = np.asarray(ff, float)
    aa = np.asarray(aa, float)
    lap = int(npt / 25)
    npt = npt + 1
    Ht = np.zeros(npt)
    df = np.diff(ff)
    nb = 1
    Ht[0] = aa[0]
    for i in range(len(ff) - 1):
        if df[i] == 0:
            nb = int(nb - lap / 2)
            ne = nb + lap
        else:
            ne = int(ff[i + 1] * npt)
        j = np.arange(nb, ne + 1)
        inc = 0 if ne == nb else (j - nb) / (ne - nb)
        Ht[nb - 1:ne] = inc * aa[i + 1] + (1 - inc) * aa[i]
        nb = ne + 1
    Ht = np.r_[Ht, Ht[npt - 2:0:-1]]
    n = len(Ht)
    n2 = (n + 1) // 2
    nr = 4 * na
    nt = np.arange(nr)
    R = np.real(np.fft.ifft(Ht * Ht))
    R = R[:nr] * (0.54 + 0.46 * np.cos(np.pi * nt / (nr - 1)))
    Rwindow = np.r_[0.5, np.ones(n2 - 1), np.zeros(n - n2)]
    A = _polystab(_denf(R, na))
    Qh = _numf(np.r_[R[0] / 2, R[1:nr]], A, na)
    with np.errstate(divide="ignore", invalid="ignore"):
        Ss = 2 * np.real(np.fft.fft(Qh, n) / np.fft.fft(A, n))
        hh = np.fft.ifft(np.exp(np.fft.fft(Rwindow * np.fft.ifft(np.log(Ss.astype(complex))))))
    B = _numf(np.real(hh[:nr]), A, na)
    return B, A


def envelope_grid(freq, amp):
    """arbitrary_magnitude_filter.h:63-84: the 256 (frequency, magnitude) points handed to yulewalk"""
    pts = []

    def insert(p):  # frequency_domain_envelope::insert: lower_bound => BEFORE equal frequencies
        k = len(pts)
        for i, q in enumerate(pts):
            if not q[0] < p[0]:
                k = i
                break
        pts.insert(k, p)

    for p in zip(map(float, freq), map(float, amp)):
        insert(p)
    # remove_outside_frequency_range(env, [0, 1]): drops f < 0 and f > 1
    pts[:] = [p for p in pts if 0.0 <= p[0] <= 1.0]
    insert((0.0, 0.0))
    insert((1.0, 0.0))
    xs = [p[0] for p in pts]
    ys = [p[1] for p in pts]

    def interp(a):  # cosine_interp.h:51-76 with linear_interp_functor
        import bisect
        it = bisect.bisect_left(xs, a)
        if it == 0:
            return ys[0]
        if it == len(xs):
            return ys[-1]
        x1, x2, y1, y2 = xs[it - 1], xs[it], ys[it - 1], ys[it]
        return y1 + ((a - x1) / (x2 - x1)) * (y2 - y1)

    f = [i / 255.0 for i in range(256)]
    m = [interp(v) for v in f]
    return f, m


def arbitrary_magnitude_filter(freq, amp, order=ORDER):
    """arbitrary_magnitude_filter.h:63-95: points (frequency 0..1 = dc..nyquist, amplitude)"""
    f, m = envelope_grid(freq, amp)
    return yulewalk(order, f, m)


def reflectance_filter(absorption, sample_rate):
    """compute_reflectance_filter_coefficients (fitted_boundary.h:79-104) -> (b, a)"""
    centres = band_centres_hz() / sample_rate * 2
    reflectance = np.sqrt(1 - np.asarray(absorption, float))
    b, a = arbitrary_magnitude_filter(centres, reflectance)
    if not is_stable(a):
        raise RuntimeError("Unable to generate stable boundary filter.")
    return b, a


def to_impedance(b, a):
    """to_impedance_coefficients (fitted_boundary.h:20-50)"""
    b, a = np.asarray(b, float), np.asarray(a, float)
    rb, ra = a + b, a - b
    if ra[0]:
        norm = 1.0 / ra[0]
        rb, ra = rb * norm, ra * norm
    return rb, ra
