// lrs_design.cpp -- locally-reacting-surface filter design behind the C ABI
// (wvb_lrs_*). Host code, like its reference counterpart:
//
//   compute_reflectance_filter_coefficients   src/waveguide/include/waveguide/fitted_boundary.h:79-104
//   arbitrary_magnitude_filter<6>             src/waveguide/include/waveguide/arbitrary_magnitude_filter.h:63-95
//   frequency_domain_envelope                 src/waveguide/src/frequency_domain_envelope.cpp:27-62
//   interp / linear_interp                    src/core/include/core/cosine_interp.h:17-76
//   band centres                              src/hrtf/lib/include/hrtf/multiband.h:11-20,
//                                             src/frequency_domain/src/envelope.cpp:49-56
//   is_stable                                 src/waveguide/include/waveguide/stable.h:11-50
//   to_impedance_coefficients / to_flat       src/waveguide/include/waveguide/fitted_boundary.h:20-50,72-75
//
// The reference delegates the fit to itpp::yulewalk (IT++, fetched at configure time,
// un-vendored; plus BLAS/LAPACK). This file carries its own implementation of that
// published algorithm (modified Yule-Walker, Friedlander & Porat; the MATLAB
// `yulewalk.m` formulation) with a radix-2 FFT, Householder least squares and an
// Aberth root finder, so the dependency disappears. It reproduces the nine coefficient
// sets the reference checked in (bin/boundary_test/output.soft/coefficients.txt) to
// better than 1e-10 (tests/test_lrs_design.py).
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <vector>

#include "common.h"

using namespace wvb;

namespace {

using cd = std::complex<double>;
using vec = std::vector<double>;
constexpr int ORDER = 6;

// ---- small numerics --------------------------------------------------------------
void fft(std::vector<cd>& x, bool inverse) {
    const size_t n = x.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(x[i], x[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const double ang = 2 * M_PI / double(len) * (inverse ? 1 : -1);
        for (size_t i = 0; i < n; i += len) {
            for (size_t k = 0; k < len / 2; ++k) {
                const cd w(std::cos(ang * double(k)), std::sin(ang * double(k)));
                const cd u = x[i + k], v = x[i + k + len / 2] * w;
                x[i + k] = u + v;
                x[i + k + len / 2] = u - v;
            }
        }
    }
    if (inverse) {
        for (cd& v : x) v /= double(n);
    }
}

// min ||A x - b||, A rows x cols (row-major), Householder QR
vec lstsq(vec A, vec b, size_t rows, size_t cols) {
    for (size_t k = 0; k < cols; ++k) {
        double norm = 0;
        for (size_t i = k; i < rows; ++i) norm += A[i * cols + k] * A[i * cols + k];
        norm = std::sqrt(norm);
        if (norm == 0) continue;
        const double alpha = A[k * cols + k] > 0 ? -norm : norm;
        vec v(rows - k);
        for (size_t i = k; i < rows; ++i) v[i - k] = A[i * cols + k];
        v[0] -= alpha;
        double vv = 0;
        for (double t : v) vv += t * t;
        if (vv == 0) continue;
        for (size_t j = k; j < cols; ++j) {
            double s = 0;
            for (size_t i = k; i < rows; ++i) s += v[i - k] * A[i * cols + j];
            s = 2 * s / vv;
            for (size_t i = k; i < rows; ++i) A[i * cols + j] -= s * v[i - k];
        }
        double s = 0;
        for (size_t i = k; i < rows; ++i) s += v[i - k] * b[i];
        s = 2 * s / vv;
        for (size_t i = k; i < rows; ++i) b[i] -= s * v[i - k];
    }
    vec x(cols, 0.0);
    for (size_t kk = cols; kk-- > 0;) {
        double s = b[kk];
        for (size_t j = kk + 1; j < cols; ++j) s -= A[kk * cols + j] * x[j];
        const double d = A[kk * cols + kk];
        x[kk] = d != 0 ? s / d : 0.0;  // rank-deficient column: minimum-effort choice
    }
    return x;
}

// all roots of a[0] z^n + ... + a[n] (Aberth-Ehrlich, then one Newton polish each)
std::vector<cd> poly_roots(vec a) {
    size_t lead = 0;
    while (lead < a.size() && a[lead] == 0) ++lead;
    a.erase(a.begin(), a.begin() + lead);
    size_t zeros = 0;
    while (a.size() > 1 && a.back() == 0) {
        a.pop_back();
        ++zeros;
    }
    std::vector<cd> z;
    if (a.size() > 1) {
        const size_t n = a.size() - 1;
        auto eval = [&](cd x, cd& d) {
            cd p = a[0];
            d = 0;
            for (size_t i = 1; i <= n; ++i) {
                d = d * x + p;
                p = p * x + a[i];
            }
            return p;
        };
        const double r0 = std::pow(std::fabs(a[n] / a[0]), 1.0 / double(n));
        z.resize(n);
        for (size_t k = 0; k < n; ++k) z[k] = std::polar(r0 > 0 ? r0 : 1.0, 2 * M_PI * double(k) / double(n) + 0.4);
        for (int it = 0; it < 500; ++it) {
            double moved = 0;
            for (size_t k = 0; k < n; ++k) {
                cd d;
                const cd p = eval(z[k], d);
                if (p == cd(0)) continue;
                const cd w = p / d;
                cd s = 0;
                for (size_t j = 0; j < n; ++j) {
                    if (j != k) s += 1.0 / (z[k] - z[j]);
                }
                const cd step = w / (1.0 - w * s);
                z[k] -= step;
                moved = std::max(moved, std::abs(step) / std::max(std::abs(z[k]), 1e-300));
            }
            if (moved < 1e-16) break;
        }
        for (size_t k = 0; k < n; ++k) {
            cd d;
            const cd p = eval(z[k], d);
            if (d != cd(0)) z[k] -= p / d;
        }
    }
    z.insert(z.end(), zeros, cd(0));
    return z;
}

// polystab: roots outside the unit circle are reflected inside
vec polystab(const vec& a) {
    size_t lead = 0;
    while (lead < a.size() && a[lead] == 0) ++lead;
    if (lead == a.size()) return a;
    std::vector<cd> v = poly_roots(a);
    for (cd& r : v) {
        if (std::abs(r) > 1) r = 1.0 / std::conj(r);
    }
    std::vector<cd> p{cd(1)};
    for (const cd& r : v) {  // poly(v)
        p.push_back(0);
        for (size_t i = p.size() - 1; i > 0; --i) p[i] -= r * p[i - 1];
    }
    vec out(p.size());
    for (size_t i = 0; i < p.size(); ++i) out[i] = a[lead] * p[i].real();
    return out;
}

// filter(1, a, [1 0 0 ...])
vec impulse(const vec& a, size_t n) {
    vec h(n, 0.0);
    for (size_t i = 0; i < n; ++i) {
        double acc = i == 0 ? 1.0 : 0.0;
        for (size_t k = 1; k < a.size() && k <= i; ++k) acc -= a[k] * h[i - k];
        h[i] = acc / a[0];
    }
    return h;
}

// numerator B given the impulse response h of B/A and the denominator A
vec numf(const vec& h, const vec& a, size_t nb) {
    const size_t nh = h.size();
    const vec impr = impulse(a, nh);
    vec T(nh * (nb + 1), 0.0);
    for (size_t c = 0; c <= nb; ++c) {
        for (size_t r = c; r < nh; ++r) T[r * (nb + 1) + c] = impr[r - c];
    }
    return lstsq(T, h, nh, nb + 1);
}

// denominator from covariances (modified Yule-Walker equations, least squares)
vec denf(const vec& R, size_t na) {
    const size_t nr = R.size(), rows = nr - 1 - na;
    vec Rm(rows * na), rhs(rows);
    for (size_t i = 0; i < rows; ++i) {
        for (size_t j = 0; j < na; ++j) {
            const long lag = long(na + i) - long(j);
            Rm[i * na + j] = R[size_t(lag < 0 ? -lag : lag)];
        }
        rhs[i] = -R[na + 1 + i];
    }
    vec x = lstsq(Rm, rhs, rows, na);
    x.insert(x.begin(), 1.0);
    return x;
}

// yulewalk(N, f, m) on the 512-point grid
void yulewalk(size_t na, const vec& ff, const vec& aa, vec& B, vec& A) {
    const long npt = 512 + 1;
    const long lap = 512 / 25;
    vec Ht(size_t(npt), 0.0);
    long nb = 1;
    Ht[0] = aa[0];
    for (size_t i = 0; i + 1 < ff.size(); ++i) {
        long ne;
        if (ff[i + 1] - ff[i] == 0) {
            nb = long(double(nb) - double(lap) / 2);
            ne = nb + lap;
        } else {
            ne = long(ff[i + 1] * double(npt));
        }
        WVB_REQUIRE(nb >= 1 && ne <= npt, WVB_ERR_INVALID, "yulewalk: frequency grid out of range");
        for (long j = nb; j <= ne; ++j) {
            const double inc = ne == nb ? 0.0 : double(j - nb) / double(ne - nb);
            Ht[size_t(j - 1)] = inc * aa[i + 1] + (1 - inc) * aa[i];
        }
        nb = ne + 1;
    }
    std::vector<cd> H;
    for (double v : Ht) H.emplace_back(v * v, 0.0);
    for (long k = npt - 2; k >= 1; --k) H.emplace_back(Ht[size_t(k)] * Ht[size_t(k)], 0.0);
    const size_t n = H.size();  // 1024
    const size_t n2 = (n + 1) / 2;
    const size_t nr = 4 * na;
    fft(H, true);
    vec R(nr);
    for (size_t t = 0; t < nr; ++t) {
        R[t] = H[t].real() * (0.54 + 0.46 * std::cos(M_PI * double(t) / double(nr - 1)));
    }
    A = polystab(denf(R, na));
    vec R2 = R;
    R2[0] = R[0] / 2;
    const vec Qh = numf(R2, A, na);
    std::vector<cd> fq(n, cd(0)), fa(n, cd(0));
    for (size_t i = 0; i < Qh.size(); ++i) fq[i] = Qh[i];
    for (size_t i = 0; i < A.size(); ++i) fa[i] = A[i];
    fft(fq, false);
    fft(fa, false);
    std::vector<cd> c(n);
    for (size_t i = 0; i < n; ++i) c[i] = std::log(cd(2 * (fq[i] / fa[i]).real(), 0.0));
    fft(c, true);
    for (size_t i = 0; i < n; ++i) c[i] *= i == 0 ? 0.5 : (i < n2 ? 1.0 : 0.0);
    fft(c, false);
    for (cd& v : c) v = std::exp(v);
    fft(c, true);
    vec hh(nr);
    for (size_t i = 0; i < nr; ++i) hh[i] = c[i].real();
    B = numf(hh, A, na);
}

struct point {
    double frequency, amplitude;
};

// arbitrary_magnitude_filter<6>
void magnitude_filter(std::vector<point> in, wvb_coefficients_canonical* out) {
    std::vector<point> env;
    auto insert = [&](point p) {  // lower_bound: before points of equal frequency
        auto it = std::lower_bound(env.begin(), env.end(), p,
                                   [](const point& a, const point& b) { return a.frequency < b.frequency; });
        env.insert(it, p);
    };
    for (const point& p : in) insert(p);
    env.erase(std::remove_if(env.begin(), env.end(),
                             [](const point& p) { return !(0.0 <= p.frequency && p.frequency <= 1.0); }),
              env.end());
    insert({0.0, 0.0});
    insert({1.0, 0.0});
    vec f(256), m(256);
    for (int i = 0; i < 256; ++i) {
        const double a = i / (256 - 1.0);
        f[size_t(i)] = a;
        // interp(b, e, a, linear_interp_functor) -- cosine_interp.h:51-76
        auto it = std::lower_bound(env.begin(), env.end(), a,
                                   [](const point& p, double v) { return p.frequency < v; });
        if (it == env.begin()) {
            m[size_t(i)] = env.front().amplitude;
        } else if (it == env.end()) {
            m[size_t(i)] = env.back().amplitude;
        } else {
            const point a1 = *(it - 1), a2 = *it;
            m[size_t(i)] = a1.amplitude + ((a - a1.frequency) / (a2.frequency - a1.frequency)) *
                                                  (a2.amplitude - a1.amplitude);
        }
    }
    vec B, A;
    yulewalk(ORDER, f, m, B, A);
    WVB_REQUIRE(B.size() == ORDER + 1 && A.size() == ORDER + 1, WVB_ERR_INVALID,
                "yulewalk returned %zu/%zu coefficients", B.size(), A.size());
    for (int i = 0; i <= ORDER; ++i) {
        out->b[i] = B[size_t(i)];
        out->a[i] = A[size_t(i)];
    }
}

bool stable(const double* a, size_t n) {  // stable.h:11-50
    vec cur(a, a + n);
    while (cur.size() > 1) {
        const double rci = cur.back();
        if (1 <= std::fabs(rci)) return false;
        const size_t size = cur.size() - 1;
        vec next(size);
        for (size_t i = 0; i < size; ++i) next[i] = (cur[i] - cur[size - i] * rci) / (1 - rci * rci);
        cur.swap(next);
    }
    return true;
}

template <class F>
wvb_status guarded(F&& f) {
    try {
        f();
        return WVB_OK;
    } catch (const status_error& e) {
        return e.code;
    } catch (const std::exception& e) {
        set_last_error("%s", e.what());
        return WVB_ERR_INVALID;
    }
}

}  // namespace

extern "C" {

wvb_status wvb_lrs_arbitrary_magnitude_filter(const double* frequency, const double* amplitude, uint32_t n,
                                              wvb_coefficients_canonical* out) {
    if (!out || (n && (!frequency || !amplitude))) return WVB_ERR_INVALID;
    return guarded([&] {
        std::vector<point> pts(n);
        for (uint32_t i = 0; i < n; ++i) pts[i] = {frequency[i], amplitude[i]};
        magnitude_filter(pts, out);
    });
}

int wvb_lrs_is_stable(const double* a, uint32_t n) { return a && n ? (stable(a, n) ? 1 : 0) : 1; }

wvb_status wvb_lrs_reflectance_filter(const double absorption[8], double sample_rate,
                                      wvb_coefficients_canonical* out) {
    if (!absorption || !out) return WVB_ERR_INVALID;
    return guarded([&] {
        std::vector<point> pts(8);
        for (int b = 0; b < 8; ++b) {
            // band_centre_frequency(b, 8, [20, 20000]) / sample_rate * 2
            const double hz = 20.0 * std::pow(20000.0 / 20.0, double(2 * b + 1) / double(2 * 8));
            pts[size_t(b)] = {hz / sample_rate * 2, std::sqrt(1 - absorption[b])};
        }
        magnitude_filter(pts, out);
        // the reference retries 1000 times with identical inputs, then throws
        WVB_REQUIRE(stable(out->a, ORDER + 1), WVB_ERR_INVALID, "Unable to generate stable boundary filter.");
    });
}

void wvb_lrs_to_impedance(const wvb_coefficients_canonical* c, wvb_coefficients_canonical* out) {
    wvb_coefficients_canonical r;
    for (int i = 0; i <= ORDER; ++i) {
        r.b[i] = c->a[i] + c->b[i];
        r.a[i] = c->a[i] - c->b[i];
    }
    if (r.a[0]) {
        const double norm = 1.0 / r.a[0];
        for (int i = 0; i <= ORDER; ++i) {
            r.b[i] *= norm;
            r.a[i] *= norm;
        }
    }
    *out = r;
}

void wvb_lrs_flat(double absorption, wvb_coefficients_canonical* out) {
    wvb_coefficients_canonical c;
    std::memset(&c, 0, sizeof c);
    c.b[0] = std::sqrt(1 - absorption);  // absorption_to_pressure_reflectance
    c.a[0] = 1;
    wvb_lrs_to_impedance(&c, out);
}

}  // extern "C"
