// TEST INFRASTRUCTURE ONLY. Stands in for FFTW 3's single-precision interface, which the
// reference's frequency_domain library links (src/frequency_domain/src/{filter,plan,buffer,
// traits}.cpp) and which is not in this image. Only what those files call is here:
//   fftwf_alloc_real / fftwf_alloc_complex / fftwf_free,
//   fftwf_plan_dft_r2c_1d / fftwf_plan_dft_c2r_1d (flags ignored), fftwf_execute, fftwf_destroy_plan,
// with FFTW's published contract: both transforms unnormalised, r2c writes bins 0..n/2 of
// X[k] = sum_j x[j] exp(-2 pi i jk/n), c2r reads that half spectrum as Hermitian (the imaginary
// parts of bin 0 and, for even n, bin n/2 are ignored) and writes the n real samples.
// The arithmetic this stand-in decides is the transform's rounding alone: it is evaluated in double
// (radix-2 for power-of-two n, which is all the reference asks for; a direct sum otherwise) and
// rounded once to float, so it sits within float rounding of any correct single-precision FFT.
// Tests that go through it therefore carry a float tolerance, never bit identity.
#pragma once
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstdlib>
#include <vector>

typedef float fftwf_complex[2];

struct wvb_stub_fftwf_plan {
    size_t n;
    float* real;
    fftwf_complex* cplx;
    bool forward;
};
typedef wvb_stub_fftwf_plan* fftwf_plan;

#define FFTW_ESTIMATE (1U << 6)

inline void* wvb_stub_fftwf_malloc(size_t bytes) {
    void* p = nullptr;
    return posix_memalign(&p, 64, bytes ? bytes : 1) == 0 ? p : nullptr;
}
inline float* fftwf_alloc_real(size_t n) { return static_cast<float*>(wvb_stub_fftwf_malloc(n * sizeof(float))); }
inline fftwf_complex* fftwf_alloc_complex(size_t n) {
    return static_cast<fftwf_complex*>(wvb_stub_fftwf_malloc(n * sizeof(fftwf_complex)));
}
inline void fftwf_free(void* p) { std::free(p); }

inline fftwf_plan fftwf_plan_dft_r2c_1d(int n, float* in, fftwf_complex* out, unsigned) {
    return new wvb_stub_fftwf_plan{size_t(n), in, out, true};
}
inline fftwf_plan fftwf_plan_dft_c2r_1d(int n, fftwf_complex* in, float* out, unsigned) {
    return new wvb_stub_fftwf_plan{size_t(n), out, in, false};
}
inline void fftwf_destroy_plan(fftwf_plan p) { delete p; }

// in-place complex transform of v, sign = -1 forward, +1 backward, unnormalised
inline void wvb_stub_dft(std::vector<std::complex<double>>& v, int sign) {
    const size_t n = v.size();
    const double pi = 3.14159265358979323846;
    if (n & (n - 1)) {
        std::vector<std::complex<double>> out(n);
        for (size_t k = 0; k < n; ++k) {
            std::complex<double> acc{0, 0};
            for (size_t j = 0; j < n; ++j) {
                const double a = sign * 2 * pi * double((j * k) % n) / double(n);
                acc += v[j] * std::complex<double>{std::cos(a), std::sin(a)};
            }
            out[k] = acc;
        }
        v.swap(out);
        return;
    }
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(v[i], v[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        std::vector<std::complex<double>> w(len / 2);
        for (size_t k = 0; k < len / 2; ++k) {
            const double a = sign * 2 * pi * double(k) / double(len);
            w[k] = {std::cos(a), std::sin(a)};
        }
        for (size_t i = 0; i < n; i += len) {
            for (size_t k = 0; k < len / 2; ++k) {
                const std::complex<double> a = v[i + k], b = v[i + k + len / 2] * w[k];
                v[i + k] = a + b;
                v[i + k + len / 2] = a - b;
            }
        }
    }
}

inline void fftwf_execute(const fftwf_plan p) {
    const size_t n = p->n;
    if (!n) return;
    std::vector<std::complex<double>> v(n);
    if (p->forward) {
        for (size_t i = 0; i < n; ++i) v[i] = {double(p->real[i]), 0.0};
        wvb_stub_dft(v, -1);
        for (size_t k = 0; k <= n / 2; ++k) {
            p->cplx[k][0] = float(v[k].real());
            p->cplx[k][1] = float(v[k].imag());
        }
    } else {
        for (size_t k = 0; k <= n / 2; ++k) {
            const bool self = k == 0 || 2 * k == n;
            v[k] = {double(p->cplx[k][0]), self ? 0.0 : double(p->cplx[k][1])};
            if (!self) v[n - k] = std::conj(v[k]);
        }
        wvb_stub_dft(v, +1);
        for (size_t i = 0; i < n; ++i) p->real[i] = float(v[i].real());
    }
}
