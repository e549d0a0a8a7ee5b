// pp_kernels.cuh -- device code of the post-processing that turns the ray path's energy
// histogram into an audio-rate signal, and joins it with the waveguide's.
//
//   reference                                                              here
//   raytracer/src/stochastic/postprocessing.cpp:16-50 generate_dirac_sequence  pp_exponentials + pp_dirac_walk
//   .../postprocessing.cpp:57-97 weight_sequence                              pp_weight_sequence
//   frequency_domain/multiband_filter.h:49-93 + src/filter.cpp:22-47          pp_fft_pass (Stockham, radix 2)
//     (FFTW r2c/c2r in the reference)                                          + pp_band_envelopes
//   core/mixdown.h:12-24 mixdown                                               pp_mixdown
//   combined/postprocess.h:33-60 crossover_filter                             pp_lo_hi_envelopes + pp_sum2
//   combined/postprocess.h:117-134 left_hanning window                        pp_left_hanning
//
// All arithmetic that decides WHERE an event lands (the Poisson walk) is fp64 in a fixed
// operation order with its own logarithm, so the sequence is bit-identical to the oracle's;
// the FFT filter is float like the reference's (tolerance stated in tests/test_pp_gpu.py).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "rt_kernels.cuh"  // philox

namespace wvb {
namespace pp {

// -log(x) for x in (0, 1], in plain IEEE double arithmetic (no libm, no contraction):
// x = m 2^e, m in [sqrt(1/2), sqrt(2)), s = (m - 1) / (m + 1), log m = 2 s (1 + s^2/3 + ... + s^24/25)
__device__ __forceinline__ double neg_log_fixed(double x) {
    int e;
    double m = frexp(x, &e);  // m in [0.5, 1)
    if (m < 0.70710678118654752440) {
        m = m * 2.0;
        e -= 1;
    }
    const double s = (m - 1.0) / (m + 1.0);
    const double z = s * s;
    double p = 1.0 / 25.0;
#pragma unroll
    for (int k = 23; k >= 1; k -= 2) p = p * z + 1.0 / (double)k;
    const double lm = 2.0 * s * p;
    const double le = (double)e * 0.693147180369123816490 + (double)e * 1.90821492927058770002e-10;
    return -(le + lm);
}

// unit-rate exponential variates E_k = -log(1 - u_k), u_k from Philox4x32-10(seed; k, 0, 7, 0):
// the part of interval_size (postprocessing.h:38-44) that does not depend on t
__global__ void pp_exponentials(unsigned long long seed, uint32_t n, double* __restrict__ out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t o0, o1;
    rt::philox(k, 0u, 7u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), o0, o1);
    const double u = ((double)(o0 >> 5) * 67108864.0 + (double)(o1 >> 6)) * (1.0 / 9007199254740992.0);
    out[k] = neg_log_fixed(1.0 - u);
}

// generate_dirac_sequence's loop (postprocessing.cpp:38-48): t += E_k / min(c t^2, 10000),
// sample floor(t * rate) gets +-1 by the parity of floor(2 t rate). One thread: the walk is a
// recurrence in t; its expensive part (random numbers, logarithms) was done in parallel above.
// n_used = events consumed; *overflow is raised when the variates ran out before max_time.
__global__ void pp_dirac_walk(const double* __restrict__ exps, uint32_t n_exps, double constant, double t0,
                              double max_time, double sample_rate, float* __restrict__ seq, uint32_t len,
                              uint32_t* __restrict__ n_used, int* __restrict__ overflow) {
    if (blockIdx.x || threadIdx.x) return;
    double t = t0;
    uint32_t k = 0;
    while (t < max_time) {
        const double sample_index = t * sample_rate;
        const unsigned long long idx = (unsigned long long)sample_index;
        const unsigned long long twice = (unsigned long long)(2 * sample_index);
        if (idx < len) seq[idx] = (twice & 1ull) ? -1.0f : 1.0f;
        if (k >= n_exps) {
            *overflow = 1;
            break;
        }
        const double mean = fmin(constant * (t * t), 10000.0);
        t += exps[k++] / mean;
    }
    *n_used = k;
}

// weight_sequence (postprocessing.cpp:57-97): one block per histogram bin; the bin's energy is
// spread over the dirac events that fall inside it. hist: [n_bins][8] doubles (the device
// histogram of the ray path) read as the float bands the reference stores.
// out: [len][8] float, len = min(sequence length, size_t(n_bins * seq_rate / hist_rate))
__global__ void pp_weight_sequence(const double* __restrict__ hist, uint32_t n_bins, double hist_rate,
                                   const float* __restrict__ seq, uint32_t len, double seq_rate,
                                   double acoustic_impedance, float* __restrict__ out) {
    const uint32_t bin = blockIdx.x;
    if (bin >= n_bins) return;
    const unsigned long long b0 = (unsigned long long)((double)bin * seq_rate / hist_rate);
    const unsigned long long b1 = (unsigned long long)((double)(bin + 1) * seq_rate / hist_rate);
    const uint32_t beg = (uint32_t)min(b0, (unsigned long long)len), end = (uint32_t)min(b1, (unsigned long long)len);
    __shared__ float ss_shared;
    __shared__ float partial[32];
    float ss = 0.0f;  // +-1 / 0 values: the sum is exact in float in any order
    for (uint32_t j = beg + threadIdx.x; j < end; j += blockDim.x) ss += seq[j] * seq[j];
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) partial[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.0f;
        for (int w = 0; w < (int)((blockDim.x + 31) / 32); ++w) tot += partial[w];
        ss_shared = tot;
    }
    __syncthreads();
    const float squared_summed = ss_shared;
    double scale[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        if (squared_summed != 0.0f) {
            // histogram[i] / squared_summed in float, then intensity_to_pressure in double
            const float q = (float)hist[(size_t)bin * 8 + b] / squared_summed;
            const double v = (double)q * acoustic_impedance;
            scale[b] = copysign(sqrt(fabs(v)), (double)q);
        } else {
            scale[b] = 0.0;
        }
    }
    for (uint32_t j = beg + threadIdx.x; j < end; j += blockDim.x) {
        const double s = (double)seq[j];
#pragma unroll
        for (int b = 0; b < 8; ++b) out[(size_t)j * 8 + b] = (float)(s * scale[b]);
    }
}

// ---- FFT: Stockham autosort, radix 2, out of place, batched ------------------------------------
// One pass: y[q + s (2p + r)] from x[q + s (p + r m)], n = 2 m s. Twiddles are evaluated in
// double (sincospi) and rounded once. batch = blockIdx.y.
__global__ void pp_fft_pass(const float2* __restrict__ x, float2* __restrict__ y, uint32_t n, uint32_t s,
                            int inverse) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;  // butterfly index in [0, n/2)
    const uint32_t half = n >> 1;
    if (i >= half) return;
    const size_t base = (size_t)blockIdx.y * n;
    const uint32_t q = i % s, p = i / s;      // p in [0, m), m = half / s
    const uint32_t m = half / s;
    double sn, cs;
    sincospi((inverse ? 1.0 : -1.0) * (double)p / (double)m, &sn, &cs);  // w = exp(-+ 2 pi i p / (2 m))
    const float2 w = make_float2((float)cs, (float)sn);
    const float2 a = x[base + q + s * p];
    const float2 b = x[base + q + s * (p + m)];
    const float2 d = make_float2(a.x - b.x, a.y - b.y);
    y[base + q + s * (2 * p)] = make_float2(a.x + b.x, a.y + b.y);
    y[base + q + s * (2 * p + 1)] = make_float2(d.x * w.x - d.y * w.y, d.x * w.y + d.y * w.x);
}

// real signal, element `band` of `stride` interleaved floats -> zero-padded complex batch
__global__ void pp_load(const float* __restrict__ in, uint32_t len, uint32_t stride, float2* __restrict__ x,
                        uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t band = blockIdx.y;
    x[(size_t)band * n + i] = make_float2(i < len ? in[(size_t)i * stride + band] : 0.0f, 0.0f);
}

// frequency_domain/src/envelope.cpp:24-110 with l = 0
__device__ __forceinline__ double band_edge0(double p, double P) { return ((p / P) + 1) / 2; }
__device__ __forceinline__ double lopass_mag(double f, double edge, double wf) {
    const double aw = edge * wf;
    if (f < edge - aw) return 1;
    if (f < edge + aw) {
        if (aw == 0) return (f - edge) < 0 ? 1.0 : 0.0;
        const double c = cos(M_PI * band_edge0(f - edge, aw) / 2);
        return c * c;
    }
    return 0;
}
__device__ __forceinline__ double hipass_mag(double f, double edge, double wf) {
    const double aw = edge * wf;
    if (f < edge - aw) return 0;
    if (f < edge + aw) {
        if (aw == 0) return 0 <= (f - edge) ? 1.0 : 0.0;
        const double s = sin(M_PI * band_edge0(f - edge, aw) / 2);
        return s * s;
    }
    return 1;
}

struct Edges {
    double e[9];
};
// multiband_filter's callback (multiband_filter.h:66-78): bin i of band b is scaled by the band-pass
// magnitude at freq = i / n (a float in the reference, filter.cpp:31); bins above n/2 are the
// mirror image (the reference's r2c transform only holds the lower half).
// mode 0: band-pass per batch entry; mode 1: batch 0 low-pass, batch 1 high-pass at e[0]
__global__ void pp_envelopes(float2* __restrict__ x, uint32_t n, Edges edges, double width_factor, int mode) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t band = blockIdx.y;
    const uint32_t k = i <= n / 2 ? i : n - i;
    const double f = (double)((float)k / (float)n);
    double amp;
    if (mode == 0) amp = lopass_mag(f, edges.e[band + 1], width_factor) * hipass_mag(f, edges.e[band], width_factor);
    else amp = band == 0 ? lopass_mag(f, edges.e[0], width_factor) : hipass_mag(f, edges.e[0], width_factor);
    const float a = (float)amp;
    float2 v = x[(size_t)band * n + i];
    v.x *= a;
    v.y *= a;
    x[(size_t)band * n + i] = v;
}

// real part / n of the first len samples, summed over the batch in band order (mixdown,
// core/mixdown.h:12-15; sum_vectors for the crossover)
__global__ void pp_mixdown(const float2* __restrict__ x, uint32_t n, uint32_t bands, float* __restrict__ out,
                           uint32_t len) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    float acc = 0.0f;
    for (uint32_t b = 0; b < bands; ++b) acc += x[(size_t)b * n + i].x / (float)n;
    out[i] = acc;
}

// crossover_filter's sum (combined/postprocess.h:46-59): each filtered signal keeps its own
// length (run_filter returns distance(b, e) samples, the filter's tail past the end is dropped),
// core::sum_vectors adds them over the longer of the two
__global__ void pp_sum2(const float2* __restrict__ x, uint32_t n, uint32_t n_lo, uint32_t n_hi,
                        float* __restrict__ out, uint32_t len) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    const float a = i < n_lo ? x[i].x / (float)n : 0.0f;
    const float b = i < n_hi ? x[(size_t)n + i].x / (float)n : 0.0f;
    out[i] = a + b;
}

// left_hanning (core/sinc.h:75-82) multiplied in: hanning_point(f) = 0.5 - 0.5 cos(2 pi f)
__global__ void pp_left_hanning(float* __restrict__ sig, uint32_t window) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= window) return;
    const double f = (double)i / (2 * ((double)window - 1.0));
    const float w = (float)(0.5 - 0.5 * cos(2 * M_PI * f));
    sig[i] = w * sig[i];
}

}  // namespace pp
}  // namespace wvb
