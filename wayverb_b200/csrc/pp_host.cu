// pp_host.cu -- host side of the post-processing entry points (include/wvb200.h, wvb_pp_*):
// the ray path's energy histogram -> dirac sequence -> weighted multiband signal -> band-pass
// filter bank -> mono signal (raytracer/src/stochastic/postprocessing.cpp:29-112), and the
// crossover that joins it with the waveguide's signal (combined/postprocess.h:33-136).
// Everything runs on the device; the FFT is a batched Stockham radix-2 transform of our own
// (pp_kernels.cuh), the reference's is FFTW behind frequency_domain::filter.
#include <cuda_runtime.h>

#include <cmath>
#include <vector>

#include "common.h"
#include "pp_kernels.cuh"

using namespace wvb;

namespace {

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

void use_device(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_last_error("no CUDA device visible (this library has no CPU fallback)");
        throw status_error{WVB_ERR_NO_DEVICE};
    }
    WVB_REQUIRE(device >= 0 && device < ndev, WVB_ERR_NO_DEVICE, "device %d of %d", device, ndev);
    WVB_CUDA(cudaSetDevice(device));
}

// best_fft_length(len) << 2 (multiband_filter.h:35-43,60-61)
uint32_t padded_fft_length(uint64_t len) {
    const double l2 = std::ceil(std::log2((double)len));
    const uint64_t n = (uint64_t)std::pow(2.0, l2) << 2;
    WVB_REQUIRE(n <= (1ull << 26), WVB_ERR_UNSUPPORTED, "signal of %llu samples is too long to filter",
                (unsigned long long)len);
    return (uint32_t)n;
}

// in-place result ends up in the returned buffer (a or b)
float2* fft(float2* a, float2* b, uint32_t n, uint32_t batch, bool inverse, cudaStream_t st) {
    float2* x = a;
    float2* y = b;
    const uint32_t half = n / 2;
    const dim3 grid((half + 255) / 256, batch);
    // Stockham: the stride doubles every pass, from 1 to n / 2
    for (uint32_t s = 1; s < n; s <<= 1) {
        pp::pp_fft_pass<<<grid, 256, 0, st>>>(x, y, n, s, inverse ? 1 : 0);
        std::swap(x, y);
    }
    return x;
}

// hrtf_band_params(sample_rate) (hrtf/multiband.h:12-36): 8 bands over 20 Hz - 20 kHz, overlap 1
void band_params(double sample_rate, pp::Edges* edges, double* width_factor) {
    const double lo = 20.0, hi = 20000.0;
    for (int i = 0; i <= 8; ++i) edges->e[i] = lo * std::pow(hi / lo, i / 8.0) / sample_rate;  // envelope.cpp:46-49
    const double base = std::pow(hi / lo, 1.0 / 8);                                             // envelope.cpp:5-16
    *width_factor = (base - 1) / (base + 1) * 1.0;
}

struct device_signal {
    dev_buf<float> data;
    uint32_t len = 0;
};

// generate_dirac_sequence on the device -> seq (len = ceil(max_time * rate))
void dirac_sequence(const wvb_pp_params* p, double rate, double max_time, dev_buf<float>& seq, uint32_t* len_out,
                    uint32_t* events_out, cudaStream_t st) {
    WVB_REQUIRE(p->room_volume > 0 && p->speed_of_sound > 0 && rate > 0 && max_time >= 0, WVB_ERR_INVALID,
                "bad dirac sequence parameters");
    const double constant = 4 * M_PI * std::pow(p->speed_of_sound, 3.0) / p->room_volume;  // postprocessing.cpp:16-19
    const double t0 = std::pow(2.0 * std::log(2.0) / constant, 1.0 / 3.0);                  // :25-27
    const uint64_t len = (uint64_t)std::ceil(max_time * rate);
    WVB_REQUIRE(len < (1ull << 31), WVB_ERR_UNSUPPORTED, "sequence too long");
    seq.alloc((size_t)std::max<uint64_t>(len, 1), true);
    // events: at most 10000 per second on average; 1.5x + slack covers the Poisson spread by far,
    // the walk reports if it ever ran out
    const uint32_t n_exps = (uint32_t)std::min<double>(1.5 * 10000.0 * max_time + 4096.0, 2.0e9);
    dev_buf<double> exps;
    exps.alloc(n_exps, false);
    dev_buf<uint32_t> used;
    used.alloc(1, true);
    dev_buf<int> overflow;
    overflow.alloc(1, true);
    pp::pp_exponentials<<<(n_exps + 255) / 256, 256, 0, st>>>(p->seed, n_exps, exps.p);
    pp::pp_dirac_walk<<<1, 32, 0, st>>>(exps.p, n_exps, constant, t0, max_time, rate, seq.p, (uint32_t)len, used.p,
                                         overflow.p);
    int h_over = 0;
    uint32_t h_used = 0;
    WVB_CUDA(cudaMemcpyAsync(&h_over, overflow.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    WVB_CUDA(cudaMemcpyAsync(&h_used, used.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    WVB_CUDA(cudaStreamSynchronize(st));
    WVB_CUDA(cudaGetLastError());
    WVB_REQUIRE(!h_over, WVB_ERR_UNSUPPORTED, "dirac sequence ran out of random variates");
    *len_out = (uint32_t)len;
    if (events_out) *events_out = h_used;
}

// multiband_filter + mixdown of d_multi [len][8] -> d_out [len]
void multiband_mixdown(const float* d_multi, uint32_t len, double sample_rate, float* d_out, cudaStream_t st) {
    if (!len) return;
    const uint32_t n = padded_fft_length(len);
    dev_buf<float2> a, b;
    a.alloc((size_t)n * 8, false);
    b.alloc((size_t)n * 8, false);
    pp::pp_load<<<dim3((n + 255) / 256, 8), 256, 0, st>>>(d_multi, len, 8, a.p, n);
    float2* x = fft(a.p, b.p, n, 8, false, st);
    pp::Edges edges;
    double wf;
    band_params(sample_rate, &edges, &wf);
    pp::pp_envelopes<<<dim3((n + 255) / 256, 8), 256, 0, st>>>(x, n, edges, wf, 0);
    float2* y = fft(x, x == a.p ? b.p : a.p, n, 8, true, st);
    pp::pp_mixdown<<<(len + 255) / 256, 256, 0, st>>>(y, n, 8, d_out, len);
    WVB_CUDA(cudaStreamSynchronize(st));
    WVB_CUDA(cudaGetLastError());
}

}  // namespace

extern "C" {

wvb_status wvb_pp_dirac_sequence(const wvb_pp_params* p, double sample_rate, double max_time, float* out,
                                 uint64_t capacity, uint64_t* count, uint32_t* events) {
    if (!p || !count) return WVB_ERR_INVALID;
    return guarded([&] {
        use_device(p->device);
        dev_buf<float> seq;
        uint32_t len = 0;
        dirac_sequence(p, sample_rate, max_time, seq, &len, events, nullptr);
        *count = len;
        if (!out) return;
        WVB_REQUIRE(capacity >= len, WVB_ERR_INVALID, "out holds %llu samples, %u needed",
                    (unsigned long long)capacity, len);
        WVB_CUDA(cudaMemcpy(out, seq.p, (size_t)len * sizeof(float), cudaMemcpyDeviceToHost));
    });
}

wvb_status wvb_pp_stochastic(const double* histogram, uint32_t n_bins, const wvb_pp_params* p, float* out,
                             uint64_t capacity, uint64_t* count, float* weighted_out) {
    if (!histogram || !p || !count) return WVB_ERR_INVALID;
    return guarded([&] {
        use_device(p->device);
        WVB_REQUIRE(p->histogram_sample_rate > 0 && p->output_sample_rate > 0 && p->acoustic_impedance > 0,
                    WVB_ERR_INVALID, "bad sample rates / impedance");
        // raytracer/postprocess.h: the sequence lasts as long as the histogram
        const double max_time = p->max_time > 0 ? p->max_time : n_bins / p->histogram_sample_rate;
        dev_buf<float> seq;
        uint32_t seq_len = 0;
        dirac_sequence(p, p->output_sample_rate, max_time, seq, &seq_len, nullptr, nullptr);
        // weight_sequence: ret.resize(min(ret.size(), convert_index(histogram.size())))
        const uint64_t ideal = (uint64_t)((double)n_bins * p->output_sample_rate / p->histogram_sample_rate);
        const uint32_t len = (uint32_t)std::min<uint64_t>(seq_len, ideal);
        *count = len;
        if (!out) return;
        WVB_REQUIRE(capacity >= len, WVB_ERR_INVALID, "out holds %llu samples, %u needed",
                    (unsigned long long)capacity, len);
        if (!len) return;
        dev_buf<double> d_hist;
        d_hist.upload(histogram, (size_t)n_bins * 8);
        dev_buf<float> d_multi, d_out;
        d_multi.alloc((size_t)len * 8, true);
        d_out.alloc(len, false);
        pp::pp_weight_sequence<<<n_bins, 128>>>(d_hist.p, n_bins, p->histogram_sample_rate, seq.p, len,
                                                p->output_sample_rate, p->acoustic_impedance, d_multi.p);
        WVB_CUDA(cudaGetLastError());
        if (weighted_out) {
            WVB_CUDA(cudaMemcpy(weighted_out, d_multi.p, (size_t)len * 8 * sizeof(float), cudaMemcpyDeviceToHost));
        }
        multiband_mixdown(d_multi.p, len, p->output_sample_rate, d_out.p, nullptr);
        WVB_CUDA(cudaMemcpy(out, d_out.p, (size_t)len * sizeof(float), cudaMemcpyDeviceToHost));
    });
}

wvb_status wvb_pp_multiband_mixdown(const float* multiband, uint64_t length, double sample_rate, int32_t device,
                                    float* out) {
    if (!multiband || !out) return WVB_ERR_INVALID;
    return guarded([&] {
        use_device(device);
        if (!length) return;
        dev_buf<float> d_in, d_out;
        d_in.upload(multiband, (size_t)length * 8);
        d_out.alloc((size_t)length, false);
        multiband_mixdown(d_in.p, (uint32_t)length, sample_rate, d_out.p, nullptr);
        WVB_CUDA(cudaMemcpy(out, d_out.p, (size_t)length * sizeof(float), cudaMemcpyDeviceToHost));
    });
}

wvb_status wvb_pp_crossover(const float* lo, uint64_t n_lo, const float* hi, uint64_t n_hi, double cutoff,
                            double width, uint64_t window_length, int32_t device, float* out, uint64_t capacity) {
    if ((!lo && n_lo) || (!hi && n_hi) || !out) return WVB_ERR_INVALID;
    return guarded([&] {
        use_device(device);
        const uint64_t len = std::max(n_lo, n_hi);  // core::sum_vectors: the longer of the two
        WVB_REQUIRE(capacity >= len, WVB_ERR_INVALID, "out holds %llu samples, %llu needed",
                    (unsigned long long)capacity, (unsigned long long)len);
        if (!len) return;
        WVB_REQUIRE(width >= 0 && width <= 1, WVB_ERR_INVALID, "Width_factor must be between 0 and 1.");
        const uint32_t n = padded_fft_length(len);
        // batch 0: the low (waveguide) signal, batch 1: the high (ray) signal
        std::vector<float> both(2 * len, 0.0f);
        for (uint64_t i = 0; i < n_lo; ++i) both[2 * i] = lo[i];
        for (uint64_t i = 0; i < n_hi; ++i) both[2 * i + 1] = hi[i];
        dev_buf<float> d_in, d_out;
        d_in.upload(both.data(), both.size());
        d_out.alloc((size_t)len, false);
        dev_buf<float2> a, b;
        a.alloc((size_t)n * 2, false);
        b.alloc((size_t)n * 2, false);
        pp::pp_load<<<dim3((n + 255) / 256, 2), 256>>>(d_in.p, (uint32_t)len, 2, a.p, n);
        float2* x = fft(a.p, b.p, n, 2, false, nullptr);
        pp::Edges e{};
        e.e[0] = cutoff;
        pp::pp_envelopes<<<dim3((n + 255) / 256, 2), 256>>>(x, n, e, width, 1);
        float2* y = fft(x, x == a.p ? b.p : a.p, n, 2, true, nullptr);
        pp::pp_sum2<<<((uint32_t)len + 255) / 256, 256>>>(y, n, (uint32_t)n_lo, (uint32_t)n_hi, d_out.p,
                                                        (uint32_t)len);
        const uint32_t w = (uint32_t)std::min<uint64_t>(window_length, len);
        if (w) pp::pp_left_hanning<<<(w + 255) / 256, 256>>>(d_out.p, w);
        WVB_CUDA(cudaGetLastError());
        WVB_CUDA(cudaMemcpy(out, d_out.p, (size_t)len * sizeof(float), cudaMemcpyDeviceToHost));
    });
}

}  // extern "C"
