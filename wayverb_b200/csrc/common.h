// common.h -- shared host-side plumbing for libwvb200.so (error reporting,
// CUDA call checking, a tiny parallel_for). No reference counterpart: the
// reference's equivalent is the OpenCL wrapper layer in src/core/include/core/cl/,
// which this library replaces with the CUDA runtime rather than ports.
#pragma once

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../../include/wvb200.h"

namespace wvb {

void set_last_error(const char* fmt, ...);

struct status_error {
    wvb_status code;
};

#define WVB_CUDA(expr)                                                          \
    do {                                                                        \
        cudaError_t e_ = (expr);                                                \
        if (e_ != cudaSuccess) {                                                \
            ::wvb::set_last_error("%s failed: %s (%s:%d)", #expr,               \
                                  cudaGetErrorString(e_), __FILE__, __LINE__);  \
            throw ::wvb::status_error{WVB_ERR_CUDA};                            \
        }                                                                       \
    } while (0)

#define WVB_REQUIRE(cond, code, ...)                                            \
    do {                                                                        \
        if (!(cond)) {                                                          \
            ::wvb::set_last_error(__VA_ARGS__);                                 \
            throw ::wvb::status_error{code};                                    \
        }                                                                       \
    } while (0)

// Runs body(i) for i in [0, n) on up to hardware_concurrency threads
// (contiguous blocks). Used for the one-off host-side mesh digestion.
inline void parallel_for(int64_t n, const std::function<void(int64_t)>& body) {
    unsigned hw = std::thread::hardware_concurrency();
    int64_t nt = std::max<int64_t>(1, std::min<int64_t>(hw ? hw : 4, n));
    if (nt == 1) {
        for (int64_t i = 0; i < n; ++i) body(i);
        return;
    }
    std::vector<std::thread> th;
    for (int64_t t = 0; t < nt; ++t) {
        th.emplace_back([=, &body] {
            const int64_t b = n * t / nt, e = n * (t + 1) / nt;
            for (int64_t i = b; i < e; ++i) body(i);
        });
    }
    for (auto& t : th) t.join();
}

template <typename T>
struct dev_buf {
    T* p = nullptr;
    size_t n = 0;
    dev_buf() = default;
    dev_buf(const dev_buf&) = delete;
    dev_buf& operator=(const dev_buf&) = delete;
    ~dev_buf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    void alloc(size_t count, bool zero, size_t* tally = nullptr) {
        release();
        n = count;
        if (!count) return;
        WVB_CUDA(cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T)));
        if (zero) {
            // The handles work on cudaStreamNonBlocking streams, which do not order against the
            // legacy default stream a plain cudaMemset runs on, and a device memset is
            // asynchronous to the host: wait for it here so that whatever stream touches the
            // buffer next finds it zeroed.
            WVB_CUDA(cudaMemsetAsync(p, 0, count * sizeof(T), cudaStreamLegacy));
            WVB_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
        }
        if (tally) *tally += count * sizeof(T);
    }
    void upload(const T* src, size_t count, size_t* tally = nullptr) {
        alloc(count, false, tally);
        if (count) WVB_CUDA(cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyHostToDevice));
    }
};

}  // namespace wvb
