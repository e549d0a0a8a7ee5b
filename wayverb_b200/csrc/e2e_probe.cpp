// e2e_probe.cpp -- bench.py's end-to-end leg THROUGH the reference-facing C++ entry point.
//
// Compiled (host-only, g++ -std=c++14) into libwvb200_probe.so next to libwvb200.so. It
// instantiates the shim's `wayverb::waveguide::run` template (reference waveguide.h:36-126)
// with the stock `preprocessor::hard_source` (hard_source.h:9-37) and
// `postprocessor::node` (node.cpp:14-18) collected by a `callback_accumulator`, exactly the
// combination `canonical.h:55-81` drives, on a synthetic cuboid mesh with HOST vectors
// (condensed nodes, coefficients, boundary index arrays built on the host and handed over as
// the reference's `mesh`), and stamps the host clock inside the callbacks so that the timed
// region is K steps of {write_value (8 B H2D), kernel launch, 4 B flag D2H, read_value
// (8 B D2H), swap} with nothing resident beforehand but the mesh.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <vector>

#include "../../include/wayverb_b200/waveguide.hpp"

using namespace wayverb;
using clk = std::chrono::steady_clock;

extern "C" {

// returns 0 on success; ms[0] = host milliseconds of the K timed steps, ms[1] = of the whole
// run() call (handle creation, mesh upload and the warm-up included); checksum = sum of the
// receiver samples of the timed steps (keeps the reads observable)
int wvb_probe_waveguide_run(int dx, int dy, int dz, const wvb_coefficients_canonical* coeffs,
                            int device, unsigned warmup, unsigned steps, size_t source_node,
                            size_t receiver_node, double ms[2], double* checksum) {
    try {
        waveguide::coefficients_canonical c{};
        for (int i = 0; i < 7; ++i) {
            c.b[i] = coeffs->b[i];
            c.a[i] = coeffs->a[i];
        }
        const auto mesh = waveguide::make_cuboid_mesh(dx, dy, dz, 0.05f, c);
        const core::compute_context cc{device};
        std::vector<double> signal(size_t(warmup) + steps, 0.0);
        signal[0] = 1.0;
        auto source = waveguide::preprocessor::make_hard_source(source_node, signal.begin(), signal.end());
        core::callback_accumulator<waveguide::postprocessor::node> receiver{receiver_node};
        clk::time_point t_begin, t_end;
        const std::atomic_bool keep_going{true};
        const auto t_call = clk::now();
        const auto done = waveguide::run(
                cc, mesh,
                [&](cl::CommandQueue& q, cl::Buffer& b, size_t step) {
                    if (step == warmup) t_begin = clk::now();
                    return source(q, b, step);
                },
                [&](cl::CommandQueue& q, const cl::Buffer& b, size_t step) {
                    receiver(q, b, step);
                    if (step + 1 == size_t(warmup) + steps) t_end = clk::now();
                },
                keep_going);
        const auto t_ret = clk::now();
        if (done != size_t(warmup) + steps) return 2;
        double sum = 0;
        for (size_t i = warmup; i < receiver.get_output().size(); ++i) sum += receiver.get_output()[i];
        if (checksum) *checksum = sum;
        ms[0] = std::chrono::duration<double, std::milli>(t_end - t_begin).count();
        ms[1] = std::chrono::duration<double, std::milli>(t_ret - t_call).count();
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "wvb_probe_waveguide_run: %s\n", e.what());
        return 1;
    }
}

}  // extern "C"
