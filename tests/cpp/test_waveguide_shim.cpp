// Mirrors src/waveguide/tests/waveguide_tests.cpp:43-139 (run_waveguide) against the
// shim: a box mesh, soft source, four postprocessor::node receivers through the
// generic callback path, the step-counter check (:105), and the same run through the
// device-side stock path; prints the receiver traces for the python test to compare
// with the oracle. Also checks the exception mapping of waveguide.h:100-119.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "wayverb_b200/waveguide.hpp"

using namespace wayverb;
using namespace wayverb::core;
using namespace wayverb::waveguide;

// nan_in_waveguide.cpp:15-71 / canonical.h:65-81 style: gaussian excitation, directional
// receiver; prints "pressure ix iy iz" per step
static int gaussian_directional(int steps, double b0) {
    const compute_context cc{};
    coefficients_canonical c{};
    c.b[0] = b0;
    c.a[0] = 1.0;
    auto m = make_cuboid_mesh(28, 26, 24, 0.05f, c);
    const auto rcv = compute_index(m.get_descriptor(), 17, 12, 10);
    callback_accumulator<postprocessor::directional_receiver> acc{m.get_descriptor(), 11776.0, 1.1765,
                                                                   rcv};
    const auto done = run(cc, m, preprocessor::gaussian{m.get_descriptor(), 0.6f, 0.65f, 0.55f, 0.1f, size_t(steps)},
                          [&](auto& queue, const auto& buffer, auto step) { acc(queue, buffer, step); }, true);
    if (done != size_t(steps)) return 7;
    for (const auto& o : acc.get_output()) {
        std::printf("%.9g %.9g %.9g %.9g\n", o.pressure, o.intensity[0], o.intensity[1], o.intensity[2]);
    }
    try {  // a receiver on the mesh edge has no six neighbours
        postprocessor::directional_receiver bad{m.get_descriptor(), 1.0, 1.0, 0};
        return 8;
    } catch (const std::runtime_error&) {
    }
    return 0;
}

// canonical.h:24-110: hard source with the calibrated unit impulse, directional receiver,
// simulation_time seconds -- once through the per-step callback path, once with nothing to
// call per step (one wvb_wg_run); both must give the same samples. Prints them.
static int canonical_run(double b0) {
    const compute_context cc{};
    coefficients_canonical c{};
    c.b[0] = b0;
    c.a[0] = 1.0;
    auto m = make_cuboid_mesh(28, 26, 24, 0.05f, c);
    const core::environment env{};
    const core::vec3 source{0.52f, 0.61f, 0.48f}, receiver{0.86f, 0.59f, 0.51f};
    const double sr = compute_sample_rate(m.get_descriptor(), env.speed_of_sound);
    const double seconds = 120.2 / sr;  // ceil -> 121 steps
    size_t calls = 0, total_seen = 0;
    const auto a = canonical(cc, m, source, receiver, env, 500.0, seconds, true,
                             [&](auto&, const auto&, auto step, auto total) {
                                 calls += (step == calls);
                                 total_seen = total;
                             });
    const auto b = canonical(cc, m, source, receiver, env, 500.0, seconds, true, no_pressure_callback{});
    if (a.size() != 1 || b.size() != 1) return 9;
    if (calls != 121 || total_seen != 121) return 10;
    const auto& x = a[0].band.directional;
    const auto& y = b[0].band.directional;
    if (x.size() != 121 || y.size() != 121 || a[0].band.sample_rate != sr || a[0].valid_hz.max != 500.0) return 11;
    for (size_t i = 0; i < x.size(); ++i) {
        if (std::memcmp(&x[i], &y[i], sizeof x[i])) return 12;  // callback path == device path
        std::printf("%.9g %.9g %.9g %.9g\n", x[i].pressure, x[i].intensity[0], x[i].intensity[1], x[i].intensity[2]);
    }
    std::printf("# source %zu receiver %zu input %.9g\n", compute_index(m.get_descriptor(), source),
                compute_index(m.get_descriptor(), receiver),
                float(rectilinear_calibration_factor(m.get_descriptor().spacing, env.acoustic_impedance)));
    try {  // a position outside the mesh
        canonical(cc, m, core::vec3{-5, 0, 0}, receiver, env, 500.0, seconds, true, no_pressure_callback{});
        return 13;
    } catch (const std::runtime_error&) {
    }
    std::atomic_bool stop{false};
    if (!canonical(cc, m, source, receiver, env, 500.0, seconds, stop, no_pressure_callback{}).empty()) return 14;
    // canonical.h:140-177, multi-band: one run per band with that band's flat coefficients; each band
    // must equal the single-band run of a mesh carrying those coefficients
    {
        struct surf { struct { float s[8]; } absorption; };
        const std::vector<surf> surfaces{{{{0.05f, 0.1f, 0.2f, 0.3f, 0.4f, 0.5f, 0.6f, 0.7f}}}};
        const multiple_band_constant_spacing_parameters params{3, 500.0, 0.6};
        const auto bands = canonical(cc, m, surfaces, source, receiver, env, params, seconds, true, no_pressure_callback{});
        if (bands.size() != 3) return 17;
        for (size_t band = 0; band < 3; ++band) {
            auto mb = m;
            mb.set_coefficients(to_flat_coefficients(surfaces[0].absorption.s[band]));
            const auto one = canonical(cc, mb, source, receiver, env, 500.0, seconds, true, no_pressure_callback{});
            if (one.size() != 1 || one[0].band.directional.size() != bands[band].band.directional.size()) return 18;
            if (std::memcmp(one[0].band.directional.data(), bands[band].band.directional.data(),
                            one[0].band.directional.size() * sizeof(one[0].band.directional[0]))) return 19;
            if (std::fabs(bands[band].valid_hz.min - 20.0 * std::pow(1000.0, band / 8.0)) > 1e-9 ||
                std::fabs(bands[band].valid_hz.max - 20.0 * std::pow(1000.0, (band + 1) / 8.0)) > 1e-9) return 20;
        }
        if (!std::memcmp(bands[0].band.directional.data(), bands[2].band.directional.data(),
                         bands[0].band.directional.size() * sizeof(bands[0].band.directional[0]))) return 21;
    }
    // cancellation DURING the device-side run: keep_going is polled every 64 steps (the reference
    // polls every step, waveguide.h:80) and the run returns the steps completed so far
    {
        util::aligned::vector<double> signal(300, 0.0), out;
        signal[0] = 1.0;
        const auto src = compute_index(m.get_descriptor(), source);
        const auto done = run_stock(cc, m, src, signal, false, {src}, out, &stop);
        if (done != 64) return 15;
        std::atomic_bool go{true};
        if (run_stock(cc, m, src, signal, false, {src}, out, &go) != 300) return 16;
    }
    return 0;
}

int main(int argc, char** argv) {
    const int steps = argc > 1 ? atoi(argv[1]) : 60;
    if (argc > 3 && std::string(argv[3]) == "gaussian") {
        return gaussian_directional(steps, strtod(argv[2], nullptr));
    }
    if (argc > 3 && std::string(argv[3]) == "canonical") return canonical_run(strtod(argv[2], nullptr));
    const compute_context cc{};
    coefficients_canonical c{};
    c.b[0] = argc > 2 ? strtod(argv[2], nullptr) : 39.0;  // impedance b0 of a flat surface
    c.a[0] = 1.0;
    auto m = make_cuboid_mesh(30, 24, 40, 0.1f, c);
    const auto source_index = compute_index(m.get_descriptor(), 15, 12, 8);
    if (!is_inside(m, source_index)) return 2;
    std::vector<double> input(steps, 0.0);
    input[0] = 1.0;
    auto prep = preprocessor::make_soft_source(source_index, input.begin(), input.end());
    std::vector<callback_accumulator<postprocessor::node>> holders;
    for (int z : {12, 18, 24, 30}) holders.emplace_back(compute_index(m.get_descriptor(), 15, 12, z));
    size_t callback_counter = 0;
    bool counter_ok = true;
    const auto done = run(
            cc, m, [&](auto& queue, auto& buffer, auto step) { return prep(queue, buffer, step); },
            [&](auto& queue, const auto& buffer, auto step) {
                for (auto& i : holders) i(queue, buffer, step);
                counter_ok = counter_ok && (step == callback_counter++);
            },
            true);
    if (done != size_t(steps) || !counter_ok) return 3;

    std::vector<double> out;
    std::vector<size_t> rcv;
    for (auto& h : holders) rcv.push_back(h.get_callback().get_output_node());
    if (run_stock(cc, m, source_index, input, true, rcv, out) != size_t(steps)) return 4;
    for (int s = 0; s < steps; ++s) {
        for (size_t r = 0; r < holders.size(); ++r) {
            const double a = holders[r].get_output()[s];
            if (a != out[s * holders.size() + r]) return 5;  // callback path == device path
            std::printf("%.17g%c", a, r + 1 == holders.size() ? '\n' : ' ');
        }
    }
    std::vector<double> bad{std::nan("")};
    try {
        run(cc, m, preprocessor::make_hard_source(source_index, bad.begin(), bad.end()),
            [](auto&, const auto&, auto) {}, true);
        return 6;
    } catch (const exceptions::value_is_nan&) {
    }
    return 0;
}
