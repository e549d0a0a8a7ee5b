// Drives wayverb::raytracer::run (the shim of raytracer.h:188-266) the way
// raytracer/canonical.h:44-75 does: random directions, image-source order 4,
// directional histogram off/on, 32 visual rays; checks the protocol-level
// behaviour (segments, per-step callback, keep_going, result shapes, energy).
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <tuple>
#include <random>
#include <string>
#include <vector>

#include "wayverb_b200/raytracer.hpp"

using namespace wayverb;

static core::flattened_scene shoebox(float sx, float sy, float sz, unsigned side) {
    core::flattened_scene s;
    const float X[2] = {0, sx}, Y[2] = {0, sy}, Z[2] = {0, sz};
    for (int i = 0; i < 8; ++i) s.vertices.push_back(cl_float3{{X[i & 1], Y[(i >> 1) & 1], Z[(i >> 2) & 1], 0}});
    const unsigned q[6][4] = {{0, 1, 3, 2}, {4, 6, 7, 5}, {0, 4, 5, 1}, {2, 3, 7, 6}, {0, 2, 6, 4}, {1, 5, 7, 3}};
    for (auto& f : q) {
        s.triangles.push_back(core::triangle{0, f[0], f[1], f[2]});
        s.triangles.push_back(core::triangle{0, f[0], f[2], f[3]});
    }
    core::surface<8> surf{};
    for (int b = 0; b < 8; ++b) { surf.absorption.s[b] = 0.1f; surf.scattering.s[b] = 0.1f; }
    s.surfaces.push_back(surf);
    s.aabb_min = {-0.1f, -0.1f, -0.1f};
    s.aabb_max = {sx + 0.1f, sy + 0.1f, sz + 0.1f};
    s.side = side;
    // every triangle in every voxel: a legal (if slow) superset assignment
    const unsigned cells = side * side * side;
    s.voxel_index.resize(cells);
    for (unsigned c = 0; c < cells; ++c) {
        s.voxel_index[c] = cl_uint(s.voxel_index.size());
        s.voxel_index.push_back(cl_uint(s.triangles.size()));
        for (unsigned t = 0; t < s.triangles.size(); ++t) s.voxel_index.push_back(t);
    }
    return s;
}

int main(int argc, char** argv) {
    const auto scene = shoebox(5.56f, 3.97f, 2.81f, 4);
    const core::compute_context cc{};
    const core::vec3 source{1, 1, 1}, receiver{2, 3, 1.5f};
    std::mt19937 eng{7};
    std::uniform_real_distribution<float> uz(-1, 1), ut(-3.14159265f, 3.14159265f);
    const size_t rays = (1 << 15) + 1000;  // two full segments + a tail
    std::vector<core::vec3> dirs(rays);
    for (auto& d : dirs) {
        const float z = uz(eng), th = ut(eng), t = std::sqrt(1 - z * z);
        d = {t * std::cos(th), z, t * std::sin(th)};
    }
    int calls = 0;
    auto res = raytracer::run(dirs.begin(), dirs.end(), cc, scene, source, receiver, core::environment{},
                              true, [&](auto group, auto groups) { calls += (groups == 2 && group == size_t(calls)); },
                              4, 0.1f, 1000.0f, false, 32, 1234);
    if (!res.completed || calls != 2) return 2;
    if (raytracer::compute_optimum_reflection_number(scene) != 132) return 3;
    if (res.first_reflections.size() != 4 || res.first_reflections[0].size() != rays) return 4;
    if (res.visual.size() != 132 || res.visual[0].size() != 32) return 5;
    for (auto& r : res.first_reflections[0]) {
        if (!r.keep_going) return 6;  // closed box: nobody escapes
    }
    double total = 0;
    for (auto& b : res.histogram.histogram) {
        for (float v : b.s) {
            if (!(v >= 0) || !std::isfinite(v)) return 7;
            total += v;
        }
    }
    if (!(total > 0) || res.histogram.histogram.empty() || res.dropped_impulses) return 8;

    // directional run sums to the same energy
    auto dres = raytracer::run(dirs.begin(), dirs.end(), cc, scene, source, receiver, core::environment{},
                               true, [](auto, auto) {}, 4, 0.1f, 1000.0f, true, 0, 1234);
    double dtotal = 0;
    for (auto& row : dres.directional.histogram.table) {
        for (auto& cell : row) {
            for (auto& b : cell) {
                for (float v : b.s) dtotal += v;
            }
        }
    }
    if (std::fabs(dtotal / total - 1) > 1e-4) return 9;

    // keep_going = false stops after the first full segment (raytracer.h:255-257)
    std::atomic_bool stop{false};
    int seen = 0;
    auto partial = raytracer::run(dirs.begin(), dirs.end(), cc, scene, source, receiver, core::environment{},
                                  stop, [&](auto, auto) { ++seen; }, 4, 0.1f, 1000.0f, false, 0, 1);
    if (partial.completed || seen != 1) return 10;
    // the reference's own signature: a tuple of callbacks -> optional<tuple<results...>>
    // (canonical.cpp:9-20: image-source input + directional histogram + visual)
    {
        int steps_seen = 0;
        auto tup = raytracer::run(
                dirs.begin(), dirs.end(), cc, scene, source, receiver, core::environment{}, true,
                [&](auto, auto) { ++steps_seen; },
                std::make_tuple(raytracer::reflection_processor::make_image_source_input{4},
                                raytracer::reflection_processor::make_stochastic_histogram{rays, 4 + 1, 0.1f, 1000.0f},
                                raytracer::reflection_processor::make_visual{32}),
                1234);
        if (!tup || steps_seen != 2) return 11;
        const auto& paths = std::get<0>(*tup);
        const auto& hist = std::get<1>(*tup);
        const auto& visual = std::get<2>(*tup);
        if (paths.size() != rays || paths[0].size() != 4) return 12;
        if (visual.size() != 132 || visual[0].size() != 32) return 13;  // no bound declared: every step
        // same seed, same directions, same gate (order + 1) as the explicit-parameter run above
        if (hist.histogram.size() != res.histogram.histogram.size()) return 14;
        for (size_t b = 0; b < hist.histogram.size(); ++b) {
            for (int k = 0; k < 8; ++k) {
                const float x = hist.histogram[b].s[k], y = res.histogram.histogram[b].s[k];
                if (std::fabs(x - y) > 1e-6f * std::fabs(y) + 1e-30f) return 15;
            }
        }
        for (size_t i = 0; i < rays; i += 997) {
            if (std::memcmp(&paths[i][0], &res.first_reflections[0][i], sizeof(raytracer::reflection))) return 16;
        }
    }
    // canonical.cpp:9-20 as the reference spells it: make_image_source evaluated on the device
    {
        auto tup = raytracer::run(
                dirs.begin(), dirs.end(), cc, scene, source, receiver, core::environment{}, true,
                [&](auto, auto) {},
                std::make_tuple(raytracer::reflection_processor::make_image_source{4},
                                raytracer::reflection_processor::make_directional_histogram{rays, 4 + 1, 0.1f, 1000.0f},
                                raytracer::reflection_processor::make_visual{32}),
                1234);
        if (!tup) return 17;
        const auto& imps = std::get<0>(*tup);
        if (imps.size() < 7) return 18;  // six first-order images + direct at the very least
        const auto& direct = imps.back();  // image_source.cpp:53-58
        if (direct.position.s[0] != source.x || direct.position.s[1] != source.y || direct.position.s[2] != source.z) return 19;
        bool saw_x0_image = false;
        for (const auto& i : imps) {
            const float dx = i.position.s[0] - receiver.x, dy = i.position.s[1] - receiver.y, dz = i.position.s[2] - receiver.z;
            if (std::fabs(std::sqrt(dx * dx + dy * dy + dz * dz) - i.distance) > 1e-4f) return 20;
            saw_x0_image |= std::fabs(i.position.s[0] + source.x) < 1e-5f && std::fabs(i.position.s[1] - source.y) < 1e-5f &&
                            std::fabs(i.position.s[2] - source.z) < 1e-5f;
        }
        if (!saw_x0_image) return 21;
        if (std::get<2>(*tup).size() != 132) return 22;
        if (argc > 1) {  // hand directions and impulses to the Python test, which re-derives them with the oracle
            const std::string dir = argv[1];
            FILE* f = std::fopen((dir + "/dirs.f32").c_str(), "wb");
            if (!f) return 23;
            std::fwrite(dirs.data(), sizeof(core::vec3), dirs.size(), f);
            std::fclose(f);
            f = std::fopen((dir + "/impulses.bin").c_str(), "wb");
            if (!f) return 23;
            std::fwrite(imps.data(), sizeof(imps[0]), imps.size(), f);
            std::fclose(f);
        }
        std::printf("IMAGE_SOURCES %zu\n", imps.size());
    }
    if (argc > 5) {  // image_source::run (image_source/run.h:12-47): compile check only; the same calls with
                     // order = depth run through the C ABI in tests/test_is_gpu.py (exact shoebox KAT)
        const auto imps = raytracer::image_source::run(dirs.begin(), dirs.begin() + 100, cc, scene, source, receiver,
                                                       core::environment{}, 1);
        std::printf("%zu\n", imps.size());
    }
    std::printf("RT_SHIM_OK energy=%g bins=%zu\n", total, res.histogram.histogram.size());
    return 0;
}
