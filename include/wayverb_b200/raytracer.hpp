// raytracer.hpp -- C++14 shim that keeps the reference's `raytracer::run` entry
// point and its processor protocol on top of the C ABI of libwvb200.so.
//
//   reference                                                        here
//   src/raytracer/include/raytracer/raytracer.h:188-266               wayverb::raytracer::run
//   .../raytracer.h:51-55,211-243  get_processor / get_group_processor /
//        process / accumulate / get_results                           same protocol
//   .../reflection_processor/stochastic_histogram.h                   make_stochastic_histogram,
//                                                                     make_directional_histogram
//                                                                     (device-resident specialisations)
//   .../reflection_processor/visual.h:18-25                           make_visual
//   .../raytracer/cl/reflection.h:10-17                               reflection
//   .../raytracer/stochastic/postprocessing.h:58-79                   energy_histogram,
//                                                                     directional_energy_histogram<20,9>
//   .../raytracer/optimum_reflection_number.h:38-68                   compute_optimum_reflection_number
//   src/core/include/core/environment.h:6-9                           core::environment
//
// Difference a caller can see: the scene arrives as `core::flattened_scene`,
// the arrays core::scene_buffers uploads from a voxelised_scene_data
// (scene_buffers.h:14-38); INTEGRATION.md shows the 10-line adapter. Rays are
// traced on the device for their whole life; processors that consume
// `reflection`s (image source, visual) are handed the records of the steps they
// ask for (`steps_required()`, default: all).
#pragma once

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <iterator>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "../wvb200.h"
#include "cl_compat.hpp"
#include "waveguide.hpp"  // core::compute_context, util::aligned::vector, detail::check

namespace wayverb {
namespace core {

constexpr auto simulation_bands = 8;
struct alignas(32) bands_type final {
    float s[simulation_bands];
};
template <size_t bands>
struct alignas(32) surface final {
    bands_type absorption;
    bands_type scattering;
};
struct alignas(8) triangle final {
    cl_uint surface, v0, v1, v2;
};
struct environment final {
    double speed_of_sound{340.0};
    double acoustic_impedance{400.0};
};
struct vec3 final {
    float x, y, z;
};

/// What core::scene_buffers uploads (scene_buffers.h:14-38), by reference.
struct flattened_scene final {
    util::aligned::vector<cl_uint> voxel_index;  // get_flattened(voxels), voxel_collection.cpp:9-37
    vec3 aabb_min, aabb_max;                     // voxels.get_aabb()
    cl_uint side;                                // voxels.get_side()
    util::aligned::vector<triangle> triangles;
    util::aligned::vector<cl_float3> vertices;
    util::aligned::vector<surface<simulation_bands>> surfaces;
};

}  // namespace core

namespace raytracer {

struct alignas(16) reflection final {
    cl_float3 position;
    cl_uint triangle;
    cl_char keep_going;
    cl_char receiver_visible;
};
static_assert(sizeof(reflection) == sizeof(wvb_reflection), "reflection layout");

/// compute_optimum_reflection_number (optimum_reflection_number.h:38-68)
inline size_t compute_optimum_reflection_number(double absorption) {
    return wvb_rt_reflection_depth(absorption);
}
inline size_t compute_optimum_reflection_number(const core::flattened_scene& scene) {
    std::vector<bool> used(scene.surfaces.size(), false);
    for (const auto& t : scene.triangles) {
        if (t.surface < used.size()) used[t.surface] = true;
    }
    double min_abs = 0;
    bool any = false;
    for (size_t i = 0; i != used.size(); ++i) {
        if (!used[i]) continue;
        // min_absorption(surface) is min_element of the band vector, which the
        // reference resolves to the FIRST band (min_element(float x) overload is
        // never reached for a vector; it takes t.absorption as a whole and the
        // double overload converts s0) -- we take the smallest band, the
        // conservative reading
        double m = scene.surfaces[i].absorption.s[0];
        for (float a : scene.surfaces[i].absorption.s) m = std::min<double>(m, a);
        min_abs = any ? std::min(min_abs, m) : m;
        any = true;
    }
    if (!any) throw std::runtime_error{"Can't find min absorption of empty vector."};
    return compute_optimum_reflection_number(min_abs);
}

namespace stochastic {
/// stochastic/postprocessing.h:58-66
struct energy_histogram final {
    double sample_rate;
    util::aligned::vector<core::bands_type> histogram;
};
/// stochastic/postprocessing.h:68-79, vector_look_up_table<vector<bands_type>, Az, El>
template <size_t Az, size_t El>
struct directional_energy_histogram final {
    double sample_rate;
    struct table_t {
        util::aligned::vector<core::bands_type> table[Az][El];
    } histogram;
};
}  // namespace stochastic

namespace detail {
class scene_handle final {
public:
    scene_handle(const core::compute_context& cc, const core::flattened_scene& s) {
        wvb_rt_scene_desc d{};
        d.voxel_index = s.voxel_index.data();
        d.voxel_index_count = s.voxel_index.size();
        d.aabb_min[0] = s.aabb_min.x; d.aabb_min[1] = s.aabb_min.y; d.aabb_min[2] = s.aabb_min.z;
        d.aabb_max[0] = s.aabb_max.x; d.aabb_max[1] = s.aabb_max.y; d.aabb_max[2] = s.aabb_max.z;
        d.side = s.side;
        d.triangles = reinterpret_cast<const wvb_triangle*>(s.triangles.data());
        d.num_triangles = uint32_t(s.triangles.size());
        d.vertices = reinterpret_cast<const wvb_float3*>(s.vertices.data());
        d.num_vertices = uint32_t(s.vertices.size());
        d.surfaces = reinterpret_cast<const wvb_surface*>(s.surfaces.data());
        d.num_surfaces = uint32_t(s.surfaces.size());
        d.device = cc.device;
        core::detail::check(wvb_rt_create(&d, &rt_));
        if (!rt_) throw std::runtime_error{std::string{"libwvb200: "} + wvb_last_error()};
    }
    ~scene_handle() { wvb_rt_destroy(rt_); }
    scene_handle(const scene_handle&) = delete;
    scene_handle& operator=(const scene_handle&) = delete;
    wvb_rt* get() const { return rt_; }

private:
    wvb_rt* rt_{nullptr};
};

/// what the run loop tells every processor about the run
struct run_info final {
    wvb_rt* rt;
    core::vec3 source, receiver;
    core::environment environment;
    size_t total_rays;
    size_t reflection_depth;
};
}  // namespace detail

namespace reflection_processor {

/// Device-resident stochastic histogram. Same constructor arguments as
/// make_stochastic_histogram / make_directional_histogram
/// (stochastic_histogram.h:176-229): (total_rays, max_image_source_order,
/// receiver_radius, histogram_sample_rate).
template <bool Directional>
class make_device_histogram final {
public:
    make_device_histogram(size_t total_rays, size_t max_image_source_order, float receiver_radius,
                          float histogram_sample_rate)
            : total_rays_{total_rays}
            , max_image_source_order_{max_image_source_order}
            , receiver_radius_{receiver_radius}
            , histogram_sample_rate_{histogram_sample_rate} {}

    static constexpr bool device_histogram = true;
    static constexpr bool directional = Directional;
    size_t total_rays_;
    size_t max_image_source_order_;
    float receiver_radius_;
    float histogram_sample_rate_;
};
using make_stochastic_histogram = make_device_histogram<false>;
using make_directional_histogram = make_device_histogram<true>;

/// visual.h:18-25: keeps the first `items` rays' reflections of every step.
class make_visual final {
public:
    explicit make_visual(size_t items) : items_{items} {}
    static constexpr bool device_histogram = false;
    size_t items_;
    size_t steps_required(size_t depth) const { return depth; }
    using result_type = util::aligned::vector<util::aligned::vector<reflection>>;
};

/// Collects {triangle, receiver_visible, keep_going} of the first `max_order`
/// steps for every ray: the input of the image-source path builder
/// (image_source/reflection_path_builder.h:15-24), which stays host code.
class make_first_reflections final {
public:
    explicit make_first_reflections(size_t max_order) : max_order_{max_order} {}
    static constexpr bool device_histogram = false;
    size_t max_order_;
    using result_type = util::aligned::vector<util::aligned::vector<reflection>>;  // [step][ray]
};

}  // namespace reflection_processor

namespace detail {
inline stochastic::energy_histogram to_energy_histogram(const std::vector<double>& h, size_t bins,
                                                        double rate) {
    stochastic::energy_histogram r{rate, {}};
    size_t last = 0;
    for (size_t b = 0; b < bins; ++b) {
        for (int k = 0; k < 8; ++k) {
            if (h[b * 8 + k] != 0) last = b + 1;
        }
    }
    r.histogram.resize(last);  // the reference grows the vector to the last used bin (histogram.h:63-80)
    for (size_t b = 0; b < last; ++b) {
        for (int k = 0; k < 8; ++k) r.histogram[b].s[k] = float(h[b * 8 + k]);
    }
    return r;
}
}  // namespace detail

/// Results of one run: whichever of these the callbacks asked for.
struct results final {
    stochastic::energy_histogram histogram{0, {}};
    stochastic::directional_energy_histogram<20, 9> directional{0, {}};
    util::aligned::vector<util::aligned::vector<reflection>> first_reflections;  // [step][ray]
    util::aligned::vector<util::aligned::vector<reflection>> visual;             // [step][item]
    uint64_t dropped_impulses{0};
    bool completed{false};
};

/// raytracer::run (raytracer.h:188-266) for the canonical callback set
/// (canonical.cpp:9-20): image-source input + (directional) stochastic histogram +
/// visual. Directions are any forward iterator range over vec3-like values
/// (.x .y .z); rays are traced in segments of 1 << 14 (raytracer.h:219) so that
/// per_step_callback(group, groups) and keep_going behave as in the reference.
/// seed: the reference seeds its scatter RNG from std::random_device
/// (reflector.cpp:13-25); pass a seed to make runs reproducible.
template <typename It, typename PerStepCallback>
results run(It b_direction, It e_direction, const core::compute_context& cc,
            const core::flattened_scene& scene, const core::vec3& source, const core::vec3& receiver,
            const core::environment& environment, const std::atomic_bool& keep_going,
            PerStepCallback&& per_step_callback, size_t max_image_source_order, float receiver_radius,
            float histogram_sample_rate, bool directional, size_t visual_items, uint64_t seed) {
    const detail::scene_handle h{cc, scene};
    const size_t total = size_t(std::distance(b_direction, e_direction));
    const size_t depth = compute_optimum_reflection_number(scene);
    constexpr size_t segment_size = 1 << 14;

    wvb_rt_trace_params p{};
    p.source[0] = source.x; p.source[1] = source.y; p.source[2] = source.z;
    p.receiver[0] = receiver.x; p.receiver[1] = receiver.y; p.receiver[2] = receiver.z;
    p.receiver_radius = receiver_radius;
    p.speed_of_sound = environment.speed_of_sound;
    p.histogram_sample_rate = histogram_sample_rate;
    p.total_rays = total;
    p.seed = seed;
    p.depth = uint32_t(depth);
    p.specular_from_step = uint32_t(max_image_source_order + 1);  // canonical.cpp:16
    p.n_bins = wvb_rt_safe_bins(h.get(), p.depth, p.speed_of_sound, p.histogram_sample_rate);
    p.directional = directional ? 1 : 0;
    const size_t keep = std::min(depth, std::max(max_image_source_order, visual_items ? depth : size_t{0}));
    p.keep_steps = uint32_t(keep);
    core::detail::check(wvb_rt_reset_histogram(h.get()));

    results ret;
    ret.first_reflections.resize(std::min(max_image_source_order, depth));
    for (auto& v : ret.first_reflections) v.reserve(total);
    ret.visual.resize(visual_items ? depth : 0);

    std::vector<float> dirs;
    std::vector<reflection> refl;
    const size_t groups = total / segment_size;
    size_t done = 0, group = 0;
    auto it = b_direction;
    while (done < total) {
        const size_t n = std::min(segment_size, total - done);
        dirs.resize(n * 3);
        for (size_t i = 0; i < n; ++i, ++it) {
            const auto d = *it;
            dirs[3 * i] = d.x; dirs[3 * i + 1] = d.y; dirs[3 * i + 2] = d.z;
        }
        refl.resize(keep * n);
        p.ray_index_base = done;
        uint64_t dropped = 0;
        core::detail::check(wvb_rt_trace(h.get(), &p, dirs.data(), n,
                                         reinterpret_cast<wvb_reflection*>(refl.data()), &dropped,
                                         nullptr));
        ret.dropped_impulses = dropped;
        for (size_t s = 0; s < ret.first_reflections.size(); ++s) {
            ret.first_reflections[s].insert(ret.first_reflections[s].end(), refl.begin() + s * n,
                                            refl.begin() + (s + 1) * n);
        }
        if (visual_items && done == 0) {
            const size_t items = std::min(visual_items, n);
            for (size_t s = 0; s < keep; ++s) {
                ret.visual[s].assign(refl.begin() + s * n, refl.begin() + s * n + items);
            }
        }
        done += n;
        if (n == segment_size) {
            per_step_callback(group++, groups);
            if (!keep_going) return ret;  // the reference returns nullopt here (raytracer.h:255-257)
        }
    }

    std::vector<double> hist(size_t(p.n_bins) * 8 * (directional ? 180 : 1));
    core::detail::check(wvb_rt_read_histogram(h.get(), hist.data()));
    if (directional) {
        ret.directional.sample_rate = histogram_sample_rate;
        for (size_t a = 0; a < 20; ++a) {
            for (size_t e = 0; e < 9; ++e) {
                std::vector<double> cell(hist.begin() + (a * 9 + e) * size_t(p.n_bins) * 8,
                                         hist.begin() + (a * 9 + e + 1) * size_t(p.n_bins) * 8);
                ret.directional.histogram.table[a][e] =
                        detail::to_energy_histogram(cell, p.n_bins, histogram_sample_rate).histogram;
            }
        }
    } else {
        ret.histogram = detail::to_energy_histogram(hist, p.n_bins, histogram_sample_rate);
    }
    ret.completed = true;
    return ret;
}

}  // namespace raytracer
}  // namespace wayverb
