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
#include <experimental/optional>
#include <limits>
#include <random>
#include <type_traits>
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
// core::environment and core::vec3 live in waveguide.hpp (both paths use them)

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

/// raytracer/cl/structs.h:37-44
template <size_t channels>
struct alignas(1 << 5) impulse final {
    core::bands_type volume;  // channels == simulation_bands == 8 is the only instantiation in use
    cl_float3 position;
    cl_float distance;
};
static_assert(sizeof(impulse<8>) == sizeof(wvb_impulse), "impulse<8> layout");

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
        // min_absorption(surface) = min_element(surface.absorption): argument-dependent lookup takes the
        // global min_element of core/cl/traits.h for the band vector, i.e. the smallest band
        // (held against the reference's own header in tests/cpp/test_hostmath_pin.cpp)
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

// =====================================================================================
// The reference's own signature: run(b, e, cc, voxelised, source, receiver, environment,
// keep_going, per_step_callback, callbacks) -> optional<tuple<results...>>
// (raytracer.h:188-266) with the tuple-of-processors protocol (:51-55,211-243):
//   callbacks[k].get_processor(cc, source, receiver, environment, voxelised)
//   processor.get_group_processor(num_directions)
//   group.process(begin, end, buffers, step, total)     // once per step per segment
//   processor.accumulate(group)                          // once per segment
//   processor.get_results()
// Host processors written against the reference (image source, visual, ...) work
// unchanged: after a segment has been traced on the device they are fed the
// `reflection`s of each step they ask for. A processor may declare
// `size_t steps_required(size_t depth) const` to bound what is copied back (the
// reference's image-source processor only uses step < max_order, visual every
// step). Processors that mark themselves `device_resident` (the stochastic
// histograms below) are evaluated inside the trace kernel instead.
// =====================================================================================
}  // namespace raytracer

namespace core {
/// Stands in for core::scene_buffers (spatial_division/scene_buffers.h:11-60) in the
/// processors' `process(b, e, buffers, step, total)` signature.
struct scene_buffers final {
    wvb_rt* handle;
};
}  // namespace core

namespace raytracer {
namespace reflection_processor {

/// Device-resident stochastic histogram following the processor protocol; same
/// constructor arguments as the reference's make_stochastic_histogram /
/// make_directional_histogram (stochastic_histogram.h:176-229).
template <bool Directional>
class device_histogram_processor final {
public:
    static constexpr bool device_resident = true;
    static constexpr bool directional = Directional;
    using result_type = typename std::conditional<Directional, stochastic::directional_energy_histogram<20, 9>,
                                                  stochastic::energy_histogram>::type;
    device_histogram_processor(size_t total_rays, size_t max_image_source_order, float receiver_radius,
                               float histogram_sample_rate)
            : total_rays{total_rays}
            , max_image_source_order{max_image_source_order}
            , receiver_radius{receiver_radius}
            , histogram_sample_rate{histogram_sample_rate} {}
    struct group final {
        template <typename It>
        void process(It, It, const core::scene_buffers&, size_t, size_t) {}
    };
    group get_group_processor(size_t) const { return {}; }
    void accumulate(const group&) {}
    result_type get_results() const { return results; }

    size_t total_rays, max_image_source_order;
    float receiver_radius, histogram_sample_rate;
    result_type results{};
};

template <bool Directional>
class make_histogram final {  // see the aliases below
public:
    make_histogram(size_t total_rays, size_t max_image_source_order, float receiver_radius,
                   float histogram_sample_rate)
            : p_{total_rays, max_image_source_order, receiver_radius, histogram_sample_rate} {}
    template <typename Scene>
    device_histogram_processor<Directional> get_processor(const core::compute_context&, const core::vec3&,
                                                          const core::vec3&, const core::environment&,
                                                          const Scene&) const {
        return p_;
    }

private:
    device_histogram_processor<Directional> p_;
};

using make_stochastic_histogram = make_histogram<false>;
using make_directional_histogram = make_histogram<true>;

/// visual.h:18-71 restated on the protocol: the first `items` reflections of every
/// step, from the first segment only (visual.cpp accumulate keeps the first).
class visual_group_processor final {
public:
    explicit visual_group_processor(size_t items) : items_{items} {}
    template <typename It>
    void process(It b, It e, const core::scene_buffers&, size_t, size_t) {
        const size_t n = std::min<size_t>(items_, size_t(std::distance(b, e)));
        data_.emplace_back(b, b + n);
    }
    const util::aligned::vector<util::aligned::vector<reflection>>& get_results() const { return data_; }

private:
    size_t items_;
    util::aligned::vector<util::aligned::vector<reflection>> data_;  // [step][item]
};
class visual_processor final {
public:
    explicit visual_processor(size_t items) : items_{items} {}
    visual_group_processor get_group_processor(size_t) const { return visual_group_processor{items_}; }
    void accumulate(const visual_group_processor& g) {
        if (results_.empty()) results_ = g.get_results();
    }
    util::aligned::vector<util::aligned::vector<reflection>> get_results() const { return results_; }

private:
    size_t items_;
    util::aligned::vector<util::aligned::vector<reflection>> results_;
};
class make_visual final {
public:
    explicit make_visual(size_t items) : items_{items} {}
    template <typename Scene>
    visual_processor get_processor(const core::compute_context&, const core::vec3&, const core::vec3&,
                                   const core::environment&, const Scene&) const {
        return visual_processor{items_};
    }

private:
    size_t items_;
};

/// The input of the image-source model (reflection_processor/image_source.h:17-26 pushes
/// the reflections of steps < max_order into reflection_path_builder): per ray, the
/// reflections of its first max_order steps. The path tree / validation stays host code.
class first_reflections_group_processor final {
public:
    first_reflections_group_processor(size_t max_order, size_t items) : max_order_{max_order}, paths_(items) {}
    size_t steps_required(size_t depth) const { return std::min(max_order_, depth); }
    template <typename It>
    void process(It b, It e, const core::scene_buffers&, size_t step, size_t) {
        if (step < max_order_) {
            size_t i = 0;
            for (auto it = b; it != e; ++it, ++i) paths_[i].push_back(*it);
        }
    }
    const util::aligned::vector<util::aligned::vector<reflection>>& get_results() const { return paths_; }

private:
    size_t max_order_;
    util::aligned::vector<util::aligned::vector<reflection>> paths_;  // [ray][step]
};
class first_reflections_processor final {
public:
    explicit first_reflections_processor(size_t max_order) : max_order_{max_order} {}
    size_t steps_required(size_t depth) const { return std::min(max_order_, depth); }
    first_reflections_group_processor get_group_processor(size_t n) const { return {max_order_, n}; }
    void accumulate(const first_reflections_group_processor& g) {
        results_.insert(results_.end(), g.get_results().begin(), g.get_results().end());
    }
    util::aligned::vector<util::aligned::vector<reflection>> get_results() const { return results_; }

private:
    size_t max_order_;
    util::aligned::vector<util::aligned::vector<reflection>> results_;
};
class make_image_source_input final {
public:
    explicit make_image_source_input(size_t max_order) : max_order_{max_order} {}
    template <typename Scene>
    first_reflections_processor get_processor(const core::compute_context&, const core::vec3&,
                                              const core::vec3&, const core::environment&,
                                              const Scene&) const {
        return first_reflections_processor{max_order_};
    }

private:
    size_t max_order_;
};

/// reflection_processor::make_image_source (reflection_processor/image_source.h:15-86)
/// evaluated on the device: the first `max_order` reflections of every ray feed the
/// path tree straight from the trace kernel, validation and pressure calculation run
/// as kernels (wvb_is_*), get_results() returns the reference's impulses in the
/// reference's order. Same constructor as the reference.
class device_image_source_processor final {
public:
    static constexpr bool device_resident = true;
    explicit device_image_source_processor(size_t max_order) : max_order{max_order} {}
    struct group final {
        template <typename It>
        void process(It, It, const core::scene_buffers&, size_t, size_t) {}
    };
    group get_group_processor(size_t) const { return {}; }
    void accumulate(const group&) {}
    util::aligned::vector<impulse<8>> get_results() const { return results; }

    size_t max_order;
    util::aligned::vector<impulse<8>> results;
};
class make_image_source final {
public:
    explicit make_image_source(size_t max_order) : max_order_{max_order} {}
    template <typename Scene>
    device_image_source_processor get_processor(const core::compute_context&, const core::vec3&,
                                                const core::vec3&, const core::environment&,
                                                const Scene&) const {
        return device_image_source_processor{max_order_};
    }

private:
    size_t max_order_;
};

}  // namespace reflection_processor

namespace detail {

template <typename T, typename = void>
struct is_device_resident : std::false_type {};
template <typename T>
struct is_device_resident<T, typename std::enable_if<T::device_resident>::type> : std::true_type {};

template <typename T>
auto steps_required_of(const T& t, size_t depth, int) -> decltype(t.steps_required(depth)) {
    return t.steps_required(depth);
}
template <typename T>
size_t steps_required_of(const T&, size_t depth, long) {
    return depth;
}

template <typename Tuple, typename F, size_t... Ix>
void for_each_in(Tuple& t, F&& f, std::index_sequence<Ix...>) {
    using expand = int[];
    (void)expand{0, ((void)f(std::get<Ix>(t), std::integral_constant<size_t, Ix>{}), 0)...};
}
template <typename... Ts, typename F>
void for_each_in(std::tuple<Ts...>& t, F&& f) {
    for_each_in(t, std::forward<F>(f), std::index_sequence_for<Ts...>{});
}

// device-histogram parameters found among the processors (at most one is honoured)
struct histogram_request final {
    bool present{false};
    bool directional{false};
    size_t max_order{0};
    float radius{0.1f};
    float rate{1000.0f};
};
template <bool D>
void note_histogram(const reflection_processor::device_histogram_processor<D>& p, histogram_request& r) {
    if (!r.present) r = {true, D, p.max_image_source_order, p.receiver_radius, p.histogram_sample_rate};
}
template <typename T>
void note_histogram(const T&, histogram_request&) {}

inline void store_histogram(reflection_processor::device_histogram_processor<false>& p,
                            const std::vector<double>& h, size_t bins) {
    p.results = to_energy_histogram(h, bins, p.histogram_sample_rate);
}
inline void store_histogram(reflection_processor::device_histogram_processor<true>& p,
                            const std::vector<double>& h, size_t bins) {
    p.results.sample_rate = p.histogram_sample_rate;
    for (size_t a = 0; a < 20; ++a) {
        for (size_t e = 0; e < 9; ++e) {
            std::vector<double> cell(h.begin() + (a * 9 + e) * bins * 8, h.begin() + (a * 9 + e + 1) * bins * 8);
            p.results.histogram.table[a][e] = to_energy_histogram(cell, bins, p.histogram_sample_rate).histogram;
        }
    }
}
template <typename T>
void store_histogram(T&, const std::vector<double>&, size_t) {}

// image-source request found among the processors (at most one is honoured)
struct image_source_request final {
    bool present{false};
    size_t max_order{0};
};
inline void note_image_source(const reflection_processor::device_image_source_processor& p,
                              image_source_request& r) {
    if (!r.present) r = {true, p.max_order};
}
template <typename T>
void note_image_source(const T&, image_source_request&) {}
inline void store_image_source(reflection_processor::device_image_source_processor& p,
                               const util::aligned::vector<impulse<8>>& v) {
    p.results = v;
}
template <typename T>
void store_image_source(T&, const util::aligned::vector<impulse<8>>&) {}

struct is_handle final {
    wvb_is* h{nullptr};
    is_handle() = default;
    is_handle(const is_handle&) = delete;
    is_handle& operator=(const is_handle&) = delete;
    ~is_handle() {
        if (h) wvb_is_destroy(h);
    }
};

}  // namespace detail

namespace detail {
template <typename Tuple, size_t... Ix>
auto make_processors(Tuple& cbs, std::index_sequence<Ix...>, const core::compute_context& cc,
                     const core::vec3& source, const core::vec3& receiver, const core::environment& env,
                     const core::flattened_scene& scene) {
    return std::make_tuple(std::get<Ix>(cbs).get_processor(cc, source, receiver, env, scene)...);
}
template <typename Tuple, size_t... Ix>
auto collect_results(Tuple& procs, std::index_sequence<Ix...>) {
    return std::make_tuple(std::get<Ix>(procs).get_results()...);
}
}  // namespace detail

template <typename It, typename PerStepCallback, typename... Callbacks>
auto run(It b_direction, It e_direction, const core::compute_context& cc,
         const core::flattened_scene& voxelised, const core::vec3& source, const core::vec3& receiver,
         const core::environment& environment, const std::atomic_bool& keep_going,
         PerStepCallback&& per_step_callback, std::tuple<Callbacks...> callbacks,
         uint64_t seed = std::random_device{}()) {
    const detail::scene_handle h{cc, voxelised};
    const core::scene_buffers buffers{h.get()};
    const auto seq = std::index_sequence_for<Callbacks...>{};

    auto processors = detail::make_processors(callbacks, seq, cc, source, receiver, environment, voxelised);
    using return_type = decltype(detail::collect_results(processors, seq));

    const size_t total = size_t(std::distance(b_direction, e_direction));
    constexpr size_t segment_size = 1 << 14;                               // raytracer.h:219
    const size_t depth = compute_optimum_reflection_number(voxelised);     // :220-221

    detail::histogram_request hist;
    detail::image_source_request img;
    size_t keep = 0;
    detail::for_each_in(processors, [&](auto& p, auto) {
        detail::note_histogram(p, hist);
        detail::note_image_source(p, img);
        if (!detail::is_device_resident<std::decay_t<decltype(p)>>::value) {
            keep = std::max(keep, detail::steps_required_of(p, depth, 0));
        }
    });

    wvb_rt_trace_params p{};
    p.source[0] = source.x; p.source[1] = source.y; p.source[2] = source.z;
    p.receiver[0] = receiver.x; p.receiver[1] = receiver.y; p.receiver[2] = receiver.z;
    p.receiver_radius = hist.radius;
    p.speed_of_sound = environment.speed_of_sound;
    p.histogram_sample_rate = hist.rate;
    p.total_rays = total;
    p.seed = seed;
    p.depth = uint32_t(depth);
    p.specular_from_step = uint32_t(hist.max_order);  // stochastic_histogram.h:98-101
    p.n_bins = hist.present ? wvb_rt_safe_bins(h.get(), p.depth, p.speed_of_sound, p.histogram_sample_rate) : 1;
    p.directional = hist.directional ? 1 : 0;
    p.keep_steps = uint32_t(keep);
    core::detail::check(wvb_rt_reset_histogram(h.get()));

    detail::is_handle is;
    const size_t is_order = std::min(img.max_order, depth);
    if (img.present) {
        wvb_is_desc d{};
        d.source[0] = source.x; d.source[1] = source.y; d.source[2] = source.z;
        d.receiver[0] = receiver.x; d.receiver[1] = receiver.y; d.receiver[2] = receiver.z;
        d.acoustic_impedance = environment.acoustic_impedance;
        d.flip_phase = 0;   // image_source.cpp:49
        d.with_direct = 1;  // image_source.cpp:53-58
        d.max_elements = std::max<uint64_t>(1, uint64_t(total) * is_order);
        core::detail::check(wvb_is_create(h.get(), &d, &is.h));
    }

    // The reference traces 16384-ray segments one after the other (raytracer.h:219-262). The device
    // wants far more rays in flight than that, and the result does not depend on how the rays are
    // batched (random streams are keyed by global ray index), so several segments are traced in ONE
    // device call and then handed to the processors segment by segment, with the per-segment
    // callback and keep_going test in the reference's order. The batch is bounded by the
    // reflection records that have to come back for host-side processors (<= 256 MB).
    const size_t per_segment_bytes = std::max<size_t>(1, keep) * sizeof(reflection) * segment_size;
    const size_t segments_per_batch = std::max<size_t>(1, std::min<size_t>(64, (size_t(256) << 20) / per_segment_bytes));
    std::vector<float> dirs;
    std::vector<reflection> refl;
    // traces rays [base, base + n) and feeds them to the processors in segments of segment_size;
    // returns false when keep_going was cleared after a full segment
    const size_t groups = total / segment_size;
    size_t group = 0;
    const auto run_batch = [&](It b, size_t n, size_t base) -> bool {
        dirs.resize(n * 3);
        auto it = b;
        for (size_t i = 0; i < n; ++i, ++it) {
            const auto d = *it;
            dirs[3 * i] = d.x; dirs[3 * i + 1] = d.y; dirs[3 * i + 2] = d.z;
        }
        refl.resize(keep * n);
        p.ray_index_base = base;
        wvb_reflection* const host_refl = keep ? reinterpret_cast<wvb_reflection*>(refl.data()) : nullptr;
        if (is.h) {
            core::detail::check(wvb_is_trace(is.h, &p, dirs.data(), n, uint32_t(is_order), host_refl, nullptr,
                                             nullptr));
        } else {
            core::detail::check(wvb_rt_trace(h.get(), &p, dirs.data(), n, host_refl, nullptr, nullptr));
        }
        for (size_t off = 0; off < n; off += segment_size) {
            const size_t len = std::min(segment_size, n - off);
            detail::for_each_in(processors, [&](auto& proc, auto) {
                auto grp = proc.get_group_processor(len);
                if (!detail::is_device_resident<std::decay_t<decltype(proc)>>::value) {
                    const size_t steps = std::min(keep, detail::steps_required_of(proc, depth, 0));
                    for (size_t s = 0; s < steps; ++s) {
                        grp.process(refl.begin() + s * n + off, refl.begin() + s * n + off + len, buffers, s, depth);
                    }
                }
                proc.accumulate(grp);
            });
            if (len == segment_size) {  // the tail segment has no callback (raytracer.h:246-262)
                per_step_callback(group, groups);
                ++group;
                if (!keep_going) return false;
            }
        }
        return true;
    };

    const size_t batch = segments_per_batch * segment_size;
    auto it = b_direction;
    for (size_t done = 0; done < total;) {
        const size_t n = std::min(batch, total - done);
        if (!run_batch(it, n, done)) return std::experimental::optional<return_type>{};
        std::advance(it, n);
        done += n;
    }

    if (hist.present) {
        std::vector<double> hbuf(size_t(p.n_bins) * 8 * (hist.directional ? 180 : 1));
        core::detail::check(wvb_rt_read_histogram(h.get(), hbuf.data()));
        detail::for_each_in(processors, [&](auto& proc, auto) { detail::store_histogram(proc, hbuf, p.n_bins); });
    }
    if (is.h) {
        uint64_t count = 0;
        core::detail::check(wvb_is_results(is.h, nullptr, 0, &count, nullptr, nullptr));
        util::aligned::vector<impulse<8>> imps(count);
        static_assert(sizeof(impulse<8>) == sizeof(wvb_impulse), "impulse<8> layout");
        core::detail::check(wvb_is_results(is.h, reinterpret_cast<wvb_impulse*>(imps.data()), count, &count,
                                           nullptr, nullptr));
        detail::for_each_in(processors, [&](auto& proc, auto) { detail::store_image_source(proc, imps); });
    }
    return std::experimental::make_optional(detail::collect_results(processors, seq));
}


// ---- the reference's own parameter types (raytracer.h:188-201) --------------------------------
/// `flatten` is the customisation point between a caller's scene type and the arrays the device
/// walks. This generic version works for anything with the interface of the reference's
/// `core::voxelised_scene_data<cl_float3, surface<8>>` (voxelised_scene_data.h:45-46):
///   get_voxels()      -> voxel_collection<3>: get_aabb().get_min()/get_max() (.x .y .z),
///                        get_side(), get_voxel(index3{x, y, z}) -> range of triangle indices
///   get_scene_data()  -> get_triangles(), get_vertices(), get_surfaces()
/// and restates get_flattened (voxel_collection.cpp:9-37) + make_scene_buffers
/// (scene_buffers.h:14-38). A caller whose type differs provides its own
/// `flatten(const T&) -> core::flattened_scene` next to T (found by ADL).
template <typename Voxelised>
auto flatten(const Voxelised& voxelised)
        -> decltype(voxelised.get_voxels().get_side(), voxelised.get_scene_data().get_triangles(),
                    core::flattened_scene{}) {
    core::flattened_scene out;
    const auto& voxels = voxelised.get_voxels();
    const auto side = size_t(voxels.get_side());
    const auto aabb = voxels.get_aabb();
    out.aabb_min = core::vec3{float(aabb.get_min().x), float(aabb.get_min().y), float(aabb.get_min().z)};
    out.aabb_max = core::vec3{float(aabb.get_max().x), float(aabb.get_max().y), float(aabb.get_max().z)};
    out.side = cl_uint(side);
    out.voxel_index.assign(side * side * side, 0);
    using index3 = std::decay_t<decltype(voxelised.voxel_index_type())>;
    for (size_t x = 0; x != side; ++x) {
        for (size_t y = 0; y != side; ++y) {
            for (size_t z = 0; z != side; ++z) {
                out.voxel_index[x * side * side + y * side + z] = cl_uint(out.voxel_index.size());
                const auto& v = voxels.get_voxel(index3(x, y, z));
                out.voxel_index.emplace_back(cl_uint(v.size()));
                for (const auto& i : v) out.voxel_index.emplace_back(cl_uint(i));
            }
        }
    }
    const auto& scene = voxelised.get_scene_data();
    for (const auto& t : scene.get_triangles()) out.triangles.push_back(core::triangle{t.surface, t.v0, t.v1, t.v2});
    for (const auto& v : scene.get_vertices()) out.vertices.push_back(cl_float3{{v.s[0], v.s[1], v.s[2], 0.0f}});
    for (const auto& sf : scene.get_surfaces()) {
        core::surface<core::simulation_bands> o{};
        for (int b = 0; b < core::simulation_bands; ++b) {
            o.absorption.s[b] = sf.absorption.s[b];
            o.scattering.s[b] = sf.scattering.s[b];
        }
        out.surfaces.push_back(o);
    }
    return out;
}

/// raytracer::run with the reference's parameter list: the scene is any type `flatten` accepts
/// (the reference's voxelised_scene_data among them), source and receiver any vector with
/// .x .y .z (glm::vec3). The processors' get_processor still receives the flattened scene.
template <typename It, typename Scene, typename Vec3, typename PerStepCallback, typename... Callbacks,
          typename = std::enable_if_t<!std::is_same<std::decay_t<Scene>, core::flattened_scene>::value>>
auto run(It b_direction, It e_direction, const core::compute_context& cc, const Scene& voxelised,
         const Vec3& source, const Vec3& receiver, const core::environment& environment,
         const std::atomic_bool& keep_going, PerStepCallback&& per_step_callback,
         std::tuple<Callbacks...> callbacks, uint64_t seed = std::random_device{}()) {
    return run(b_direction, e_direction, cc, flatten(voxelised), core::vec3{float(source.x), float(source.y), float(source.z)},
               core::vec3{float(receiver.x), float(receiver.y), float(receiver.z)}, environment, keep_going,
               std::forward<PerStepCallback>(per_step_callback), std::move(callbacks), seed);
}

/// image_source/run.h:12-47: raytracer::run with only the image-source processor, every
/// reflection step eligible (max order = the reflection depth).
namespace image_source {
template <typename It>
auto run(It b, It e, const core::compute_context& cc, const core::flattened_scene& voxelised,
         const core::vec3& source, const core::vec3& receiver, const core::environment& environment,
         uint64_t seed = std::random_device{}()) {
    const auto callbacks = std::make_tuple(reflection_processor::make_image_source{std::numeric_limits<size_t>::max()});
    auto results = raytracer::run(b, e, cc, voxelised, source, receiver, environment, true,
                                  [](auto /*i*/, auto /*steps*/) {}, callbacks, seed);
    if (!results) throw std::runtime_error{"Raytracer failed to generate results."};
    return std::move(std::get<0>(*results));
}
}  // namespace image_source

}  // namespace raytracer
}  // namespace wayverb
