// waveguide.hpp -- C++14 shim that keeps the reference's `waveguide::run` entry
// point and the types it takes, on top of the C ABI of libwvb200.so.
//
//   reference                                            here
//   src/waveguide/include/waveguide/waveguide.h:36-126    wayverb::waveguide::run
//   .../waveguide/mesh.h:12-26, setup.h:27-85             mesh, vectors
//   .../waveguide/mesh_descriptor.h:14-20                 mesh_descriptor (+ compute_index/locator)
//   .../waveguide/cl/structs.h, filter_structs.h          condensed_node, coefficients_canonical, ...
//   src/core/include/core/cl/common.h:13-57               core::compute_context, read_value, write_value,
//                                                         read_from_buffer, items_in_buffer
//   .../waveguide/preprocessor/{hard,soft}_source.h       preprocessor::make_hard_source / make_soft_source
//   .../waveguide/postprocessor/node.h                    postprocessor::node
//   src/core/include/core/exceptions.h:22-30              core::exceptions::value_is_inf / value_is_nan
//
// Everything is header-only; link with -lwvb200. The per-step order of effects
// is the reference's: pre -> flag reset -> kernel -> flag check/throw -> post ->
// swap. Pressures are fp64 on the device; `read_value<cl_float>` and friends
// convert, so callers that keep float (e.g. directional_receiver) are unchanged.
#pragma once

#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <iterator>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../wvb200.h"
#include "cl_compat.hpp"
#include "core.hpp"

namespace wayverb {
namespace waveguide {

// ---- PODs: the layouts are the contract (see include/wvb200.h for file:line) ----
using error_code = cl_int;
constexpr cl_int id_success = 0, id_inf_error = 1 << 0, id_nan_error = 1 << 1,
                 id_outside_range_error = 1 << 2, id_outside_mesh_error = 1 << 3,
                 id_suspicious_boundary_error = 1 << 4;
using boundary_type = cl_int;
constexpr cl_int id_none = 0, id_inside = 1 << 0, id_nx = 1 << 1, id_px = 1 << 2, id_ny = 1 << 3,
                 id_py = 1 << 4, id_nz = 1 << 5, id_pz = 1 << 6, id_reentrant = 1 << 7;
constexpr auto no_neighbor = ~cl_uint{0};
constexpr size_t biquad_order{2}, biquad_sections{3};
using filt_real = cl_double;

struct alignas(8) condensed_node final {
    cl_int boundary_type{};
    cl_uint boundary_index{};
};
template <size_t o>
struct alignas(8) coefficients final {
    static constexpr auto order = o;
    filt_real b[order + 1]{};
    filt_real a[order + 1]{};
};
using coefficients_canonical = coefficients<biquad_order * biquad_sections>;
template <size_t n>
struct alignas(4) boundary_index_array final {
    cl_uint array[n];
};
using boundary_index_array_1 = boundary_index_array<1>;
using boundary_index_array_2 = boundary_index_array<2>;
using boundary_index_array_3 = boundary_index_array<3>;
struct boundary_index_data final {
    util::aligned::vector<boundary_index_array_1> b1;
    util::aligned::vector<boundary_index_array_2> b2;
    util::aligned::vector<boundary_index_array_3> b3;
};
static_assert(sizeof(condensed_node) == sizeof(wvb_condensed_node), "condensed_node layout");
static_assert(sizeof(coefficients_canonical) == sizeof(wvb_coefficients_canonical),
              "coefficients_canonical layout");

// ---- boundary filter design: fitted_boundary.h, arbitrary_magnitude_filter.h, stable.h --
/// frequency_domain_envelope.h:9-31 (points kept in ascending frequency order)
class frequency_domain_envelope final {
public:
    struct point final {
        double frequency;
        double amplitude;
    };
    using const_iterator = std::vector<point>::const_iterator;
    const_iterator cbegin() const { return points.cbegin(); }
    const_iterator cend() const { return points.cend(); }
    void insert(point p) {
        points.insert(std::lower_bound(points.begin(), points.end(), p,
                                       [](const point& a, const point& b) { return a.frequency < b.frequency; }),
                      p);
    }

private:
    std::vector<point> points;
};
template <size_t N>
frequency_domain_envelope make_frequency_domain_envelope(const std::array<double, N>& frequency,
                                                         const std::array<double, N>& amplitude) {
    frequency_domain_envelope ret;
    for (size_t i = 0; i != N; ++i) ret.insert({frequency[i], amplitude[i]});
    return ret;
}
/// arbitrary_magnitude_filter.h:63-95 (order 6 is the only one wayverb instantiates);
/// the Yule-Walker fit runs in libwvb200.so instead of IT++
template <size_t N>
coefficients<N> arbitrary_magnitude_filter(const frequency_domain_envelope& env) {
    static_assert(N == coefficients_canonical::order, "order 6 only");
    std::vector<double> f, a;
    for (auto it = env.cbegin(); it != env.cend(); ++it) {
        f.push_back(it->frequency);
        a.push_back(it->amplitude);
    }
    coefficients<N> ret;
    core::detail::check(wvb_lrs_arbitrary_magnitude_filter(f.data(), a.data(), uint32_t(f.size()),
                                                          reinterpret_cast<wvb_coefficients_canonical*>(&ret)));
    return ret;
}
/// stable.h:37-50
template <typename T>
bool is_stable(const T& a) {
    const std::vector<double> v(std::begin(a), std::end(a));
    return wvb_lrs_is_stable(v.data(), uint32_t(v.size())) != 0;
}
/// fitted_boundary.h:34-50
inline coefficients_canonical to_impedance_coefficients(const coefficients_canonical& c) {
    coefficients_canonical ret;
    wvb_lrs_to_impedance(reinterpret_cast<const wvb_coefficients_canonical*>(&c),
                         reinterpret_cast<wvb_coefficients_canonical*>(&ret));
    return ret;
}
/// fitted_boundary.h:72-75
inline coefficients_canonical to_flat_coefficients(double absorption) {
    coefficients_canonical ret;
    wvb_lrs_flat(absorption, reinterpret_cast<wvb_coefficients_canonical*>(&ret));
    return ret;
}
/// fitted_boundary.h:79-104; throws std::runtime_error("Unable to generate stable boundary
/// filter.") like the reference
template <typename T>
coefficients_canonical compute_reflectance_filter_coefficients(const T& absorption, double sample_rate) {
    double a[8];
    size_t i = 0;
    for (auto it = std::begin(absorption); it != std::end(absorption) && i < 8; ++it, ++i) a[i] = *it;
    for (; i < 8; ++i) a[i] = 0;
    coefficients_canonical ret;
    const auto s = wvb_lrs_reflectance_filter(a, sample_rate, reinterpret_cast<wvb_coefficients_canonical*>(&ret));
    if (s != WVB_OK) throw std::runtime_error{wvb_last_error()};
    return ret;
}

struct alignas(16) mesh_descriptor final {
    static constexpr auto no_neighbor = ~cl_uint{0};
    cl_float3 min_corner;
    cl_int3 dimensions;
    cl_float spacing;
};
static_assert(sizeof(mesh_descriptor) == 48, "mesh_descriptor layout (SURVEY 8a)");

inline size_t compute_num_nodes(const mesh_descriptor& d) {
    return size_t(d.dimensions.s[0]) * size_t(d.dimensions.s[1]) * size_t(d.dimensions.s[2]);
}
/// compute_index (mesh_descriptor.cpp:7-10): x fastest
inline size_t compute_index(const mesh_descriptor& d, int x, int y, int z) {
    return size_t(x) + size_t(y) * d.dimensions.s[0] +
           size_t(z) * d.dimensions.s[0] * size_t(d.dimensions.s[1]);
}
/// compute_locator (mesh_descriptor.cpp:16-20)
inline std::array<int, 3> compute_locator(const mesh_descriptor& d, size_t index) {
    const size_t dx = d.dimensions.s[0], dy = d.dimensions.s[1], dz = d.dimensions.s[2];
    return {{int(index % dx), int((index / dx) % dy), int((index / dx / dy) % dz)}};
}

constexpr bool is_inside(const condensed_node& c) { return c.boundary_type & id_inside; }

class vectors final {
public:
    vectors(util::aligned::vector<condensed_node> nodes,
            util::aligned::vector<coefficients_canonical> coefficients,
            boundary_index_data boundary_index_data)
            : condensed_nodes_(std::move(nodes))
            , coefficients_(std::move(coefficients))
            , boundary_index_data_(std::move(boundary_index_data)) {}

    template <size_t n>
    const util::aligned::vector<boundary_index_array<n>>& get_boundary_indices() const;
    const util::aligned::vector<condensed_node>& get_condensed_nodes() const {
        return condensed_nodes_;
    }
    const util::aligned::vector<coefficients_canonical>& get_coefficients() const {
        return coefficients_;
    }
    void set_coefficients(coefficients_canonical c) {
        for (auto& i : coefficients_) i = c;
    }
    void set_coefficients(util::aligned::vector<coefficients_canonical> c) {
        if (c.size() != coefficients_.size()) {
            throw std::runtime_error(
                    "Size of new coefficients vector must be equal to the existing one in order "
                    "to maintain object invariants.");
        }
        coefficients_ = std::move(c);
    }

private:
    util::aligned::vector<condensed_node> condensed_nodes_;
    util::aligned::vector<coefficients_canonical> coefficients_;
    boundary_index_data boundary_index_data_;
};
template <>
inline const util::aligned::vector<boundary_index_array<1>>& vectors::get_boundary_indices<1>() const {
    return boundary_index_data_.b1;
}
template <>
inline const util::aligned::vector<boundary_index_array<2>>& vectors::get_boundary_indices<2>() const {
    return boundary_index_data_.b2;
}
template <>
inline const util::aligned::vector<boundary_index_array<3>>& vectors::get_boundary_indices<3>() const {
    return boundary_index_data_.b3;
}

class mesh final {
public:
    mesh(mesh_descriptor descriptor, vectors vectors)
            : descriptor_(descriptor), vectors_(std::move(vectors)) {}
    const mesh_descriptor& get_descriptor() const { return descriptor_; }
    const vectors& get_structure() const { return vectors_; }
    void set_coefficients(coefficients_canonical c) { vectors_.set_coefficients(c); }
    void set_coefficients(util::aligned::vector<coefficients_canonical> c) {
        vectors_.set_coefficients(std::move(c));
    }

private:
    mesh_descriptor descriptor_;
    vectors vectors_;
};

inline bool is_inside(const mesh& m, size_t node_index) {
    return is_inside(m.get_structure().get_condensed_nodes()[node_index]);
}

/// Synthetic box room through wvb_mesh_cuboid (what compute_mesh, mesh.cpp:53-141,
/// yields for a cuboid), one surface.
inline mesh make_cuboid_mesh(int dx, int dy, int dz, float spacing, coefficients_canonical c) {
    mesh_descriptor d{};
    d.dimensions.s[0] = dx; d.dimensions.s[1] = dy; d.dimensions.s[2] = dz;
    d.spacing = spacing;
    util::aligned::vector<condensed_node> nodes(size_t(dx) * dy * dz);
    const int32_t dim[3] = {dx, dy, dz};
    uint64_t counts[3] = {0, 0, 0};
    core::detail::check(wvb_mesh_cuboid(dim, 0, dz, reinterpret_cast<wvb_condensed_node*>(nodes.data()),
                                        counts));
    boundary_index_data b;
    b.b1.assign(counts[0], boundary_index_array_1{{0}});
    b.b2.assign(counts[1], boundary_index_array_2{{0, 0}});
    b.b3.assign(counts[2], boundary_index_array_3{{0, 0, 0}});
    return mesh{d, vectors{std::move(nodes), {c}, std::move(b)}};
}

// ---- the entry point -------------------------------------------------------------------
namespace detail {
class handle final {
public:
    handle(const core::compute_context& cc, const mesh& m) {
        const auto& v = m.get_structure();
        wvb_wg_desc d{};
        for (int i = 0; i < 3; ++i) d.dim[i] = m.get_descriptor().dimensions.s[i];
        d.z_begin = 0;
        d.z_end = d.dim[2];
        d.nodes = reinterpret_cast<const wvb_condensed_node*>(v.get_condensed_nodes().data());
        d.nodes_z0 = 0;
        d.nodes_nz = d.dim[2];
        d.coefficients =
                reinterpret_cast<const wvb_coefficients_canonical*>(v.get_coefficients().data());
        d.num_coefficients = uint32_t(v.get_coefficients().size());
        d.boundary_index[0] = reinterpret_cast<const uint32_t*>(v.get_boundary_indices<1>().data());
        d.boundary_index[1] = reinterpret_cast<const uint32_t*>(v.get_boundary_indices<2>().data());
        d.boundary_index[2] = reinterpret_cast<const uint32_t*>(v.get_boundary_indices<3>().data());
        d.boundary_count[0] = v.get_boundary_indices<1>().size();
        d.boundary_count[1] = v.get_boundary_indices<2>().size();
        d.boundary_count[2] = v.get_boundary_indices<3>().size();
        d.device = cc.device;
        d.rank = 0;
        d.nranks = 1;
        core::detail::check(wvb_wg_create(&d, &wg_));
        if (!wg_) throw std::runtime_error{std::string{"libwvb200: "} + wvb_last_error()};
    }
    ~handle() { wvb_wg_destroy(wg_); }
    handle(const handle&) = delete;
    handle& operator=(const handle&) = delete;
    wvb_wg* get() const { return wg_; }

private:
    wvb_wg* wg_{nullptr};
};
}  // namespace detail

/// waveguide::run (waveguide.h:36-126).
///   pre(cl::CommandQueue&, cl::Buffer&, size_t step) -> bool   (false ends the run)
///   post(cl::CommandQueue&, const cl::Buffer&, size_t step)
/// returns the number of steps completed; throws what the reference throws.
template <typename step_preprocessor, typename step_postprocessor>
size_t run(const core::compute_context& cc,
           const mesh& mesh,
           step_preprocessor&& pre,
           step_postprocessor&& post,
           const std::atomic_bool& keep_going) {
    const detail::handle h{cc, mesh};
    cl::CommandQueue queue{h.get()};
    cl::Buffer current{h.get(), compute_num_nodes(mesh.get_descriptor())};

    auto step = 0u;
    for (; pre(queue, current, step) && keep_going; ++step) {
        int32_t error_flag = 0;
        core::detail::check(wvb_wg_launch(h.get(), &error_flag));
        if (error_flag) {
            if (error_flag & id_inf_error) {
                throw core::exceptions::value_is_inf(
                        "Pressure value is inf, check filter coefficients.");
            }
            if (error_flag & id_nan_error) {
                throw core::exceptions::value_is_nan(
                        "Pressure value is nan, check filter coefficients.");
            }
            if (error_flag & id_outside_mesh_error) {
                throw std::runtime_error("Tried to read non-existant node.");
            }
            if (error_flag & id_suspicious_boundary_error) {
                throw std::runtime_error("Suspicious boundary read.");
            }
        }
        post(queue, static_cast<const cl::Buffer&>(current), step);
        core::detail::check(wvb_wg_swap(h.get()));
    }
    return step;
}

/// The same loop for the stock processors, executed on the device in one call
/// (wvb_wg_run): hard/soft source at one node, postprocessor::node at each
/// receiver. out[step * receivers.size() + r].
namespace detail {
inline int32_t poll_keep_going(void* user) { return static_cast<const std::atomic_bool*>(user)->load() ? 1 : 0; }
}  // namespace detail
inline size_t run_stock(const core::compute_context& cc, const mesh& mesh, size_t source_node,
                        const util::aligned::vector<double>& signal, bool soft,
                        const util::aligned::vector<size_t>& receivers,
                        util::aligned::vector<double>& out, const std::atomic_bool* keep_going = nullptr) {
    const detail::handle h{cc, mesh};
    util::aligned::vector<uint64_t> rcv(receivers.begin(), receivers.end());
    out.assign(signal.size() * receivers.size(), 0.0);
    wvb_wg_run_params p{};
    p.source_node = source_node;
    p.signal = signal.data();
    p.n_steps = uint32_t(signal.size());
    p.soft = soft ? 1 : 0;
    p.receiver_nodes = rcv.data();
    p.n_receivers = uint32_t(rcv.size());
    p.out = out.data();
    p.check_interval = 64;
    // polled every 64 steps: the reference tests keep_going every step (waveguide.h:80) and
    // returns the steps completed so far
    p.keep_going = keep_going ? &detail::poll_keep_going : nullptr;
    p.keep_going_user = const_cast<std::atomic_bool*>(keep_going);
    uint32_t done = 0;
    int32_t flags = 0;
    core::detail::check(wvb_wg_run(h.get(), &p, &done, &flags));
    if (flags & id_inf_error) throw core::exceptions::value_is_inf("Pressure value is inf, check filter coefficients.");
    if (flags & id_nan_error) throw core::exceptions::value_is_nan("Pressure value is nan, check filter coefficients.");
    if (flags & id_outside_mesh_error) throw std::runtime_error("Tried to read non-existant node.");
    if (flags & id_suspicious_boundary_error) throw std::runtime_error("Suspicious boundary read.");
    return done;
}

// ---- stock processors (same behaviour as the reference's) ---------------------------------
#ifdef WVB_WITH_REFERENCE_HEADERS
}  // namespace waveguide
}  // namespace wayverb
// Overlay mode: the stock processors are the reference's OWN headers, compiled unmodified
// against include/compat (core/cl/common.h -> core.hpp). Needs the reference's
// src/waveguide/include on the include path, after include/compat.
#include "waveguide/postprocessor/node.h"
#include "waveguide/preprocessor/hard_source.h"
#include "waveguide/preprocessor/soft_source.h"
namespace wayverb {
namespace waveguide {
namespace preprocessor {
#else
namespace preprocessor {
/// hard_source.h:9-37: overwrite the node with the next sample every step.
template <typename It>
class hard_source final {
public:
    hard_source(size_t node, It begin, It end) : node_{node}, begin_{begin}, end_{end} {}
    bool operator()(cl::CommandQueue& queue, cl::Buffer& buffer, size_t) {
        if (begin_ == end_) return false;
        core::write_value(queue, buffer, node_, *begin_++);
        return true;
    }

private:
    size_t node_;
    It begin_, end_;
};
template <typename It>
auto make_hard_source(size_t node, It begin, It end) {
    return hard_source<It>{node, begin, end};
}
/// soft_source.h:9-39: add the next sample to the node every step.
template <typename It>
class soft_source final {
public:
    soft_source(size_t node, It begin, It end) : node_{node}, begin_{begin}, end_{end} {}
    bool operator()(cl::CommandQueue& queue, cl::Buffer& buffer, size_t) {
        if (begin_ == end_) return false;
        const auto current_pressure = core::read_value<cl_double>(queue, buffer, node_);
        core::write_value(queue, buffer, node_, current_pressure + *begin_++);
        return true;
    }

private:
    size_t node_;
    It begin_, end_;
};
template <typename It>
auto make_soft_source(size_t node, It begin, It end) {
    return soft_source<It>{node, begin, end};
}
#endif  // WVB_WITH_REFERENCE_HEADERS
/// gaussian.cpp:10-53: at step 0 the whole field is set to a 3-d gaussian centred
/// on `centre_pos` (evaluated on the host in the reference's mixed float/double
/// arithmetic, written as float values); the run lasts `steps` steps.
class gaussian final {
public:
    static float compute(float dx, float dy, float dz, float sdev) {
        const float len = std::sqrt(dx * dx + dy * dy + dz * dz);
        return float(std::exp(-std::pow(double(len), 2) / (2 * std::pow(double(sdev), 2))) /
                     std::pow(sdev * std::sqrt(2 * M_PI), 3));
    }
    gaussian(const mesh_descriptor& descriptor, float cx, float cy, float cz, float sdev, size_t steps)
            : descriptor_(descriptor), cx_{cx}, cy_{cy}, cz_{cz}, sdev_{sdev}, steps_{steps} {}
    bool operator()(cl::CommandQueue&, cl::Buffer& buffer, size_t step) const {
        if (step == steps_) return false;
        if (step == 0) {
            const auto nodes = core::items_in_buffer<cl_double>(buffer);
            util::aligned::vector<double> pressures;
            pressures.reserve(nodes);
            for (size_t i = 0; i != nodes; ++i) {
                const auto l = compute_locator(descriptor_, i);
                // compute_position (mesh_descriptor.cpp:27-34): min_corner + locator * spacing
                const float px = descriptor_.min_corner.s[0] + float(l[0]) * descriptor_.spacing;
                const float py = descriptor_.min_corner.s[1] + float(l[1]) * descriptor_.spacing;
                const float pz = descriptor_.min_corner.s[2] + float(l[2]) * descriptor_.spacing;
                pressures.emplace_back(compute(px - cx_, py - cy_, pz - cz_, sdev_));
            }
            core::detail::check(wvb_wg_write_field(buffer.handle(), pressures.data()));
        }
        return true;
    }

private:
    mesh_descriptor descriptor_;
    float cx_, cy_, cz_, sdev_;
    size_t steps_;
};

}  // namespace preprocessor

/// compute_neighbors (mesh_descriptor.cpp:36-57): nx, px, ny, py, nz, pz or no_neighbor
inline std::array<cl_uint, 6> compute_neighbors(const mesh_descriptor& d, size_t index) {
    const auto l = compute_locator(d, index);
    const int dl[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
    std::array<cl_uint, 6> ret;
    for (int p = 0; p < 6; ++p) {
        const int x = l[0] + dl[p][0], y = l[1] + dl[p][1], z = l[2] + dl[p][2];
        const bool inside = 0 <= x && 0 <= y && 0 <= z && x < d.dimensions.s[0] &&
                            y < d.dimensions.s[1] && z < d.dimensions.s[2];
        ret[p] = inside ? cl_uint(compute_index(d, x, y, z)) : mesh_descriptor::no_neighbor;
    }
    return ret;
}

namespace postprocessor {
#ifndef WVB_WITH_REFERENCE_HEADERS
/// node.cpp:14-18: the pressure at one node each step.
class node final {
public:
    explicit node(size_t output_node) : output_node_{output_node} {}
    using return_type = double;
    return_type operator()(cl::CommandQueue& queue, const cl::Buffer& buffer, size_t) const {
        return core::read_value<cl_double>(queue, buffer, output_node_);
    }
    size_t get_output_node() const { return output_node_; }

private:
    size_t output_node_;
};
#endif  // WVB_WITH_REFERENCE_HEADERS (else: the reference's postprocessor/node.h + node.cpp)
/// directional_receiver.cpp:10-69: pressure + intensity (velocity from the pressure
/// gradient of the six neighbours, integrated in double on the host). Reads are
/// `cl_float` like the reference's, i.e. the fp64 device values are converted.
class directional_receiver final {
public:
    directional_receiver(const mesh_descriptor& mesh_descriptor, double sample_rate,
                         double ambient_density, size_t output_node)
            : mesh_spacing_{mesh_descriptor.spacing}
            , sample_rate_{sample_rate}
            , ambient_density_{ambient_density}
            , output_node_{output_node}
            , surrounding_nodes_(compute_neighbors(mesh_descriptor, output_node)) {
        for (const auto& i : surrounding_nodes_) {
            if (i == ~cl_uint{0}) {
                throw std::runtime_error(
                        "Can't place directional_receiver at this node as it is adjacent to a "
                        "boundary.");
            }
        }
    }
    struct output final {
        float intensity[3];
        float pressure;
    };
    using return_type = output;
    return_type operator()(cl::CommandQueue& queue, const cl::Buffer& buffer, size_t) {
        const auto pressure = core::read_value<cl_float>(queue, buffer, output_node_);
        std::array<cl_float, 6> surrounding;
        for (size_t i = 0; i != 6; ++i) {
            surrounding[i] = cl_float(
                    (core::read_value<cl_float>(queue, buffer, surrounding_nodes_[i]) - pressure) /
                    mesh_spacing_);
        }
        const double m[3] = {(surrounding[1] - surrounding[0]) * 0.5,
                             (surrounding[3] - surrounding[2]) * 0.5,
                             (surrounding[5] - surrounding[4]) * 0.5};
        output ret{};
        for (int k = 0; k < 3; ++k) {
            velocity_[k] -= m[k] / (ambient_density_ * sample_rate_);
            ret.intensity[k] = float(velocity_[k] * double(pressure));
        }
        ret.pressure = pressure;
        return ret;
    }
    /// the same evaluation from values gathered on the device: v[0] the node, v[1..6] its
    /// neighbours in port order (fp64 device values, converted to cl_float like read_value does)
    return_type from_gathered(const double* v) {
        const auto pressure = cl_float(v[0]);
        std::array<cl_float, 6> surrounding;
        for (size_t i = 0; i != 6; ++i) surrounding[i] = cl_float((cl_float(v[1 + i]) - pressure) / mesh_spacing_);
        const double m[3] = {(surrounding[1] - surrounding[0]) * 0.5, (surrounding[3] - surrounding[2]) * 0.5,
                             (surrounding[5] - surrounding[4]) * 0.5};
        output ret{};
        for (int k = 0; k < 3; ++k) {
            velocity_[k] -= m[k] / (ambient_density_ * sample_rate_);
            ret.intensity[k] = float(velocity_[k] * double(pressure));
        }
        ret.pressure = pressure;
        return ret;
    }
    size_t get_output_node() const { return output_node_; }

private:
    double mesh_spacing_;
    double sample_rate_;
    double ambient_density_;
    size_t output_node_;
    std::array<cl_uint, 6> surrounding_nodes_;
    double velocity_[3]{0, 0, 0};
};

}  // namespace postprocessor

// ---- canonical.h: the combination the engine drives -------------------------------------
/// config.cpp:19-21 / mesh_descriptor.cpp:72-74
inline double compute_sample_rate(const mesh_descriptor& d, double speed_of_sound) {
    return 1 / (double(d.spacing) / (speed_of_sound * std::sqrt(3.0)));
}
/// mesh_descriptor.cpp:12-25: nearest node of a position (glm::round of float quotients)
inline size_t compute_index(const mesh_descriptor& d, const core::vec3& pos) {
    const float q[3] = {(pos.x - d.min_corner.s[0]) / d.spacing, (pos.y - d.min_corner.s[1]) / d.spacing,
                        (pos.z - d.min_corner.s[2]) / d.spacing};
    return compute_index(d, int(std::round(q[0])), int(std::round(q[1])), int(std::round(q[2])));
}
/// mesh_descriptor.cpp:27-30
inline core::vec3 compute_position(const mesh_descriptor& d, const std::array<int, 3>& locator) {
    return {d.min_corner.s[0] + float(locator[0]) * d.spacing, d.min_corner.s[1] + float(locator[1]) * d.spacing,
            d.min_corner.s[2] + float(locator[2]) * d.spacing};
}
/// calibration.h:21-31
inline double rectilinear_calibration_factor(double grid_spacing, double acoustic_impedance) {
    return std::sqrt(acoustic_impedance / (4 * M_PI)) / (0.3405 * grid_spacing);
}
/// bandpass_band.h:11-20
struct band final {
    util::aligned::vector<postprocessor::directional_receiver::output> directional;
    double sample_rate;
};
struct bandpass_band final {
    struct band band;
    struct {
        double min, max;
    } valid_hz;
};
/// Tag for "no per-step pressure callback": lets canonical_impl run the whole loop on the
/// device (wvb_wg_run) and evaluate the directional receiver afterwards from the gathered
/// pressures -- the same values, in the same order, as the per-step path.
struct no_pressure_callback final {
    template <typename Q, typename B>
    void operator()(Q&, const B&, size_t, size_t) const {}
};

namespace detail {
inline size_t checked_node(const mesh& mesh, const core::vec3& pt) {
    const auto ret = compute_index(mesh.get_descriptor(), pt);
    if (ret >= compute_num_nodes(mesh.get_descriptor()) || !is_inside(mesh, ret)) {
        throw std::runtime_error{"Source/receiver node position appears to be outside mesh."};
    }
    return ret;
}
inline util::aligned::vector<float> canonical_input(const mesh& mesh, double ideal_steps,
                                                    const core::environment& environment) {
    auto raw = util::aligned::vector<float>(size_t(ideal_steps), 0.0f);
    if (!raw.empty()) {
        raw.front() = float(rectilinear_calibration_factor(mesh.get_descriptor().spacing,
                                                           environment.acoustic_impedance));
    }
    return raw;
}

/// canonical.h:24-81
template <typename Callback>
bool canonical_impl(const core::compute_context& cc, const mesh& mesh, double simulation_time,
                    const core::vec3& source, const core::vec3& receiver, const core::environment& environment,
                    const std::atomic_bool& keep_going, Callback&& callback, band& out) {
    const auto sample_rate = compute_sample_rate(mesh.get_descriptor(), environment.speed_of_sound);
    const auto ideal_steps = std::ceil(sample_rate * simulation_time);
    const auto input = canonical_input(mesh, ideal_steps, environment);
    core::callback_accumulator<postprocessor::directional_receiver> output_accumulator{
            mesh.get_descriptor(), sample_rate, core::get_ambient_density(environment),
            checked_node(mesh, receiver)};
    const auto steps = run(cc, mesh,
                           preprocessor::make_hard_source(checked_node(mesh, source), input.begin(), input.end()),
                           [&](auto& queue, const auto& buffer, auto step) {
                               output_accumulator(queue, buffer, step);
                               callback(queue, buffer, step, size_t(ideal_steps));
                           },
                           keep_going);
    if (double(steps) != ideal_steps) return false;
    out = band{std::move(output_accumulator.get_output()), sample_rate};
    return true;
}

/// The same with nothing to call per step: one wvb_wg_run, receiver evaluated afterwards.
inline bool canonical_impl(const core::compute_context& cc, const mesh& mesh, double simulation_time,
                           const core::vec3& source, const core::vec3& receiver,
                           const core::environment& environment, const std::atomic_bool& keep_going,
                           no_pressure_callback, band& out) {
    if (!keep_going) return false;
    const auto sample_rate = compute_sample_rate(mesh.get_descriptor(), environment.speed_of_sound);
    const auto ideal_steps = std::ceil(sample_rate * simulation_time);
    const auto input = canonical_input(mesh, ideal_steps, environment);
    const size_t rcv = checked_node(mesh, receiver);
    postprocessor::directional_receiver dr{mesh.get_descriptor(), sample_rate,
                                           core::get_ambient_density(environment), rcv};
    util::aligned::vector<size_t> nodes{rcv};
    for (auto n : compute_neighbors(mesh.get_descriptor(), rcv)) nodes.push_back(n);
    const util::aligned::vector<double> signal(input.begin(), input.end());
    util::aligned::vector<double> gathered;
    const auto steps = run_stock(cc, mesh, checked_node(mesh, source), signal, false, nodes, gathered, &keep_going);
    if (double(steps) != ideal_steps) return false;
    out.sample_rate = sample_rate;
    out.directional.clear();
    out.directional.reserve(steps);
    for (size_t s = 0; s < steps; ++s) out.directional.push_back(dr.from_gathered(&gathered[s * 7]));
    return true;
}
}  // namespace detail

/// canonical.h:85-110 (single band): hard source at the node nearest `source`, directional
/// receiver at the node nearest `receiver`, `simulation_time` seconds. Empty on cancellation.
template <typename PressureCallback>
util::aligned::vector<bandpass_band> canonical(const core::compute_context& cc, const mesh& mesh,
                                               const core::vec3& source, const core::vec3& receiver,
                                               const core::environment& environment, double cutoff,
                                               double simulation_time, const std::atomic_bool& keep_going,
                                               PressureCallback&& pressure_callback) {
    band b;
    if (!detail::canonical_impl(cc, mesh, simulation_time, source, receiver, environment, keep_going,
                                std::forward<PressureCallback>(pressure_callback), b)) {
        return {};
    }
    util::aligned::vector<bandpass_band> ret(1);
    ret[0].band = std::move(b);
    ret[0].valid_hz.min = 0.0;
    ret[0].valid_hz.max = cutoff;
    return ret;
}

/// simulation_parameters.h:9-44
struct single_band_parameters final {
    double cutoff;
    double usable_portion;
};
struct multiple_band_constant_spacing_parameters final {
    size_t bands;
    double cutoff;
    double usable_portion;
};
/// simulation_parameters.h:58-72
constexpr double compute_cutoff_frequency(double sample_rate, double usable_portion) {
    return sample_rate * 0.25 * usable_portion;
}
constexpr double compute_sampling_frequency(double cutoff, double usable_portion) {
    return cutoff / (0.25 * usable_portion);
}
/// hrtf_band_params_hz().edges (hrtf/multiband.h:25-28, frequency_domain/envelope.cpp:46-49):
/// 8 bands over 20 Hz - 20 kHz
inline double hrtf_band_edge_hz(size_t edge) { return 20.0 * std::pow(20000.0 / 20.0, double(edge) / 8.0); }

/// canonical.h:140-177, the multi-band variant: the waveguide is run once per band with every
/// surface's flat coefficients for that band (set_flat_coefficients_for_band, :128-137), each
/// band valid between the band's edges. `surfaces`: one entry per coefficient set of the mesh
/// (the scene's surfaces, `.absorption.s[band]`). Empty on cancellation.
template <typename Surfaces, typename PressureCallback>
util::aligned::vector<bandpass_band> canonical(const core::compute_context& cc, mesh m, const Surfaces& surfaces,
                                               const core::vec3& source, const core::vec3& receiver,
                                               const core::environment& environment,
                                               const multiple_band_constant_spacing_parameters& sim_params,
                                               double simulation_time, const std::atomic_bool& keep_going,
                                               PressureCallback&& pressure_callback) {
    util::aligned::vector<bandpass_band> ret;
    for (size_t b = 0; b != sim_params.bands; ++b) {
        util::aligned::vector<coefficients_canonical> coeffs;
        for (const auto& surface : surfaces) coeffs.push_back(to_flat_coefficients(surface.absorption.s[b]));
        m.set_coefficients(std::move(coeffs));
        band rendered;
        if (!detail::canonical_impl(cc, m, simulation_time, source, receiver, environment, keep_going,
                                    pressure_callback, rendered)) {
            return {};
        }
        bandpass_band bb;
        bb.band = std::move(rendered);
        bb.valid_hz.min = hrtf_band_edge_hz(b);
        bb.valid_hz.max = hrtf_band_edge_hz(b + 1);
        ret.push_back(std::move(bb));
    }
    return ret;
}

}  // namespace waveguide
}  // namespace wayverb
