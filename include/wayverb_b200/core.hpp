// core.hpp -- the part of the reference's `core` library the hot-path entry points and their
// stock processors are written against, re-pointed at libwvb200.so:
//
//   reference                                              here
//   src/core/include/core/cl/common.h:10-57                 core::device_type, compute_context,
//                                                           items_in_buffer, read_from_buffer,
//                                                           read_value, write_value
//   src/core/include/core/exceptions.h:22-30                core::exceptions::value_is_inf / value_is_nan
//   src/core/include/core/environment.h:6-13                core::environment, get_ambient_density
//   src/core/include/core/callback_accumulator.h            core::callback_accumulator
//   src/utilities/include/utilities/aligned/vector.h        util::aligned::vector
//
// include/compat/ holds forwarding headers under the reference's own paths
// (core/cl/common.h, core/cl/include.h, core/exceptions.h, core/environment.h,
// core/callback_accumulator.h, utilities/aligned/vector.h) so that UNMODIFIED reference
// headers -- waveguide/preprocessor/hard_source.h, soft_source.h, postprocessor/node.h and
// node.cpp -- compile against this file with `-I include/compat`
// (tests/cpp/test_overlay.cpp does exactly that).
#pragma once

#include <algorithm>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../wvb200.h"
#include "cl_compat.hpp"

namespace util {
namespace aligned {
template <typename T>
using vector = std::vector<T>;  // utilities/aligned/vector.h: std::vector with an aligned allocator
}  // namespace aligned
}  // namespace util

namespace wayverb {
namespace core {

enum class device_type { cpu, gpu };

/// cl::Context + cl::Device in the reference (cl/common.h:13-22); here a CUDA
/// device ordinal. There is no CPU device: device_type::cpu is refused.
class compute_context final {
public:
    compute_context() = default;
    explicit compute_context(int cuda_device) : device{cuda_device} {}
    explicit compute_context(device_type type) {
        if (type != device_type::gpu) {
            throw std::runtime_error{"wayverb_b200 has no CPU path: a B200 is required."};
        }
    }
    int device{0};
};

namespace exceptions {
struct value_is_nan final : public std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct value_is_inf final : public std::runtime_error {
    using std::runtime_error::runtime_error;
};
}  // namespace exceptions

/// core/environment.h:6-13
struct environment final {
    double speed_of_sound{340.0};
    double acoustic_impedance{400.0};
};
constexpr double get_ambient_density(const environment& s) { return s.acoustic_impedance / s.speed_of_sound; }
/// stands in for glm::vec3 in signatures
struct vec3 final {
    float x, y, z;
};

namespace detail {
inline void check(wvb_status s) {
    if (s != WVB_OK && s != WVB_ERR_SIM) {
        throw std::runtime_error{std::string{"libwvb200: "} + wvb_last_error()};
    }
}
}  // namespace detail

template <typename T>
size_t items_in_buffer(const cl::Buffer& buffer) {
    return buffer.items();
}

template <typename T>
T read_value(cl::CommandQueue&, const cl::Buffer& buffer, size_t index) {
    double v = 0;
    detail::check(wvb_wg_read_f64(buffer.handle(), index, &v, nullptr));
    return static_cast<T>(v);
}

template <typename T>
void write_value(cl::CommandQueue&, cl::Buffer& buffer, size_t index, T val) {
    detail::check(wvb_wg_write_f64(buffer.handle(), index, static_cast<double>(val)));
}

template <typename T>
util::aligned::vector<T> read_from_buffer(cl::CommandQueue&, const cl::Buffer& buffer);

template <>
inline util::aligned::vector<double> read_from_buffer<double>(cl::CommandQueue&,
                                                              const cl::Buffer& buffer) {
    util::aligned::vector<double> ret(buffer.items());
    detail::check(wvb_wg_read_field(buffer.handle(), ret.data()));
    return ret;
}
template <>
inline util::aligned::vector<float> read_from_buffer<float>(cl::CommandQueue&,
                                                            const cl::Buffer& buffer) {
    util::aligned::vector<float> ret(buffer.items());
    detail::check(wvb_wg_read_field_f32(buffer.handle(), ret.data()));
    return ret;
}

/// callback_accumulator (core/callback_accumulator.h): collects what a
/// postprocessor returns each step.
template <typename Callback>
class callback_accumulator final {
public:
    template <typename... Ts>
    explicit callback_accumulator(Ts&&... ts) : callback_{std::forward<Ts>(ts)...} {}
    template <typename... Ts>
    void operator()(Ts&&... ts) {
        output_.emplace_back(callback_(std::forward<Ts>(ts)...));
    }
    const auto& get_output() const { return output_; }
    const Callback& get_callback() const { return callback_; }

private:
    Callback callback_;
    util::aligned::vector<typename Callback::return_type> output_;
};

}  // namespace core
}  // namespace wayverb
