// cl_compat.hpp -- the few OpenCL host names the reference's headers spell out,
// re-pointed at the CUDA-backed handles of libwvb200.so.
//
// The reference's callback signatures are `(cl::CommandQueue&, cl::Buffer&, size_t)`
// (waveguide.h:80,121; preprocessor/hard_source.h:17; postprocessor/node.cpp:14) and
// its PODs are spelled with cl_float3 / cl_int3 (mesh_descriptor.h:14-20). Keeping
// those names lets the reference's processors and call sites compile unchanged
// against this shim instead of against cl.hpp.
#pragma once

#include <cstddef>
#include <cstdint>

#include "../wvb200.h"

typedef float cl_float;
typedef double cl_double;
typedef int32_t cl_int;
typedef uint32_t cl_uint;
typedef int8_t cl_char;
struct alignas(16) cl_float3 {
    float s[4];
};
struct alignas(16) cl_int3 {
    int32_t s[4];
};

namespace cl {

/// Stands in for the in-order command queue waveguide::run creates
/// (waveguide.h:46): every operation of a wvb_wg handle is ordered on the
/// handle's own CUDA stream.
class CommandQueue final {
public:
    explicit CommandQueue(wvb_wg* wg = nullptr) : wg_{wg} {}
    wvb_wg* handle() const { return wg_; }

private:
    wvb_wg* wg_;
};

/// Stands in for the `current` pressure buffer handed to the callbacks. It is a
/// borrowed view (valid during the callback), element type double on the device;
/// core::read_value / write_value convert to the caller's T.
class Buffer final {
public:
    Buffer(wvb_wg* wg = nullptr, size_t items = 0) : wg_{wg}, items_{items} {}
    wvb_wg* handle() const { return wg_; }
    size_t items() const { return items_; }

private:
    wvb_wg* wg_;
    size_t items_;
};

}  // namespace cl
