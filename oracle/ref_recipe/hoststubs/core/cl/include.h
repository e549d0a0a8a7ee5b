// TEST INFRASTRUCTURE ONLY. Stands in for core/cl/include.h (which includes CL/cl.hpp): the
// OpenCL host scalar / vector typedefs the compiled headers name.
#pragma once
#include <cstdint>
typedef uint32_t cl_uint;
typedef int32_t cl_int;
typedef float cl_float;
struct alignas(16) cl_float3 {
    float s[4];
};
