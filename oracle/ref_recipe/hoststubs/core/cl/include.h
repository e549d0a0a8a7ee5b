// TEST INFRASTRUCTURE ONLY. Stands in for core/cl/include.h (which includes CL/cl.hpp): the
// OpenCL host scalar / vector typedefs the compiled headers name, laid out as Khronos'
// cl_platform.h publishes them -- `T s[N]` first, size and alignment sizeof(T) * N, the 3-vectors
// typedefs of the 4-vectors. No OpenCL runtime is declared: none of the code compiled behind this
// header touches a context, queue or buffer.
#pragma once
#include <cstdint>
typedef int8_t cl_char;
typedef uint8_t cl_uchar;
typedef int16_t cl_short;
typedef uint16_t cl_ushort;
typedef int32_t cl_int;
typedef uint32_t cl_uint;
typedef int64_t cl_long;
typedef uint64_t cl_ulong;
typedef float cl_float;
typedef double cl_double;

#define WVB_STUB_CL_VECTOR(T, N) \
    struct alignas(sizeof(T) * N) T##N { \
        T s[N]; \
    };
#define WVB_STUB_CL_FAMILY(T) \
    WVB_STUB_CL_VECTOR(T, 2) WVB_STUB_CL_VECTOR(T, 4) WVB_STUB_CL_VECTOR(T, 8) WVB_STUB_CL_VECTOR(T, 16) \
    typedef T##4 T##3;
WVB_STUB_CL_FAMILY(cl_char)
WVB_STUB_CL_FAMILY(cl_uchar)
WVB_STUB_CL_FAMILY(cl_short)
WVB_STUB_CL_FAMILY(cl_ushort)
WVB_STUB_CL_FAMILY(cl_int)
WVB_STUB_CL_FAMILY(cl_uint)
WVB_STUB_CL_FAMILY(cl_long)
WVB_STUB_CL_FAMILY(cl_ulong)
WVB_STUB_CL_FAMILY(cl_float)
WVB_STUB_CL_FAMILY(cl_double)
#undef WVB_STUB_CL_FAMILY
#undef WVB_STUB_CL_VECTOR
