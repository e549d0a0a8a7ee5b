// oracle/ref_recipe/cl_prelude.hpp
//
// TEST INFRASTRUCTURE ONLY (see oracle/wg_oracle.cpp's header for who may load
// what is built from this).
//
// A small OpenCL-C 1.x emulation layer for g++, so that the reference's kernel
// source -- extracted VERBATIM at build time from the raw-string literals under
// /root/reference by oracle/ref_recipe/build.py -- compiles and runs on the host.
// Nothing in here restates the reference's algorithm: it only supplies what an
// OpenCL compiler supplies (address-space keywords, vector types, work-item ids,
// atomics, the handful of builtins the kernels call).
//
// Builtins whose rounding OpenCL leaves to the platform (dot, cross, length,
// normalize, distance, sin, cos) are DEFINED here the way oracle/rt_oracle.cpp and
// the CUDA kernels define them (see DESIGN.md "Precision"); everything else is the
// IEEE operation of the same name.
//
// CLC_REAL selects the element type of float3 (float for the reference's own
// arithmetic, double for the fp64 build that `#define float double` produces).
#pragma once

#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <type_traits>

#ifndef CLC_REAL
#define CLC_REAL float
#endif

namespace clc {

typedef unsigned int uint;
typedef unsigned long ulong;
typedef unsigned char uchar;
typedef unsigned short ushort;
typedef CLC_REAL real_t;
using std::size_t;

// ---- work-item functions -----------------------------------------------------
static thread_local size_t g_global_id = 0;
static size_t g_global_size = 0;
inline size_t get_global_id(int) { return g_global_id; }
inline size_t get_global_size(int) { return g_global_size; }

// ---- atomics ---------------------------------------------------------------------
inline int atomic_or(volatile int* p, int v) { return __atomic_fetch_or(const_cast<int*>(p), v, __ATOMIC_RELAXED); }
inline int atomic_xchg(volatile int* p, int v) { return __atomic_exchange_n(const_cast<int*>(p), v, __ATOMIC_RELAXED); }

// ---- scalar builtins ---------------------------------------------------------------
inline int popcount(int v) { return __builtin_popcount((unsigned)v); }
inline int popcount(uint v) { return __builtin_popcount(v); }
inline float sqrt(float v) { return __builtin_sqrtf(v); }
inline double sqrt(double v) { return __builtin_sqrt(v); }
inline float fabs(float v) { return __builtin_fabsf(v); }
inline double fabs(double v) { return __builtin_fabs(v); }
inline float floor(float v) { return __builtin_floorf(v); }
inline double floor(double v) { return __builtin_floor(v); }
inline float ceil(float v) { return __builtin_ceilf(v); }
inline double ceil(double v) { return __builtin_ceil(v); }
inline int isinf(float v) { return __builtin_isinf(v) ? 1 : 0; }
inline int isinf(double v) { return __builtin_isinf(v) ? 1 : 0; }
inline int isnan(float v) { return __builtin_isnan(v) ? 1 : 0; }
inline int isnan(double v) { return __builtin_isnan(v) ? 1 : 0; }
// scalar signbit: 1 if the sign bit is set, else 0 (OpenCL 1.2 s6.12.6) -- the VECTOR
// form further down returns -1 / 0. The reference relies on this difference
// (SURVEY.md s3.4 item 9).
inline int signbit(float v) { return __builtin_signbit(v) ? 1 : 0; }
inline int signbit(double v) { return __builtin_signbit(v) ? 1 : 0; }
inline float max(float a, float b) { return a < b ? b : a; }
inline float min(float a, float b) { return b < a ? b : a; }
inline double max(double a, double b) { return a < b ? b : a; }
inline double min(double a, double b) { return b < a ? b : a; }
inline int max(int a, int b) { return a < b ? b : a; }
inline int min(int a, int b) { return b < a ? b : a; }

// platform-defined in OpenCL; fixed here exactly as in oracle/rt_oracle.cpp:sincos_fixed
inline void sincos_fixed(float theta, float* s, float* c) {
    const float kf = __builtin_rintf(theta * 0.636619746685028076f);
    const int k = int(kf);
    float r = theta - kf * 1.57079625129699707f;
    r = r - kf * 7.54978941586159635e-08f;
    const float r2 = r * r;
    float ps = -1.9515295891e-4f;
    ps = ps * r2 + 8.3321608736e-3f;
    ps = ps * r2 + -1.6666654611e-1f;
    const float sin_r = r + (r * r2) * ps;
    float pc = 2.443315711809948e-5f;
    pc = pc * r2 + -1.388731625493765e-3f;
    pc = pc * r2 + 4.166664568298827e-2f;
    const float cos_r = (1.0f - 0.5f * r2) + (r2 * r2) * pc;
    switch (k & 3) {
        case 0: *s = sin_r; *c = cos_r; break;
        case 1: *s = cos_r; *c = -sin_r; break;
        case 2: *s = -sin_r; *c = -cos_r; break;
        default: *s = -cos_r; *c = sin_r; break;
    }
}
inline float sin(float t) { float s, c; sincos_fixed(t, &s, &c); return s; }
inline float cos(float t) { float s, c; sincos_fixed(t, &s, &c); return c; }

// ---- 3-vectors (16-byte aligned, 4 lanes of storage like cl_float3 / cl_int3) --------
template <typename T>
struct alignas(4 * sizeof(T)) vec3 {
    T x, y, z, w_;
    vec3() : x(0), y(0), z(0), w_(0) {}
    template <typename A, typename = std::enable_if_t<std::is_arithmetic<A>::value>>
    vec3(A a) : x(T(a)), y(T(a)), z(T(a)), w_(0) {}
    template <typename A, typename B, typename C>
    vec3(A a, B b, C c) : x(T(a)), y(T(b)), z(T(c)), w_(0) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
};
typedef vec3<real_t> float3;
typedef vec3<int> int3;
typedef vec3<uint> uint3;

#define CLC_VEC3_OP(op)                                                                          \
    template <typename T>                                                                        \
    inline vec3<T> operator op(vec3<T> a, vec3<T> b) { return vec3<T>(a.x op b.x, a.y op b.y, a.z op b.z); } \
    template <typename T, typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value>>  \
    inline vec3<T> operator op(vec3<T> a, S s) { return vec3<T>(a.x op T(s), a.y op T(s), a.z op T(s)); }    \
    template <typename T, typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value>>  \
    inline vec3<T> operator op(S s, vec3<T> b) { return vec3<T>(T(s) op b.x, T(s) op b.y, T(s) op b.z); }    \
    template <typename T, typename R>                                                            \
    inline vec3<T>& operator op##=(vec3<T>& a, R b) { a = a op b; return a; }
CLC_VEC3_OP(+)
CLC_VEC3_OP(-)
CLC_VEC3_OP(*)
CLC_VEC3_OP(/)
#undef CLC_VEC3_OP
template <typename T>
inline vec3<T> operator-(vec3<T> a) { return vec3<T>(-a.x, -a.y, -a.z); }
// int3 * float3 appears once in the reference (closest_triangle_in_voxel, never called
// by a kernel); AMD's compiler took it as a conversion, so do we.
inline float3 operator*(int3 a, float3 b) { return float3(real_t(a.x) * b.x, real_t(a.y) * b.y, real_t(a.z) * b.z); }

// relational operators on vectors give -1 (all bits set) / 0 per lane
#define CLC_VEC3_REL(op)                                                                 \
    template <typename T>                                                                \
    inline int3 operator op(vec3<T> a, vec3<T> b) { return int3(a.x op b.x ? -1 : 0, a.y op b.y ? -1 : 0, a.z op b.z ? -1 : 0); }
CLC_VEC3_REL(<)
CLC_VEC3_REL(<=)
CLC_VEC3_REL(>)
CLC_VEC3_REL(>=)
CLC_VEC3_REL(==)
CLC_VEC3_REL(!=)
#undef CLC_VEC3_REL
inline int any(int3 v) { return (v.x < 0 || v.y < 0 || v.z < 0) ? 1 : 0; }
inline int all(int3 v) { return (v.x < 0 && v.y < 0 && v.z < 0) ? 1 : 0; }
inline int3 signbit(float3 v) { return int3(__builtin_signbit(v.x) ? -1 : 0, __builtin_signbit(v.y) ? -1 : 0, __builtin_signbit(v.z) ? -1 : 0); }
inline int3 isnan(float3 v) { return int3(v.x != v.x ? -1 : 0, v.y != v.y ? -1 : 0, v.z != v.z ? -1 : 0); }
template <typename T>
inline vec3<T> select(vec3<T> a, vec3<T> b, int3 c) { return vec3<T>(c.x < 0 ? b.x : a.x, c.y < 0 ? b.y : a.y, c.z < 0 ? b.z : a.z); }
inline float3 fabs(float3 v) { return float3(fabs(v.x), fabs(v.y), fabs(v.z)); }
inline float3 floor(float3 v) { return float3(floor(v.x), floor(v.y), floor(v.z)); }
inline float3 ceil(float3 v) { return float3(ceil(v.x), ceil(v.y), ceil(v.z)); }
template <typename T>
inline vec3<T> max(vec3<T> a, vec3<T> b) { return vec3<T>(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
template <typename T>
inline vec3<T> min(vec3<T> a, vec3<T> b) { return vec3<T>(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline int3 convert_int3(float3 v) { return int3(int(v.x), int(v.y), int(v.z)); }
inline float3 convert_float3(int3 v) { return float3(real_t(v.x), real_t(v.y), real_t(v.z)); }
// platform-defined rounding in OpenCL; fixed as in rt_oracle.cpp
inline real_t dot(float3 a, float3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float3 cross(float3 a, float3 b) { return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline real_t length(float3 a) { return sqrt(dot(a, a)); }
inline float3 normalize(float3 a) { return a * (real_t(1) / sqrt(dot(a, a))); }
inline real_t distance(float3 a, float3 b) { return length(a - b); }

// ---- float8 (bands_type) -----------------------------------------------------------------
struct alignas(32) float8 {
    float s0, s1, s2, s3, s4, s5, s6, s7;
    float8() : s0(0), s1(0), s2(0), s3(0), s4(0), s5(0), s6(0), s7(0) {}
    template <typename A, typename = std::enable_if_t<std::is_arithmetic<A>::value>>
    float8(A a) : s0(float(a)), s1(float(a)), s2(float(a)), s3(float(a)), s4(float(a)), s5(float(a)), s6(float(a)), s7(float(a)) {}
    float& operator[](int i) { return (&s0)[i]; }
    const float& operator[](int i) const { return (&s0)[i]; }
};
#define CLC_VEC8_OP(op)                                                                         \
    inline float8 operator op(float8 a, float8 b) { float8 r; for (int i = 0; i < 8; ++i) r[i] = a[i] op b[i]; return r; } \
    template <typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value>>             \
    inline float8 operator op(float8 a, S s) { return a op float8(s); }                          \
    template <typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value>>             \
    inline float8 operator op(S s, float8 b) { return float8(s) op b; }
CLC_VEC8_OP(+)
CLC_VEC8_OP(-)
CLC_VEC8_OP(*)
CLC_VEC8_OP(/)
#undef CLC_VEC8_OP
inline float8 sqrt(float8 a) { float8 r; for (int i = 0; i < 8; ++i) r[i] = sqrt(a[i]); return r; }

}  // namespace clc

// ---- OpenCL-C keywords ----------------------------------------------------------------------
#define kernel
#define __kernel
#define global
#define __global
#define constant const
#define __constant const
#define local
#define __local
#define M_PI_F 3.14159274101257324f
