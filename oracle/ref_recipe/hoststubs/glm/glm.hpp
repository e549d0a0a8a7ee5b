// oracle/ref_recipe/hoststubs/glm/glm.hpp -- TEST INFRASTRUCTURE ONLY.
//
// Stand-in for the handful of GLM 0.9.8.1 names the reference's HOST scene code uses
// (tri_cube_intersection.cpp, box.cpp, geometric.cpp, triangle_vec.cpp, ndim_tree.h,
// voxel_collection.h/.cpp, voxelised_scene_data.h, indexing.h, utilities/range.h, and the
// image-source files raytracer/src/image_source/{tree,postprocess_branches,exact}.cpp), so that those
// files can be compiled here UNMODIFIED (GLM is fetched at configure time by the reference,
// config/dependencies.cmake, and is not in this image). Semantics follow GLM's generic (non-SIMD)
// code paths:
//   operators        componentwise
//   dot(a, b)        tmp = a * b; tmp.x + tmp.y + tmp.z            (func_geometric.inl compute_dot)
//   cross(x, y)      (x.y*y.z - y.y*x.z, x.z*y.x - y.z*x.x, x.x*y.y - y.x*x.y)
//   normalize(v)     v * inversesqrt(dot(v, v)),  inversesqrt(x) = 1 / sqrt(x)
//   length(v)        sqrt(dot(v, v));  distance(a, b) = length(b - a)
//   mix(a, b, t)     a + t * (b - a) for a float t;  componentwise select (t ? b : a) for a bool vector
//   lessThan / lessThanEqual / equal / isnan     componentwise, to a bool vector; any / all fold it
//   vec -> ivec      componentwise static_cast (truncation), ceil / round componentwise std::ceil /
//                    std::round (GLM_HAS_CXX11_STL: func_common.inl takes ::std::round)
//   min / max        (y < x) ? y : x  /  (x < y) ? y : x, componentwise
// These are the only places where this file, and not the reference's source, decides arithmetic.
#pragma once

#include <cassert>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <limits>   // GLM pulls these in; core/almost_equal.h (numeric_limits) relies on that
#include <type_traits>

namespace glm {

template <typename T>
struct tvec2 {
    T x, y;
    constexpr tvec2() : x{}, y{} {}
    constexpr explicit tvec2(T s) : x(s), y(s) {}
    constexpr tvec2(T a, T b) : x(a), y(b) {}
    template <typename U>
    constexpr tvec2(const tvec2<U>& v) : x(T(v.x)), y(T(v.y)) {}
    template <typename V, typename = decltype(V::z)>
    constexpr tvec2(const V& v) : x(T(v.x)), y(T(v.y)) {}  // GLM_EXPLICIT is empty: vec3 -> vec2 truncates
    T& operator[](size_t i) { return i == 0 ? x : y; }
    constexpr const T& operator[](size_t i) const { return i == 0 ? x : y; }
    tvec2& operator+=(const tvec2& o) { x += o.x; y += o.y; return *this; }
    tvec2& operator-=(const tvec2& o) { x -= o.x; y -= o.y; return *this; }
    tvec2& operator*=(const tvec2& o) { x *= o.x; y *= o.y; return *this; }
    tvec2& operator/=(const tvec2& o) { x /= o.x; y /= o.y; return *this; }
};

template <typename T>
struct tvec3 {
    T x, y, z;
    constexpr tvec3() : x{}, y{}, z{} {}
    constexpr explicit tvec3(T s) : x(s), y(s), z(s) {}
    constexpr tvec3(T a, T b, T c) : x(a), y(b), z(c) {}
    template <typename A, typename B, typename C>
    constexpr tvec3(A a, B b, C c) : x(T(a)), y(T(b)), z(T(c)) {}
    template <typename U>
    constexpr tvec3(const tvec3<U>& v) : x(T(v.x)), y(T(v.y)), z(T(v.z)) {}
    T& operator[](size_t i) { return i == 0 ? x : (i == 1 ? y : z); }
    constexpr const T& operator[](size_t i) const { return i == 0 ? x : (i == 1 ? y : z); }
    tvec3& operator+=(const tvec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    tvec3& operator-=(const tvec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    tvec3& operator*=(const tvec3& o) { x *= o.x; y *= o.y; z *= o.z; return *this; }
    tvec3& operator/=(const tvec3& o) { x /= o.x; y /= o.y; z /= o.z; return *this; }
    template <typename U> tvec3& operator*=(U s) { x *= T(s); y *= T(s); z *= T(s); return *this; }
    template <typename U> tvec3& operator/=(U s) { x /= T(s); y /= T(s); z /= T(s); return *this; }
};

using vec2 = tvec2<float>;
using uvec2 = tvec2<unsigned>;
using ivec2 = tvec2<int>;
using vec3 = tvec3<float>;
using uvec3 = tvec3<unsigned>;
using ivec3 = tvec3<int>;
using dvec3 = tvec3<double>;
struct mat4 final {   // named in core/orientation.h's declarations only; nothing compiled here touches a matrix
    float m[16];
};
using bvec3 = tvec3<bool>;
using bvec2 = tvec2<bool>;

#define GLM_STUB_BINOP(op)                                                                                     \
    template <typename T> constexpr tvec3<T> operator op(const tvec3<T>& a, const tvec3<T>& b) {               \
        return tvec3<T>(a.x op b.x, a.y op b.y, a.z op b.z);                                                   \
    }                                                                                                          \
    template <typename T> constexpr tvec3<T> operator op(const tvec3<T>& a, T s) {                             \
        return tvec3<T>(a.x op s, a.y op s, a.z op s);                                                         \
    }                                                                                                          \
    template <typename T> constexpr tvec3<T> operator op(T s, const tvec3<T>& a) {                             \
        return tvec3<T>(s op a.x, s op a.y, s op a.z);                                                         \
    }
GLM_STUB_BINOP(+)
GLM_STUB_BINOP(-)
GLM_STUB_BINOP(*)
GLM_STUB_BINOP(/)
GLM_STUB_BINOP(%)
GLM_STUB_BINOP(&)
GLM_STUB_BINOP(<<)
#undef GLM_STUB_BINOP
template <typename T> constexpr tvec3<T> operator-(const tvec3<T>& a) { return tvec3<T>(-a.x, -a.y, -a.z); }
template <typename T> constexpr bool operator==(const tvec3<T>& a, const tvec3<T>& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
template <typename T> constexpr bool operator!=(const tvec3<T>& a, const tvec3<T>& b) { return !(a == b); }
#define GLM_STUB_BINOP2(op)                                                                                    \
    template <typename T> constexpr tvec2<T> operator op(const tvec2<T>& a, const tvec2<T>& b) {               \
        return tvec2<T>(a.x op b.x, a.y op b.y);                                                               \
    }                                                                                                          \
    template <typename T> constexpr tvec2<T> operator op(const tvec2<T>& a, T s) { return tvec2<T>(a.x op s, a.y op s); } \
    template <typename T> constexpr tvec2<T> operator op(T s, const tvec2<T>& a) { return tvec2<T>(s op a.x, s op a.y); }
GLM_STUB_BINOP2(+)
GLM_STUB_BINOP2(-)
GLM_STUB_BINOP2(*)
GLM_STUB_BINOP2(/)
GLM_STUB_BINOP2(%)
GLM_STUB_BINOP2(&)
GLM_STUB_BINOP2(<<)
#undef GLM_STUB_BINOP2
template <typename T> constexpr bool operator==(const tvec2<T>& a, const tvec2<T>& b) { return a.x == b.x && a.y == b.y; }
template <typename T> constexpr bool operator!=(const tvec2<T>& a, const tvec2<T>& b) { return !(a == b); }

template <typename T> constexpr tvec3<T> min(const tvec3<T>& x, const tvec3<T>& y) {
    return tvec3<T>(y.x < x.x ? y.x : x.x, y.y < x.y ? y.y : x.y, y.z < x.z ? y.z : x.z);
}
template <typename T> constexpr tvec3<T> max(const tvec3<T>& x, const tvec3<T>& y) {
    return tvec3<T>(x.x < y.x ? y.x : x.x, x.y < y.y ? y.y : x.y, x.z < y.z ? y.z : x.z);
}
template <typename T> constexpr tvec2<T> min(const tvec2<T>& x, const tvec2<T>& y) { return tvec2<T>(y.x < x.x ? y.x : x.x, y.y < x.y ? y.y : x.y); }
template <typename T> constexpr tvec2<T> max(const tvec2<T>& x, const tvec2<T>& y) { return tvec2<T>(x.x < y.x ? y.x : x.x, x.y < y.y ? y.y : x.y); }
inline vec3 abs(const vec3& v) { return vec3(std::fabs(v.x), std::fabs(v.y), std::fabs(v.z)); }
inline float dot(const vec3& a, const vec3& b) {
    const vec3 tmp(a * b);
    return tmp.x + tmp.y + tmp.z;
}
inline vec3 cross(const vec3& x, const vec3& y) {
    return vec3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
inline vec3 normalize(const vec3& v) { return v * inversesqrt(dot(v, v)); }
inline float length(const vec3& v) { return std::sqrt(dot(v, v)); }
inline float distance(const vec3& a, const vec3& b) { return length(b - a); }
// the scalar forms (func_geometric.inl: length(x) = abs(x), distance(p0, p1) = length(p1 - p0))
inline float length(float x) { return std::fabs(x); }
inline double length(double x) { return std::fabs(x); }
inline float distance(float a, float b) { return length(b - a); }
inline double distance(double a, double b) { return length(b - a); }
inline vec3 mix(const vec3& a, const vec3& b, float t) { return a + t * (b - a); }
inline vec3 round(const vec3& v) { return vec3(std::round(v.x), std::round(v.y), std::round(v.z)); }
inline vec3 ceil(const vec3& v) { return vec3(std::ceil(v.x), std::ceil(v.y), std::ceil(v.z)); }
template <typename T> constexpr bvec3 lessThan(const tvec3<T>& a, const tvec3<T>& b) { return bvec3(a.x < b.x, a.y < b.y, a.z < b.z); }
template <typename T> constexpr bvec3 lessThanEqual(const tvec3<T>& a, const tvec3<T>& b) { return bvec3(a.x <= b.x, a.y <= b.y, a.z <= b.z); }
template <typename T> constexpr bvec3 equal(const tvec3<T>& a, const tvec3<T>& b) { return bvec3(a.x == b.x, a.y == b.y, a.z == b.z); }
template <typename T> constexpr bvec2 lessThan(const tvec2<T>& a, const tvec2<T>& b) { return bvec2(a.x < b.x, a.y < b.y); }
inline bvec3 isnan(const vec3& v) { return bvec3(std::isnan(v.x), std::isnan(v.y), std::isnan(v.z)); }
template <typename T> constexpr tvec3<T> mix(const tvec3<T>& x, const tvec3<T>& y, const bvec3& a) {
    return tvec3<T>(a.x ? y.x : x.x, a.y ? y.y : x.y, a.z ? y.z : x.z);
}
constexpr bool any(const bvec3& v) { return v.x || v.y || v.z; }
constexpr bool all(const bvec3& v) { return v.x && v.y && v.z; }
constexpr bool any(const bvec2& v) { return v.x || v.y; }
constexpr bool all(const bvec2& v) { return v.x && v.y; }

}  // namespace glm
