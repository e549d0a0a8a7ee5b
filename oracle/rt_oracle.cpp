// oracle/rt_oracle.cpp
//
// TEST INFRASTRUCTURE ONLY -- NOT PART OF THE PRODUCT PATH (see wg_oracle.cpp's
// header for who may load this).
//
// CPU restatement (C++17 + OpenMP over rays) of wayverb's stochastic
// ray-reflection loop, written by reading the reference's OpenCL kernel strings
// and host drivers. Paths relative to /root/reference:
//
//   reflections kernel             src/raytracer/src/program.cpp:59-153
//   init_reflections               src/raytracer/src/program.cpp:51-57
//   sphere_point / lambert_*       src/raytracer/src/cl/brdf.cpp:7-35
//   mean                           src/raytracer/src/cl/brdf.cpp:100-103
//   stochastic kernel              src/raytracer/src/stochastic/program.cpp:58-152
//   init_stochastic_path_info      src/raytracer/src/stochastic/program.cpp:51-56
//   Moller-Trumbore + helpers      src/core/src/cl/geometry.cpp:7-164
//   voxel DDA, point visibility    src/core/src/cl/voxel.cpp:7-95,227-258
//   flattened voxel layout         src/core/src/spatial_division/voxel_collection.cpp:9-37
//   compute_ray_energy             src/raytracer/include/raytracer/stochastic/finder.h:18-25,
//                                  src/raytracer/src/stochastic/finder.cpp:7-15
//   segment/depth loop             src/raytracer/include/raytracer/raytracer.h:188-266
//   histogram binning              src/raytracer/include/raytracer/reflection_processor/stochastic_histogram.h:17-32,70-111
//   image-source stage             see is_oracle.inc (included at the end)
//   direction -> LUT cell          src/core/include/core/vector_look_up_table.h:53-116, src/core/src/az_el.cpp:53-68
//
// All ray arithmetic is fp32 in the reference and here; operation order is
// written out explicitly and the file is built with -ffp-contract=off.
//
// Two things the reference leaves to the platform are DEFINED here (and
// identically in the CUDA code) so that runs are reproducible bit for bit:
//   * random numbers: the reference draws from std::default_random_engine
//     seeded by std::random_device per step (reflector.cpp:13-25) -- nothing is
//     reproducible there. We use Philox4x32-10 keyed by the seed, counter =
//     (global ray index, step, stream).
//   * sin/cos/normalize: OpenCL's cos/sin/normalize have implementation-defined
//     rounding. We use a fixed Cody-Waite + polynomial sincos for theta in
//     [-pi, pi] and normalize(v) = v * (1 / sqrt(dot(v, v))).
// Parity pinning: PINNED to reference-run output. The reference's own `reflections` and
// `stochastic` kernels with their geometry / voxel / brdf sources are compiled for the host into
// oracle/_ref by oracle/ref_recipe/build.py; tests/test_ref_pin_rt.py feeds both sides the same
// directions and random stream and asserts the reflection records of every step bit-identical
// and the histograms equal to 1e-12. The HOST side runs too: the reference's own raytracer::run
// template with reflector.cpp, finder.cpp and the reflection processors
// (tests/test_ref_pin_ray_run.py), its histogram binning and look-up-table indexing
// (incremental_histogram, vector_look_up_table::index), compute_ray_energy and
// compute_optimum_reflection_number (tests/test_ref_pin_hostmath.py). On top: the reference's CPU-twin tests restated in
// tests/test_rt_oracle_kats.py (brute-force == voxel traversal, analytic shoebox image sources,
// energy equivalence).

#ifdef _OPENMP
#include <omp.h>
#endif
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

struct f3 {
    float x, y, z;
};
struct cl_f3 {  // cl_float3: 16 bytes
    float x, y, z, w;
};
struct ray_t {  // core::ray, 32 B              core/cl/geometry_structs.h:9-12
    cl_f3 position, direction;
};
struct triangle_t {  // core::triangle, 16 B    core/cl/triangle.h:8-13
    uint32_t surface, v0, v1, v2;
};
struct surface_t {  // core::surface<8>, 64 B   core/cl/scene_structs.h:24-30
    float absorption[8];
    float scattering[8];
};
struct reflection_t {  // raytracer::reflection, 32 B   raytracer/cl/reflection.h:10-17
    cl_f3 position;
    uint32_t triangle;
    int8_t keep_going;
    int8_t receiver_visible;
    int8_t pad_[10];
};
static_assert(sizeof(ray_t) == 32 && sizeof(triangle_t) == 16 && sizeof(surface_t) == 64 &&
                      sizeof(reflection_t) == 32,
              "reference POD layouts");

struct inter_t {  // triangle_inter + index  core/cl/geometry_structs.h:30-60
    float t, u, v;
    uint32_t index;
};

// ---- fixed-order float3 helpers ------------------------------------------------
inline f3 mk(float x, float y, float z) { return {x, y, z}; }
inline f3 add(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline f3 sub(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline f3 mul(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline f3 mulv(f3 a, f3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline f3 divv(f3 a, f3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline float dot(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline f3 cross(f3 a, f3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float length(f3 a) { return std::sqrt(dot(a, a)); }
inline f3 normalize(f3 a) { return mul(a, 1.0f / std::sqrt(dot(a, a))); }
inline float comp(f3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// ---- Philox4x32-10 ----------------------------------------------------------------
inline void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                   uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = uint64_t(0xD2511F53u) * c0;
        const uint64_t p1 = uint64_t(0xCD9E8D57u) * c2;
        const uint32_t n0 = uint32_t(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = uint32_t(p1);
        const uint32_t n2 = uint32_t(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = uint32_t(p0);
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// z in [-1, 1), theta in [-pi, pi): the ranges of core::direction_rng
// (core/azimuth_elevation.h:15-29)
inline void direction_rng(uint64_t seed, uint32_t ray, uint32_t step, uint32_t stream, float* z,
                          float* theta) {
    uint32_t o[4];
    philox(ray, step, stream, 0u, uint32_t(seed), uint32_t(seed >> 32), o);
    const float u0 = float(o[0] >> 8) * 5.9604644775390625e-08f;  // 2^-24
    const float u1 = float(o[1] >> 8) * 5.9604644775390625e-08f;
    *z = 2.0f * u0 - 1.0f;
    *theta = (2.0f * u1 - 1.0f) * 3.14159274101257324f;
}

// ---- sincos on [-pi, pi] -------------------------------------------------------------
inline void sincos_fixed(float theta, float* s, float* c) {
    const float kf = std::rint(theta * 0.636619746685028076f);  // 2/pi
    const int k = int(kf);
    float r = theta - kf * 1.57079625129699707f;  // pi/2 high part
    r = r - kf * 7.54978941586159635e-08f;        // pi/2 low part
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

// sphere_point                  brdf.cpp:7-11
inline f3 sphere_point(float z, float theta) {
    const float t = std::sqrt(1 - z * z);
    float s, c;
    sincos_fixed(theta, &s, &c);
    return {t * c, z, t * s};
}

// ---- scene ------------------------------------------------------------------------------
struct scene_t {
    std::vector<uint32_t> voxel_index;
    f3 c0, c1;  // aabb
    uint32_t side = 0;
    std::vector<triangle_t> triangles;
    std::vector<cl_f3> vertices;
    std::vector<surface_t> surfaces;
    f3 vert(uint32_t i) const { return {vertices[i].x, vertices[i].y, vertices[i].z}; }
};

// almost_equal                  geometry.cpp:7-11
inline bool almost_equal(float x, float y, float ulp) {
    const float abs_diff = std::fabs(x - y);
    return abs_diff < FLT_EPSILON * std::fabs(x + y) * ulp || abs_diff < FLT_MIN;
}
constexpr float ULP = 10.0f;

// triangle_vert_intersection    geometry.cpp:20-54
inline void tri_intersection(f3 v0, f3 v1, f3 v2, f3 pos, f3 dir, float* t, float* u_out,
                             float* v_out) {
    *t = *u_out = *v_out = 0;
    const f3 e0 = sub(v1, v0);
    const f3 e1 = sub(v2, v0);
    const f3 pvec = cross(dir, e1);
    const float det = dot(e0, pvec);
    if (almost_equal(det, 0, ULP)) return;
    const float invdet = 1.0f / det;
    const f3 tvec = sub(pos, v0);
    const float u = invdet * dot(tvec, pvec);
    if (u < 0.0f || 1.0f < u) return;
    const f3 qvec = cross(tvec, e0);
    const float v = invdet * dot(dir, qvec);
    if (v < 0.0f || 1.0f < v + u) return;
    const float tt = invdet * dot(e1, qvec);
    if (tt < 0 || almost_equal(tt, 0, ULP)) return;
    *t = tt; *u_out = u; *v_out = v;
}

// ray_triangle_group_intersection + INTERSECTION_ACCUMULATOR   geometry.cpp:103-148
inline inter_t group_intersection(const scene_t& sc, f3 pos, f3 dir, const uint32_t* indices,
                                  uint32_t n, uint32_t avoid) {
    inter_t ret{0, 0, 0, 0};
    for (uint32_t i = 0; i != n; ++i) {
        const uint32_t ti = indices[i];
        if (ti != avoid) {
            const triangle_t tri = sc.triangles[ti];
            float t, u, v;
            tri_intersection(sc.vert(tri.v0), sc.vert(tri.v1), sc.vert(tri.v2), pos, dir, &t, &u, &v);
            if (t && (!ret.t || t < ret.t)) {
                ret.index = ti;
                ret.t = t; ret.u = u; ret.v = v;
            }
        }
    }
    return ret;
}

// brute force over all triangles (ray_triangle_intersection, geometry.cpp:116-128);
// the reference's CPU twin test compares it with the voxel traversal
inline inter_t brute_intersection(const scene_t& sc, f3 pos, f3 dir, uint32_t avoid) {
    inter_t ret{0, 0, 0, 0};
    for (uint32_t ti = 0; ti != sc.triangles.size(); ++ti) {
        if (ti != avoid) {
            const triangle_t tri = sc.triangles[ti];
            float t, u, v;
            tri_intersection(sc.vert(tri.v0), sc.vert(tri.v1), sc.vert(tri.v2), pos, dir, &t, &u, &v);
            if (t && (!ret.t || t < ret.t)) {
                ret.index = ti;
                ret.t = t; ret.u = u; ret.v = v;
            }
        }
    }
    return ret;
}

// VOXEL_TRAVERSAL_ALGORITHM + voxel_traversal     voxel.cpp:22-95
inline inter_t voxel_traversal(const scene_t& sc, f3 pos, f3 dir, uint32_t avoid) {
    const float sidef = float(sc.side);
    const f3 vd = mk((sc.c1.x - sc.c0.x) / sidef, (sc.c1.y - sc.c0.y) / sidef,
                     (sc.c1.z - sc.c0.z) / sidef);
    const f3 rel = divv(sub(pos, sc.c0), vd);
    int ind[3] = {int(std::floor(rel.x)), int(std::floor(rel.y)), int(std::floor(rel.z))};
    const int side = int(sc.side);
    if (!(0 <= ind[0] && 0 <= ind[1] && 0 <= ind[2] && ind[0] < side && ind[1] < side &&
          ind[2] < side)) {
        return inter_t{0, 0, 0, 0};
    }
    const f3 lo = add(sc.c0, mulv(mk(float(ind[0]), float(ind[1]), float(ind[2])), vd));
    const f3 hi = add(sc.c0, mulv(mk(float(ind[0] + 1), float(ind[1] + 1), float(ind[2] + 1)), vd));
    int step[3], just_out[3];
    float t_max[3], t_delta[3];
    for (int i = 0; i < 3; ++i) {
        const float d = comp(dir, i);
        const bool neg = std::signbit(d);
        step[i] = neg ? -1 : 1;
        just_out[i] = neg ? -1 : side;
        const float boundary = neg ? comp(lo, i) : comp(hi, i);
        const float tmp = std::fabs((boundary - comp(pos, i)) / d);
        t_max[i] = std::isnan(tmp) ? INFINITY : tmp;
        t_delta[i] = std::fabs(comp(vd, i) / d);
    }
    for (;;) {
        int min_i = 0;
        for (int i = 1; i != 3; ++i) {
            if (t_max[i] < t_max[min_i]) min_i = i;
        }
        const uint32_t voxel_offset =
                sc.voxel_index[size_t(ind[0]) * side * side + size_t(ind[1]) * side + ind[2]];
        const uint32_t num = sc.voxel_index[voxel_offset];
        const uint32_t* begin = sc.voxel_index.data() + voxel_offset + 1;
        const float max_dist = t_max[min_i];
        const inter_t state = group_intersection(sc, pos, dir, begin, num, avoid);
        if (state.t && state.t <= max_dist) return state;
        ind[min_i] += step[min_i];
        if (ind[min_i] == just_out[min_i]) break;
        t_max[min_i] += t_delta[min_i];
    }
    return inter_t{0, 0, 0, 0};
}

// count_intersections           voxel.cpp:97-133: number of triangle crossings along a
// ray, ~0u if any crossing is degenerate (near an edge or vertex, geometry.cpp:13-18)
inline uint32_t count_intersections(const scene_t& sc, f3 pos, f3 dir) {
    const float sidef = float(sc.side);
    const f3 vd = mk((sc.c1.x - sc.c0.x) / sidef, (sc.c1.y - sc.c0.y) / sidef,
                     (sc.c1.z - sc.c0.z) / sidef);
    const f3 rel = divv(sub(pos, sc.c0), vd);
    int ind[3] = {int(std::floor(rel.x)), int(std::floor(rel.y)), int(std::floor(rel.z))};
    const int side = int(sc.side);
    uint32_t count = 0;
    if (!(0 <= ind[0] && 0 <= ind[1] && 0 <= ind[2] && ind[0] < side && ind[1] < side &&
          ind[2] < side)) {
        return count;
    }
    const f3 lo = add(sc.c0, mulv(mk(float(ind[0]), float(ind[1]), float(ind[2])), vd));
    const f3 hi = add(sc.c0, mulv(mk(float(ind[0] + 1), float(ind[1] + 1), float(ind[2] + 1)), vd));
    int step[3], just_out[3];
    float t_max[3], t_delta[3];
    for (int i = 0; i < 3; ++i) {
        const float d = comp(dir, i);
        const bool neg = std::signbit(d);
        step[i] = neg ? -1 : 1;
        just_out[i] = neg ? -1 : side;
        const float boundary = neg ? comp(lo, i) : comp(hi, i);
        const float tmp = std::fabs((boundary - comp(pos, i)) / d);
        t_max[i] = std::isnan(tmp) ? INFINITY : tmp;
        t_delta[i] = std::fabs(comp(vd, i) / d);
    }
    float prev_max = 0;
    for (;;) {
        int min_i = 0;
        for (int i = 1; i != 3; ++i) {
            if (t_max[i] < t_max[min_i]) min_i = i;
        }
        const uint32_t voxel_offset =
                sc.voxel_index[size_t(ind[0]) * side * side + size_t(ind[1]) * side + ind[2]];
        const uint32_t num = sc.voxel_index[voxel_offset];
        const uint32_t* begin = sc.voxel_index.data() + voxel_offset + 1;
        const float max_dist = t_max[min_i];
        for (uint32_t i = 0; i != num; ++i) {
            const triangle_t tri = sc.triangles[begin[i]];
            float t, u, v;
            tri_intersection(sc.vert(tri.v0), sc.vert(tri.v1), sc.vert(tri.v2), pos, dir, &t, &u, &v);
            if (t) {
                // is_degenerate (geometry.cpp:13-18)
                if (almost_equal(u, 0, ULP) || almost_equal(v, 0, ULP) || almost_equal(u + v, 1, ULP)) {
                    return ~0u;
                }
                if (prev_max < t && t <= max_dist) count += 1;
            }
        }
        ind[min_i] += step[min_i];
        if (ind[min_i] == just_out[min_i]) break;
        prev_max = t_max[min_i];
        t_max[min_i] += t_delta[min_i];
    }
    return count;
}

// the 32 fixed probe directions of voxel_inside        voxel.cpp:156-189
static const float kInsideDirections[32][3] = {
        {-0.427602f, 0.791267f, -0.437096f},  {-0.832527f, -0.545442f, 0.0969113f},
        {0.633363f, 0.413131f, 0.65435f},     {0.985873f, 0.140209f, 0.0916325f},
        {0.384519f, 0.0309011f, -0.9226f},    {-0.532584f, -0.0244727f, 0.846023f},
        {0.844848f, 0.230031f, -0.483029f},   {-0.186143f, -0.291698f, -0.938223f},
        {-0.108511f, -0.861706f, 0.495669f},  {0.0951741f, 0.959367f, -0.265625f},
        {0.407194f, 0.907127f, -0.106369f},   {0.521731f, -0.00522727f, -0.853094f},
        {0.369627f, 0.218276f, 0.903179f},    {-0.518837f, 0.815586f, -0.25618f},
        {-0.954901f, 0.105507f, 0.277548f},   {0.63419f, 0.768703f, 0.0830607f},
        {-0.0258027f, 0.998294f, 0.052379f},  {-0.868361f, 0.473347f, 0.147958f},
        {0.346294f, -0.131168f, 0.928911f},   {-0.635896f, 0.649019f, 0.417624f},
        {0.293121f, 0.235495f, -0.926619f},   {-0.55088f, -0.0237137f, -0.834247f},
        {-0.661022f, -0.653122f, -0.369434f}, {0.224176f, -0.351092f, 0.909109f},
        {0.456587f, 0.736627f, -0.498907f},   {0.965231f, 0.154753f, 0.210667f},
        {0.626034f, -0.245898f, 0.740011f},   {0.435825f, 0.794758f, -0.422393f},
        {0.662049f, 0.713267f, 0.23009f},     {0.261843f, -0.620862f, 0.738897f},
        {0.23673f, 0.714889f, 0.657946f},     {-0.404007f, 0.699316f, 0.589691f},
};

// voxel_inside + single_ray_inside     voxel.cpp:135-154,191-225
inline bool voxel_inside(const scene_t& sc, f3 pt) {
    for (int i = 0; i != 32; ++i) {
        const f3 dir = mk(kInsideDirections[i][0], kInsideDirections[i][1], kInsideDirections[i][2]);
        const uint32_t n = count_intersections(sc, pt, dir);
        if (n == ~0u) continue;
        return (n % 2) != 0;
    }
    return false;
}

// point_triangle_distance_squared      boundary_coefficient_program.cpp:16-135
inline float point_triangle_distance_squared(f3 v0, f3 v1, f3 v2, f3 point) {
    const f3 diff = sub(point, v0);
    const f3 e0 = sub(v1, v0);
    const f3 e1 = sub(v2, v0);
    const float a00 = dot(e0, e0);
    const float a01 = dot(e0, e1);
    const float a11 = dot(e1, e1);
    const float b0 = -dot(diff, e0);
    const float b1 = -dot(diff, e1);
    const float det = a00 * a11 - a01 * a01;
    float t0 = a01 * b1 - a11 * b0;
    float t1 = a01 * b0 - a00 * b1;
    if (t0 + t1 <= det) {
        if (t0 < 0) {
            if (t1 < 0) {
                if (b0 < 0) {
                    t1 = 0;
                    if (a00 <= -b0) t0 = 1;
                    else t0 = -b0 / a00;
                } else {
                    t0 = 0;
                    if (0 <= b1) t1 = 0;
                    else if (a11 <= -b1) t1 = 1;
                    else t1 = -b1 / a11;
                }
            } else {
                t0 = 0;
                if (0 <= b1) t1 = 0;
                else if (a11 <= -b1) t1 = 1;
                else t1 = -b1 / a11;
            }
        } else if (t1 < 0) {
            t1 = 0;
            if (0 <= b0) t0 = 0;
            else if (a00 <= -b0) t0 = 1;
            else t0 = -b0 / a00;
        } else {
            const float invDet = 1 / det;
            t0 *= invDet;
            t1 *= invDet;
        }
    } else {
        if (t0 < 0) {
            const float tmp0 = a01 + b0;
            const float tmp1 = a11 + b1;
            if (tmp0 < tmp1) {
                const float numer = tmp1 - tmp0;
                const float denom = a00 - 2 * a01 + a11;
                if (denom <= numer) { t0 = 1; t1 = 0; }
                else { t0 = numer / denom; t1 = 1 - t0; }
            } else {
                t0 = 0;
                if (tmp1 <= 0) t1 = 1;
                else if (0 <= b1) t1 = 0;
                else t1 = -b1 / a11;
            }
        } else if (t1 < 0) {
            const float tmp0 = a01 + b1;
            const float tmp1 = a00 + b0;
            if (tmp0 < tmp1) {
                const float numer = tmp1 - tmp0;
                const float denom = a00 - 2 * a01 + a11;
                if (denom <= numer) { t1 = 1; t0 = 0; }
                else { t1 = numer / denom; t0 = 1 - t1; }
            } else {
                t1 = 0;
                if (tmp1 <= 0) t0 = 1;
                else if (0 <= b0) t0 = 0;
                else t0 = -b0 / a00;
            }
        } else {
            const float numer = a11 + b1 - a01 - b0;
            if (numer <= 0) { t0 = 0; t1 = 1; }
            else {
                const float denom = a00 - 2 * a01 + a11;
                if (denom <= numer) { t0 = 1; t1 = 0; }
                else { t0 = numer / denom; t1 = 1 - t0; }
            }
        }
    }
    const f3 closest = add(add(v0, mul(e0, t0)), mul(e1, t1));
    const f3 d = sub(point, closest);
    return dot(d, d);
}

// slow_closest_triangle         boundary_coefficient_program.cpp:222-241 (what the 1d finder uses, :338)
inline uint32_t slow_closest_triangle(const scene_t& sc, f3 pt) {
    uint32_t ret = 0;
    float distance = INFINITY;
    for (uint32_t i = 0; i != sc.triangles.size(); ++i) {
        const triangle_t t = sc.triangles[i];
        const float nd = point_triangle_distance_squared(sc.vert(t.v0), sc.vert(t.v1), sc.vert(t.v2), pt);
        if (nd < distance) {
            ret = i;
            distance = nd;
        }
    }
    return ret;
}

// voxel_point_intersection      voxel.cpp:227-258
inline bool point_visible(const scene_t& sc, f3 begin, f3 point, uint32_t avoid) {
    const f3 b2p = sub(point, begin);
    const float mag = length(b2p);
    const f3 direction = normalize(b2p);
    const inter_t inter = voxel_traversal(sc, begin, direction, avoid);
    return !inter.t || mag < inter.t;
}

// triangle_normal               geometry.cpp:69-81
inline f3 triangle_normal(const scene_t& sc, triangle_t tri) {
    const f3 v0 = sc.vert(tri.v0);
    return normalize(cross(sub(sc.vert(tri.v1), v0), sub(sc.vert(tri.v2), v0)));
}
// reflect                       geometry.cpp:83-86
inline f3 reflect(f3 normal, f3 direction) {
    return sub(direction, mul(mul(normal, 2), dot(direction, normal)));
}
// OpenCL scalar signbit(): 1 if the sign bit is set, else 0
inline float signbit_scalar(float x) { return std::signbit(x) ? 1.0f : 0.0f; }

// line_segment_sphere_intersection   geometry.cpp:155-164
inline bool segment_sphere(f3 p1, f3 p2, f3 sc, float r) {
    const f3 diff = sub(p2, p1);
    const float u = dot(sub(sc, p1), diff) / dot(diff, diff);
    if (u < 0 || 1 < u) return false;
    const f3 closest = sub(add(p1, mul(diff, u)), sc);
    return dot(closest, closest) < r * r;
}

// ---- directional cell: vector_look_up_table<.., 20, 9>::index ---------------------------
// compute_azimuth_elevation (az_el.cpp:53-68) + azimuth/elevation_to_index
// (vector_look_up_table.h:53-77,112-116). Host code in the reference: libm
// atan2f/asinf on float, the rest in double.
// core::degrees (vector_look_up_table.h:22): `radians * 180 / M_PI` on a float, returned as float --
// the product is rounded to float, the quotient is taken in double and rounded to float again
inline float degrees_f(float radians) { return float(double(radians * 180) / M_PI); }

inline void lut_index(f3 v, int* az_cell, int* el_cell) {
    float az = std::atan2(v.x, -v.z);
    const float el = std::asin(v.y);
    if (almost_equal(el, float(-M_PI / 2), 10) || almost_equal(el, float(M_PI / 2), 10)) az = 0;
    double a = degrees_f(-az);
    a += (360.0 / 20) / 2;
    while (a < 0) a += 360;
    *az_cell = int(size_t(a / (360.0 / 20)) % 20);
    double e = degrees_f(el);
    e += 90 + (180.0 / 10) / 2;
    while (e < 0) e += 360;
    size_t adj = size_t(e / (180.0 / 10)) % 20;
    if (adj < 1) adj = 1;
    if (adj > 9) adj = 9;
    *el_cell = int(adj - 1);
}

struct trace_params {
    float source[3];
    float receiver[3];
    float receiver_radius;
    float pad0;
    double speed_of_sound;
    double histogram_rate;
    uint64_t total_rays;       // compute_ray_energy's N (all rays of the whole run)
    uint64_t seed;
    uint64_t ray_index_base;   // global index of dirs[0] (slabs of one run share a seed)
    uint32_t depth;            // reflection_depth
    uint32_t specular_from_step;  // specular impulses are binned when step >= this
    uint32_t n_bins;           // histogram length (bins beyond are counted, not stored)
    uint32_t directional;      // 0: [bins][8]; 1: [20][9][bins][8]
    uint32_t keep_steps;       // reflections of steps < keep_steps are written out
    uint32_t pad1;
};

}  // namespace

extern "C" {

struct rto_scene {
    scene_t sc;
};

rto_scene* rto_scene_create(const uint32_t* voxel_index, size_t n_index, const float* aabb6,
                            uint32_t side, const void* triangles, size_t n_tri, const void* vertices,
                            size_t n_vert, const void* surfaces, size_t n_surf) {
    auto* s = new rto_scene;
    s->sc.voxel_index.assign(voxel_index, voxel_index + n_index);
    s->sc.c0 = {aabb6[0], aabb6[1], aabb6[2]};
    s->sc.c1 = {aabb6[3], aabb6[4], aabb6[5]};
    s->sc.side = side;
    s->sc.triangles.resize(n_tri);
    std::memcpy(s->sc.triangles.data(), triangles, n_tri * sizeof(triangle_t));
    s->sc.vertices.resize(n_vert);
    std::memcpy(s->sc.vertices.data(), vertices, n_vert * sizeof(cl_f3));
    s->sc.surfaces.resize(n_surf);
    std::memcpy(s->sc.surfaces.data(), surfaces, n_surf * sizeof(surface_t));
    return s;
}
void rto_scene_destroy(rto_scene* s) { delete s; }

// closest hit for n rays (pos[3], dir[3] packed as 6 floats), voxel or brute force
void rto_closest_hit(const rto_scene* s, const float* rays6, size_t n, int brute, uint32_t* tri_out,
                     float* t_out) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; ++i) {
        const f3 p{rays6[6 * i], rays6[6 * i + 1], rays6[6 * i + 2]};
        const f3 d{rays6[6 * i + 3], rays6[6 * i + 4], rays6[6 * i + 5]};
        const inter_t r = brute ? brute_intersection(s->sc, p, d, ~0u) : voxel_traversal(s->sc, p, d, ~0u);
        tri_out[i] = r.t ? r.index : ~0u;
        t_out[i] = r.t;
    }
}

// compute_ray_energy            finder.h:18-25, finder.cpp:7-15
float rto_ray_energy(uint64_t total_rays, const float* source, const float* receiver,
                     float receiver_radius) {
    const f3 d = sub(mk(source[0], source[1], source[2]), mk(receiver[0], receiver[1], receiver[2]));
    const float dist = length(d);
    const float sin_y = receiver_radius / std::fmax(receiver_radius, dist);
    const float cos_y = std::sqrt(1 - sin_y * sin_y);
    return float(2.0 / (4 * M_PI * double(total_rays) * dist * dist * (1 - cos_y)));
}

// initial directions the way the product generates them when the caller gives
// none: sphere_point(z, theta) from Philox stream 1 (random_unit_vector,
// core/azimuth_elevation.h:31-35)
void rto_directions(uint64_t seed, uint64_t base, size_t n, float* out3) {
    for (size_t i = 0; i < n; ++i) {
        float z, th;
        direction_rng(seed, uint32_t(base + i), 0u, 1u, &z, &th);
        const f3 d = sphere_point(z, th);
        out3[3 * i] = d.x; out3[3 * i + 1] = d.y; out3[3 * i + 2] = d.z;
    }
}

// set_node_inside (mesh_setup_program.cpp:110-140) for every node of a mesh
// descriptor: position = min_corner + locator * spacing (cl/utils.cpp:71-74), float
void rto_nodes_inside(const rto_scene* s, const float* min_corner, const int32_t* dim, float spacing,
                      uint8_t* inside_out) {
    const long long nn = (long long)dim[0] * dim[1] * dim[2];
#pragma omp parallel for schedule(dynamic, 1024)
    for (long long i = 0; i < nn; ++i) {
        const int x = int(i % dim[0]), y = int((i / dim[0]) % dim[1]), z = int(i / dim[0] / dim[1]);
        const f3 p = mk(min_corner[0] + float(x) * spacing, min_corner[1] + float(y) * spacing,
                        min_corner[2] + float(z) * spacing);
        inside_out[i] = voxel_inside(s->sc, p) ? 1 : 0;
    }
}
// the surface the 1d finder assigns to a node position (boundary_coefficient_program.cpp:323-342)
void rto_closest_surface(const rto_scene* s, const float* points3, size_t n, uint32_t* surface_out,
                         uint32_t* triangle_out) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; ++i) {
        const uint32_t t = slow_closest_triangle(s->sc, mk(points3[3 * i], points3[3 * i + 1], points3[3 * i + 2]));
        if (triangle_out) triangle_out[i] = t;
        surface_out[i] = s->sc.triangles[t].surface;
    }
}

// the (z, theta) pair ray `base + i` consumes at `step` (stream 0): what the reference's host
// draws per step and uploads as `rng` (reflector.cpp:13-25,35). Exposed so that the reference's
// own kernels (oracle/_ref) can be fed the identical stream.
void rto_step_rng(uint64_t seed, uint64_t base, size_t n, uint32_t step, float* out2) {
    for (size_t i = 0; i < n; ++i) direction_rng(seed, uint32_t(base + i), step, 0u, &out2[2 * i], &out2[2 * i + 1]);
}

void rto_sincos(const float* theta, size_t n, float* s, float* c) {
    for (size_t i = 0; i < n; ++i) sincos_fixed(theta[i], &s[i], &c[i]);
}

void rto_lut_index(const float* v3, size_t n, int32_t* az, int32_t* el) {
    for (size_t i = 0; i < n; ++i) {
        int a, e;
        lut_index(mk(v3[3 * i], v3[3 * i + 1], v3[3 * i + 2]), &a, &e);
        az[i] = a; el[i] = e;
    }
}

// The whole loop of raytracer::run for n rays (raytracer.h:223-244), with the
// stochastic histogram processor folded in. hist: double accumulators laid out
// [bins][8] or [20][9][bins][8]; *dropped counts impulses beyond n_bins.
// reflections_out: [keep_steps][n] reflection records (may be null).
void rto_trace(const rto_scene* s, const void* params_v, const float* dirs3, size_t n, double* hist,
               uint64_t* dropped, void* reflections_out) {
    trace_params P;
    std::memcpy(&P, params_v, sizeof P);
    const scene_t& sc = s->sc;
    const f3 source = mk(P.source[0], P.source[1], P.source[2]);
    const f3 receiver = mk(P.receiver[0], P.receiver[1], P.receiver[2]);
    const float energy = rto_ray_energy(P.total_rays, P.source, P.receiver, P.receiver_radius);
    reflection_t* rout = static_cast<reflection_t*>(reflections_out);
    uint64_t drop_total = 0;

#pragma omp parallel for schedule(dynamic, 256) reduction(+ : drop_total)
    for (long long ri = 0; ri < (long long)n; ++ri) {
        // reflector ctor + init_reflections (program.cpp:51-57)
        f3 rpos = source;
        f3 rdir = mk(dirs3[3 * ri], dirs3[3 * ri + 1], dirs3[3 * ri + 2]);
        bool keep_going = true;
        uint32_t prev_tri = ~0u;
        // init_stochastic_path_info (stochastic/program.cpp:51-56)
        float volume[8];
        for (int b = 0; b < 8; ++b) volume[b] = energy;
        f3 path_pos = source;
        float path_dist = 0;

        auto deposit = [&](const float* vol, f3 position, float distance) {
            // finder drops distance == 0 (finder.h:65-76); histogram_sum (:17-32)
            if (!distance) return;
            const double time = double(distance) / P.speed_of_sound;
            const size_t bin = size_t(time * P.histogram_rate);
            if (bin >= P.n_bins) {
                drop_total++;
                return;
            }
            size_t base = bin * 8;
            if (P.directional) {
                int az, el;
                lut_index(normalize(sub(position, receiver)), &az, &el);
                base = ((size_t(az) * 9 + el) * P.n_bins + bin) * 8;
            }
            for (int b = 0; b < 8; ++b) {
#pragma omp atomic
                hist[base + b] += double(vol[b]);
            }
        };

        for (uint32_t step = 0; step < P.depth; ++step) {
            // ---- reflections kernel (program.cpp:59-153) ----
            reflection_t refl;
            std::memset(&refl, 0, sizeof refl);
            f3 hit = mk(0, 0, 0);
            uint32_t hit_tri = 0;
            bool visible = false;
            bool alive = false;
            if (keep_going) {
                const inter_t ci = voxel_traversal(sc, rpos, rdir, prev_tri);
                if (ci.t) {
                    alive = true;
                    hit = add(rpos, mul(rdir, ci.t));
                    hit_tri = ci.index;
                    const triangle_t tri = sc.triangles[ci.index];
                    f3 tnorm = triangle_normal(sc, tri);
                    const f3 specular = reflect(tnorm, rdir);
                    tnorm = mul(tnorm, signbit_scalar(dot(tnorm, specular)));
                    visible = point_visible(sc, hit, receiver, ci.index);
                    float z, theta;
                    direction_rng(P.seed, uint32_t(P.ray_index_base + ri), step, 0u, &z, &theta);
                    const f3 rnd = sphere_point(z, theta);
                    const surface_t& sf = sc.surfaces[tri.surface];
                    const float* sv = sf.scattering;
                    const float scatter =
                            (sv[0] + sv[1] + sv[2] + sv[3] + sv[4] + sv[5] + sv[6] + sv[7]) / 8;
                    // lambert_vector / lambert_scattering (brdf.cpp:20-35)
                    const f3 l = mul(rnd, signbit_scalar(dot(rnd, tnorm)));
                    const f3 next = normalize(add(mul(l, scatter), mul(specular, 1 - scatter)));
                    refl.position = {hit.x, hit.y, hit.z, 0};
                    refl.triangle = ci.index;
                    refl.keep_going = 1;
                    refl.receiver_visible = visible ? 1 : 0;
                    rpos = hit;
                    rdir = next;
                }
            }
            keep_going = alive;
            prev_tri = refl.triangle;
            if (rout && step < P.keep_steps) rout[size_t(step) * n + ri] = refl;

            // ---- stochastic kernel (stochastic/program.cpp:58-152) ----
            if (!alive) continue;
            const triangle_t tri = sc.triangles[hit_tri];
            const surface_t& sf = sc.surfaces[tri.surface];
            float outgoing[8];
            for (int b = 0; b < 8; ++b) outgoing[b] = volume[b] * (1 - sf.absorption[b]);
            const f3 last_position = path_pos;
            const f3 this_position = hit;
            const float last_distance = path_dist;
            const float this_distance = last_distance + length(sub(last_position, this_position));
            float last_volume[8];
            for (int b = 0; b < 8; ++b) last_volume[b] = volume[b];
            for (int b = 0; b < 8; ++b) volume[b] = outgoing[b];
            path_pos = this_position;
            path_dist = this_distance;

            // specular ("intersected") output :108-118, binned from specular_from_step on
            if (segment_sphere(last_position, this_position, receiver, P.receiver_radius)) {
                const float total = last_distance + length(sub(receiver, last_position));
                if (step >= P.specular_from_step) deposit(last_volume, last_position, total);
            }
            // diffuse rain :121-151
            if (visible) {
                const f3 to_receiver = sub(receiver, this_position);
                const float trd = length(to_receiver);
                const float total = this_distance + trd;
                const f3 tnorm = triangle_normal(sc, tri);
                const float cos_angle = std::fabs(dot(tnorm, normalize(to_receiver)));
                const float sin_y = P.receiver_radius / std::fmax(P.receiver_radius, trd);
                const float angle_correction = 1 - std::sqrt(1 - sin_y * sin_y);
                float out[8];
                for (int b = 0; b < 8; ++b) {
                    out[b] = ((angle_correction * 2) * cos_angle) * (outgoing[b] * sf.scattering[b]);
                }
                deposit(out, this_position, total);
            }
        }
    }
    if (dropped) *dropped = drop_total;
}

int rto_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

size_t rto_params_size() { return sizeof(trace_params); }

}  // extern "C"

#include "is_oracle.inc"
