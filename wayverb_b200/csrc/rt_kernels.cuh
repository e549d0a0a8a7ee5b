// rt_kernels.cuh -- sm_100a device code of the stochastic ray-reflection loop.
//
// The reference drives, per reflection step and per 16384-ray segment, one
// `reflections` launch (src/raytracer/src/program.cpp:59-153), one `stochastic`
// launch (src/raytracer/src/stochastic/program.cpp:58-152), three bulk H2D and
// three bulk D2H copies, host RNG and a host histogram loop
// (raytracer.h:223-244, reflector.cpp:31-51, stochastic/finder.h:48-79,
// reflection_processor/stochastic_histogram.h:70-111). Here one thread owns one
// ray for its whole life: closest hit by voxel DDA, receiver visibility, next
// direction, energy bookkeeping and histogram binning all happen in registers,
// and only the histogram (fp64 atomics) and the optional first-steps
// reflection records touch memory.
//
// fp32 arithmetic in the reference's operation order (-fmad=false). Random
// numbers and sin/cos are the fixed definitions shared with the oracle (see
// oracle/rt_oracle.cpp's header): Philox4x32-10 keyed by (seed), counter =
// (global ray index, step, stream); Cody-Waite + polynomial sincos.
#pragma once

#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>

#ifndef RT_MIN_BLOCKS
#define RT_MIN_BLOCKS 8
#endif

namespace wvb {
namespace rt {

struct f3 {
    float x, y, z;
};

// reference PODs as they arrive (scene_buffers.h:14-38)
struct alignas(16) TriPod {  // core::triangle
    uint32_t surface, v0, v1, v2;
};
struct alignas(16) ReflectionPod {  // raytracer::reflection, 32 B  (raytracer/cl/reflection.h:10-17)
    float px, py, pz, pw;
    uint32_t triangle;
    int8_t keep_going;
    int8_t receiver_visible;
    int8_t pad_[10];
};
static_assert(sizeof(ReflectionPod) == 32, "reflection layout");

// per-triangle data derived once at create time with the same fp32 operations
// the per-ray code would use: v0, e0 = v1 - v0, e1 = v2 - v0 and the unit
// normal (triangle_verts_normal, geometry.cpp:69-74).
struct alignas(16) TriPre {
    float v0x, v0y, v0z, nx;
    float e0x, e0y, e0z, ny;
    float e1x, e1y, e1z, nz;
};
// One entry of a voxel's triangle run, gathered at create time so that the walk
// needs one dependent load per triangle instead of index -> triangle -> vertices:
// the triangle's precomputed data next to its index (64 B, one 128-B line holds two).
struct alignas(16) VoxEntry {
    TriPre pre;
    uint32_t tri;
    uint32_t pad_[3];
};
static_assert(sizeof(VoxEntry) == 64, "VoxEntry size");

struct Scene {
    const uint2* cells;        // per voxel (x*side*side + y*side + z): first entry, entry count
    const VoxEntry* entries;   // the runs of voxel_collection.cpp:9-37, in the same order
    const uint32_t* voxel_index;
    const TriPod* triangles;
    const TriPre* pre;
    const float* surfaces;  // 16 floats per surface: absorption[8], scattering[8]
    f3 c0, c1;
    uint32_t side;
    uint32_t n_triangles;
};

struct Params {
    f3 source, receiver;
    float receiver_radius;
    float ray_energy;
    double speed_of_sound;
    double histogram_rate;
    unsigned long long seed;
    unsigned long long ray_index_base;
    uint32_t depth;
    uint32_t specular_from_step;
    uint32_t n_bins;
    uint32_t directional;
    uint32_t keep_steps;
};

__device__ __forceinline__ f3 mk(float x, float y, float z) { return {x, y, z}; }
__device__ __forceinline__ f3 add(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ f3 sub(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ f3 mul(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ float dot(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ f3 cross(f3 a, f3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ float length(f3 a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ f3 normalize(f3 a) { return mul(a, 1.0f / sqrtf(dot(a, a))); }

// ---- Philox4x32-10 ---------------------------------------------------------------
__device__ __forceinline__ void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                       uint32_t k0, uint32_t k1, uint32_t& o0, uint32_t& o1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0;
        const uint32_t n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    o0 = c0;
    o1 = c1;
}
__device__ __forceinline__ void direction_rng(unsigned long long seed, uint32_t ray, uint32_t step,
                                              uint32_t stream, float& z, float& theta) {
    uint32_t o0, o1;
    philox(ray, step, stream, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), o0, o1);
    const float u0 = (float)(o0 >> 8) * 5.9604644775390625e-08f;
    const float u1 = (float)(o1 >> 8) * 5.9604644775390625e-08f;
    z = 2.0f * u0 - 1.0f;
    theta = (2.0f * u1 - 1.0f) * 3.14159274101257324f;
}

__device__ __forceinline__ void sincos_fixed(float theta, float& s, float& c) {
    const float kf = rintf(theta * 0.636619746685028076f);
    const int k = (int)kf;
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
        case 0: s = sin_r; c = cos_r; break;
        case 1: s = cos_r; c = -sin_r; break;
        case 2: s = -sin_r; c = -cos_r; break;
        default: s = -cos_r; c = sin_r; break;
    }
}

// sphere_point                  brdf.cpp:7-11
__device__ __forceinline__ f3 sphere_point(float z, float theta) {
    const float t = sqrtf(1 - z * z);
    float s, c;
    sincos_fixed(theta, s, c);
    return {t * c, z, t * s};
}

// almost_equal                  geometry.cpp:7-11
__device__ __forceinline__ bool almost_equal(float x, float y, float ulp) {
    const float abs_diff = fabsf(x - y);
    return abs_diff < FLT_EPSILON * fabsf(x + y) * ulp || abs_diff < FLT_MIN;
}

// triangle_vert_intersection    geometry.cpp:20-54   (returns t, 0 = no hit)
__device__ __forceinline__ float tri_intersection(const TriPre& T, f3 pos, f3 dir) {
    const f3 e0 = mk(T.e0x, T.e0y, T.e0z);
    const f3 e1 = mk(T.e1x, T.e1y, T.e1z);
    const f3 pvec = cross(dir, e1);
    const float det = dot(e0, pvec);
    if (almost_equal(det, 0, 10.0f)) return 0.0f;
    const float invdet = 1.0f / det;
    const f3 tvec = sub(pos, mk(T.v0x, T.v0y, T.v0z));
    const float u = invdet * dot(tvec, pvec);
    if (u < 0.0f || 1.0f < u) return 0.0f;
    const f3 qvec = cross(tvec, e0);
    const float v = invdet * dot(dir, qvec);
    if (v < 0.0f || 1.0f < v + u) return 0.0f;
    const float t = invdet * dot(e1, qvec);
    if (t < 0 || almost_equal(t, 0, 10.0f)) return 0.0f;
    return t;
}

// VOXEL_TRAVERSAL_ALGORITHM + voxel_traversal + ray_triangle_group_intersection
// (voxel.cpp:22-95, geometry.cpp:103-148). Returns t (0 = none) and the index.
// t_stop: the walk ends (returning "no hit") before visiting a voxel whose entry
// distance already exceeds t_stop. With INFINITY this is the reference's walk.
// The visibility query passes the distance to the receiver: a hit accepted in a
// voxel entered at e > t_stop has t >= e > t_stop (voxel lists are conservative
// supersets of the triangles touching the voxel, so a nearer hit would have been
// accepted in the voxel that contains it), and `!t || mag < t` is true either way.
__device__ __forceinline__ float voxel_traversal(const Scene& sc, f3 pos, f3 dir, uint32_t avoid,
                                                 uint32_t& index, float t_stop = INFINITY) {
    index = 0;
    // lanes that entered together leave together: the loop below is paced by a
    // warp vote, so every iteration all unfinished lanes do one unit of work side by
    // side (see the note at the loop)
    const unsigned lanes = __activemask();
    bool done = false;
    float result = 0.0f;
    const float sidef = (float)sc.side;
    const f3 vd = mk((sc.c1.x - sc.c0.x) / sidef, (sc.c1.y - sc.c0.y) / sidef,
                     (sc.c1.z - sc.c0.z) / sidef);
    const f3 rel = mk((pos.x - sc.c0.x) / vd.x, (pos.y - sc.c0.y) / vd.y, (pos.z - sc.c0.z) / vd.z);
    int ix = (int)floorf(rel.x), iy = (int)floorf(rel.y), iz = (int)floorf(rel.z);
    const int side = (int)sc.side;
    if (!(0 <= ix && 0 <= iy && 0 <= iz && ix < side && iy < side && iz < side)) {
        done = true;
        ix = iy = iz = 0;
    }
    const f3 lo = mk(sc.c0.x + (float)ix * vd.x, sc.c0.y + (float)iy * vd.y, sc.c0.z + (float)iz * vd.z);
    const f3 hi = mk(sc.c0.x + (float)(ix + 1) * vd.x, sc.c0.y + (float)(iy + 1) * vd.y,
                     sc.c0.z + (float)(iz + 1) * vd.z);
    const bool ngx = signbit(dir.x), ngy = signbit(dir.y), ngz = signbit(dir.z);
    const int stx = ngx ? -1 : 1, sty = ngy ? -1 : 1, stz = ngz ? -1 : 1;
    const int jox = ngx ? -1 : side, joy = ngy ? -1 : side, joz = ngz ? -1 : side;
    float tmx = fabsf(((ngx ? lo.x : hi.x) - pos.x) / dir.x);
    float tmy = fabsf(((ngy ? lo.y : hi.y) - pos.y) / dir.y);
    float tmz = fabsf(((ngz ? lo.z : hi.z) - pos.z) / dir.z);
    if (isnan(tmx)) tmx = INFINITY;
    if (isnan(tmy)) tmy = INFINITY;
    if (isnan(tmz)) tmz = INFINITY;
    const float tdx = fabsf(vd.x / dir.x), tdy = fabsf(vd.y / dir.y), tdz = fabsf(vd.z / dir.z);
    // The reference nests "for each voxel { for each triangle }". On a SIMT machine
    // that nest leaves most lanes idle (ncu: 3.8 of 32 threads active per issued
    // instruction) because lanes reach their few non-empty voxels at different
    // times. The same walk is therefore run as a flat state machine: every
    // iteration a lane either enters a voxel, tests ONE triangle of the current
    // voxel, or leaves the voxel -- identical visiting and testing order, identical
    // arithmetic, identical result. The vote in the loop condition is a convergence
    // point: without it the compiler threads the branches back into the nested
    // loops.
    uint32_t i = 0, num = 0;
    const VoxEntry* begin = sc.entries;
    float best_t = 0.0f, tmin = 0.0f;
    uint32_t best_i = 0;
    int min_i = 0;
    bool enter = true;
    while (__any_sync(lanes, !done)) {
        if (!done) {
            if (enter) {
                min_i = 0;
                tmin = tmx;
                if (tmy < tmin) { min_i = 1; tmin = tmy; }
                if (tmz < tmin) { min_i = 2; tmin = tmz; }
                const uint2 cell = sc.cells[(size_t)ix * side * side + (size_t)iy * side + iz];
                num = cell.y;
                begin = sc.entries + cell.x;
                i = 0;
                best_t = 0.0f;
                enter = false;
            }
            if (i < num) {
                const uint32_t ti = begin[i].tri;
                const TriPre& T = begin[i].pre;
                ++i;
                if (ti != avoid) {
                    const float t = tri_intersection(T, pos, dir);
                    if (t && (!best_t || t < best_t)) {
                        best_i = ti;
                        best_t = t;
                    }
                }
            }
            if (i >= num) {
                if (best_t && best_t <= tmin) {
                    index = best_i;
                    result = best_t;
                    done = true;
                } else if (tmin > t_stop) {
                    done = true;  // the next voxel starts beyond the point of interest
                } else {
                    if (min_i == 0) {
                        ix += stx;
                        if (ix == jox) done = true;
                        tmx += tdx;
                    } else if (min_i == 1) {
                        iy += sty;
                        if (iy == joy) done = true;
                        tmy += tdy;
                    } else {
                        iz += stz;
                        if (iz == joz) done = true;
                        tmz += tdz;
                    }
                    enter = true;
                }
            }
        }
    }
    return result;
}

// voxel_point_intersection      voxel.cpp:227-258
__device__ __forceinline__ bool point_visible(const Scene& sc, f3 begin, f3 point, uint32_t avoid) {
    const f3 b2p = sub(point, begin);
    const float mag = length(b2p);
    const f3 direction = normalize(b2p);
    uint32_t idx;
    const float t = voxel_traversal(sc, begin, direction, avoid, idx, mag);
    return !t || mag < t;
}

// line_segment_sphere_intersection   geometry.cpp:155-164
__device__ __forceinline__ bool segment_sphere(f3 p1, f3 p2, f3 c, float r) {
    const f3 diff = sub(p2, p1);
    const float u = dot(sub(c, p1), diff) / dot(diff, diff);
    if (u < 0 || 1 < u) return false;
    const f3 closest = sub(add(p1, mul(diff, u)), c);
    return dot(closest, closest) < r * r;
}

__device__ __forceinline__ float signbit_scalar(float x) { return signbit(x) ? 1.0f : 0.0f; }

// vector_look_up_table<..,20,9>::index (vector_look_up_table.h:53-116, az_el.cpp:53-68)
__device__ __forceinline__ void lut_index(f3 v, int& az_cell, int& el_cell) {
    float az = atan2f(v.x, -v.z);
    const float el = asinf(v.y);
    if (almost_equal(el, -1.57079637050628662f, 10.0f) || almost_equal(el, 1.57079637050628662f, 10.0f))
        az = 0;
    const double deg = 180.0 / 3.14159265358979323846;
    double a = (double)(-az) * deg;
    a += (360.0 / 20) / 2;
    while (a < 0) a += 360;
    az_cell = (int)((unsigned long long)(a / (360.0 / 20)) % 20ull);
    double e = (double)el * deg;
    e += 90 + (180.0 / 10) / 2;
    while (e < 0) e += 360;
    unsigned long long adj = (unsigned long long)(e / (180.0 / 10)) % 20ull;
    if (adj < 1) adj = 1;
    if (adj > 9) adj = 9;
    el_cell = (int)(adj - 1);
}

// one impulse into the histogram (finder.h:65-76 drops distance == 0;
// energy_histogram_sum, stochastic_histogram.h:17-32)
__device__ __forceinline__ void deposit(const Params& P, double* __restrict__ hist,
                                        unsigned long long* __restrict__ dropped,
                                        const float (&vol)[8], f3 position, float distance) {
    if (!distance) return;
    const double time = (double)distance / P.speed_of_sound;
    const size_t bin = (size_t)(time * P.histogram_rate);
    if (bin >= P.n_bins) {
        atomicAdd(dropped, 1ull);
        return;
    }
    size_t base = bin * 8;
    if (P.directional) {
        int az, el;
        lut_index(normalize(sub(position, P.receiver)), az, el);
        base = (((size_t)az * 9 + el) * P.n_bins + bin) * 8;
    }
#pragma unroll
    for (int b = 0; b < 8; ++b) atomicAdd(hist + base + b, (double)vol[b]);
}

// per-triangle precompute (same ops as triangle_normal / triangle_vert_intersection)
static __global__ void rt_precompute(const TriPod* __restrict__ tris, const float4* __restrict__ verts,
                              TriPre* __restrict__ out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const TriPod t = tris[i];
        const float4 a = verts[t.v0], b = verts[t.v1], c = verts[t.v2];
        const f3 v0 = mk(a.x, a.y, a.z);
        const f3 e0 = sub(mk(b.x, b.y, b.z), v0);
        const f3 e1 = sub(mk(c.x, c.y, c.z), v0);
        const f3 nrm = normalize(cross(e0, e1));
        out[i] = TriPre{v0.x, v0.y, v0.z, nrm.x, e0.x, e0.y, e0.z, nrm.y, e1.x, e1.y, e1.z, nrm.z};
    }
}

// voxel runs -> (cells, entries)
static __global__ void rt_build_entries(const uint32_t* __restrict__ voxel_index, const TriPre* __restrict__ pre,
                                 const uint32_t* __restrict__ first, uint2* __restrict__ cells,
                                 VoxEntry* __restrict__ entries, uint32_t n_cells) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_cells) {
        const uint32_t off = voxel_index[c];
        const uint32_t num = voxel_index[off];
        const uint32_t e0 = first[c];
        cells[c] = make_uint2(e0, num);
        for (uint32_t i = 0; i < num; ++i) {
            const uint32_t ti = voxel_index[off + 1 + i];
            VoxEntry e;
            e.pre = pre[ti];
            e.tri = ti;
            e.pad_[0] = e.pad_[1] = e.pad_[2] = 0;
            entries[e0 + i] = e;
        }
    }
}

// directions from Philox stream 1 (random_unit_vector, core/azimuth_elevation.h:31-35)
static __global__ void rt_directions(unsigned long long seed, unsigned long long base, uint32_t n,
                              float* __restrict__ out3) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        float z, th;
        direction_rng(seed, (uint32_t)(base + i), 0u, 1u, z, th);
        const f3 d = sphere_point(z, th);
        out3[3 * (size_t)i] = d.x;
        out3[3 * (size_t)i + 1] = d.y;
        out3[3 * (size_t)i + 2] = d.z;
    }
}

// closest hit of n rays (test hook mirroring reflector_tests.cpp's comparison)
static __global__ void rt_closest_hit(Scene sc, const float* __restrict__ rays6, uint32_t n,
                               uint32_t* __restrict__ tri_out, float* __restrict__ t_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const f3 p = mk(rays6[6 * (size_t)i], rays6[6 * (size_t)i + 1], rays6[6 * (size_t)i + 2]);
        const f3 d = mk(rays6[6 * (size_t)i + 3], rays6[6 * (size_t)i + 4], rays6[6 * (size_t)i + 5]);
        uint32_t idx;
        const float t = voxel_traversal(sc, p, d, ~0u, idx);
        tri_out[i] = t ? idx : ~0u;
        t_out[i] = t;
    }
}

// the whole life of one ray: raytracer.h:223-244 with reflections (program.cpp:59-153),
// stochastic (stochastic/program.cpp:58-152) and the histogram processor folded in
static __global__ void __launch_bounds__(128, RT_MIN_BLOCKS)
rt_trace(Scene sc, Params P, const float* __restrict__ dirs3, uint32_t n, double* __restrict__ hist,
         unsigned long long* __restrict__ dropped, ReflectionPod* __restrict__ refl_out) {
    const uint32_t ri = blockIdx.x * blockDim.x + threadIdx.x;
    if (ri >= n) return;
    f3 rpos = P.source;
    f3 rdir = mk(dirs3[3 * (size_t)ri], dirs3[3 * (size_t)ri + 1], dirs3[3 * (size_t)ri + 2]);
    bool keep_going = true;
    uint32_t prev_tri = ~0u;
    float volume[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) volume[b] = P.ray_energy;
    f3 path_pos = P.source;
    float path_dist = 0;

    for (uint32_t step = 0; step < P.depth; ++step) {
        ReflectionPod refl = {};
        f3 hit = mk(0, 0, 0);
        uint32_t hit_tri = 0;
        bool visible = false, alive = false;
        f3 tnorm_raw = mk(0, 0, 0);
        if (keep_going) {
            uint32_t idx;
            const float t = voxel_traversal(sc, rpos, rdir, prev_tri, idx);
            if (t) {
                alive = true;
                hit = add(rpos, mul(rdir, t));
                hit_tri = idx;
                const TriPre T = sc.pre[idx];
                tnorm_raw = mk(T.nx, T.ny, T.nz);
                // reflect (geometry.cpp:83-86)
                const f3 specular = sub(rdir, mul(mul(tnorm_raw, 2), dot(rdir, tnorm_raw)));
                const f3 tnorm = mul(tnorm_raw, signbit_scalar(dot(tnorm_raw, specular)));
                visible = point_visible(sc, hit, P.receiver, idx);
                float z, theta;
                direction_rng(P.seed, (uint32_t)(P.ray_index_base + ri), step, 0u, z, theta);
                const f3 rnd = sphere_point(z, theta);
                const float* sv = sc.surfaces + 16 * (size_t)sc.triangles[idx].surface + 8;
                const float scatter =
                        (sv[0] + sv[1] + sv[2] + sv[3] + sv[4] + sv[5] + sv[6] + sv[7]) / 8;
                const f3 l = mul(rnd, signbit_scalar(dot(rnd, tnorm)));
                const f3 next = normalize(add(mul(l, scatter), mul(specular, 1 - scatter)));
                refl.px = hit.x; refl.py = hit.y; refl.pz = hit.z;
                refl.triangle = idx;
                refl.keep_going = 1;
                refl.receiver_visible = visible ? 1 : 0;
                rpos = hit;
                rdir = next;
            }
        }
        keep_going = alive;
        prev_tri = refl.triangle;
        if (refl_out && step < P.keep_steps) refl_out[(size_t)step * n + ri] = refl;
        if (!alive) continue;

        const float* sf = sc.surfaces + 16 * (size_t)sc.triangles[hit_tri].surface;
        float outgoing[8], last_volume[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            last_volume[b] = volume[b];
            outgoing[b] = volume[b] * (1 - sf[b]);
            volume[b] = outgoing[b];
        }
        const f3 last_position = path_pos;
        const float last_distance = path_dist;
        const float this_distance = last_distance + length(sub(last_position, hit));
        path_pos = hit;
        path_dist = this_distance;

        if (segment_sphere(last_position, hit, P.receiver, P.receiver_radius)) {
            const float total = last_distance + length(sub(P.receiver, last_position));
            if (step >= P.specular_from_step) deposit(P, hist, dropped, last_volume, last_position, total);
        }
        if (visible) {
            const f3 to_receiver = sub(P.receiver, hit);
            const float trd = length(to_receiver);
            const float total = this_distance + trd;
            const float cos_angle = fabsf(dot(tnorm_raw, normalize(to_receiver)));
            const float sin_y = P.receiver_radius / fmaxf(P.receiver_radius, trd);
            const float angle_correction = 1 - sqrtf(1 - sin_y * sin_y);
            float out[8];
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                out[b] = ((angle_correction * 2) * cos_angle) * (outgoing[b] * sf[8 + b]);
            }
            deposit(P, hist, dropped, out, hit, total);
        }
    }
}

// ---------------------------------------------------------------------------
// Wavefront form of the same loop: one launch per reflection, rays re-binned in between.
//
// rt_trace keeps 11 of 32 lanes busy on the concert hall (ncu): the lanes of a warp own
// unrelated rays, so at any moment some step through empty voxels, some test triangles and some
// are done. Here every ray's state lives in device memory; between two reflections the rays are
// counting-sorted by (voxel of their origin, cell of their direction on an octahedral map), so that a warp's lanes start in the same voxel heading the same way and walk the same
// cells. The receiver-visibility ray of a hit is cast at the START of the next pass, when the
// rays are sorted by the voxel of that hit (all lanes then aim at the receiver from one voxel).
// Per ray the arithmetic, its order and the random numbers (keyed by global ray index and step)
// are those of rt_trace, so reflections and impulses are identical; only the order in which the
// impulses reach the histogram's atomics differs (as it already does from run to run).
// ---------------------------------------------------------------------------
struct WaveState {
    float4* pos;      // x, y, z of the ray origin (= last hit = path position), w = path distance
    float4* dir;      // x, y, z, w = bits of the triangle just left (~0: none)
    float4* vol;      // [2][n]: energy per band
    uint32_t* alive;  // 1 while the ray is alive
    uint32_t* keys;   // sort key of the next pass, per ray
    uint32_t* perm;   // rays in the order of this pass
    uint32_t* bins;   // counting-sort histogram / offsets, n_bins + 1 entries
    uint32_t key_voxel_shift;  // voxel coordinates are coarsened by this many bits in the key
    uint32_t key_side_bits;    // bits per (coarsened) voxel coordinate
    uint32_t key_dir_bits;     // bits per axis of the octahedral direction map
    uint32_t dead_key;         // key of a dead ray: the last bin
};

__device__ __forceinline__ uint32_t wave_key(const Scene& sc, const WaveState& W, f3 pos, f3 dir) {
    const float sidef = (float)sc.side;
    const int side = (int)sc.side;
    int ix = (int)floorf((pos.x - sc.c0.x) / ((sc.c1.x - sc.c0.x) / sidef));
    int iy = (int)floorf((pos.y - sc.c0.y) / ((sc.c1.y - sc.c0.y) / sidef));
    int iz = (int)floorf((pos.z - sc.c0.z) / ((sc.c1.z - sc.c0.z) / sidef));
    ix = min(max(ix, 0), side - 1) >> W.key_voxel_shift;
    iy = min(max(iy, 0), side - 1) >> W.key_voxel_shift;
    iz = min(max(iz, 0), side - 1) >> W.key_voxel_shift;
    // octahedral map of the direction onto [-1, 1]^2, 8 x 8 cells
    const float inv = 1.0f / (fabsf(dir.x) + fabsf(dir.y) + fabsf(dir.z));
    float u = dir.x * inv, v = dir.y * inv;
    if (dir.z < 0) {
        const float uu = (1.0f - fabsf(v)) * (u < 0 ? -1.0f : 1.0f);
        const float vv = (1.0f - fabsf(u)) * (v < 0 ? -1.0f : 1.0f);
        u = uu;
        v = vv;
    }
    const int cells = 1 << W.key_dir_bits;
    const int cu = min(max((int)((u * 0.5f + 0.5f) * (float)cells), 0), cells - 1);
    const int cv = min(max((int)((v * 0.5f + 0.5f) * (float)cells), 0), cells - 1);
    const uint32_t voxel = (((uint32_t)ix << W.key_side_bits) | (uint32_t)iy) << W.key_side_bits | (uint32_t)iz;
    return (voxel << (2 * W.key_dir_bits)) | (uint32_t)(cu * cells + cv);
}

// pass 0 set-up: every ray at the source with its direction
static __global__ void rt_wave_init(Scene sc, Params P, WaveState W, const float* __restrict__ dirs3, uint32_t n) {
    const uint32_t ri = blockIdx.x * blockDim.x + threadIdx.x;
    if (ri >= n) return;
    const f3 d = mk(dirs3[3 * (size_t)ri], dirs3[3 * (size_t)ri + 1], dirs3[3 * (size_t)ri + 2]);
    W.pos[ri] = make_float4(P.source.x, P.source.y, P.source.z, 0.0f);
    W.dir[ri] = make_float4(d.x, d.y, d.z, __uint_as_float(~0u));
    W.vol[ri] = make_float4(P.ray_energy, P.ray_energy, P.ray_energy, P.ray_energy);
    W.vol[(size_t)n + ri] = make_float4(P.ray_energy, P.ray_energy, P.ray_energy, P.ray_energy);
    W.alive[ri] = 1u;
    const uint32_t key = wave_key(sc, W, P.source, d);
    W.keys[ri] = key;
    atomicAdd(W.bins + key, 1u);
}

// exclusive scan of the bin counts, in place: 1024 threads x 4 bins per block
static __global__ void __launch_bounds__(1024) rt_scan_blocks(uint32_t* __restrict__ bins, uint32_t n_bins,
                                                              uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t warp_sums[32];
    const uint32_t base = blockIdx.x * 4096u + threadIdx.x * 4u;
    uint32_t v[4], s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[k] = base + k < n_bins ? bins[base + k] : 0u;
        s += v[k];
    }
    uint32_t incl = s;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= (unsigned)o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t w = warp_sums[threadIdx.x], wi = w;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (threadIdx.x >= (unsigned)o) wi += t;
        }
        warp_sums[threadIdx.x] = wi - w;  // exclusive prefix of the warps
        if (threadIdx.x == 31) block_sums[blockIdx.x] = wi;
    }
    __syncthreads();
    uint32_t run = warp_sums[threadIdx.x >> 5] + incl - s;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (base + k < n_bins) bins[base + k] = run;
        run += v[k];
    }
}
static __global__ void __launch_bounds__(1024) rt_scan_sums(uint32_t* __restrict__ block_sums, uint32_t n) {
    // n <= 1024 * 8: one block, serial per thread + block scan
    __shared__ uint32_t tot[1024];
    const uint32_t per = (n + 1023u) / 1024u;
    uint32_t s = 0;
    for (uint32_t k = 0; k < per; ++k) {
        const uint32_t i = threadIdx.x * per + k;
        if (i < n) s += block_sums[i];
    }
    tot[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const uint32_t t = threadIdx.x >= (unsigned)o ? tot[threadIdx.x - o] : 0u;
        __syncthreads();
        tot[threadIdx.x] += t;
        __syncthreads();
    }
    uint32_t run = tot[threadIdx.x] - s;
    for (uint32_t k = 0; k < per; ++k) {
        const uint32_t i = threadIdx.x * per + k;
        if (i < n) {
            const uint32_t c = block_sums[i];
            block_sums[i] = run;
            run += c;
        }
    }
}
static __global__ void __launch_bounds__(1024) rt_scan_add(uint32_t* __restrict__ bins, uint32_t n_bins,
                                                           const uint32_t* __restrict__ block_sums) {
    const uint32_t base = blockIdx.x * 4096u + threadIdx.x * 4u;
    const uint32_t add = block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (base + k < n_bins) bins[base + k] += add;
    }
}
static __global__ void rt_scatter(const uint32_t* __restrict__ keys, uint32_t* __restrict__ bins,
                                  uint32_t* __restrict__ perm, uint32_t n) {
    const uint32_t ri = blockIdx.x * blockDim.x + threadIdx.x;
    if (ri < n) perm[atomicAdd(bins + keys[ri], 1u)] = ri;
}

// one reflection of every ray, in sorted order. step == P.depth: only the deferred visibility
// of the last hit.
static __global__ void __launch_bounds__(128, RT_MIN_BLOCKS)
rt_wave(Scene sc, Params P, WaveState W, uint32_t n, uint32_t step, double* __restrict__ hist,
        unsigned long long* __restrict__ dropped, ReflectionPod* __restrict__ refl_out, uint32_t* __restrict__ next_bins) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t ri = W.perm[i];
    if (!W.alive[ri]) {
        if (step < P.depth) {
            if (refl_out && step < P.keep_steps) refl_out[(size_t)step * n + ri] = ReflectionPod{};
            W.keys[ri] = W.dead_key;
            atomicAdd(next_bins + W.dead_key, 1u);
        }
        return;
    }
    const float4 p4 = W.pos[ri], d4 = W.dir[ri];
    f3 rpos = mk(p4.x, p4.y, p4.z);
    f3 rdir = mk(d4.x, d4.y, d4.z);
    float path_dist = p4.w;
    uint32_t prev_tri = __float_as_uint(d4.w);
    float volume[8];
    {
        const float4 a = W.vol[ri], b = W.vol[(size_t)n + ri];
        volume[0] = a.x; volume[1] = a.y; volume[2] = a.z; volume[3] = a.w;
        volume[4] = b.x; volume[5] = b.y; volume[6] = b.z; volume[7] = b.w;
    }
    // ---- the previous reflection's visibility ray and diffuse contribution ----------------------
    if (step > 0) {
        const bool visible = point_visible(sc, rpos, P.receiver, prev_tri);
        if (refl_out && step - 1 < P.keep_steps) refl_out[(size_t)(step - 1) * n + ri].receiver_visible = visible ? 1 : 0;
        if (visible) {
            const TriPre T = sc.pre[prev_tri];
            const f3 tnorm_raw = mk(T.nx, T.ny, T.nz);
            const float* sf = sc.surfaces + 16 * (size_t)sc.triangles[prev_tri].surface;
            const f3 to_receiver = sub(P.receiver, rpos);
            const float trd = length(to_receiver);
            const float total = path_dist + trd;
            const float cos_angle = fabsf(dot(tnorm_raw, normalize(to_receiver)));
            const float sin_y = P.receiver_radius / fmaxf(P.receiver_radius, trd);
            const float angle_correction = 1 - sqrtf(1 - sin_y * sin_y);
            float out[8];
#pragma unroll
            for (int b = 0; b < 8; ++b) out[b] = ((angle_correction * 2) * cos_angle) * (volume[b] * sf[8 + b]);
            deposit(P, hist, dropped, out, rpos, total);
        }
    }
    if (step >= P.depth) return;
    // ---- this reflection ---------------------------------------------------------------------------
    ReflectionPod refl = {};
    uint32_t idx;
    const float t = voxel_traversal(sc, rpos, rdir, prev_tri, idx);
    uint32_t key = W.dead_key;
    if (t) {
        const f3 hit = add(rpos, mul(rdir, t));
        const TriPre T = sc.pre[idx];
        const f3 tnorm_raw = mk(T.nx, T.ny, T.nz);
        const f3 specular = sub(rdir, mul(mul(tnorm_raw, 2), dot(rdir, tnorm_raw)));
        const f3 tnorm = mul(tnorm_raw, signbit_scalar(dot(tnorm_raw, specular)));
        float z, theta;
        direction_rng(P.seed, (uint32_t)(P.ray_index_base + ri), step, 0u, z, theta);
        const f3 rnd = sphere_point(z, theta);
        const float* sf = sc.surfaces + 16 * (size_t)sc.triangles[idx].surface;
        const float* sv = sf + 8;
        const float scatter = (sv[0] + sv[1] + sv[2] + sv[3] + sv[4] + sv[5] + sv[6] + sv[7]) / 8;
        const f3 l = mul(rnd, signbit_scalar(dot(rnd, tnorm)));
        const f3 next = normalize(add(mul(l, scatter), mul(specular, 1 - scatter)));
        refl.px = hit.x; refl.py = hit.y; refl.pz = hit.z;
        refl.triangle = idx;
        refl.keep_going = 1;
        float last_volume[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            last_volume[b] = volume[b];
            volume[b] = volume[b] * (1 - sf[b]);
        }
        const f3 last_position = rpos;
        const float last_distance = path_dist;
        const float this_distance = last_distance + length(sub(last_position, hit));
        if (segment_sphere(last_position, hit, P.receiver, P.receiver_radius)) {
            const float total = last_distance + length(sub(P.receiver, last_position));
            if (step >= P.specular_from_step) deposit(P, hist, dropped, last_volume, last_position, total);
        }
        W.pos[ri] = make_float4(hit.x, hit.y, hit.z, this_distance);
        W.dir[ri] = make_float4(next.x, next.y, next.z, __uint_as_float(idx));
        W.vol[ri] = make_float4(volume[0], volume[1], volume[2], volume[3]);
        W.vol[(size_t)n + ri] = make_float4(volume[4], volume[5], volume[6], volume[7]);
        key = wave_key(sc, W, hit, next);
    } else {
        W.alive[ri] = 0u;
    }
    if (refl_out && step < P.keep_steps) refl_out[(size_t)step * n + ri] = refl;
    W.keys[ri] = key;
    atomicAdd(next_bins + key, 1u);
}

}  // namespace rt
}  // namespace wvb
