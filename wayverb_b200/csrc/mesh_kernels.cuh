// mesh_kernels.cuh -- sm_100a device code of the waveguide mesh construction
// ("next" rank 1: the step immediately before the hot path).
//
//   set_node_inside            src/waveguide/src/mesh_setup_program.cpp:110-140
//     voxel_inside / single_ray_inside / count_intersections   src/core/src/cl/voxel.cpp:97-225
//   set_node_boundary_type     mesh_setup_program.cpp:14-108,142-172
//   boundary_coefficient_finder_1d/2d/3d   src/waveguide/src/boundary_coefficient_program.cpp:310-484
//     slow_closest_triangle :222-241, point_triangle_distance_squared :16-135
//
// fp32 in the reference's operation order (-fmad=false), bit-identical to the
// oracle's restatement. Numbering of boundary_index (boundary_coefficient_finder.cpp:
// 12-19,129) is host code in the reference and stays host code here.
#pragma once

#include "rt_kernels.cuh"
#include "../../include/wvb200.h"

namespace wvb {
namespace mesh {

using rt::f3;
using rt::mk;

struct Desc {
    f3 min_corner;
    int dx, dy, dz;
    float spacing;
};

__device__ __forceinline__ f3 node_position(const Desc& d, long long i, int& x, int& y, int& z) {
    x = (int)(i % d.dx);
    y = (int)((i / d.dx) % d.dy);
    z = (int)(i / d.dx / d.dy);
    // compute_node_position (cl/utils.cpp:71-74): min_corner + convert_float3(locator) * spacing
    return mk(d.min_corner.x + (float)x * d.spacing, d.min_corner.y + (float)y * d.spacing,
              d.min_corner.z + (float)z * d.spacing);
}

// triangle_vert_intersection returning u, v as well (geometry.cpp:20-54)
__device__ __forceinline__ float tri_intersection_uv(const rt::TriPre& T, f3 pos, f3 dir, float& u_out,
                                                     float& v_out) {
    const f3 e0 = mk(T.e0x, T.e0y, T.e0z);
    const f3 e1 = mk(T.e1x, T.e1y, T.e1z);
    const f3 pvec = rt::cross(dir, e1);
    const float det = rt::dot(e0, pvec);
    if (rt::almost_equal(det, 0, 10.0f)) return 0.0f;
    const float invdet = 1.0f / det;
    const f3 tvec = rt::sub(pos, mk(T.v0x, T.v0y, T.v0z));
    const float u = invdet * rt::dot(tvec, pvec);
    if (u < 0.0f || 1.0f < u) return 0.0f;
    const f3 qvec = rt::cross(tvec, e0);
    const float v = invdet * rt::dot(dir, qvec);
    if (v < 0.0f || 1.0f < v + u) return 0.0f;
    const float t = invdet * rt::dot(e1, qvec);
    if (t < 0 || rt::almost_equal(t, 0, 10.0f)) return 0.0f;
    u_out = u;
    v_out = v;
    return t;
}

// count_intersections (voxel.cpp:97-133) as the same flat state machine as the ray
// kernel's walk: ~0u when a crossing is degenerate.
__device__ __forceinline__ uint32_t count_intersections(const rt::Scene& sc, f3 pos, f3 dir) {
    const unsigned lanes = __activemask();
    bool done = false;
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
    uint32_t count = 0, i = 0, num = 0;
    const rt::VoxEntry* begin = sc.entries;
    float prev_max = 0.0f, tmin = 0.0f;
    int min_i = 0;
    bool enter = true;
    while (__any_sync(lanes, !done)) {  // vote = convergence point, see rt::voxel_traversal
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
                enter = false;
            }
            if (i < num) {
                float u, v;
                const float t = tri_intersection_uv(begin[i].pre, pos, dir, u, v);
                ++i;
                if (t) {
                    if (rt::almost_equal(u, 0, 10.0f) || rt::almost_equal(v, 0, 10.0f) ||
                        rt::almost_equal(u + v, 1, 10.0f)) {
                        count = ~0u;  // degenerate crossing: this probe direction is undecided
                        done = true;
                    } else if (prev_max < t && t <= tmin) {
                        count += 1;
                    }
                }
            }
            if (!done && i >= num) {
                if (min_i == 0) {
                    ix += stx;
                    if (ix == jox) done = true;
                    prev_max = tmx;
                    tmx += tdx;
                } else if (min_i == 1) {
                    iy += sty;
                    if (iy == joy) done = true;
                    prev_max = tmy;
                    tmy += tdy;
                } else {
                    iz += stz;
                    if (iz == joz) done = true;
                    prev_max = tmz;
                    tmz += tdz;
                }
                enter = true;
            }
        }
    }
    return count;
}

__constant__ float kInsideDirections[32][3] = {
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

// set_node_inside: voxel_inside for every node (voxel.cpp:191-225)
static __global__ void __launch_bounds__(128)
mesh_inside(rt::Scene sc, Desc d, uint8_t* __restrict__ inside) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long nn = (long long)d.dx * d.dy * d.dz;
    if (i >= nn) return;
    int x, y, z;
    const f3 pt = node_position(d, i, x, y, z);
    uint8_t r = 0;
    for (int k = 0; k != 32; ++k) {
        const f3 dir = mk(kInsideDirections[k][0], kInsideDirections[k][1], kInsideDirections[k][2]);
        const uint32_t n = count_intersections(sc, pt, dir);
        if (n == ~0u) continue;
        r = (uint8_t)(n % 2);
        break;
    }
    inside[i] = r;
}

// set_node_boundary_type + test_directions (mesh_setup_program.cpp:14-108,142-172)
__device__ __forceinline__ int test_directions(const uint8_t* __restrict__ inside, const Desc& d, int x,
                                               int y, int z, const int* dirs, int n) {
    int ret = WVB_ID_NONE;
    for (int k = 0; k < n; ++k) {
        const int a = dirs[k];
        const int ax = x + ((a & WVB_ID_PX) ? 1 : 0) - ((a & WVB_ID_NX) ? 1 : 0);
        const int ay = y + ((a & WVB_ID_PY) ? 1 : 0) - ((a & WVB_ID_NY) ? 1 : 0);
        const int az = z + ((a & WVB_ID_PZ) ? 1 : 0) - ((a & WVB_ID_NZ) ? 1 : 0);
        if (ax < 0 || ay < 0 || az < 0 || ax >= d.dx || ay >= d.dy || az >= d.dz) continue;
        if (inside[((long long)az * d.dy + ay) * d.dx + ax]) {
            if (ret != WVB_ID_NONE) return WVB_ID_REENTRANT;
            ret = a;
        }
    }
    return ret;
}

static __global__ void mesh_boundary_type(const uint8_t* __restrict__ inside, Desc d,
                                   wvb_condensed_node* __restrict__ nodes) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long nn = (long long)d.dx * d.dy * d.dz;
    if (i >= nn) return;
    wvb_condensed_node n{WVB_ID_NONE, 0};
    if (inside[i]) {
        n.boundary_type = WVB_ID_INSIDE;
    } else {
        const int x = (int)(i % d.dx), y = (int)((i / d.dx) % d.dy), z = (int)(i / d.dx / d.dy);
        const int d1[6] = {WVB_ID_NX, WVB_ID_PX, WVB_ID_NY, WVB_ID_PY, WVB_ID_NZ, WVB_ID_PZ};
        const int d2[12] = {WVB_ID_NX | WVB_ID_NY, WVB_ID_NX | WVB_ID_PY, WVB_ID_PX | WVB_ID_NY,
                            WVB_ID_PX | WVB_ID_PY, WVB_ID_NX | WVB_ID_NZ, WVB_ID_NX | WVB_ID_PZ,
                            WVB_ID_PX | WVB_ID_NZ, WVB_ID_PX | WVB_ID_PZ, WVB_ID_NY | WVB_ID_NZ,
                            WVB_ID_NY | WVB_ID_PZ, WVB_ID_PY | WVB_ID_NZ, WVB_ID_PY | WVB_ID_PZ};
        const int d3[8] = {WVB_ID_NX | WVB_ID_NY | WVB_ID_NZ, WVB_ID_NX | WVB_ID_NY | WVB_ID_PZ,
                           WVB_ID_NX | WVB_ID_PY | WVB_ID_NZ, WVB_ID_NX | WVB_ID_PY | WVB_ID_PZ,
                           WVB_ID_PX | WVB_ID_NY | WVB_ID_NZ, WVB_ID_PX | WVB_ID_NY | WVB_ID_PZ,
                           WVB_ID_PX | WVB_ID_PY | WVB_ID_NZ, WVB_ID_PX | WVB_ID_PY | WVB_ID_PZ};
        int t = test_directions(inside, d, x, y, z, d1, 6);
        if (t == WVB_ID_NONE) t = test_directions(inside, d, x, y, z, d2, 12);
        if (t == WVB_ID_NONE) t = test_directions(inside, d, x, y, z, d3, 8);
        n.boundary_type = t;
    }
    nodes[i] = n;
}

// point_triangle_distance_squared (boundary_coefficient_program.cpp:16-135), with e0/e1
// taken from the precomputed TriPre (the same v1 - v0, v2 - v0)
__device__ __forceinline__ float point_triangle_distance_squared(const rt::TriPre& T, f3 point) {
    const f3 v0 = mk(T.v0x, T.v0y, T.v0z);
    const f3 diff = rt::sub(point, v0);
    const f3 e0 = mk(T.e0x, T.e0y, T.e0z);
    const f3 e1 = mk(T.e1x, T.e1y, T.e1z);
    const float a00 = rt::dot(e0, e0);
    const float a01 = rt::dot(e0, e1);
    const float a11 = rt::dot(e1, e1);
    const float b0 = -rt::dot(diff, e0);
    const float b1 = -rt::dot(diff, e1);
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
    const f3 closest = rt::add(rt::add(v0, rt::mul(e0, t0)), rt::mul(e1, t1));
    const f3 dd = rt::sub(point, closest);
    return rt::dot(dd, dd);
}

// boundary_coefficient_finder_1d (boundary_coefficient_program.cpp:310-343) for the
// compacted list of 1-d / reentrant nodes: surface of the closest triangle.
//
// The reference's kernel calls slow_closest_triangle (:222-241): every triangle, first minimum
// wins, O(nodes x triangles). VOXEL = true gives the SAME triangle through the voxel grid (the
// search the reference sketches at :243-308, made exact): voxels overlapping the box
// [pt - r, pt + r] are scanned, r doubling from half a voxel; as soon as the best distance found
// is <= r the answer is final, because every triangle at that distance or closer touches the
// ball of that radius, which lies inside the scanned box, and voxel lists name every triangle
// that overlaps them (the precondition of wvb_rt_scene_desc). Ties go to the lowest triangle
// index, which is what "first minimum wins" yields over the whole array, and distances come
// from the same point_triangle_distance_squared, so the result is identical, not just close.
template <bool VOXEL>
static __global__ void mesh_find_1d(rt::Scene sc, Desc d, const uint32_t* __restrict__ node_list, uint32_t n,
                             uint32_t* __restrict__ surface_out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int x, y, z;
    const f3 pt = node_position(d, node_list[k], x, y, z);
    uint32_t best = 0;
    float distance = INFINITY;
    if (!VOXEL) {
        for (uint32_t i = 0; i != sc.n_triangles; ++i) {
            const float nd = point_triangle_distance_squared(sc.pre[i], pt);
            if (nd < distance) {
                best = i;
                distance = nd;
            }
        }
    } else {
        const int side = (int)sc.side;
        const float vx = (sc.c1.x - sc.c0.x) / (float)side, vy = (sc.c1.y - sc.c0.y) / (float)side,
                    vz = (sc.c1.z - sc.c0.z) / (float)side;
        float r = 0.5f * fminf(vx, fminf(vy, vz));
        // the whole grid is inside this radius of any point that matters; beyond it every voxel is scanned
        const float r_all = 2.0f * (fabsf(pt.x - sc.c0.x) + fabsf(pt.y - sc.c0.y) + fabsf(pt.z - sc.c0.z) +
                                    (sc.c1.x - sc.c0.x) + (sc.c1.y - sc.c0.y) + (sc.c1.z - sc.c0.z));
        for (;;) {
            // voxel range overlapping [pt - r, pt + r], one voxel of slack for rounding, clamped
            const int x0 = max(0, (int)floorf((pt.x - r - sc.c0.x) / vx) - 1), x1 = min(side - 1, (int)floorf((pt.x + r - sc.c0.x) / vx) + 1);
            const int y0 = max(0, (int)floorf((pt.y - r - sc.c0.y) / vy) - 1), y1 = min(side - 1, (int)floorf((pt.y + r - sc.c0.y) / vy) + 1);
            const int z0 = max(0, (int)floorf((pt.z - r - sc.c0.z) / vz) - 1), z1 = min(side - 1, (int)floorf((pt.z + r - sc.c0.z) / vz) + 1);
            for (int ix = x0; ix <= x1; ++ix) {
                for (int iy = y0; iy <= y1; ++iy) {
                    for (int iz = z0; iz <= z1; ++iz) {
                        const uint2 cell = sc.cells[((size_t)ix * side + iy) * side + iz];
                        for (uint32_t e = 0; e != cell.y; ++e) {
                            const rt::VoxEntry& ve = sc.entries[cell.x + e];
                            const float nd = point_triangle_distance_squared(ve.pre, pt);
                            if (nd < distance || (nd == distance && ve.tri < best)) {
                                best = ve.tri;
                                distance = nd;
                            }
                        }
                    }
                }
            }
            if (distance <= r * r || r > r_all) break;
            r *= 2.0f;
        }
    }
    surface_out[k] = sc.triangles[best].surface;
}

// boundary_coefficient_finder_2d / _3d (boundary_coefficient_program.cpp:345-484)
template <int N>
static __global__ void mesh_find_nd(const wvb_condensed_node* __restrict__ nodes, Desc d,
                             const uint32_t* __restrict__ idx1, uint32_t* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long nn = (long long)d.dx * d.dy * d.dz;
    if (i >= nn) return;
    const int bt = nodes[i].boundary_type;
    if (__popc((unsigned)bt) != N) return;
    if ((bt & WVB_ID_INSIDE) || (bt & WVB_ID_REENTRANT)) return;
    const uint32_t this_bi = nodes[i].boundary_index;
    const int x = (int)(i % d.dx), y = (int)((i / d.dx) % d.dy), z = (int)(i / d.dx / d.dy);
    const int adj2[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
    const int adj3[12][3] = {{-1, -1, 0}, {-1, 1, 0}, {1, -1, 0}, {1, 1, 0},  {-1, 0, -1}, {-1, 0, 1},
                             {1, 0, -1},  {1, 0, 1},  {0, -1, -1}, {0, -1, 1}, {0, 1, -1}, {0, 1, 1}};
    uint32_t count = 0;
    for (uint32_t p = 0; p != 6; ++p) {
        if (!(bt & (1 << (p + 1)))) continue;
        const int nadj = N == 2 ? 6 : 12;
        for (int j = 0; j != nadj; ++j) {
            const int ax = x + (N == 2 ? adj2[j][0] : adj3[j][0]);
            const int ay = y + (N == 2 ? adj2[j][1] : adj3[j][1]);
            const int az = z + (N == 2 ? adj2[j][2] : adj3[j][2]);
            if (ax < 0 || ay < 0 || az < 0 || ax >= d.dx || ay >= d.dy || az >= d.dz) continue;
            const wvb_condensed_node a = nodes[((long long)az * d.dy + ay) * d.dx + ax];
            if (__popc((unsigned)a.boundary_type) != 1) continue;
            out[(size_t)this_bi * N + count] = idx1[a.boundary_index];
            count += 1;
            break;
        }
    }
}

}  // namespace mesh
}  // namespace wvb
