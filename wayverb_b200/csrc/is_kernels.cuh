// is_kernels.cuh -- sm_100a device code of the image-source stage.
//
// The reference collects, per ray, the triangles of its first `max_order`
// reflections (reflection_path_builder.h:16-26), merges all rays' paths into a
// multitree keyed by triangle index (multitree.h:13-36, tree.h:18-38), then walks
// the tree on the host: each node mirrors its parent's image source in the node's
// triangle and, if the node was visible from the receiver, validates the path by
// casting rays receiver -> image source -> ... -> source through the voxelised
// scene (tree.cpp:23-181). Valid paths become impulses (fast_pressure_calculator.h).
//
// Here:
//   is_insert    one thread per ray walks its path and inserts (parent node,
//                triangle) keys into an open-addressing table with 64-bit CAS: the
//                slot index IS the node id, so a node's key names its parent and the
//                table is the tree. The first ray to reach a node decides `visible`
//                (set insert does not replace, recursive_vector.h:249-255): an
//                atomicMin on (ray index << 1 | visible). Each inserting thread
//                carries the running image source and stores it in the node (all
//                writers store the same value).
//   is_validate  one thread per table slot: for a visible node walk UP the parent
//                links -- the order find_valid_path consumes the image sources in
//                (tree.cpp:121-157) -- with the host code's own ray/scene routines
//                restated on the device, and append the impulse.
//   is_chains    writes the triangle path of every emitted impulse so that the host
//                can return them in the reference's tree order.
//
// fp32 in the host code's operation order (glm 0.9.8.1 definitions, see
// oracle/is_oracle.inc); built with -fmad=false.
#pragma once

#include "rt_kernels.cuh"

namespace wvb {
namespace is {

using rt::f3;
using rt::Scene;

constexpr uint32_t ELEM_NONE = 0xffffffffu;
constexpr uint32_t ELEM_VISIBLE = 0x80000000u;
constexpr uint32_t ROOT = 0xffffffffu;
constexpr unsigned long long KEY_EMPTY = ~0ull;

struct Table {
    unsigned long long* keys;   // (parent slot << 32) | triangle, KEY_EMPTY when free
    unsigned long long* first;  // min over rays of (ray << 1 | visible)
    float* image;               // [slot][3]
    uint32_t mask;              // capacity - 1 (power of two)
};

struct Impulse {  // raytracer::impulse<8>, 64 B (raytracer/cl/structs.h:37-44)
    float volume[8];
    float position[4];
    float distance;
    uint32_t slot;   // pad_ of the reference struct: node id and path length, cleared by the host
    uint32_t depth;
    uint32_t pad_;
};
static_assert(sizeof(Impulse) == 64, "impulse<8> layout");

struct Query {
    f3 source, receiver;
    double distance_scale;  // sqrt(acoustic_impedance / (4 pi)), pressure_intensity.cpp:10-12
    float flip;             // -1 or 1 (fast_pressure_calculator.h:57)
};

__device__ __forceinline__ bool eq3(f3 a, f3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

// geo::mirror (geometric.cpp:344-348) with the precomputed unit normal
__device__ __forceinline__ f3 mirror(const rt::TriPre& T, f3 p) {
    const f3 n = rt::mk(T.nx, T.ny, T.nz);
    const float d = rt::dot(n, rt::sub(p, rt::mk(T.v0x, T.v0y, T.v0z)));
    return rt::sub(p, rt::mul(rt::mul(n, d), 2.0f));
}

// ---- tree build ------------------------------------------------------------------
// element k of ray r: from path elements [order][n] or from reflection records
__device__ __forceinline__ uint32_t load_element(const uint32_t* __restrict__ elems,
                                                 const rt::ReflectionPod* __restrict__ refl,
                                                 size_t k, size_t n, size_t r) {
    if (elems) return elems[k * n + r];
    const rt::ReflectionPod x = refl[k * n + r];
    // reflection_path_builder.h:17-24
    return x.keep_going ? (x.triangle | (x.receiver_visible ? ELEM_VISIBLE : 0u)) : ELEM_NONE;
}

static __global__ void is_insert(Table tab, Scene sc, f3 source, const uint32_t* __restrict__ elems,
                                 const rt::ReflectionPod* __restrict__ refl, uint32_t n, uint32_t order,
                                 unsigned long long ray_base, unsigned long long* __restrict__ counters) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    uint32_t parent = ROOT;
    f3 img = source;
    for (uint32_t k = 0; k < order; ++k) {
        const uint32_t e = load_element(elems, refl, k, n, r);
        if (e == ELEM_NONE) break;
        const uint32_t tri = e & ~ELEM_VISIBLE;
        if (tri >= sc.n_triangles) {  // malformed input: report, do not touch memory
            atomicAdd(counters + 3, 1ull);
            break;
        }
        const unsigned long long key = ((unsigned long long)parent << 32) | tri;
        uint32_t slot = (uint32_t)mix64(key) & tab.mask;
        for (;;) {
            const unsigned long long old = atomicCAS(tab.keys + slot, KEY_EMPTY, key);
            if (old == KEY_EMPTY) {
                atomicAdd(counters + 0, 1ull);  // nodes
                break;
            }
            if (old == key) break;
            slot = (slot + 1) & tab.mask;
        }
        atomicMin(tab.first + slot, ((ray_base + r) << 1) | (e >> 31));
        img = mirror(sc.pre[tri], img);
        tab.image[3 * (size_t)slot] = img.x;
        tab.image[3 * (size_t)slot + 1] = img.y;
        tab.image[3 * (size_t)slot + 2] = img.z;
        parent = slot;
    }
}

// ---- the host code's ray / scene routines, on the device ---------------------------
struct Ray {
    f3 pos, dir;
};
// construct_ray + geo::ray's constructor (tree.cpp:12-19, geometric.cpp:13-15):
// the direction is normalized twice
__device__ __forceinline__ bool construct_ray(f3 from, f3 to, Ray& out) {
    if (eq3(from, to)) return false;
    out.pos = from;
    out.dir = rt::normalize(rt::normalize(rt::sub(to, from)));
    return true;
}

// intersects(box, ray)         box.cpp:29-65
__device__ __forceinline__ bool box_entry_distance(f3 bmin, f3 bmax, const Ray& r, float& out) {
    const float ix = 1.0f / r.dir.x, iy = 1.0f / r.dir.y, iz = 1.0f / r.dir.z;
    float t0 = ((ix < 0 ? bmax.x : bmin.x) - r.pos.x) * ix;
    float t1 = ((ix < 0 ? bmin.x : bmax.x) - r.pos.x) * ix;
    const float ty0 = ((iy < 0 ? bmax.y : bmin.y) - r.pos.y) * iy;
    const float ty1 = ((iy < 0 ? bmin.y : bmax.y) - r.pos.y) * iy;
    if (ty1 < t0 || t1 < ty0) return false;
    t0 = ty0 < t0 ? t0 : ty0;  // std::max(ty0, t0)
    t1 = t1 < ty1 ? t1 : ty1;  // std::min(ty1, t1)
    const float tz0 = ((iz < 0 ? bmax.z : bmin.z) - r.pos.z) * iz;
    const float tz1 = ((iz < 0 ? bmin.z : bmax.z) - r.pos.z) * iz;
    if (tz1 < t0 || t1 < tz0) return false;
    const float first = tz0 < t0 ? t0 : tz0;    // std::max(tz0, t0)
    const float second = t1 < tz1 ? t1 : tz1;   // std::min(tz1, t1)
    if (0 < first) { out = first; return true; }
    if (0 < second) { out = second; return true; }
    return false;
}

__device__ __forceinline__ int to_cell(float q, int side) {
    int v;
    if (!(q > -2147483648.0f)) v = INT_MIN;
    else if (!(q < 2147483648.0f)) v = INT_MAX;
    else v = (int)q;  // toward zero, like ivec3(vec3)
    return max(0, min(side - 1, v));
}

// traverse + intersects(voxelised, ray, to_ignore)
// (voxel_collection.cpp:41-124, voxelised_scene_data.h:80-106, geometric.cpp:73-125)
__device__ __forceinline__ bool cpu_intersects(const Scene& sc, const Ray& r, uint32_t to_ignore,
                                               float& t_out, uint32_t& index_out) {
    const int side = (int)sc.side;
    const float sidef = (float)sc.side;
    const f3 vd = rt::mk((sc.c1.x - sc.c0.x) / sidef, (sc.c1.y - sc.c0.y) / sidef,
                         (sc.c1.z - sc.c0.z) / sidef);
    f3 p = r.pos;
    const bool inside = sc.c0.x < p.x && sc.c0.y < p.y && sc.c0.z < p.z && p.x < sc.c1.x &&
                        p.y < sc.c1.y && p.z < sc.c1.z;
    if (!inside) {
        float d;
        if (!box_entry_distance(sc.c0, sc.c1, r, d)) return false;
        p = rt::add(r.pos, rt::mul(r.dir, d));
    }
    int ind[3] = {to_cell((p.x - sc.c0.x) / vd.x, side), to_cell((p.y - sc.c0.y) / vd.y, side),
                  to_cell((p.z - sc.c0.z) / vd.z, side)};
    const f3 root = rt::mk(sc.c0.x + vd.x * (float)ind[0], sc.c0.y + vd.y * (float)ind[1],
                           sc.c0.z + vd.z * (float)ind[2]);
    const f3 top = rt::add(root, vd);
    const f3 dims = rt::sub(top, root);
    const float dir[3] = {r.dir.x, r.dir.y, r.dir.z};
    const float pos[3] = {r.pos.x, r.pos.y, r.pos.z};
    const float lo[3] = {root.x, root.y, root.z}, hi[3] = {top.x, top.y, top.z};
    const float dm[3] = {dims.x, dims.y, dims.z};
    int step[3], just_out[3];
    float t_max[3], t_delta[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const bool gt = 0.0f <= dir[i];
        step[i] = gt ? 1 : -1;
        just_out[i] = gt ? side : -1;
        const float tmp = fabsf(((gt ? hi[i] : lo[i]) - pos[i]) / dir[i]);
        t_max[i] = isnan(tmp) ? INFINITY : tmp;
        t_delta[i] = fabsf(dm[i] / dir[i]);
    }
    // The host code nests "for each voxel { for each triangle }"; as in rt_kernels.cuh's
    // voxel_traversal the walk runs here as a flat state machine paced by a warp vote -- per
    // iteration a lane enters a voxel, tests ONE triangle of it, or leaves it; same visiting and
    // testing order, same arithmetic -- so that the lanes of a warp work side by side instead of
    // waiting for each other's inner loops.
    const unsigned lanes = __activemask();
    bool done = false, found = false, enter = true, hit = false;
    int min_i = 0;
    float tm = 0.0f, best = 0.0f;
    uint32_t best_i = 0, k = 0, num = 0;
    const rt::VoxEntry* e = sc.entries;
    while (__any_sync(lanes, !done)) {
        if (!done) {
            if (enter) {
                min_i = 0;
                if (t_max[1] < t_max[min_i]) min_i = 1;
                if (t_max[2] < t_max[min_i]) min_i = 2;
                tm = min_i == 0 ? t_max[0] : (min_i == 1 ? t_max[1] : t_max[2]);
                const uint2 cell = sc.cells[(size_t)ind[0] * side * side + (size_t)ind[1] * side + ind[2]];
                e = sc.entries + cell.x;
                num = cell.y;
                k = 0;
                hit = false;
                best = 0.0f;
                enter = false;
            }
            if (k < num) {
                const uint32_t ti = e[k].tri;
                if (ti != to_ignore) {
                    const float t = rt::tri_intersection(e[k].pre, r.pos, r.dir);
                    if (t != 0.0f && (!hit || t < best)) {
                        hit = true;
                        best = t;
                        best_i = ti;
                    }
                }
                ++k;
            }
            if (k >= num) {
                if (hit && best <= tm) {
                    t_out = best;
                    index_out = best_i;
                    found = true;
                    done = true;
                } else {
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        if (i == min_i) {
                            ind[i] += step[i];
                            if (ind[i] == just_out[i]) done = true;
                            t_max[i] += t_delta[i];
                        }
                    }
                    enter = true;
                }
            }
        }
    }
    return found;
}

// ---- validation ------------------------------------------------------------------
// counters: [0] nodes, [1] visible nodes, [2] construct_ray failures, [3] bad elements,
//           [4] impulses emitted
// The table is at most half full and only visible nodes are validated: gather their
// slots into a dense list first (warp-aggregated append), so that the validation
// kernel starts with full warps.
static __global__ void is_collect(Table tab, uint32_t* __restrict__ list,
                                  unsigned long long* __restrict__ counters) {
    const size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool take = false;
    if (slot <= tab.mask) take = tab.keys[slot] != KEY_EMPTY && (tab.first[slot] & 1ull);
    const unsigned m = __ballot_sync(0xffffffffu, take);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(counters + 1, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (take) list[base + __popc(m & ((1u << lane) - 1))] = (uint32_t)slot;
}

static __global__ void __launch_bounds__(128)
is_validate(Table tab, Scene sc, Query q, const float* __restrict__ impedance /* [surface][8] */,
            const uint32_t* __restrict__ list, uint32_t n_list, Impulse* __restrict__ out,
            unsigned long long cap, unsigned long long* __restrict__ counters) {
    const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_list) return;
    const size_t slot0 = list[li];
    const f3 final_image = rt::mk(tab.image[3 * slot0], tab.image[3 * slot0 + 1], tab.image[3 * slot0 + 2]);
    if (eq3(q.receiver, final_image)) return;  // tree.cpp:108-111

    f3 prev_intersection = q.receiver;
    uint32_t prev_surface = ~0u;
    float volume[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) volume[b] = 1.0f;
    uint32_t depth = 0;
    uint32_t cur = (uint32_t)slot0;
    while (cur != ROOT) {
        const unsigned long long key = tab.keys[cur];
        const uint32_t tri = (uint32_t)key;
        const f3 image = rt::mk(tab.image[3 * (size_t)cur], tab.image[3 * (size_t)cur + 1],
                                tab.image[3 * (size_t)cur + 2]);
        Ray ray;
        if (!construct_ray(prev_intersection, image, ray)) {
            atomicAdd(counters + 2, 1ull);
            return;
        }
        float t;
        uint32_t idx;
        if (!cpu_intersects(sc, ray, prev_surface, t, idx) || idx != tri) return;
        const rt::TriPre& T = sc.pre[tri];
        const float d = fabsf(rt::dot(ray.dir, rt::mk(T.nx, T.ny, T.nz)));
        const float capped = d < 1.0f ? d : 1.0f;                   // std::min(1, d)
        const float cos_angle = 0.0f < capped ? capped : 0.0f;      // std::max(0, .)
        const uint32_t surface = sc.triangles[tri].surface;
        const float* imp = impedance + (size_t)surface * 8;
        const float* scat = sc.surfaces + (size_t)surface * 16 + 8;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const float tmp = imp[b] * cos_angle;  // surfaces.h:41-49
            const float reflectance = (tmp - 1) / (tmp + 1);
            const float outgoing = (volume[b] * reflectance) * (1 - scat[b]);
            volume[b] = outgoing * q.flip;
        }
        prev_intersection = rt::add(ray.pos, rt::mul(ray.dir, t));
        prev_surface = tri;
        cur = (uint32_t)(key >> 32);
        ++depth;
    }
    {
        Ray ray;
        if (!construct_ray(q.source, prev_intersection, ray)) {
            atomicAdd(counters + 2, 1ull);
            return;
        }
        float t;
        uint32_t idx;
        if (!cpu_intersects(sc, ray, ~0u, t, idx) || idx != prev_surface) return;
    }
    const unsigned long long o = atomicAdd(counters + 4, 1ull);
    if (o >= cap) return;
    Impulse imp;
    imp.distance = rt::length(rt::sub(q.receiver, final_image));
    const double p = q.distance_scale / (double)imp.distance;  // image_source.cpp:61-65
#pragma unroll
    for (int b = 0; b < 8; ++b) imp.volume[b] = (float)((double)volume[b] * p);
    imp.position[0] = final_image.x;
    imp.position[1] = final_image.y;
    imp.position[2] = final_image.z;
    imp.position[3] = 0.0f;
    imp.slot = (uint32_t)slot0;
    imp.depth = depth;
    imp.pad_ = 0;
    out[o] = imp;
}

// get_direct (get_direct.h:16-41): one thread
static __global__ void is_direct(Scene sc, Query q, Impulse* __restrict__ out, uint32_t* __restrict__ have) {
    *have = 0;
    if (eq3(q.source, q.receiver)) return;
    const f3 s2r = rt::sub(q.receiver, q.source);
    const float len = rt::length(s2r);
    Ray ray;
    ray.pos = q.source;
    ray.dir = rt::normalize(rt::normalize(s2r));
    float t;
    uint32_t idx;
    if (cpu_intersects(sc, ray, ~0u, t, idx) && !(t >= len)) return;
    Impulse imp;
    const double p = q.distance_scale / (double)len;
    for (int b = 0; b < 8; ++b) imp.volume[b] = (float)(1.0 * p);
    imp.position[0] = q.source.x;
    imp.position[1] = q.source.y;
    imp.position[2] = q.source.z;
    imp.position[3] = 0.0f;
    imp.distance = len;
    imp.slot = imp.depth = imp.pad_ = 0;
    *out = imp;
    *have = 1;
}

// triangle path (root first, +1 so that 0 pads and sorts before every triangle)
static __global__ void is_chains(Table tab, const Impulse* __restrict__ imps, uint32_t n, uint32_t width,
                                 uint32_t* __restrict__ chains) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t cur = imps[i].slot;
    uint32_t d = imps[i].depth;
    uint32_t* row = chains + (size_t)i * width;
    for (uint32_t k = d; k < width; ++k) row[k] = 0;
    while (cur != ROOT && d > 0) {
        const unsigned long long key = tab.keys[cur];
        row[--d] = (uint32_t)key + 1u;
        cur = (uint32_t)(key >> 32);
    }
}

}  // namespace is
}  // namespace wvb
