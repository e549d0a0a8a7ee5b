// oracle/scene_oracle.cpp -- TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's
// CPU legs). Never linked into or loaded by the product.
//
// CPU restatement of how the reference turns scene geometry into the voxel index the ray
// kernels walk:
//   make_voxelised_scene_data        src/core/include/core/spatial_division/voxelised_scene_data.h:27-71
//   ndim_tree<3> (the octree)        src/core/include/core/spatial_division/ndim_tree.h:14-117
//   relative_position                src/core/include/core/indexing.h:43-47
//   voxelise / voxel_collection<3>   src/core/include/core/spatial_division/voxel_collection.h:66-118
//   get_flattened                    src/core/src/spatial_division/voxel_collection.cpp:9-37
//   geo::overlaps(box, triangle)     src/core/src/geo/box.cpp:21-27
//   t_c_intersection                 src/core/src/geo/tri_cube_intersection.cpp:131-170
//   util::range, padded, centre ...  src/utilities/include/utilities/range.h:10-155
// The tree is built as the reference builds it -- every node stores its item list and owns
// eight children -- and then walked into the voxel grid; this is deliberately NOT the flat
// descent of the product (wayverb_b200/csrc/scene_host.cpp).
//
// Parity pinning: PINNED to reference-run output, with one stated caveat. The reference's host C++
// needs GLM, which is fetched at configure time and is not in this image; oracle/ref_recipe/build.py
// compiles the reference's OWN tri_cube_intersection.cpp, box.cpp, geometric.cpp, voxel_collection.cpp,
// ndim_tree.h, voxel_collection.h, voxelised_scene_data.h (make_voxelised_scene_data as written),
// indexing.h and utilities/range.h, whole, unmodified and where they lie, behind a stand-in for the
// GLM operations they use (oracle/ref_recipe/hoststubs/glm/glm.hpp: componentwise operators, dot,
// cross, normalize, min/max ... after GLM 0.9.8.1's generic code paths -- the only arithmetic that
// stand-in decides).
// tests/test_ref_pin_scene.py asserts that this file AND the product's wvb_voxelise reproduce that
// build's flattened index entry for entry on the demo concert hall, its subdivided variants and
// random triangle soups. On top: the reference's own property test (core/tests/voxel_tests.cpp:
// voxel walk == brute force) on the same scenes, tests/test_scene.py, tests/test_config5_gpu.py.
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <vector>

namespace {

struct vec3 {
    float x, y, z;
};
vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
vec3 operator*(vec3 a, vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
vec3 operator/(vec3 a, vec3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
float dot(vec3 a, vec3 b) {
    const vec3 t = a * b;
    return t.x + t.y + t.z;
}
vec3 cross(vec3 a, vec3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
vec3 abs3(vec3 a) { return {std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)}; }
vec3 min3(vec3 a, vec3 b) { return {b.x < a.x ? b.x : a.x, b.y < a.y ? b.y : a.y, b.z < a.z ? b.z : a.z}; }
vec3 max3(vec3 a, vec3 b) { return {a.x < b.x ? b.x : a.x, a.y < b.y ? b.y : a.y, a.z < b.z ? b.z : a.z}; }
vec3 splat(float v) { return {v, v, v}; }

// util::range<glm::vec3> (range.h:10-96)
struct range3 {
    vec3 mn{0, 0, 0}, mx{0, 0, 0};
    range3() = default;
    range3(vec3 a, vec3 b) : mn{min3(a, b)}, mx{max3(a, b)} {}  // maintain_invariant
};
range3 padded(const range3& r, vec3 p) { return range3{r.mn - p, r.mx + p}; }   // range.h:141-145
vec3 centre(const range3& r) { return (r.mn + r.mx) * splat(0.5f); }              // :147-150
vec3 dimensions(const range3& r) { return r.mx - r.mn; }                          // :152-155
range3 shifted(const range3& r, vec3 v) {                                         // operator+ :115-118
    range3 out = r;
    out.mn = out.mn + v;
    out.mx = out.mx + v;
    return out;
}

using triangle_vec3 = std::array<vec3, 3>;

// tri_cube_intersection.cpp:131-170
bool t_c_intersection(const triangle_vec3& v) {
    const std::array<vec3, 3> f{{v[1] - v[0], v[2] - v[1], v[0] - v[2]}};
    const vec3 axes[] = {{0, -f[0].z, f[0].y}, {0, -f[1].z, f[1].y}, {0, -f[2].z, f[2].y},
                         {f[0].z, 0, -f[0].x}, {f[1].z, 0, -f[1].x}, {f[2].z, 0, -f[2].x},
                         {-f[0].y, f[0].x, 0}, {-f[1].y, f[1].x, 0}, {-f[2].y, f[2].x, 0}};
    for (const vec3& a : axes) {
        const float coll[3] = {dot(a, v[0]), dot(a, v[1]), dot(a, v[2])};
        const float r = dot(abs3(a), splat(0.5f));
        float hi = coll[0], lo = coll[0];
        for (int i = 1; i < 3; ++i) {
            if (hi < coll[i]) hi = coll[i];
            if (coll[i] < lo) lo = coll[i];
        }
        if (std::max(-hi, lo) > r) return false;
    }
    const vec3 mm_min = min3(min3(v[0], v[1]), v[2]), mm_max = max3(max3(v[0], v[1]), v[2]);
    if (mm_max.x < -0.5f || mm_max.y < -0.5f || mm_max.z < -0.5f) return false;
    if (0.5f < mm_min.x || 0.5f < mm_min.y || 0.5f < mm_min.z) return false;
    const vec3 c = cross(f[0], f[2]);
    const float inv = 1.0f / std::sqrt(dot(c, c));
    const vec3 normal = c * splat(inv);
    const float dist = dot(normal, v[0]);
    const float r = dot(abs3(normal), splat(0.5f));
    return std::fabs(dist) <= r;
}

// box.cpp:21-27
bool overlaps(const range3& b, const triangle_vec3& t) {
    triangle_vec3 coll = t;
    for (auto& i : coll) i = (i - centre(b)) / dimensions(b);
    return t_c_intersection(coll);
}

// ndim_tree<3> (ndim_tree.h:41-117)
struct tree {
    using checker = std::function<bool(size_t, const range3&)>;
    range3 aabb;
    std::vector<size_t> items;
    std::unique_ptr<std::array<tree, 8>> nodes;

    tree() = default;
    tree(size_t depth, const checker& cb, const std::vector<size_t>& to_test, const range3& box) : aabb{box} {
        for (size_t i : to_test) {
            if (cb(i, box)) items.push_back(i);       // compute_contained_items :68-77
        }
        if (depth) {                                  // compute_nodes :79-96
            const vec3 c = centre(box);               // next_boundaries :17-35
            const range3 root{box.mn, c};
            const vec3 d = dimensions(root);
            nodes = std::make_unique<std::array<tree, 8>>();
            for (size_t i = 0; i != 8; ++i) {
                const vec3 rel{float(i & 1), float((i >> 1) & 1), float((i >> 2) & 1)};  // indexing.h:43-47
                (*nodes)[i] = tree(depth - 1, cb, items, shifted(root, d * rel));
            }
        }
    }
    size_t side() const { return nodes ? 2 * nodes->front().side() : 1; }  // :61-63
};

// voxelise (voxel_collection.h:66-85)
void voxelise(const tree& t, unsigned px, unsigned py, unsigned pz, size_t side, std::vector<std::vector<size_t>>& out) {
    if (!t.nodes) {
        out[(size_t(px) * side + py) * side + pz] = t.items;
        return;
    }
    const unsigned half = unsigned(t.side() / 2);
    for (size_t i = 0; i != 8; ++i) {
        voxelise((*t.nodes)[i], px + unsigned(i & 1) * half, py + unsigned((i >> 1) & 1) * half,
                 pz + unsigned((i >> 2) & 1) * half, side, out);
    }
}

}  // namespace

extern "C" {

// vertices: n x 4 floats (cl_float3), triangles: n x {surface, v0, v1, v2}. Returns the length of
// the flattened index; writes it when `out` is non-null and large enough. aabb6 = min, max.
size_t sco_voxelise(const float* vertices, size_t nv, const uint32_t* triangles, size_t nt, size_t depth,
                    float padding, float* aabb6, uint32_t* out, size_t capacity) {
    if (!nv || !nt) return 0;
    vec3 lo{vertices[0], vertices[1], vertices[2]}, hi = lo;
    for (size_t i = 1; i < nv; ++i) {
        const vec3 p{vertices[4 * i], vertices[4 * i + 1], vertices[4 * i + 2]};
        lo = min3(lo, p);
        hi = max3(hi, p);
    }
    const range3 aabb = padded(range3{lo, hi}, splat(padding));  // voxelised_scene_data.h:66-70
    std::vector<size_t> all(nt);
    for (size_t i = 0; i < nt; ++i) all[i] = i;
    const tree root(
            depth,
            [&](size_t item, const range3& box) {  // voxelised_scene_data.h:34-42
                const uint32_t* t = triangles + 4 * item;
                triangle_vec3 tv;
                for (int k = 0; k < 3; ++k) {
                    const float* p = vertices + 4 * size_t(t[1 + k]);
                    tv[size_t(k)] = vec3{p[0], p[1], p[2]};
                }
                return overlaps(padded(box, splat(0.001f)), tv);
            },
            all, aabb);
    const size_t side = root.side();
    std::vector<std::vector<size_t>> cells(side * side * side);
    voxelise(root, 0, 0, 0, side, cells);
    // get_flattened (voxel_collection.cpp:9-37)
    std::vector<uint32_t> ret(side * side * side);
    for (size_t x = 0; x != side; ++x) {
        for (size_t y = 0; y != side; ++y) {
            for (size_t z = 0; z != side; ++z) {
                ret[x * side * side + y * side + z] = uint32_t(ret.size());
                const auto& v = cells[(x * side + y) * side + z];
                ret.push_back(uint32_t(v.size()));
                for (size_t i : v) ret.push_back(uint32_t(i));
            }
        }
    }
    if (aabb6) {
        aabb6[0] = aabb.mn.x; aabb6[1] = aabb.mn.y; aabb6[2] = aabb.mn.z;
        aabb6[3] = aabb.mx.x; aabb6[4] = aabb.mx.y; aabb6[5] = aabb.mx.z;
    }
    if (out && capacity >= ret.size()) std::memcpy(out, ret.data(), ret.size() * sizeof(uint32_t));
    return ret.size();
}

// the overlap predicate alone, for tests: box (min, max) grown by 0.001 against one triangle
int sco_overlaps(const float* box6, const float* tri9) {
    const range3 b{vec3{box6[0], box6[1], box6[2]}, vec3{box6[3], box6[4], box6[5]}};
    const triangle_vec3 t{{vec3{tri9[0], tri9[1], tri9[2]}, vec3{tri9[3], tri9[4], tri9[5]},
                           vec3{tri9[6], tri9[7], tri9[8]}}};
    return overlaps(padded(b, splat(0.001f)), t) ? 1 : 0;
}

}  // extern "C"
