// scene_host.cpp -- the two host-side steps that turn a model file into what the ray path
// consumes (include/wvb200.h: wvb_obj_parse, wvb_voxelise).
//
//   reference                                                        here
//   src/core/src/scene_data_loader.cpp:17-70 (assimp, triangulated)   wvb_obj_parse: Wavefront OBJ only
//   src/core/include/core/spatial_division/voxelised_scene_data.h:27-43
//     + ndim_tree.h:47-117 (octree over triangle indices)
//     + voxel_collection.h:66-85 (tree -> side^3 voxels)
//     + src/core/src/spatial_division/voxel_collection.cpp:9-37       wvb_voxelise
//
// Host code in the reference, host code here (it runs once per scene). The octree is not
// materialised: a depth-first descent carries the surviving triangle list of each node
// (children only test what their parent kept, exactly as ndim_tree's constructor does), the
// 64 depth-2 subtrees run on separate threads, and leaves write their list straight into the
// voxel they cover. The float arithmetic of the box subdivision and of the box/triangle
// overlap test (geo/box.cpp:21-27, geo/tri_cube_intersection.cpp:131-170) is kept operation
// by operation -- the file is compiled with -ffp-contract=off -- so that the lists are the
// ones the reference's octree produces, not merely conservative.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "common.h"

namespace {

struct v3 {
    float x, y, z;
};
inline v3 operator+(v3 a, v3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline v3 operator-(v3 a, v3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline v3 operator*(v3 a, v3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline v3 operator/(v3 a, v3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline v3 operator*(v3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot(v3 a, v3 b) {  // glm::dot: componentwise product, then x + y + z
    const v3 t = a * b;
    return t.x + t.y + t.z;
}
inline v3 cross(v3 a, v3 b) {  // glm::cross
    return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
inline v3 vabs(v3 a) { return {std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)}; }
inline v3 vmin(v3 a, v3 b) { return {std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)}; }
inline v3 vmax(v3 a, v3 b) { return {std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)}; }

struct box {
    v3 lo, hi;
};

// t_c_intersection (geo/tri_cube_intersection.cpp:131-170): separating-axis test of a triangle
// against the unit cube centred on the origin
bool tri_hits_unit_cube(const v3 v[3]) {
    const v3 f[3] = {v[1] - v[0], v[2] - v[1], v[0] - v[2]};
    const v3 axes[9] = {{0, -f[0].z, f[0].y}, {0, -f[1].z, f[1].y}, {0, -f[2].z, f[2].y},
                        {f[0].z, 0, -f[0].x}, {f[1].z, 0, -f[1].x}, {f[2].z, 0, -f[2].x},
                        {-f[0].y, f[0].x, 0}, {-f[1].y, f[1].x, 0}, {-f[2].y, f[2].x, 0}};
    const v3 half{0.5f, 0.5f, 0.5f};
    for (const v3& a : axes) {
        const float p0 = dot(a, v[0]), p1 = dot(a, v[1]), p2 = dot(a, v[2]);
        const float r = dot(vabs(a), half);
        const float mx = std::max(std::max(p0, p1), p2), mn = std::min(std::min(p0, p1), p2);
        if (std::max(-mx, mn) > r) return false;
    }
    const v3 lo = vmin(vmin(v[0], v[1]), v[2]), hi = vmax(vmax(v[0], v[1]), v[2]);
    if (hi.x < -0.5f || hi.y < -0.5f || hi.z < -0.5f) return false;
    if (0.5f < lo.x || 0.5f < lo.y || 0.5f < lo.z) return false;
    const v3 c = cross(f[0], f[2]);
    const v3 normal = c * (1.0f / std::sqrt(dot(c, c)));  // glm::normalize = v * inversesqrt(dot)
    const float dist = dot(normal, v[0]);
    const float r = dot(vabs(normal), half);
    return std::fabs(dist) <= r;  // false for a degenerate triangle (nan), like the reference
}

// the item_checker of voxelised_scene_data.h:34-42: overlaps(padded(aabb, 0.001), triangle),
// geo/box.cpp:21-27
bool overlaps_padded(const box& b, const v3 tri[3]) {
    const v3 pad{0.001f, 0.001f, 0.001f};
    const box p{b.lo - pad, b.hi + pad};
    const v3 centre = (p.lo + p.hi) * 0.5f;
    const v3 dim = p.hi - p.lo;
    const v3 t[3] = {(tri[0] - centre) / dim, (tri[1] - centre) / dim, (tri[2] - centre) / dim};
    return tri_hits_unit_cube(t);
}

struct voxeliser {
    const float* verts;  // cl_float3: 4 floats per vertex
    const wvb_triangle* tris;
    uint32_t side;
    std::vector<std::vector<uint32_t>> cells;  // [x * side * side + y * side + z]

    void triangle(uint32_t i, v3 out[3]) const {
        const uint32_t idx[3] = {tris[i].v0, tris[i].v1, tris[i].v2};
        for (int k = 0; k < 3; ++k) out[k] = {verts[4 * idx[k]], verts[4 * idx[k] + 1], verts[4 * idx[k] + 2]};
    }
    // compute_contained_items (ndim_tree.h:68-77)
    std::vector<uint32_t> keep(const std::vector<uint32_t>& from, const box& b) const {
        std::vector<uint32_t> out;
        v3 t[3];
        for (uint32_t i : from) {
            triangle(i, t);
            if (overlaps_padded(b, t)) out.push_back(i);
        }
        return out;
    }
    // next_boundaries (ndim_tree.h:17-35): child i sits at offset (i & 1, i >> 1 & 1, i >> 2 & 1)
    static box child(const box& parent, unsigned i) {
        const v3 c = (parent.lo + parent.hi) * 0.5f;
        const v3 d = c - parent.lo;
        const v3 rel{float(i & 1u), float((i >> 1) & 1u), float((i >> 2) & 1u)};
        return {parent.lo + d * rel, c + d * rel};
    }
    // node of the tree with its own (already filtered) items, `levels` levels above the leaves,
    // covering voxels [px, px + 2^levels) x ...
    void descend(const box& b, const std::vector<uint32_t>& items, unsigned levels, uint32_t px, uint32_t py,
                 uint32_t pz) {
        if (!levels) {
            cells[((size_t)px * side + py) * side + pz] = items;
            return;
        }
        const uint32_t half = 1u << (levels - 1);
        for (unsigned i = 0; i < 8; ++i) {
            const box cb = child(b, i);
            descend(cb, keep(items, cb), levels - 1, px + (i & 1u) * half, py + ((i >> 1) & 1u) * half,
                    pz + ((i >> 2) & 1u) * half);
        }
    }
};

struct task {
    box b;
    std::vector<uint32_t> items;
    unsigned levels;
    uint32_t px, py, pz;
};

template <class F>
wvb_status guarded(F&& f) {
    try {
        f();
        return WVB_OK;
    } catch (const wvb::status_error& e) {
        return e.code;
    } catch (const std::exception& e) {
        wvb::set_last_error("%s", e.what());
        return WVB_ERR_INVALID;
    }
}

}  // namespace

extern "C" {

wvb_status wvb_voxelise(const wvb_float3* vertices, uint32_t num_vertices, const wvb_triangle* triangles,
                        uint32_t num_triangles, uint32_t octree_depth, float padding, float aabb_min[3],
                        float aabb_max[3], uint32_t* index_out, uint64_t capacity, uint64_t* count) {
    if (!vertices || !triangles || !count || !aabb_min || !aabb_max) return WVB_ERR_INVALID;
    return guarded([&] {
        WVB_REQUIRE(num_vertices > 0 && num_triangles > 0, WVB_ERR_INVALID, "No geometry found in scene file.");
        WVB_REQUIRE(octree_depth <= 8, WVB_ERR_UNSUPPORTED, "octree depth %u > 8", octree_depth);
        const float* v = reinterpret_cast<const float*>(vertices);
        for (uint32_t i = 0; i < num_triangles; ++i) {
            WVB_REQUIRE(triangles[i].v0 < num_vertices && triangles[i].v1 < num_vertices &&
                                triangles[i].v2 < num_vertices,
                        WVB_ERR_INVALID, "triangle %u refers to a vertex beyond %u", i, num_vertices);
        }
        // compute_aabb + padded (voxelised_scene_data.h:66-70)
        v3 lo{v[0], v[1], v[2]}, hi = lo;
        for (uint32_t i = 1; i < num_vertices; ++i) {
            const v3 p{v[4 * i], v[4 * i + 1], v[4 * i + 2]};
            lo = vmin(lo, p);
            hi = vmax(hi, p);
        }
        const v3 pad{padding, padding, padding};
        const box root{lo - pad, hi + pad};
        aabb_min[0] = root.lo.x; aabb_min[1] = root.lo.y; aabb_min[2] = root.lo.z;
        aabb_max[0] = root.hi.x; aabb_max[1] = root.hi.y; aabb_max[2] = root.hi.z;

        voxeliser vx{v, triangles, 1u << octree_depth, {}};
        vx.cells.resize((size_t)vx.side * vx.side * vx.side);
        std::vector<uint32_t> all(num_triangles);
        for (uint32_t i = 0; i < num_triangles; ++i) all[i] = i;
        // the root and the first two levels serially, then one task per depth-2 subtree
        std::vector<task> tasks{{root, vx.keep(all, root), octree_depth, 0, 0, 0}};
        for (int round = 0; round < 2; ++round) {
            std::vector<task> next;
            for (const task& t : tasks) {
                if (!t.levels) {
                    next.push_back(t);
                    continue;
                }
                const uint32_t half = 1u << (t.levels - 1);
                for (unsigned i = 0; i < 8; ++i) {
                    const box cb = voxeliser::child(t.b, i);
                    next.push_back({cb, vx.keep(t.items, cb), t.levels - 1, t.px + (i & 1u) * half,
                                    t.py + ((i >> 1) & 1u) * half, t.pz + ((i >> 2) & 1u) * half});
                }
            }
            tasks.swap(next);
        }
        wvb::parallel_for((int64_t)tasks.size(), [&](int64_t k) {
            const task& t = tasks[(size_t)k];
            vx.descend(t.b, t.items, t.levels, t.px, t.py, t.pz);
        });
        // get_flattened (voxel_collection.cpp:9-37)
        uint64_t total = vx.cells.size();
        for (const auto& c : vx.cells) total += 1 + c.size();
        *count = total;
        if (!index_out) return;
        WVB_REQUIRE(capacity >= total, WVB_ERR_INVALID, "index_out holds %llu entries, %llu needed",
                    (unsigned long long)capacity, (unsigned long long)total);
        WVB_REQUIRE(total < 0xffffffffull, WVB_ERR_UNSUPPORTED, "flattened index exceeds 32-bit offsets");
        uint64_t pos = vx.cells.size();
        for (size_t c = 0; c < vx.cells.size(); ++c) {
            index_out[c] = (uint32_t)pos;
            index_out[pos++] = (uint32_t)vx.cells[c].size();
            for (uint32_t t : vx.cells[c]) index_out[pos++] = t;
        }
    });
}

// Wavefront OBJ: `v x y z`, `f a b c ...` (a = v, v/vt, v/vt/vn or v//vn; 1-based, negative =
// relative to the vertices read so far; polygons fan-triangulated like aiProcess_Triangulate
// does for convex faces), `usemtl name` (material index = order of first use; faces before
// any usemtl get material 0, named "default"). Everything else is skipped.
// Two-pass: call with null outputs to get the counts.
wvb_status wvb_obj_parse(const char* text, uint64_t length, wvb_float3* vertices, uint64_t* num_vertices,
                         wvb_triangle* triangles, uint64_t* num_triangles, char* material_names,
                         uint64_t* material_names_length) {
    if (!text || !num_vertices || !num_triangles) return WVB_ERR_INVALID;
    return guarded([&] {
        std::vector<std::string> materials;
        uint32_t current = 0;
        bool any_default = false;
        uint64_t nv = 0, nt = 0;
        const uint64_t cap_v = vertices ? *num_vertices : 0, cap_t = triangles ? *num_triangles : 0;
        const char* p = text;
        const char* end = text + length;
        uint64_t line_no = 0;
        while (p < end) {
            const char* eol = static_cast<const char*>(memchr(p, '\n', size_t(end - p)));
            if (!eol) eol = end;
            std::string line(p, eol);
            p = eol + 1;
            ++line_no;
            if (!line.empty() && line.back() == '\r') line.pop_back();
            size_t s = line.find_first_not_of(" \t");
            if (s == std::string::npos || line[s] == '#') continue;
            if (line.compare(s, 2, "v ") == 0 || line.compare(s, 2, "v\t") == 0) {
                float x = 0, y = 0, z = 0;
                WVB_REQUIRE(sscanf(line.c_str() + s + 2, "%f %f %f", &x, &y, &z) == 3, WVB_ERR_INVALID,
                            "OBJ line %llu: malformed vertex", (unsigned long long)line_no);
                if (vertices) {
                    WVB_REQUIRE(nv < cap_v, WVB_ERR_INVALID, "vertex buffer too small");
                    vertices[nv] = wvb_float3{x, y, z, 0.0f};
                }
                ++nv;
            } else if (line.compare(s, 2, "f ") == 0 || line.compare(s, 2, "f\t") == 0) {
                std::vector<uint32_t> idx;
                const char* q = line.c_str() + s + 2;
                while (*q) {
                    while (*q == ' ' || *q == '\t') ++q;
                    if (!*q) break;
                    char* e = nullptr;
                    const long long raw = strtoll(q, &e, 10);
                    WVB_REQUIRE(e != q && raw != 0, WVB_ERR_INVALID, "OBJ line %llu: malformed face",
                                (unsigned long long)line_no);
                    const long long abs_i = raw > 0 ? raw - 1 : (long long)nv + raw;
                    WVB_REQUIRE(abs_i >= 0 && abs_i < (long long)nv, WVB_ERR_INVALID,
                                "OBJ line %llu: face refers to vertex %lld of %llu",
                                (unsigned long long)line_no, raw, (unsigned long long)nv);
                    idx.push_back((uint32_t)abs_i);
                    q = e;
                    while (*q && *q != ' ' && *q != '\t') ++q;  // skip /vt/vn
                }
                WVB_REQUIRE(idx.size() >= 3, WVB_ERR_INVALID, "OBJ line %llu: face with %zu vertices",
                            (unsigned long long)line_no, idx.size());
                if (materials.empty()) {
                    materials.push_back("default");
                    any_default = true;
                    current = 0;
                }
                for (size_t k = 1; k + 1 < idx.size(); ++k) {
                    if (triangles) {
                        WVB_REQUIRE(nt < cap_t, WVB_ERR_INVALID, "triangle buffer too small");
                        triangles[nt] = wvb_triangle{current, idx[0], idx[k], idx[k + 1]};
                    }
                    ++nt;
                }
            } else if (line.compare(s, 7, "usemtl ") == 0) {
                size_t b = line.find_first_not_of(" \t", s + 7);
                std::string name = b == std::string::npos ? std::string{} : line.substr(b);
                while (!name.empty() && (name.back() == ' ' || name.back() == '\t')) name.pop_back();
                auto it = std::find(materials.begin(), materials.end(), name);
                if (it == materials.end()) {
                    materials.push_back(name);
                    current = (uint32_t)materials.size() - 1;
                } else {
                    current = (uint32_t)(it - materials.begin());
                }
            }
        }
        (void)any_default;
        WVB_REQUIRE(nv > 0 && nt > 0, WVB_ERR_INVALID, "No geometry found in scene file.");
        std::string names;
        for (const auto& m : materials) {
            names += m;
            names += '\n';
        }
        if (material_names && material_names_length) {
            WVB_REQUIRE(*material_names_length >= names.size(), WVB_ERR_INVALID, "material name buffer too small");
            memcpy(material_names, names.data(), names.size());
        }
        if (material_names_length) *material_names_length = names.size();
        *num_vertices = nv;
        *num_triangles = nt;
    });
}

}  // extern "C"
