// The shim as an OVERLAY on the reference tree, not a parallel universe.
//
// Built with  -DWVB_WITH_REFERENCE_HEADERS -I include/compat -I include
//             -I /root/reference/src/waveguide/include -I /root/reference/src/utilities/include
// together with the reference's own src/waveguide/src/postprocessor/node.cpp. What is compiled
// UNMODIFIED from /root/reference:
//   waveguide/preprocessor/hard_source.h   (hard_source.h:9-35)
//   waveguide/preprocessor/soft_source.h   (soft_source.h:9-39)
//   waveguide/postprocessor/node.h + src/postprocessor/node.cpp
//   utilities/map_to_vector.h (pulled in by node.cpp)
// They resolve "core/cl/common.h", "core/cl/include.h", "utilities/aligned/vector.h" to the
// forwarding headers of include/compat, i.e. to libwvb200.so; waveguide.hpp then does not
// define its own copies. waveguide::run is driven with those reference-made processors.
//
// Second half: raytracer::run with the reference's parameter TYPES (raytracer.h:188-201) -- a
// stand-in for core::voxelised_scene_data<cl_float3, surface<8>> (same member interface;
// the real one needs glm + the octree) and a glm-like vec3 -- through the `flatten`
// customisation point, against the flattened_scene overload.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <vector>

#include "wayverb_b200/raytracer.hpp"
#include "wayverb_b200/waveguide.hpp"

using namespace wayverb;

// prove the processors in use are the reference's: its node returns float, the shim's own
// returns double (and is not declared in overlay mode)
static_assert(std::is_same<waveguide::postprocessor::node::return_type, float>::value,
              "postprocessor::node is not the reference's");

static int waveguide_half(int steps, double b0, bool soft) {
    const core::compute_context cc{};
    waveguide::coefficients_canonical c{};
    c.b[0] = b0;
    c.a[0] = 1.0;
    const auto m = waveguide::make_cuboid_mesh(30, 24, 20, 0.05f, c);
    const auto src = waveguide::compute_index(m.get_descriptor(), 14, 11, 9);
    const auto rcv = waveguide::compute_index(m.get_descriptor(), 19, 13, 8);
    std::vector<float> sig(size_t(steps), 0.0f);  // the engine's input signal is float (canonical.h:55-63)
    sig[0] = 1.0f;
    if (soft) sig[2] = -0.5f;
    core::callback_accumulator<waveguide::postprocessor::node> out{rcv};
    const std::atomic_bool keep_going{true};
    size_t done;
    if (soft) {
        done = waveguide::run(cc, m, waveguide::preprocessor::make_soft_source(src, sig.begin(), sig.end()),
                              [&](auto& q, const auto& b, auto step) { out(q, b, step); }, keep_going);
    } else {
        done = waveguide::run(cc, m, waveguide::preprocessor::make_hard_source(src, sig.begin(), sig.end()),
                              [&](auto& q, const auto& b, auto step) { out(q, b, step); }, keep_going);
    }
    if (done != size_t(steps) || out.get_output().size() != size_t(steps)) return 3;
    if (out.get_callback().get_output_node() != rcv) return 4;
    std::printf("# src %zu rcv %zu\n", src, rcv);
    for (float v : out.get_output()) std::printf("%.9g\n", v);
    return 0;
}

// ---- stand-ins with the reference's interfaces -------------------------------------------------
namespace ref_like {
struct vec3 {  // glm::vec3
    float x, y, z;
};
struct uvec3 {  // glm::uvec3 (indexing::index_t<3>)
    unsigned x, y, z;
    uvec3(size_t a, size_t b, size_t c) : x(unsigned(a)), y(unsigned(b)), z(unsigned(c)) {}
};
struct range3 {  // util::range<glm::vec3>
    vec3 mn, mx;
    vec3 get_min() const { return mn; }
    vec3 get_max() const { return mx; }
};
struct triangle {
    cl_uint surface, v0, v1, v2;
};
struct surface8 {
    core::bands_type absorption, scattering;
};
struct scene_data {  // generic_scene_data<cl_float3, surface<8>>
    std::vector<triangle> triangles;
    std::vector<cl_float3> vertices;
    std::vector<surface8> surfaces;
    const std::vector<triangle>& get_triangles() const { return triangles; }
    const std::vector<cl_float3>& get_vertices() const { return vertices; }
    const std::vector<surface8>& get_surfaces() const { return surfaces; }
};
struct voxel_collection {  // voxel_collection<3>
    range3 aabb;
    size_t side;
    std::vector<std::vector<size_t>> cells;  // [x][y][z] flattened
    range3 get_aabb() const { return aabb; }
    size_t get_side() const { return side; }
    const std::vector<size_t>& get_voxel(uvec3 i) const { return cells[(i.x * side + i.y) * side + i.z]; }
};
struct voxelised_scene_data {
    scene_data scene;
    voxel_collection voxels;
    const scene_data& get_scene_data() const { return scene; }
    const voxel_collection& get_voxels() const { return voxels; }
    static uvec3 voxel_index_type() { return uvec3(0, 0, 0); }
};
}  // namespace ref_like

static int raytracer_half() {
    // a shoebox, voxelised by the library like the engine does (depth 3 here)
    ref_like::voxelised_scene_data v;
    const float sx = 5.56f, sy = 3.97f, sz = 2.81f;
    const float X[2] = {0, sx}, Y[2] = {0, sy}, Z[2] = {0, sz};
    for (int i = 0; i < 8; ++i) v.scene.vertices.push_back(cl_float3{{X[i & 1], Y[(i >> 1) & 1], Z[(i >> 2) & 1], 0}});
    const unsigned q[6][4] = {{0, 1, 3, 2}, {4, 6, 7, 5}, {0, 4, 5, 1}, {2, 3, 7, 6}, {0, 2, 6, 4}, {1, 5, 7, 3}};
    for (auto& f : q) {
        v.scene.triangles.push_back({0, f[0], f[1], f[2]});
        v.scene.triangles.push_back({0, f[0], f[2], f[3]});
    }
    ref_like::surface8 surf{};
    for (int b = 0; b < 8; ++b) { surf.absorption.s[b] = 0.1f; surf.scattering.s[b] = 0.1f; }
    v.scene.surfaces.push_back(surf);
    const unsigned depth = 3;
    float lo[3], hi[3];
    uint64_t n = 0;
    static_assert(sizeof(ref_like::triangle) == sizeof(wvb_triangle), "triangle layout");
    const auto* verts = reinterpret_cast<const wvb_float3*>(v.scene.vertices.data());
    const auto* tris = reinterpret_cast<const wvb_triangle*>(v.scene.triangles.data());
    if (wvb_voxelise(verts, 8, tris, 12, depth, 0.1f, lo, hi, nullptr, 0, &n) != WVB_OK) return 30;
    std::vector<uint32_t> flat(n);
    if (wvb_voxelise(verts, 8, tris, 12, depth, 0.1f, lo, hi, flat.data(), n, &n) != WVB_OK) return 31;
    v.voxels.aabb = {{lo[0], lo[1], lo[2]}, {hi[0], hi[1], hi[2]}};
    v.voxels.side = 1u << depth;
    v.voxels.cells.resize(v.voxels.side * v.voxels.side * v.voxels.side);
    for (size_t c = 0; c < v.voxels.cells.size(); ++c) {
        const uint32_t o = flat[c];
        v.voxels.cells[c].assign(flat.begin() + o + 1, flat.begin() + o + 1 + flat[o]);
    }
    // the customisation point reproduces what the library flattened
    const auto fs = raytracer::flatten(v);
    if (fs.voxel_index.size() != flat.size() || std::memcmp(fs.voxel_index.data(), flat.data(), flat.size() * 4)) return 32;
    if (fs.side != 8 || fs.triangles.size() != 12 || fs.vertices.size() != 8 || fs.surfaces.size() != 1) return 33;

    const core::compute_context cc{};
    const ref_like::vec3 source{1, 1, 1}, receiver{2, 3, 1.5f};
    std::vector<core::vec3> dirs;
    for (int i = 0; i < 20000; ++i) {
        const float z = -1 + 2 * ((i * 0.6180339887f) - std::floor(i * 0.6180339887f));
        const float th = 2.39996323f * float(i), t = std::sqrt(1 - z * z);
        dirs.push_back({t * std::cos(th), z, t * std::sin(th)});
    }
    const auto callbacks = [&] {
        return std::make_tuple(raytracer::reflection_processor::make_image_source{4},
                               raytracer::reflection_processor::make_stochastic_histogram{dirs.size(), 5, 0.1f, 1000.0f});
    };
    // raytracer.h:188-201 spelled with the caller's own types ...
    auto a = raytracer::run(dirs.begin(), dirs.end(), cc, v, source, receiver, core::environment{}, true,
                            [](auto, auto) {}, callbacks(), 99);
    // ... and with the already-flattened scene
    auto b = raytracer::run(dirs.begin(), dirs.end(), cc, fs, core::vec3{1, 1, 1}, core::vec3{2, 3, 1.5f},
                            core::environment{}, true, [](auto, auto) {}, callbacks(), 99);
    if (!a || !b) return 34;
    const auto& ia = std::get<0>(*a);
    const auto& ib = std::get<0>(*b);
    if (ia.size() != ib.size() || ia.size() < 7 || std::memcmp(ia.data(), ib.data(), ia.size() * sizeof(ia[0]))) return 35;
    const auto& ha = std::get<1>(*a).histogram;
    const auto& hb = std::get<1>(*b).histogram;
    if (ha.size() != hb.size() || ha.empty()) return 36;
    double ea = 0, eb = 0;
    for (size_t i = 0; i < ha.size(); ++i) {
        for (int k = 0; k < 8; ++k) {
            ea += ha[i].s[k];
            eb += hb[i].s[k];
        }
    }
    if (!(ea > 0) || std::fabs(ea - eb) > 1e-6 * eb) return 37;
    std::printf("OVERLAY_RT_OK impulses=%zu energy=%g\n", ia.size(), ea);
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && !std::strcmp(argv[1], "rt")) return raytracer_half();
    const int steps = argc > 1 ? std::atoi(argv[1]) : 60;
    const double b0 = argc > 2 ? std::atof(argv[2]) : 0.8;
    const bool soft = argc > 3 && !std::strcmp(argv[3], "soft");
    return waveguide_half(steps, b0, soft);
}
