// Holds the host-side closed-form functions of the shim and of the C ABI against the REFERENCE'S OWN
// source for them (waveguide/src/mesh_descriptor.cpp, config.cpp, calibration.h, fitted_boundary.h,
// stable.h, filters.cpp, raytracer/optimum_reflection_number.h, stochastic/finder.{h,cpp}), compiled from
// /root/reference into oracle/_ref/lib_ref.so by oracle/ref_recipe/build.py (driver:
// oracle/ref_recipe/hostmath_driver.inc). Host code only -- runs without a GPU. Equality is exact
// (==) everywhere: these are the same few floating-point operations or they are not.
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <random>

#include "wayverb_b200/raytracer.hpp"
#include "wayverb_b200/waveguide.hpp"

extern "C" {
size_t refk_hm_sizeof_mesh_descriptor();
size_t refk_hm_index_of_position(const float*, const int32_t*, float, const float*, int32_t*);
void refk_hm_node(const float*, const int32_t*, float, size_t, int32_t*, float*, uint32_t*);
void refk_hm_rates(float, double, double*);
double refk_hm_calibration_factor(float, double);
size_t refk_hm_reflection_number(double);
size_t refk_hm_reflection_number_of_scene(const uint32_t*, size_t, size_t, const float*, size_t);
float refk_hm_ray_energy(size_t, const float*, const float*, float);
void refk_hm_to_impedance(const double*, const double*, double*, double*);
void refk_hm_to_flat(double, double*, double*);
int refk_hm_is_stable(const double*);
}

using namespace wayverb;

#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) {                                                     \
            std::printf("FAILED line %d: %s\n", __LINE__, #cond);          \
            return 1;                                                      \
        }                                                                  \
    } while (0)

int main() {
    std::mt19937_64 engine{2024};
    std::uniform_real_distribution<float> unit{0.0f, 1.0f};
    CHECK(refk_hm_sizeof_mesh_descriptor() == sizeof(waveguide::mesh_descriptor));

    // ---- mesh_descriptor.cpp: index / locator / position / neighbours / sample rate -----------
    for (int trial = 0; trial < 40; ++trial) {
        waveguide::mesh_descriptor d{};
        const float mc[3] = {unit(engine) * 8 - 4, unit(engine) * 8 - 4, unit(engine) * 8 - 4};
        const int32_t dims[3] = {2 + int(unit(engine) * 30), 2 + int(unit(engine) * 30), 2 + int(unit(engine) * 30)};
        const float spacing = 0.01f + unit(engine) * 0.3f;
        for (int k = 0; k < 3; ++k) {
            d.min_corner.s[k] = mc[k];
            d.dimensions.s[k] = dims[k];
        }
        d.spacing = spacing;
        const size_t nodes = waveguide::compute_num_nodes(d);
        for (size_t i = 0; i < nodes; i += 1 + nodes / 997) {
            int32_t loc[3];
            float pos[3];
            uint32_t nb[6];
            refk_hm_node(mc, dims, spacing, i, loc, pos, nb);
            const auto l = waveguide::compute_locator(d, i);
            CHECK(l[0] == loc[0] && l[1] == loc[1] && l[2] == loc[2]);
            const auto p = waveguide::compute_position(d, l);
            CHECK(p.x == pos[0] && p.y == pos[1] && p.z == pos[2]);
            const auto n = waveguide::compute_neighbors(d, i);
            for (int k = 0; k < 6; ++k) CHECK(n[k] == nb[k]);
        }
        // nearest node of arbitrary positions inside the mesh, incl. points at the half-way marks
        for (int k = 0; k < 2000; ++k) {
            float pos[3];
            for (int a = 0; a < 3; ++a) {
                const float cell = unit(engine) * float(dims[a] - 1);
                pos[a] = mc[a] + (k % 5 == 0 ? std::floor(cell) + 0.5f : cell) * spacing;
            }
            int32_t loc[3];
            const size_t want = refk_hm_index_of_position(mc, dims, spacing, pos, loc);
            if (loc[0] < 0 || loc[1] < 0 || loc[2] < 0 || loc[0] >= dims[0] || loc[1] >= dims[1] || loc[2] >= dims[2])
                continue;
            CHECK(waveguide::compute_index(d, core::vec3{pos[0], pos[1], pos[2]}) == want);
        }
        for (double c : {300.0, 340.0, 343.2, 399.9}) {
            double r[4];
            refk_hm_rates(spacing, c, r);
            CHECK(waveguide::compute_sample_rate(d, c) == r[0]);
        }
        for (double z : {400.0, 413.3, 1.0}) {
            CHECK(waveguide::rectilinear_calibration_factor(d.spacing, z) == refk_hm_calibration_factor(spacing, z));
        }
    }

    // ---- optimum_reflection_number.h ------------------------------------------------------------
    for (int k = 0; k < 20000; ++k) {
        const double a = k < 100 ? 0.01 * (k + 1) * 0.99 : double(unit(engine)) * 0.98 + 0.005;
        CHECK(raytracer::compute_optimum_reflection_number(a) == refk_hm_reflection_number(a));
    }
    {
        // a scene whose least absorbent surface is not referenced by any triangle, and whose smallest
        // band is not the first one
        core::flattened_scene scene;
        scene.vertices.resize(3);
        scene.surfaces.resize(3);
        const float abs_bands[3][8] = {{0.3f, 0.2f, 0.25f, 0.4f, 0.5f, 0.6f, 0.7f, 0.8f},
                                       {0.02f, 0.02f, 0.02f, 0.02f, 0.02f, 0.02f, 0.02f, 0.02f},
                                       {0.5f, 0.45f, 0.4f, 0.35f, 0.31f, 0.33f, 0.6f, 0.7f}};
        float surfaces[3][16] = {};
        for (int s = 0; s < 3; ++s)
            for (int b = 0; b < 8; ++b) {
                scene.surfaces[s].absorption.s[b] = abs_bands[s][b];
                scene.surfaces[s].scattering.s[b] = 0.1f;
                surfaces[s][b] = abs_bands[s][b];
                surfaces[s][8 + b] = 0.1f;
            }
        scene.triangles = {{0, 0, 1, 2}, {2, 0, 1, 2}};
        const uint32_t tris[8] = {0, 0, 1, 2, 2, 0, 1, 2};
        const size_t want = refk_hm_reflection_number_of_scene(tris, 2, 3, &surfaces[0][0], 3);
        CHECK(raytracer::compute_optimum_reflection_number(scene) == want);
        CHECK(want == refk_hm_reflection_number(double(0.2f)));
    }

    // ---- compute_ray_energy ---------------------------------------------------------------------
    for (int k = 0; k < 20000; ++k) {
        const float s[3] = {unit(engine) * 10, unit(engine) * 10, unit(engine) * 10};
        float r[3] = {unit(engine) * 10, unit(engine) * 10, unit(engine) * 10};
        if (k % 50 == 0) {  // receiver within its own radius of the source
            for (int a = 0; a < 3; ++a) r[a] = s[a] + 0.01f * unit(engine);
        }
        const float radius = 0.05f + unit(engine) * 0.5f;
        const uint64_t rays = 1000 + uint64_t(unit(engine) * 2.0e6f);
        CHECK(wvb_rt_ray_energy(rays, s, r, radius) == refk_hm_ray_energy(size_t(rays), s, r, radius));
    }

    // ---- fitted_boundary.h / stable.h -----------------------------------------------------------
    for (int k = 0; k < 5000; ++k) {
        waveguide::coefficients_canonical c{};
        for (int i = 0; i < 7; ++i) {
            c.b[i] = double(unit(engine)) * 2 - 1;
            c.a[i] = (double(unit(engine)) * 2 - 1) * (k % 3 == 0 ? 0.2 : 1.0);
        }
        c.a[0] = 1;
        if (k % 97 == 0) {  // a[0] + ... such that the impedance denominator's a[0] is 0: no normalisation
            c.b[0] = 1;
        }
        double b7[7], a7[7];
        refk_hm_to_impedance(c.b, c.a, b7, a7);
        const auto got = waveguide::to_impedance_coefficients(c);
        for (int i = 0; i < 7; ++i) CHECK(got.b[i] == b7[i] && got.a[i] == a7[i]);
        CHECK(waveguide::is_stable(c.a) == (refk_hm_is_stable(c.a) != 0));
    }
    for (int k = 0; k < 1000; ++k) {
        const double absorption = k == 0 ? 0.0 : double(unit(engine)) * 0.999;
        double b7[7], a7[7];
        refk_hm_to_flat(absorption, b7, a7);
        const auto got = waveguide::to_flat_coefficients(absorption);
        for (int i = 0; i < 7; ++i) CHECK(got.b[i] == b7[i] && got.a[i] == a7[i]);
    }
    std::printf("HOSTMATH_PIN_OK\n");
    return 0;
}
