// The reference's boundary-filter helpers through the shim (fitted_boundary.h,
// arbitrary_magnitude_filter.h, stable.h): the call sequence of mesh.cpp:126-138 /
// boundary_test.cpp:311-334 for the three materials of the golden file. Host code
// only -- runs without a GPU. Prints the coefficients for the Python test.
#include <array>
#include <cmath>
#include <cstdio>

#include "wayverb_b200/waveguide.hpp"

using namespace wayverb::waveguide;

int main() {
    const std::array<std::array<double, 8>, 3> materials{{
            {{0.08, 0.08, 0.2, 0.5, 0.4, 0.4, 0.36}},
            {{0.15, 0.15, 0.11, 0.1, 0.07, 0.06, 0.06}},
            {{0.02, 0.02, 0.03, 0.03, 0.03, 0.04, 0.07}},
    }};
    for (const auto& m : materials) {
        const auto refl = compute_reflectance_filter_coefficients(m, 8000.0);
        if (!is_stable(refl.a)) return 2;
        const auto imp = to_impedance_coefficients(refl);
        for (double v : refl.b) std::printf("%.17g ", v);
        for (double v : refl.a) std::printf("%.17g ", v);
        for (double v : imp.b) std::printf("%.17g ", v);
        for (double v : imp.a) std::printf("%.17g ", v);
        std::printf("\n");
    }
    // tests/fitted_boundary.cpp:31-40
    constexpr std::array<double, 5> centres{{0.2, 0.4, 0.6, 0.8, 1.0}};
    constexpr std::array<double, 5> amplitudes{{0, 1, 0.5, 1, 0}};
    const auto c = arbitrary_magnitude_filter<6>(make_frequency_domain_envelope(centres, amplitudes));
    if (!is_stable(c.a)) return 3;
    // tests/arbitrary_magnitude_filter.cpp:15: the empty envelope
    if (!is_stable(arbitrary_magnitude_filter<6>(frequency_domain_envelope{}).a)) return 4;
    const auto flat = to_flat_coefficients(0.1);
    if (flat.a[0] != 1.0 || !(flat.b[0] > 37.9 && flat.b[0] < 38.0)) return 5;
    // mesh_tests.cpp:53-70: index <-> locator <-> position round trips over a whole descriptor
    {
        mesh_descriptor d{};
        d.min_corner.s[0] = -1.03f; d.min_corner.s[1] = 0.5f; d.min_corner.s[2] = 2.2f;
        d.dimensions.s[0] = 17; d.dimensions.s[1] = 13; d.dimensions.s[2] = 11;
        d.spacing = 0.1f;
        const size_t lim = compute_num_nodes(d);
        if (lim != 17u * 13u * 11u) return 6;
        for (size_t i = 0; i != lim; ++i) {
            const auto loc = compute_locator(d, i);
            if (compute_index(d, loc[0], loc[1], loc[2]) != i) return 7;
            if (compute_index(d, compute_position(d, loc)) != i) return 8;
        }
        if (std::fabs(compute_sample_rate(d, 340.0) - 340.0 * std::sqrt(3.0) / double(d.spacing)) > 1e-9) return 9;
    }
    std::printf("LRS_SHIM_OK\n");
    return 0;
}
