// The reference's boundary-filter helpers through the shim (fitted_boundary.h,
// arbitrary_magnitude_filter.h, stable.h): the call sequence of mesh.cpp:126-138 /
// boundary_test.cpp:311-334 for the three materials of the golden file. Host code
// only -- runs without a GPU. Prints the coefficients for the Python test.
#include <array>
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
    std::printf("LRS_SHIM_OK\n");
    return 0;
}
