// oracle/ref_recipe/reftest_support.cpp -- TEST INFRASTRUCTURE ONLY.
// What the reference's own unit tests need from a platform when they are run on the host stand-ins
// (build_tests.py): a default compute_context (core/src/cl/common.cpp picks an OpenCL device; there is
// none to pick), the host-compiled kernels registered by name before the first test runs, and a fit behind
// the IT++ stand-in's yulewalk: the library's own (wvb_lrs_arbitrary_magnitude_filter, host code; handed
// the 256-point grid as its envelope) -- the tests that reach it build meshes with fitted walls and look
// at the simulation, not at the coefficients.
#include "core/cl/common.h"
#include "itpp/signal/filter_design.h"

#include "wvb200.h"

extern "C" {
void refk_register_waveguide_kernels();
void refk_register_ray_kernels_c();
}

namespace wayverb {
namespace core {
compute_context::compute_context() {}
compute_context::compute_context(device_type) {}
}  // namespace core
}  // namespace wayverb

namespace {
void fit_with_the_library(int order, int points, const double* f, const double* m, double* b, double* a) {
    wvb_coefficients_canonical c{};
    if (order != 6 || wvb_lrs_arbitrary_magnitude_filter(f, m, uint32_t(points), &c) != WVB_OK) {
        throw std::runtime_error{"the library's filter fit failed"};
    }
    for (int k = 0; k <= order; ++k) {
        b[k] = c.b[k];
        a[k] = c.a[k];
    }
}
const struct registrar {
    registrar() {
        refk_register_waveguide_kernels();
        refk_register_ray_kernels_c();
        refk_yulewalk_hook() = fit_with_the_library;
    }
} registrar_instance;
}  // namespace
