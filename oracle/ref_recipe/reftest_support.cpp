// oracle/ref_recipe/reftest_support.cpp -- TEST INFRASTRUCTURE ONLY.
// What the reference's own unit tests need from a platform when they are run on the host stand-ins
// (build_tests.py): a default compute_context (core/src/cl/common.cpp picks an OpenCL device; there is
// none to pick), the host-compiled kernels registered by name before the first test runs, and a fit behind
// the IT++ stand-in's yulewalk: the library's own (wvb_lrs_arbitrary_magnitude_filter, host code; handed
// the 256-point grid as its envelope) -- the tests that reach it build meshes with fitted walls and look
// at the simulation, not at the coefficients.
#include "core/cl/common.h"
#include "core/scene_data_loader.h"
#include "itpp/signal/filter_design.h"

#include <fstream>
#include <sstream>

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

// core/src/scene_data_loader.cpp reads model files through assimp (not in the image). For the reference's
// tests that run on its own Wavefront models the loader is stood in for by the library's OBJ reader
// (wvb_obj_parse, host code): vertices, triangles with a material index each, material names.
namespace wayverb {
namespace core {
class scene_data_loader::impl final {
public:
    std::experimental::optional<scene_data> data;
};
scene_data_loader::scene_data_loader() : pimpl_{std::make_unique<impl>()} {}
scene_data_loader::scene_data_loader(const std::string& fpath) : pimpl_{std::make_unique<impl>()} { load(fpath); }
scene_data_loader::scene_data_loader(scene_data_loader&&) noexcept = default;
scene_data_loader& scene_data_loader::operator=(scene_data_loader&&) noexcept = default;
scene_data_loader::~scene_data_loader() noexcept = default;
void scene_data_loader::load(const std::string& f) {
    std::ifstream file{f, std::ios::binary};
    if (!file) throw std::runtime_error{"Couldn't load scene.\n" + f};
    std::stringstream text;
    text << file.rdbuf();
    const std::string raw = text.str();
    uint64_t nv = 0, nt = 0, nn = 0;
    if (wvb_obj_parse(raw.data(), raw.size(), nullptr, &nv, nullptr, &nt, nullptr, &nn) != WVB_OK)
        throw std::runtime_error{"No geometry found in scene file."};
    util::aligned::vector<cl_float3> vertices(nv);
    util::aligned::vector<triangle> triangles(nt);
    std::string names(nn, '\0');
    static_assert(sizeof(cl_float3) == sizeof(wvb_float3) && sizeof(triangle) == sizeof(wvb_triangle), "POD layouts");
    if (wvb_obj_parse(raw.data(), raw.size(), reinterpret_cast<wvb_float3*>(vertices.data()), &nv,
                      reinterpret_cast<wvb_triangle*>(triangles.data()), &nt, &names[0], &nn) != WVB_OK)
        throw std::runtime_error{"No geometry found in scene file."};
    util::aligned::vector<std::string> surfaces;
    std::stringstream lines{names};
    for (std::string line; std::getline(lines, line);) surfaces.push_back(line);
    pimpl_->data = make_scene_data(std::move(triangles), std::move(vertices), std::move(surfaces));
}
void scene_data_loader::save(const std::string&) const { throw std::runtime_error{"not in this build"}; }
void scene_data_loader::clear() { pimpl_->data = std::experimental::nullopt; }
std::string scene_data_loader::get_extensions() const { return "*.obj"; }
const std::experimental::optional<scene_data_loader::scene_data>& scene_data_loader::get_scene_data() const {
    return pimpl_->data;
}
}  // namespace core
}  // namespace wayverb

// audio_file's implementation writes through libsndfile (not in the image). The tests that call it only
// leave .wav files behind for a listener; nothing is asserted on them, so writing is a no-op here.
#include "audio_file/audio_file.h"
namespace audio_file {
template <typename T>
void write_interleaved(const char*, const T*, size_t, int, int, format, bit_depth) {}
template void write_interleaved<float>(const char*, const float*, size_t, int, int, format, bit_depth);
template void write_interleaved<double>(const char*, const double*, size_t, int, int, format, bit_depth);
}  // namespace audio_file

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
