"""oracle/ref_recipe/build.py -- builds oracle/_ref/lib_ref.so from the REFERENCE'S OWN kernel source.

TEST INFRASTRUCTURE ONLY. Nothing in wayverb_b200/ may import or load what this produces.

The reference's device code is OpenCL-C held in C++ raw-string literals. Its own build cannot run in
this image (glm, OpenCL-CLHPP, FFTW, IT++, libsamplerate ... are fetched at configure time), but the
kernel strings compile for the host, and so do its host sources behind stand-ins for those
dependencies (HOST_UNITS below, hoststubs/, hostcl/). For the kernels this recipe

  1. reads the raw-string literals out of the files under /root/reference where they lie
     (never copied into the repository: every output goes to oracle/_ref/, which is
     git-ignored but travels to the GPU box with the snapshot),
  2. joins them in the order the reference's `program::program` constructors join them
     (src/waveguide/src/program.cpp:533-555, src/raytracer/src/program.cpp:157-174,
     src/raytracer/src/stochastic/program.cpp:156-175, src/waveguide/src/mesh_setup_program.cpp:176-194,
     src/waveguide/src/boundary_coefficient_program.cpp:487-507),
  3. applies the MECHANICAL rewrites listed in `rewrite()` -- syntax only, each one documented --
  4. and compiles the result with g++ behind cl_prelude.hpp (an OpenCL-C emulation layer) and a
     per-program driver (*.inc: extern "C" launchers that loop over the NDRange).

Build flags: -O2 -ffp-contract=off (no FMA contraction), so what executes is the operation order
of the reference source.

The host units are plain translation units that #include the reference's .cpp files by path (whole
files; single functions only where the rest of a file needs something that is not stood in for), then a
driver (*_driver.inc) with the extern "C" entry points oracle/refk.py binds.

Usage: python oracle/ref_recipe/build.py [--force]        (WVB_REFERENCE_ROOT=<checkout> for another tree)
"""
from __future__ import annotations

import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.normpath(os.path.join(HERE, "..", "_ref"))
LIB = os.path.join(OUT, "lib_ref.so")
REF = os.environ.get("WVB_REFERENCE_ROOT", "/root/reference")

RAW = re.compile(r'R"\((.*?)\)"', re.S)


def _read(rel: str) -> str:
    with open(os.path.join(REF, rel), "r", encoding="utf-8", errors="replace") as f:
        return f.read()


def raw_strings(rel: str) -> list[str]:
    return RAW.findall(_read(rel))


def representation(rel: str, type_name: str) -> str:
    """The `cl_representation<...type_name>` raw string of a header."""
    text = _read(rel)
    m = re.search(r"cl_representation<\s*(?:\w+::)*" + re.escape(type_name) + r"\s*(?:<[^>]*>)?\s*>\s*final\s*\{(.*?)\};",
                  text, re.S)
    if not m:
        raise RuntimeError("no cl_representation<%s> in %s" % (type_name, rel))
    r = RAW.search(m.group(1))
    if not r:
        raise RuntimeError("cl_representation<%s> in %s holds no raw string" % (type_name, rel))
    return r.group(1)


def filter_struct_representations() -> dict[str, str]:
    """src/waveguide/src/cl/filter_structs.cpp builds four representation strings by
    concatenating raw strings with std::to_string(<order expression>). Evaluate those
    expressions from the constants of include/waveguide/cl/filter_structs.h:9-10,59-66."""
    hdr = _read("src/waveguide/include/waveguide/cl/filter_structs.h")
    order = int(re.search(r"biquad_order\{(\d+)\}", hdr).group(1))
    sections = int(re.search(r"biquad_sections\{(\d+)\}", hdr).group(1))
    orders = {"memory_biquad": order, "coefficients_biquad": order,
              "memory_canonical": order * sections, "coefficients_canonical": order * sections}
    text = _read("src/waveguide/src/cl/filter_structs.cpp")
    out = {}
    for m in re.finditer(r"cl_representation<waveguide::(\w+)>::value\s*\{(.*?)\};", text, re.S):
        name, body = m.group(1), m.group(2)

        def number(expr):
            e = re.sub(r"waveguide::(\w+)::order", lambda k: str(orders[k.group(1)]), expr)
            assert re.fullmatch(r"[\d\s+\-*()]+", e), e
            return str(eval(e))  # digits and + - * ( ) only

        pieces = []
        for piece in re.finditer(r'R"\((.*?)\)"|std::to_string\(((?:[^()]|\([^()]*\))*)\)', body, re.S):
            pieces.append(piece.group(1) if piece.group(1) is not None else number(piece.group(2)))
        out[name] = "".join(pieces)
    assert set(out) == set(orders), out.keys()
    return out, order, sections


# ---- the mechanical rewrites -------------------------------------------------------------------
VECTOR_TYPES = ("int3", "float3", "uint3", "float8", "bands_type")


def rewrite(src: str, fp64: bool) -> str:
    """Syntax-only rewrites from OpenCL-C to what g++ parses. None changes an operation.

    R1  `(int3)(a, b, c)` / `(float3)(v)` vector literals      -> constructor calls `int3(a, b, c)`
    R2  C99 compound literals `(T){...}` of the program's own   -> C++ brace initialisation `T{...}`
        struct types (names collected from its typedefs)
    R3  fp64 build only: the `f` suffix of floating literals is  -> `1.0f` becomes `1.0`
        dropped, and `float` is #defined to `double` by the
        translation unit's header (see `unit()`), which is how
        an fp64 build of the same source reads.
    """
    for t in VECTOR_TYPES:
        src = re.sub(r"\(\s*%s\s*\)\s*\(" % t, t + "(", src)                      # R1
    types = set(re.findall(r"\}\s*(\w+)\s*;", src)) | set(re.findall(r"typedef\s+\w+\s+(\w+)\s*;", src))
    types -= {"return", "break"}
    control = {"switch", "if", "while", "for"}   # `switch (boundary_type) {` is not a literal

    def literal(m):
        return m.group(0) if m.group(1) in control else (m.group(1) or "") + m.group(2) + m.group(3) + "{"

    for t in sorted(types, key=len, reverse=True):
        src = re.sub(r"(\b\w+\b)?(\s*)\(\s*(%s)\s*\)\s*\{" % re.escape(t), literal, src)   # R2
    if fp64:
        src = re.sub(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?)f\b", r"\1", src)  # R3
    return src


def unit(name: str, parts: list[str], driver: str, fp64: bool = False, prefix: str = "") -> str:
    """One translation unit: prelude, the program's sources in the reference's order, the driver."""
    body = "\n".join(parts)
    body = rewrite(body, fp64)
    head = ['// GENERATED by oracle/ref_recipe/build.py from /root/reference -- do not commit.',
            '#define CLC_REAL %s' % ("double" if fp64 else "float"),
            '#include "%s"' % os.path.join(HERE, "cl_prelude.hpp"),
            '#define REFK(n) refk_%s##n' % prefix,
            'namespace clc { namespace %s {' % name]
    if fp64:
        head.append("#define float double")
    tail = ['#include "%s"' % os.path.join(HERE, driver), "} }"]
    return "\n".join(head) + "\n" + body + "\n" + "\n".join(tail) + "\n"


def waveguide_program(fp64: bool) -> str:
    """src/waveguide/src/program.cpp:533-555 (constructor order)."""
    fs, order, sections = filter_struct_representations()
    W = "src/waveguide/include/waveguide/"
    # filter_constants is assembled from std::to_string in src/waveguide/src/cl/filters.cpp:8-11
    filter_constants = ("#define BIQUAD_SECTIONS %d\n#define BIQUAD_ORDER %d\n#define CANONICAL_FILTER_ORDER %d\n"
                        % (sections, order, sections * 2))
    filters = raw_strings("src/waveguide/src/cl/filters.cpp")
    assert len(filters) == 1
    utils = raw_strings("src/waveguide/src/cl/utils.cpp")
    assert len(utils) == 1
    prog = raw_strings("src/waveguide/src/program.cpp")
    assert len(prog) == 1
    parts = [
        filter_constants,
        representation(W + "cl/filter_structs.h", "filt_real"),
        fs["memory_biquad"], fs["coefficients_biquad"], fs["memory_canonical"], fs["coefficients_canonical"],
        representation(W + "cl/filter_structs.h", "biquad_memory_array"),
        representation(W + "cl/filter_structs.h", "biquad_coefficients_array"),
        representation(W + "mesh_descriptor.h", "mesh_descriptor"),
        representation(W + "cl/structs.h", "error_code"),
        representation(W + "cl/structs.h", "condensed_node"),
        representation(W + "cl/structs.h", "boundary_data"),
        representation(W + "cl/structs.h", "boundary_data_array_1"),
        representation(W + "cl/structs.h", "boundary_data_array_2"),
        representation(W + "cl/structs.h", "boundary_data_array_3"),
        representation(W + "cl/utils.h", "boundary_type"),
        filters[0], utils[0], prog[0],
    ]
    return unit("wg_f64" if fp64 else "wg_f32", parts, "wg_driver.inc", fp64, "f64_" if fp64 else "f32_")


C = "src/core/include/core/cl/"
R = "src/raytracer/include/raytracer/cl/"


def _core_structs() -> list[str]:
    return [
        representation(C + "scene_structs.h", "bands_type"),
        representation(C + "scene_structs.h", "surface"),
        representation(C + "triangle.h", "triangle"),
        representation(C + "scene_structs.h", "triangle_verts"),
        representation(C + "voxel_structs.h", "aabb"),
        representation(C + "geometry_structs.h", "ray"),
        representation(C + "geometry_structs.h", "triangle_inter"),
        representation(C + "geometry_structs.h", "intersection"),
    ]


def _one(rel: str) -> str:
    s = raw_strings(rel)
    assert len(s) == 1, (rel, len(s))
    return s[0]


def raytracer_program() -> str:
    """src/raytracer/src/program.cpp:157-174."""
    parts = _core_structs() + [
        representation(R + "structs.h", "reflection"),
        representation(R + "structs.h", "impulse"),
        _one("src/core/src/cl/geometry.cpp"), _one("src/core/src/cl/voxel.cpp"),
        _one("src/raytracer/src/cl/brdf.cpp"), _one("src/raytracer/src/program.cpp"),
    ]
    return unit("rt", parts, "rt_driver.inc")


def stochastic_program() -> str:
    """src/raytracer/src/stochastic/program.cpp:156-175."""
    parts = _core_structs() + [
        representation(R + "structs.h", "impulse"),
        representation(R + "structs.h", "reflection"),
        representation(R + "structs.h", "stochastic_path_info"),
        _one("src/core/src/cl/geometry.cpp"), _one("src/core/src/cl/voxel.cpp"),
        _one("src/raytracer/src/cl/brdf.cpp"), _one("src/raytracer/src/stochastic/program.cpp"),
    ]
    return unit("stoch", parts, "stoch_driver.inc")


def mesh_setup_program() -> str:
    """src/waveguide/src/mesh_setup_program.cpp:176-194."""
    W = "src/waveguide/include/waveguide/"
    parts = _core_structs() + [
        representation(W + "cl/utils.h", "boundary_type"),
        representation(W + "cl/structs.h", "condensed_node"),
        representation(W + "mesh_descriptor.h", "mesh_descriptor"),
        _one("src/core/src/cl/geometry.cpp"), _one("src/core/src/cl/voxel.cpp"),
        _one("src/waveguide/src/cl/utils.cpp"), _one("src/waveguide/src/mesh_setup_program.cpp"),
    ]
    return unit("mesh_setup", parts, "mesh_setup_driver.inc")


def boundary_coefficient_program() -> str:
    """src/waveguide/src/boundary_coefficient_program.cpp:487-507."""
    W = "src/waveguide/include/waveguide/"
    parts = [
        representation(W + "mesh_descriptor.h", "mesh_descriptor"),
        representation(W + "cl/utils.h", "boundary_type"),
        representation(W + "cl/structs.h", "condensed_node"),
        representation(W + "cl/boundary_index_array.h", "boundary_index_array_1"),
        representation(W + "cl/boundary_index_array.h", "boundary_index_array_2"),
        representation(W + "cl/boundary_index_array.h", "boundary_index_array_3"),
        representation(C + "voxel_structs.h", "aabb"),
        representation(C + "geometry_structs.h", "ray"),
        representation(C + "geometry_structs.h", "triangle_inter"),
        representation(C + "geometry_structs.h", "intersection"),
        representation(C + "scene_structs.h", "triangle_verts"),
        representation(C + "triangle.h", "triangle"),
        _one("src/core/src/cl/geometry.cpp"), _one("src/core/src/cl/voxel.cpp"),
        _one("src/waveguide/src/cl/utils.cpp"), _one("src/waveguide/src/boundary_coefficient_program.cpp"),
    ]
    return unit("bcoef", parts, "bcoef_driver.inc")


def function_source(rel: str, signature_regex: str) -> str:
    """The text of one function definition of a reference .cpp file, from its signature to the
    matching closing brace (taken where it lies, like the kernel strings)."""
    text = _read(rel)
    m = re.search(signature_regex, text)
    if not m:
        raise RuntimeError("no %s in %s" % (signature_regex, rel))
    i = text.index("{", m.end() - 1)
    depth, j = 0, i
    while True:
        if text[j] == "{":
            depth += 1
        elif text[j] == "}":
            depth -= 1
            if depth == 0:
                break
        j += 1
    return text[m.start():j + 1]


def host_scene_program() -> str:
    """The reference's HOST scene code, whole files #included where they lie behind the GLM stand-in:
    the octree voxeliser (tri_cube_intersection.cpp, box.cpp, ndim_tree.h, voxel_collection.h/.cpp,
    voxelised_scene_data.h), the CPU ray casts (geometric.cpp, triangle_vec.cpp, the voxel walk of
    voxel_collection.cpp:41-124) and the image-source stage (raytracer/src/image_source/{tree,
    postprocess_branches,exact}.cpp, pressure_intensity.cpp). exact.h prints its shell counts to
    std::cout; nothing else is said or changed."""
    src = os.path.join(REF, "src")
    files = [
        ("core", "src", "geo", "tri_cube_intersection.cpp"), ("core", "src", "geo", "triangle_vec.cpp"),
        ("core", "src", "geo", "geometric.cpp"), ("core", "src", "geo", "box.cpp"),
        ("core", "src", "spatial_division", "voxel_collection.cpp"), ("core", "src", "pressure_intensity.cpp"),
        ("raytracer", "src", "image_source", "tree.cpp"), ("raytracer", "src", "image_source", "postprocess_branches.cpp"),
        ("raytracer", "src", "image_source", "exact.cpp"), ("core", "src", "az_el.cpp"),
    ]
    sh = "src/raytracer/include/raytracer/reflection_processor/stochastic_histogram.h"
    # the three binning pieces of stochastic_histogram.h:17-39 (the rest of that header runs the OpenCL finder)
    sums = [function_source(sh, r"template <typename T, typename U, typename Alloc>\s*void energy_histogram_sum\("),
            function_source(sh, r"template <typename T, typename U, typename Alloc, size_t Az, size_t El>\s*"
                                r"void energy_histogram_sum\("),
            function_source(sh, r"struct energy_histogram_sum_functor final \{") + ";"]
    return "\n".join([
        "// GENERATED by oracle/ref_recipe/build.py from /root/reference -- do not commit.",
        "#include <algorithm>", "#include <array>", "#include <cmath>", "#include <cstring>", "#include <functional>",
        "#include <future>", "#include <iostream>", "#include <memory>", "#include <numeric>", "#include <random>",
        "#include <stdexcept>", "#include <vector>", "#include <experimental/optional>",
    ] + ['#include "%s"' % os.path.join(src, *f) for f in files] + [
        '#include "raytracer/image_source/get_direct.h"',
        '#include "raytracer/image_source/reflection_path_builder.h"',
        '#include "raytracer/histogram.h"', '#include "raytracer/stochastic/postprocessing.h"',
        '#include "core/vector_look_up_table.h"',
        "namespace wayverb { namespace raytracer { namespace reflection_processor {"] + sums + ["} } }",
        '#include "%s"' % os.path.join(HERE, "scene_driver.inc"),
    ]) + "\n"


def host_math_program() -> str:
    """The reference's small closed-form HOST functions that the product's host side restates: whole
    files where they compile (mesh_descriptor.cpp, config.cpp, frequency_domain_envelope.cpp, filters.cpp,
    fitted_boundary.h & co. over an IT++ stand-in that forwards the fit)."""
    src = os.path.join(REF, "src")
    wg = os.path.join(src, "waveguide", "src")
    return "\n".join([
        "// GENERATED by oracle/ref_recipe/build.py from /root/reference -- do not commit.",
        "#include <algorithm>", "#include <array>", "#include <cmath>", "#include <cstring>", "#include <functional>",
        "#include <iostream>", "#include <memory>", "#include <numeric>", "#include <stdexcept>", "#include <vector>",
        '#include "%s"' % os.path.join(wg, "mesh_descriptor.cpp"),
        '#include "%s"' % os.path.join(wg, "config.cpp"),
        '#include "%s"' % os.path.join(wg, "frequency_domain_envelope.cpp"),
        '#include "%s"' % os.path.join(wg, "filters.cpp"),
        '#include "waveguide/fitted_boundary.h"', '#include "waveguide/calibration.h"',
        '#include "raytracer/optimum_reflection_number.h"', '#include "core/cl/scene_structs.h"',
        # compute_ray_energy: stochastic/finder.cpp is compiled whole in ref_hostray.cpp (it needs the cl.hpp stand-in)
        "namespace wayverb { namespace raytracer { namespace stochastic {",
        "float compute_ray_energy(size_t, const glm::vec3&, const glm::vec3&, float);", "} } }",
        '#include "%s"' % os.path.join(HERE, "hostmath_driver.inc"),
    ]) + "\n"


def single_file_unit(*rel) -> "callable":
    """A translation unit that is one reference .cpp file, #included where it lies (files that each
    define a file-local `source` string cannot share a unit)."""
    def make() -> str:
        return "\n".join([
            "// GENERATED by oracle/ref_recipe/build.py from /root/reference -- do not commit.",
            "#include <algorithm>", "#include <array>", "#include <cmath>", "#include <cstring>", "#include <functional>",
            "#include <iostream>", "#include <memory>", "#include <numeric>", "#include <stdexcept>", "#include <vector>",
            '#include "%s"' % os.path.join(REF, "src", *rel),
        ]) + "\n"
    return make


def host_run_program() -> str:
    """The reference's waveguide HOST code -- the waveguide::run template and canonical.h, the program
    classes, setup.cpp, the stock processors, and mesh construction (mesh.cpp, boundary_adjust.cpp,
    boundary_coefficient_finder.cpp with their programs) -- whole files #included where they lie over
    the host-memory cl.hpp stand-in (hostcl/)."""
    src = os.path.join(REF, "src")
    wg = os.path.join(src, "waveguide", "src")
    files = [("cl", "filter_structs.cpp"), ("cl", "filters.cpp"), ("cl", "utils.cpp"), ("program.cpp",), ("setup.cpp",),
             ("postprocessor", "node.cpp"), ("postprocessor", "directional_receiver.cpp"),
             ("preprocessor", "gaussian.cpp"), ("boundary_coefficient_finder.cpp",), ("boundary_adjust.cpp",),
             ("mesh.cpp",)]
    core_cl = [os.path.join(src, "core", "src", "cl", f) for f in ("geometry.cpp", "voxel.cpp")]
    return "\n".join([
        "// GENERATED by oracle/ref_recipe/build.py from /root/reference -- do not commit.",
        "#include <algorithm>", "#include <array>", "#include <cmath>", "#include <cstring>", "#include <functional>",
        "#include <iostream>", "#include <memory>", "#include <numeric>", "#include <stdexcept>", "#include <vector>",
    ] + ['#include "%s"' % f for f in core_cl] + ['#include "%s"' % os.path.join(wg, *f) for f in files] + [
        '#include "waveguide/waveguide.h"', '#include "waveguide/canonical.h"',
        '#include "waveguide/preprocessor/hard_source.h"',
        '#include "waveguide/preprocessor/soft_source.h"', '#include "core/callback_accumulator.h"',
    ] + [
        '#include "%s"' % os.path.join(HERE, "hostrun_driver.inc"),
    ]) + "\n"


def host_ray_program() -> str:
    """The reference's ray HOST code -- the raytracer::run template, reflector.cpp, stochastic/finder.cpp,
    the reflection processors, canonical.cpp -- whole files #included where they lie over the host-memory
    cl.hpp stand-in. While reflector.cpp is read std::random_device is spelled as a device that counts
    up from a chosen seed (hostray_driver.inc says why); nothing else changes."""
    src = os.path.join(REF, "src")
    rt = os.path.join(src, "raytracer", "src")
    files = [("stochastic", "finder.cpp"), ("reflection_processor", "image_source.cpp"),
             ("reflection_processor", "stochastic_histogram.cpp"), ("reflection_processor", "visual.cpp"),
             ("canonical.cpp",)]
    return "\n".join([
        "// GENERATED by oracle/ref_recipe/build.py from /root/reference -- do not commit.",
        "#include <algorithm>", "#include <array>", "#include <cmath>", "#include <cstring>", "#include <functional>",
        "#include <future>", "#include <iostream>", "#include <memory>", "#include <numeric>", "#include <random>",
        "#include <stdexcept>", "#include <vector>",
        "static unsigned refk_ray_seed_value = 1;",
        "namespace std { struct refk_counting_device { unsigned operator()() const { return refk_ray_seed_value++; } }; }",
        "#define random_device refk_counting_device",
        '#include "%s"' % os.path.join(rt, "reflector.cpp"),
        "#undef random_device",
        '#include "%s"' % os.path.join(src, "core", "src", "azimuth_elevation.cpp"),
    ] + ['#include "%s"' % os.path.join(rt, *f) for f in files] + [
        '#include "raytracer/raytracer.h"',
        '#include "%s"' % os.path.join(HERE, "hostray_driver.inc"),
    ]) + "\n"


def host_pp_program() -> str:
    """The reference's HOST post-processing code: raytracer/src/stochastic/postprocessing.cpp and the
    whole frequency_domain library as files (#included where they lie) over the FFTW stand-in of
    hoststubs/fftw3.h, plus crossover_filter of combined/postprocess.h:33-60 as a single function
    (the rest of that header needs the waveguide / raytracer engines). std::random_device is spelled
    as a device that returns a chosen seed while postprocessing.cpp is read, nothing else changes."""
    src = os.path.join(REF, "src")
    crossover = function_source("src/combined/include/combined/postprocess.h",
                                r"template <typename LoIt, typename HiIt>\s*auto crossover_filter\(")
    fd = os.path.join(src, "frequency_domain", "src")
    return "\n".join([
        "// GENERATED by oracle/ref_recipe/build.py from /root/reference -- do not commit.",
        "#include <algorithm>", "#include <array>", "#include <cmath>", "#include <complex>", "#include <cstring>",
        "#include <functional>", "#include <iostream>", "#include <memory>", "#include <numeric>", "#include <random>",
        "#include <stdexcept>", "#include <vector>",
        "static unsigned refk_pp_seed_value = 1;",
        "namespace std { struct refk_seeded_device { unsigned operator()() const { return refk_pp_seed_value; } }; }",
        "#define random_device refk_seeded_device",
        '#include "%s"' % os.path.join(src, "raytracer", "src", "stochastic", "postprocessing.cpp"),
        "#undef random_device",
        '#include "%s"' % os.path.join(fd, "envelope.cpp"),
        '#include "%s"' % os.path.join(fd, "traits.cpp"),
        '#include "%s"' % os.path.join(fd, "buffer.cpp"),
        '#include "%s"' % os.path.join(fd, "plan.cpp"),
        '#include "%s"' % os.path.join(fd, "filter.cpp"),
        '#include "utilities/aligned/vector.h"',
        '#include "core/sinc.h"',
        '#include "core/sum_ranges.h"',
        "namespace wayverb { namespace combined {", crossover, "} }",
        '#include "%s"' % os.path.join(HERE, "pp_driver.inc"),
    ]) + "\n"


HOST_UNITS = {"ref_scene.cpp": host_scene_program, "ref_pp.cpp": host_pp_program, "ref_hostmath.cpp": host_math_program, "ref_hostrun.cpp": host_run_program,
              "ref_hostrun_setup_program.cpp": single_file_unit("waveguide", "src", "mesh_setup_program.cpp"),
              "ref_hostrun_bcoef_program.cpp": single_file_unit("waveguide", "src", "boundary_coefficient_program.cpp"),
              "ref_hostray.cpp": host_ray_program,
              "ref_hostray_program.cpp": single_file_unit("raytracer", "src", "program.cpp"),
              "ref_hostray_stoch_program.cpp": single_file_unit("raytracer", "src", "stochastic", "program.cpp"),
              "ref_hostray_brdf.cpp": single_file_unit("raytracer", "src", "cl", "brdf.cpp")}
HOSTCL_UNITS = {"ref_hostrun.cpp", "ref_hostrun_setup_program.cpp", "ref_hostrun_bcoef_program.cpp", "ref_hostray.cpp",
                "ref_hostray_program.cpp", "ref_hostray_stoch_program.cpp", "ref_hostray_brdf.cpp"}     # compiled with hostcl/ (the cl.hpp stand-in) in front
HOST_INCLUDES = ["-I", os.path.join(HERE, "hoststubs"), "-I", os.path.join(REF, "src", "core", "include"),
                 "-I", os.path.join(REF, "src", "utilities", "include"),
                 "-I", os.path.join(REF, "src", "raytracer", "include"),
                 "-I", os.path.join(REF, "src", "frequency_domain", "include"),
                 "-I", os.path.join(REF, "src", "hrtf", "lib", "include"),
                 "-I", os.path.join(REF, "src", "waveguide", "include")]

UNITS = {
    "ref_wg_f32.cpp": lambda: waveguide_program(False),
    "ref_wg_f64.cpp": lambda: waveguide_program(True),
    "ref_rt.cpp": raytracer_program,
    "ref_stoch.cpp": stochastic_program,
    "ref_mesh_setup.cpp": mesh_setup_program,
    "ref_bcoef.cpp": boundary_coefficient_program,
}

CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
# -fpermissive: OpenCL-C is C, where int -> enum converts implicitly (e.g. `return -1;` from a
# function returning PortDirection, program.cpp:99); g++ needs the flag to accept that.
FLAGS = ["-std=gnu++17", "-O2", "-fPIC", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-fpermissive", "-w"]


def have_reference() -> bool:
    return os.path.isdir(os.path.join(REF, "src", "waveguide"))


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(d, f)) > t for d, _, files in os.walk(HERE) for f in files
               if not f.endswith(".pyc"))


def build(force: bool = False) -> str | None:
    """Returns the path of lib_ref.so, or None when neither the reference nor a prebuilt
    library is present (the GPU box has no /root/reference: it uses the prebuilt file)."""
    if not have_reference():
        return LIB if os.path.exists(LIB) else None
    if not force and not stale():
        return LIB
    os.makedirs(OUT, exist_ok=True)
    objs = []
    for fname, make in UNITS.items():
        path = os.path.join(OUT, fname)
        with open(path, "w") as f:
            f.write(make())
        obj = path[:-4] + ".o"
        r = subprocess.run([CXX] + FLAGS + ["-c", path, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed on %s:\n%s" % (path, r.stderr[-6000:]))
        objs.append(obj)
    for fname, make in HOST_UNITS.items():
        path = os.path.join(OUT, fname)
        with open(path, "w") as f:
            f.write(make())
        obj = path[:-4] + ".o"
        flags = ["-std=gnu++14", "-O2", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-w"]
        first = ["-I", os.path.join(HERE, "hostcl")] if fname in HOSTCL_UNITS else []
        r = subprocess.run([CXX] + flags + first + HOST_INCLUDES + ["-c", path, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed on %s:\n%s" % (path, r.stderr[-6000:]))
        objs.append(obj)
    r = subprocess.run([CXX, "-shared", "-fopenmp", "-o", LIB] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
