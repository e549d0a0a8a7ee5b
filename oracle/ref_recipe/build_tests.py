"""oracle/ref_recipe/build_tests.py -- builds the REFERENCE'S OWN unit tests against oracle/_ref/lib_ref.so.

TEST INFRASTRUCTURE ONLY. The stand-ins this recipe relies on (hoststubs/: GLM, FFTW, IT++, googletest;
hostcl/: cl.hpp) decide a handful of things and say so in their headers. This is the check on them: the
reference's own test files for the code on and around the path -- src/core/tests, src/raytracer/tests,
src/waveguide/tests, src/frequency_domain/tests -- are compiled UNMODIFIED, where they lie, on top of the
same stand-ins and the same host build of the reference's sources, and run. If a vector operator, the
host-memory command queue, a kernel launcher or the FFT were wrong, the reference's own expectations of
its geometry, ray kernels, image sources, histograms, waveguide and filter bank would fail.

Left out, and why:
  * tests that load material / audio files through cereal or libsndfile (filter, reconstruction,
    multiband_filter, boundary_tests ...): neither library is here. rectangular_kernel only WRITES .wav
    files for a listener (no assertion on them): audio_file's writer is a no-op in reftest_support.cpp and
    the test runs; waveguide_tests needs make_transparent, whose mesh_impulse_response.h is generated. The tests that only load MODELS (voxel_tests, mesh_tests, mesh_setup_tests, stochastic_tests) do
    run: assimp's loader is stood in for by the library's own OBJ reader (group "models");
  * gpu_geometry_tests: it builds an OpenCL program from kernel source written inside the test file;
  * equal_energy.cpp: g++ 13 stops with an internal compiler error on it;
  * waveguide_init / verify_compensation_signal: need the generated mesh_impulse_response.h;
  * arbitrary_magnitude_filter / fitted_boundary tests: they test IT++'s fit itself (and cereal);
  * tests of code that is not on the path (dc blocker, schroeder, attenuators, orientation, ...).
Known result: core/tests/vector_look_up_table.cpp's `index` and `pointing` cases expect +z to be "front";
az_el.cpp has -z (compute_azimuth = atan2(x, -z), compute_pointing -> (0, 0, -1) for azimuth 0), so the
reference fails these two itself. The following cases are statistical, seeded from std::random_device:
multiband.noise (eight 20 % bounds on 40-bin estimates: passes about three times in five) and
tri_cube_tests.comparison (two float implementations of the same predicate on 2^20 random triangles:
about one disagreement per two million triangles, so it passes about three times in five), and
image_source.fast_pressure draws source and receiver anywhere in the room and needs 10 000 random rays to
find every exact image source within 10 m, matched inside a window of neighbours by distance: it throws
"No approximate matches." about one time in ten. The reflector fixture compares the kernel's hit
positions with the CPU twin's for randomly drawn rays through ten layers of reflections to 1e-5 and asks
the voxel walk and the brute-force search for the same triangle: a ray that grazes an edge breaks it about
once in eighty runs.

Usage: python oracle/ref_recipe/build_tests.py [--all]  -> oracle/_ref/reftest_{core,raytracer,waveguide,frequency_domain,models}
       (--all: also the six-minute nan_in_waveguide run)
"""
from __future__ import annotations

import importlib.util
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("_wvb_ref_recipe_build", os.path.join(HERE, "build.py"))
recipe = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(recipe)

REPO = os.path.normpath(os.path.join(HERE, "..", ".."))
SRC = os.path.join(recipe.REF, "src")

GROUPS = {
    "core": [("core", "tests", f + ".cpp") for f in
             ("main", "tri_cube_tests", "geo_tests", "box_tests", "indexing_tests", "vector_look_up_table",
              "recursive_vector", "cosine_interp", "pressure_intensity")],
    "raytracer": [("raytracer", "tests", f + ".cpp") for f in
                  ("main", "reflector_tests", "image_source", "multitree", "histogram", "pressure", "brdf",
                   "build_program")],
    "waveguide": [("waveguide", "tests", f + ".cpp") for f in ("main", "build_program", "rectangular_kernel")],
    "frequency_domain": [("frequency_domain", "tests", f + ".cpp") for f in ("main", "multiband", "convolution")] +
                        [("frequency_domain", "src", "convolver.cpp")],
    # the tests that run on the reference's own Wavefront models (demo/assets/test_models): assimp's loader is
    # stood in for by the library's OBJ reader (reftest_support.cpp)
    "models": [("core", "tests", "main.cpp"), ("core", "tests", "voxel_tests.cpp"),
               ("waveguide", "tests", "mesh_tests.cpp"), ("waveguide", "tests", "mesh_setup_tests.cpp"),
               ("raytracer", "tests", "stochastic_tests.cpp"), ("utilities", "src", "progress_bar.cpp")],
}
MODELS = os.path.join(recipe.REF, "demo", "assets", "test_models")
DEFINES = ['-DSCRATCH_PATH="/tmp"', '-DOBJ_PATH="%s"' % os.path.join(MODELS, "vault.obj"),
           '-DOBJ_PATH_TUNNEL="%s"' % os.path.join(MODELS, "echo_tunnel.obj"),
           '-DOBJ_PATH_BEDROOM="%s"' % os.path.join(MODELS, "bedroom.obj"),
           '-DOBJ_PATH_BAD_BOX="%s"' % os.path.join(MODELS, "small_square.obj")]
# built and run only on request (`--all`): nan_in_waveguide steps a 56-million-node mesh 432 times through the
# host-compiled kernel -- six minutes on eight cores (profiles/r02_reference_own_tests.txt has the run)
SLOW_GROUPS = {
    "waveguide_slow": [("waveguide", "tests", f + ".cpp") for f in ("main", "nan_in_waveguide")] +
                      [("utilities", "src", "progress_bar.cpp")],
}
# cases the reference's own code does not satisfy (see the docstring)
KNOWN_STALE = {"vector_look_up_table.index", "vector_look_up_table.pointing"}
STATISTICAL = {"multiband.noise", "tri_cube_tests.comparison", "image_source.fast_pressure",
               "reflector_fixture.multi_layer_reflections", "reflector_fixture.locations"}


def exe(group: str) -> str:
    return os.path.join(recipe.OUT, "reftest_" + group)


def build(force: bool = False, slow: bool = False) -> dict[str, str] | None:
    """-> {group: executable}, or None without /root/reference (the executables travel, like lib_ref.so)"""
    groups = dict(GROUPS, **(SLOW_GROUPS if slow else {}))
    if not recipe.have_reference():
        out = {g: exe(g) for g in groups}
        return out if all(os.path.exists(p) for p in out.values()) else None
    lib = recipe.build()
    out = {}
    includes = ["-I", os.path.join(HERE, "hostcl")] + recipe.HOST_INCLUDES + \
               ["-I", os.path.join(SRC, "frequency_domain", "src"), "-I", os.path.join(SRC, "audio_file", "include"),
                "-I", os.path.join(REPO, "include")]
    wvb = os.path.join(REPO, "wayverb_b200")
    for group, files in groups.items():
        target = exe(group)
        out[group] = target
        sources = [os.path.join(SRC, *f) for f in files] + [os.path.join(HERE, "reftest_support.cpp")]
        newest = max(os.path.getmtime(p) for p in sources + [lib, os.path.abspath(__file__)])
        if not force and os.path.exists(target) and os.path.getmtime(target) >= newest:
            continue
        # -include gtest/gtest.h: the standard headers that googletest, GLM and cl.hpp bring in before the
        # reference's own headers are read (several of those use <limits>, <tuple>, size_t without naming them)
        cmd = [recipe.CXX, "-std=gnu++14", "-O1", "-w", "-ffp-contract=off", "-include", "gtest/gtest.h"] + includes + \
              DEFINES + sources + \
              ["-o", target, "-L" + recipe.OUT, "-l:lib_ref.so", "-Wl,-rpath," + recipe.OUT, "-Wl,-rpath,$ORIGIN",
               "-L" + wvb, "-lwvb200", "-Wl,-rpath," + wvb, "-Wl,-rpath,$ORIGIN/../../wayverb_b200", "-fopenmp", "-pthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed on the reference's %s tests:\n%s" % (group, r.stderr[-6000:]))
    return out


def run(group: str, timeout: int = 1800) -> dict[str, bool]:
    """-> {case: passed}"""
    r = subprocess.run([exe(group)], capture_output=True, text=True, timeout=timeout)
    result = {}
    for line in r.stdout.splitlines():
        if line.startswith("[  OK  ] "):
            result[line[9:].strip()] = True
        elif line.startswith("[FAILED] "):
            result[line[9:].strip()] = False
    if not result:
        raise RuntimeError("no test ran:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    return result


if __name__ == "__main__":
    built = build(force="--force" in sys.argv, slow="--all" in sys.argv)
    for g in built or {}:
        res = run(g)
        print(g, sum(res.values()), "of", len(res), "passed;", "failed:", sorted(k for k, v in res.items() if not v))
