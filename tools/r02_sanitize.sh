#!/bin/bash
# compute-sanitizer over the kernels added in round 2 (small cases): temporal blocking, wavefront ray
# loop, voxel-grid finder, post-processing, image-source walk
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
T="tests/test_wg_gpu.py::test_temporal_blocking_is_bit_identical tests/test_wg_gpu.py::test_temporal_blocking_l_shaped_room_and_random_field tests/test_rt_gpu.py::test_wavefront_dead_rays_directional_and_segments tests/test_pp.py tests/test_is_gpu.py tests/test_mesh_build.py tests/test_edge_cases_gpu.py"
timeout 1500 $CS --tool memcheck --error-exitcode 9 python -m pytest $T -m gpu -q -x > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck.txt
timeout 900 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_wg_gpu.py::test_temporal_blocking_is_bit_identical tests/test_pp.py -m gpu -q -x > gpurun_out/r02_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck.txt
timeout 900 $CS --tool initcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_initcheck.txt 2>&1; echo "initcheck rc=$?" >> gpurun_out/r02_sanitizer_initcheck.txt
tail -4 gpurun_out/r02_sanitizer_memcheck.txt; tail -4 gpurun_out/r02_sanitizer_racecheck.txt; tail -4 gpurun_out/r02_sanitizer_initcheck.txt
