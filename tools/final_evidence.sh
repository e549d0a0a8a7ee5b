#!/bin/bash
# Round-end evidence on one B200: smoke, GPU tests, bench, ncu launch list and one
# full capture of each waveguide kernel. Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 700 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:wg_air -s 3 -c 1 -f -o gpurun_out/prof_air_final python tools/profile_wg.py tma 6 > gpurun_out/ncu_air.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:wg_boundary -s 3 -c 1 -f -o gpurun_out/prof_bnd_final python tools/profile_wg.py tma 6 > gpurun_out/ncu_bnd.log 2>&1
tail -3 gpurun_out/smoke.log; tail -3 gpurun_out/pytest.log; cat gpurun_out/bench.json; tail -2 gpurun_out/bench.err
