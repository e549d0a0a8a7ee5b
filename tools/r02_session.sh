#!/bin/bash
# one-GPU development session of round 2: boundary-kernel placement A/B on the same box, then the
# ncu evidence (launch list of the bench command, full captures of the waveguide and ray kernels)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
D=$PWD/wayverb_b200/libwvb200_dbg.so
if [ -f $D ]; then
WVB_LIB=$D timeout 900 python tools/ab_lib.py WVB_WG_BCARVE=-1 default \
  WVB_WG_AIRFIRST=0,WVB_WG_BTHREADS=64,WVB_WG_BPIPE=1 \
  WVB_WG_AIRFIRST=0,WVB_WG_BTHREADS=64,WVB_WG_BPIPE=2 \
  WVB_WG_AIRFIRST=0,WVB_WG_BTHREADS=128,WVB_WG_BPIPE=1,WVB_WG_BMINB=4 \
  WVB_WG_AIRFIRST=0 WVB_WG_AIRFIRST=0,WVB_WG_BPIPE=2 2 > gpurun_out/r02_ab_boundary2.txt 2>&1
cat gpurun_out/r02_ab_boundary2.txt
fi
if [ "$1" != "noprof" ]; then
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wg_air -s 3 -c 1 -f -o gpurun_out/r02_prof_air python tools/profile_wg.py tma 6 > gpurun_out/r02_ncu_air.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wg_boundary -s 3 -c 1 -f -o gpurun_out/r02_prof_bnd python tools/profile_wg.py tma 6 > gpurun_out/r02_ncu_bnd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rt_trace -s 1 -c 1 -f -o gpurun_out/r02_prof_rt_hall python tools/profile_rt.py 262144 hall > gpurun_out/r02_ncu_rt.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:is_validate -c 1 -f -o gpurun_out/r02_prof_is_hall python tools/profile_rt.py 262144 hall > gpurun_out/r02_ncu_is.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/r02_ncu_rt.log
fi
