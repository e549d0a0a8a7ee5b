#!/bin/bash
# 2-GPU session: multi-GPU parity check, then the ghost-plane transports A/B, then one full bench line
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 900 $TR --master-port 29511 tests/multi_gpu_check.py > gpurun_out/r02_multi_gpu_check_n$N.txt 2>&1
echo "check rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r02_multi_gpu_check_n$N.txt | tail -12
for h in nccl p2p p2p+overlap nccl+overlap; do
  timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 500 --warmup 5 --halo $h --no-extras > gpurun_out/r02_bench_n${N}_$h.json 2> gpurun_out/r02_bench_n${N}_$h.err
  echo "$h rc=$?"; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_n${N}_$h.json").read().strip().splitlines()[-1])
    print("$h", d["value"], d["ms_per_step"], d["config"]["parallelism"], d.get("multi_gpu_parity",{}).get("identical"))
except Exception as e:
    print("$h failed", e); print(open("gpurun_out/r02_bench_n${N}_$h.err").read()[-1500:])
P
done
timeout 900 $TR --master-port 29513 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/r02_bench_n${N}_full.json 2> gpurun_out/r02_bench_n${N}_full.err
echo "full rc=$?"; cat gpurun_out/r02_bench_n${N}_full.json; tail -c 800 gpurun_out/r02_bench_n${N}_full.err
