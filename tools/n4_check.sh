#!/bin/bash
# N-GPU session: parity check (interior ranks have two neighbours), then one full bench line
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 900 $TR --master-port 29511 tests/multi_gpu_check.py > gpurun_out/r02_multi_gpu_check_n$N.txt 2>&1
echo "check rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r02_multi_gpu_check_n$N.txt | tail -12
timeout 900 $TR --master-port 29513 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/r02_bench_n${N}_full.json 2> gpurun_out/r02_bench_n${N}_full.err
echo "full rc=$?"; cat gpurun_out/r02_bench_n${N}_full.json; tail -c 600 gpurun_out/r02_bench_n${N}_full.err
timeout 600 $TR --master-port 29514 bench.py --gpus $N --steps 200 --warmup 5 --halo nccl --no-extras > gpurun_out/r02_bench_n${N}_nccl.json 2> gpurun_out/r02_bench_n${N}_nccl.err
echo "nccl rc=$?"; cut -c1-400 gpurun_out/r02_bench_n${N}_nccl.json
timeout 300 python bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_ref_n$N.json 2>&1; cut -c1-300 gpurun_out/r02_ref_n$N.json
