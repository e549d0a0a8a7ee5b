"""Stress of the multi-GPU per-step path alone (see per_step_check in multi_gpu_check.py), repeated:
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/multi_gpu_per_step.py [reps]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import multi_gpu_check as M  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    ok = M.per_step_check(rank, world, local) and ok
res = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(res, 0)
dist.barrier()
dist.destroy_process_group()
if not int(res.item()):
    sys.exit(1)
if rank == 0:
    print("PER_STEP_OK")
