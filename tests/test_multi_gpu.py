"""N-rank NCCL correctness, collected by pytest: when the box has >= 2 GPUs, tests/multi_gpu_check.py
(z-slabs with the per-step ghost-plane exchange, and ray shards with the histogram all-reduce,
each compared with the single-domain CPU oracle) runs under torchrun on 2 ranks. On a
one-GPU box the test is skipped; bench.py's `multi_gpu_parity` gate covers N = 2/4/8 there."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_slabs_and_ray_shards_match_the_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (this box has %d)" % torch.cuda.device_count())
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTI_GPU_CHECK_OK" in r.stdout
