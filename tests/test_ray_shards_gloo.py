"""N>1 host logic of the ray path on CPU: world_size-2 (and 3) gloo runs of the ray
sharding (wayverb_b200/slab.py ray_range == wvb_rt_allreduce_histogram), with the CPU
oracle standing in for the trace kernel. Rank r traces global rays [begin, end) with
`ray_index_base = begin`; the all-reduced histogram must equal the single-domain
oracle's (fp64 sums in a different order: 1e-12 relative to the largest bin), and
the per-ray reflections must be identical."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wayverb_b200 import scene as S
from wayverb_b200.slab import ray_range
from oracle import rto

TOTAL, DEPTH, SEED = 5000, 12, 21
SRC, RCV = (1.1, 1.2, 1.3), (3.0, 2.0, 4.5)


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def room():
    return S.box_scene((4.0, 3.0, 6.0), subdiv=2, side=8, surfaces=[S.make_surface(0.1, 0.2)])


def worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = rto.Scene(room())
    b, e = ray_range(TOTAL, rank, world)
    dirs = rto.directions(SEED, e - b, base=b)  # the generated directions are keyed by global index too
    hist, refl, dropped = o.trace(dirs, SRC, RCV, DEPTH, total_rays=TOTAL, seed=SEED, ray_index_base=b,
                                  keep_steps=3, n_bins=400)
    t = torch.from_numpy(hist)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)  # what wvb_rt_allreduce_histogram does over NCCL
    np.save(os.path.join(out_dir, "hist%d.npy" % rank), t.numpy())
    np.save(os.path.join(out_dir, "refl%d.npy" % rank), refl)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_rays_reproduce_single_domain(world, tmp_path):
    mp.spawn(worker, args=(world, free_port(), str(tmp_path)), nprocs=world, join=True)
    o = rto.Scene(room())
    want_h, want_r, _ = o.trace(rto.directions(SEED, TOTAL), SRC, RCV, DEPTH, seed=SEED, keep_steps=3, n_bins=400)
    scale = np.abs(want_h).max()
    assert scale > 0
    for r in range(world):
        got = np.load(tmp_path / ("hist%d.npy" % r))
        assert np.abs(got - want_h).max() <= 1e-12 * scale  # every rank holds the whole-job histogram
    got_r = np.concatenate([np.load(tmp_path / ("refl%d.npy" % r)) for r in range(world)], axis=1)
    assert np.array_equal(got_r.view(np.uint8), want_r.view(np.uint8))


def test_ray_ranges_tile_the_job():
    for total, n in ((10, 3), (1 << 20, 8), (7, 8)):
        edges = [ray_range(total, r, n) for r in range(n)]
        assert edges[0][0] == 0 and edges[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
