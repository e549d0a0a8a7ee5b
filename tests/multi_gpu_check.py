"""Multi-GPU parity check (run under torchrun on N >= 2 GPUs of one node):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py

Every rank owns a z-slab of one mesh (NCCL ghost-plane exchange inside
libwvb200.so); rank 0 gathers the slabs and compares them, and the receiver
traces, with the single-domain CPU oracle: they must be identical.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wayverb_b200 as wvb  # noqa: E402
from wayverb_b200 import _lib  # noqa: E402
from oracle import wgo  # noqa: E402


def plaster():
    s = json.load(open(os.path.join(ROOT, "tests", "golden", "lrs_coefficients.json")))["sets"][0]["impedance"]
    c = np.zeros((), _lib.COEFF_DT)
    c["b"], c["a"] = s["b"], s["a"]
    return c


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for dims, kernel in (((140, 40, 12 * world + 5), _lib.KERNEL_TMA), ((37, 29, 8 * world + 3), _lib.KERNEL_DIRECT)):
        # an ncclUniqueId serves exactly one communicator: a fresh one per handle
        box = [wvb.waveguide.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
        dx, dy, dz = dims
        z0, z1 = wvb.slab_range(dz, rank, world)
        lo, hi = max(z0 - 1, 0), min(z1 + 1, dz)
        mesh = wvb.cuboid_mesh(dims, [plaster()], z0=lo, nz=hi - lo)
        steps = 3 * dz
        sig = np.zeros(steps)
        sig[:3] = [1.0, 0.0, -1.0]
        # source on the last plane of rank 0's slab, receivers in the first and last slab
        b0 = wvb.slab_range(dz, 0, world)[1]
        src = mesh.index(dx // 2, dy // 2, b0 - 1)
        rcv = [mesh.index(5, 6, 3), mesh.index(dx - 6, dy - 5, dz - 4), src]
        with wvb.Waveguide(mesh, device=local, z_range=(z0, z1), rank=rank, nranks=world,
                           nccl_unique_id=uid, kernel=kernel) as g:
            done, out, flag = g.run_device(src, sig, rcv, soft=True, check_interval=7)
            field = g.field()
            info = g.info()
        assert done == steps and flag == 0, (done, flag)
        out_t = torch.from_numpy(out).cuda()
        dist.all_reduce(out_t)  # receivers not owned by a rank are written as 0
        parts = [None] * world
        dist.all_gather_object(parts, field)
        if rank == 0:
            om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [plaster()])
            sim = wgo.Sim(om)
            _, want_out, wflag = sim.run(src, sig, rcv, soft=True)
            want = sim.field()
            got = np.concatenate(parts)
            same_f = np.array_equal(got, want)
            same_o = np.array_equal(out_t.cpu().numpy(), want_out)
            rms = np.sqrt(np.mean((got - want) ** 2)) / np.abs(want).max()
            print("dims %s kernel %s ranks %d: field identical=%s (rel RMS %.1e), traces identical=%s"
                  % (dims, info["kernel_variant"], world, same_f, rms, same_o), flush=True)
            ok = ok and same_f and same_o and wflag == 0
    res = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(res, 0)
    dist.barrier()
    dist.destroy_process_group()
    if not int(res.item()):
        sys.exit(1)
    if rank == 0:
        print("MULTI_GPU_CHECK_OK")


if __name__ == "__main__":
    main()
