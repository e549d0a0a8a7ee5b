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


def ray_check(rank, world, local):
    """the ray path over NCCL: every rank traces its share, wvb_rt_allreduce_histogram sums;
    rank 0 compares with the single-domain oracle (and with one GPU tracing everything)"""
    from wayverb_b200 import scene
    from wayverb_b200.slab import ray_range
    from oracle import rto
    total, depth, seed = 200000, 20, 33
    src, rcv = (1.1, 1.2, 1.3), (3.0, 2.0, 4.5)
    sc = scene.box_scene((4.0, 3.0, 6.0), subdiv=3, side=8, surfaces=[scene.make_surface(0.1, 0.2)])
    box = [wvb.waveguide.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    b, e = ray_range(total, rank, world)
    with wvb.RayTracer(sc, device=local) as g:
        g.comm_init(box[0], rank, world)
        g.trace(None, src, rcv, depth, n_rays=e - b, total_rays=total, seed=seed, ray_index_base=b, n_bins=500)
        g.allreduce_histogram()
        got = g.histogram()
    ok = True
    if rank == 0:
        o = rto.Scene(sc)
        want, _, _ = o.trace(rto.directions(seed, total), src, rcv, depth, seed=seed, n_bins=500)
        with wvb.RayTracer(sc, device=local) as g1:
            g1.trace(None, src, rcv, depth, n_rays=total, seed=seed, n_bins=500)
            one = g1.histogram()
        scale = np.abs(want).max()
        err_o = np.abs(got - want).max() / scale
        err_1 = np.abs(got - one).max() / scale
        ok = err_o <= 1e-9 and err_1 <= 1e-9 and scale > 0
        print("rays %d over %d ranks: all-reduced histogram vs oracle %.1e, vs one GPU %.1e (rel. to max bin)"
              % (total, world, err_o, err_1), flush=True)
    return ok


def per_step_check(rank, world, local):
    """the per-step path on slabs (what waveguide::run drives): write / launch / read / swap on every
    rank, the error flag gathered over peer memory and the receiver answered from the launch's cache;
    receiver samples against the single-domain oracle, and an inf planted on rank 0 must raise the
    flag on EVERY rank in the same step"""
    ok = True
    for halo in (_lib.HALO_AUTO, _lib.HALO_NCCL):
        dims = (140, 40, 10 * world + 4)
        box = [wvb.waveguide.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        dx, dy, dz = dims
        z0, z1 = wvb.slab_range(dz, rank, world)
        lo, hi = max(z0 - 1, 0), min(z1 + 1, dz)
        mesh = wvb.cuboid_mesh(dims, [plaster()], z0=lo, nz=hi - lo)
        src = mesh.index(dx // 2, dy // 2, wvb.slab_range(dz, 0, world)[1] - 1)
        rcv = mesh.index(dx // 2 + 3, dy // 2 - 2, dz - 4)          # in the last slab
        steps = 3 * dz
        got = np.zeros(steps)
        with wvb.Waveguide(mesh, device=local, z_range=(z0, z1), rank=rank, nranks=world, nccl_unique_id=box[0],
                           kernel=_lib.KERNEL_TMA, flags=halo) as g:
            for s in range(steps):
                g.write(src, 1.0 if s == 0 else 0.0)
                assert g.launch() == 0
                got[s] = g.read(rcv)      # 0 on ranks that do not own the node
                g.swap()
            # error flag: inf on rank 0's source node; every rank must see it after the same launch
            g.write(src, float("inf"))
            flags = []
            for s in range(3):
                flags.append(g.launch())
                g.swap()
            halo_name = g.info()["halo"]
        t = torch.from_numpy(got).cuda()
        dist.all_reduce(t)
        fl = torch.tensor(flags, device="cuda")
        all_fl = [torch.zeros_like(fl) for _ in range(world)]
        dist.all_gather(all_fl, fl)
        if rank == 0:
            om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [plaster()])
            sim = wgo.Sim(om)
            sig = np.zeros(steps)
            sig[0] = 1.0
            _, want, _ = sim.run(src, sig, [rcv], soft=False)
            same = np.array_equal(t.cpu().numpy(), want[:, 0])
            same_flags = all(torch.equal(all_fl[0], f) for f in all_fl) and int(all_fl[0][0]) != 0
            print("per-step path, halo %s, ranks %d: receiver samples identical=%s, error flags equal on all ranks=%s %s"
                  % (halo_name, world, same, same_flags, all_fl[0].tolist()), flush=True)
            ok = ok and same and same_flags and np.abs(want).max() > 0
    return ok


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    cases = [((140, 40, 12 * world + 5), _lib.KERNEL_TMA, _lib.HALO_AUTO),
             ((140, 40, 12 * world + 5), _lib.KERNEL_TMA, _lib.HALO_NCCL),
             ((140, 40, 12 * world + 5), _lib.KERNEL_TMA, _lib.HALO_P2P | _lib.HALO_OVERLAP),
             ((37, 29, 8 * world + 3), _lib.KERNEL_DIRECT, _lib.HALO_P2P),
             ((37, 29, 8 * world + 3), _lib.KERNEL_DIRECT, _lib.HALO_NCCL | _lib.HALO_OVERLAP)]
    for dims, kernel, halo in cases:
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
                           nccl_unique_id=uid, kernel=kernel, flags=halo) as g:
            done, out, flag = g.run_device(src, sig, rcv, soft=True, check_interval=7)
            # a second run on the same handle: plain steps (the captured per-step graph where
            # the transport allows one) continuing from the state the first run left
            flag2 = g.step(9)
            field = g.field()
            info = g.info()
        assert done == steps and flag == 0 and flag2 == 0, (done, flag, flag2)
        out_t = torch.from_numpy(out).cuda()
        dist.all_reduce(out_t)  # receivers not owned by a rank are written as 0
        parts = [None] * world
        dist.all_gather_object(parts, field)
        if rank == 0:
            om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [plaster()])
            sim = wgo.Sim(om)
            _, want_out, wflag = sim.run(src, sig, rcv, soft=True)
            wflag |= sim.step(9)
            want = sim.field()
            got = np.concatenate(parts)
            same_f = np.array_equal(got, want)
            same_o = np.array_equal(out_t.cpu().numpy(), want_out)
            rms = np.sqrt(np.mean((got - want) ** 2)) / np.abs(want).max()
            print("dims %s kernel %s halo %s ranks %d: field identical=%s (rel RMS %.1e), traces identical=%s"
                  % (dims, info["kernel_variant"], info["halo"], world, same_f, rms, same_o), flush=True)
            ok = ok and same_f and same_o and wflag == 0
    ok = per_step_check(rank, world, local) and ok
    ok = ray_check(rank, world, local) and ok
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
