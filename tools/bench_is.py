"""Secondary measurement: the image-source stage on the BASELINE
config 5 shape (1 M rays, a 30 000-triangle hall, image-source order 4). Prints one
JSON line: tree nodes validated per second on the GPU (validation kernel by CUDA
events, and the whole results() call incl. ordering and copies) next to the CPU
oracle's tree walk on a bounded sample of the same rays."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import wayverb_b200 as wvb  # noqa: E402
from wayverb_b200 import scene  # noqa: E402


def main():
    rays = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1 << 20
    order = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 4
    subdiv, side = 50, 32
    absorption = [0.1, 0.1, 0.12, 0.15, 0.2, 0.25, 0.3, 0.35]
    sc = scene.box_scene((30.0, 12.0, 45.0), subdiv=subdiv, side=side,
                         surfaces=[scene.make_surface(absorption, 0.0)])
    src, rcv = [8.0, 3.0, 10.0], [20.0, 7.0, 35.0]
    out = {"metric": "image-source tree nodes validated/s", "rays": rays, "order": order,
           "triangles": int(sc.triangles.size), "voxel_side": side}
    seg = 1 << 14
    with wvb.RayTracer(sc) as g:
        with wvb.ImageSource(g, src, rcv, max_elements=seg * order) as s:  # warm-up
            s.trace(None, depth=order, order=order, n_rays=seg, seed=1)
            s.results()
        with wvb.ImageSource(g, src, rcv, max_elements=rays * order) as s:
            t0 = time.perf_counter()
            for b in range(0, rays, seg):  # the segments of raytracer::run (raytracer.h:219-244)
                n = min(seg, rays - b)
                s.trace(None, depth=order, order=order, n_rays=n, total_rays=rays, seed=2, ray_index_base=b)
            t_push = time.perf_counter() - t0
            t0 = time.perf_counter()
            imps, stats, ms = s.results()
            t_res = time.perf_counter() - t0
        out.update({"tree_nodes": int(stats[0]), "visible_nodes": int(stats[1]), "impulses": int(imps.size),
                    "trace_and_insert_s": t_push, "validate_kernel_ms": ms, "results_call_s": t_res,
                    "gpu_value": float(stats[0]) / (ms * 1e-3), "e2e_value": float(stats[0]) / t_res})
        if "--cpu" in sys.argv:
            from oracle import rto
            n = min(rays, 1 << 16)
            # the same rays' elements, via the GPU trace (bit-identical to the oracle's, tests/test_rt_gpu.py)
            refl, _, _ = g.trace(None, src, rcv, order, n_rays=n, total_rays=rays, seed=2, keep_steps=order)
            elems = rto.path_elements(refl, order)
            o = rto.Scene(sc)
            t0 = time.perf_counter()
            want, st = rto.image_source(o, elems, src, rcv)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": float(st[0]) / dt, "unit": "tree nodes validated/s", "cores": 1,
                                   "kind": "port", "tree_nodes": int(st[0]), "seconds": dt,
                                   "sample": "first %d rays, oracle/is_oracle.inc (single thread; the reference "
                                             "runs one std::async per first-order branch)" % n}
            with wvb.ImageSource(g, src, rcv, max_elements=n * order) as s:
                s.push_elements(elems)
                got, _, _ = s.results()
            out["sample_identical_to_oracle"] = bool(got.shape == want.shape and
                                                     np.array_equal(got.view(np.uint8), want.view(np.uint8)))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
