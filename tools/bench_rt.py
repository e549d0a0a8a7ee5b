"""Secondary measurement: the ray loop (BASELINE config 5 shape: 1 M rays, a
many-triangle hall, depth from the minimum absorption). Prints one JSON line with
ray-reflections/s on the GPU (kernel time by CUDA events and end-to-end through the
C ABI incl. copies) next to the CPU oracle on a bounded sample."""
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
    rays = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    subdiv = int(sys.argv[2]) if len(sys.argv) > 2 else 50  # 6 * 50 * 50 * 2 = 30 000 triangles
    side = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    absorption = [0.1, 0.1, 0.12, 0.15, 0.2, 0.25, 0.3, 0.35]
    t0 = time.perf_counter()
    sc = scene.box_scene((30.0, 12.0, 45.0), subdiv=subdiv, side=side,
                         surfaces=[scene.make_surface(absorption, 0.3)])
    build_s = time.perf_counter() - t0
    src, rcv = [8.0, 3.0, 10.0], [20.0, 7.0, 35.0]
    depth = wvb.reflection_depth(min(absorption))
    out = {"metric": "ray-reflections/s", "rays": rays, "depth": depth, "triangles": int(sc.triangles.size),
           "voxel_side": side, "scene_build_s": build_s}
    with wvb.RayTracer(sc) as g:
        g.trace(None, src, rcv, depth, n_rays=1 << 14, seed=1)  # warm-up
        g.reset_histogram()
        t0 = time.perf_counter()
        _, dropped, ms = g.trace(None, src, rcv, depth, n_rays=rays, seed=2)
        h = g.histogram()
        wall = time.perf_counter() - t0
        # with host-provided directions (the iterator range raytracer::run receives)
        d = g.directions(3, rays)
        g.reset_histogram()
        t0 = time.perf_counter()
        g.trace(d, src, rcv, depth, seed=3, keep_steps=4)
        h2 = g.histogram()
        wall2 = time.perf_counter() - t0
    out.update({"gpu_kernel_ms": ms, "gpu_value": rays * depth / (ms * 1e-3),
                "e2e_value_device_dirs": rays * depth / wall,
                "e2e_value_host_dirs_keep4": rays * depth / wall2,
                "dropped": int(dropped), "hist_energy": float(h.sum()), "hist_bins": int(h.shape[0])})
    if "--cpu" in sys.argv:
        from oracle import rto
        n = 1 << 14
        o = rto.Scene(sc)
        dd = rto.directions(5, n)
        t0 = time.perf_counter()
        o.trace(dd, src, rcv, depth, total_rays=rays, seed=2)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n * depth / dt, "unit": "ray-reflections/s", "cores": rto.num_threads(),
                               "kind": "port", "sample": "%d rays x %d steps, oracle/rt_oracle.cpp" % (n, depth)}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
