"""Measurement of the mesh-construction kernels: build the
waveguide mesh of a shoebox scene on the device vs the oracle on the host cores."""
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
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    box = (8.0, 6.0, 10.0)
    sc = scene.box_scene(box, subdiv=16, side=32, per_wall_surfaces=True,
                         surfaces=[scene.make_surface(0.1, 0.1), scene.make_surface(0.2, 0.1),
                                   scene.make_surface(0.3, 0.1)])
    spacing = np.float32(max(box) / (n - 8))
    mc = np.array([-3.5 * spacing] * 3, np.float32)
    dims = tuple(min(n, int(np.ceil((b - float(mc[0])) / spacing)) + 3) for b in box)
    c = np.zeros(3, wvb._lib.COEFF_DT)
    c["b"][:, 0], c["a"][:, 0] = 39.0, 1.0
    out = {"metric": "mesh nodes classified / s", "dims": dims, "nodes": int(np.prod(dims)),
           "triangles": int(sc.triangles.size)}
    with wvb.RayTracer(sc) as g:
        wvb.build_mesh((16, 16, 16), mc, spacing * 8, c, scene=g)  # warm-up
        t0 = time.perf_counter()
        m, ins = wvb.build_mesh(dims, mc, spacing, c, scene=g, return_inside=True)
        dt = time.perf_counter() - t0
    out.update({"gpu_s": dt, "gpu_value": out["nodes"] / dt, "inside_nodes": int(ins.sum()),
                "boundary_nodes": [int(b.shape[0]) for b in m.b]})
    if "--cpu" in sys.argv:
        from oracle import rto, wgo
        o = rto.Scene(sc)
        small = tuple(max(16, d // 4) for d in dims)
        t0 = time.perf_counter()
        ins_o = o.nodes_inside(mc, small, spacing * 4)
        z, y, x = np.indices(ins_o.shape)
        pts = np.stack([mc[0] + x.astype(np.float32) * spacing * 4, mc[1] + y.astype(np.float32) * spacing * 4,
                        mc[2] + z.astype(np.float32) * spacing * 4], -1).reshape(-1, 3)
        bt = wgo.classify(ins_o)["boundary_type"]
        need = np.array([bin(int(v)).count("1") == 1 and v != 1 for v in bt])
        o.closest_surface(pts[need])
        dto = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": float(np.prod(small)) / dto, "unit": "mesh nodes / s",
                               "cores": rto.num_threads(), "kind": "port",
                               "sample": "%s nodes (inside test + classification + 1-d finder), oracle" % (small,)}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
