"""Runs one ray trace for ncu (development tool)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wayverb_b200 as wvb  # noqa: E402
from wayverb_b200 import scene  # noqa: E402

rays = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
which = sys.argv[2] if len(sys.argv) > 2 else "box"
if which == "box":      # the round-1 workload: 30 000 triangles, depth 132
    sc = scene.box_scene((30.0, 12.0, 45.0), subdiv=50, side=32,
                         surfaces=[scene.make_surface([0.1, 0.1, 0.12, 0.15, 0.2, 0.25, 0.3, 0.35], 0.3)])
    src, rcv, depth = [8.0, 3.0, 10.0], [20.0, 7.0, 35.0], 132
else:                   # BASELINE config 5's geometry: "hall" (322 triangles) or "hall3" (20 608)
    sc, meta = scene.concert_hall(3 if which == "hall3" else 0)
    src, rcv, depth = meta["source"], meta["receiver"], wvb.reflection_depth(meta["min_absorption"])
with wvb.RayTracer(sc) as g, wvb.ImageSource(g, src, rcv, max_elements=rays * 4) as s:
    g.trace(None, src, rcv, depth, n_rays=1 << 12, seed=1)
    for mode, name in ((1, "ray-life"), (2, "wavefront"), (1, "ray-life"), (2, "wavefront")):
        g.reset_histogram()
        _, dropped, ms = g.trace(None, src, rcv, depth, n_rays=rays, seed=2, mode=mode)
        print(which, name, rays, depth, "%.3f ms" % ms, "%.1f M ray-reflections/s" % (rays * depth / ms / 1e3), flush=True)
    s.trace(None, depth=depth, order=4, n_rays=rays, seed=2)
    imp, stats, vms = s.results()
    print("image sources", imp.size, stats.tolist(), vms, "ms validate")
