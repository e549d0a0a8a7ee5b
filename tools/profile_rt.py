"""Runs one ray trace for ncu (development tool)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wayverb_b200 as wvb  # noqa: E402
from wayverb_b200 import scene  # noqa: E402

rays = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
sc = scene.box_scene((30.0, 12.0, 45.0), subdiv=50, side=32,
                     surfaces=[scene.make_surface([0.1, 0.1, 0.12, 0.15, 0.2, 0.25, 0.3, 0.35], 0.3)])
with wvb.RayTracer(sc) as g:
    _, dropped, ms = g.trace(None, [8.0, 3.0, 10.0], [20.0, 7.0, 35.0], 132, n_rays=rays, seed=2)
    print(rays, ms, rays * 132 / ms / 1e3, "ray-reflections/s")
