"""Temporal blocking under a sustained load (the board power-caps after ~1 s at 512^3): ms per step
of 3000-step runs with and without WVB_WG_TEMPORAL2, alternating (development tool)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import wayverb_b200 as wvb  # noqa: E402
from wayverb_b200 import _lib  # noqa: E402

s = json.load(open(os.path.join(ROOT, "tests", "golden", "lrs_coefficients.json")))["sets"][0]["impedance"]
c = np.zeros((), _lib.COEFF_DT)
c["b"], c["a"] = s["b"], s["a"]
dims = (512, 512, 512)
m = wvb.cuboid_mesh(dims, [c])
nodes = dims[0] * dims[1] * dims[2]
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
handles = {"single": wvb.Waveguide(m, kernel=_lib.KERNEL_TMA), "tb2": wvb.Waveguide(m, kernel=_lib.KERNEL_TMA, flags=_lib.TEMPORAL2)}
for g in handles.values():
    g.write(m.index(256, 256, 256), 1.0)
    g.time_steps(20)
for rep in range(3):
    for name, g in handles.items():
        ms = g.time_steps(steps)[0] / steps
        print(rep, name, "%.4f ms/step" % ms, "%.0f Mnode-updates/s" % (nodes / ms / 1e3), flush=True)
