"""Temporal blocking A/B on one box: ms per step of wvb_wg_time_steps with and without
WVB_WG_TEMPORAL2, 512^3 plaster (development tool)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import wayverb_b200 as wvb  # noqa: E402
from wayverb_b200 import _lib  # noqa: E402

s = json.load(open(os.path.join(ROOT, "tests", "golden", "lrs_coefficients.json")))["sets"][0]["impedance"]
c = np.zeros((), _lib.COEFF_DT)
c["b"], c["a"] = s["b"], s["a"]
dims = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "512,512,512").split(","))
m = wvb.cuboid_mesh(dims, [c])
nodes = dims[0] * dims[1] * dims[2]
fields = {}
for name, flags in (("single", 0), ("tb2", _lib.TEMPORAL2), ("single", 0), ("tb2", _lib.TEMPORAL2)):
    with wvb.Waveguide(m, kernel=_lib.KERNEL_TMA, flags=flags) as g:
        g.write(m.index(dims[0] // 2, dims[1] // 2, dims[2] // 2), 1.0)
        g.time_steps(10)
        best = min(g.time_steps(100)[0] for _ in range(3)) / 100
        fields[name] = g.field()
        print(name, "%.4f ms/step" % best, "%.0f Mnode-updates/s" % (nodes / best / 1e3), g.info()["tile"], flush=True)
print("identical:", np.array_equal(fields["single"], fields["tb2"]))
