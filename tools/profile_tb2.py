"""A few temporal-blocking pairs at 512^3 for ncu (development tool)."""
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
m = wvb.cuboid_mesh((512, 512, 512), [c])
with wvb.Waveguide(m, kernel=_lib.KERNEL_TMA, flags=_lib.TEMPORAL2) as g:
    g.write(m.index(256, 256, 256), 1.0)
    assert g.step(int(sys.argv[1]) if len(sys.argv) > 1 else 8) == 0
    print(g.info())
