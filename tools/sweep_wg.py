"""Times the waveguide step for kernel variants / tile parameters on one GPU.
Development tool (not the bench): prints Mnode-updates/s and the 32 B/node
bandwidth figure for each configuration."""
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import wayverb_b200 as wvb  # noqa: E402
from wayverb_b200 import _lib  # noqa: E402


def plaster():
    s = json.load(open(os.path.join(ROOT, "tests", "golden", "lrs_coefficients.json")))["sets"][0]["impedance"]
    c = np.zeros((), _lib.COEFF_DT)
    c["b"], c["a"] = s["b"], s["a"]
    return c


def main():
    dims = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "512,512,512").split(","))
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    m = wvb.cuboid_mesh(dims, [plaster()])
    nodes = dims[0] * dims[1] * dims[2]
    configs = [("direct", None, None, z) for z in (0, 4, 16)]
    for ty, st, zc in itertools.product((8, 16), (4, 5, 6), (0, 3, 7, 9, 16)):
        if ty == 16 and st == 6:
            continue
        configs.append(("tma", ty, st, zc))
    for kern, ty, st, zc in configs:
        os.environ["WVB_WG_KERNEL"] = kern
        os.environ["WVB_WG_ZCHUNKS"] = str(zc)
        if ty:
            os.environ["WVB_WG_TY"] = str(ty)
            os.environ["WVB_WG_STAGES"] = str(st)
        try:
            with wvb.Waveguide(m) as g:
                g.write(m.index(dims[0] // 2, dims[1] // 2, dims[2] // 2), 1.0)
                g.time_steps(5)
                best = min(g.time_steps(steps)[0] for _ in range(3))
                info = g.info()
            ms = best / steps
            print("%-6s ty=%s st=%s zc=%-3d | %.4f ms/step  %8.1f Mnode/s  %7.1f GB/s@32B  tile=%s" %
                  (kern, ty, st, info["tile"][2], ms, nodes / ms / 1e3, nodes * 32 / ms / 1e6, info["tile"]),
                  flush=True)
        except Exception as e:  # noqa: BLE001
            print(kern, ty, st, zc, "FAILED", e, flush=True)


if __name__ == "__main__":
    main()
