"""Times the waveguide step for kernel variants / tile parameters on one GPU (needs a library built
with -DWVB_DEBUG_KNOBS and WVB_LIB pointing at it: the shipped library ignores the WVB_WG_* knobs).
Development tool (not the bench): prints Mnode-updates/s and the 32 B/node
bandwidth figure for each configuration."""
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
    configs = []
    base = dict(WVB_WG_KERNEL="tma", WVB_WG_TY=8, WVB_WG_MINB=1, WVB_WG_DIV=1, WVB_WG_STAGES=5, WVB_WG_ZCHUNKS=12,
                WVB_WG_OVERLAP=1)
    for bp, af in ((0, 1), (1, 0), (2, 0), (1, 1), (4, 0), (3, 0)):
        configs.append(dict(base, WVB_WG_BPERSIST=bp, WVB_WG_AIRFIRST=af))
    only = os.environ.get("SWEEP_ONLY")
    for cfg in configs:
        if only and cfg["WVB_WG_KERNEL"] != only:
            continue
        for k, v in cfg.items():
            os.environ[k] = str(v)
        kern = cfg["WVB_WG_KERNEL"]
        tag = " ".join("%s=%s" % (k[7:].lower(), v) for k, v in cfg.items() if k != "WVB_WG_KERNEL")
        try:
            with wvb.Waveguide(m) as g:
                g.write(m.index(dims[0] // 2, dims[1] // 2, dims[2] // 2), 1.0)
                g.time_steps(5)
                best = min(g.time_steps(steps)[0] for _ in range(3))
                info = g.info()
            ms = best / steps
            print("%-6s %-40s | %.4f ms/step  %8.1f Mnode/s  %7.1f GB/s@32B  tile=%s" %
                  (kern, tag, ms, nodes / ms / 1e3, nodes * 32 / ms / 1e6, info["tile"]), flush=True)
        except Exception as e:  # noqa: BLE001
            print(kern, tag, "FAILED", e, flush=True)


if __name__ == "__main__":
    main()
