"""A/B of two builds of libwvb200.so on the SAME box (boxes differ by ~8 %). Arms given as
environment settings need a library built with -DWVB_DEBUG_KNOBS (the shipped one ignores them):
    WVB_LIB_OUT=$PWD/wayverb_b200/libwvb200_dbg.so WVB_NVCC_DEFS=-DWVB_DEBUG_KNOBS python -m wayverb_b200.build
    WVB_LIB=$PWD/wayverb_b200/libwvb200_dbg.so python tools/ab_lib.py default WVB_WG_BPIPE=2
It alternates the libraries in fresh processes and prints the step time of each.
usage: ab_lib.py A B [C ...] [rounds]   where each arm is a library path or a
comma-separated list of environment settings (WVB_WG_BPIPE=4,WVB_WG_BMINB=4)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, json, numpy as np
sys.path.insert(0, %r)
import wayverb_b200 as wvb
from wayverb_b200 import _lib
s = json.load(open(%r))["sets"][0]["impedance"]
c = np.zeros((), _lib.COEFF_DT); c["b"], c["a"] = s["b"], s["a"]
m = wvb.cuboid_mesh((512, 512, 512), [c])
with wvb.Waveguide(m) as g:
    g.write(m.index(256, 256, 256), 1.0)
    g.time_steps(20)
    best = min(g.time_steps(100)[0] for _ in range(3)) / 100
    k = g.time_kernels(100)
    print(json.dumps({"step_ms": best, "air_ms": k[0] / 100, "bnd_ms": k[1] / 100}))
''' % (ROOT, os.path.join(ROOT, "tests", "golden", "lrs_coefficients.json"))


def main():
    args = sys.argv[1:]
    rounds = 3
    if args and args[-1].isdigit():
        rounds = int(args.pop())
    for r in range(rounds):
        for lib in args:
            if "=" in lib or lib == "default":
                env = dict(os.environ)
                env.update(kv.split("=", 1) for kv in lib.split(",") if "=" in kv)
            else:
                env = dict(os.environ, WVB_LIB=os.path.abspath(lib))
            out = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
            print(r, os.path.basename(lib), out.stdout.strip() or out.stderr[-300:], flush=True)


if __name__ == "__main__":
    main()
