"""Step rate of a launch-bound mesh (BASELINE config 1: 33 x 27 x 22 nodes, the
5 x 4 x 3 m shoebox at a 500 Hz cutoff) with and without CUDA-graph batching, and of
the device-side run loop vs the per-step callback loop."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    import wayverb_b200 as wvb
    from wayverb_b200 import _lib
    c = np.zeros((), _lib.COEFF_DT)
    r = np.sqrt(0.9)
    c["b"][0], c["a"][0] = (1 + r) / (1 - r), 1.0
    for dims in ((33, 27, 22), (61, 50, 38)):
        m = wvb.cuboid_mesh(dims, [c])
        n = dims[0] * dims[1] * dims[2]
        for graph in (1, 0):
            os.environ["WVB_WG_GRAPH"] = str(graph)
            with wvb.Waveguide(m) as g:
                src = m.index(dims[0] // 2, dims[1] // 2, dims[2] // 2)
                g.write(src, 1.0)
                g.step(50)
                t0 = time.perf_counter()
                g.step(4000)
                dt = time.perf_counter() - t0
                sig = np.zeros(4000)
                sig[0] = 1.0
                g.run_device(src, sig[:64], [src])
                t0 = time.perf_counter()
                g.run_device(src, sig, [src, src + 3])
                dt2 = time.perf_counter() - t0
                print("dims %s graph=%d: step() %.2f us/step (%.0f Mnode/s), run_device %.2f us/step" %
                      (dims, graph, dt / 4000 * 1e6, n * 4000 / dt / 1e6, dt2 / 4000 * 1e6), flush=True)
        os.environ["WVB_WG_GRAPH"] = "1"
        trace = []
        t0 = time.perf_counter()
        wvb.run(m, wvb.hard_source(m.index(dims[0] // 2, dims[1] // 2, dims[2] // 2), [1.0] + [0.0] * 999),
                wvb.node_receiver(m.index(5, 5, 5), trace))
        dt3 = time.perf_counter() - t0
        print("dims %s: callback run() %.2f us/step (create included)" % (dims, dt3 / 1000 * 1e6), flush=True)


if __name__ == "__main__":
    main()
