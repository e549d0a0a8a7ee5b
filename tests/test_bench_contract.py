"""bench.py's control flow and the JSON contract of its line, on CPU: the GPU-facing pieces
(handles, the C++ probe, NVML, the ray row, the CPU baseline) are replaced by stand-ins, main()
runs, and the printed line must carry every key the driver reads. Also: a failing optional row
must not cost the headline on a single GPU."""
import contextlib
import importlib
import io
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeWG:
    def __init__(self):
        self.launches = 0

    def info(self):
        self.launches += 40
        return {"kernel_launches": self.launches, "kernel_variant": "tma", "tile": (128, 8, 12),
                "boundary_nodes": (1548384, 6096, 8), "device_bytes": 1, "halo": "none"}

    def write(self, *a):
        pass

    def time_steps(self, n):
        return 0.56 * n, 0

    def launch(self):
        return 0

    def read(self, n):
        return 0.0

    def swap(self):
        pass

    def run_device(self, src, sig, rcv):
        return len(sig), np.zeros((len(sig), 1)), 0

    def time_kernels(self, n):
        return 0.5 * n, 0.07 * n

    def close(self):
        pass


class FakeMesh:
    def index(self, x, y, z):
        return 1


def run_bench(monkeypatch, ray):
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    importlib.reload(bench)

    class Ctx(bench.Ctx):
        def init(self):
            class T:
                class cuda:
                    @staticmethod
                    def synchronize():
                        pass
            self.torch = T

    class Sampler:
        def __init__(self, index):
            pass

        def start(self):
            pass

        def mark(self, a, b):
            pass

        def summary(self):
            return {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "samples": 3, "reasons": []}

    class Probe:
        def wvb_probe_waveguide_run(self, *a):
            a[-2][0], a[-2][1] = 11.5, 200.0
            return 0

    monkeypatch.setattr(bench, "Ctx", Ctx)
    monkeypatch.setattr(bench, "ClockSampler", Sampler)
    monkeypatch.setattr(bench, "slab_handle", lambda ctx, gdims, coeffs, kernel=None: (FakeMesh(), FakeWG(), (0, gdims[2])))
    monkeypatch.setattr(bench, "probe_lib", lambda: Probe())
    monkeypatch.setattr(bench, "ray_row", ray)
    monkeypatch.setattr(bench, "cpu_baseline", lambda: {"value": 1.0, "unit": "Mnode-updates/s", "cores": 1,
                                                        "kind": "port", "sample": "stand-in"})
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "20", "--warmup", "5"])
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        bench.main()
    lines = [l for l in buf.getvalue().splitlines() if l.strip()]
    assert len(lines) == 1, "exactly ONE JSON line"
    return json.loads(lines[0])


def test_line_carries_the_contract(monkeypatch):
    d = run_bench(monkeypatch, lambda ctx, sampler, with_cpu: {"metric": "ray-reflections/s", "reflections_per_s": 1.0})
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 20 and d["warmup"] == 5 and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
    assert d["roofline"]["bound"] == "hbm" and d["roofline"]["bytes_per_node"] == 24.5
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-12
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert d["gpu_launches"] > 0
    # value = nodes / ms_per_step: 512^3 nodes at the stand-in's 0.56 ms
    assert abs(d["value"] - 512 ** 3 / 0.56 / 1e3) < 1e-6 * d["value"]
    assert "config4" in d and "slab256" in d and "ray" in d


def test_a_failing_optional_row_does_not_cost_the_headline(monkeypatch):
    def boom(ctx, sampler, with_cpu):
        raise RuntimeError("boom")
    d = run_bench(monkeypatch, boom)
    assert d["ray"] == {"error": "RuntimeError: boom"}
    assert d["value"] > 0 and "roofline" in d


def test_reference_arm_runs_on_the_host_and_prints_the_contract():
    """`bench.py --impl reference` for real (CPU only): the reference's own kernel source from
    oracle/_ref (or the oracle port where that is absent), every host thread even when the
    launcher exported OMP_NUM_THREADS=1 like torch.distributed.run does"""
    import subprocess
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                        "--warmup", "3"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"] == "Mnode-updates/s (fp64)" and d["unit"] == "Mnode-updates/s"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    threads = len(os.sched_getaffinity(0))
    assert d["cpu_baseline"]["cores"] == threads, "the arm must not inherit OMP_NUM_THREADS=1"
