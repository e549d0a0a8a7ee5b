import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")
    # the product library is built in-tree (nvcc cross-compiles without a GPU); a fresh
    # clone has none yet. The oracle builds itself on first use (oracle/wgo.py, rto.py).
    from wayverb_b200 import build as wvb_build
    wvb_build.build_lib(force=False)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
