import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the built libraries travel with the repo snapshot; on a clean checkout compile them first
    # (nvcc cross-compiles sm_100a without a GPU) -- the loader itself never falls back to a CPU path
    if not os.path.exists(os.path.join(ROOT, "vstrains_b200", "libvspe.so")):
        import __graft_entry__
        __graft_entry__.build()


class Golden:
    def __init__(self, path):
        z = np.load(path)
        self.name = os.path.basename(path)[:-4]
        self.gfa = z["gfa"].tobytes()
        self.fwd = z["fwd"].tobytes()
        self.rve = z["rve"].tobytes()
        self.k = int(z["k"])
        self.status = int(z["status"])
        self.pe_info = z["pe_info"].tobytes()
        self.st_info = z["st_info"].tobytes()

    def __repr__(self):
        return self.name


def golden_cases():
    return [Golden(p) for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))]


def pytest_generate_tests(metafunc):
    if "golden" in metafunc.fixturenames:
        cases = golden_cases()
        metafunc.parametrize("golden", cases, ids=[c.name for c in cases])


def parse_info(data: bytes, ids):
    """pe_info/st_info text -> dense int64 matrix (test helper)."""
    n = len(ids)
    m = np.zeros((n, n), dtype=np.int64)
    lines = data.decode().split("\n")
    assert lines[-1] == ""
    assert len(lines) - 1 == n * n
    for t, line in enumerate(lines[:-1]):
        m[t // n, t % n] = int(line.rsplit(":", 1)[1])
    return m
