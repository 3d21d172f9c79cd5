import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "toy-robust-backend-slam_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session", autouse=True)
def built():
    """Everything native is built once per session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as ge
    ge.build()
    return True


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_case(name):
    import numpy as np
    from dcs_b200 import Graph
    path = os.path.join(GOLDEN, name + ".npz")
    return Graph.load_npz(path), np.load(path)
