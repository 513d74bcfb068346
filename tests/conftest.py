import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TESTS = os.path.join(ROOT, "tests")
for p in (ROOT, TESTS):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(TESTS, "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    """-> (EpochTable holding only the kept epochs, list of sha256 hex digests, head bytes)"""
    from gps_sdr_sim_b200.table import EpochTable
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    table = EpochTable.load_npz(path)
    z = np.load(path)
    return table, [str(s) for s in z["sha256"]], z["head"]


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_required():
    if not has_gpu():
        pytest.skip("no CUDA device")
