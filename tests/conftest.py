"""pytest configuration: markers and shared fixtures.

`-m "not gpu"` runs everywhere (oracle vs golden vectors, host logic on a NumPy fake of the C ABI,
ABI symbol checks); `-m gpu` needs a B200 and runs the parity tests through the C ABI.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def unpack_folds(flat, offs):
    parts = [flat[offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]
    return [(parts[2 * i], parts[2 * i + 1]) for i in range(len(parts) // 2)]


@pytest.fixture(scope="session")
def golden():
    return load_golden
