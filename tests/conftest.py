import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))


@pytest.fixture(params=[0, 1, 2])
def seed(request):
    """The reference's rng fixture seeds (tests/conftest.py:14-16)."""
    return request.param


def fixture_data(seed):
    """data / missing_data fixtures of the reference (tests/conftest.py:19-21, tests/test_gpu.py:16-20)."""
    rng = np.random.default_rng(seed)
    data = (rng.uniform(size=(10, 1000)) < 0.05).astype(np.int8)
    missing = data.copy()
    inds = rng.integers(0, missing.size, size=int(0.01 * missing.size))
    missing.flat[inds] = -1
    return data, missing.clip(-1, 1)
