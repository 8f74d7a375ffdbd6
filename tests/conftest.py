import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def xf():
    from yuki_b200 import transforms
    return transforms


@pytest.fixture(scope="session")
def gpu_ctx():
    from yuki_b200 import api
    ctx = api.Context(0)  # raises YukiGpuError when no device: GPU tests must not pass on a fallback
    yield ctx
    ctx.close()


def rel_rmse(a, b):
    """Parity metric of SURVEY.md §8d: sqrt(mean((g-o)^2)) / mean(o) over RGB."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.sqrt(np.mean((a - b) ** 2)) / max(np.mean(b), 1e-30))
