import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def relerr(a, b, floor=1e-150):
    """max |a/b - 1| over entries where the reference b exceeds `floor`; entries below must also be below in a."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    big = np.abs(b) > floor
    r = 0.0
    if np.any(big):
        r = float(np.max(np.abs(a[big] / b[big] - 1.0)))
    return r


@pytest.fixture(scope="session")
def gpu_engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from acropolis_b200.input import shared_input_interface
    from acropolis_b200.engine import Engine
    return Engine.get(shared_input_interface())


@pytest.fixture(scope="session")
def xcheck_engine():
    """Engine bound to the TEST build of the library (-DACRO_CROSSCHECK): the product kernels plus the per-(E,E',T)
    cross-check modes kernel_mode = PLAIN_GL / PER_T, which the product library rejects."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from acropolis_b200.input import shared_input_interface
    from acropolis_b200.engine import Engine
    return Engine.get(shared_input_interface(), variant="xcheck")
