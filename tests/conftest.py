import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


# BASELINE.json north_star tolerances (relative L2 against the oracle): the 16-bit tensor-core path, the FP32 CUDA-core validation mode
TOL_16BIT = 5e-3
TOL_FP32 = 1e-5


def rel_l2(a, b):
    import numpy as np
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.fixture(scope="session")
def emd():
    import denoiser
    return denoiser.emd
