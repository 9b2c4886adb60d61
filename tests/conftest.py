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
def oracle_built():
    import oracle
    oracle.build()
    return oracle


def rel_block_err(a, b):
    """Per-row max |a-b| relative to the row's inf-norm of b (SURVEY 8(d) parity check)."""
    a = np.asarray(a).reshape(a.shape[0], -1)
    b = np.asarray(b).reshape(b.shape[0], -1)
    scale = np.maximum(np.max(np.abs(b), axis=1), 1e-300)
    return np.max(np.abs(a - b), axis=1) / scale
