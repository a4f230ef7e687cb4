import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "8")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_small.npz"))


@pytest.fixture(scope="session")
def ref32():
    from oracle import ref_lib
    if not ref_lib.available(32):
        pytest.skip("oracle/_ref/libref32.so not built")
    return ref_lib.RefLib(32)
