import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the product library and the oracle once per session (nvcc cross-compiles on CPU)."""
    from gsasr_b200 import build as gbuild
    from oracle import oracle

    gbuild.build()
    oracle.build()
    yield


def golden(name):
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", name), allow_pickle=False)
