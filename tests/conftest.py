import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "closed_forms.npz"))


@pytest.fixture(scope="session")
def orc():
    import oracle as o  # oracle/oracle.py (test infrastructure)
    o.build()
    return o


@pytest.fixture(scope="session")
def P():
    import pdmp_b200
    return pdmp_b200
