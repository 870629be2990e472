import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """Committed golden vectors generated from the reference itself (tests/golden/make_golden.py)."""
    import numpy as np
    path = os.path.join(HERE, "golden", "aln_golden.npz")
    return np.load(path, allow_pickle=False)


@pytest.fixture(scope="session")
def small_index(golden):
    import refload as R
    T = golden["genome"]
    return T, R.bwa.index.build_index(T)
