import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    here = os.path.join(ROOT, "tests", "golden")
    return {n: np.load(os.path.join(here, f"{n}.npz")) for n in ("ref_sparse", "ref_dense", "ref_edge", "ref_c1_256")}
