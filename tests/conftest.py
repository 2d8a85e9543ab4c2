import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def sample_iq():
    import numpy as np
    return np.fromfile(os.path.join(GOLDEN, "sample_cl78.bin"), dtype=np.uint8)


@pytest.fixture(scope="session")
def built():
    """Make sure the in-tree shared libraries exist (compiles on CPU, no GPU needed)."""
    import __graft_entry__ as g
    g.build()
    return True
