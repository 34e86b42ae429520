import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_n21.npz"))


@pytest.fixture(scope="session")
def product_lib():
    """The CUDA library must exist in-tree; build it if this checkout has not been built yet."""
    import landing_controller_b200 as lc
    if not os.path.exists(lc.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return lc.load_library()
