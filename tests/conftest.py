import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def fv():
    import __graft_entry__ as g
    g.build_library()
    return g.load_package()


@pytest.fixture(scope="session")
def orc():
    from oracle import fv_oracle
    fv_oracle.build()
    return fv_oracle


@pytest.fixture(scope="session")
def fourfractures():
    import numpy as np
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "fourfractures.npz")))
