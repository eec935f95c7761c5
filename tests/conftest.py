import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_devices():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests are skipped (not failed) on a box without a CUDA device, so the plain
    `pytest tests/` run is green on the CPU container as well as `-m "not gpu"`."""
    if _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (gpu-marked test; the library has no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
