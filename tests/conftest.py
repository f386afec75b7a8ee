import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_metrics():
    return np.load(GOLDEN / "metrics_loss.npz")


@pytest.fixture(scope="session")
def golden_corrupt():
    return np.load(GOLDEN / "corrupt.npz")


@pytest.fixture(scope="session")
def golden_model():
    return np.load(GOLDEN / "model.npz")


@pytest.fixture(scope="session")
def lib():
    """The C-ABI library; the GPU tests call the product through it (via the msd_b200 shims or directly)."""
    import msd_b200
    return msd_b200._lib.load()
