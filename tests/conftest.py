import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # GPU tests are skipped (not failed) when no device is visible, e.g. in the build container.
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def example_data():
    """tests/example.train / example.test of the reference, committed verbatim as a golden fixture."""
    import numpy as np
    g = os.path.join(ROOT, "tests", "golden")
    tr = np.loadtxt(os.path.join(g, "example.train"))
    te = np.loadtxt(os.path.join(g, "example.test"))
    return (tr[:, 0].astype(np.int32), tr[:, 1].astype(np.int32), tr[:, 2].astype(np.float32),
            te[:, 0].astype(np.int32), te[:, 1].astype(np.int32), te[:, 2].astype(np.float32))
