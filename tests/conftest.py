import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def alley_pair(golden_dir):
    import cv2
    a = cv2.imread(os.path.join(golden_dir, "alley_0001_gray.png"), cv2.IMREAD_UNCHANGED)
    b = cv2.imread(os.path.join(golden_dir, "alley_0002_gray.png"), cv2.IMREAD_UNCHANGED)
    assert a is not None and a.ndim == 2
    return a, b
