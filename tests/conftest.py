import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def small_cfg():
    """A reduced WDSR (2 blocks) that keeps every layer type of cfg/p16t9c85r12."""
    return dict(scale=3, numFilters=32, kernelSize=(3, 3, 3), numResBlocks=2, expRate=8, decayRate=0.8,
                numImgLR=9, patchSizeLR=16, isGrayScale=True)


@pytest.fixture(scope="session")
def full_cfg():
    return dict(scale=3, numFilters=32, kernelSize=(3, 3, 3), numResBlocks=12, expRate=8, decayRate=0.8,
                numImgLR=9, patchSizeLR=16, isGrayScale=True)
