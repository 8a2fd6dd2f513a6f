import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200)")
    config.addinivalue_line("markers", "slow: long-running statistical test")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) on a machine without CUDA.  On a GPU box nothing is skipped: a missing
    libnkfb200.so must fail loudly there (there is no CPU fallback to fall through to)."""
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200): torch.cuda.is_available() is False")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
