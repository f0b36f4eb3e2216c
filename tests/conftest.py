import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a host without a CUDA device skips the `gpu` tests instead of failing in them (the product
    has no CPU path, so there is nothing they could run on).  On the GPU box nothing is skipped."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device: `gpu` tests run on the B200 box (python -m pytest tests -m gpu)')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)
