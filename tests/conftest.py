import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs the live reference under /root/reference")


def _cuda_ok():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    from oracle import refshim

    have_ref = refshim.available()
    have_gpu = None
    for item in items:
        if "ref" in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason="live reference not present"))
        if "gpu" in item.keywords:
            if have_gpu is None:
                have_gpu = _cuda_ok()
            if not have_gpu:
                item.add_marker(pytest.mark.skip(reason="no CUDA device"))
