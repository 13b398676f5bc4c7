import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "requires_reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    ref = os.path.isdir("/root/reference/HHI/models")
    has_timeout = config.pluginmanager.hasplugin("timeout")
    for item in items:
        if "gpu" in item.keywords and has_timeout and item.get_closest_marker("timeout") is None:
            item.add_marker(pytest.mark.timeout(180))      # a GPU-side deadlock must fail the test, not hang the box
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "requires_reference" in item.keywords and not ref:
            item.add_marker(pytest.mark.skip(reason="/root/reference not present"))
