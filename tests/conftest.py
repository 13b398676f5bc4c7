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
    from oracle import ref_shims
    ref = ref_shims.reference_available()      # the source tree, or its byte-compiled modules under oracle/_ref (GPU box)
    has_timeout = config.pluginmanager.hasplugin("timeout")
    for item in items:
        if "gpu" in item.keywords and has_timeout and item.get_closest_marker("timeout") is None:
            item.add_marker(pytest.mark.timeout(180))      # a GPU-side deadlock must fail the test, not hang the box
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "requires_reference" in item.keywords and not ref:
            item.add_marker(pytest.mark.skip(reason="neither /root/reference nor oracle/_ref present"))


@pytest.fixture(autouse=True)
def _reset_dropout_epoch(request):
    """Trainers switch the library's device-resident dropout epoch on for the whole process (and advance it); the tests
    that regenerate the documented masks on the host assume epoch 0.  Reset before every GPU test (trainers re-enable it in
    their constructors)."""
    if "gpu" in request.keywords:
        import torch
        if torch.cuda.is_available():
            from egot2_b200 import _lib as L
            L.call("egot2_dropout_epoch_host", 0)
            L.call("egot2_dropout_epoch_enable", 0)
    yield
