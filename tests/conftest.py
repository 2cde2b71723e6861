import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: large-size case (still part of -m gpu)")


@pytest.fixture(autouse=True)
def _current_device_is_gpu0(request):
    """GPU tests assume cuda:0 is current when they start (streams and torch allocations follow
    the current device); a multi-GPU test that made another device current must not leak that
    into the next test, whatever the order the files run in."""
    if request.node.get_closest_marker("gpu") is not None:
        import torch
        if torch.cuda.is_available():
            torch.cuda.set_device(0)
    yield
