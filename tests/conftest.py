import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(autouse=True)
def _no_aborted_kernels(request):
    """After every GPU test: no tensor-core kernel gave up on a bounded wait (include/nic.h: nic_pipeline_status)."""
    yield
    if request.node.get_closest_marker("gpu") is not None:
        import torch
        if torch.cuda.is_available():
            from neural_image_compression_b200 import _lib
            assert _lib.load().nic_pipeline_status() == 0, "a tensor-core kernel aborted on an expired pipeline wait"
