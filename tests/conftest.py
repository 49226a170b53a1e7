import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib():
    """The C-ABI library; GPU tests call the product through it."""
    from geossl_b200 import _lib
    return _lib.load()


@pytest.fixture(params=["simt", "tc_fp16"])
def filter_mode(request):
    """Runs a test once on the exact fp32 CUDA-core filter kernels and once on the tcgen05 tensor-core ones."""
    from geossl_b200 import ops
    old = ops.FILTER_MODE
    ops.FILTER_MODE = request.param
    yield request.param
    ops.FILTER_MODE = old
