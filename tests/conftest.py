import os
import sys
import warnings

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    warnings.filterwarnings("ignore", message="no CLIP weights given")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def weights_dir():
    return os.path.join(ROOT, "lossyless_b200", "hub", "weights")


TAGS = ["beta1e-01", "beta5e-02", "beta1e-02"]


@pytest.fixture(scope="session")
def native_lib():
    """The built shared library; CPU tests only touch its host-side entry points."""
    from lossyless_b200 import build, _native

    build.build()
    return _native.load()


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("a test marked gpu ran without a GPU")
    from lossyless_b200 import _native

    _native.check(_native.load().ll_check_device(0), "ll_check_device")
    return torch.device("cuda:0")
