import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def lib_built():
    """Build libspittle_b200.so in-tree if it is missing (nvcc cross-compiles without a GPU)."""
    from spittle_b200 import build, capi
    if not os.path.exists(capi.LIB_PATH):
        build.build()
    return capi.LIB_PATH


@pytest.fixture(scope="session")
def cuda_dev(lib_built):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test running without a CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="session")
def model_dir(tmp_path_factory):
    d = os.environ.get("SB_MODEL_DIR") or str(tmp_path_factory.mktemp("models"))
    return d
