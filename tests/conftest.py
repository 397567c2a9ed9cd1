import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _load_build_module():
    # build.py is loaded by path: importing the package itself requires the library to exist already
    spec = importlib.util.spec_from_file_location("_lr_build", os.path.join(ROOT, "multimodal_lipread_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The CUDA library is built in-tree (nvcc cross-compiles without a GPU); a prebuilt, up-to-date
    # .so is left alone.
    _load_build_module().build()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("this test is marked gpu and needs a CUDA device")
    return torch.device("cuda:0")
