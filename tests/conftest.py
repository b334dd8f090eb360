import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_ok():
    try:
        import torch
        # the PyTorch references in the GPU tests must be true fp32 (no TF32 convolutions / matmuls)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_ok():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def rb():
    """The product package (its directory name is not an identifier, hence the alias module)."""
    return importlib.import_module("resenc_b200")


@pytest.fixture(scope="session")
def built_lib():
    import __graft_entry__ as g
    g.build()
    return importlib.import_module("resenc_b200")._lib.load()
