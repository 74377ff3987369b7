import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (GPU tests run under gpurun)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): builds oracle/librfx_oracle.so on first use."""
    from oracle import pyoracle
    pyoracle.build(ref=False)
    return pyoracle


@pytest.fixture(scope="session")
def rfx_lib():
    """The product's C-ABI library, built in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    from reflaxman_b200 import build, capi
    build.build()
    return capi.load()
