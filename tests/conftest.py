import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    # Build the product library and the oracle once per session (cross-compiles without a GPU).
    import __graft_entry__

    __graft_entry__.build()


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    return torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _auto_kernel_variant(request):
    """GPU tests may force a kernel variant (gsdrB200SetKernelVariant is process-wide): reset it around every test."""
    if "gpu" not in request.keywords:
        yield
        return
    import gsdr_b200 as g

    g.set_kernel_variant(-1)
    yield
    g.set_kernel_variant(-1)
