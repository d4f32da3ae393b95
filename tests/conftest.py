import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run under gpurun on a B200)")
    config.addinivalue_line("markers", "reference: needs /root/reference (dev container only)")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir(REFERENCE)
    skip_ref = pytest.mark.skip(reason="/root/reference not present on this box")
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def cuda_dev():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test running without a CUDA device")
    return torch.device("cuda:0")
