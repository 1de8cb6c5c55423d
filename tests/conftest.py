import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built_lib():
    """The C-ABI library, (re)built if sources changed and nvcc is present."""
    from gan_playground_b200 import build

    if os.path.exists("/usr/local/cuda/bin/nvcc"):
        return build.build()
    assert os.path.exists(build.LIB), "libgpb200.so missing and no nvcc to build it"
    return build.LIB


def load_golden(name):
    import torch

    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def unpack_grads(d):
    return {k: v["q"].float() * v["scale"] for k, v in d.items()}
