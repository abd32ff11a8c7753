import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line("markers", "slow: CPU test that takes more than ~30 s")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests need a CUDA device and the built library: skip (not error) when either is missing."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    lib = os.path.exists(os.path.join(ROOT, "samrs_b200", "libsamrs_b200.so"))
    if have and lib:
        return
    skip = pytest.mark.skip(reason="no CUDA device" if not have else "libsamrs_b200.so not built")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
