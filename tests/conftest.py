import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ref_raymarching():
    from oracle.build_ref import load_ref
    return load_ref("_ref_raymarching")


@pytest.fixture(scope="session")
def ref_gridencoder():
    from oracle.build_ref import load_ref
    return load_ref("_ref_gridencoder")


@pytest.fixture(scope="session")
def ref_shencoder():
    from oracle.build_ref import load_ref
    return load_ref("_ref_shencoder")
