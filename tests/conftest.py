import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


REF_PY = os.path.join(ROOT, "baseline", "_ref")
if os.path.isdir(os.path.join(REF_PY, "tskit")) and REF_PY not in sys.path:
    sys.path.insert(0, REF_PY)  # the unmodified reference Python package (drop-in tests)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


HAS_GPU = _has_gpu()


def pytest_collection_modifyitems(config, items):
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def wf_small():
    from tskit_b200.tables import Tables
    return Tables.load(os.path.join(ROOT, "tests", "data", "wf_200_500_100000.npz")).ensure_derived()


@pytest.fixture(scope="session")
def wf_1k():
    from tskit_b200.tables import Tables
    return Tables.load(os.path.join(ROOT, "tests", "data", "wf_1000_4000_10000000.npz")).ensure_derived()
