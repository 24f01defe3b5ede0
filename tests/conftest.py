import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


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
def golden_dir():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="session")
def fold_on_disk(tmp_path_factory):
    """The seeded synthetic fold of tests/cases.py::FOLD_ARGS in the reference's on-disk layout."""
    import cases
    from multimodal_error_detection_b200 import synthetic
    fold = synthetic.make_fold(**cases.FOLD_ARGS)
    path = tmp_path_factory.mktemp("fold")
    return synthetic.write_fold(fold, str(path)) + "/", fold
