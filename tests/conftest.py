import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected with -m gpu; when they are selected on a box without a GPU they must fail loudly rather
    # than skip (a silent skip would read as "parity green").  A plain `pytest tests` (no -m expression) on a box
    # without a CUDA device deselects them instead, so the CPU suite can be run without remembering -m "not gpu".
    if config.getoption("markexpr"):
        return
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    keep, drop = [], []
    for it in items:
        (drop if it.get_closest_marker("gpu") else keep).append(it)
    if drop:
        config.hook.pytest_deselected(items=drop)
        items[:] = keep


@pytest.fixture(scope="session")
def repo_root():
    return ROOT
