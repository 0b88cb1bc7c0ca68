import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False


def pytest_collection_modifyitems(config, items):
    # a GPU test on a box without a GPU is an error in how the suite was invoked, except when the
    # whole suite is run unfiltered on the CPU container: then GPU tests are skipped, not failed.
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device (GPU tests run with -m gpu on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
