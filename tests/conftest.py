import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a CPU-only box reports the gpu tests as skipped, not as 50 fixture errors.  On a box WITH a
    GPU nothing is skipped and a missing libasb200.so is a hard error (the product has no CPU fallback)."""
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device (run on the B200 box with -m gpu)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def engine():
    """One GPU engine for the whole session.  Fails loudly if the CUDA library or device is missing."""
    from amplicon_sorter_b200.engine import Engine

    eng = Engine(0)
    yield eng
    eng.close()
