"""Test configuration: `gpu` marker, repo root on sys.path, library build once per session."""

import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def built_lib():
    """libfava_b200.so, (re)built in-tree if stale.  nvcc cross-compiles without a GPU."""
    from fava_b200.build import build_library

    return build_library()


@pytest.fixture(scope="session")
def cuda_device(built_lib):
    import torch

    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test selected but no CUDA device is visible")
    return torch.device("cuda", 0)
