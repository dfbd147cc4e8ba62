import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(autouse=True, scope="session")
def _bounded_cpu_threads():
    """The CPU oracle (torch convolutions on 64x64 images) is fastest with 8-16 intra-op threads and collapses on the 100+
    cores of a GPU box; every test that runs it gets a bounded pool."""
    import torch

    torch.set_num_threads(min(16, max(1, os.cpu_count() or 1)))
    yield
