import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` on the GPU box)")


@pytest.fixture(scope="session")
def worker():
    """One b200zk context on cuda:0.  The product has no CPU fallback: without a GPU this raises."""
    import zcash_gpu_thesis_b200 as zk

    w = zk.Worker(0)
    yield w
    w.close()
