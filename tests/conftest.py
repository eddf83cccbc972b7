import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        # a kernel that never returns must not hold the GPU box until the outer limit: pytest-timeout's "thread" method
        # ends the process (and with it the CUDA context) even when the main thread is blocked inside the driver
        try:
            import pytest_timeout  # noqa: F401
            for item in items:
                if "gpu" in item.keywords and item.get_closest_marker("timeout") is None:
                    item.add_marker(pytest.mark.timeout(1500, method="thread"))
        except ImportError:
            pass
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def built_lib():
    """Path of the in-tree library; builds it when the tree has none (nvcc cross-compiles without a GPU)."""
    from simt_b200 import build
    return build.build()
