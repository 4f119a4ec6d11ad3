import os
import sys
import pytest

# The data-parallel tests run several RANKS of one process on one GPU, each on its own streams, and the ranks' kernels wait on one another.
# Streams of a process share the device's hardware work queues (8 by default): two streams mapped onto one queue execute in order, so a kernel
# could be parked behind the very kernel that waits for it.  32 queues keep the ranks' streams apart (must be set before the CUDA context exists;
# one process per GPU — the product's layout — never meets this).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """-m gpu tests are skipped (not failed) when collected on a box without a GPU."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(autouse=True)
def _release_device_buffers():
    yield
    try:
        import gpu_util
        gpu_util.release()
    except Exception:
        pass
