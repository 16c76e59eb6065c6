import os
import sys

import pytest

# several slab ranks are driven from one process on separate streams (tests/test_gpu_slab.py): give every
# stream its own hardware queue so a kernel waiting for a peer never blocks that peer's launch
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# ... and load every kernel at start-up: a lazy module load blocks the host until running kernels finish,
# which would stall the launch of the very kernel a spinning peer rank is waiting for
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def rel_max(a, b):
    """max-norm relative difference used for every fp64 parity check."""
    import numpy as np

    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback)")
    return torch.device("cuda:0")


class no_gc_during_collective:
    """Several emulated ranks are enqueued from ONE host thread: while rank 0's kernels spin (bounded) on flags that
    rank 1's kernels will raise, the host must not block.  A garbage-collected context from an earlier test would call
    cudaFree (a device-wide synchronisation) right there, so collect first and keep the collector off meanwhile."""

    def __enter__(self):
        import gc

        gc.collect()
        gc.disable()

    def __exit__(self, *a):
        import gc

        gc.enable()
