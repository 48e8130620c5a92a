import os
import sys

# The in-process multi-shard tests run several handles (4 streams each) with spin-wait flag barriers on
# ONE device: give every stream its own hardware queue so a spinning barrier cannot block a peer's kernels.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# ... and load all kernels up front: a lazy module load synchronises the context and would dead-lock against
# another in-process shard's spinning barrier (not an issue with one process per GPU).
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
