import os
import sys

import numpy as np
import pytest

# Several shards of one field on ONE GPU (the single-GPU tests of the peer-to-peer halo protocol) wait for each other from inside
# their kernels, so their streams must never share a hardware queue: CUDA's default is 8 queues per device, and two streams that
# alias on one queue would serialise a waiting kernel before the kernel it waits for.  One queue per stream (the maximum is 32).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "vignette_golden.json")) as f:
        return json.load(f)
