import os
import sys

import pytest

# tests/test_gpu_mg.py and the CLI's SVINET_SHARDS_ON_ONE_GPU mode put several shards -- three streams each, with
# flag-wait kernels at their heads -- on ONE device; with the default 8 hardware queues, streams alias and a waiting
# kernel can block the very stream it waits for.  (A real multi-GPU run has one shard per device.)  Must be set before
# CUDA initialises.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")
