import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# The pinned staging buffer of the pipelined upload is normally built by a background thread (the first calls of a
# process go by plain DMA meanwhile).  The tests want the host-conversion path from the first call on.
os.environ.setdefault("UMPA_STAGE_SYNC", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cuda_ok():
    import torch
    return torch.cuda.is_available()
