import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: long-running CPU oracle check")


@pytest.fixture(scope="session")
def net():
    """One GPU party for the whole session.  Fails (not skips) when the CUDA library or a GPU is missing:
    the product has no CPU fallback, and a silent skip would hide 'native code not loaded'."""
    from distributed_groth16_b200 import Net
    n = Net(0)
    n.use_torch_stream(0)          # slot 0 launches on torch's current stream: tensor ops and kernels stay ordered
    yield n
    n.close()


@pytest.fixture(scope="session")
def cref():
    from oracle import cref as c
    c.build()
    return c
