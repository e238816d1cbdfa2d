import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100a device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture
def deterministic():
    """One tcgen05.mma issuing warp in the GEMM / conv kernel for the duration of a test that compares two runs bit for bit
    (the default two-issuer mode accumulates in an order the hardware does not fix: last bits may differ between runs)."""
    import torch
    if not torch.cuda.is_available():
        yield
        return
    from ivideogpt_b200 import ops
    ops.set_deterministic(True)
    try:
        yield
    finally:
        ops.set_deterministic(False)
