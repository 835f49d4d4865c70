import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "triton-vm_b200", "py"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def backend():
    import tvm_b200
    if not os.path.exists(tvm_b200.LIB_PATH):
        tvm_b200.build()
    b = tvm_b200.Backend(0)  # raises loudly without a GPU: there is no CPU fallback
    yield b
    b.close()


def rand_bfes(rng, shape):
    """uniform canonical field elements (numpy Generator)"""
    import numpy as np
    P = (1 << 64) - (1 << 32) + 1
    a = rng.integers(0, P, size=shape, dtype=np.uint64, endpoint=False)
    return a
