import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session")
def c_oracle():
    from oracle import oracle as O
    return O.COracle()


@pytest.fixture(scope="session")
def golden_cpu():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "dq_refcpu.npz"))


@pytest.fixture(scope="session")
def golden_gpu():
    import numpy as np
    p = os.path.join(ROOT, "tests", "golden", "dq_refgpu.npz")
    if not os.path.exists(p):
        pytest.skip("tests/golden/dq_refgpu.npz not generated yet (tests/golden/make_golden_gpu.py on a B200)")
    return np.load(p)
