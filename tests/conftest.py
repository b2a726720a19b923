import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


@pytest.fixture
def cpu_backend():
    """Install the oracle-backed checker backend (host-logic tests on a machine without a GPU)."""
    import vae_connexe_b200.lib as L
    from tests.cpu_backend import OracleKernels
    prev = L._kernels
    L.set_test_backend(OracleKernels())
    yield L._kernels
    L.set_test_backend(prev)
