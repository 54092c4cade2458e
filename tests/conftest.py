import os
import sys

import pytest

# the peer-exchange tests run several spin-waiting kernels of one process concurrently: every stream needs its own
# hardware queue (two streams sharing one would order a rank's kernel behind the kernel that waits for it).
# Read by the driver when the CUDA context is created, so it is set before anything imports torch.cuda.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
