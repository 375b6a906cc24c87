import os
import sys

import pytest

# The oracle is OpenMP code with many short parallel regions: on a box whose visible core count exceeds what the
# container may use, a full-width team spins at every barrier. A small passive team is faster everywhere.
os.environ.setdefault("OMP_NUM_THREADS", str(max(1, min(8, len(os.sched_getaffinity(0))))))
os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _has_gpu():
    try:
        from isaac_ros_nvblox_b200 import _lib
        return _lib.load().nvb_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def built():
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def gpu(built):
    if not _has_gpu():
        pytest.fail("this test is marked gpu but no CUDA device is visible (no CPU fallback exists)")
    return True
