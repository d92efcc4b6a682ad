import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def pytest_sessionstart(session):
    """The CUDA library is built in-tree and git-ignored: build it if this checkout does not have it yet
    (nvcc cross-compiles sm_100a without a GPU)."""
    lib = os.path.join(ROOT, "algebra_b200", "libalgebra_b200.so")
    if not os.path.exists(lib):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "algebra_b200", "csrc"), "-j8"])
