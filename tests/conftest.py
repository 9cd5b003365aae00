import ctypes
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_lib():
    """CPU oracle (oracle/liboracle.so): the checker; built on demand with the committed Makefile."""
    path = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    return ctypes.CDLL(path)


@pytest.fixture(scope="session")
def cuda_lib():
    """Product library; fails loudly if it has not been built (no fallback of any kind)."""
    import eqtlbma_b200
    return eqtlbma_b200.load_library()
