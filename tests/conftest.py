import os
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO / "stable-diffusion.cpp_b200"))
sys.path.insert(0, str(REPO))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu() -> bool:
    try:
        import ctypes
        from sdb200 import B200_SO
        lib = ctypes.CDLL(str(B200_SO))
        lib.ggml_backend_b200_get_device_count.restype = ctypes.c_int
        return lib.ggml_backend_b200_get_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def harness():
    from sdb200 import Harness
    return Harness()


@pytest.fixture(scope="session")
def cpu_oracle(harness):
    """The reference's own ggml CPU backend (checker)."""
    from oracle.cpu_ref import load_cpu_oracle
    load_cpu_oracle(harness)
    return harness


@pytest.fixture(scope="session")
def b200(cpu_oracle):
    """Harness with both the CPU oracle and the B200 plugin registered.  Fails loudly without the CUDA library/GPU."""
    devs = cpu_oracle.load_b200()
    return cpu_oracle, devs[0]
