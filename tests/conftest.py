import os
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

ORACLE_DIR = REPO / "oracle"
REF_SO = ORACLE_DIR / "_ref" / "libhbref.so"
PORT_SO = ORACLE_DIR / "_ref" / "liboracle_port.so"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu():
    try:
        import ctypes
        from handbrake_b200 import LIBHBCU
        lib = ctypes.CDLL(str(LIBHBCU))
        return lib.hbcu_device_count() > 0
    except OSError:
        return False


@pytest.fixture(scope="session")
def ref():
    """the reference's own filters (oracle/_ref/libhbref.so), built from /root/reference when present"""
    from handbrake_b200.hblib import FilterLib
    if not REF_SO.exists():
        pytest.skip("oracle/_ref/libhbref.so not built (needs /root/reference at build time)")
    return FilterLib(REF_SO)


@pytest.fixture(scope="session")
def cuda_filters():
    import handbrake_b200
    if not _have_gpu():
        pytest.fail("GPU test selected but no CUDA device is visible: there is no CPU fallback")
    return handbrake_b200.filters()
