import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def built():
    """Make sure the product library and the C oracle exist (compiles only when stale or missing)."""
    from pheniqs_b200.build import build
    build()
    from oracle import oracle
    if not os.path.exists(oracle.PORT_LIBRARY):
        oracle.build("port")
    if not oracle.ref_available() and os.path.exists("/root/reference/pamld.cpp"):
        oracle.build("ref")
    return True
