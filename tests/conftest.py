import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def _ensure_built():
    """Build the product library and the checkers if they are missing (CPU-only compile)."""
    from elementary_b200.runtime import LIB_PATH
    from oracle import oracle as orc
    if not (os.path.exists(LIB_PATH) and orc.port_available()):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session", autouse=True)
def built():
    _ensure_built()
    yield
