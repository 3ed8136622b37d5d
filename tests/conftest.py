import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    import mpskit_b200 as m
    return m.default_context()


@pytest.fixture(scope="session")
def dbg_ctx():
    """context that routes every contraction through the one-thread-per-element cross-check kernel"""
    import mpskit_b200 as m
    from mpskit_b200 import _lib
    return m.Context(0, _lib.FLAG_DEBUG_GEMM)
