import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def orc():
    from oracle import orc as _orc
    _orc.load()
    return _orc


@pytest.fixture(scope="session")
def rlr():
    """The product library, through its ctypes binding (built in-tree if stale)."""
    import rust_local_rag_b200  # noqa: F401
    from rust_local_rag_b200 import binding
    binding.load()
    return binding
