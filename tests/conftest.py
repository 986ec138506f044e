import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def ggq():
    """The product library; building it is part of __graft_entry__.build()."""
    import gguf_b200
    from gguf_b200._lib import SO_PATH
    if not os.path.exists(SO_PATH):
        import __graft_entry__ as ge
        ge.build()
    return gguf_b200
