import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """build libirt_b200.so (nvcc cross-compiles without a GPU) and the oracle if they are missing"""
    import irt_b200
    if not os.path.exists(irt_b200.LIB_PATH) or not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        import __graft_entry__ as ge
        ge.build()
    yield


@pytest.fixture(scope="session")
def orc():
    """canonical CPU oracle (test infrastructure)"""
    from oracle.oracle import Oracle, build
    build()
    return Oracle("canonical")


@pytest.fixture(scope="session")
def wl():
    import irt_b200.workloads as w
    return w
