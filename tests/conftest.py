import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """native pieces are built once per session (no-op when up to date)"""
    import __graft_entry__ as ge
    ge.build()
    return True


@pytest.fixture(scope="session")
def orc(built):
    from oracle.oracle import Oracle
    return Oracle("libm")


@pytest.fixture(scope="session")
def orc_pm(built):
    from oracle.oracle import Oracle
    return Oracle("pm")


@pytest.fixture(scope="session")
def ref(built):
    from oracle.oracle import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref not built (reference tree absent and no prebuilt copy)")
    return Reference()


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
