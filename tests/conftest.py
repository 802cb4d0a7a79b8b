import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def native():
    """Build (if needed) and import the native library; never falls back to anything."""
    from pysparse_b200 import build
    build.build_all()
    from pysparse_b200 import _abi
    return _abi


@pytest.fixture(scope="session")
def port():
    import oracle
    oracle.build(ref=False)
    return oracle.Oracle("port")


@pytest.fixture(scope="session")
def ref():
    import oracle
    if os.path.isdir("/root/reference/pysparse"):
        oracle.build(ref=True)
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libpsp_ref.so not available (needs /root/reference to build)")
    return oracle.Oracle("ref")


@pytest.fixture(scope="session")
def checker(port):
    """The strongest CPU checker available: the compiled reference, else the restatement."""
    import oracle
    return oracle.Oracle("ref") if oracle.have_ref() else port
