import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT, os.path.join(ROOT, "multiple-object-tracking_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    import oraclelib
    oraclelib.build_port()
    return oraclelib.Oracle("port")


@pytest.fixture(scope="session")
def ref():
    import oraclelib
    if not oraclelib.have_ref():
        if os.path.isdir("/root/reference"):
            oraclelib.build_ref()
        else:
            pytest.skip("oracle/_ref not prebuilt and /root/reference absent")
    return oraclelib.Oracle("ref")


@pytest.fixture(scope="session")
def oracle():
    """The strongest oracle available: the compiled reference if prebuilt, else the C restatement."""
    import oraclelib
    oraclelib.build_port()
    return oraclelib.Oracle(oraclelib.best())
