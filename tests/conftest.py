import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def fds_bytes():
    with open(os.path.join(ROOT, "tests", "golden", "schemas.binpb"), "rb") as fh:
        return fh.read()


@pytest.fixture(scope="session")
def oracle(fds_bytes):
    import orc
    return orc.Schema(fds_bytes)


@pytest.fixture(scope="session")
def hsim(fds_bytes):
    import hostsim
    return hostsim.Schema(fds_bytes)


@pytest.fixture(scope="session")
def engine():
    import ggrmcp_b200
    e = ggrmcp_b200.Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="session")
def schema(engine, fds_bytes):
    return engine.register(fds_bytes)
