import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU parity oracle (test infrastructure; see oracle/)."""
    from oracle import pyoracle

    if os.environ.get("BB200_ORACLE_NATIVE"):  # check that the -march=native timing build gives the same answers
        assert pyoracle.use_native_build()
    pyoracle.build()
    return pyoracle
