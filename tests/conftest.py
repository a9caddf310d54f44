import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """liboracle.so -- this repo's CPU restatement of the reference path (test infrastructure)."""
    from oracle.binding import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref_engine():
    """The unmodified reference compiled into oracle/_ref (absent if that build did not happen)."""
    from oracle import binding
    if not binding.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference, present only in the build container)")
    return binding.RefEngine()


@pytest.fixture(scope="session")
def ref_cli():
    from oracle import binding
    if not binding.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference, present only in the build container)")
    return binding.RefCli()


@pytest.fixture()
def handle():
    """A fresh icp_handle on cuda:0.  Fails loudly (no skip) if the CUDA library is missing."""
    from iterativeclosestpoint_b200.engine import Handle
    h = Handle(0)
    yield h
    h.close()
