import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as ge
    return ge.load_package()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))


@pytest.fixture(scope="session")
def orc():
    import oracle_lib as ol
    o = ol.oracle()
    o.set_threads(1)
    return o


@pytest.fixture(scope="session")
def be(pkg):
    """One backend (device 0) for the whole GPU session.  No CPU fallback: creation raises without a B200."""
    b = pkg.Backend(0)
    yield b
    b.close()
