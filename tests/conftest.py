import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle_mod():
    """The parity checkers (oracle/), built on demand.  Test infrastructure only."""
    from oracle import oracle
    if not oracle.have_port():
        oracle.build()
    return oracle


@pytest.fixture(scope="session")
def ref_fixtures():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "ref_fixtures.npz")
    return np.load(path, allow_pickle=False)
