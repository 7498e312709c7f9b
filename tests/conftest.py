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
    """The CPU oracle (test infrastructure only)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import newton_oracle
    newton_oracle.lib()
    return newton_oracle


@pytest.fixture(scope="session")
def pf():
    import cracks_b200
    cracks_b200.build_library()
    return cracks_b200
