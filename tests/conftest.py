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
    if os.environ.get("PF_EMULATED_LIBRARY") == "1":
        # developer aid for a box without a GPU: `PF_EMULATED_LIBRARY=1 pytest tests/test_gpu_parity.py` runs the
        # GPU parity tests through the CPU emulation of the library (tests/emu/, test infrastructure) to catch
        # regressions of the host logic / kernel sources before GPU time is spent.  Never set by the suites.
        sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
        import build_emulated_library
        so = build_emulated_library.build()
        from cracks_b200 import api
        api.library_path = lambda: so
        api._LIB = None
        return cracks_b200
    cracks_b200.build_library()
    return cracks_b200
