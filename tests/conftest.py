import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """Compile (if stale) and load the C-ABI library; no CUDA call is made here."""
    import sqpsolver_jl_b200.capi as capi

    capi.build()
    return capi.lib()


@pytest.fixture()
def engine(built_lib):
    import sqpsolver_jl_b200.capi as capi

    eng = capi.Engine(0)
    yield eng
    eng.close()
