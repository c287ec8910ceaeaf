import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session")
def native_lib():
    """The in-tree CUDA library; built on demand (nvcc cross-compiles without a GPU)."""
    from eventql_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        from eventql_b200 import build
        build.build()
    return capi.lib()


@pytest.fixture(scope="session")
def gpu_ctx(native_lib):
    from eventql_b200 import capi
    ctx = capi.Context(0)   # raises if there is no device: the CUDA path has no fallback
    yield ctx
    ctx.close()
