import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Both shared libraries must exist; build them if a fresh checkout lacks them."""
    import glpk_js_b200
    import oracle_lib
    if not os.path.exists(glpk_js_b200.LIB_PATH):
        glpk_js_b200.build()
    if not os.path.exists(oracle_lib.LIB_PATH):
        oracle_lib.build()
    yield
