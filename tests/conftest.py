import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle and (if missing) the product library once per session."""
    from oracle import core
    core.build_lib()
    lib = os.path.join(ROOT, "dune_copasi_b200", "libdune_copasi_b200.so")
    if not os.path.exists(lib):
        from dune_copasi_b200 import build
        build.build()
    yield
