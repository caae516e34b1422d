import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Build libbiggan_b200.so (and the oracle helpers) once if it is missing."""
    so = os.path.join(ROOT, "biggan_tensorflow_b200", "libbiggan_b200.so")
    if not os.path.exists(so):
        import __graft_entry__
        __graft_entry__.build()
    yield
