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


@pytest.fixture(scope="session")
def host_so(tmp_path_factory):
    """The test-only host implementation of the C ABI (tests/host_emul/host_lib.cpp), compiled once per session."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "host_emul"))
    import host_env
    return host_env.build(str(tmp_path_factory.mktemp("host_lib") / "host_lib.so"))
