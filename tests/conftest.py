import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _experimental_library():
    try:
        from learn_path_tracing_b200 import _lib
        return _lib.has_experimental()
    except Exception:
        return False


# tests of the `make EXPERIMENTAL=1` kernel forms exist only for a library that has them (no skips in the default suite)
collect_ignore = [] if _experimental_library() else ["test_gpu_experimental.py"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import ptoracle
    ptoracle.load()
    return ptoracle


@pytest.fixture(scope="session")
def ctx():
    """The CUDA context; GPU tests fail loudly (no fallback) when the library or device is missing."""
    import learn_path_tracing_b200 as L
    return L.default_context()
