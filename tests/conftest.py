import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def gpu_lib():
    """The CUDA library on a box with a GPU; fails (never skips) when the extension is missing."""
    from phylowgs_b200 import build
    import phylowgs_b200 as P
    if not os.path.exists(build.LIB):
        build.build_library()
    assert P.device_count() > 0, "no CUDA device visible to libpwgs_b200.so"
    return P
