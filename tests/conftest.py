import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library must exist (CPU tests load it too); build it in-tree if it is missing."""
    from phasespace_b200 import build

    if not os.path.exists(build.LIB):
        build.build()
    return build.LIB
