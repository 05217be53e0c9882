import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def built_lib():
    """Path of the in-tree shared library; built on demand so the CPU suite can check its symbols."""
    from bode_b200 import _ext
    if not os.path.exists(_ext.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _ext.LIB_PATH
