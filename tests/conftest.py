import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    from oracle import orc_py
    orc_py.build()
    return orc_py


@pytest.fixture(scope="session")
def enc():
    """The CUDA library; building it here keeps `pytest -m gpu` self-contained on a fresh box."""
    import __graft_entry__ as ge
    from media_b200 import enc as e
    if not os.path.exists(e.LIB_PATH):
        ge.build()
    e.lib()
    return e
