import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Libraries are built in-tree by __graft_entry__.build(); build lazily when missing."""
    pkg = os.path.join(ROOT, "pearray_b200")
    if not (os.path.exists(os.path.join(pkg, "libprb200.so")) and os.path.exists(os.path.join(pkg, "libprb200_host.so"))
            and os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so"))):
        import __graft_entry__ as g
        g.build()


def scene_path(name):
    return os.path.join(ROOT, "scenes", name)
