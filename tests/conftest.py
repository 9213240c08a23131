import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
if str(ROOT / "tests") not in sys.path:
    sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library is built in-tree by __graft_entry__.build(); build it here if a fresh checkout lacks it."""
    from raytracinginoneweekendinrust_b200 import _build
    _build.build_cuda()
    yield


@pytest.fixture(scope="session")
def orc():
    import support
    return support.oracle_lib()


@pytest.fixture(scope="session")
def hostsim():
    import support
    return support.hostsim_lib()
