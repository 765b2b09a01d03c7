import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def tss():
    import tss_loader
    return tss_loader.load()


@pytest.fixture(scope="session")
def orc():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc as _orc
    _orc.lib()
    return _orc
