import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(autouse=True)
def _reset_specdec_options(request):
    """GPU tests flip library options (specdec_set_option); every test starts from and leaves the library defaults,
    so the path bench.py times (default options) is the path the other tests exercise."""
    if request.node.get_closest_marker("gpu") is None:
        yield
        return
    from specdec_b200 import _lib
    lib = _lib.lib()
    assert lib.specdec_set_option(b"reset", 1) == 0
    yield
    assert lib.specdec_set_option(b"reset", 1) == 0
