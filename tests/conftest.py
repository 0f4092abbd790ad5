import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line(
        'markers', 'gpu: needs a B200 (run with -m gpu on the GPU box)')


@pytest.fixture(scope='session')
def engine():
    from woltka_b200.engine import Engine
    eng = Engine(0)
    yield eng
    eng.close()


@pytest.fixture
def engine_factory():
    """A function that makes a fresh Engine on device 0 (the caller closes it)."""
    from woltka_b200.engine import Engine
    return lambda: Engine(0)
