import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line(
        'markers', 'gpu: needs a B200 (run with -m gpu on the GPU box)')


def _have_device():
    """True when the library loads and sees at least one CUDA device."""
    try:
        import ctypes as C
        from woltka_b200 import _lib
        n = C.c_int(0)
        return _lib.load().wk_device_count(C.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """Tests that need a B200 are skipped (not failed) on a box without a
    device or driver, so a plain `pytest tests` works everywhere."""
    if _have_device():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def engine():
    from woltka_b200.engine import Engine
    if not _have_device():
        pytest.skip('no CUDA device')
    eng = Engine(0)
    yield eng
    eng.close()


@pytest.fixture
def engine_factory():
    """A function that makes a fresh Engine on device 0 (the caller closes it)."""
    from woltka_b200.engine import Engine
    if not _have_device():
        pytest.skip('no CUDA device')
    return lambda: Engine(0)


class _Knobs:
    """wk_set_option on the session's engine, restored when the test ends."""

    def __init__(self, eng):
        self.eng, self.used = eng, set()

    def set(self, name, value):
        self.used.add(name)
        self.eng.set_option(name, value)

    def reset(self):
        for name in self.used:
            self.eng.set_option(name, 0)


@pytest.fixture
def knobs(engine):
    k = _Knobs(engine)
    yield k
    k.reset()
