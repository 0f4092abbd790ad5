"""The C-ABI shared library loads and exports every symbol that
include/woltka_b200.h declares (no compute calls: runs without a GPU)."""
import ctypes
import os
import re

from woltka_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, 'include', 'woltka_b200.h')) as f:
        text = f.read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(wk_[a-z_0-9]+)\s*\(', text)))


def test_library_exports_the_header():
    names = declared_symbols()
    assert len(names) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(lib, name), f'{name} is declared but not exported'


def test_python_binding_covers_the_header():
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.wk_abi_version() == 1


def test_no_gpu_is_a_loud_error():
    """Without a device the product fails with the library's message instead
    of falling back to anything."""
    import pytest
    lib = _lib.load()
    n = ctypes.c_int(0)
    rc = lib.wk_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip('a GPU is present')
    from woltka_b200.engine import Engine, WoltkaB200Error
    with pytest.raises(WoltkaB200Error):
        Engine(0)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, 'woltka_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert 'oracle' not in src.replace('no oracle', ''), \
                    f'{fn} mentions the oracle'
