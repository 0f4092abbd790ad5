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


def test_wire_format_packer():
    """Engine.pack_columns (host side of wk_classify_packed[_bits]): one head
    bit per record; subjects as a little-endian bit stream of 8 / 10 / 12 / 14
    / 20 / 24 / 28 bits each, else uint16 / uint32 arrays."""
    import numpy as np
    from woltka_b200.engine import Engine
    rng = np.random.default_rng(0)
    for top, width, stream in ((200, 8, True), (1000, 10, True), (4000, 12, True),
                               (9999, 14, True), (60000, 16, False),
                               (10 ** 6, 20, True), (10 ** 7, 24, True),
                               (2 * 10 ** 8, 28, True), (2 ** 31 - 1, 32, False)):
        for n in (1, 3, 4, 5, 63, 64, 65, 1000):
            k = rng.integers(1, 4, n)
            q = np.repeat(np.arange(n), k)[:n].astype(np.int32)
            s = rng.integers(0, top + 1, n).astype(np.int64)
            s[0] = top
            p = Engine.pack_columns(q, s, pinned=False)
            assert (p.width, p.stream, p.n) == (width, stream, n)
            heads = np.unpackbits(p.bits.view(np.uint8), bitorder='little')[:n]
            assert heads.tolist() == [1] + (q[1:] != q[:-1]).astype(int).tolist()
            if stream:
                bits = np.unpackbits(p.subj.view(np.uint8), bitorder='little')
                got = bits[:n * width].reshape(n, width).astype(np.uint64)
                got = (got << np.arange(width, dtype=np.uint64)).sum(axis=1)
                assert got.tolist() == s.tolist()
                # readable one word past the last subject (the device loads
                # the word behind a subject that ends on a word boundary)
                assert p.subj.nbytes >= (n * width + 7) // 8 + 8
            else:
                assert p.subj[:n].tolist() == s.tolist()
            assert p.nbytes == (n + 63) // 64 * 8 + (n * width + 7) // 8
    # n_subjects fixes the width whatever the chunk holds
    p = Engine.pack_columns(np.zeros(4, np.int32), np.arange(4), pinned=False,
                            n_subjects=10000)
    assert p.width == 14 and p.stream
    p = Engine.pack_columns(np.zeros(0, np.int32), np.zeros(0, np.int32),
                            pinned=False)
    assert p.n == 0
