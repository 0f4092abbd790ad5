"""GPU parity of the subject-coverage store (wk_cover_add / wk_cover_merge /
wk_cover_fetch) against the pure-Python restatement of range.merge_ranges
(oracle/pyport.py; reference vectors: tests/test_range.py of the reference)."""
import numpy as np
import pytest

from oracle import pyport

pytestmark = pytest.mark.gpu


def _expected(sample, subject, beg, end):
    store = {}
    for sm, sb, b, e in zip(sample.tolist(), subject.tolist(), beg.tolist(),
                            end.tolist()):
        store.setdefault((sm, sb), []).extend((b, e))
    rows = []
    for (sm, sb), ranges in sorted(store.items()):
        m = pyport.merge_ranges(ranges)
        rows.extend((sm, sb, m[k], m[k + 1]) for k in range(0, len(m), 2))
    return np.asarray(rows, dtype=np.int32).reshape(-1, 4)


def _got(eng):
    return np.stack(eng.cover_ranges(), axis=1)


def test_reference_vectors(engine_factory):
    # range.merge_ranges docstring and tests/test_range.py:83-104
    for ranges, exp in [([1, 3, 2, 4, 6, 8, 7, 9], [1, 4, 6, 9]),
                        ([1, 2, 3, 4, 5, 6], [1, 2, 3, 4, 5, 6]),
                        ([1, 3, 3, 5, 6, 8], [1, 5, 6, 8]),      # touching
                        ([5, 9, 1, 20, 2, 3], [1, 20]),          # nested
                        ([4, 4, 4, 4], [4, 4]),                  # empty, equal
                        ([], [])]:
        assert pyport.merge_ranges(ranges) == exp
        eng = engine_factory()
        n = len(ranges) // 2
        eng.cover_add(np.zeros(n), np.full(n, 7), ranges[0::2], ranges[1::2])
        got = _got(eng)
        assert got[:, 2:].reshape(-1).tolist() == exp
        assert all(got[:, 0] == 0) and all(got[:, 1] == 7)
        eng.close()


@pytest.mark.parametrize('n,span', [(1, 10), (1000, 50), (200_000, 300),
                                    (200_000, 40_000)])
def test_random_intervals(engine_factory, n, span):
    rng = np.random.default_rng(n + span)
    sample = rng.integers(0, 5, n).astype(np.int32)
    subject = rng.integers(0, 40, n).astype(np.int32)
    beg = rng.integers(0, 2_000_000, n).astype(np.int32)
    end = beg + rng.integers(0, span, n).astype(np.int32)
    eng = engine_factory()
    # in pieces, with a merge in between (the store re-merges merged ranges)
    cut = n // 3
    eng.cover_add(sample[:cut], subject[:cut], beg[:cut], end[:cut])
    eng.cover_ranges()
    eng.cover_add(sample[cut:], subject[cut:], beg[cut:], end[cut:])
    exp = _expected(sample, subject, beg, end)
    got = _got(eng)
    assert got.shape == exp.shape and np.array_equal(got, exp)
    # idempotent
    assert np.array_equal(_got(eng), exp)
    eng.close()


def test_limits_and_bad_input(engine_factory):
    from woltka_b200.engine import WoltkaB200Error
    eng = engine_factory()
    top = (1 << 31) - 1
    eng.cover_add([4095, 0], [(1 << 21) - 1, 0], [top - 5, 0], [top, top])
    got = _got(eng)
    assert got.tolist() == [[0, 0, 0, top], [4095, (1 << 21) - 1, top - 5, top]]
    for bad in ([4096], [0]), ([0], [1 << 21]):
        with pytest.raises(WoltkaB200Error):
            eng.cover_add(bad[0], bad[1], [1], [2])
    with pytest.raises(WoltkaB200Error):
        eng.cover_add([0], [0], [-1], [2])
    # the failed calls added nothing
    assert _got(eng).tolist() == got.tolist()
    eng.close()
