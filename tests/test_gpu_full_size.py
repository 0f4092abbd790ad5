"""GPU parity at BASELINE.json's full single-GPU sizes (1e8 records), where
the oracle is too slow to run in a test: size-independent properties of the
count tables.

  additivity     counts(stream) == sum of counts(shards) for any cut at query
                 boundaries (util.sum_dict over chunks, workflow.py:1058; the
                 reference's documented split-and-merge, doc/perform.md:70-98)
  conservation   with --unassigned every query adds exactly one unit
                 (classify.py:163-170: the 1/k' shares of a query sum to 1)
  idempotence    the same stream twice == twice the table
  two kernels    the run-per-lane kernel and the window kernel, which share no
                 per-query logic, give identical tables
and the oracle itself on the first 2e6 records of the same stream."""
import numpy as np
import pytest

from tests import cases
from woltka_b200 import synth
from woltka_b200._lib import UNITS, KIND_NONE_ID

pytestmark = pytest.mark.gpu
N = 100_000_000


@pytest.fixture(scope='module')
def stream():
    import torch
    dev = torch.device('cuda', 0)
    q, s, _, nq = synth.gen_hits(N, seed=1002, device=dev)
    return q, s, nq


def _cuts(q, parts):
    """Record offsets that cut the stream into `parts` uneven shards at query
    boundaries."""
    import torch
    n = q.numel()
    cuts = [0]
    for i in range(1, parts):
        x = n * i * i // (parts * parts) + 17 * i
        while x < n and int(q[x]) == int(q[x - 1]):
            x += 1
        cuts.append(max(x, cuts[-1]))
    cuts.append(n)
    return cuts


@pytest.mark.parametrize('entries,mode', [(['genus'], 'default'),
                                          (['genus'], 'major+unassigned'),
                                          (['phylum', 'genus', 'species'],
                                           'above+unassigned')])
def test_classify_1e8_properties(engine, stream, entries, mode):
    q, s, nq = stream
    case = cases.Case(synth.Taxonomy(seed=42))
    kinds, tab, _ = case.tables(entries)
    fl = cases.MODES[mode] | cases.F_UNASSIGNED
    engine.set_tree(case.ft.parent, 0)
    engine.set_plan(kinds, fl, 0.8, 1, case.NF)
    engine.set_subjects(tab, case.sub_node)

    def run(a, b, times=1):
        for _ in range(times):
            engine.classify_device(q.data_ptr() + 4 * a, s.data_ptr() + 4 * a,
                                   b - a)
        return engine.fetch_counts()

    engine.reset_counts()
    whole = run(0, N)
    assert engine.last_kernel() == (
        'classify_seg_kernel' if mode == 'default' else
        'classify_multi_kernel')      # --above, --major 80
    assert not len(engine.fetch_overflow()[0])
    # conservation: one unit per query and entry
    assert np.array_equal(whole.sum(axis=(1, 2)),
                          np.full(len(entries), nq * UNITS, dtype=np.int64))
    # additivity over uneven shards (cuts are 16-byte aligned or not: both)
    cuts = _cuts(q, 7)
    cuts = [c - c % 4 if int(q[c - c % 4]) != int(q[c - c % 4 - 1]) else c
            for c in cuts[:-1]] + [N]
    engine.reset_counts()
    for a, b in zip(cuts[:-1], cuts[1:]):
        if a % 4 == 0:
            shard = run(a, b)
        else:                      # device columns must be 16-byte aligned
            import torch
            qq, ss = q[a:b].clone(), s[a:b].clone()
            torch.cuda.synchronize()   # the engine has its own stream
            engine.classify_device(qq.data_ptr(), ss.data_ptr(), b - a)
            torch.cuda.synchronize()
            shard = engine.fetch_counts()
    assert np.array_equal(shard, whole)
    # idempotence
    assert np.array_equal(run(0, N), 2 * whole)
    # the window kernel on the same stream
    engine.set_tuning(0, 1, 0)
    try:
        engine.reset_counts()
        other = run(0, N)
        assert engine.last_kernel() == 'classify_kernel'
    finally:
        engine.set_tuning(0, 0, 0)
    assert np.array_equal(other, whole)
    # and the oracle on the head of the stream
    m = 2_000_000
    while int(q[m]) == int(q[m - 1]):
        m += 1
    engine.reset_counts()
    head = run(0, m)
    exp = cases.run_oracle(case, entries, fl, 0.8, q[:m].cpu().numpy(),
                           s[:m].cpu().numpy(), n_threads=4)[0]
    assert np.array_equal(head, exp)


def test_ordinal_1e8_properties():
    """cfg3 at full size: 1e8 reads x 5M genes; additivity over two shards,
    idempotence, and conservation of the per-query shares."""
    import torch
    from woltka_b200.engine import Engine
    dev = torch.device('cuda', 0)
    coff, gb, ge = synth.gen_genes()
    G = len(gb)
    cols = synth.gen_reads(N, seed=1003, device=dev)
    rq, rc, rb, re_, rl, nq = cols
    eng = Engine(0)
    try:
        eng.set_plan(np.array([KIND_NONE_ID]), 0, 0.0, 1, G)
        eng.set_subjects(None, None, G)
        eng.ordinal_set_genes(coff, gb, ge, np.arange(G, dtype=np.int32))

        def run(a, b):
            ptrs = [x.data_ptr() + 4 * a for x in (rq, rc, rb, re_, rl)]
            eng.ordinal_device(ptrs, b - a, 0.8)
            return eng.fetch_counts()

        whole = run(0, N)
        assert not len(eng.fetch_overflow()[0])
        # every query with at least one gene adds exactly one unit
        total = int(whole.sum())
        assert total % UNITS == 0 and 0 < total // UNITS <= nq
        cut = N // 2 - (N // 2) % 4
        while int(rq[cut]) == int(rq[cut - 1]) or cut % 4:
            cut += 1
        eng.reset_counts()
        run(0, cut)
        both = run(cut, N)
        assert np.array_equal(both, whole)
        assert np.array_equal(run(0, N), 2 * whole)
    finally:
        eng.close()
