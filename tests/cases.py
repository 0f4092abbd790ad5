"""Seeded integer-world cases shared by the GPU parity tests and smoke()."""
import numpy as np

from woltka_b200 import synth
from woltka_b200._lib import (KIND_NONE, KIND_FREE, KIND_RANK, KIND_NONE_ID,
                              F_UNIQ, F_ABOVE, F_MAJOR, F_UNASSIGNED)
from woltka_b200.synth import Case, MODES  # noqa: F401  (the bench uses them too)


def random_hits(case, n_qry, seed, kmax=16, p=0.48, long_every=0,
                long_len=100, window=20):
    """Random records over ALL subjects of the case, with duplicates and,
    optionally, a query of `long_len` hits every `long_every` queries."""
    rng = np.random.default_rng(seed)
    k = np.minimum(rng.geometric(p, n_qry), kmax)
    if long_every:
        k[long_every - 1::long_every] = long_len
    q = np.repeat(np.arange(n_qry, dtype=np.int32), k)
    n = len(q)
    first = rng.integers(0, case.V, n_qry)
    start = np.cumsum(k) - k
    pos = np.arange(n) - start[q]
    s = (first[q] + np.where(pos == 0, 0, rng.integers(0, window, n))) % case.V
    dup = (rng.random(n) < 0.05) & (pos > 0)
    s = np.where(dup, np.roll(s, 1), s)
    return q, s.astype(np.int32)


def run_engine(eng, case, entries, flags, major_th, q, s, n_samples=1,
               q_sample=None, q_stratum=None, sample=0, subok=False,
               root=0, chunks=1):
    kinds, tab, _ = case.tables(entries, subok)
    eng.set_tree(case.ft.parent, root)
    eng.set_plan(kinds, flags, major_th, n_samples, case.NF)
    eng.set_subjects(tab, case.sub_node)
    # split at query boundaries into `chunks` calls
    n = len(q)
    cuts = [0]
    for c in range(1, chunks):
        x = n * c // chunks
        while 0 < x < n and q[x] == q[x - 1]:
            x += 1
        cuts.append(max(x, cuts[-1]))
    cuts.append(n)
    for a, b in zip(cuts[:-1], cuts[1:]):
        eng.classify_chunk(q[a:b], s[a:b], q_sample, q_stratum, sample)
    return collect(eng, n_samples, case.NF)


def collect(eng, n_samples, NF):
    """Engine results in the canonical form oracle.classify returns."""
    units = eng.fetch_counts()
    cell, strat, den = eng.fetch_overflow()
    NF1 = NF + 1
    overflow = []
    for c, t, d in zip(cell.tolist(), strat.tolist(), den.tolist()):
        es, f = divmod(c, NF1)
        overflow.append((es // n_samples, es % n_samples, t, f, d))
    overflow.sort()
    strata = {}
    e, sm, st, f, u = eng.fetch_strata()
    for i in range(len(e)):
        strata[(int(e[i]), int(sm[i]), int(st[i]), int(f[i]))] = int(u[i])
    return units, overflow, strata


def run_oracle(case, entries, flags, major_th, q, s, n_samples=1,
               q_sample=None, q_stratum=None, sample=0, subok=False, root=0,
               n_threads=1):
    from oracle import oracle as O
    kinds, _, trk = case.tables(entries, subok)
    return O.classify(q, s, parent=case.ft.parent,
                      node_rank=case.ft.node_rank, root=root,
                      sub_node=case.sub_node, sub_feat=case.sub_feat,
                      kinds=kinds, target_rank=trk, flags=flags,
                      major_th=major_th, subok=subok, n_samples=n_samples,
                      n_features=case.NF, q_sample=q_sample,
                      q_stratum=q_stratum, sample=sample, n_threads=n_threads)
