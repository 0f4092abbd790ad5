"""GPU parity: CUDA classify path (through the C-ABI) vs the CPU oracle,
bit-exact on the integer units tables."""
import numpy as np
import pytest

from tests import cases
from woltka_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def small_case():
    tax = synth.Taxonomy(seed=7, level_sizes=[1, 2, 5, 12, 30, 60, 150, 400],
                         n_genomes=900)
    return cases.Case(tax, n_extra=40, internal_subjects=60, seed=3)


@pytest.fixture(scope='module')
def big_case():
    return cases.Case(synth.Taxonomy(seed=42))


def _lean_kernel(entries, mode, th=0.8):
    """The kernel a plan of one kind takes (ranks only, or `none` only, staged
    tables, no strata, no read map): default or --uniq at one entry ->
    classify_seg_kernel; several ranks, --above, --major above one half ->
    classify_multi_kernel (all ranks in one pass); --major up to one half
    (ties between top taxa possible) -> classify_fast_kernel.
    assign_none knows no --major / --above."""
    none = all(x == 'none' for x in entries)
    if not none and any(x in ('none', 'free') for x in entries):
        return 'classify_kernel'
    if not none and (mode.startswith('above') or
                     (mode.startswith('major') and th > 0.5) or (
            len(entries) > 1 and mode.startswith(('default', 'uniq')))):
        return 'classify_multi_kernel'    # all ranks in one pass
    if none or mode.startswith(('default', 'uniq')):
        return 'classify_seg_kernel'
    return 'classify_fast_kernel'


def _same(a, b):
    ua, oa, sa = a
    ub, ob, sb = b
    assert ua.shape == ub.shape
    bad = np.argwhere(ua != ub)
    assert len(bad) == 0, f'{len(bad)} cells differ, first {bad[:5]}'
    assert oa == ob
    assert sa == sb


@pytest.mark.parametrize('mode', list(cases.MODES))
@pytest.mark.parametrize('entries', [['genus'], ['none'], ['free'],
                                     ['phylum', 'genus', 'species'],
                                     ['none', 'free', 'family']])
def test_modes_small(engine, small_case, mode, entries):
    q, s = cases.random_hits(small_case, 20000, seed=11)
    flags = cases.MODES[mode]
    th = 0.8
    _same(cases.run_engine(engine, small_case, entries, flags, th, q, s),
          cases.run_oracle(small_case, entries, flags, th, q, s))


@pytest.mark.parametrize('th', [0.54, 0.55, 0.6, 0.81, 0.5, 0.3, 1.0])
def test_major_thresholds(engine, small_case, th):
    # includes th <= 0.5 (first-seen top item) and the fp-sensitive values
    q, s = cases.random_hits(small_case, 30000, seed=5, kmax=12, p=0.3,
                             window=8)
    fl = cases.MODES['major+unassigned']
    _same(cases.run_engine(engine, small_case, ['genus', 'family'], fl, th, q, s),
          cases.run_oracle(small_case, ['genus', 'family'], fl, th, q, s))


@pytest.mark.parametrize('mode', ['default', 'major', 'above', 'uniq'])
def test_long_queries(engine, small_case, mode):
    # queries of 100 and 700 hits take the slow path, k > 16 exercises the
    # overflow list (denominators that do not divide 720720)
    fl = cases.MODES[mode]
    for ll in (33, 100, 700):
        q, s = cases.random_hits(small_case, 3000, seed=ll, kmax=31,
                                 p=0.15, long_every=97, long_len=ll,
                                 window=300)
        ent = ['none', 'free', 'genus', 'phylum']
        _same(cases.run_engine(engine, small_case, ent, fl, 0.6, q, s),
              cases.run_oracle(small_case, ent, fl, 0.6, q, s))


def test_subok_and_no_root(engine, small_case):
    q, s = cases.random_hits(small_case, 20000, seed=2)
    for subok in (False, True):
        for root in (0, -1):
            _same(cases.run_engine(engine, small_case, ['free'], 0, 0, q, s,
                                   subok=subok, root=root),
                  cases.run_oracle(small_case, ['free'], 0, 0, q, s,
                                   subok=subok, root=root))


def test_samples_and_chunks(engine, small_case):
    q, s = cases.random_hits(small_case, 50000, seed=9)
    rng = np.random.default_rng(1)
    nq = int(q.max()) + 1
    q_sample = rng.integers(-1, 7, nq).astype(np.int32)  # -1 = dropped
    ent = ['genus', 'none']
    ref = cases.run_oracle(small_case, ent, 0, 0, q, s, n_samples=7,
                           q_sample=q_sample)
    for chunks in (1, 3, 17):
        _same(cases.run_engine(engine, small_case, ent, 0, 0, q, s,
                               n_samples=7, q_sample=q_sample, chunks=chunks),
              ref)
    # scalar sample
    _same(cases.run_engine(engine, small_case, ent, 0, 0, q, s, n_samples=7,
                           sample=5),
          cases.run_oracle(small_case, ent, 0, 0, q, s, n_samples=7, sample=5))


def test_strata(engine, small_case):
    q, s = cases.random_hits(small_case, 40000, seed=21)
    rng = np.random.default_rng(4)
    nq = int(q.max()) + 1
    q_sample = rng.integers(0, 3, nq).astype(np.int32)
    q_stratum = rng.integers(-1, 50, nq).astype(np.int32)
    for mode in ('default', 'uniq+unassigned'):
        fl = cases.MODES[mode]
        _same(cases.run_engine(engine, small_case, ['genus', 'none'], fl, 0,
                               q, s, n_samples=3, q_sample=q_sample,
                               q_stratum=q_stratum),
              cases.run_oracle(small_case, ['genus', 'none'], fl, 0, q, s,
                               n_samples=3, q_sample=q_sample,
                               q_stratum=q_stratum))


def test_edge_inputs(engine, small_case):
    ent = ['genus']
    # empty chunk
    e = np.zeros(0, dtype=np.int32)
    _same(cases.run_engine(engine, small_case, ent, 0, 0, e, e),
          cases.run_oracle(small_case, ent, 0, 0, e, e))
    # single record, and sizes around the tile / window boundaries
    for n in (1, 2, 31, 32, 33, 63, 64, 65, 4095, 4096, 4097, 8191, 8193):
        q, s = cases.random_hits(small_case, n, seed=n)
        _same(cases.run_engine(engine, small_case, ent, 0, 0, q, s),
              cases.run_oracle(small_case, ent, 0, 0, q, s))
    # one query spanning several tiles; all records one subject
    q = np.zeros(10000, dtype=np.int32)
    s = np.full(10000, 5, dtype=np.int32)
    _same(cases.run_engine(engine, small_case, ent, 0, 0, q, s),
          cases.run_oracle(small_case, ent, 0, 0, q, s))
    # every record its own query
    q = np.arange(9999, dtype=np.int32)
    s = (q * 7 % small_case.V).astype(np.int32)
    _same(cases.run_engine(engine, small_case, ent, 0, 0, q, s),
          cases.run_oracle(small_case, ent, 0, 0, q, s))


def test_bad_subject_is_an_error(engine, small_case):
    from woltka_b200.engine import WoltkaB200Error
    q = np.array([0, 1], dtype=np.int32)
    s = np.array([0, small_case.V + 5], dtype=np.int32)
    with pytest.raises(WoltkaB200Error):
        cases.run_engine(engine, small_case, ['genus'], 0, 0, q, s)


@pytest.mark.parametrize('mode', ['default', 'major', 'uniq', 'above'])
def test_cfg2_shape_1e6(engine, big_case, mode):
    # the §8(d) generator at 1e6 records on the 21,603-node taxonomy
    qi, si, _, nq = synth.gen_hits(1_000_000, seed=1002)
    q, s = qi.numpy(), si.numpy()
    fl = cases.MODES[mode]
    ent = ['phylum', 'genus', 'species']
    _same(cases.run_engine(engine, big_case, ent, fl, 0.8, q, s),
          cases.run_oracle(big_case, ent, fl, 0.8, q, s, n_threads=4))


@pytest.mark.parametrize('cache_slots', [0, 4096, 256, -1])
def test_every_count_sink(engine, small_case, cache_slots):
    # 0 = automatic (direct table here), >0 = hashed write-back cache,
    # -1 = straight to global memory
    q, s = cases.random_hits(small_case, 60000, seed=31)
    ent = ['phylum', 'genus', 'none', 'free']
    engine.set_tuning(0, 0, cache_slots)
    try:
        for mode in ('default', 'major+unassigned', 'above'):
            fl = cases.MODES[mode]
            _same(cases.run_engine(engine, small_case, ent, fl, 0.7, q, s,
                                   n_samples=2, sample=1),
                  cases.run_oracle(small_case, ent, fl, 0.7, q, s,
                                   n_samples=2, sample=1))
    finally:
        engine.set_tuning(0, 0, 0)


def test_carry_out_of_the_32bit_low_words(engine, small_case):
    # > 5958 unit counts into one cell overflow a 32-bit low word
    n = 200_000
    q = np.arange(n, dtype=np.int32)
    s = np.full(n, 3, dtype=np.int32)
    for cache_slots in (0, 1024):
        engine.set_tuning(0, 0, cache_slots)
        try:
            _same(cases.run_engine(engine, small_case, ['genus'], 0, 0, q, s),
                  cases.run_oracle(small_case, ['genus'], 0, 0, q, s))
        finally:
            engine.set_tuning(0, 0, 0)


def test_lca_on_a_tree_that_is_not_level_ordered(engine):
    """parent[i] < i holds for a depth-first numbering too; the kernel then
    falls back from the level-synchronous LCA to the pairwise climb."""
    from oracle import oracle as O
    from woltka_b200._lib import KIND_FREE, KIND_RANK, F_ABOVE
    rng = np.random.default_rng(17)
    T = 700
    # random tree, then renumber in depth-first preorder
    par0 = np.zeros(T, dtype=np.int64)
    for i in range(1, T):
        par0[i] = rng.integers(max(0, i - 40), i)
    children = [[] for _ in range(T)]
    for i in range(1, T):
        children[par0[i]].append(i)
    order, stack = [], [0]
    while stack:
        v = stack.pop()
        order.append(v)
        stack.extend(reversed(children[v]))
    new = np.empty(T, dtype=np.int64)
    new[order] = np.arange(T)
    parent = np.zeros(T, dtype=np.int32)
    for old in range(T):
        parent[new[old]] = new[par0[old]]
    depth = np.zeros(T, dtype=np.int64)
    for i in range(1, T):
        depth[i] = depth[parent[i]] + 1
    assert np.any(np.diff(depth) < 0)          # really not level ordered
    node_rank = (depth == 3).astype(np.int32) - 1   # rank 0 at depth 3
    sub_node = np.arange(T, dtype=np.int32)
    anc = np.full(T, -1, dtype=np.int32)
    for i in range(T):
        v = i
        while True:
            if node_rank[v] == 0:
                anc[i] = v
                break
            if parent[v] == v:
                break
            v = parent[v]
    nq = 20000
    k = np.minimum(rng.geometric(0.4, nq), 12)
    q = np.repeat(np.arange(nq, dtype=np.int32), k)
    first = rng.integers(0, T, nq)
    s = ((first[q] + rng.integers(0, 30, len(q))) % T).astype(np.int32)
    par_tab = parent[sub_node]
    kinds = np.array([KIND_FREE, KIND_RANK], dtype=np.int32)
    tab = np.stack([par_tab, anc]).astype(np.int32)
    engine.set_tree(parent, 0)
    engine.set_plan(kinds, F_ABOVE, 0.0, 1, T)
    engine.set_subjects(tab, sub_node)
    engine.classify_chunk(q, s)
    got = engine.fetch_counts()
    exp, ovf, _ = O.classify(q, s, parent=parent, node_rank=node_rank, root=0,
                             sub_node=sub_node, sub_feat=sub_node, kinds=kinds,
                             target_rank=[0, 0], flags=F_ABOVE, n_features=T)
    assert np.array_equal(got, exp)


@pytest.mark.parametrize('noseg', ['', '1'])
@pytest.mark.parametrize('mode', ['default', 'above', 'major+unassigned',
                                  'uniq+unassigned'])
@pytest.mark.parametrize('entries', [['genus'], ['phylum', 'genus', 'species'],
                                     ['none']])
def test_contiguous_samples_take_the_run_per_lane_kernel(engine, small_case,
                                                         entries, mode,
                                                         knobs, noseg):
    """Samples that follow one another in the stream (one file per sample):
    the lane-per-record and the run-per-lane kernel work segment by segment;
    a dropped sample (-1) in the middle and long queries are part of the
    stream."""
    if noseg:
        knobs.set('no_seg', 1)
        knobs.set('no_multi', 1)
    q, s = cases.random_hits(small_case, 60000, seed=77, long_every=5000,
                             long_len=90)
    nq = int(q.max()) + 1
    rng = np.random.default_rng(8)
    q_sample = np.sort(rng.integers(0, 6, nq)).astype(np.int32)
    q_sample[q_sample == 3] = -1          # a dropped sample in the middle
    fl = cases.MODES[mode]
    ref = cases.run_oracle(small_case, entries, fl, 0.7, q, s, n_samples=6,
                           q_sample=q_sample)
    for chunks in (1, 4):
        _same(cases.run_engine(engine, small_case, entries, fl, 0.7, q, s,
                               n_samples=6, q_sample=q_sample, chunks=chunks),
              ref)
    assert engine.last_kernel() == (
        'classify_fast_kernel' if noseg == '1' else
        _lean_kernel(entries, mode, 0.7))


def test_which_kernel_runs(engine, small_case):
    """One entry in default / --uniq mode takes classify_seg_kernel; several
    ranks, --above and --major above one half classify_multi_kernel; --major up
    to one half classify_fast_kernel; strata, mixed kinds and read maps take
    classify_kernel."""
    q, s = cases.random_hits(small_case, 5000, seed=1)
    cases.run_engine(engine, small_case, ['genus'], 0, 0, q, s)
    assert engine.last_kernel() == 'classify_seg_kernel'
    cases.run_engine(engine, small_case, ['genus'], cases.MODES['major'], 0.8, q, s)
    assert engine.last_kernel() == 'classify_multi_kernel'
    cases.run_engine(engine, small_case, ['genus'], cases.MODES['major'], 0.5, q, s)
    assert engine.last_kernel() == 'classify_fast_kernel'
    cases.run_engine(engine, small_case, ['phylum', 'genus'], 0, 0, q, s)
    assert engine.last_kernel() == 'classify_multi_kernel'
    cases.run_engine(engine, small_case, ['genus'], cases.MODES['above'], 0, q, s)
    assert engine.last_kernel() == 'classify_multi_kernel'
    cases.run_engine(engine, small_case, ['phylum', 'genus'],
                     cases.MODES['major'], 0.8, q, s)
    assert engine.last_kernel() == 'classify_multi_kernel'
    cases.run_engine(engine, small_case, ['none', 'free'], 0, 0, q, s)
    assert engine.last_kernel() == 'classify_kernel'
    nq = int(q.max()) + 1
    strat = np.zeros(nq, dtype=np.int32)
    cases.run_engine(engine, small_case, ['genus'], 0, 0, q, s,
                     q_stratum=strat, q_sample=np.zeros(nq, dtype=np.int32))
    assert engine.last_kernel() == 'classify_strata_kernel'
    cases.run_engine(engine, small_case, ['genus'], cases.MODES['major'], 0.8,
                     q, s, q_stratum=strat, q_sample=np.zeros(nq, dtype=np.int32))
    assert engine.last_kernel() == 'classify_kernel'
    engine.set_tuning(0, 1, 0)
    try:
        a = cases.run_engine(engine, small_case, ['genus'], 0, 0, q, s)
        assert engine.last_kernel() == 'classify_kernel'
    finally:
        engine.set_tuning(0, 0, 0)
    _same(a, cases.run_engine(engine, small_case, ['genus'], 0, 0, q, s))


@pytest.mark.parametrize('noseg', ['', '1'])
def test_run_per_lane_kernel_at_1e6(engine, big_case, knobs, noseg):
    # one-entry plans on the 21,603-node taxonomy, every mode, both kinds;
    # with and without the lane-per-record kernel of the default / --uniq mode
    if noseg:
        knobs.set('no_seg', 1)
        knobs.set('no_multi', 1)
    qi, si, _, nq = synth.gen_hits(1_000_000, seed=1002)
    q, s = qi.numpy(), si.numpy()
    for ent in (['genus'], ['species'], ['none']):
        for mode in ('default', 'uniq', 'major', 'above', 'uniq+unassigned'):
            fl = cases.MODES[mode]
            _same(cases.run_engine(engine, big_case, ent, fl, 0.8, q, s),
                  cases.run_oracle(big_case, ent, fl, 0.8, q, s, n_threads=4))
            assert engine.last_kernel() == (
                'classify_fast_kernel' if noseg else _lean_kernel(ent, mode))


def test_cfg5_shape_gene_to_ko_stratified_by_genus(engine):
    """BASELINE.json configs[4] at test size: subjects are genes, the rank is
    'ko' through a gene -> KO map read as a two-level tree (tree.read_map /
    find_rank), counts are keyed by (genus stratum, KO) and split over samples
    (classify.counter_strat, classify.py:216-249).  The gene table is too large
    to stage as uint16 rows at full size; here V = 60,000 > 65,535 - so the
    int32 table path of classify_kernel runs."""
    from oracle import oracle as O
    from woltka_b200._lib import KIND_RANK
    rng = np.random.default_rng(1005)
    n_genomes, genes_per, n_ko, n_genus, n_samples = 1200, 50, 900, 300, 8
    V = n_genomes * genes_per
    # nodes: 0 = root, 1..n_ko = KOs (rank 0), then one node per gene
    T = 1 + n_ko + V
    parent = np.zeros(T, dtype=np.int32)
    node_rank = np.full(T, -1, dtype=np.int32)
    node_rank[1:1 + n_ko] = 0
    ko_of_gene = np.where(rng.random(V) < 0.6, rng.integers(1, n_ko + 1, V), 0)
    parent[1 + n_ko:] = ko_of_gene          # unannotated genes hang off the root
    sub_node = (1 + n_ko + np.arange(V)).astype(np.int32)
    tab = np.where(ko_of_gene > 0, ko_of_gene, -1).astype(np.int32)[None, :]
    # records: k hits per query on genes of neighbouring genomes
    nq = 150_000
    k = np.minimum(rng.geometric(0.48, nq), 16)
    q = np.repeat(np.arange(nq, dtype=np.int32), k)
    first = rng.integers(0, n_genomes, nq)
    genome = (first[q] + rng.integers(0, 3, len(q))) % n_genomes
    s = (genome * genes_per + rng.integers(0, genes_per, len(q))).astype(np.int32)
    genus_of_genome = rng.integers(0, n_genus, n_genomes)
    q_stratum = np.where(rng.random(nq) < 0.8, genus_of_genome[first], -1).astype(np.int32)
    q_sample = (np.arange(nq) * n_samples // nq).astype(np.int32)
    kinds = np.array([KIND_RANK], dtype=np.int32)
    engine.set_tree(parent, 0)
    engine.set_plan(kinds, 0, 0.0, n_samples, T)
    engine.set_subjects(tab, sub_node)
    engine.classify_chunk(q, s, q_sample, q_stratum, 0)
    got = cases.collect(engine, n_samples, T)
    exp = O.classify(q, s, parent=parent, node_rank=node_rank, root=0,
                     sub_node=sub_node, sub_feat=sub_node, kinds=kinds,
                     target_rank=[0], flags=0, n_samples=n_samples,
                     n_features=T, q_sample=q_sample, q_stratum=q_stratum,
                     n_threads=4)
    _same(got, exp)
    assert len(got[2]) > 10000          # (sample, genus, KO) cells
    assert engine.last_kernel() == 'classify_strata_kernel'


@pytest.mark.parametrize('r', ['5', '9'])
@pytest.mark.parametrize('block', [0, 768, 512])
def test_short_runs_and_fewer_warps(engine, small_case, knobs, r, block):
    """The run-per-lane kernel at every run length it is built for (13 is the
    default) and with fewer warps per CTA, as chosen for large tables."""
    knobs.set('sweep_r', int(r))
    knobs.set('no_seg', 1)
    knobs.set('no_multi', 1)
    q, s = cases.random_hits(small_case, 30000, seed=int(r), long_every=3000,
                             long_len=60)
    engine.set_tuning(0, block, 0)
    try:
        for ent, mode in ((['genus'], 'default'), (['none'], 'uniq'),
                          (['phylum', 'genus', 'species'], 'above'),
                          (['family', 'genus'], 'major+unassigned')):
            fl = cases.MODES[mode]
            _same(cases.run_engine(engine, small_case, ent, fl, 0.8, q, s),
                  cases.run_oracle(small_case, ent, fl, 0.8, q, s))
            assert engine.last_kernel() == 'classify_fast_kernel'
    finally:
        engine.set_tuning(0, 0, 0)


@pytest.mark.parametrize('noseg', ['', '1'])
def test_randomised_shapes_both_kernels_agree(engine, small_case, knobs,
                                              noseg):
    """Many small random streams whose sizes straddle the warp-tile and run
    boundaries (32 x 13 = 416 records, runs of 13; tiles of 512), with long
    queries and repeats placed at random: lane-per-record kernel ==
    run-per-lane kernel == window kernel == oracle."""
    if noseg:
        knobs.set('no_seg', 1)
        knobs.set('no_multi', 1)
    rng = np.random.default_rng(2026)
    ents = (['genus'], ['none'], ['phylum', 'species'])
    modes = ('default', 'uniq', 'major', 'above', 'major+unassigned')
    for it in range(40):
        nq = int(rng.choice([1, 3, 40, 200, 205, 247, 416, 420, 832, 1000, 3000]))
        p = float(rng.choice([0.2, 0.48, 0.9]))
        every = int(rng.choice([0, 7, 50]))
        q, s = cases.random_hits(small_case, nq, seed=1000 + it, kmax=45, p=p,
                                 long_every=every,
                                 long_len=int(rng.choice([38, 40, 41, 43, 90])),
                                 window=int(rng.choice([2, 20, 400])))
        ent = ents[it % len(ents)]
        mode = modes[it % len(modes)]
        fl = cases.MODES[mode]
        th = float(rng.choice([0.5, 0.51, 0.8]))
        exp = cases.run_oracle(small_case, ent, fl, th, q, s)
        got = cases.run_engine(engine, small_case, ent, fl, th, q, s)
        assert engine.last_kernel() == (
            'classify_fast_kernel' if noseg == '1' else
            _lean_kernel(ent, mode, th))
        _same(got, exp)
        engine.set_tuning(0, 1, 0)
        try:
            _same(cases.run_engine(engine, small_case, ent, fl, th, q, s), exp)
        finally:
            engine.set_tuning(0, 0, 0)


@pytest.mark.parametrize('NF', [5000, 3_000_000])
@pytest.mark.parametrize('mode', ['default', 'uniq+unassigned'])
@pytest.mark.parametrize('noseg', ['', '1'])
def test_rank_none_without_a_table(engine, mode, NF, knobs, noseg):
    """feature == subject (KIND_NONE_ID: an OGU table, or genes of the ordinal
    path): the lane-per-record kernel, or the run-per-lane kernel with 24-bit
    codes; counts in the private table when it fits (NF = 5000) and straight
    to global memory when it does not (NF = 3e6)."""
    if noseg:
        knobs.set('no_seg', 1)
        knobs.set('no_multi', 1)
    from oracle import oracle as O
    from woltka_b200._lib import KIND_NONE_ID
    rng = np.random.default_rng(NF)
    nq = 40000
    k = np.minimum(rng.geometric(0.4, nq), 20)
    k[::997] = 75                                   # long queries
    q = np.repeat(np.arange(nq, dtype=np.int32), k)
    first = rng.integers(0, NF, nq)
    s = ((first[q] + rng.integers(0, 6, len(q)) * 65537) % NF).astype(np.int32)
    kinds = np.array([KIND_NONE_ID], dtype=np.int32)
    fl = cases.MODES[mode]
    engine.set_plan(kinds, fl, 0.0, 2, NF)
    engine.set_subjects(None, None, NF)
    engine.classify_chunk(q, s, None, None, 1)
    assert engine.last_kernel() == ('classify_fast_kernel' if noseg else
                                    'classify_seg_kernel')
    got = cases.collect(engine, 2, NF)
    exp = O.classify(q, s, kinds=kinds, flags=fl, n_samples=2, n_features=NF,
                     sample=1, n_threads=4)
    _same(got, exp)


@pytest.mark.parametrize('sub', ['36', '516', '4000'])
@pytest.mark.parametrize('noseg', ['', '1'])
def test_host_chunk_in_sub_chunks(engine, small_case, knobs, sub, noseg):
    """wk_classify_chunk launches one kernel per sub-chunk of the host columns
    (records [r0, r1), not cut at query boundaries: a query belongs to the
    launch that holds its first record).  Small sub-chunks put those cuts, the
    512-record warp tiles and the 32-record windows in every relative
    position, with queries longer than a window and longer than a sub-chunk."""
    knobs.set('cls_sub', int(sub))
    if noseg:
        knobs.set('no_seg', 1)
        knobs.set('no_multi', 1)
    q, s = cases.random_hits(small_case, 2500, seed=int(sub), kmax=31, p=0.3,
                             long_every=211, long_len=70, window=6)
    for ent, mode in ((['genus'], 'default'), (['genus'], 'uniq+unassigned'),
                      (['none'], 'default'), (['species'], 'major')):
        fl = cases.MODES[mode]
        _same(cases.run_engine(engine, small_case, ent, fl, 0.8, q, s),
              cases.run_oracle(small_case, ent, fl, 0.8, q, s))
        assert engine.last_kernel() == (
            'classify_fast_kernel' if noseg else _lean_kernel(ent, mode))


@pytest.mark.parametrize('sub', [0, 64, 192, 4096])
def test_packed_wire_format(engine, small_case, knobs, sub):
    """wk_classify_packed[_bits]: head bits + subjects (bit stream of 8..28
    bits each, or uint16 / uint32) expanded on the device must give what the int32 SoA columns give — with sub-chunks
    that cut queries (also queries longer than two sub-chunks), a per-query
    sample column, and an out-of-range subject reported as an error."""
    from woltka_b200.engine import Engine
    if sub:
        knobs.set('cls_sub', sub)
    q, s = cases.random_hits(small_case, 3000, seed=sub + 1, kmax=31, p=0.3,
                             long_every=211, long_len=300, window=6)
    # query ids with gaps and in no order: only q[i] != q[i+1] matters
    ids = np.random.default_rng(2).permutation(10 * (int(q.max()) + 1))
    nq = int(q.max()) + 1
    q_sample = (np.arange(nq) * 3 // nq).astype(np.int32)
    for ent, mode in ((['genus'], 'default'), (['none'], 'uniq+unassigned'),
                      (['phylum', 'genus', 'species'], 'above'),
                      (['species'], 'major')):
        fl = cases.MODES[mode]
        ref = cases.run_engine(engine, small_case, ent, fl, 0.8, q, s,
                               n_samples=3, q_sample=q_sample)
        # subjects as a bit stream of every width the packer writes, and as
        # uint16 / uint32 arrays
        for n_subjects in (None, 1 << 10, 1 << 12, 1 << 14, 1 << 16, 1 << 20,
                           1 << 24, 1 << 28, 1 << 31):
            packed = Engine.pack_columns(ids[q].astype(np.int32), s,
                                         n_subjects=n_subjects)
            assert packed.stream == (n_subjects not in (1 << 16, 1 << 31))
            engine.reset_counts()
            engine.classify_packed(packed, q_sample, 0)
            _same(cases.collect(engine, 3, small_case.NF), ref)
    engine.reset_counts()
    bad = Engine.pack_columns(q, np.where(np.arange(len(s)) == 77,
                                          small_case.V + 5, s))
    with pytest.raises(Exception, match='subject index'):
        engine.classify_packed(bad, q_sample, 0)
    engine.classify_packed(Engine.pack_columns(q[:0], s[:0]), None, 0)  # empty


@pytest.mark.parametrize('gtab', [0, 1, 2])
@pytest.mark.parametrize('entries', [['genus'], ['none'], ['species', 'genus']])
def test_stratified_one_kind_plans(engine, small_case, knobs, entries, gtab):
    """classify_strata_kernel (counts keyed by the query's stratum,
    classify.counter_strat classify.py:216-249): interleaved samples, queries
    without a stratum, dropped samples, 17 distinct hits (shares of 1/17 go to
    the overflow list), queries longer than a window, several chunks; table
    staged in shared memory or read through L2; updates straight into the
    strata table or (2) staged per table region and applied region by region
    (what a table much larger than L2 gets)."""
    knobs.set('strata_gtab', int(gtab > 0))
    knobs.set('strata_part', int(gtab == 2))
    q, s = cases.random_hits(small_case, 30000, seed=31 + gtab, long_every=997,
                             long_len=17)
    k = np.bincount(q)
    pos = np.arange(len(q)) - (np.cumsum(k) - k)[q]
    lng = k[q] == 17
    s[lng] = ((q[lng] * 7 + pos[lng]) % small_case.V).astype(np.int32)
    q2, s2 = cases.random_hits(small_case, 2000, seed=5, long_every=301, long_len=70)
    q = np.concatenate([q, q2 + int(q.max()) + 1])
    s = np.concatenate([s, s2])
    nq = int(q.max()) + 1
    rng = np.random.default_rng(9)
    q_sample = rng.integers(-1, 4, nq).astype(np.int32)
    q_stratum = rng.integers(-1, 40, nq).astype(np.int32)
    for mode in ('default', 'uniq', 'uniq+unassigned'):
        fl = cases.MODES[mode]
        ref = cases.run_oracle(small_case, entries, fl, 0, q, s, n_samples=4,
                               q_sample=q_sample, q_stratum=q_stratum)
        assert ref[2] and (mode != 'default' or ref[1])
        for chunks in (1, 5):
            _same(cases.run_engine(engine, small_case, entries, fl, 0, q, s,
                                   n_samples=4, q_sample=q_sample,
                                   q_stratum=q_stratum, chunks=chunks), ref)
            assert engine.last_kernel() == 'classify_strata_kernel'
