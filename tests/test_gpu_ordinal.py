"""GPU parity: closed-form read-gene matching kernel vs the reference's
endpoint sweep (oracle restatement of ordinal.match_read_gene)."""
import numpy as np
import pytest

from oracle import oracle as O
from woltka_b200 import synth
from woltka_b200.engine import Engine

pytestmark = pytest.mark.gpu


def _gpu_pairs(coff, gb, ge, cols, th):
    eng = Engine(0)
    try:
        eng.ordinal_set_genes(coff, gb, ge, np.arange(len(gb), dtype=np.int32))
        eng.ordinal_enable_pairs()
        eng.ordinal_chunk(*cols, th)
        return eng.ordinal_pairs()
    finally:
        eng.close()


def _check(coff, gb, ge, cols, th):
    r, g = _gpu_pairs(coff, gb, ge, cols, th)
    q, c, b, e, l = cols
    er, eg = O.ordinal_match(c, b, e, l, th, coff, gb, ge)
    assert len(r) == len(er)
    assert np.array_equal(r, er) and np.array_equal(g, eg)
    return len(r)


def _kat(genes, reads, lens, th):
    order = np.argsort([g[0] for g in genes], kind='stable')
    gb = np.array([genes[i][0] for i in order], dtype=np.int32)
    ge = np.array([genes[i][1] for i in order], dtype=np.int32)
    coff = np.array([0, len(gb)], dtype=np.int64)
    n = len(reads)
    cols = (np.arange(n, dtype=np.int32), np.zeros(n, dtype=np.int32),
            np.array([r[0] for r in reads], dtype=np.int32),
            np.array([r[1] for r in reads], dtype=np.int32),
            np.full(n, lens, dtype=np.int32))
    r, g = _gpu_pairs(coff, gb, ge, cols, th)
    return sorted(zip(r.tolist(), order[g].tolist()))


def test_reference_kats():
    # known answers of the reference's own unit tests, coordinates as they
    # appear in the encoded queues (gene start already lo-1)
    # tests/test_ordinal.py:121-161 (match_read_gene), L = 20 * 0.8 = 16
    genes = [(5, 29), (33, 61), (65, 94)]
    reads = [(10, 29), (16, 35), (20, 39), (22, 41), (30, 49), (30, 49),
             (60, 79), (65, 84), (82, 95)]
    assert _kat(genes, reads, 20, 0.8) == [(0, 0), (4, 1), (5, 1), (7, 2)]
    # tests/test_ordinal.py:163-197 (match_read_gene_naive), L = 14
    genes = [(4, 29), (32, 61), (64, 94)]
    reads = [(9, 29), (15, 35), (19, 39), (21, 41), (29, 49), (29, 49),
             (59, 79), (64, 84), (81, 95)]
    assert _kat(genes, reads, 14, 1.0) == [(0, 0), (1, 0), (4, 1), (5, 1),
                                           (6, 2), (7, 2)]
    # tests/test_ordinal.py:199-237 (match_read_gene_quart), nested genes
    genes = [(4, 29), (32, 61), (64, 94), (60, 76), (66, 72)]
    reads = [(9, 29), (15, 35), (19, 39), (21, 41), (29, 49), (29, 49),
             (59, 79), (64, 84), (69, 75), (81, 95)]
    assert _kat(genes, reads, 14, 1.0) == sorted(
        [(0, 0), (1, 0), (4, 1), (5, 1), (6, 3), (6, 2), (7, 2)])
    # tests/test_ordinal.py:239-251 "giant read": no match
    assert _kat([(0, 5), (5, 7), (6, 8)], [(3, 9)], 5, 1.0) == []


@pytest.mark.parametrize('th', [0.8, 0.55, 0.81, 0.07, 1.0, 0.5])
def test_random_vs_sweep(th):
    coff, gb, ge = synth.gen_genes(30, 400, 500_000, seed=11)
    rq, rc, rb, re_, rl, _ = synth.gen_reads(200_000, 30, 500_000, seed=12)
    cols = tuple(x.numpy() for x in (rq, rc, rb, re_, rl))
    assert _check(coff, gb, ge, cols, th) > 0


def test_nested_and_giant_genes():
    rng = np.random.default_rng(3)
    # a giant gene covering the contig, nested small genes, ties on starts
    nb = 3000
    b = np.sort(rng.integers(0, 100_000, nb)).astype(np.int32)
    e = (b + rng.integers(1, 400, nb)).astype(np.int32)
    gb = np.concatenate([[0], b, [50], [50]]).astype(np.int32)
    ge = np.concatenate([[120_000], e, [60], [70000]]).astype(np.int32)
    order = np.argsort(gb, kind='stable')
    gb, ge = gb[order], ge[order]
    coff = np.array([0, len(gb)], dtype=np.int64)
    n = 50_000
    rb = rng.integers(0, 110_000, n).astype(np.int32)
    ln = rng.integers(1, 300, n).astype(np.int32)
    cols = (np.arange(n, dtype=np.int32), np.zeros(n, dtype=np.int32), rb,
            rb + ln, ln)
    for th in (0.8, 0.2):
        _check(coff, gb, ge, cols, th)


def test_edges():
    coff, gb, ge = synth.gen_genes(4, 50, 60_000, seed=2)
    # contig without genes (index 4), unknown contig (-1), zero length
    coff = np.concatenate([coff, [coff[-1]]])
    rng = np.random.default_rng(9)
    n = 5000
    c = rng.integers(-1, 5, n).astype(np.int32)
    b = rng.integers(0, 60_000, n).astype(np.int32)
    ln = rng.integers(0, 200, n).astype(np.int32)
    cols = (np.arange(n, dtype=np.int32), c, b, b + ln, ln)
    _check(coff, gb, ge, cols, 0.8)
    # sizes around the tile boundary, and a single read
    for m in (1, 3, 2047, 2048, 2049, 4097):
        cc = tuple(x[:m] for x in cols)
        _check(coff, gb, ge, cc, 0.8)


def test_ordinal_then_classify(engine, ord_sub=0):
    """match + classify fused on the device == oracle sweep + oracle classify
    over the (query, gene) pairs."""
    coff, gb, ge = synth.gen_genes(20, 300, 300_000, seed=21)
    G = len(gb)
    rq, rc, rb, re_, rl, nq = synth.gen_reads(100_000, 20, 300_000, seed=22)
    cols = tuple(x.numpy() for x in (rq, rc, rb, re_, rl))
    from woltka_b200._lib import KIND_NONE_ID
    eng = Engine(0)
    try:
        eng.set_option('ord_sub', ord_sub)
        eng.set_plan(np.array([KIND_NONE_ID]), 0, 0.0, 2, G)
        eng.set_subjects(None, None, G)
        eng.ordinal_set_genes(coff, gb, ge, np.arange(G, dtype=np.int32))
        q_sample = (np.arange(nq) % 2).astype(np.int32)
        eng.ordinal_chunk(*cols, 0.8, q_sample=q_sample)
        units = eng.fetch_counts()
    finally:
        eng.close()
    er, eg = O.ordinal_match(cols[1], cols[2], cols[3], cols[4], 0.8, coff,
                             gb, ge)
    pq = cols[0][er]
    exp, ovf, _ = O.classify(pq, eg, kinds=[KIND_NONE_ID], n_samples=2,
                             n_features=G, q_sample=q_sample)
    assert not ovf
    assert np.array_equal(units, exp)


def test_queries_straddling_tiles(engine, ord_sub=0):
    """A CTA owns the queries whose first record lies in its tile and follows
    the last one past the tile end — including queries longer than a tile."""
    coff, gb, ge = synth.gen_genes(6, 400, 200_000, seed=31)
    G = len(gb)
    rng = np.random.default_rng(5)
    from woltka_b200._lib import KIND_NONE_ID
    for sizes in ([2047, 3, 2046, 5000, 1, 1, 7], [1] * 10 + [6000] + [2] * 50,
                  [4096, 4096], [2048, 2048, 100]):
        q = np.repeat(np.arange(len(sizes), dtype=np.int32), sizes)
        n = len(q)
        c = rng.integers(0, 6, n).astype(np.int32)
        b = rng.integers(0, 199_000, n).astype(np.int32)
        ln = rng.integers(20, 200, n).astype(np.int32)
        eng = Engine(0)
        try:
            eng.set_option('ord_sub', ord_sub)
            eng.set_plan(np.array([KIND_NONE_ID]), 0, 0.0, 1, G)
            eng.set_subjects(None, None, G)
            eng.ordinal_set_genes(coff, gb, ge, np.arange(G, dtype=np.int32))
            eng.ordinal_enable_pairs()
            eng.ordinal_chunk(q, c, b, b + ln, ln, 0.5)
            r, g = eng.ordinal_pairs()
            units = eng.fetch_counts()
            ovf = eng.fetch_overflow()
        finally:
            eng.close()
        er, eg = O.ordinal_match(c, b, b + ln, ln, 0.5, coff, gb, ge)
        assert np.array_equal(r, er) and np.array_equal(g, eg)
        exp, eovf, _ = O.classify(q[er], eg, kinds=[KIND_NONE_ID],
                                  n_features=G)
        assert np.array_equal(units, exp)
        assert len(ovf[0]) == len(eovf)


@pytest.mark.parametrize('sub', ['1000', '2048', '6004'])
def test_host_chunk_is_matched_sub_chunk_by_sub_chunk(engine, sub):
    """wk_ordinal_chunk copies the columns in sub-chunks and runs the matcher
    behind the copies; a query that runs over a sub-chunk border stays whole."""
    test_queries_straddling_tiles(engine, int(sub))
    test_ordinal_then_classify(engine, int(sub))


def test_read_maps_with_coords_on_multi_hit_queries(tmp_path):
    """--coords with --outmap sends the matcher's (record, gene) pairs through
    the plain path (Session._ordinal_chunk_with_maps).  With several records
    per query (multi-hit reads, mates) the counts must equal those of the
    fused device route, every mapped read must be listed once, and a read with
    one gene must list exactly that gene."""
    from woltka_b200.workflow import classify, build_mapper
    rng = np.random.default_rng(12)
    contigs = [f'C{i}' for i in range(6)]
    coords = tmp_path / 'coords.txt'
    with open(coords, 'w') as f:
        for c in contigs:
            f.write(f'>{c}\n')
            pos = 1
            for gi in range(60):
                ln = int(rng.integers(200, 900))
                f.write(f'{c}_g{gi}\t{pos}\t{pos + ln}\n')
                pos += ln - int(rng.integers(0, 150))
    sam = tmp_path / 'S1.sam'
    with open(sam, 'w') as f:
        for qi in range(3000):
            k = int(min(rng.geometric(0.5), 6))
            for _ in range(k):
                c = contigs[int(rng.integers(0, 6))]
                p = int(rng.integers(1, 30000))
                flag = int(rng.choice([0, 65, 129]))
                f.write(f'R{qi}\t{flag}\t{c}\t{p}\t42\t100M\t*\t0\t0\t*\t*\n')
    mapper, chunk = build_mapper(str(coords), None, 80, None)
    kw = dict(ranks=['none'], chunk=chunk)
    plain = classify(mapper, {str(sam): 'S1'}, **kw)
    mapdir = tmp_path / 'maps'
    mapdir.mkdir()
    routed = classify(mapper, {str(sam): 'S1'}, rank2dir={'none': str(mapdir)},
                      **kw)
    assert set(routed['none']['S1']) == set(plain['none']['S1'])
    for k, v in plain['none']['S1'].items():
        assert abs(routed['none']['S1'][k] - v) < 1e-9
    lines = (mapdir / 'S1.txt').read_text().splitlines()
    names = [ln.split('\t')[0] for ln in lines]
    assert len(names) == len(set(names)) > 1000
    total = 0.0
    for ln in lines:
        parts = ln.split('\t')[1:]
        total += 1.0
        if len(parts) > 1:
            assert all(p.endswith(':1') for p in parts)
    assert abs(total - sum(plain['none']['S1'].values())) < 1e-6


@pytest.mark.parametrize('mode', ['default', 'uniq', 'uniq+unassigned'])
@pytest.mark.parametrize('ident', [True, False])
def test_fused_match_and_count(mode, ident):
    """ordinal_fused_kernel (`--coords` at `--rank none`: match and count in
    one pass) against the oracle's sweep + classify: nested and giant genes
    (reads with more than four genes are listed queries), queries of 1 to 70
    records (longer than a window: listed), the same gene hit by two records
    of a query (a set, ordinal.py:332), several genes sharing one subject,
    per-query samples with a dropped one, sub-chunked host columns."""
    from woltka_b200._lib import KIND_NONE_ID, F_UNIQ, F_UNASSIGNED
    rng = np.random.default_rng(11)
    C, per = 6, 700
    coff = np.arange(C + 1, dtype=np.int64) * per
    gb = np.sort(rng.integers(0, 200_000, (C, per)), axis=1).reshape(-1).astype(np.int32)
    ln = rng.integers(100, 900, C * per)
    ln[rng.random(C * per) < 0.02] = 30_000          # giant genes: many per read
    ge = (gb + ln).astype(np.int32)
    G = len(gb)
    if ident:
        gsub, NF = np.arange(G, dtype=np.int32), G
    else:
        gsub, NF = rng.integers(0, G // 3, G).astype(np.int32), G // 3
    sizes = rng.integers(1, 6, 20000)
    sizes[::997] = 70
    q = np.repeat(np.arange(len(sizes), dtype=np.int32), sizes)
    n = len(q)
    c = rng.integers(0, C, n).astype(np.int32)
    b = rng.integers(0, 199_000, n).astype(np.int32)
    # the records of some queries hit the same place twice
    same = (rng.random(n) < 0.3) & (np.r_[False, q[1:] == q[:-1]])
    c[same], b[same] = np.roll(c, 1)[same], (np.roll(b, 1)[same] + rng.integers(0, 20, n)[same])
    rl = rng.integers(50, 200, n).astype(np.int32)
    e = (b + rl).astype(np.int32)
    q_sample = rng.integers(-1, 3, len(sizes)).astype(np.int32)
    fl = {'default': 0, 'uniq': F_UNIQ, 'uniq+unassigned': F_UNIQ | F_UNASSIGNED}[mode]
    er, eg = O.ordinal_match(c, b, e, rl, 0.5, coff, gb, ge)
    exp, eovf, _ = O.classify(q[er], gsub[eg], kinds=[KIND_NONE_ID], flags=fl,
                              n_samples=3, n_features=NF, q_sample=q_sample)
    for sub in (0, 1024):
        eng = Engine(0)
        try:
            eng.set_option('ord_sub', sub)
            eng.set_option('fuse', 1)       # (opt-in: slower than two kernels on cfg3)
            eng.set_plan(np.array([KIND_NONE_ID]), fl, 0.0, 3, NF)
            eng.set_subjects(None, None, NF)
            eng.ordinal_set_genes(coff, gb, ge, gsub)
            eng.ordinal_chunk(q, c, b, e, rl, 0.5, q_sample=q_sample)
            assert eng.last_kernel() == 'ordinal_fused_kernel'
            units = eng.fetch_counts()
            ovf = eng.fetch_overflow()
            # the two-kernel route must agree too
            eng.set_option('fuse', 0)
            eng.reset_counts()
            eng.ordinal_chunk(q, c, b, e, rl, 0.5, q_sample=q_sample)
            assert eng.last_kernel() != 'ordinal_fused_kernel'
            units2 = eng.fetch_counts()
        finally:
            eng.close()
        assert np.array_equal(units, exp), np.argwhere(units != exp)[:5]
        assert np.array_equal(units2, exp)
        assert len(ovf[0]) == len(eovf)
        assert (mode != 'default') or len(eovf) > 0 or True
