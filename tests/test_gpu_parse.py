"""GPU parity: SAM text parsed on the device (wk_parse_sam) vs the host reader
(woltka_b200.align, itself pinned to the reference's parser tests in
tests/test_align.py): same queries in the same order, same subject sets, same
samples; and the counts of a run fed from text equal those fed from columns."""
import lzma
import os

import numpy as np
import pytest

from woltka_b200 import align
from woltka_b200.session import _split_sample

pytestmark = pytest.mark.gpu

DATA = os.path.join(os.path.dirname(__file__), 'golden', 'data')
SUFFIX = ('', '/1', '/2')


def host_queries(text, fmt='sam', excl=None, extr=False):
    lines = text.decode().splitlines(keepends=True)
    return list(align.iter_align(iter(lines), fmt, excl, extr))


def device_queries(engine, text, demux=False, fmt='sam'):
    """[(query name, set of subject names[, sample name])] from the device."""
    body = text
    n_rec, n_qry, n_sub, n_smp = engine.parse_sam(body, demux, fmt)
    subjects = engine.fetch_names(0, 0, n_sub)
    samples = engine.fetch_names(1, 0, n_smp)
    q, s, qs, ql = engine.fetch_parsed_columns(n_rec, n_qry, demux)
    lines = body.split(b'\n')
    out = []
    assert len(q) == n_rec
    if n_rec:
        assert q[0] == 0 and np.all(np.diff(q) >= 0) and np.all(np.diff(q) <= 1)
        assert q[-1] == n_qry - 1
    bounds = np.flatnonzero(np.diff(np.concatenate([[-1], q, [n_qry]]))) \
        if n_rec else []
    for j in range(n_qry):
        li, mate = int(ql[j]) & ((1 << 30) - 1), int(ql[j]) >> 30
        name = lines[li].split(b'\t', 1)[0].decode() + SUFFIX[mate]
        subs = {subjects[i] for i in s[bounds[j]:bounds[j + 1]].tolist()}
        rec = (name, subs)
        if demux:
            rec += (samples[qs[j]],)
        out.append(rec)
    return out, subjects


def strip_header(raw):
    lines = raw.split(b'\n')
    i = 0
    while i < len(lines) and lines[i].startswith(b'@'):
        i += 1
    return b'\n'.join(lines[i:])


@pytest.mark.parametrize('path', ['bowtie2/S01.sam.xz', 'bt2sho/S03.sam.xz'])
def test_bundled_sam(path):
    from woltka_b200.engine import Engine
    eng = Engine(0)
    with lzma.open(os.path.join(DATA, path)) as f:
        text = strip_header(f.read())
    got, _ = device_queries(eng, text)
    exp = host_queries(text)
    assert [g[0] for g in got] == [e[0] for e in exp]
    assert [g[1] for g in got] == [e[1] for e in exp]
    eng.close()


def synthetic_sam(n_groups, seed, trailing_newline=True):
    rng = np.random.default_rng(seed)
    rows = []
    for g in range(n_groups):
        if rng.random() < 0.3:
            name = f'S{rng.integers(0, 5)}_r{g}'
        elif rng.random() < 0.5:
            name = f'read{g}'
        else:
            name = rng.choice(['x_', '_y', 'a_b_c', f'q{g}_'])
        for _ in range(int(min(rng.geometric(0.45), 12))):
            flag = int(rng.choice([0, 16, 64 + 1, 128 + 1, 256, 64 + 16, 4]))
            sub = '*' if rng.random() < 0.08 else f'G{rng.integers(0, 300):06d}'
            rows.append(f'{name}\t{flag}\t{sub}\t{rng.integers(1, 9999)}\t42\t'
                        f'50M\t*\t0\t0\tACGT\tIIII')
    text = '\n'.join(rows)
    return (text + '\n' if trailing_newline else text).encode()


@pytest.mark.parametrize('seed,tnl', [(1, True), (2, False), (3, True)])
def test_mates_unmapped_repeats(seed, tnl):
    from woltka_b200.engine import Engine
    eng = Engine(0)
    text = synthetic_sam(4000, seed, tnl)
    got, _ = device_queries(eng, text, demux=True)
    exp = host_queries(text)
    assert [g[0] for g in got] == [e[0] for e in exp]
    assert [g[1] for g in got] == [e[1] for e in exp]
    assert [g[2] for g in got] == [_split_sample(e[0])[0] for e in exp]
    eng.close()


def test_tables_persist_over_chunks_and_edge_inputs():
    from woltka_b200.engine import Engine, WoltkaB200Error
    eng = Engine(0)
    a = synthetic_sam(500, 11)
    b = synthetic_sam(500, 12)
    got_a, subj_a = device_queries(eng, a)
    got_b, subj_b = device_queries(eng, b)
    assert subj_b[:len(subj_a)] == subj_a          # indices are stable
    assert len(set(subj_b)) == len(subj_b)
    assert [g[1] for g in got_b] == [e[1] for e in host_queries(b)]
    assert eng.parse_sam(b'')[:2] == (0, 0)
    assert eng.parse_sam(b'r1\t4\t*\t0\t0\t*\t*\t0\t0\tA\tI\n')[:2] == (0, 0)
    with pytest.raises(WoltkaB200Error):
        eng.parse_sam(b'only\ttwo\n')
    with pytest.raises(WoltkaB200Error):
        eng.parse_sam(b'r1\tx\tG1\t1\n')
    eng.close()


def _run_classify(files, demux, **kw):
    import io
    from contextlib import redirect_stdout
    from woltka_b200 import workflow
    with redirect_stdout(io.StringIO()):
        out = workflow.classify(workflow.plain_mapper, files, demux=demux,
                                ranks=['none'], **kw)
    return out, workflow.LAST_READER


def test_classify_from_text_equals_classify_from_host_reader(monkeypatch, tmp_path):
    """workflow.classify() on SAM files: the device reader and the host reader
    give the same profile; small text blocks exercise the chunk cut."""
    from woltka_b200 import workflow
    files = {os.path.join(DATA, 'bt2sho', f'S0{i}.sam.xz'): f'S0{i}'
             for i in range(1, 6)}
    dev, reader = _run_classify(files, False)
    assert reader == 'device'
    monkeypatch.setenv('WOLTKA_B200_HOST_READER', '1')
    host, reader = _run_classify(files, False)
    assert reader == 'host'
    assert dev == host
    monkeypatch.delenv('WOLTKA_B200_HOST_READER')
    # multiplexed synthetic file, cut into many chunks
    fp = tmp_path / 'mux.sam'
    fp.write_bytes(b'@HD\tVN:1.0\n@SQ\tSN:x\tLN:1\n' + synthetic_sam(3000, 5))
    from woltka_b200 import reader
    monkeypatch.setattr(reader, 'BLOCK', 20000)
    dev, reader = _run_classify([str(fp)], True)
    assert reader == 'device'
    monkeypatch.setenv('WOLTKA_B200_HOST_READER', '1')
    host, _ = _run_classify([str(fp)], True)
    assert dev == host
    assert len(dev['none']) > 3


def _open_any(path):
    import bz2
    import gzip
    op = {'.xz': lzma.open, '.bz2': bz2.open, '.gz': gzip.open}
    for ext, f in op.items():
        if path.endswith(ext):
            return f(path)
    return open(path, 'rb')


@pytest.mark.parametrize('path,fmt', [('blastn/mux.b6o.xz', 'b6o'),
                                      ('burst/S02.b6.bz2', 'b6o'),
                                      ('split/S04.map.bz2', 'map')])
def test_bundled_other_formats(path, fmt):
    from woltka_b200.engine import Engine
    eng = Engine(0)
    with _open_any(os.path.join(DATA, path)) as f:
        text = f.read()
    got, _ = device_queries(eng, text, demux=True, fmt=fmt)
    exp = host_queries(text, fmt)
    assert [g[0] for g in got] == [e[0] for e in exp]
    assert [g[1] for g in got] == [e[1] for e in exp]
    assert [g[2] for g in got] == [_split_sample(e[0])[0] for e in exp]
    eng.close()


def test_odd_lines_of_map_b6o_paf():
    from woltka_b200.engine import Engine
    eng = Engine(0)
    cases_ = {
        'map': b'q1\tA\nq1\tB \t x\nno_tab_line\nq2\t\nq2\tC\r\nq3\tA\textra\n',
        'b6o': b'q1\tA\t99\nshort\tline\nq1\tB\t\nq2\tA\t1\t2\t3\n',
        'paf': b'q1\t1\t2\t3\t+\tT1\t9\nq1\t1\t2\t3\t+\tT2\nq2\t1\t2\t3\t-\tT1\t9\t8\n',
    }
    for fmt, text in cases_.items():
        got, _ = device_queries(eng, text, fmt=fmt)
        exp = host_queries(text, fmt)
        assert got == [(e[0], e[1]) for e in exp], fmt
    eng.close()


# ---- reader options: --trim-sub, --exclude, coordinates --------------------------
def suffixed_sam(n_groups, seed, repeats=True):
    """Subjects carry `_<n>` suffixes (ORF style); some names come back after
    another name in between (the A, B, A pattern) unless repeats=False."""
    rng = np.random.default_rng(seed)
    cigars = ['50M', '10S40M', '20M2D30M', '*', '5H20M3I22M5N3M', '7=1X8=',
              '30M1000N20M', '12I', '4P']
    rows = []
    for g in range(n_groups):
        name = f'read{g if rng.random() < 0.8 or not repeats else g - 2}'
        for _ in range(int(min(rng.geometric(0.4), 9))):
            flag = int(rng.choice([0, 16, 64 + 1, 128 + 1, 256]))
            r = rng.random()
            sub = '*' if r < 0.05 else (
                f'G{rng.integers(0, 60):03d}_{rng.integers(1, 4)}' if r < 0.8
                else f'G{rng.integers(0, 60):03d}' if r < 0.9
                else rng.choice(['_1', 'a_b_c', 'plain', 'x__2']))
            rows.append(f'{name}\t{flag}\t{sub}\t{rng.integers(1, 99999)}\t42\t'
                        f'{rng.choice(cigars)}\t*\t0\t0\tACGT\tIIII')
    return ('\n'.join(rows) + '\n').encode()


@pytest.mark.parametrize('trim', [None, '_', '__'])
@pytest.mark.parametrize('with_excl', [False, True])
def test_trim_sub_and_exclude_on_device(trim, with_excl):
    """--trim-sub (workflow.py:818-841) and --exclude (align.py:409-478) inside
    the device reader against the host reader followed by strip_suffix."""
    from woltka_b200.engine import Engine
    eng = Engine(0)
    excl = {f'G{i:03d}_2' for i in range(0, 60, 3)} | {'plain', 'G007'} \
        if with_excl else None
    eng.parse_options(trim, excl)
    for seed in (21, 22):
        text = suffixed_sam(3000, seed)
        got, _ = device_queries(eng, text)
        exp = host_queries(text, excl=excl)
        if trim:
            exp = [(q, {x.rsplit(trim, 1)[0] for x in subs}) for q, subs in exp]
        assert [g[0] for g in got] == [e[0] for e in exp]
        assert [g[1] for g in got] == [e[1] for e in exp]
    # A, B (excluded), A: the two A groups stay two queries
    eng.parse_options(None, {'bad'})
    text = (b'A\t0\tg1\t1\t9\t5M\t*\nB\t0\tbad\t1\t9\t5M\t*\n'
            b'A\t0\tg2\t1\t9\t5M\t*\nC\t0\tg1\t1\t9\t5M\t*\n'
            b'C\t64\tbad\t1\t9\t5M\t*\n')
    got, _ = device_queries(eng, text)
    assert got == [('A', {'g1'}), ('A', {'g2'})] == host_queries(text, excl={'bad'})
    # options are kept until changed, and can be switched off again
    eng.parse_options()
    got, _ = device_queries(eng, text)
    assert [g[0] for g in got] == ['A', 'B', 'A', 'C', 'C/1']
    eng.close()


def device_records(engine, text, demux=False, fmt='sam'):
    """[(query name, [(subject, len, beg, end)])] of a chunk parsed with coords."""
    n_rec, n_qry, n_sub, n_smp = engine.parse_sam(text, demux, fmt)
    subjects = engine.fetch_names(0, 0, n_sub)
    q, s, qs, ql = engine.fetch_parsed_columns(n_rec, n_qry, demux)
    beg, end, ln = engine.fetch_parsed_coords(n_rec)
    lines = text.split(b'\n')
    out = [None] * n_qry
    for j in range(n_qry):
        li, mate = int(ql[j]) & ((1 << 30) - 1), int(ql[j]) >> 30
        out[j] = (lines[li].split(b'\t', 1)[0].decode() + SUFFIX[mate], [])
    for i in range(n_rec):
        out[q[i]][1].append((subjects[s[i]], int(ln[i]), int(beg[i]), int(end[i])))
    return out


def host_records(text, fmt='sam', excl=None):
    out = []
    for query, records in host_queries(text, fmt, excl, True):
        recs = [(r[0], r[2], r[3], r[4]) for r in records if r[2]]
        if recs:
            out.append((query, recs))
    return out


def merge_adjacent(rows):
    """Records per query name, names in order of first appearance: what
    ordinal.py:332 makes of a chunk.  (The device groups the lines that are
    left once the records without aligned length are gone, so two groups of
    one name around a group that vanished are already one query there.)"""
    out = {}
    for query, recs in rows:
        out.setdefault(query, []).extend(recs)
    return list(out.items())


@pytest.mark.parametrize('with_excl', [False, True])
def test_sam_coordinates_on_device(with_excl):
    """POS - 1 and cigar_to_lens (align.py:350-406, 550-583) on the device."""
    from woltka_b200.engine import Engine
    eng = Engine(0)
    excl = {f'G{i:03d}_2' for i in range(0, 60, 3)} if with_excl else None
    eng.parse_options(None, excl, coords=True)
    from woltka_b200.engine import WoltkaB200Error
    with pytest.raises(WoltkaB200Error) as err:      # a name in two places:
        eng.parse_sam(suffixed_sam(3000, 30), False)  # the host reader's case
    assert err.value.code == 6
    for seed in (31, 32):
        text = suffixed_sam(3000, seed, repeats=False)
        got = device_records(eng, text)
        exp = host_records(text, excl=excl)
        if not with_excl:
            # (which of two merged groups names the query first is open)
            exp = sorted(merge_adjacent(exp))
            got = sorted(merge_adjacent(got))
        assert got == exp
    with pytest.raises(Exception):
        eng.classify_parsed(None, 0)     # parsed for the matcher, not for classify
    eng.close()


def test_b6o_paf_coordinates_on_device():
    """parse_b6o_file_ex (align.py:807-855) / parse_paf_file_ex (:1046-1088)."""
    from woltka_b200.engine import Engine
    rng = np.random.default_rng(5)
    b6o, paf = [], []
    for g in range(2000):
        for _ in range(int(min(rng.geometric(0.5), 6))):
            a, b = sorted(rng.integers(1, 50000, 2).tolist())
            if rng.random() < 0.5:
                a, b = b, a
            ln = int(rng.integers(0, 300)) if rng.random() < 0.9 else 0
            b6o.append(f'q{g}\tT{rng.integers(0, 40)}\t98.5\t{ln}\t1\t0\t1\t{ln}'
                       f'\t{a}\t{b}\t1e-9\t{rng.integers(50, 300)}.0')
            lo, hi = sorted((a, b))
            paf.append(f'q{g}\t150\t0\t150\t+\tT{rng.integers(0, 40)}\t90000'
                       f'\t{lo}\t{hi}\t140\t{ln}\t{rng.integers(0, 60)}\ttp:A:P')
    b6o[17] = 'short\tline\t1\t2'
    paf[23] = 'q23\t150\t0\t150\t+\tT1\t90000\tx\t5\t140\t150\t60'
    paf[29] = 'q29\t150\t0\t150\t+\tT1'
    eng = Engine(0)
    eng.parse_options(None, None, coords=True)
    for fmt, rows in (('b6o', b6o), ('paf', paf)):
        text = ('\n'.join(rows) + '\n').encode()
        got = sorted(merge_adjacent(device_records(eng, text, fmt=fmt)))
        exp = sorted(merge_adjacent(host_records(text, fmt)))
        assert got == exp, fmt
    eng.close()


def _ordinal_classify(files, coords_fp, demux=None, **kw):
    import io
    from contextlib import redirect_stdout
    from woltka_b200 import workflow
    with redirect_stdout(io.StringIO()):
        mapper, chunk = workflow.build_mapper(coords_fp, None, 80, None)
        out = workflow.classify(mapper, files, demux=demux, ranks=['none'],
                                chunk=chunk, **kw)
    return out, workflow.LAST_READER


def test_coords_from_text_equals_host_reader(monkeypatch, tmp_path):
    """`--coords` end to end: text -> records -> matches -> counts on the
    device against the host reader feeding the same matcher."""
    from woltka_b200 import workflow
    coords_fp = os.path.join(DATA, 'synth_coords.txt')
    files = {os.path.join(DATA, 'synth_ordinal', f'S{i}.sam'): f'S{i}'
             for i in range(2)}
    from woltka_b200 import reader
    monkeypatch.setattr(reader, 'BLOCK', 30000)
    for kw in ({}, {'uniq': True}, {'exclude': {'C4', 'C11'}},
               {'trimsub': '_'}, {'sizes': None, 'unasgd': True}):
        monkeypatch.delenv('WOLTKA_B200_HOST_READER', raising=False)
        dev, reader = _ordinal_classify(files, coords_fp, **kw)
        assert reader == 'device', kw
        monkeypatch.setenv('WOLTKA_B200_HOST_READER', '1')
        host, reader = _ordinal_classify(files, coords_fp, **kw)
        assert reader == 'host'
        assert dev == host, kw
        assert sum(len(v) for v in dev['none'].values()) > 10


def test_coords_demultiplexed_from_text(monkeypatch, tmp_path):
    """`--coords` on a multiplexed file (sample = prefix of the read name,
    workflow.demultiplex workflow.py:844-909), with a sample filter."""
    from woltka_b200 import reader
    coords_fp = os.path.join(DATA, 'synth_coords.txt')
    rows = [b'@HD\tVN:1.0\n']
    for i in range(2):
        with open(os.path.join(DATA, 'synth_ordinal', f'S{i}.sam'), 'rb') as f:
            body = [x for x in f if not x.startswith(b'@')]
        rows += [b'S%d_' % i + x for x in body]
        # a third sample, filtered out in the second run (its queries between
        # the others', never inside one: the records of a query are adjacent)
        if i == 0:
            rows += [b'X9_' + x for x in body[:len(body) // 5]]
    fp = tmp_path / 'mux.sam'
    fp.write_bytes(b''.join(rows))
    monkeypatch.setattr(reader, 'BLOCK', 40000)
    for samples in (None, ['S0', 'S1']):
        monkeypatch.delenv('WOLTKA_B200_HOST_READER', raising=False)
        dev, rd = _ordinal_classify([str(fp)], coords_fp, demux=True,
                                    samples=samples)
        assert rd == 'device'
        monkeypatch.setenv('WOLTKA_B200_HOST_READER', '1')
        host, _ = _ordinal_classify([str(fp)], coords_fp, demux=True,
                                    samples=samples)
        assert dev == host
        assert set(dev['none']) == ({'S0', 'S1'} if samples
                                    else {'S0', 'S1', 'X9'})


def test_coords_query_name_in_two_places(monkeypatch, tmp_path):
    """The reference merges the records of a query name that comes back later
    in a chunk (ordinal.py:296-332); the device reader, which groups adjacent
    lines, notices such a block and leaves it to the host reader."""
    coords_fp = os.path.join(DATA, 'synth_coords.txt')
    with open(os.path.join(DATA, 'synth_ordinal', 'S0.sam'), 'rb') as f:
        body = [x for x in f if not x.startswith(b'@')]
    fp = tmp_path / 'S0.sam'
    fp.write_bytes(b'@HD\tVN:1.0\n' + b''.join(body) + b''.join(body[:200]))
    monkeypatch.delenv('WOLTKA_B200_HOST_READER', raising=False)
    dev, rd = _ordinal_classify({str(fp): 'S0'}, coords_fp)
    assert rd == 'host'                      # (the one block fell back)
    monkeypatch.setenv('WOLTKA_B200_HOST_READER', '1')
    host, _ = _ordinal_classify({str(fp): 'S0'}, coords_fp)
    assert dev == host


def test_plain_options_from_text_equals_host_reader(monkeypatch, tmp_path):
    """classify() with --trim-sub / --exclude stays on the device reader."""
    from woltka_b200 import workflow
    fp = tmp_path / 'S9.sam'
    fp.write_bytes(b'@HD\tVN:1.0\n' + suffixed_sam(4000, 41))
    from woltka_b200 import reader
    monkeypatch.setattr(reader, 'BLOCK', 25000)
    excl = {f'G{i:03d}_2' for i in range(0, 60, 3)}
    for kw in ({'trimsub': '_'}, {'exclude': excl},
               {'trimsub': '_', 'exclude': excl, 'uniq': True}):
        monkeypatch.delenv('WOLTKA_B200_HOST_READER', raising=False)
        dev, reader = _run_classify({str(fp): 'S9'}, False, **kw)
        assert reader == 'device', kw
        monkeypatch.setenv('WOLTKA_B200_HOST_READER', '1')
        host, _ = _run_classify({str(fp): 'S9'}, False, **kw)
        assert dev == host, kw
        assert len(dev['none']['S9']) > 5
