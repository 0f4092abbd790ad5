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


def host_queries(text, fmt='sam'):
    lines = text.decode().splitlines(keepends=True)
    return list(align.iter_align(iter(lines), fmt))


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
    orig = workflow._text_chunks
    monkeypatch.setattr(workflow, '_text_chunks',
                        lambda p, header=True: orig(p, block=20000, header=header))
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
