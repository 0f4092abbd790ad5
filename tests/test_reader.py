"""Host side of the block reader (woltka_b200/reader.py): every byte of the
body reaches the parser exactly once, blocks start at line starts, '@' header
lines are skipped, whatever the block size; a stand-in for the device cuts
the blocks the way wk_parse_block does (start of the last query group)."""
import gzip
import os
import random

import numpy as np
import pytest

from woltka_b200 import reader


@pytest.fixture
def plain_buffers(monkeypatch):
    import woltka_b200.engine as engine
    monkeypatch.setattr(reader, '_pinned_pair', lambda room, block: (
        [np.empty(room + block, np.uint8) for _ in (0, 1)], None))
    monkeypatch.setattr(engine, 'pinned_empty', lambda n, dt: np.empty(n, dt))


def device_cut(text):
    """Bytes wk_parse_block consumes of a block that is not the last."""
    end = text.rfind(b'\n') + 1
    lines = text[:end].splitlines(True)
    if not lines:
        return 0
    name, k = lines[-1].split(b'\t', 1)[0], len(lines)
    while k > 0 and lines[k - 1].split(b'\t', 1)[0] == name:
        k -= 1
    return sum(map(len, lines[:k]))


@pytest.mark.parametrize('ext', ['', '.gz'])
def test_blocks_cover_the_body_once(tmp_path, plain_buffers, ext):
    random.seed(1)
    head = [b'@HD\tx\n', b'@SQ\t' + b'y' * 5000 + b'\n']
    body = []
    for g in range(3000):
        for k in range(random.randint(1, 5)):
            body.append(b'r%d\t0\tG%d\t%s\n' % (g, k, b'z' * random.randint(0, 200)))
    body.append(b'last\t0\tG\tno newline')
    fp = str(tmp_path / ('a.sam' + ext))
    with (gzip.open if ext else open)(fp, 'wb') as f:
        f.write(b''.join(head + body))
    for block, room in ((1000, 64), (4096, 1024), (1 << 20, None), (300, 16)):
        rd = reader.BlockReader(fp, header=True, block=block, room=room)
        out, finals = [], 0
        for view, final in rd:
            text = view.tobytes()
            assert not out or out[-1][-1:] in (b'', b'\n')
            used = len(text) if final else device_cut(text)
            finals += final
            out.append(text[:used])
            rd.consumed(used)
        assert finals == 1
        assert b''.join(out) == b''.join(body), (ext, block)


def test_header_only_and_empty_files(tmp_path, plain_buffers):
    for data in (b'', b'@HD\tonly\n', b'@HD\tno newline'):
        fp = str(tmp_path / 'h.sam')
        with open(fp, 'wb') as f:
            f.write(data)
        assert list(reader.BlockReader(fp, header=True, block=256)) == []
    fp = str(tmp_path / 'm.map')
    with open(fp, 'wb') as f:
        f.write(b'@q\tS\n')          # not a header in a format without one
    got = [(v.tobytes(), f) for v, f in reader.BlockReader(fp, header=False)]
    assert got == [(b'@q\tS\n', True)]


def test_header_longer_than_a_block_and_host_cut(tmp_path, plain_buffers):
    """Every leading '@' line of a SAM file is dropped even when a block ends
    inside one (align.py:296-300 skips them line by line); the host's cut
    (the fallback when the device reader's tables are full) never splits a
    query (align.py:73-79)."""
    from woltka_b200.workflow import _host_cut
    hdr = ''.join(f'@SQ\tSN:contig{i}\tLN:{1000 + i}\n' for i in range(20))
    body = ''.join(f'r{i // 2}\t0\tG{i % 5}\t1\t42\t50M\t*\t0\t0\tA\tI\n'
                   for i in range(40))
    fp = tmp_path / 'x.sam'
    fp.write_text(hdr + body)
    for block in (7, 16, 50, 64, 1000, 1 << 20):
        rd = reader.BlockReader(str(fp), header=True, block=block, room=4)
        chunks = []
        for view, final in rd:
            text = view.tobytes()
            used = len(text) if final else _host_cut(text)
            assert used == (len(text) if final else device_cut(text))
            rd.consumed(used)
            if used:
                chunks.append(text[:used])
        assert b''.join(chunks) == body.encode(), block
        for a, b in zip(chunks, chunks[1:]):
            assert a.splitlines()[-1].split(b'\t')[0] != \
                b.splitlines()[0].split(b'\t')[0]


def test_two_readers_at_once_do_not_share_buffers(tmp_path, monkeypatch):
    """The process keeps one pair of page-locked buffers; a second reader
    that starts while the first is still iterating gets its own."""
    import woltka_b200.engine as engine
    monkeypatch.setattr(engine, 'pinned_empty', lambda n, dt: np.empty(n, dt))
    monkeypatch.setattr(reader, '_buffers', {})
    monkeypatch.setattr(reader, '_busy', set())
    a, b = tmp_path / 'a.map', tmp_path / 'b.map'
    a.write_bytes(b''.join(b'a%d\tS\n' % i for i in range(2000)))
    b.write_bytes(b''.join(b'b%d\tT\n' % i for i in range(2000)))
    ra = reader.BlockReader(str(a), header=False, block=4096, room=256)
    rb = reader.BlockReader(str(b), header=False, block=4096, room=256)
    out_a, out_b = [], []
    ita, itb = iter(ra), iter(rb)
    for (va, fa), (vb, fb) in zip(ita, itb):
        ta, tb = va.tobytes(), vb.tobytes()
        ua = len(ta) if fa else ta.rfind(b'\n') + 1
        ub = len(tb) if fb else tb.rfind(b'\n') + 1
        out_a.append(ta[:ua])
        out_b.append(tb[:ub])
        ra.consumed(ua)
        rb.consumed(ub)
    assert b''.join(out_a) == a.read_bytes() and b''.join(out_b) == b.read_bytes()
    for it in (ita, itb):
        it.close()
    assert not reader._busy
