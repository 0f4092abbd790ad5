"""Alignment files as blocks of bytes for the device reader.

Replaces the file handle iteration + chunk packing of the reference's mappers
(align.plain_mapper, /root/reference/woltka/align.py:47-115, and
ordinal.ordinal_mapper, ordinal.py:167-240) on the way INTO the GPU: the host
only moves bytes.  A block is read straight into page-locked memory (several
threads, `os.preadv`, for plain files; one decompressing thread otherwise)
while the device parses the previous one; where a block is cut — at a line
end, and never inside a query — is decided by the device (wk_parse_block),
which returns how many bytes it consumed; the rest is copied in front of the
next block.
"""
import bz2
import gzip
import lzma
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_OPENERS = {'.gz': gzip.open, '.bz2': bz2.open, '.xz': lzma.open,
            '.lzma': lzma.open}

BLOCK = 64 << 20      # bytes read per block
ROOM = 8 << 20        # space in front of a block for what the last one left
_THREADS = 4

# page-locked buffers are expensive to make (the pages are pinned one by one):
# a process keeps the pair it made
_buffers = {}
_pool = None


def _executor():
    global _pool
    if _pool is None:
        _pool = ThreadPoolExecutor(max_workers=_THREADS + 1)
    return _pool


_busy = set()          # keys of cached pairs a reader is iterating over


def _pinned_pair(room, block):
    """Two page-locked buffers of room + block bytes: the process's cached
    pair when it is free (one reader at a time uses it), else a new pair."""
    from .engine import pinned_empty
    key = (room, block)
    if key in _busy:
        return [pinned_empty(room + block, np.uint8) for _ in (0, 1)], None
    if key not in _buffers:
        for k in [k for k in _buffers if k not in _busy]:
            del _buffers[k]
        _buffers[key] = [pinned_empty(room + block, np.uint8) for _ in (0, 1)]
    _busy.add(key)
    return _buffers[key], key


class BlockReader:
    """for view, final in reader: used = parse(view, final); reader.consumed(used)

    `view` is a uint8 array over page-locked memory holding whole lines from
    the first line that has not been consumed yet; leading '@' lines of a SAM
    file are skipped (align.py:296-300)."""

    def __init__(self, fp, header=True, block=None, room=None):
        self.fp = fp
        self.header = header
        self.block = block or BLOCK
        self.room = room or min(ROOM, max(self.block // 4, 1 << 12))
        self.plain = not any(fp.endswith(ext) for ext in _OPENERS)
        self._used = None

    # -- reading ---------------------------------------------------------------
    def _open(self):
        if self.plain:
            self.fd = os.open(self.fp, os.O_RDONLY)
            self.size = os.fstat(self.fd).st_size
            self.offset = 0
        else:
            for ext, opener in _OPENERS.items():
                if self.fp.endswith(ext):
                    self.fh = opener(self.fp, 'rb')

    def _close(self):
        if self.plain:
            os.close(self.fd)
        else:
            self.fh.close()

    def _read_plain(self, dest, offset):
        """dest[:] <- file[offset:offset+len(dest)], in parallel slices."""
        n = len(dest)
        step = max((n + _THREADS - 1) // _THREADS, 1 << 20)
        mv = memoryview(dest)

        def part(a):
            got, want = 0, min(step, n - a)
            while got < want:
                k = os.preadv(self.fd, [mv[a + got:a + want]], offset + a + got)
                if k <= 0:
                    raise OSError(f'short read of {self.fp}')
                got += k
        list(_executor().map(part, range(0, n, step)))
        return n

    def _read_stream(self, dest):
        mv, got = memoryview(dest), 0
        while got < len(dest):
            k = self.fh.readinto(mv[got:])
            if not k:
                break
            got += k
        return got

    def _fetch(self, buf):
        """Start filling buf[room:room+block]; the future gives (n, final)."""
        dest = buf[self.room:self.room + self.block]
        if self.plain:
            off = self.offset
            n = max(min(self.block, self.size - off), 0)
            self.offset += n
            final = self.offset >= self.size
            return _executor().submit(
                lambda: (self._read_plain(dest[:n], off) if n else 0, final))

        def work():
            n = self._read_stream(dest)
            return n, n < len(dest)
        return _executor().submit(work)

    # -- iteration ---------------------------------------------------------------
    def consumed(self, used):
        self._used = int(used)

    def __iter__(self):
        self._open()
        pending = held = None
        try:
            bufs, held = _pinned_pair(self.room, self.block)
            cur = 0
            pending = self._fetch(bufs[cur])
            start = self.room
            in_header = self.header
            while True:
                buf = bufs[cur]
                (n, final), pending = pending.result(), None
                end = self.room + n
                if not final:
                    pending = self._fetch(bufs[1 - cur])
                if in_header:
                    start, in_header = self._skip_header(buf, start, end)
                if not in_header and end > start:
                    self._used = None
                    yield buf[start:end], final
                    if self._used is None and not final:
                        raise RuntimeError('BlockReader.consumed() not called')
                    start += self._used or 0
                if final:
                    return
                left = end - start
                if left > self.room:
                    # a query (or a header) longer than the room in front of a
                    # block: a pair of buffers with more room, the block that
                    # is being fetched moves over
                    n2, final2 = pending.result()
                    room = self.room
                    while room < left:
                        room *= 2
                    from .engine import pinned_empty
                    new = [pinned_empty(room + self.block, np.uint8)
                           for _ in (0, 1)]
                    new[1 - cur][room:room + n2] = \
                        bufs[1 - cur][self.room:self.room + n2]
                    bufs, self.room = new, room
                    pending = _executor().submit(lambda r=(n2, final2): r)
                bufs[1 - cur][self.room - left:self.room] = buf[start:end]
                start = self.room - left
                cur = 1 - cur
        finally:
            if pending is not None:
                try:
                    pending.result()     # nobody writes into the buffers now
                except Exception:
                    pass
            _busy.discard(held)
            self._close()

    def _skip_header(self, buf, start, end):
        """Skip complete '@' lines; (new start, still inside the header)."""
        while start < end and buf[start] == 64:          # '@'
            stop = min(end, start + (1 << 16))
            nl = np.flatnonzero(buf[start:stop] == 10)
            while not len(nl) and stop < end:
                a, stop = stop, min(end, stop + (1 << 20))
                nl = np.flatnonzero(buf[a:stop] == 10) + (a - start)
            if not len(nl):
                return start, True       # the line goes on in the next block
            start += int(nl[0]) + 1
        return start, start >= end
