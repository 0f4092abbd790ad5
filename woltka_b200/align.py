"""Alignment readers feeding the GPU path (host side, text -> queries).

Mirrors the reader interface of the reference (`plain_mapper`, `iter_align`,
`infer_align_format`; /root/reference/woltka/align.py:47-223) so that
`build_mapper()` / `classify()` are drop-in, but is organised differently:
one table of per-format field extractors and ONE grouping loop, instead of
the reference's sixteen hand-unrolled parser functions.

Grouping rules restated from the reference:
  * a query is a run of ADJACENT lines with the same query name
    (align.py:325-339, :656-666, :794-802, :1036-1044);
  * SAM: '@' lines before the body are skipped, RNAME '*' is skipped, the
    mate is (FLAG >> 6) & 3 and mates 1/2 become queries "name/1", "name/2",
    emitted after the unpaired pool (align.py:295-347);
  * plain mode collects a SET of subjects, extra mode (ordinal) a LIST of
    (subject, score, length, beg, end) with 0-based begin / exclusive end
    (align.py:382-398, :832, :1067);
  * with an exclusion set, a query name any of whose records hits an excluded
    subject is dropped entirely, all mates included (align.py:409-478).
    Deviation: the reference's parse_sam_file_ex_ft forgets that flag on the
    very last query of a file (align.py:542-547); this reader drops it like
    every other query.
"""
import re
from functools import lru_cache
from itertools import chain

__all__ = ['plain_mapper', 'iter_align', 'infer_align_format', 'cigar_to_lens']


# ---- per-format field extractors ---------------------------------------------
# each returns None (skip the line) or (qname, mate, subject, record);
# `record` is only built when extr is true.

def _sam(line, extr):
    if extr:
        qname, flag, rname, pos, _, cigar, _ = line.split('\t', 6)
    else:
        qname, flag, rname, _ = line.split('\t', 3)
    if rname == '*':
        return None
    mate = int(flag) >> 6 & 3
    if not extr:
        return qname, mate, rname, None
    beg = int(pos) - 1
    length, span = cigar_to_lens(cigar)
    return qname, mate, rname, (rname, None, length, beg, beg + span)


def _b6o(line, extr):
    if not extr:
        try:
            qseqid, sseqid, _ = line.split('\t', 2)
        except ValueError:
            return None
        return qseqid, 0, sseqid, None
    x = line.split('\t')
    if len(x) < 12:
        return None
    length, score = int(x[3]), float(x[11])
    a, b = int(x[8]), int(x[9])
    lo, hi = (a, b) if a <= b else (b, a)
    return x[0], 0, x[1], (x[1], score, length, lo - 1, hi)


def _paf(line, extr):
    if not extr:
        try:
            qname, _, _, _, _, tname, _ = line.split('\t', 6)
        except ValueError:
            return None
        return qname, 0, tname, None
    x = line.split('\t')
    try:
        rec = (x[5], int(x[11]), int(x[10]), int(x[7]), int(x[8]))
    except (IndexError, ValueError):
        return None
    return x[0], 0, x[5], rec


def _map(line, extr):
    query, found, rest = line.partition('\t')
    if not found:
        return None
    subject = rest.partition('\t')[0].rstrip()
    return query, 0, subject, None


_EXTRACT = {'sam': _sam, 'b6o': _b6o, 'paf': _paf, 'map': _map}
_SUFFIX = ('', '/1', '/2')


_CIGAR_OP = re.compile(r'([^MDIHNPSX=]*)([MDIHNPSX=])')


@lru_cache(maxsize=128)
def cigar_to_lens(cigar):
    """(aligned length, reference span) of a CIGAR string: M/=/X count for
    both, D/N only for the span; other operations are ignored and a CIGAR
    without any operation ('*') gives (0, 0) (align.py:550-583)."""
    align = extra = 0
    for num, op in _CIGAR_OP.findall(cigar):
        if op in 'M=X':
            align += int(num)
        elif op in 'DN':
            extra += int(num)
    return align, align + extra


def _grouped(lines, fmt, extr, excl):
    """The one grouping loop: yields (query, set | list) per query x mate."""
    try:
        extract = _EXTRACT[fmt]
    except KeyError:
        raise ValueError(f'Invalid format code: "{fmt}".')
    if fmt == 'sam':
        lines = _skip_sam_header(lines)
    new_pool = (lambda: ([], [], [])) if extr else \
        (lambda: (set(), set(), set()))
    this, keep, pool = None, True, new_pool()
    for line in lines:
        rec = extract(line, extr)
        if rec is None:
            continue
        qname, mate, subject, payload = rec
        if qname != this:
            if keep:
                for m in range(3):
                    if pool[m]:
                        yield this + _SUFFIX[m], pool[m]
            this, keep, pool = qname, True, new_pool()
        if not keep:
            continue
        if excl and subject in excl:
            keep = False
        elif extr:
            pool[mate].append(payload)
        else:
            pool[mate].add(subject)
    if keep and this is not None:
        for m in range(3):
            if pool[m]:
                yield this + _SUFFIX[m], pool[m]


def _skip_sam_header(lines):
    for line in lines:
        if line[0] != '@':
            yield line
            break
    yield from lines


def infer_align_format(fh):
    """Guess the format from the first line (align.py:153-223)."""
    try:
        line = next(fh)
    except StopIteration:
        raise ValueError('Alignment file is empty or unreadable.')
    if line.split()[0] in ('@HD', '@PG'):
        return 'sam', [line]
    row = line.rstrip().split('\t')
    if len(row) == 2:
        return 'map', [line]
    if len(row) >= 12:
        if all(row[i].isdigit() for i in range(3, 10)):
            return 'b6o', [line]
        if row[4] in '+-' and all(row[i].isdigit() for i in
                                  (1, 2, 3, 6, 7, 8, 9, 10, 11)):
            return 'paf', [line]
    if len(row) >= 11 and all(row[i].isdigit() for i in (1, 3, 4)):
        return 'sam', [line]
    raise ValueError('Cannot determine alignment file format.')


def iter_align(fh, fmt=None, excl=None, extr=None):
    """Iterator of (query, subjects) [extr: (query, records)]
    (align.py:118-150)."""
    if not fmt:
        fmt, head = infer_align_format(fh)
        fh = chain(iter(head), fh)
    return _grouped(iter(fh), fmt, bool(extr), excl or None)


def plain_mapper(fh, fmt=None, excl=None, n=1024):
    """Yield (queries, subject sets) in chunks of n queries
    (align.py:47-115)."""
    qryque, subque = [], []
    for query, subjects in iter_align(fh, fmt, excl):
        qryque.append(query)
        subque.append(subjects)
        if len(qryque) == n:
            yield qryque, subque
            qryque, subque = [], []
    if qryque:
        yield qryque, subque
