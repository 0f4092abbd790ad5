"""Subject coverage (`--outcov`): which ranges of every subject are covered by
at least one alignment, per sample.

Replaces /root/reference/woltka/range.py: `range_mapper` (:28-76) keeps the
mapper protocol (subjects as a dict subject -> [start1, end1, start2, ...]);
the accumulation and merging of `parse_ranges` / `merge_ranges` /
`calc_coverage` (:79-180) run on the GPU (wk_cover_add / wk_cover_merge:
radix sort of (sample, subject, start) keys, one running maximum of the ends,
a range opens where a start exceeds it); `write_coverage` (:183-254) is
restated for the output files.
"""
from os import makedirs
from os.path import join

import numpy as np

from .align import iter_align


def range_mapper(fh, fmt=None, excl=None, n=1000):
    """Like plain_mapper, but every subject comes with the ranges the query
    covers on it (range.py:28-76)."""
    qryque, subque = [], []
    for query, records in iter_align(fh, fmt, excl, True):
        ranges = {}
        for subject, _, _, start, end in records:
            ranges.setdefault(subject, []).extend((start, end))
        qryque.append(query)
        subque.append(ranges)
        if len(qryque) == n:
            yield qryque, subque
            qryque, subque = [], []
    if qryque:
        yield qryque, subque


_NAMED_FORMATS = {'bed': (0, 0), 'gff': (1, 0)}


def coverage_offsets(fmt):
    """(start offset, end offset) of an output coordinate format: 'bed'
    (default) = 0-based exclusive, 'gff' = 1-based inclusive, or `<n>e` /
    `<n>i` = n-based exclusive / inclusive (range.py:229-245, same error
    text)."""
    if fmt is None:
        return 0, 0
    if fmt.lower() in _NAMED_FORMATS:
        return _NAMED_FORMATS[fmt.lower()]
    try:
        if fmt[-1] not in 'ie':
            raise ValueError
        base = int(fmt[:-1])
    except (ValueError, IndexError):
        raise ValueError(f'Invalid coverage format: {fmt}.')
    return base, base - (fmt[-1] == 'i')


class Coverage:
    """Interval store of one classify() call on the engine `eng`."""

    def __init__(self, eng):
        self.eng = eng
        self.samples, self.subjects = {}, {}
        self._cols = ([], [], [], [])
        self._n = 0

    def add(self, sample, ranges):
        """The subject -> [start, end, ...] dict of one query of `sample`
        (range.parse_ranges, range.py:140-151)."""
        si = self.samples.setdefault(sample, len(self.samples))
        sm, sb, bg, en = self._cols
        for subject, rr in ranges.items():
            ji = self.subjects.setdefault(subject, len(self.subjects))
            for k in range(0, len(rr), 2):
                sm.append(si)
                sb.append(ji)
                bg.append(rr[k])
                en.append(rr[k + 1])
        self._n = len(sm)
        if self._n >= 1 << 20:
            self.flush()

    def flush(self):
        if self._n:
            self.eng.cover_add(*[np.asarray(c, dtype=np.int32)
                                 for c in self._cols])
            self._cols = ([], [], [], [])
            self._n = 0

    def result(self):
        """{sample: {subject: [start1, end1, ...]}} (range.calc_coverage)."""
        self.flush()
        sm, sb, bg, en = self.eng.cover_ranges()
        snames = list(self.samples)
        jnames = list(self.subjects)
        res = {name: {} for name in snames}
        for s, j, b, e in zip(sm.tolist(), sb.tolist(), bg.tolist(),
                              en.tolist()):
            res[snames[s]].setdefault(jnames[j], []).extend((b, e))
        return res

    def write(self, outdir, fmt=None):
        """One `<sample>.cov` per sample: subject, start, end per line
        (range.write_coverage, range.py:247-254)."""
        begoff, endoff = coverage_offsets(fmt)
        covers = self.result()
        makedirs(outdir, exist_ok=True)
        for sample, cover in sorted(covers.items()):
            with open(join(outdir, f'{sample}.cov'), 'w') as fh:
                for subject, ranges in sorted(cover.items()):
                    for k in range(0, len(ranges), 2):
                        print(subject, ranges[k] + begoff,
                              ranges[k + 1] + endoff, sep='\t', file=fh)
