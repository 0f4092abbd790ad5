"""Host side of the coordinate-matching ("ordinal") path.

Keeps the reference's seams (/root/reference/woltka/ordinal.py):
`load_gene_coords(fh, sort)` returns the same `(coords, idmap, isdup)` triple
with the same int64 endpoint encoding (bits 0-21 gene index, bit 22 is-gene,
bit 23 is-end, bits 24+ coordinate; ordinal.py:464-465), and
`ordinal_mapper(fh, coords, idmap, fmt, excl, n, th, prefix)` follows the
mapper protocol, so `build_mapper()`'s partial exposes the same keywords
(`coords`, `idmap`, `prefix`, `th`) the reference's callers read
(ordinal.py:828-830, tests/test_workflow.py:293-312).

What differs is who matches reads to genes: the reference sorts a merged
endpoint queue per contig and sweeps it (ordinal.py:314-332); here the gene
table is flattened once (`GeneIndex`) and every chunk of reads is matched on
the GPU by wk_ordinal_chunk (closed-form overlap predicate, ordinal.py:644-646).
"""
import numpy as np

from .align import iter_align

__all__ = ['load_gene_coords', 'load_gene_coords_cached', 'ordinal_mapper',
           'GeneIndex', 'iter_records',
           'reference_read_order']

_IDX = (1 << 22) - 1


def load_gene_coords(fh, sort=False):
    """Read a gene coordinates file (ordinal.py:338-430).

    Lines: `>contig` or `#contig` start a contig (`>>` / `##` lines are
    genome headers and ignored), `gene<TAB>beg<TAB>end` adds a gene.
    Returns (coords: contig -> int64 endpoint codes, idmap: contig -> gene ids,
    isdup: whether a gene id occurs twice).
    """
    raw, idmap = {}, {}
    cur = None
    seen, isdup = set(), False
    for line in fh:
        c0 = line[0]
        if c0 in '>#':
            if line[1] != c0:
                name = line[1:].strip()
                cur = raw[name] = []
                idmap[name] = []
                cur_ids = idmap[name]
            continue
        try:
            gene, beg, end = line.rstrip().split('\t')
        except ValueError:
            raise ValueError(
                f'Cannot extract coordinates from line: "{line}".')
        cur.append(beg)
        cur.append(end)
        cur_ids.append(gene)
        if not isdup:
            if gene in seen:
                isdup = True
            else:
                seen.add(gene)
    if cur is None:
        raise ValueError('No coordinate was read from file.')
    coords = {}
    for name, flat in raw.items():
        try:
            arr = np.asarray(flat, dtype=np.int64)
        except ValueError:
            raise ValueError('Invalid coordinate(s) found.')
        a, b = arr[0::2], arr[1::2]
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        idx = np.arange(len(lo), dtype=np.int64)
        codes = np.empty(2 * len(lo), dtype=np.int64)
        codes[0::2] = ((lo - 1) << 24) + (1 << 22) + idx   # gene start
        codes[1::2] = (hi << 24) + (3 << 22) + idx         # gene end
        if sort:
            codes.sort(kind='stable')
        coords[name] = codes
    return coords, idmap, isdup


def load_gene_coords_cached(fp, opener, cache_dir):
    """load_gene_coords(sort=True) of file `fp` through a binary cache
    (the text parse is the one-off cost that grows with
    the database, ordinal.py:338-430).  The cache file is keyed by the path,
    size and mtime of `fp`; anything unreadable or stale is rebuilt."""
    import hashlib
    import os
    st = os.stat(fp)
    key = hashlib.sha1(f'{os.path.abspath(fp)}|{st.st_size}|{st.st_mtime_ns}'
                       .encode()).hexdigest()[:20]
    cfp = os.path.join(cache_dir, f'coords-{key}.npz')
    try:
        with np.load(cfp, allow_pickle=False) as z:
            names = str(z['names']).split('\n')
            off = z['off']
            codes = z['codes']
            ids = str(z['ids']).split('\n')
            isdup = bool(z['isdup'])
        if len(names) + 1 != len(off) or len(ids) * 2 != len(codes):
            raise ValueError('inconsistent cache')
        coords = {n: codes[2 * off[i]:2 * off[i + 1]]
                  for i, n in enumerate(names)}
        idmap = {n: ids[off[i]:off[i + 1]] for i, n in enumerate(names)}
        return coords, idmap, isdup
    except (OSError, KeyError, ValueError):
        pass
    with opener(fp) as fh:
        coords, idmap, isdup = load_gene_coords(fh, sort=True)
    names = list(coords)
    if any('\n' in x for x in names) or \
            any('\n' in g for n in names for g in idmap[n]):
        return coords, idmap, isdup
    off = np.cumsum([0] + [len(idmap[n]) for n in names]).astype(np.int64)
    try:
        os.makedirs(cache_dir, exist_ok=True)
        tmp = f'{cfp}.{os.getpid()}.tmp.npz'
        np.savez(tmp, names=np.asarray('\n'.join(names)), off=off,
                 codes=np.concatenate([coords[n] for n in names])
                 if names else np.zeros(0, dtype=np.int64),
                 ids=np.asarray('\n'.join(g for n in names
                                          for g in idmap[n])),
                 isdup=np.asarray(isdup))
        os.replace(tmp, cfp)
    except OSError:
        pass
    return coords, idmap, isdup


class GeneIndex:
    """Flat, device-ready form of `(coords, idmap, prefix)`: contigs in dict
    order, genes of a contig sorted by start, gbeg = lo-1, gend = hi."""

    def __init__(self, coords, idmap, prefix=False):
        self.contig_names = list(coords)
        self.contig_index = {c: i for i, c in enumerate(self.contig_names)}
        offs, begs, ends, gids, local = [0], [], [], [], []
        for name in self.contig_names:
            codes = np.asarray(coords[name], dtype=np.int64)
            idx = codes & _IDX
            isend = (codes >> 23) & 1
            pos = codes >> 24
            n = len(codes) // 2
            gb = np.empty(n, dtype=np.int64)
            ge = np.empty(n, dtype=np.int64)
            gb[idx[isend == 0]] = pos[isend == 0]
            ge[idx[isend == 1]] = pos[isend == 1]
            order = np.argsort(gb, kind='stable')
            begs.append(gb[order])
            ends.append(ge[order])
            local.append(order)
            ids = idmap[name]
            pfx = name + '_' if prefix else ''
            gids.extend(pfx + ids[i] for i in order)
            offs.append(offs[-1] + n)
        self.contig_off = np.asarray(offs, dtype=np.int64)
        cat = (lambda xs: np.concatenate(xs) if xs else
               np.zeros(0, dtype=np.int64))
        gb, ge = cat(begs), cat(ends)
        if len(gb) and (gb.min() < -(1 << 31) or ge.max() >= (1 << 31)):
            raise ValueError('Gene coordinates exceed the 32-bit range.')
        self.gbeg = gb.astype(np.int32)
        self.gend = ge.astype(np.int32)
        self.gene_ids = gids
        # index of every gene within its contig as the file listed it (the
        # index the reference packs into its endpoint codes)
        self.local_idx = cat(local).astype(np.int64)
        self._bound = None

    def bind(self, session):
        """Intern the gene identifiers as subjects of `session` and upload the
        table to its engine(s) — once per session."""
        if self._bound is session:
            return
        subj = self.subjects(session)
        for eng in session.engines:
            eng.ordinal_set_genes(self.contig_off, self.gbeg, self.gend, subj)
        self._bound = session

    def subjects(self, session):
        """Subject index of every gene in `session` (interning the gene
        identifiers on first use)."""
        if getattr(self, '_subj_of', None) is not session:
            bulk = getattr(session, 'subjects_bulk', None)
            self._subj = bulk(self.gene_ids) if bulk else np.fromiter(
                (session.subject(g) for g in self.gene_ids), dtype=np.int32,
                count=len(self.gene_ids))
            self._subj_of = session
        return self._subj


def reference_read_order(genes, q, cidx, beg, end, pos, pair_rec, pair_gene):
    """Query indices in the order the reference's result dict of one chunk
    lists them (ordinal.flush_chunk, ordinal.py:296-332): contigs in the order
    they first appear among the chunk's records; inside a contig with more than
    five records the matches come out of the endpoint sweep — a pair is
    emitted at the earlier of its two END events, several reads at one gene
    end in the order of their starts (match_read_gene, ordinal.py:476-582) —
    and with five or fewer, read by read (match_read_gene_quart, :650-811); a
    query is listed where its first match is.

    q, cidx, beg, end: per record (any order); pos: position of the record in
    the chunk as read; pair_rec / pair_gene: the matches (record, gene)."""
    q, cidx = np.asarray(q), np.asarray(cidx)
    beg, end = np.asarray(beg, dtype=np.int64), np.asarray(end, dtype=np.int64)
    pos = np.asarray(pos, dtype=np.int64)
    pair_rec, pair_gene = np.asarray(pair_rec), np.asarray(pair_gene)
    if not len(pair_rec):
        return np.zeros(0, dtype=np.int64)
    # contigs by first appearance, records per contig
    by_pos = np.argsort(pos, kind='stable')
    seen, first = np.unique(cidx[by_pos], return_index=True)
    corder = np.full(int(cidx.max()) + 2, 0, dtype=np.int64)
    corder[seen] = np.argsort(np.argsort(first))
    count = np.bincount(cidx[cidx >= 0], minlength=len(corder))
    # first emission of every matched record
    r, g = pair_rec, pair_gene
    re_code = (end[r] << 24) + (1 << 23) + pos[r]
    ge_code = (genes.gend[g].astype(np.int64) << 24) + (3 << 22) + \
        genes.local_idx[g]
    ev = np.minimum(re_code, ge_code)
    sweep = count[cidx[r]] > 5
    k1 = np.where(sweep, ev, pos[r])
    k2 = np.where(sweep, (beg[r] << 24) + pos[r], 0)
    k0 = corder[cidx[r]]
    # smallest key per query
    order = np.lexsort((k2, k1, k0))
    qs = q[r][order]
    _, firsts = np.unique(qs, return_index=True)
    return qs[np.sort(firsts)]


def iter_records(fh, fmt=None, excl=None, n=2**20):
    """Chunks of alignment records with coordinates:
    (qnames, contigs, beg, end, length) lists, at most n records per chunk, a
    query never split across chunks, zero-length hits dropped
    (ordinal.py:219-240)."""
    qn, cn, bg, en, ln = [], [], [], [], []
    for query, records in iter_align(fh, fmt, excl, True):
        if len(qn) + len(records) > n and qn:
            yield qn, cn, bg, en, ln
            qn, cn, bg, en, ln = [], [], [], [], []
        for subject, _, length, beg, end in records:
            if length:
                qn.append(query)
                cn.append(subject)
                bg.append(beg)
                en.append(end)
                ln.append(length)
    yield qn, cn, bg, en, ln


def ordinal_mapper(fh, coords, idmap, fmt=None, excl=None, n=2**20, th=0.8,
                   prefix=False, _engine_factory=None):
    """Mapper protocol over the GPU matcher: yields (queries, gene-id sets)
    per chunk, reads without a gene omitted (ordinal.py:167-240).

    `classify()` does not go through this generator — it feeds the record
    chunks to the fused match+classify device path — but other callers of the
    mapper protocol get the reference's behaviour from it.
    """
    if _engine_factory is None:
        from .engine import Engine
        _engine_factory = Engine
    genes = GeneIndex(coords, idmap, prefix)
    eng = _engine_factory(0)
    try:
        eng.ordinal_set_genes(genes.contig_off, genes.gbeg, genes.gend,
                              np.arange(len(genes.gbeg), dtype=np.int32))
        eng.ordinal_enable_pairs()
        for qn, cn, bg, en, ln in iter_records(fh, fmt, excl, n):
            res = {}
            if qn:
                cidx = np.asarray([genes.contig_index.get(c, -1) for c in cn],
                                  dtype=np.int32)
                eng.ordinal_chunk(np.arange(len(qn), dtype=np.int32), cidx,
                                  bg, en, ln, th)
                r, g = eng.ordinal_pairs()
                for ri, gi in zip(r.tolist(), g.tolist()):
                    res.setdefault(qn[ri], set()).add(genes.gene_ids[gi])
            yield res.keys(), res.values()
    finally:
        eng.close()
