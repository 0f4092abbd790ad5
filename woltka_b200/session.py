"""String world <-> integer world for one `classify()` call.

The reference keeps everything as Python str / set / dict
(/root/reference/woltka/workflow.py:268-335).  The GPU engine wants dense
int32 indices, so this module owns the vocabularies and the per-subject
tables, converts `(qryque, subque)` chunks to SoA columns, and turns the
device count tables back into `{rank: {sample: {feature: count}}}`.

Index spaces (see include/woltka_b200.h):
  subjects  first-seen order over the whole run (after --trim-sub)
  features  tree nodes (breadth-first order) then subjects not in the tree;
            the table has spare columns that are claimed as subjects appear
  samples   first-seen order
  strata    first-seen order of stratum labels
"""
from collections import Counter
from fractions import Fraction

import numpy as np

from ._lib import (UNITS, MAX_ENTRIES, KIND_NONE, KIND_FREE, KIND_RANK,
                   KIND_NONE_ID, F_UNIQ, F_ABOVE, F_MAJOR, F_UNASSIGNED,
                   F_SIZES)
from .hierarchy import FlatTree

_SEP = '_'
_NO_STRATUM = -2     # device copy of a subject for reads without a stratum


def _split_sample(query):
    """workflow.demultiplex (workflow.py:889-893): sample = text before the
    first '_' when something follows it, else ''."""
    left, _, right = query.partition(_SEP)
    return (left, right) if right else ('', right or left)


def _as_bytes_array(text):
    if isinstance(text, np.ndarray):
        return text
    return np.frombuffer(text, dtype=np.uint8)


class Session:
    def __init__(self, ranks, tree=None, rankdic=None, root=None, uniq=False,
                 major=None, above=False, subok=False, unasgd=False,
                 trimsub=None, engine_factory=None, device=0, rank2dir=None,
                 outzip=None, namedic=None, sizes=None, stratified=False):
        if engine_factory is None:
            from .engine import Engine
            engine_factory = Engine
        self._engine_factory, self._device = engine_factory, device
        self._matcher = None
        self.ranks = list(ranks)
        self.order = list(dict.fromkeys(self.ranks))   # unique, first-seen
        self.mult = Counter(self.ranks)
        self.trimsub = trimsub
        self.subok = subok
        self.rank2dir = rank2dir
        self.outzip = outzip if outzip != 'none' else None
        self.namedic = namedic
        # (a tree from loaders.build_hierarchy brings its flat form along)
        flat = getattr(tree, 'flat_tree', None)
        self.ft = flat and flat(rankdic, root)
        if self.ft is None and tree is not None:
            self.ft = FlatTree.from_dicts(tree, rankdic, root)
        self.T = self.ft.n_nodes if self.ft else 0

        # how each entry is assigned (workflow.py:1017-1032)
        self.kinds = []
        for rank in self.order:
            if rank is None or rank == 'none' or tree is None:
                # (feature == subject needs no table, except when a subject
                # has one device copy per stratum: --sizes with --stratify)
                self.kinds.append(KIND_NONE_ID if tree is None and
                                  not (sizes and stratified) else KIND_NONE)
            elif rank == 'free':
                self.kinds.append(KIND_FREE)
            else:
                self.kinds.append(KIND_RANK)
        need_lca = any(k == KIND_FREE for k in self.kinds) or (
            above and not major and any(k == KIND_RANK for k in self.kinds))
        if need_lca and self.ft is not None and self.ft.n_roots > 1:
            raise ValueError('Hierarchy has more than one root; run '
                             'fill_root before LCA-based assignment.')
        self.flags = ((F_UNIQ if uniq else 0) | (F_ABOVE if above else 0) |
                      (F_MAJOR if major else 0) |
                      (F_UNASSIGNED if unasgd else 0) |
                      (F_SIZES if sizes else 0))
        # --sizes: the device keeps (subject, feature) -> units, weighted here
        # (classify.counter_size, classify.py:174-213)
        self.sizes = sizes
        self.major_th = float(major) if major else 0.0

        # vocabularies
        self._sub_index = {}         # name -> subject index, built on demand
        self._sub_done = 0           # (entries of sub_name it covers)
        self.sub_node = []
        self.sub_feat = []
        # --sizes with --stratify: a subject seen in stratum t is its own
        # device subject (same tables), so that the (subject, feature) shares
        # of the device also tell the stratum (classify.counter_size_strat,
        # classify.py:252-297)
        self.sub_name = []           # subject index -> name
        self.sub_stratum = []        # subject index -> stratum index or -1
        self._pair_index = {}        # (subject index, stratum) -> index
        self.extra_names = []        # features beyond the tree
        self._extra_index = {}       # the same for extra_names
        self._extra_done = 0
        self.sample_index = {}
        self.sample_names = []
        self.file_index = 0          # set by classify() before each file
        self.sample_seen = {}        # sample -> (file index, order in file)
        self.stratum_index = {}
        self.stratum_names = []
        self.uses_strata = False
        self._tab_rows = [[] for _ in self.order]
        self._rank_tabs = [self.ft.anc_at_rank(r) if k == KIND_RANK else None
                           for r, k in zip(self.order, self.kinds)]
        self._dirty = True

        # engines: groups of at most MAX_ENTRIES entries
        self.NF_cap = self.T + 1024
        self.S_cap = 8
        self.groups = [list(range(i, min(i + MAX_ENTRIES, len(self.order))))
                       for i in range(0, len(self.order), MAX_ENTRIES)]
        self.engines = []
        for grp in self.groups:
            eng = engine_factory(device)
            if self.ft is not None:
                eng.set_tree(self.ft.parent, self.ft.root)
            eng.set_plan(np.array([self.kinds[i] for i in grp], dtype=np.int32),
                         self.flags, self.major_th, self.S_cap, self.NF_cap)
            if rank2dir is not None:
                eng.set_assign_output(True)
            self.engines.append(eng)

    def close(self):
        for eng in self.engines:
            eng.close()
        self.engines = []
        if self._matcher is not None:
            self._matcher.close()
            self._matcher = None

    # -- vocabularies ------------------------------------------------------
    @property
    def sub_index(self):
        """name -> subject index.  subjects_bulk() only appends to the lists;
        the dict catches up when somebody asks by name."""
        k = self._sub_done
        if k != len(self.sub_name):
            names, strata = self.sub_name[k:], self.sub_stratum[k:]
            if any(t != -1 for t in strata):
                # (the per-stratum copies of subject_in_stratum have no name
                # of their own)
                for i, (name, t) in enumerate(zip(names, strata), k):
                    if t == -1:
                        self._sub_index[name] = i
            else:
                self._sub_index.update(zip(names, range(k, k + len(names))))
            self._sub_done = len(self.sub_name)
        return self._sub_index

    @property
    def extra_index(self):
        k = self._extra_done
        if k != len(self.extra_names):
            self._extra_index.update(zip(
                self.extra_names[k:],
                range(self.T + k, self.T + len(self.extra_names))))
            self._extra_done = len(self.extra_names)
        return self._extra_index

    def subject(self, name, trimmed=False):
        """Index of a subject string (after trimming), interning it.
        `trimmed`: the device reader has already cut the name."""
        if self.trimsub and not trimmed:
            name = name.rsplit(self.trimsub, 1)[0]   # workflow.py:840-841
        index = self._sub_index if self._sub_done == len(self.sub_name) \
            else self.sub_index
        idx = index.get(name)
        if idx is not None:
            return idx
        idx = index[name] = len(self.sub_node)
        node = self.ft.node_of(name) if self.ft else -1
        if node >= 0:
            feat = node
        else:
            extra = self.extra_index
            feat = extra.get(name)
            if feat is None:
                feat = extra[name] = self.T + len(self.extra_names)
                self.extra_names.append(name)
                self._extra_done += 1
        self.sub_node.append(node)
        self.sub_feat.append(feat)
        self.sub_name.append(name)
        self._sub_done += 1
        self.sub_stratum.append(-1)
        for e, kind in enumerate(self.kinds):
            if kind == KIND_RANK:
                v = int(self._rank_tabs[e][node]) if node >= 0 else -1
            elif kind == KIND_FREE:
                # single-hit result (classify.py:75)
                v = feat if self.subok else (
                    int(self.ft.parent[node]) if node >= 0 else -1)
            else:
                v = feat
            self._tab_rows[e].append(v)
        self._dirty = True
        return idx

    def subjects_bulk(self, names):
        """subject() for many names at once (the gene identifiers of a
        coordinates file): int32 indices; the per-name work is a dict
        look-up, the table rows are filled by numpy."""
        if self.trimsub:
            sep = self.trimsub
            names = [x.rsplit(sep, 1)[0] for x in names]
        base = len(self.sub_node)
        if base == 0:
            # the first subjects of the run: nothing to look up, and the name
            # index is only built if somebody asks by name later
            new = list(dict.fromkeys(names))
            if len(new) == len(names):
                out = np.arange(len(names), dtype=np.int32)
            else:
                order = dict(zip(new, range(len(new))))
                out = np.fromiter(map(order.__getitem__, names),
                                  dtype=np.int32, count=len(names))
        else:
            index = self.sub_index
            new = [x for x in dict.fromkeys(names) if x not in index]
            index.update(zip(new, range(base, base + len(new))))
            self._sub_done += len(new)   # (sub_name is extended below)
            out = np.fromiter(map(index.__getitem__, names), dtype=np.int32,
                              count=len(names))
        if not new:
            return out
        if self.ft is not None:
            node = self.ft.nodes_of(new)
        else:
            node = np.full(len(new), -1, dtype=np.int32)
        feat = node.copy()
        unknown = np.flatnonzero(node < 0)
        if len(unknown):
            missing = new if len(unknown) == len(new) else \
                [new[k] for k in unknown.tolist()]
            at = self.T + len(self.extra_names)
            if not self.extra_names:
                self.extra_names.extend(missing)
                feat[unknown] = np.arange(at, at + len(missing), dtype=np.int32)
            else:
                extra = self.extra_index
                fresh = [x for x in missing if x not in extra]
                extra.update(zip(fresh, range(at, at + len(fresh))))
                self.extra_names.extend(fresh)
                self._extra_done += len(fresh)
                feat[unknown] = np.fromiter(map(extra.__getitem__, missing),
                                            dtype=np.int32, count=len(missing))
        self.sub_node.extend(node.tolist())
        self.sub_feat.extend(feat.tolist())
        self.sub_name.extend(new)
        self.sub_stratum.extend([-1] * len(new))
        known = node >= 0
        safe = np.where(known, node, 0)
        for e, kind in enumerate(self.kinds):
            if kind == KIND_RANK:
                v = np.where(known, self._rank_tabs[e][safe], -1)
            elif kind == KIND_FREE and not self.subok:
                v = np.where(known, np.asarray(self.ft.parent)[safe], -1)
            else:
                v = feat
            self._tab_rows[e].extend(v.tolist())
        self._dirty = True
        return out

    def subject_in_stratum(self, base, stratum):
        """Device subject standing for subject index `base` seen in a query of
        `stratum`: a copy of its table entries under a new index."""
        key = (base, stratum)
        idx = self._pair_index.get(key)
        if idx is None:
            idx = self._pair_index[key] = len(self.sub_node)
            current = self._sub_done == len(self.sub_name)
            self.sub_node.append(self.sub_node[base])
            self.sub_feat.append(self.sub_feat[base])
            self.sub_name.append(self.sub_name[base])
            self.sub_stratum.append(stratum)
            if current:
                self._sub_done += 1      # (a copy: not in the name index)
            for row in self._tab_rows:
                row.append(row[base])
            self._dirty = True
        return idx

    def sample(self, name):
        idx = self.sample_index.get(name)
        if idx is None:
            idx = self.sample_index[name] = len(self.sample_names)
            self.sample_names.append(name)
            self.sample_seen[name] = (self.file_index, idx)
        return idx

    def stratum(self, name):
        idx = self.stratum_index.get(name)
        if idx is None:
            idx = self.stratum_index[name] = len(self.stratum_names)
            self.stratum_names.append(name)
        return idx

    def _sync_tables(self):
        """Push grown vocabularies / tables to the device(s)."""
        nf = self.T + len(self.extra_names)
        ns = len(self.sample_names)
        if nf > self.NF_cap or ns > self.S_cap:
            while nf > self.NF_cap:
                self.NF_cap = self.NF_cap * 2
            while ns > self.S_cap:
                self.S_cap *= 2
            for eng in self.engines:
                eng.resize_counts(self.S_cap, self.NF_cap)
            self._dirty = True
        if not self._dirty:
            return
        V = len(self.sub_node)
        sub_node = np.asarray(self.sub_node, dtype=np.int32)
        for grp, eng in zip(self.groups, self.engines):
            if all(self.kinds[i] == KIND_NONE_ID for i in grp):
                eng.set_subjects(None, None, V)
                continue
            tab = np.asarray([self._tab_rows[i] for i in grp],
                             dtype=np.int32).reshape(len(grp), V)
            eng.set_subjects(tab, sub_node if self.ft is not None else None, V)
        self._dirty = False

    # -- chunks -------------------------------------------------------------
    def _chunk_columns(self, qryque, subque, demux, sample_name, samples,
                       strata_of):
        """Index columns of a chunk (everything but --sizes with --stratify):
        per-query work as list comprehensions, the subjects of the whole
        chunk interned in one call.  Same numbering as query-by-query
        interning: names get their indices in order of first appearance."""
        if demux:
            pairs = [_split_sample(x) for x in qryque]
            if samples is not None:
                keep = [p[0] in samples for p in pairs]
                if not all(keep):
                    pairs = [p for p, k in zip(pairs, keep) if k]
                    subque = [x for x, k in zip(subque, keep) if k]
            snames = [p[0] for p in pairs]
            reads = [p[1] for p in pairs]
        else:
            snames, reads = None, list(qryque)
            subque = list(subque)
        nq = len(reads)
        flat = [x for subs in subque for x in subs]
        if not flat:
            return None
        q_stratum = None
        if strata_of is not None:
            # counter_strat skips reads without a stratum but the sample
            # still gets its (possibly empty) profile (workflow.py:1058)
            if demux:
                labels = [strata_of(a).get(b) for a, b in zip(snames, reads)]
            else:
                labels = list(map(strata_of(sample_name).get, reads))
            stratum = self.stratum
            q_stratum = [stratum(x) if x is not None else -1 for x in labels]
        if demux:
            q_sample = list(map(self.sample, snames))
        else:
            q_sample = [self.sample(sample_name)] * nq
        lens = np.fromiter(map(len, subque), dtype=np.int64, count=nq)
        s = self.subjects_bulk(flat)
        q = np.repeat(np.arange(nq, dtype=np.int32), lens)
        starts = (np.cumsum(lens) - lens).tolist()
        return q, s, q_sample, q_stratum, reads, starts

    def _chunk_columns_sized_strata(self, qryque, subque, demux, sample_name,
                                    samples, strata_of):
        """--sizes with --stratify: the stratum travels with the subject."""
        q, s, q_sample, q_stratum = [], [], [], []
        reads, starts = [], []
        nq = 0
        for query, subjects in zip(qryque, subque):
            if demux:
                sname, read = _split_sample(query)
                if samples is not None and sname not in samples:
                    continue
            else:
                sname, read = sample_name, query
            label = strata_of(sname).get(read)
            stratum = self.stratum(label) if label is not None else -1
            q_sample.append(self.sample(sname))
            q_stratum.append(stratum)
            reads.append(read)
            starts.append(len(q))
            # reads without a stratum contribute nothing
            # (classify.py:283-284) — they only stay in the stream (as copies
            # whose shares are dropped) when read maps are written, which
            # list every read
            if stratum >= 0 or self.rank2dir is not None:
                for idx in {self.subject(sub) for sub in subjects}:
                    q.append(nq)
                    s.append(self.subject_in_stratum(
                        idx, stratum if stratum >= 0 else _NO_STRATUM))
            nq += 1
        if not q:
            return None
        return q, s, q_sample, None, reads, starts

    def add_chunk(self, qryque, subque, demux, sample_name, samples=None,
                  strata_of=None):
        """One `(qryque, subque)` chunk of the mapper protocol
        (workflow.py:304-335): demultiplex, intern, classify on the GPU.

        strata_of(sample_name) -> dict read -> stratum label, or None.
        Returns the number of queries in the chunk."""
        use_strata = strata_of is not None
        sized_strata = bool(self.sizes) and use_strata
        if not sized_strata:
            cols = self._chunk_columns(qryque, subque, demux, sample_name,
                                       samples, strata_of)
            if cols is None:
                return
            q, s, q_sample, q_stratum, reads, starts = cols
        else:
            cols = self._chunk_columns_sized_strata(
                qryque, subque, demux, sample_name, samples, strata_of)
            if cols is None:
                return
            q, s, q_sample, q_stratum, reads, starts = cols
        self.uses_strata = self.uses_strata or use_strata
        self._sync_tables()
        q = np.asarray(q, dtype=np.int32)
        s = np.asarray(s, dtype=np.int32)
        q_sample = np.asarray(q_sample, dtype=np.int32)
        q_stratum = np.asarray(q_stratum, dtype=np.int32) \
            if q_stratum is not None else None
        packed = None
        for eng in self.engines:
            if hasattr(eng, 'classify_packed') and self.rank2dir is None:
                # the compact wire format: one head bit per record + uint16 /
                # uint32 subjects (wk_classify_packed)
                if packed is None:
                    packed = eng.pack_columns(q, s, pinned=False)
                eng.classify_packed(packed, q_sample, 0, q_stratum)
            else:
                eng.classify_chunk(q, s, q_sample, q_stratum)
        if self.rank2dir is not None:
            starts.append(len(q))
            self._write_readmaps(len(q), reads, starts, q_sample)

    # -- SAM text parsed on the device (wk_parse_sam) ---------------------------
    def can_parse_on_device(self):
        """One engine, no read maps, a --trim-sub separator the reader can
        hold, and an engine that has the device reader (the test stand-in
        engine does not)."""
        return (len(self.engines) == 1 and self.rank2dir is None and
                len((self.trimsub or '').encode()) <= 8 and
                not getattr(self, 'device_reader_off', False) and
                hasattr(self.engines[0], 'parse_sam'))

    def configure_reader(self, exclude=None, coords=False):
        """Options of the device reader for this run (wk_parse_options):
        --trim-sub applies to the subjects of the plain path only — with
        --coords the reference cuts the gene identifiers, which `subject()`
        does when they are interned."""
        key = (bool(coords), id(exclude))
        if getattr(self, '_reader_key', None) != key:
            self.engines[0].parse_options(
                None if coords else self.trimsub, exclude, coords)
            self._reader_key = key

    def add_text_chunk_host(self, text, demux, sample_name, samples=None,
                            fmt='sam', n=1024, exclude=None):
        """The same chunk of text through the host reader (the device
        reader's tables were full).  Subjects the device named keep their
        indices; from here on the host numbers the new ones, so the device
        reader is not used again in this run."""
        from .align import plain_mapper
        self.device_reader_off = True
        lines = iter(text.decode().splitlines(True))
        nqry = 0
        for qryque, subque in plain_mapper(lines, fmt=fmt, excl=exclude, n=n):
            nqry += len(qryque)
            self.add_chunk(qryque, subque, demux, sample_name, samples)
        return nqry

    def add_text_chunk(self, text, demux, sample_name, samples=None,
                       fmt='sam'):
        """A chunk that ends where a query ends (see add_text_block);
        returns the number of queries in it."""
        return self.add_text_block(text, True, demux, sample_name, samples,
                                   fmt)[1]

    def add_text_block(self, text, final, demux, sample_name, samples=None,
                       fmt='sam'):
        """One block of alignment text (bytes or a uint8 array, best
        page-locked): lines -> queries -> indices on the GPU (align.py:258-347,
        workflow.py:844-909); the host only names the subjects and samples
        that appear for the first time.  Unless `final`, the device leaves the
        last query group and any cut line for the next block
        (wk_parse_block).  Returns (bytes consumed, queries classified)."""
        eng = self.engines[0]
        if not hasattr(self, '_dev_sub'):
            if self.sub_node:
                raise RuntimeError('device and host readers cannot be mixed '
                                   'in one run')
            self._dev_sub = 0          # subjects named so far
            self._dev_smp = []         # device sample index -> plan sample
        used, n_rec, n_qry, n_sub, n_smp = eng.parse_block(
            _as_bytes_array(text), demux, fmt, final)
        for name in eng.fetch_names(0, self._dev_sub, n_sub):
            if self.subject(name, trimmed=True) != self._dev_sub:
                raise RuntimeError('subject numbering diverged')
            self._dev_sub += 1
        if not n_rec:
            return used, 0
        if demux:
            self._name_device_samples(eng, n_smp, samples)
            self._sync_tables()
            eng.classify_parsed(np.asarray(self._dev_smp, dtype=np.int32))
        else:
            si = self.sample(sample_name)
            self._sync_tables()
            eng.classify_parsed(None, si)
        return used, n_qry

    def _name_device_samples(self, eng, n_smp, samples):
        for name in eng.fetch_names(1, len(self._dev_smp), n_smp):
            keep = samples is None or name in samples
            self._dev_smp.append(self.sample(name) if keep else -1)

    def add_text_chunk_ordinal(self, text, genes, th, demux, sample_name,
                               samples=None, fmt='sam'):
        return self.add_text_block_ordinal(text, True, genes, th, demux,
                                           sample_name, samples, fmt)[1]

    def add_text_block_ordinal(self, text, final, genes, th, demux,
                               sample_name, samples=None, fmt='sam'):
        """One block of alignment text for the coordinate matcher: the reader
        (configured with coords) leaves query / contig / beg / end / len
        columns on the device (align.py:350-406, 550-583, 807-855, 1046-1088),
        the host names the contigs that appear for the first time, and the
        matcher + classify kernels run on the resident columns
        (ordinal.py:167-335, workflow.py:304-335).

        A query is a run of adjacent lines with one name (per mate).  The
        reference also merges a name that comes back later in the same chunk
        (ordinal.py:332): a block with a name in two places raises
        WoltkaB200Error code 6 before anything is counted and the caller
        hands it to the host reader.  Returns
        (bytes consumed, queries in the block)."""
        eng = self.engines[0]
        if not hasattr(self, '_dev_contig'):
            self._dev_contig = []      # device subject index -> contig index
            self._dev_smp = []
        used, n_rec, n_qry, n_sub, n_smp = eng.parse_block(
            _as_bytes_array(text), demux, fmt, final)
        lookup = genes.contig_index.get
        for name in eng.fetch_names(0, len(self._dev_contig), n_sub):
            self._dev_contig.append(lookup(name, -1))
        if not n_rec:
            return used, 0
        genes.bind(self)
        cmap = np.asarray(self._dev_contig, dtype=np.int32)
        if demux:
            self._name_device_samples(eng, n_smp, samples)
            self._sync_tables()
            eng.ordinal_parsed(cmap, th,
                               np.asarray(self._dev_smp, dtype=np.int32))
        else:
            si = self.sample(sample_name)
            self._sync_tables()
            eng.ordinal_parsed(cmap, th, None, si)
        return used, n_qry

    def _write_readmaps(self, n_rec, reads, starts, q_sample, stops=None):
        """Append this chunk's read-to-taxon lines (file.write_readmap,
        file.py:469-500, called at workflow.py:1042-1046): one line per
        assigned query; a unique assignment prints the taxon, a list prints
        `taxon:count` sorted by count (high to low) then name."""
        from os.path import join
        from .engine import ASSIGN_UNIQ
        from .workflow import openzip
        namedic = self.namedic
        for grp, eng in zip(self.groups, self.engines):
            asg = eng.fetch_assignments(n_rec)
            for e, gi in enumerate(grp):
                rank = self.order[gi]
                lines = {}
                row = asg[e]
                for j, read in enumerate(reads):
                    vals = row[starts[j]:stops[j] if stops else starts[j + 1]]
                    vals = vals[vals >= 0]
                    if not len(vals):
                        continue
                    if vals[0] & ASSIGN_UNIQ:
                        name = self.feature_name(int(vals[0]) & ~ASSIGN_UNIQ)
                        if namedic and name in namedic:
                            name = namedic[name]
                        text = f'{read}\t{name}'
                    else:
                        counts = {}
                        for v in vals.tolist():
                            name = self.feature_name(v)
                            counts[name] = counts.get(name, 0) + 1
                        parts = [read]
                        for name, c in sorted(counts.items(),
                                              key=lambda x: (-x[1], x[0])):
                            if namedic and name in namedic:
                                name = namedic[name]
                            parts.append(f'{name}:{c}')
                        text = '\t'.join(parts)
                    lines.setdefault(int(q_sample[j]), []).append(text)
                for si, rows in lines.items():
                    fp = join(self.rank2dir[rank],
                              f'{self.sample_names[si]}.txt')
                    if self.outzip:
                        fp = f'{fp}.{self.outzip}'
                    with openzip(fp, 'at') as fh:
                        # a rank listed twice gets its lines twice: the
                        # reference runs assign_readmap per entry of `ranks`
                        # (workflow.py:333-335)
                        fh.write(('\n'.join(rows) + '\n') * self.mult[rank])

    def add_ordinal_chunk(self, genes, qnames, contigs, beg, end, length, th,
                          demux, sample_name, samples=None, strata_of=None):
        """One chunk of alignment records with coordinates: reads are matched
        to genes and classified on the GPU (ordinal.py:167-335 followed by
        workflow.py:304-335).  `qnames`/`contigs` are per-record strings."""
        use_strata = strata_of is not None
        qindex = {}
        q = np.empty(len(qnames), dtype=np.int32)
        q_sample, q_stratum, reads = [], [], []
        for i, query in enumerate(qnames):
            j = qindex.get(query)
            if j is None:
                j = qindex[query] = len(q_sample)
                if demux:
                    sname, read = _split_sample(query)
                    if samples is not None and sname not in samples:
                        sname = None
                else:
                    sname, read = sample_name, query
                reads.append(read)
                if sname is None and demux:
                    q_sample.append(-1)
                    q_stratum.append(-1)
                else:
                    stratum = 0
                    if use_strata:
                        label = strata_of(sname).get(read)
                        stratum = self.stratum(label) \
                            if label is not None else -1
                    q_sample.append(self.sample(sname))
                    q_stratum.append(stratum)
            q[i] = j
        if not len(q):
            return
        # ordinal.py:332 merges equal query names anywhere in the chunk: make
        # the records of one query contiguous
        order = None
        if np.any(np.diff(q) < 0):
            order = np.argsort(q, kind='stable')
        cidx = np.asarray([genes.contig_index.get(c, -1) for c in contigs],
                          dtype=np.int32)
        cols = [q, cidx, np.asarray(beg, dtype=np.int32),
                np.asarray(end, dtype=np.int32),
                np.asarray(length, dtype=np.int32)]
        if order is not None:
            cols = [c[order] for c in cols]
        self.uses_strata = self.uses_strata or use_strata
        genes.bind(self)
        self._sync_tables()
        q_sample = np.asarray(q_sample, dtype=np.int32)
        q_stratum = np.asarray(q_stratum, dtype=np.int32) if use_strata \
            else None
        if self.rank2dir is not None or (self.sizes and use_strata):
            pos = order if order is not None else np.arange(len(q))
            return self._ordinal_chunk_with_maps(genes, cols, th, q_sample,
                                                 q_stratum, reads, pos)
        for eng in self.engines:
            eng.ordinal_chunk(*cols, th, q_sample, q_stratum)

    def _ordinal_chunk_with_maps(self, genes, cols, th, q_sample, q_stratum,
                                 reads, pos):
        """--coords with --outmap: the matcher alone runs first (a plan-less
        engine), its (record, gene) pairs come back and go through the plain
        path, whose kernel also writes the per-record assignment column the
        read maps are made of (workflow.py:1042-1046 on the gene sets of
        ordinal.flush_chunk).  The lines come in the order the reference's
        per-contig sweep lists the reads (ordinal.reference_read_order);
        `pos` is the position of every record in the chunk as read."""
        if getattr(self, '_matcher', None) is None:
            self._matcher = self._engine_factory(self._device)
            self._matcher.ordinal_set_genes(
                genes.contig_off, genes.gbeg, genes.gend,
                np.arange(len(genes.gbeg), dtype=np.int32))
            self._matcher.ordinal_enable_pairs()
        self._matcher.ordinal_chunk(*cols, th)
        r, g = self._matcher.ordinal_pairs()      # (record, gene), sorted
        listed = None
        if self.rank2dir is not None:
            from .ordinal import reference_read_order
            listed = reference_read_order(genes, cols[0], cols[1], cols[2],
                                          cols[3], pos, r, g)
        # record -> query: the records of a query are contiguous and the query
        # indices ascend, so the pairs of a query stay contiguous
        r = cols[0][r]
        live = q_sample[r] >= 0
        r, g = r[live], g[live]
        if not len(r):
            return
        s2 = genes.subjects(self)[g]
        if self.sizes and q_stratum is not None:
            # --sizes with --stratify: the stratum travels with the subject
            if self.rank2dir is None:
                live = q_stratum[r] >= 0
                r, s2 = r[live], s2[live]
            s2 = np.fromiter((self.subject_in_stratum(
                                  a, b if b >= 0 else _NO_STRATUM)
                              for a, b in zip(s2.tolist(),
                                              q_stratum[r].tolist())),
                             dtype=np.int32, count=len(r))
            q_stratum = None
            self._sync_tables()
            if not len(r):
                return
        for eng in self.engines:
            eng.classify_chunk(r, s2, q_sample, q_stratum)
        if self.rank2dir is not None:
            listed = listed[q_sample[listed] >= 0]
            starts = np.searchsorted(r, listed, 'left')
            stops = np.searchsorted(r, listed, 'right')
            self._write_readmaps(len(r), [reads[j] for j in listed.tolist()],
                                 starts.tolist(), q_sample[listed],
                                 stops.tolist())

    def _sized_results(self, data):
        """--sizes: sum over subjects of weight x the exact share the subject
        contributed to the feature (the device's (subject, feature) table)."""
        names, strat_of = self.sub_name, self.sub_stratum
        NF1 = self.NF_cap + 1
        for grp, eng in zip(self.groups, self.engines):
            shares = {}   # (e, sample, feature, subject) -> Fraction
            e_, s_, t_, f_, u_ = eng.fetch_strata()
            for e, s, t, f, u in zip(e_.tolist(), s_.tolist(), t_.tolist(),
                                     f_.tolist(), u_.tolist()):
                shares[(e, s, f, t)] = Fraction(u, UNITS)
            ocell, ostrat, oden = eng.fetch_overflow()
            for cell, t, den in zip(ocell.tolist(), ostrat.tolist(),
                                    oden.tolist()):
                es, f = divmod(cell, NF1)
                k = (es // self.S_cap, es % self.S_cap, f, t)
                shares[k] = shares.get(k, 0) + Fraction(1, den)
            try:
                for (e, s, f, t) in sorted(shares):
                    if strat_of[t] == _NO_STRATUM:
                        continue
                    rank = self.order[grp[e]]
                    prof = data[rank][self.sample_names[s]]
                    name = self.feature_name(f)
                    if strat_of[t] >= 0:
                        name = (self.stratum_names[strat_of[t]], name)
                    prof[name] = prof.get(name, 0) + \
                        self.sizes[names[t]] * float(shares[(e, s, f, t)]) * \
                        self.mult[rank]
            except KeyError:
                raise ValueError('One or more subjects are not found in the '
                                 'size map.')
        return data

    # -- results --------------------------------------------------------------
    def feature_name(self, f):
        if f == self.NF_cap:
            return 'Unassigned'
        if f < self.T:
            return self.ft.ids[f]
        return self.extra_names[f - self.T]

    def _dense_results(self, data, grp, units, final=False):
        """Profiles of one engine from its units table [E, S, NF+1]; `final`:
        a count that is not whole comes as the correctly rounded double
        straight away (x / UNITS in float64 is exactly that while x < 2^53)
        instead of as a Fraction."""
        named = (list(self.ft.ids) if self.ft is not None else []) + \
            self.extra_names
        n_named = len(named)
        for e in range(units.shape[0]):
            rank = self.order[grp[e]]
            mult = self.mult[rank]
            for si in range(min(units.shape[1], len(self.sample_names))):
                row = units[e, si]
                f = np.flatnonzero(row)
                if not len(f):
                    continue
                v = row[f].astype(np.int64)
                whole = v % UNITS == 0
                # (feature NF = 'Unassigned' lies past the named ones)
                keys = [named[i] if i < n_named else self.feature_name(i)
                        for i in f.tolist()]
                vals = (v // UNITS * mult).tolist()
                part = np.flatnonzero(~whole)
                if len(part):
                    x = v[part] * mult
                    if final and int(x.max()) < (1 << 53):
                        shares = (x / UNITS).tolist()
                    else:
                        shares = [Fraction(y, UNITS) for y in x.tolist()]
                    for k, y in zip(part.tolist(), shares):
                        vals[k] = y
                data[rank][self.sample_names[si]].update(zip(keys, vals))

    def results(self):
        """{rank: {sample: {feature | (stratum, feature): count}}} with exact
        counts: int when integral, else the correctly rounded double."""
        return finalize(self.exact_results(final=True))

    def exact_results(self, final=False):
        """The same dict with the counts as exact Fractions (floats with
        --sizes): what `distributed.merge_profiles` sums across ranks."""
        data = {rank: {} for rank in self.order}
        for rank in self.order:
            for sname in self.sample_names:
                data[rank][sname] = {}
        NF1 = self.NF_cap + 1
        if self.sizes:
            return self._sized_results(data)
        for grp, eng in zip(self.groups, self.engines):
            cells = {}    # (e, sample, stratum | None, feature) -> Fraction
            ocell, ostrat, oden = eng.fetch_overflow()
            if self.uses_strata:
                e_, s_, t_, f_, u_ = eng.fetch_strata()
                for e, s, t, f, u in zip(e_.tolist(), s_.tolist(), t_.tolist(),
                                         f_.tolist(), u_.tolist()):
                    cells[(e, s, t, f)] = Fraction(u, UNITS)
            elif len(ocell) == 0:
                # no stratified cells, no overflow shares: whole profiles at
                # once (a whole count stays an int - also an exact rational)
                self._dense_results(data, grp, eng.fetch_counts(), final)
                continue
            else:
                units = eng.fetch_counts()
                for e, s, f in zip(*np.nonzero(units)):
                    cells[(int(e), int(s), None, int(f))] = Fraction(
                        int(units[e, s, f]), UNITS)
            for cell, t, den in zip(ocell.tolist(), ostrat.tolist(),
                                    oden.tolist()):
                es, f = divmod(cell, NF1)
                k = (es // self.S_cap, es % self.S_cap,
                     t if self.uses_strata else None, f)
                cells[k] = cells.get(k, 0) + Fraction(1, den)
            for (e, s, t, f), v in cells.items():
                rank = self.order[grp[e]]
                name = self.feature_name(f)
                if t is not None:
                    name = (self.stratum_names[t], name)
                data[rank][self.sample_names[s]][name] = v * self.mult[rank]
        return data


def finalize(data):
    """Exact cells -> what the reference holds: int where the count is whole,
    else the correctly rounded double (in place; returns `data`)."""
    for samples in data.values():
        for prof in samples.values():
            for key, v in [kv for kv in prof.items() if type(kv[1]) is Fraction]:
                prof[key] = int(v) if v.denominator == 1 else float(v)
    return data
