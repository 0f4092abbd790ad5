"""An Engine look-alike backed by the CPU oracle (oracle/woltka_oracle.c).

TEST INFRASTRUCTURE: lets the `-m "not gpu"` suite drive the whole host layer
(woltka_b200.workflow / session / align / ordinal) end to end without a GPU,
and pins the oracle against golden outputs of the real reference.  It is
injected through the private `_engine_factory` argument; product code never
imports it.

For RANK entries the oracle does NOT use the host-built lookup tables: it
walks the tree itself (tree.find_rank) from the node ranks, so the host's
table construction is checked too.
"""
import numpy as np

from oracle import oracle as O
from woltka_b200._lib import (KIND_NONE, KIND_FREE, KIND_RANK, KIND_NONE_ID,
                              UNITS, F_UNIQ, F_ABOVE, F_MAJOR, F_SIZES,
                              F_UNASSIGNED)
from woltka_b200.hierarchy import FlatTree


def make_factory(tree=None, rankdic=None, root=None, ranks=None, subok=False,
                 log=None):
    """Factory with the string-level context the oracle needs."""
    ft = FlatTree.from_dicts(tree, rankdic, root) if tree is not None else None
    order = list(dict.fromkeys(ranks or []))
    state = {'next': 0}

    def factory(device=0):
        eng = OracleEngine(ft, subok)
        eng._all_ranks = order
        eng._state = state
        if log is not None:
            log.append(eng)
        return eng

    factory.state = state
    return factory


class OracleEngine:
    def __init__(self, ft=None, subok=False):
        self.ft = ft
        self.subok = subok
        self.parent = None
        self.root = -1
        self.units = None
        self.overflow = []
        self.strata = {}
        self.genes = None
        self.pairs = (np.zeros(0, np.int32), np.zeros(0, np.int32))
        self.closed = False
        self.n_launch = 0

    # -- plumbing ----------------------------------------------------------
    def close(self):
        self.closed = True

    def set_stream(self, s):
        pass

    def launch_count(self):
        return 0

    # -- model -------------------------------------------------------------
    def set_tree(self, parent, root):
        self.parent = np.ascontiguousarray(parent, dtype=np.int32)
        self.root = -1 if root is None else int(root)

    def set_plan(self, kinds, flags=0, major_th=0.0, n_samples=1,
                 n_features=0):
        self.kinds = np.ascontiguousarray(kinds, dtype=np.int32)
        self.flags, self.major_th = flags, major_th
        self.E, self.S, self.NF = len(kinds), n_samples, n_features
        self.units = np.zeros((self.E, self.S, self.NF + 1), dtype=np.int64)
        # rank ids wanted by the RANK entries of this group
        self.trk = np.zeros(self.E, dtype=np.int32)
        if self.ft is not None and getattr(self, '_all_ranks', None):
            first = self._state['next']
            self._state['next'] += self.E
            names = self._all_ranks[first:first + self.E]
            for e, (k, r) in enumerate(zip(self.kinds, names)):
                if k == KIND_RANK:
                    self.trk[e] = self.ft.rank_id(r)

    def resize_counts(self, n_samples, n_features):
        new = np.zeros((self.E, n_samples, n_features + 1), dtype=np.int64)
        new[:, :self.S, :self.NF] = self.units[:, :, :self.NF]
        new[:, :self.S, n_features] = self.units[:, :, self.NF]
        self.units, self.S, self.NF = new, n_samples, n_features

    def set_subjects(self, tab, sub_node=None, n_subjects=None):
        self.tab = None if tab is None else np.ascontiguousarray(
            tab, dtype=np.int32).reshape(self.E, -1)
        self.sub_node = None if sub_node is None else np.ascontiguousarray(
            sub_node, dtype=np.int32)
        self.V = n_subjects if n_subjects is not None else (
            self.tab.shape[1] if self.tab is not None else len(self.sub_node))

    def _sub_feat(self):
        if self.tab is not None:
            for e, k in enumerate(self.kinds):
                if k == KIND_NONE or (k == KIND_FREE and self.subok):
                    return self.tab[e]
        return np.zeros(self.V, dtype=np.int32)

    # -- classify ----------------------------------------------------------
    def _one(self, subs, flags):
        """Units [E, NF+1] and overflow of ONE query with subjects `subs`."""
        sub_node = self.sub_node if self.sub_node is not None else \
            np.full(self.V, -1, dtype=np.int32)
        units, ovf, _ = O.classify(
            np.zeros(len(subs), dtype=np.int32),
            np.asarray(subs, dtype=np.int32), parent=self.parent,
            node_rank=None if self.ft is None else self.ft.node_rank,
            root=self.root, sub_node=sub_node, sub_feat=self._sub_feat(),
            kinds=self.kinds, target_rank=self.trk, flags=flags,
            major_th=self.major_th, subok=self.subok, n_samples=1,
            n_features=self.NF, sample=0)
        return units[:, 0, :], ovf

    def _query_results(self, subs, flags):
        """Per entry the assignment of ONE query with distinct subjects
        `subs`: ('unique', feature) | ('list', [(subject, taxon), ...]) | None
        — the str / list / None of classify.assign_* — from single-query calls
        of the oracle."""
        plain = not (flags & (F_UNIQ | F_ABOVE | F_MAJOR))
        own = self.__dict__.setdefault('_own_feature', {})
        units, _ = self._one(subs, flags)
        out = []
        for e, kind in enumerate(self.kinds):
            listed = None
            if (kind == KIND_RANK and plain) or \
                    (kind in (KIND_NONE, KIND_NONE_ID) and
                     not flags & F_UNIQ):
                # taxon of every subject alone (find_rank / the subject)
                taxa = []
                for sub in subs:
                    if (e, sub) not in own:
                        u1, _ = self._one([sub], flags & ~F_UNASSIGNED)
                        nz = np.flatnonzero(u1[e])
                        own[(e, sub)] = int(nz[0]) if len(nz) else None
                    taxa.append(own[(e, sub)])
                if len(set(taxa)) > 1 or \
                        (kind != KIND_RANK and len(subs) > 1):
                    listed = [(sub, t) for sub, t in zip(subs, taxa)
                              if t is not None]
            if listed is not None:
                out.append(('list', listed) if listed else None)
                continue
            nz = np.flatnonzero(units[e])
            if not len(nz):
                out.append(None)
                continue
            assert len(nz) == 1 and units[e, nz[0]] == UNITS
            out.append(('unique', int(nz[0])))
        return out

    @staticmethod
    def _queries(qidx):
        q = np.asarray(qidx)
        cuts = np.flatnonzero(np.diff(q)) + 1
        return zip(np.r_[0, cuts].tolist(), np.r_[cuts, len(q)].tolist())

    def _classify_sized(self, qidx, sidx, q_sample, sample):
        """WK_F_SIZES: exact (subject, feature) shares per query
        (classify.counter_size, classify.py:174-213): a uniquely assigned
        query gives 1/k of its unit to each of its k subjects, a list gives
        1/k' to every listed subject under its own taxon.  Results in the
        strata table with the subject in the stratum's place, like the
        kernels."""
        flags = self.flags & ~F_SIZES
        un = (lambda f: None if f == self.NF else f)
        for a, b in self._queries(qidx):
            qi = int(qidx[a])
            samp = int(q_sample[qi]) if q_sample is not None else sample
            if samp < 0 or samp >= self.S:
                continue
            subs = list(dict.fromkeys(np.asarray(sidx[a:b]).tolist()))
            for e, res in enumerate(self._query_results(subs, flags)):
                if res is None:
                    continue
                if res[0] == 'unique':
                    pairs, k = [(sub, res[1]) for sub in subs], len(subs)
                else:
                    pairs, k = res[1], len(res[1])
                for sub, f in pairs:
                    if UNITS % k == 0:
                        key = (e, samp, sub, un(f))
                        self.strata[key] = self.strata.get(key, 0) + UNITS // k
                    else:
                        self.overflow.append((e, samp, sub, un(f), k))

    # -- read maps: the per-record assignment column of classify_kernel --------
    def set_assign_output(self, enable=True):
        self.want_assign = bool(enable)

    def _record_assignments(self, qidx, sidx):
        from woltka_b200.engine import ASSIGN_UNIQ
        sidx = np.asarray(sidx)
        asg = np.full((self.E, len(sidx)), -1, dtype=np.int32)
        for a, b in self._queries(qidx):
            first = {}
            for i in range(a, b):
                first.setdefault(int(sidx[i]), i)
            for e, res in enumerate(self._query_results(
                    list(first), self.flags & ~F_SIZES)):
                if res is None:
                    continue
                if res[0] == 'unique':
                    asg[e, a] = res[1] | ASSIGN_UNIQ
                else:
                    for sub, t in res[1]:
                        asg[e, first[sub]] = t
        self._assign = asg

    def fetch_assignments(self, n_rec):
        assert self._assign.shape[1] == n_rec
        return self._assign

    def classify_chunk(self, qidx, sidx, q_sample=None, q_stratum=None,
                       sample=0):
        if getattr(self, 'want_assign', False):
            self._record_assignments(qidx, sidx)
        if self.flags & F_SIZES:
            return self._classify_sized(qidx, sidx, q_sample, sample)
        sub_node = self.sub_node if self.sub_node is not None else \
            np.full(self.V, -1, dtype=np.int32)
        units, ovf, strata = O.classify(
            qidx, sidx, parent=self.parent,
            node_rank=None if self.ft is None else self.ft.node_rank,
            root=self.root, sub_node=sub_node, sub_feat=self._sub_feat(),
            kinds=self.kinds, target_rank=self.trk, flags=self.flags,
            major_th=self.major_th, subok=self.subok, n_samples=self.S,
            n_features=self.NF, q_sample=q_sample, q_stratum=q_stratum,
            sample=sample)
        self.units += units
        # keep features capacity-independent: None marks 'Unassigned'
        un = (lambda f: None if f == self.NF else f)
        self.overflow.extend((e, s, t, un(f), d) for e, s, t, f, d in ovf)
        for (e, s, t, f), v in strata.items():
            k = (e, s, t, un(f))
            self.strata[k] = self.strata.get(k, 0) + v

    # -- ordinal -----------------------------------------------------------
    def ordinal_set_genes(self, contig_off, gbeg, gend, gene_subject):
        self.genes = (np.ascontiguousarray(contig_off, dtype=np.int64),
                      np.ascontiguousarray(gbeg, dtype=np.int32),
                      np.ascontiguousarray(gend, dtype=np.int32),
                      np.ascontiguousarray(gene_subject, dtype=np.int32))

    def ordinal_enable_pairs(self):
        return len(self.pairs[0])

    def ordinal_pairs(self):
        return self.pairs

    def ordinal_chunk(self, qidx, contig, beg, end, length, th, q_sample=None,
                      q_stratum=None, sample=0):
        coff, gb, ge, gsub = self.genes
        qidx = np.ascontiguousarray(qidx, dtype=np.int32)
        r, g = O.ordinal_match(contig, beg, end, length, th, coff, gb, ge)
        self.pairs = (r, g)
        if self.units is None or not len(r):
            return
        self.classify_chunk(qidx[r], gsub[g], q_sample, q_stratum, sample)

    # -- subject coverage ----------------------------------------------------
    def cover_add(self, sample, subject, beg, end):
        store = self.__dict__.setdefault('cover', {})
        for sm, sb, b, e in zip(*[np.asarray(x).tolist()
                                  for x in (sample, subject, beg, end)]):
            store.setdefault((sm, sb), []).extend((b, e))

    def cover_ranges(self):
        from oracle import pyport
        rows = []
        for (sm, sb), ranges in sorted(self.__dict__.get('cover', {}).items()):
            merged = pyport.merge_ranges(ranges)
            rows.extend((sm, sb, merged[k], merged[k + 1])
                        for k in range(0, len(merged), 2))
        cols = np.asarray(rows, dtype=np.int32).reshape(-1, 4)
        return [np.ascontiguousarray(cols[:, k]) for k in range(4)]

    # -- results -----------------------------------------------------------
    def fetch_counts(self):
        return self.units.copy()

    def fetch_overflow(self):
        NF1 = self.NF + 1
        cell, strat, den = [], [], []
        for e, s, t, f, d in self.overflow:
            f = self.NF if f is None else f
            cell.append((e * self.S + s) * NF1 + f)
            strat.append(t)
            den.append(d)
        return (np.asarray(cell, np.int64), np.asarray(strat, np.int32),
                np.asarray(den, np.int32))

    def fetch_strata(self):
        e, s, t, f, u = [], [], [], [], []
        for (e_, s_, t_, f_), val in self.strata.items():
            e.append(e_)
            s.append(s_)
            t.append(t_)
            f.append(self.NF if f_ is None else f_)
            u.append(val)
        return (np.asarray(e, np.int32), np.asarray(s, np.int32),
                np.asarray(t, np.int32), np.asarray(f, np.int64),
                np.asarray(u, np.int64))
