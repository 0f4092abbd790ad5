"""Flattening of the reference's dict hierarchy into device arrays.

The reference keeps `tree` (child -> parent, tree.py:14-25), `rankdic`
(taxon -> rank) and `root` as Python dicts and walks them per subject
(tree.find_rank tree.py:467-510, tree.find_lca :513-566).  Here the tree is
numbered once in breadth-first order (parent index < child index, the order
the LCA kernel relies on) and every per-subject walk becomes a table:

    parent[T]            int32, parent[root] == root
    anc_at_rank(r)[T]    int32, find_rank(node, r) for every node, -1 = None

`anc_at_rank` is filled level by level (vectorised), each level reusing the
answer of its parents, which is exactly the upward walk of find_rank read
top-down.
"""
import numpy as np


class FlatTree:
    def __init__(self, ids, parent, node_rank, rank_names, level_off, root):
        self.ids = ids                  # index -> identifier (list of str)
        self.parent = parent            # int32 [T]
        self.node_rank = node_rank      # int32 [T], index into rank_names, -1
        self.rank_names = rank_names    # list of str
        self.level_off = level_off      # BFS level boundaries (len = depth+1)
        self.root = root                # index of the root passed by caller
        self.index = None               # identifier -> index (built on demand)
        # from_dicts: identifier -> position in the dict it was given, and
        # that position -> index; saves a second identifier dict
        self._pos = None
        self._pos_to_index = None
        self._anc = {}

    @property
    def n_nodes(self):
        return len(self.parent)

    def node_of(self, name):
        if self.index is None:
            if self._pos is not None:
                j = self._pos.get(name)
                return -1 if j is None else int(self._pos_to_index[j])
            self.index = {x: i for i, x in enumerate(self.ids)}
        return self.index.get(name, -1)

    def nodes_of(self, names):
        """node_of for a list of names: int32 array, -1 = not in the tree."""
        if self.index is None and self._pos is not None:
            get = self._pos.get
            pos = np.fromiter((get(x, -1) for x in names), dtype=np.int64,
                              count=len(names))
            if not len(self._pos_to_index):      # an empty hierarchy
                return np.full(len(names), -1, dtype=np.int32)
            out = self._pos_to_index[np.where(pos < 0, 0, pos)].astype(np.int32)
            out[pos < 0] = -1
            return out
        if self.index is None:
            self.index = {x: i for i, x in enumerate(self.ids)}
        get = self.index.get
        return np.fromiter((get(x, -1) for x in names), dtype=np.int32,
                           count=len(names))

    def rank_id(self, rank):
        try:
            return self.rank_names.index(rank)
        except ValueError:
            return -2  # no node carries this rank: find_rank is None for all

    # -- builders ----------------------------------------------------------
    @classmethod
    def from_dicts(cls, tree, rankdic=None, root=None):
        """Number a `tree` dict breadth-first from its self-parent node(s).

        Raises KeyError if a parent is not itself a key of the tree (the
        reference fails on the same input with KeyError while walking,
        tree.py:510) and ValueError for nodes that never reach a root.
        """
        rankdic = rankdic or {}
        keys = list(tree)
        n = len(keys)
        index = dict(zip(keys, range(n)))
        try:
            par = np.fromiter(map(index.__getitem__, tree.values()),
                              dtype=np.int64, count=n)
        except KeyError as err:          # a parent that is no key (or None)
            raise KeyError(err.args[0]) from None
        me = np.arange(n, dtype=np.int64)
        is_root = par == me
        # children of every node, in dict order: a stable sort by parent
        kids = np.argsort(np.where(is_root, n, par), kind='stable')
        n_kids = np.bincount(par[~is_root], minlength=n)
        first = np.concatenate([[0], np.cumsum(n_kids)[:-1]])
        # breadth first: a level is the children of the level above, parent
        # by parent
        order = np.empty(n, dtype=np.int64)       # new index -> old index
        new_par = np.empty(n, dtype=np.int64)     # in new indices
        frontier = np.flatnonzero(is_root)
        level_off, done = [0], 0
        new_par[:len(frontier)] = np.arange(len(frontier))
        while len(frontier):
            m = len(frontier)
            order[done:done + m] = frontier
            cnt = n_kids[frontier]
            total = int(cnt.sum())
            if done + m + total > n:
                break                              # (cannot happen in a forest)
            # position of every child in `kids`
            starts = np.repeat(first[frontier], cnt)
            within = np.arange(total) - np.repeat(np.cumsum(cnt) - cnt, cnt)
            nxt = kids[starts + within]
            new_par[done + m:done + m + total] = np.repeat(
                np.arange(done, done + m), cnt)
            done += m
            level_off.append(done)
            frontier = nxt
        if done != n:
            raise ValueError('Hierarchy contains nodes that do not descend '
                             'from a root (cycle or dangling parent).')
        ids = [keys[i] for i in order.tolist()]
        rank_names, rank_index = [], {}
        node_rank = np.full(n, -1, dtype=np.int32)
        if rankdic:
            get = rankdic.get
            ranks = [get(x) for x in ids]
            for r in dict.fromkeys(ranks):         # first appearance order
                if r is not None:
                    rank_index[r] = len(rank_names)
                    rank_names.append(r)
            rank_index[None] = -1
            node_rank = np.fromiter(map(rank_index.__getitem__, ranks),
                                    dtype=np.int32, count=n)
        ft = cls(ids, new_par.astype(np.int32), node_rank, rank_names,
                 level_off, -1)
        inv = np.empty(n, dtype=np.int32)
        inv[order] = np.arange(n, dtype=np.int32)
        ft._pos, ft._pos_to_index = index, inv
        ft.n_roots = int(is_root.sum())
        ft.root = ft.node_of(root) if root is not None else -1
        return ft

    @classmethod
    def from_arrays(cls, parent, node_rank, rank_names, level_off, ids=None,
                    root=0):
        parent = np.ascontiguousarray(parent, dtype=np.int32)
        ft = cls(ids, parent, np.ascontiguousarray(node_rank, dtype=np.int32),
                 list(rank_names), list(level_off), root)
        ft.n_roots = int((parent == np.arange(len(parent))).sum())
        return ft

    # -- tables ------------------------------------------------------------
    def anc_at_rank(self, rank):
        """find_rank(node, rank) for every node (tree.py:467-510)."""
        if rank in self._anc:
            return self._anc[rank]
        rid = self.rank_id(rank)
        T = self.n_nodes
        anc = np.full(T, -1, dtype=np.int32)
        own = self.node_rank == rid
        idx = np.arange(T, dtype=np.int32)
        for lv in range(len(self.level_off) - 1):
            a, b = self.level_off[lv], self.level_off[lv + 1]
            if lv == 0:
                anc[a:b] = np.where(own[a:b], idx[a:b], -1)
            else:
                anc[a:b] = np.where(own[a:b], idx[a:b],
                                    anc[self.parent[a:b]])
        self._anc[rank] = anc
        return anc
