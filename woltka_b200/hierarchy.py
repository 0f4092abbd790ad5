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
        self.index = None               # identifier -> index (lazy for arrays)
        self._anc = {}

    @property
    def n_nodes(self):
        return len(self.parent)

    def node_of(self, name):
        if self.index is None:
            self.index = {x: i for i, x in enumerate(self.ids)}
        return self.index.get(name, -1)

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
        children = {}
        roots = []
        for node, par in tree.items():
            if par == node:
                roots.append(node)
                continue
            if par not in tree:
                raise KeyError(par)
            children.setdefault(par, []).append(node)
        ids, parent, level_off = [], [], [0]
        frontier = [(r, -1) for r in roots]
        while frontier:
            nxt = []
            for node, pidx in frontier:
                idx = len(ids)
                ids.append(node)
                parent.append(idx if pidx < 0 else pidx)
                for ch in children.get(node, ()):
                    nxt.append((ch, idx))
            level_off.append(len(ids))
            frontier = nxt
        if len(ids) != len(tree):
            raise ValueError('Hierarchy contains nodes that do not descend '
                             'from a root (cycle or dangling parent).')
        rank_names = []
        rank_index = {}
        node_rank = np.full(len(ids), -1, dtype=np.int32)
        for i, node in enumerate(ids):
            r = rankdic.get(node)
            if r is None:
                continue
            j = rank_index.get(r)
            if j is None:
                j = rank_index[r] = len(rank_names)
                rank_names.append(r)
            node_rank[i] = j
        ft = cls(ids, np.asarray(parent, dtype=np.int32), node_rank,
                 rank_names, level_off, -1)
        ft.index = {x: i for i, x in enumerate(ids)}
        ft.n_roots = len(roots)
        ft.root = ft.index.get(root, -1) if root is not None else -1
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
