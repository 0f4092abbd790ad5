"""Classification-system loaders: files -> the reference's dicts AND the flat
arrays the device uses, with a binary cache.

`build_hierarchy` has the signature and results of the reference's
(/root/reference/woltka/workflow.py:698-815): `(tree, rankdic, namedic, root)`
built from NCBI names / nodes dumps, Newick trees, lineage strings, rank
columns and plain maps (tree.py:48-388), conflicting entries raise
(util.update_dict, util.py:49-75), the root is filled in (tree.fill_root,
tree.py:302-388).  Two things are added on the way:

  * the returned `tree` carries the breadth-first numbered arrays
    (`tree.flat`, a hierarchy.FlatTree) that `classify()` uploads — they are
    made once here instead of once per run from the dict;
  * with `WOLTKA_B200_CACHE=<dir>` (or `cache_dir=`) everything is stored in
    one binary file keyed by the input files' paths, sizes and mtimes: a
    WoL / NCBI scale taxonomy (millions of nodes) loads in the time it takes
    to read that file instead of being parsed line by line again.
"""
import gc
import hashlib
import os
import pickle
import re
import sys
from os.path import basename, splitext

import numpy as np

from .hierarchy import FlatTree

__all__ = ['build_hierarchy', 'read_names', 'read_nodes', 'read_newick',
           'read_lineage', 'read_columns', 'read_map_1st', 'fill_root',
           'TreeDict']

# tree.py:29-45
_CODE2RANK = {'k': 'kingdom', 'p': 'phylum', 'c': 'class', 'o': 'order',
              'f': 'family', 'g': 'genus', 's': 'species', 't': 'strain',
              'd': 'kingdom'}
_NOTAX = {'', '0', 'unclassified', 'unassigned'}
_ZIPEXT = {'.gz', '.gzip', '.bz2', '.bzip2', '.xz', '.lz', '.lzma'}


class TreeDict(dict):
    """The `tree` dict of the reference plus its flat form (`flat`), valid
    for the (rankdic, root) it was made with and while the dict keeps its
    size."""
    flat = None
    flat_for = None     # (id(rankdic), root, len(tree))

    def flat_tree(self, rankdic, root):
        key = (id(rankdic), root, len(self))
        return self.flat if self.flat is not None and \
            self.flat_for == key else None


_LINE_END_BLANKS = re.compile(r'[^\S\n]+$', re.M)


def _dmp_rows(fh):
    """Fields of every line of a names.dmp / nodes.dmp style table, as
    `line.rstrip().replace('\t|', '').split('\t')` gives them (tree.py:66, 98)
    - done on the whole text at once: trailing blanks of every line go first,
    then the bars, then the lines and fields are split."""
    text = fh.read()
    if text.endswith('\n'):
        text = text[:-1]
    if not text:
        return []
    text = _LINE_END_BLANKS.sub('', text).replace('\t|', '')
    # (millions of small lists: the cyclic collector would walk them again
    # and again while they are made)
    was_on = gc.isenabled()
    gc.disable()
    try:
        return [line.split('\t') for line in text.split('\n')]
    finally:
        if was_on:
            gc.enable()


def read_names(fh):
    """Taxon names: NCBI names.dmp (scientific names only) or a plain map
    (tree.py:48-71)."""
    return {x[0]: x[1] for x in _dmp_rows(fh)
            if len(x) < 4 or x[3] == 'scientific name'}


def read_nodes(fh):
    """Taxon -> parent and (where given) -> rank: NCBI nodes.dmp or a plain
    table (tree.py:74-103)."""
    rows = _dmp_rows(fh)
    return ({x[0]: x[1] for x in rows},
            {x[0]: x[2] for x in rows if len(x) > 2})


def read_newick(fh):
    """Child -> parent of a Newick tree (tree.py:106-158): innermost
    parentheses are resolved one after the other; internal nodes need unique
    labels; the last parent found is the root (its own parent)."""
    nwk = ''.join(x.strip() for x in fh).rstrip(';')
    res = {}
    innermost = re.compile(r'\([^()]+\)')
    label_end = re.compile(r'[,)]')

    def label_id(label):
        return label.split(':', 1)[0].strip('"\'')

    parent = None
    while True:
        m = innermost.search(nwk)
        if m is None:
            break
        tail = nwk[m.end(0):]
        parent = label_id(label_end.split(tail, 1)[0])
        if parent == '':
            raise ValueError('Missing internal node ID.')
        for child in (label_id(x) for x in m.group(0)[1:-1].split(',')):
            if child in res:
                raise ValueError(f'Found non-unique node ID: "{child}".')
            res[child] = parent
        nwk = nwk[:m.start(0)] + tail
    res[parent] = parent      # (a file without parentheses fails here with
    return res                #  the reference too: `parent` is unbound there)


def _last_value(values):
    for x in reversed(values):
        if x is not None:
            return x
    return None


def read_columns(fh):
    """Rank-per-column table (tree.py:161-228): header names the ranks; each
    row maps its first field to its lowest taxon and every taxon to the one
    left of it; a taxon met again must agree."""
    tree, rankdic = {}, {}
    ranks = next(fh).rstrip().split('\t')[1:]
    for line in fh:
        row = line.rstrip().split('\t')
        lineage = [None if x in _NOTAX else x for x in row[1:]]
        tree[row[0]] = _last_value(lineage)
        for i, taxon in enumerate(lineage):
            if taxon is None:
                continue
            rank, parent = ranks[i], _last_value(lineage[:i])
            if taxon in tree and taxon in rankdic:
                if tree[taxon] != parent or rankdic[taxon] != rank:
                    raise ValueError(f'Conflict at taxon "{taxon}".')
            elif taxon in tree and tree[taxon] != parent:
                raise ValueError(f'Conflict at taxon "{taxon}".')
            else:
                # (reached through KeyError in the reference: either dict
                # misses the taxon)
                tree[taxon], rankdic[taxon] = parent, rank
    return tree, rankdic


_RANK_PREFIX = re.compile(r'([a-z])__.*')


def read_lineage(fh):
    """Greengenes-style lineage strings (tree.py:231-299): a taxon is named
    by its whole lineage so far; empty levels (`p__`, unclassified...) are
    skipped but stay in the names below them; `x__` prefixes give ranks."""
    tree, rankdic = {}, {}
    for line in fh:
        if line.startswith('#'):
            continue
        id_, lineage = line.rstrip().split('\t')
        parent = this = None
        for taxon in lineage.split(';'):
            taxon = taxon.strip()
            this = f'{this};{taxon}' if this else taxon
            if taxon.lower() in _NOTAX or taxon[1:] == '__':
                continue
            tree[this] = parent
            m = _RANK_PREFIX.match(taxon)
            if m and m.group(1) in _CODE2RANK:
                rankdic[this] = _CODE2RANK[m.group(1)]
            parent = this
        tree[id_] = parent
    return tree, rankdic


def read_map_1st(fh, sep='\t'):
    """(key, first value) of every line that has the separator
    (file.py:383-406)."""
    for line in fh:
        key, found, rest = line.partition(sep)
        if found:
            yield key, rest.partition(sep)[0].rstrip()


def _path2stem(fp):
    stem, ext = splitext(basename(fp))          # file.py:131-181
    if ext in _ZIPEXT:
        stem = splitext(stem)[0]
    return stem


def _stem2rank(stem):
    """Rank named by a map file: `a_to_b`, `a-2-b`, `a2b` -> b, else the stem
    (file.py:184-219)."""
    for sep in ('-', '_'):
        parts = stem.split(sep)
        if len(parts) == 3 and parts[1] in ('to', '2'):
            return parts[2]
    parts = stem.split('2')
    return parts[1] if len(parts) == 2 else stem


def _update(dic, other):
    """dict.update that refuses to change a value (util.py:25-75)."""
    if not dic:
        dic.update(other)
        return
    for key, value in other.items():
        if key in dic:
            assert dic[key] == value, f'Conflicting values found for "{key}".'
        else:
            dic[key] = value


def fill_root(tree):
    """Seal the single top node as its own parent, or hang several top nodes
    under a new node named by the first unused positive integer
    (tree.py:302-388).  Parents that are no keys become top nodes."""
    # a top node: its own parent, no parent, or a parent that is no key (the
    # reference finds them walking up from every taxon; every one of them is
    # reached, and nothing else ends a walk)
    crown = [k for k, v in tree.items() if v is None or v == k]
    toadd = set(tree.values())
    toadd.discard(None)
    toadd.difference_update(tree)
    crown.extend(toadd)
    for node in toadd:
        tree[node] = None
    if not crown:
        return None
    if len(crown) == 1:
        root = crown[0]
    else:
        i = 1
        while str(i) in tree:
            i += 1
        root = str(i)
        for x in crown:
            tree[x] = root
    tree[root] = root
    return root


def _echo(msg, nl=True):
    sys.stdout.write(msg + ('\n' if nl else ''))
    sys.stdout.flush()


def _cache_key(groups, map_rank):
    h = hashlib.sha1()
    for kind, fps in groups:
        for fp in fps:
            st = os.stat(fp)
            h.update(f'{kind}\0{os.path.abspath(fp)}\0{st.st_size}\0'
                     f'{st.st_mtime_ns}\n'.encode())
    h.update(f'map_rank={map_rank!r};v1'.encode())
    return h.hexdigest()[:24]


def build_hierarchy(names_fps=[], nodes_fps=[], newick_fps=[], lineage_fps=[],
                    columns_fps=[], map_fps=[], map_rank=None, zippers=None,
                    cache_dir=None):
    """(tree, rankdic, namedic, root) as workflow.build_hierarchy
    (workflow.py:698-815); `tree` is a TreeDict that also holds the flat
    arrays; see the module docstring for the cache."""
    from .workflow import readzip
    names_fps, nodes_fps = list(names_fps or ()), list(nodes_fps or ())
    newick_fps, lineage_fps = list(newick_fps or ()), list(lineage_fps or ())
    columns_fps, map_fps = list(columns_fps or ()), list(map_fps or ())
    groups = (('names', names_fps), ('nodes', nodes_fps),
              ('newick', newick_fps), ('lineage', lineage_fps),
              ('columns', columns_fps), ('map', map_fps))
    is_build = any(fps for _, fps in groups)
    if is_build:
        _echo('Constructing classification system...')
    if map_rank is None:
        map_rank = bool(map_fps) and not any(
            [nodes_fps, newick_fps, lineage_fps, columns_fps])

    cache_dir = cache_dir or os.environ.get('WOLTKA_B200_CACHE')
    cfp = None
    if cache_dir and is_build:
        cfp = os.path.join(cache_dir,
                           f'hierarchy-{_cache_key(groups, map_rank)}.pkl')
        loaded = _load_cache(cfp)
        if loaded is not None:
            tree, rankdic, namedic, root = loaded
            _echo(f'  Loaded from cache: {basename(cfp)}.')
            _echo('Classification system constructed.')
            _echo(f'  Total number of classification units: {len(tree)}.')
            return tree, rankdic, namedic, root

    tree, rankdic, namedic = TreeDict(), {}, {}

    def each(fps, what):
        for fp in fps:
            _echo(f'  Parsing {what} file: {basename(fp)}...', nl=False)
            with readzip(fp, zippers) as f:
                yield fp, f
            _echo(' Done.')

    for _, f in each(names_fps, 'taxon names'):
        _update(namedic, read_names(f))
    for _, f in each(nodes_fps, 'taxon nodes'):
        tree_, rankdic_ = read_nodes(f)
        _update(tree, tree_)
        _update(rankdic, rankdic_)
    for _, f in each(newick_fps, 'Newick tree'):
        _update(tree, read_newick(f))
    for _, f in each(lineage_fps, 'lineage'):
        tree_, rankdic_ = read_lineage(f)
        _update(tree, tree_)
        _update(rankdic, rankdic_)
    for _, f in each(columns_fps, 'columns'):
        tree_, rankdic_ = read_columns(f)
        _update(tree, tree_)
        _update(rankdic, rankdic_)
    if map_rank:
        _echo('  Will extract rank name from map filename.')
    for fp, f in each(map_fps, 'simple map'):
        map_ = dict(read_map_1st(f))
        _update(tree, map_)
        if map_rank:
            rank = _stem2rank(_path2stem(fp))
            _update(rankdic, {k: rank for k in set(map_.values())})

    root = fill_root(tree)
    _attach_flat(tree, rankdic, root)
    if cfp is not None:
        _store_cache(cfp, tree, rankdic, namedic, root)
    if is_build:
        _echo('Classification system constructed.')
        _echo(f'  Total number of classification units: {len(tree)}.')
    return tree, rankdic, namedic, root


def _attach_flat(tree, rankdic, root):
    """tree.flat = FlatTree of (tree, rankdic, root) when the dict is a
    proper hierarchy (what classify() would otherwise build per run)."""
    if not tree:
        return
    try:
        tree.flat = FlatTree.from_dicts(tree, rankdic, root)
        tree.flat_for = (id(rankdic), root, len(tree))
    except (KeyError, ValueError):
        tree.flat = None       # classify() reports the problem as before


def _store_cache(cfp, tree, rankdic, namedic, root):
    os.makedirs(os.path.dirname(cfp), exist_ok=True)
    ft = tree.flat
    flat = None if ft is None else dict(
        ids=ft.ids, parent=ft.parent, node_rank=ft.node_rank,
        rank_names=ft.rank_names, level_off=list(ft.level_off), root=ft.root,
        n_roots=ft.n_roots)
    tmp = f'{cfp}.{os.getpid()}.tmp'
    with open(tmp, 'wb') as f:
        pickle.dump({'tree': dict(tree), 'rankdic': rankdic,
                     'namedic': namedic, 'root': root, 'flat': flat}, f,
                    protocol=pickle.HIGHEST_PROTOCOL)
    os.replace(tmp, cfp)


def _load_cache(cfp):
    try:
        with open(cfp, 'rb') as f:
            blob = pickle.load(f)
    except (OSError, pickle.UnpicklingError, EOFError):
        return None
    tree = TreeDict(blob['tree'])
    rankdic, namedic, root = blob['rankdic'], blob['namedic'], blob['root']
    flat = blob['flat']
    if flat is not None:
        ft = FlatTree(flat['ids'], np.asarray(flat['parent'], dtype=np.int32),
                      np.asarray(flat['node_rank'], dtype=np.int32),
                      flat['rank_names'], flat['level_off'], flat['root'])
        ft.n_roots = flat['n_roots']
        tree.flat = ft
        tree.flat_for = (id(rankdic), root, len(tree))
    return tree, rankdic, namedic, root
