"""Classification-system loaders (woltka_b200/loaders.py)
against the reference's `build_hierarchy` (workflow.py:698-815): the dicts, the
root, the messages and the errors are the reference's — from the committed
golden results (tests/golden/hierarchy.json, made by
tests/golden/make_hierarchy_golden.py from the real reference) and, where the
reference is on the box, from the reference itself on its own test data; the
flat arrays that ride along equal the ones classify() builds from the dicts;
the binary cache returns the same thing."""
import io
import json
import os
from contextlib import redirect_stdout

import numpy as np
import pytest

from tests.golden.make_hierarchy_golden import CASES, write_inputs, run_case
from woltka_b200.hierarchy import FlatTree
from woltka_b200.loaders import build_hierarchy, TreeDict, fill_root

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, 'golden', 'hierarchy.json')) as f:
    GOLDEN = json.load(f)


@pytest.fixture(scope='module')
def inputs(tmp_path_factory):
    d = tmp_path_factory.mktemp('hierarchy')
    write_inputs(str(d))
    return str(d)


@pytest.mark.parametrize('name', sorted(CASES))
def test_golden_hierarchies(inputs, name):
    got = run_case(build_hierarchy, inputs, CASES[name])
    assert json.loads(json.dumps(got)) == GOLDEN[name]


def same_flat(a, b):
    return (a.ids == b.ids and np.array_equal(a.parent, b.parent) and
            np.array_equal(a.node_rank, b.node_rank) and
            a.rank_names == b.rank_names and
            list(a.level_off) == list(b.level_off) and a.root == b.root and
            a.n_roots == b.n_roots)


@pytest.mark.parametrize('name', ['nodes_ncbi', 'newick', 'lineage', 'columns',
                                  'map_rank_from_stem', 'nodes_tworoots'])
def test_flat_arrays_ride_along_and_cache_round_trip(inputs, tmp_path, name):
    kw = {k: [os.path.join(inputs, x) for x in v] if isinstance(v, list) else v
          for k, v in CASES[name].items()}
    with redirect_stdout(io.StringIO()):
        tree, rankdic, namedic, root = build_hierarchy(**kw)
    assert isinstance(tree, TreeDict)
    flat = tree.flat_tree(rankdic, root)
    assert flat is not None
    assert same_flat(flat, FlatTree.from_dicts(dict(tree), rankdic, root))
    # not for another rank dict / root, nor after the dict has grown
    assert tree.flat_tree(dict(rankdic), root) is None
    assert tree.flat_tree(rankdic, 'other') is None
    # the cache: made on the first call, read on the second
    cache = str(tmp_path / 'cache')
    outs = []
    for _ in range(2):
        buf = io.StringIO()
        with redirect_stdout(buf):
            outs.append(build_hierarchy(cache_dir=cache, **kw))
        said = buf.getvalue()
    assert 'Loaded from cache' in said and len(os.listdir(cache)) == 1
    for t2, r2, n2, root2 in outs:
        assert (dict(t2), r2, n2, root2) == (dict(tree), rankdic, namedic, root)
        assert same_flat(t2.flat_tree(r2, root2), flat)
    tree['grown'] = root
    assert tree.flat_tree(rankdic, root) is None
    # a changed input file is a different cache entry
    fp = next(iter(v for v in kw.values() if isinstance(v, list)))[0]
    os.utime(fp, ns=(1, 1))
    with redirect_stdout(io.StringIO()):
        build_hierarchy(cache_dir=cache, **kw)
    assert len(os.listdir(cache)) == 2


def test_fill_root_variants():
    """tree.fill_root (tree.py:302-388): one top node is sealed, several get a
    new parent named by the first unused integer, dangling parents become top
    nodes."""
    t = {'a': 'b', 'b': 'c'}
    assert fill_root(t) == 'c' and t == {'a': 'b', 'b': 'c', 'c': 'c'}
    t = {'a': 'x', 'b': 'y', '1': 'x'}
    assert fill_root(t) == '2'
    assert t == {'a': 'x', 'b': 'y', '1': 'x', 'x': '2', 'y': '2', '2': '2'}
    t = {'a': None, 'b': 'a'}
    assert fill_root(t) == 'a' and t['a'] == 'a'
    assert fill_root({}) is None


def test_session_takes_the_flat_tree(inputs):
    """classify() with the loader's tree gives what it gives with plain
    dicts (the flat form is used instead of being rebuilt)."""
    from tests.oracle_engine import make_factory
    from woltka_b200 import workflow
    from woltka_b200.session import Session
    with redirect_stdout(io.StringIO()):
        tree, rankdic, namedic, root = build_hierarchy(
            nodes_fps=[os.path.join(inputs, 'nodes.dmp')])
    fac = make_factory(tree, rankdic, root, ['genus', 'free'])
    sess = Session(['genus', 'free'], tree, rankdic, root, False, None, False,
                   False, False, None, fac, 0, None, None, None, None, False)
    assert sess.ft is tree.flat
    sess.close()
    qry, sub = ['r1', 'r2', 'r3'], [{'11'}, {'11', '12'}, {'12', '21'}]

    def mapper(fh, fmt=None, excl=None, n=None):
        yield qry, sub
    res = []
    for t in (tree, dict(tree)):
        with redirect_stdout(io.StringIO()):
            res.append(workflow.classify(
                mapper, {os.path.join(inputs, 'names.tsv'): 'S'}, tree=t,
                rankdic=rankdic, root=root, ranks=['genus', 'free'],
                _engine_factory=make_factory(t, rankdic, root,
                                             ['genus', 'free'])))
    assert res[0] == res[1] and res[0]['genus']['S'] == {'10': 2.5, '20': 0.5}


def test_against_the_reference_on_its_own_data():
    from baseline.reference_arm import find_reference
    wf, where = find_reference()
    if wf is None:
        pytest.skip(where)
    d = os.path.join(where, 'woltka', 'tests', 'data')
    tax = os.path.join(d, 'taxonomy')
    cases = [dict(names_fps=[f'{tax}/names.dmp'], nodes_fps=[f'{tax}/nodes.dmp']),
             dict(newick_fps=[f'{d}/tree.nwk']),
             dict(lineage_fps=[f'{tax}/lineages.txt']),
             dict(columns_fps=[f'{tax}/rank_names.tsv']),      # (a conflict)
             dict(columns_fps=[f'{tax}/rank_tids.tsv']),
             dict(map_fps=[f'{tax}/nucl/nucl2g.txt']),
             dict(lineage_fps=[f'{tax}/nucl/nucl2lineage.txt']),
             dict(newick_fps=[f'{d}/tree.nwk'], map_fps=[f'{tax}/nucl/nucl2g.txt']),
             dict(map_fps=[f'{d}/function/uniref/uniref.map.xz',
                           f'{d}/function/go/process.tsv.xz'])]
    for kw in cases:
        assert run_case(build_hierarchy, '', kw) == \
            run_case(wf.build_hierarchy, '', kw), kw
