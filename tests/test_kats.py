"""Known answers of the reference's unit tests for the assign / count path
(/root/reference/woltka/tests/test_classify.py, test_tree.py,
test_workflow.py), driven through the drop-in classify()."""
import os
import tempfile

import pytest

from tests.oracle_engine import make_factory
from woltka_b200.workflow import classify, build_mapper

ENGINES = ['oracle', pytest.param('gpu', marks=pytest.mark.gpu)]

TREE = {'G1': 'T1', 'G2': 'T1', 'G3': 'T2', 'T1': 'T0', 'T2': 'T0',
        'T0': 'T0'}


def run(engine, queries, ranks, tree=None, rankdic=None, root=None,
        sample='S1', **kw):
    """queries: list of (name, iterable of subjects)."""
    qryque = [q for q, _ in queries]
    subque = [set(s) for _, s in queries]

    def mapper(fh, fmt=None, excl=None, n=None):
        yield qryque, subque

    fd, path = tempfile.mkstemp()
    os.close(fd)
    try:
        factory = make_factory(tree, rankdic, root, ranks,
                               kw.get('subok', False)) \
            if engine == 'oracle' else None
        return classify(mapper, {path: sample}, tree=tree, rankdic=rankdic,
                        root=root, ranks=ranks, _engine_factory=factory, **kw)
    finally:
        os.remove(path)


@pytest.mark.parametrize('engine', ENGINES)
def test_assign_readmap_kats(engine):
    # tests/test_workflow.py:549-589
    q = [('R1', ['G1']), ('R2', ['G1', 'G2']), ('R3', ['G2', 'G3'])]
    assert run(engine, q, ['none'])['none']['S1'] == \
        {'G1': 1.5, 'G2': 1.0, 'G3': 0.5}
    assert run(engine, q, ['none'], uniq=True, unasgd=True)['none']['S1'] == \
        {'G1': 1, 'Unassigned': 2}
    assert run(engine, q, ['free'], tree=TREE)['free']['S1'] == \
        {'T0': 1, 'T1': 2}
    rankdic = {'T1': 'ko', 'T2': 'ko', 'T0': 'mo'}
    assert run(engine, q, ['ko'], tree=TREE, rankdic=rankdic)['ko']['S1'] == \
        {'T1': 2.5, 'T2': 0.5}


@pytest.mark.parametrize('engine', ENGINES)
def test_assign_free_kats(engine):
    # tests/test_classify.py:42-66
    kw = dict(tree=TREE, root='T0')
    one = lambda subs, **k: run(engine, [('R', subs)], ['free'], **kw,
                                **k)['free']['S1']
    assert one(['G1', 'G2']) == {'T1': 1}
    assert one(['G1', 'G2', 'G3']) == {}          # LCA is the root
    assert one(['G1']) == {'T1': 1}
    assert one(['G1'], subok=True) == {'G1': 1}
    assert one(['Gx']) == {}                      # not in the tree
    assert one(['Gx'], subok=True) == {'Gx': 1}
    assert one(['T0']) == {'T0': 1}               # parent of root is root


@pytest.mark.parametrize('engine', ENGINES)
def test_assign_rank_kats(engine):
    # tests/test_classify.py:68-107
    rankdic = {'T1': 'general', 'T2': 'general', 'T0': 'marshal'}
    kw = dict(tree=TREE, rankdic=rankdic, root='T0')
    one = lambda subs, rank, **k: run(engine, [('R', subs)], [rank],
                                      **{**kw, **k})[rank]['S1']
    assert one(['G1', 'G2', 'G3'], 'marshal') == {'T0': 1}
    assert one(['G1', 'G2'], 'general') == {'T1': 1}
    assert one(['G1'], 'general') == {'T1': 1}
    assert one(['G1', 'G2', 'G3'], 'admiral', uniq=True) == {}
    assert one(['G1', 'G2', 'G3'], 'general', uniq=True) == {}
    assert one(['G1', 'G2', 'G3'], 'general', major=60) == {'T1': 1}
    assert one(['G1', 'G2', 'G3'], 'general', major=80) == {}
    got = one(['G1', 'G2', 'G3'], 'general')
    assert got.keys() == {'T1', 'T2'}
    assert abs(got['T1'] - 2 / 3) < 1e-12 and abs(got['T2'] - 1 / 3) < 1e-12
    assert one(['G1', 'G2', 'G3'], 'general', root=None, above=True) == \
        {'T0': 1}
    assert one(['G1', 'G2', 'G3'], 'general', above=True) == {}
    assert one(['G1', 'G2', 'G3', 'Gx'], 'general', root=None,
               above=True) == {}


@pytest.mark.parametrize('engine', ENGINES)
def test_counter_kats(engine):
    # tests/test_classify.py:109-114: 1/k' per remaining occurrence
    tree = {'G1': 'Ecoli', 'G4': 'Ecoli', 'G6': 'Ecoli', 'G2': 'Cdiff',
            'G5': 'Cdiff', 'G3': 'Strep', 'Ecoli': 'root', 'Cdiff': 'root',
            'Strep': 'root', 'root': 'root'}
    rankdic = {'Ecoli': 'species', 'Cdiff': 'species', 'Strep': 'species'}
    q = [('a', ['G1']), ('b', ['G2', 'G3']),
         ('c', ['G3', 'G4', 'G5', 'G6', 'G7']), ('d', ['G4', 'G6']),
         ('e', ['G7'])]
    got = run(engine, q, ['species'], tree=tree, rankdic=rankdic,
              root='root')['species']['S1']
    assert got == {'Ecoli': 2.5, 'Cdiff': 0.75, 'Strep': 0.75}


@pytest.mark.parametrize('engine', ENGINES)
def test_demultiplex_and_strata_kats(engine):
    # tests/test_workflow.py:513-547 (demultiplex), classify.py:216-249
    q = [('S1_R1', ['G1']), ('S1_R2', ['G2']), ('S2_R1', ['G1', 'G2']),
         ('R9', ['G3']), ('S3_', ['G3']), ('S1_R3', ['G1'])]
    got = run(engine, q, ['none'], demux=True)['none']
    assert got == {'S1': {'G1': 2, 'G2': 1}, 'S2': {'G1': 0.5, 'G2': 0.5},
                   '': {'G3': 2}}
    got = run(engine, q, ['none'], demux=True, samples=['S1', 'SX'])['none']
    assert got == {'S1': {'G1': 2, 'G2': 1}}
    # stratified: only reads present in the strata map count
    d = tempfile.mkdtemp()
    with open(os.path.join(d, 'S1.txt'), 'w') as f:
        f.write('R1\tEsch\nR2\tEsch\tx\nR3\tShig\n')
    with open(os.path.join(d, 'S2.txt'), 'w') as f:
        f.write('R1\tEsch\n')
    stratmap = {'S1': os.path.join(d, 'S1.txt'),
                'S2': os.path.join(d, 'S2.txt')}
    got = run(engine, q[:3] + q[5:], ['none'], demux=True,
              stratmap=stratmap)['none']
    assert got == {'S1': {('Esch', 'G1'): 1, ('Shig', 'G1'): 1},
                   'S2': {('Esch', 'G1'): 0.5, ('Esch', 'G2'): 0.5}}


def test_build_mapper_seams():
    # tests/test_workflow.py:293-312
    obs = build_mapper()
    assert obs[0].__name__ == 'plain_mapper' and obs[1] == 1024
    d = tempfile.mkdtemp()
    fp = os.path.join(d, 'coords.txt')
    with open(fp, 'w') as f:
        f.write('>G1\n1\t10\t20\n2\t35\t50\n')
    obs = build_mapper(fp)
    assert obs[0].func.__name__ == 'ordinal_mapper' and obs[1] == 1048576
    obs = build_mapper(fp, overlap=75, chunk=50000)
    assert obs[0].keywords['th'] == 0.75 and obs[1] == 50000
    assert set(obs[0].keywords) >= {'coords', 'idmap', 'prefix', 'th'}


def test_gene_coords_encoding():
    # tests/test_ordinal.py:358-400: bit layout of the endpoint codes
    from woltka_b200.ordinal import load_gene_coords, GeneIndex
    coords, idmap, isdup = load_gene_coords(
        ('>n1', 'g1\t5\t29', 'g2\t33\t61', 'g3\t65\t94', 'gx\t108\t135'),
        sort=True)
    assert not isdup and idmap == {'n1': ['g1', 'g2', 'g3', 'gx']}
    exp = [(5 - 1 << 24) + (1 << 22) + 0, (29 << 24) + (3 << 22) + 0,
           (33 - 1 << 24) + (1 << 22) + 1, (61 << 24) + (3 << 22) + 1,
           (65 - 1 << 24) + (1 << 22) + 2, (94 << 24) + (3 << 22) + 2,
           (108 - 1 << 24) + (1 << 22) + 3, (135 << 24) + (3 << 22) + 3]
    assert coords['n1'].tolist() == exp
    gi = GeneIndex(coords, idmap, False)
    assert gi.gbeg.tolist() == [4, 32, 64, 107]
    assert gi.gend.tolist() == [29, 61, 94, 135]
    # reversed coordinates, duplicate ids -> prefix, '##' headers ignored
    coords, idmap, isdup = load_gene_coords(
        ('##genome', '>n1', 'g1\t29\t5', '#n2', 'g1\t7\t9'), sort=True)
    assert isdup and list(coords) == ['n1', 'n2']
    gi = GeneIndex(coords, idmap, True)
    assert gi.gene_ids == ['n1_g1', 'n2_g1'] and gi.gbeg.tolist() == [4, 6]
    with pytest.raises(ValueError, match='No coordinate was read'):
        load_gene_coords(())
    with pytest.raises(ValueError, match='Cannot extract coordinates'):
        load_gene_coords(('>n1', 'g1\t5'))


def test_unsupported_options_are_loud():
    # coverage needs ranges: not with --coords (the reference crashes on the
    # gene sets, range.py:142), and the coordinate format is checked up front
    # (range.py:229-245)
    from functools import partial
    from woltka_b200.ordinal import ordinal_mapper
    om = partial(ordinal_mapper, coords={}, idmap={}, prefix=False, th=0.8)
    with pytest.raises(ValueError, match='--coords'):
        classify(om, [], ranks=['none'], outcov_dir='cov')
    with pytest.raises(ValueError, match='Invalid coverage format: xyz.'):
        classify(None, [], ranks=['none'], outcov_dir='cov', outcov_fmt='xyz')


@pytest.mark.gpu
@pytest.mark.parametrize('opts', [dict(), dict(uniq=True, unasgd=True),
                                  dict(major=60), dict(above=True)])
def test_size_weighted_counts_vs_pure_python(opts):
    """--sizes (classify.counter_size, classify.py:174-213) through the
    kernels' (subject, feature) table, against the pure-Python restatement:
    multi-rank, repeats, a 70-subject query (slow path), subjects without a
    taxon."""
    import random
    from oracle import pyport
    rnd = random.Random(5)
    tree = {'root': 'root'}
    rankdic = {}
    for g in range(12):
        tree[f'genus{g}'] = 'root'
        rankdic[f'genus{g}'] = 'genus'
        for s in range(4):
            sp = f'sp{g}_{s}'
            tree[sp] = f'genus{g}'
            rankdic[sp] = 'species'
            for x in range(3):
                tree[f'G{g}_{s}_{x}'] = sp
    leaves = [k for k in tree if k.startswith('G')]
    sizes = {x: 1 / rnd.randint(1000, 9000) for x in leaves}
    sizes.update({f'X{i}': 1 / (500 + i) for i in range(5)})
    queries = []
    for i in range(3000):
        k = 70 if i % 600 == 7 else min(1 + int(rnd.expovariate(0.7)), 12)
        base = rnd.randrange(len(leaves))
        subs = [leaves[(base + rnd.randrange(9)) % len(leaves)] for _ in range(k)]
        if rnd.random() < 0.05:
            subs.append(f'X{rnd.randrange(5)}')
        queries.append((f'R{i}', subs))
    ranks = ['genus', 'species', 'none', 'free']
    kw = dict(tree=tree, rankdic=rankdic, root='root')
    got = run('gpu', queries, ranks, sizes=sizes, **kw, **opts)
    chunks = [([q for q, _ in queries], [set(s) for _, s in queries])]
    po = dict(opts)
    if 'major' in po:
        po['major'] = po['major'] / 100
    exp = pyport.classify_chunks(chunks, ranks, tree, rankdic, 'root',
                                 sample='S1', sizes=sizes, **po)
    for r in ranks:
        g, e = got[r]['S1'], exp[r]['S1']
        assert set(g) == set(e), (r, set(g) ^ set(e))
        for key, v in e.items():
            assert abs(g[key] - v) <= 1e-9 * abs(v), (r, key, g[key], v)
    # a subject without a size is the reference's error
    with pytest.raises(ValueError, match='not found in the size map'):
        run('gpu', queries, ['none'], sizes={'G0_0_0': 1.0})


def test_gene_coords_cache(tmp_path, monkeypatch):
    """WOLTKA_B200_CACHE: build_mapper reads the coordinates file once, later
    runs load the binary table; a changed file is parsed again."""
    import numpy as np
    fp = tmp_path / 'coords.txt'
    fp.write_text('##genome\n>n1\ng1\t29\t5\ng2\t40\t90\n#n2\ng1\t7\t9\n')
    plain = build_mapper(str(fp))[0].keywords
    monkeypatch.setenv('WOLTKA_B200_CACHE', str(tmp_path / 'cache'))
    for _ in range(2):                       # miss, then hit
        kw = build_mapper(str(fp))[0].keywords
        assert list(kw['coords']) == list(plain['coords'])
        for c in plain['coords']:
            assert np.array_equal(kw['coords'][c], plain['coords'][c])
        assert kw['idmap'] == plain['idmap'] and kw['prefix'] == plain['prefix']
    assert len(list((tmp_path / 'cache').iterdir())) == 1
    fp.write_text('>n1\ng1\t29\t5\n>n3\ng9\t1\t2\ng8\t5\t9\n')
    kw = build_mapper(str(fp))[0].keywords
    assert list(kw['coords']) == ['n1', 'n3'] and kw['prefix'] is False
    assert kw['idmap'] == {'n1': ['g1'], 'n3': ['g9', 'g8']}


@pytest.mark.parametrize('engine', ENGINES)
def test_assign_readmap_seam(engine, tmp_path):
    """The per-chunk seam with the reference's own vectors
    (tests/test_workflow.py:545-603): counts are added into `data`, the read
    map is appended, sizes weigh the subjects, a missing size is the
    reference's ValueError."""
    from woltka_b200.workflow import assign_readmap
    qryq = ['R1', 'R2', 'R3']
    subq = [('G1',), ('G1', 'G2'), ('G2', 'G3')]

    def call(data, rank, **kw):
        fac = make_factory(kw.get('tree'), kw.get('rankdic'), kw.get('root'),
                           [rank], kw.get('subok', False)) \
            if engine == 'oracle' else None
        assign_readmap(qryq, subq, data, rank, 'S1', {}, _engine_factory=fac,
                       **kw)
        return data[rank]['S1']

    assert call({'none': {}}, 'none') == {'G1': 1.5, 'G2': 1.0, 'G3': 0.5}
    # added into what is there (util.sum_dict), read map appended
    data = {'none': {'S1': {'G1': 1, 'G9': 2}}}
    got = call(data, 'none', rank2dir={'none': str(tmp_path)})
    assert got == {'G1': 2.5, 'G2': 1.0, 'G3': 0.5, 'G9': 2}
    assert (tmp_path / 'S1.txt').read_text().splitlines() == \
        ['R1\tG1', 'R2\tG1:1\tG2:1', 'R3\tG2:1\tG3:1']
    assert call({'none': {}}, 'none', uniq=True, unasgd=True) == \
        {'G1': 1, 'Unassigned': 2}
    assert call({'free': {}}, 'free', tree=TREE) == {'T0': 1, 'T1': 2}
    rankdic = {'T1': 'ko', 'T2': 'ko', 'T0': 'mo'}
    assert call({'ko': {}}, 'ko', tree=TREE, rankdic=rankdic) == \
        {'T1': 2.5, 'T2': 0.5}
    sizes = {'G1': 0.3, 'G2': 0.5, 'G3': 1.0}
    got = call({'ko': {}}, 'ko', tree=TREE, rankdic=rankdic, sizes=sizes)
    assert got.keys() == {'T1', 'T2'}
    assert abs(got['T1'] - 0.95) < 1e-12 and abs(got['T2'] - 0.5) < 1e-12
    # stratified (classify.counter_strat): reads without a stratum are skipped
    got = call({'none': {}}, 'none', strata={'R1': 'Esch', 'R3': 'Shig'})
    assert got == {('Esch', 'G1'): 1, ('Shig', 'G2'): 0.5, ('Shig', 'G3'): 0.5}
    del sizes['G3']
    with pytest.raises(ValueError, match='One or more subjects are not found '
                                         'in the size map.'):
        call({'ko': {}}, 'ko', tree=TREE, rankdic=rankdic, sizes=sizes)


def test_chunk_loop_helpers_under_their_reference_names(tmp_path):
    # tests/test_workflow.py:489-547, :604-612
    from woltka_b200.workflow import demultiplex, strip_suffix, read_strata
    subs = [{'G1_1', 'G1_2', 'G2_3', 'G3'}, {'G1_1', 'G1.3', 'G4_5', 'G4_x'}]
    assert list(strip_suffix(subs, sep='_')) == \
        [{'G1', 'G2', 'G3'}, {'G1', 'G1.3', 'G4'}]
    assert list(strip_suffix([{'NC_123456.1_300', 'ABCD000001.20_101'}],
                             sep='_')) == [{'NC_123456.1', 'ABCD000001.20'}]
    assert list(strip_suffix([{'G1.1', 'G1.2', 'G2'}, {'G1.1', 'G1.3', 'G3_x'}],
                             sep='.')) == [{'G1', 'G2'}, {'G1', 'G3_x'}]
    rmap = [('S1_R1', 5), ('S1_R2', 12), ('S1_R3', 3), ('S2_R1', 10),
            ('S2_R2', 8), ('S2_R4', 7), ('S3_R2', 15), ('S3_R3', 1),
            ('S3_R4', 5)]
    exp = {'S1': [('R1', 5), ('R2', 12), ('R3', 3)],
           'S2': [('R1', 10), ('R2', 8), ('R4', 7)],
           'S3': [('R2', 15), ('R3', 1), ('R4', 5)]}
    obs = demultiplex(*zip(*rmap))
    assert list(obs) == ['S1', 'S2', 'S3']
    for s in obs:
        assert list(map(tuple, obs[s])) == list(zip(*exp[s]))
    obs = demultiplex(*zip(*rmap), sep='.')
    assert obs.keys() == {''}
    assert list(map(tuple, obs[''])) == list(zip(*rmap))
    obs = demultiplex(*zip(*rmap), samples=['S1', 'S2', 'SX'])
    assert obs.keys() == {'S1', 'S2'}
    assert list(map(tuple, obs['S2'])) == list(zip(*exp['S2']))
    fp = tmp_path / 'S1.txt'
    fp.write_text('R1\tEsch\nR2\tEsch\tx\nR3\tShig\n')
    assert read_strata(str(fp)) == {'R1': 'Esch', 'R3': 'Shig'}
    bad = tmp_path / 'tree.nwk'
    bad.write_text('(a,b);\n')
    with pytest.raises(ValueError, match='No stratification information is '
                                         'found in file: tree.nwk.'):
        read_strata(str(bad))


def test_subjects_bulk_equals_subject_by_subject():
    """Session.subjects_bulk (gene identifiers of a coordinates file) leaves
    the vocabularies and table rows exactly as one subject() per name does."""
    from woltka_b200.session import Session
    tree = {'a': 'r', 'b': 'r', 'a1': 'a', 'a2': 'a', 'b1': 'b', 'r': 'r'}
    rankdic = {'a': 'genus', 'b': 'genus', 'a1': 'species', 'a2': 'species',
               'b1': 'species'}
    names = ['a1', 'zz_1', 'a2', 'b1', 'a1', 'zz_2', 'a', 'yy', 'b1_3']
    cases = [(ranks, subok, trim, prior)
             for ranks, subok, trim in (
                 (['none', 'free', 'genus', 'species'], False, None),
                 (['free'], True, None), (['genus'], False, '_'))
             for prior in (True, False)]
    for ranks, subok, trim, prior in cases:
        def make():
            return Session(ranks, tree, rankdic, 'r', False, None, False, subok,
                           False, trim,
                           make_factory(tree, rankdic, 'r', ranks, subok), 0,
                           None, None, None, None, False)
        one, bulk = make(), make()
        if prior:
            one.subject('b')
            bulk.subject('b')
        assert [one.subject(n) for n in names] == \
            bulk.subjects_bulk(names).tolist()
        # asking by name afterwards finds the bulk entries
        for n in ('yy', 'fresh', 'a2', 'zz_9'):
            assert one.subject(n) == bulk.subject(n)
        assert one.subjects_bulk(['k1', 'a1', 'k1']).tolist() == \
            bulk.subjects_bulk(['k1', 'a1', 'k1']).tolist()
        for field in ('sub_node', 'sub_feat', 'sub_name', 'sub_stratum',
                      '_tab_rows', 'extra_names', 'sub_index', 'extra_index'):
            assert getattr(one, field) == getattr(bulk, field), field
        one.close()
        bulk.close()
