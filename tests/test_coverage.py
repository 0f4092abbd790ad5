"""Known answers of the reference's coverage tests
(/root/reference/woltka/tests/test_range.py) for woltka_b200.coverage:
range_mapper, the (sample, subject) interval store behind parse_ranges /
merge_ranges / calc_coverage, and the writer's coordinate formats."""
import os
import tempfile
from io import StringIO

import pytest

from oracle import pyport
from tests.oracle_engine import make_factory
from woltka_b200.coverage import range_mapper, Coverage, coverage_offsets

ENGINES = ['oracle', pytest.param('gpu', marks=pytest.mark.gpu)]

ALN = '\n'.join((
    'R1	G1	95	20	0	0	1	20	10	29	1	1',
    'R2	G1	95	20	0	0	1	20	16	35	1	1',
    'R3	G2	95	20	0	0	1	20	39	21	1	1',
    'R3	G3	95	20	0	0	1	20	88	70	1	1',
    'R4	G2	95	20	0	0	20	1	41	22	1	1',
    'R5	G3	95	20	0	0	20	1	30	49	1	1',
    'R5	G3	95	20	0	0	20	1	50	69	1	1',
    '# this is not an alignment'))


def test_range_mapper_kats():
    # tests/test_range.py:33-72
    exp = [('R1', {'G1': [9, 29]}), ('R2', {'G1': [15, 35]}),
           ('R3', {'G2': [20, 39], 'G3': [69, 88]}), ('R4', {'G2': [21, 41]}),
           ('R5', {'G3': [29, 49, 49, 69]})]
    for kw in ({}, {'fmt': 'b6o'}):
        obs = list(range_mapper(StringIO(ALN), **kw))
        assert len(obs) == 1
        assert list(obs[0][0]) == [x[0] for x in exp]
        assert list(obs[0][1]) == [x[1] for x in exp]
    obs = list(range_mapper(StringIO(ALN), n=3))
    assert [list(c[0]) for c in obs] == [['R1', 'R2', 'R3'], ['R4', 'R5']]
    assert list(obs[1][1]) == [x[1] for x in exp[3:]]
    obs = list(range_mapper(StringIO(ALN), excl={'G1'}))
    assert list(obs[0][0]) == [x[0] for x in exp[2:]]
    assert list(obs[0][1]) == [x[1] for x in exp[2:]]


def test_merge_ranges_kats():
    # tests/test_range.py:74-88 (the checker of the GPU store)
    assert pyport.merge_ranges([1, 3, 2, 4, 6, 8, 7, 9]) == [1, 4, 6, 9]
    assert pyport.merge_ranges([4, 6, 1, 4, 5, 9]) == [1, 9]
    assert pyport.merge_ranges([1, 2, 2, 3, 3, 4]) == [1, 4]
    assert pyport.merge_ranges([]) == []


def _engine(kind):
    if kind == 'oracle':
        return make_factory(None, None, None, ['none'], False)(0)
    from woltka_b200.engine import Engine
    return Engine(0)


@pytest.mark.parametrize('engine', ENGINES)
def test_parse_ranges_and_calc_coverage_kats(engine):
    # tests/test_range.py:90-146: parse_ranges followed by calc_coverage
    rmap = {'S1': (['R1', 'R2', 'R3'], [
                {'G1': [1, 100]},
                {'G1': [1, 50, 251, 300]},
                {'G1': [1, 50], 'G2': [101, 150]}]),
            'S2': (['R1', 'R3', 'R4'], [
                {'G2': [151, 200, 51, 100]},
                {'G3': [76, 125], 'G4': [26, 75, 101, 150]},
                {'G2': [26, 75], 'G3': [1, 50]}]),
            'S3': (['R2', 'R4', 'R5'], [
                {'G1': [1, 50], 'G2': [51, 100]},
                {'G1': [51, 100, 151, 200]},
                {'G1': [101, 150], 'G2': [1, 50]}])}
    exp = {'S1': {'G1': [1, 100, 251, 300], 'G2': [101, 150]},
           'S2': {'G2': [26, 100, 151, 200], 'G3': [1, 50, 76, 125],
                  'G4': [26, 75, 101, 150]},
           'S3': {'G1': [1, 50, 51, 100, 101, 150, 151, 200],
                  'G2': [1, 50, 51, 100]}}
    eng = _engine(engine)
    try:
        cov = Coverage(eng)
        for sample, (_, subque) in rmap.items():
            for ranges in subque:
                cov.add(sample, ranges)
            # a merge per sample, like parse_ranges' auto-compress: merged
            # ranges merge again with what follows
            cov.flush()
            eng.cover_ranges()
        assert cov.result() == exp
    finally:
        eng.close()


@pytest.mark.parametrize('engine', ENGINES)
def test_write_coverage_formats(engine):
    # tests/test_range.py:148-217: bed / gff / 1e / 0i coordinates
    cov_in = {'S1': {'G1': [0, 100, 250, 300], 'G2': [100, 150]},
              'S3': {'G1': [0, 200], 'G2': [0, 100]}}
    exp = {None: ['G1\t0\t100', 'G1\t250\t300', 'G2\t100\t150'],
           'bed': ['G1\t0\t100', 'G1\t250\t300', 'G2\t100\t150'],
           'gff': ['G1\t1\t100', 'G1\t251\t300', 'G2\t101\t150'],
           '1e': ['G1\t1\t101', 'G1\t251\t301', 'G2\t101\t151'],
           '0i': ['G1\t0\t99', 'G1\t250\t299', 'G2\t100\t149']}
    eng = _engine(engine)
    try:
        cov = Coverage(eng)
        for sample, subs in cov_in.items():
            for sub, rr in subs.items():
                cov.add(sample, {sub: rr})
        for fmt, lines in exp.items():
            d = tempfile.mkdtemp()
            cov.write(d, fmt)
            with open(os.path.join(d, 'S1.cov')) as f:
                assert f.read().splitlines() == lines
            assert sorted(os.listdir(d)) == ['S1.cov', 'S3.cov']
    finally:
        eng.close()
    with pytest.raises(ValueError, match='Invalid coverage format: hello.'):
        coverage_offsets('hello')
    with pytest.raises(ValueError, match='Invalid coverage format: xe.'):
        coverage_offsets('xe')
