"""End-to-end parity against golden outputs of the unmodified reference
(tests/golden/*.json, produced by tests/golden/make_golden.py).

Each case runs the drop-in `woltka_b200.workflow.classify()` on the same
alignment files, hierarchy dicts and options the reference was given, twice:

  * not gpu: with the CPU oracle behind the engine interface — this is what
    PINS the oracle (oracle/woltka_oracle.c) and the host layer to the
    reference;
  * gpu: with the real CUDA engine through the C-ABI.

Checked: the raw dict (`int` cells equal; fractional cells equal to the
reference's floating-point sums within 1e-9 relative) and, bit-exact, the dict
after the reference's rounding rule (util.round_dict, util.py:323-354).
"""
import json
import os
from functools import partial
from os.path import join, dirname, abspath, isdir

import pytest

from tests.oracle_engine import make_factory
from woltka_b200.workflow import classify, build_mapper

HERE = dirname(abspath(__file__))
GOLD = join(HERE, 'golden')
DATA = join(GOLD, 'data')

with open(join(GOLD, 'INDEX.json')) as f:
    CASES = json.load(f)


def dec_key(k):
    if k == '\x00':
        return None
    if '\x1f' in k:
        return tuple(k.split('\x1f'))
    return k


def dec(data):
    return {rank: {dec_key(s): {dec_key(f): v for f, v in prof.items()}
                   for s, prof in samples.items()}
            for rank, samples in data.items()}


def round_like_reference(data):
    """util.round_dict with digits=None (util.py:323-354): snap to the nearest
    half within 1e-7, then Python's round (banker's); zeros are dropped."""
    out = {}
    for rank, samples in data.items():
        out[rank] = {}
        for s, prof in samples.items():
            res = {}
            for k, v in prof.items():
                near = round(v * 2) / 2
                iv = round(near) if abs(v - near) <= 1e-7 else round(v)
                if iv:
                    res[k] = iv
            out[rank][s] = res
    return out


def run_case(name, engine_factory=None):
    with open(join(GOLD, f'{name}.json')) as f:
        case = json.load(f)
    inp = join(DATA, case['input'])
    if isinstance(case['files'], dict):
        files = {join(inp, k) if isdir(inp) else inp: v
                 for k, v in case['files'].items()}
    else:
        files = [join(inp, k) if isdir(inp) else inp for k in case['files']]
    coords = join(DATA, case['coords']) if case['coords'] else None
    covdir = None
    if case.get('cov'):
        import tempfile
        covdir = tempfile.mkdtemp()
    mapper, chunk = build_mapper(coords, covdir, case['overlap'], case['chunk'])
    if coords:
        kw = dict(mapper.keywords)
        assert kw['th'] == case['overlap'] / 100
        kw['prefix'] = case['prefix']   # decided on the full coordinates file
        mapper = partial(mapper.func, **kw)
    stratmap = None
    if case['strata']:
        sdir = join(DATA, case['strata'])
        stratmap = {}
        for fn in os.listdir(sdir):
            stratmap[fn.split('.')[0]] = join(sdir, fn)
    ranks = case['ranks']
    rank2dir = None
    if engine_factory == 'oracle':
        engine_factory = make_factory(case['tree'], case['rankdic'],
                                      case['root'], ranks, case['subok'])
    if case.get('expected_maps'):
        import tempfile
        mapdir = tempfile.mkdtemp()
        rank2dir = {}
        for r in ranks:
            rank2dir[r] = join(mapdir, str(r))
            os.makedirs(rank2dir[r])
    got = classify(
        mapper, files, samples=case['samples'], fmt=case['fmt'],
        demux=case['demux'], trimsub=case['trimsub'], tree=case['tree'],
        rankdic=case['rankdic'], root=case['root'], ranks=ranks,
        uniq=case['uniq'], major=case['major'], above=case['above'],
        subok=case['subok'], unasgd=case['unasgd'], stratmap=stratmap,
        exclude=set(case['exclude']) if case['exclude'] else None,
        chunk=chunk, _engine_factory=engine_factory, rank2dir=rank2dir,
        namedic=case.get('namedic'), sizes=case.get('sizes'),
        outcov_dir=covdir,
        outcov_fmt=None if case.get('cov') in (None, True) else case['cov'])
    if covdir is not None:
        got_cov = {}
        for fn in sorted(os.listdir(covdir)):
            with open(join(covdir, fn)) as fh:
                got_cov[fn] = fh.read()
        assert got_cov == case['expected_cov'], 'coverage files differ'
    if rank2dir is not None:
        maps = {}
        for r, d in rank2dir.items():
            maps[str(r)] = {}
            for fn in sorted(os.listdir(d)):
                with open(join(d, fn)) as fh:
                    maps[str(r)][fn[:-4]] = fh.read().splitlines()
        assert maps == case['expected_maps'], 'read maps differ'
    exp_raw = dec(case['expected_raw'])
    exp_rounded = dec(case['expected_rounded'])
    if case.get('sizes'):
        check.relative = True   # size-weighted cells are ~1e-6: compare relatively
    return got, exp_raw, exp_rounded


def check(got, exp_raw, exp_rounded):
    relative, check.relative = getattr(check, 'relative', False), False
    assert set(got) == set(exp_raw)
    for rank in exp_raw:
        assert set(got[rank]) == set(exp_raw[rank]), rank
        for s in exp_raw[rank]:
            g, e = got[rank][s], exp_raw[rank][s]
            assert set(g) == set(e), (rank, s, set(g) ^ set(e))
            for k, v in e.items():
                if isinstance(v, int):
                    assert g[k] == v, (rank, s, k, g[k], v)
                else:
                    assert abs(g[k] - v) <= 1e-9 * (abs(v) if relative else
                                                    max(1.0, abs(v))), \
                        (rank, s, k, g[k], v)
    assert round_like_reference(got) == exp_rounded


@pytest.mark.parametrize('name', CASES)
def test_golden_oracle(name):
    check(*run_case(name, 'oracle'))


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_golden_gpu(name):
    check(*run_case(name, None))
