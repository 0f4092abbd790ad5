"""The pure-Python restatement (oracle/pyport.py) against the golden outputs
of the real reference — the string-level twin of the C oracle."""
import json
import os
from os.path import join, isdir

import pytest

from oracle import pyport
from tests.test_golden import (GOLD, DATA, CASES, dec, round_like_reference,
                               check)
from woltka_b200.align import plain_mapper
from woltka_b200.workflow import readzip, _read_strata

PLAIN = []
for name in CASES:
    with open(join(GOLD, f'{name}.json')) as f:
        if not json.load(f)['coords']:
            PLAIN.append(name)


@pytest.mark.parametrize('name', PLAIN)
def test_pyport_matches_reference(name):
    with open(join(GOLD, f'{name}.json')) as f:
        case = json.load(f)
    inp = join(DATA, case['input'])
    if isinstance(case['files'], dict):
        files = {join(inp, k) if isdir(inp) else inp: v
                 for k, v in case['files'].items()}
    else:
        files = {(join(inp, k) if isdir(inp) else inp): None
                 for k in case['files']}
    ranks = case['ranks']
    total = {r: {} for r in ranks}
    maps = {} if case.get('expected_maps') else None
    excl = set(case['exclude']) if case['exclude'] else None
    strata_of = None
    if case['strata']:
        sdir = join(DATA, case['strata'])
        smap = {fn.split('.')[0]: join(sdir, fn) for fn in os.listdir(sdir)}
        cache = {}
        strata_of = lambda s: cache.setdefault(s, _read_strata(smap[s]))
    for fp in sorted(files):
        with readzip(fp) as fh:
            chunks = plain_mapper(fh, fmt=case['fmt'], excl=excl, n=1024)
            data = pyport.classify_chunks(
                chunks, ranks, case['tree'], case['rankdic'], case['root'],
                case['uniq'], case['major'] and case['major'] / 100,
                case['above'], case['subok'], case['unasgd'],
                demux=case['demux'],
                samples=set(case['samples']) if case['demux'] and
                case['samples'] else None,
                sample=files[fp], trimsub=case['trimsub'], maps=maps,
                namedic=case.get('namedic'), sizes=case.get('sizes'),
                strata_of=strata_of)
        for r in ranks:
            for s, prof in data[r].items():
                tgt = total[r].setdefault(s, {})
                for k, v in prof.items():
                    tgt[k] = tgt.get(k, 0) + v
    check.relative = bool(case.get('sizes'))
    check(total, dec(case['expected_raw']), dec(case['expected_rounded']))
    if maps is not None:
        # taxon:count lists tie-break on the taxon name, so set order of the
        # subjects cannot change a line
        got = {str(r): d for r, d in maps.items()}
        assert got == case['expected_maps']
