#!/usr/bin/env python
"""Per-kernel SASS digest of the shipped library (profiles/r2_sass_digest.txt):
for every kernel of woltka_b200/libwoltka_b200.so the instruction count, the
opcode histogram and the counts of the instructions that prove the sm_100a
path: UBLKCP (cp.async.bulk = TMA bulk copy), SYNCS.* (mbarrier), ATOMS / RED /
ATOMG (shared and global atomics), LDS.U16 (staged uint16 tables), VOTE, SHFL,
CREDUX / REDUX.

    python scratch/sass_digest.py > profiles/r2_sass_digest.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'woltka_b200', 'libwoltka_b200.so')
KEY = ['UBLKCP', 'SYNCS.ARRIVE', 'SYNCS.PHASECHK', 'SYNCS.EXCH', 'ATOMS', 'REDG', 'ATOMG',
       'RED.', 'LDS.U16', 'LDS.128', 'VOTE', 'VOTEU', 'SHFL', 'CREDUX', 'REDUX', 'MATCH',
       'LDG.E.128', 'STG.E.128', 'BAR.SYNC', 'DMUL', 'HMMA', 'UTC']


def main():
    sass = subprocess.run(['cuobjdump', '-sass', SO], capture_output=True, text=True).stdout
    arch = re.findall(r'arch = (sm_\w+)', sass)
    blocks = re.split(r'\n\s*Function : ', sass)[1:]
    names = subprocess.run(['c++filt'], input='\n'.join(b.split('\n', 1)[0] for b in blocks),
                           capture_output=True, text=True).stdout.split('\n')
    print(f'# {os.path.relpath(SO, ROOT)}: {len(blocks)} kernels, cubin arch {sorted(set(arch))}')
    print('# kernel | instructions | registers are in profiles/*details*; key opcodes; top opcodes')
    total = collections.Counter()
    for name, blk in sorted(zip(names, blocks)):
        ops = re.findall(r'/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', blk)
        if not ops:
            continue
        hist = collections.Counter(o.split('.')[0] for o in ops)
        key = collections.OrderedDict()
        for k in KEY:
            n = sum(1 for o in ops if o.startswith(k))
            if n:
                key[k] = n
        total.update({k: v for k, v in key.items()})
        short = re.sub(r'\(.*', '', name)
        short = re.sub(r'^void ', '', short).replace('wk::', '')
        top = ' '.join(f'{o}:{n}' for o, n in hist.most_common(8))
        print(f'{short} | {len(ops)} | ' + ' '.join(f'{k}={v}' for k, v in key.items()) + f' | {top}')
    print('# library totals: ' + ' '.join(f'{k}={v}' for k, v in total.items()))


if __name__ == '__main__':
    main()
