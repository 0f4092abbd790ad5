import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from tests import cases
from woltka_b200 import synth
from woltka_b200.distributed import shard_bounds, merge_engine
from woltka_b200.engine import Engine
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
tax = synth.Taxonomy(seed=7, level_sizes=[1, 2, 5, 12, 30, 60, 150, 400], n_genomes=900)
case = cases.Case(tax, n_extra=10, internal_subjects=20, seed=3)
q, s = cases.random_hits(case, 40000, seed=5, kmax=19, p=0.2)
nq = int(q.max()) + 1
rng = np.random.default_rng(1)
q_sample = np.sort(rng.integers(0, 4, nq)).astype(np.int32)
q_stratum = rng.integers(-1, 30, nq).astype(np.int32)
cuts = shard_bounds(q, world)
a, b = cuts[rank], cuts[rank + 1]
for strata in (False, True):
    eng = Engine(rank)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    kinds, tab, _ = case.tables(['genus', 'none'])
    eng.set_tree(case.ft.parent, 0)
    eng.set_plan(kinds, 0, 0.8, 4, case.NF)
    eng.set_subjects(tab, case.sub_node)
    eng.classify_chunk(q[a:b], s[a:b], q_sample, q_stratum if strata else None, 0)
    loc = cases.collect(eng, 4, case.NF)
    print(rank, 'local', strata, int(loc[0].sum()), len(loc[1]), len(loc[2]), flush=True)
    merge_engine(eng, dst=0, dense=not strata, strata=strata)
    if rank == 0:
        got = cases.collect(eng, 4, case.NF)
        exp = cases.run_oracle(case, ['genus', 'none'], 0, 0.8, q, s, n_samples=4, q_sample=q_sample,
                               q_stratum=q_stratum if strata else None)
        print('merged', strata, 'units eq', np.array_equal(got[0], exp[0]), int(got[0].sum()), int(exp[0].sum()),
              'ovf', len(got[1]), len(exp[1]), got[1] == exp[1], 'strata', len(got[2]), len(exp[2]), got[2] == exp[2], flush=True)
        if got[1] != exp[1]:
            print(got[1][:5], exp[1][:5])
    eng.close()
dist.destroy_process_group()
