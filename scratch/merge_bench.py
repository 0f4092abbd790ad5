"""Where the time of a strata merge goes (one GPU, two contexts standing for
two ranks): export of B's cells, reserve + import into A."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, '.')
import bench
from woltka_b200.engine import Engine
from woltka_b200._lib import KIND_RANK

n = int(sys.argv[1]) if len(sys.argv) > 1 else 62_500_000
dev = torch.device('cuda', 0)
engs = []
for r in range(2):
    q, s, qs, qt, nq, tab, n_ko = bench.make_cfg5(n, 1005 + r, dev, n_samples=8)
    qs = (qs + r * 8).contiguous()
    T = 1 + n_ko
    e = Engine(0)
    e.set_stream(torch.cuda.current_stream().cuda_stream)
    e.set_tree(np.zeros(T, dtype=np.int32), 0)
    e.set_plan(np.array([KIND_RANK], dtype=np.int32), 0, 0.0, 16, T)
    e.set_subjects(tab, None)
    engs.append((e, (q, s, qs, qt, nq)))
def t(label, f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize()
    print(f'{label}: {(time.perf_counter() - t0) * 1e3:.2f} ms'); return r
for rep in range(3):
    print('--- rep', rep)
    for e, (q, s, qs, qt, nq) in engs:
        t('reset', e.reset_counts)
        t('classify', lambda: e.classify_device(q.data_ptr(), s.data_ptr(), n, qs.data_ptr(), qt.data_ptr(), nq, 0))
    A, B = engs[0][0], engs[1][0]
    k, u = t('export B', B.strata_export)
    print('cells', k.numel())
    k, u = k.clone(), u.clone()
    t('reserve A', lambda: A.strata_reserve(k.numel()))
    t('import A', lambda: A.strata_import(k, u))
    t('export A (merged)', A.strata_export)
