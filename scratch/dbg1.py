import sys, numpy as np
sys.path.insert(0,'.')
from tests import cases
from woltka_b200 import synth
from woltka_b200.engine import Engine
tax = synth.Taxonomy(seed=7, level_sizes=[1, 2, 5, 12, 30, 60, 150, 400], n_genomes=900)
c = cases.Case(tax, n_extra=40, internal_subjects=60, seed=3)
eng = Engine(0)
def cmp(name, ent, fl, q, s, **kw):
    g = cases.run_engine(eng, c, ent, fl, 0.8, q, s, **kw)
    o = cases.run_oracle(c, ent, fl, 0.8, q, s, **kw)
    bad = np.argwhere(g[0]!=o[0])
    print(name, 'n=',len(q),'sum gpu',g[0].sum()/720720,'sum ora',o[0].sum()/720720,'bad cells',len(bad))
    for b in bad[:6]:
        print('   cell',b,'gpu',g[0][tuple(b)]/720720,'ora',o[0][tuple(b)]/720720)
    return g,o
for cache in (0,-1):
    eng.set_tuning(0,0,cache)
    print('cache',cache)
    q = np.arange(1000,dtype=np.int32); s=(q*7%c.V).astype(np.int32)
    cmp('own-query', ['genus'],0,q,s)
    cmp('own-query none', ['none'],0,q,s)
    q = np.repeat(np.arange(500,dtype=np.int32),2); s=(np.arange(1000)*7%c.V).astype(np.int32)
    cmp('pairs genus', ['genus'],0,q,s)
    cmp('pairs none', ['none'],0,q,s)
    q,s = cases.random_hits(c, 40, seed=11)
    cmp('rand40', ['genus'],0,q,s)
    q,s = cases.random_hits(c, 20000, seed=11)
    cmp('rand20000', ['genus'],0,q,s)
    cmp('rand20000 uniq', ['genus'],1,q,s)
