"""cProfile of the device-reader run of scratch/file_bench.py (where the host
time of file -> profile goes)."""
import cProfile, io, os, pstats, sys, tempfile, time
from contextlib import redirect_stdout
import numpy as np
sys.path.insert(0, '.')
from woltka_b200 import synth, workflow

n_rec = int(sys.argv[1]) if len(sys.argv) > 1 else 6_000_000
tax = synth.Taxonomy(seed=42)
ids = tax.ids()
tree = {ids[i]: ids[tax.parent[i]] for i in range(tax.T)}
rankdic = {ids[i]: tax.rank_names[tax.node_rank[i]] for i in range(tax.T)
           if tax.node_rank[i] >= 0}
q, s, _, nq = synth.gen_hits(n_rec, seed=1002)
q, s = q.numpy(), s.numpy()
gid = [tax.genome_id(g).encode() for g in range(tax.n_genomes)]
tail = b'\t1\t42\t150M\t*\t0\t0\t' + b'A' * 50 + b'\t' + b'I' * 50 + b'\n'
fp = os.path.join(tempfile.mkdtemp(), 'S1.sam')
with open(fp, 'wb') as f:
    f.write(b'@HD\tVN:1.0\tSO:unsorted\n')
    for a in range(0, n_rec, 500_000):
        f.write(b''.join(b'r%d\t0\t%s%s' % (qi, gid[si], tail)
                         for qi, si in zip(q[a:a + 500_000].tolist(),
                                           s[a:a + 500_000].tolist())))
def run():
    with redirect_stdout(io.StringIO()):
        return workflow.classify(workflow.plain_mapper, {fp: 'S1'}, tree=tree,
                                 rankdic=rankdic, root=ids[0], ranks=['genus'])
run()
t0 = time.perf_counter(); run(); print('warm run', time.perf_counter() - t0, 's')
pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats('cumulative').print_stats(28)
print(st.getvalue()[:6000])

# ---- the same for --coords
co, gb, ge = synth.gen_genes(1000, 1000)
d = os.path.dirname(fp)
synth.write_coords(os.path.join(d, 'coords.txt'), co, gb, ge)
n_ord = 12_000_000
qi, ci, bg, en, ln, nq3 = synth.gen_reads(n_ord)
qi, ci, bg, ln = qi.numpy(), ci.numpy(), bg.numpy(), ln.numpy()
fp3 = os.path.join(d, 'O1.sam')
with open(fp3, 'wb') as f:
    f.write(b'@HD\tVN:1.0\tSO:unsorted\n')
    for a in range(0, n_ord, 500_000):
        sl = slice(a, a + 500_000)
        f.write(b''.join(
            b'R%d\t0\tC%d\t%d\t42\t%s\t*\t0\t0\t*\t*\n' % (
                q_, c_, b_ + 1, b'150M' if l_ == 150 else b'70M2D78M2S')
            for q_, c_, b_, l_ in zip(qi[sl].tolist(), ci[sl].tolist(),
                                      bg[sl].tolist(), ln[sl].tolist())))
with redirect_stdout(io.StringIO()):
    mapper, chunk = workflow.build_mapper(os.path.join(d, 'coords.txt'), None, 80, None)
def run3():
    with redirect_stdout(io.StringIO()):
        return workflow.classify(mapper, {fp3: 'O1'}, ranks=['none'], chunk=chunk)
run3()
t0 = time.perf_counter(); run3(); print('coords warm run', time.perf_counter() - t0, 's', workflow.LAST_READER)
pr = cProfile.Profile(); pr.enable(); run3(); pr.disable()
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats('cumulative').print_stats(30)
print(st.getvalue()[:6500])
