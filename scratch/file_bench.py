"""File -> profile through the drop-in classify(): a synthetic SAM file
(cfg2 shape: 10k genomes, 21,603-node taxonomy, genus) read by the device
reader and by the host reader."""
import io, json, os, sys, time, tempfile
from contextlib import redirect_stdout
import numpy as np
sys.path.insert(0, '.')
from woltka_b200 import synth, workflow

n_rec = int(sys.argv[1]) if len(sys.argv) > 1 else 12_000_000
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
size = os.path.getsize(fp)
res = {}
for name, env in (('device_reader', None), ('host_reader', '1')):
    if env:
        os.environ['WOLTKA_B200_HOST_READER'] = env
    best = None
    for rep in range(2):
        t0 = time.perf_counter()
        with redirect_stdout(io.StringIO()):
            out = workflow.classify(workflow.plain_mapper, {fp: 'S1'}, tree=tree,
                                    rankdic=rankdic, root=ids[0], ranks=['genus'])
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        if env:
            break
    res[name] = {'seconds': best, 'records_per_s': n_rec / best,
                 'reader': workflow.LAST_READER,
                 'checksum': float(sum(out['genus']['S1'].values()))}
assert abs(res['device_reader']['checksum'] - res['host_reader']['checksum']) < 1e-6

# ---- --coords: cfg3-shaped reads against a gene table (1000 contigs x 1000 genes)
n_ord = int(sys.argv[2]) if len(sys.argv) > 2 else 30_000_000
d = os.path.dirname(fp)
co, gb, ge = synth.gen_genes(1000, 1000)
synth.write_coords(os.path.join(d, 'coords.txt'), co, gb, ge)
qi, ci, bg, en, ln, nq3 = synth.gen_reads(n_ord)
fp3 = os.path.join(d, 'O1.sam')
qi, ci, bg, ln = qi.numpy(), ci.numpy(), bg.numpy(), ln.numpy()
with open(fp3, 'wb') as f:
    f.write(b'@HD\tVN:1.0\tSO:unsorted\n')
    for a in range(0, n_ord, 500_000):
        sl = slice(a, a + 500_000)
        f.write(b''.join(
            b'R%d\t0\tC%d\t%d\t42\t%s\t*\t0\t0\t*\t*\n' % (
                q_, c_, b_ + 1, b'150M' if l_ == 150 else b'70M2D78M2S')
            for q_, c_, b_, l_ in zip(qi[sl].tolist(), ci[sl].tolist(),
                                      bg[sl].tolist(), ln[sl].tolist())))
t0 = time.perf_counter()
with redirect_stdout(io.StringIO()):
    mapper, chunk = workflow.build_mapper(os.path.join(d, 'coords.txt'), None, 80, None)
t_load = time.perf_counter() - t0
res3 = {'records': n_ord, 'queries': int(nq3), 'file_bytes': os.path.getsize(fp3),
        'genes': int(len(gb)), 'build_mapper_seconds': t_load}
os.environ.pop('WOLTKA_B200_HOST_READER', None)
for name, env, n_use in (('device_reader', None, n_ord), ('host_reader', '1', None)):
    use = fp3
    if env:
        # the host reader on a slice of the file (it is ~30x slower)
        os.environ['WOLTKA_B200_HOST_READER'] = env
        use = os.path.join(d, 'O2.sam')
        with open(fp3, 'rb') as f, open(use, 'wb') as g:
            g.write(f.read(60 << 20).rsplit(b'\nR', 1)[0] + b'\n')
        n_use = open(use, 'rb').read().count(b'\n') - 1
    best = None
    for rep in range(1 if env else 2):
        t0 = time.perf_counter()
        with redirect_stdout(io.StringIO()):
            out = workflow.classify(mapper, {use: 'O1'}, ranks=['none'], chunk=chunk)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    res3[name] = {'seconds': best, 'records': n_use, 'records_per_s': n_use / best,
                  'reader': workflow.LAST_READER,
                  'checksum': float(sum(out['none']['O1'].values()))}
os.environ.pop('WOLTKA_B200_HOST_READER', None)
res['coords'] = res3
print(json.dumps({'what': 'SAM file -> genus profile through woltka_b200.workflow.classify()',
                  'records': n_rec, 'queries': int(nq), 'file_bytes': size, **res}))
