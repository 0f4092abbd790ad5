"""File -> profile through the drop-in classify(): a synthetic SAM file
(cfg2 shape: 10k genomes, 21,603-node taxonomy, genus) read by the device
reader and by the host reader."""
import io, json, os, sys, time, tempfile
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
print(json.dumps({'what': 'SAM file -> genus profile through woltka_b200.workflow.classify()',
                  'records': n_rec, 'queries': int(nq), 'file_bytes': size, **res}))
