"""F2: an NCBI-scale taxonomy (2M nodes) through the reference's
build_hierarchy + the per-run flattening, vs woltka_b200.loaders with its
binary cache.  CPU only."""
import io, json, os, sys, tempfile, time
from contextlib import redirect_stdout
import numpy as np
sys.path.insert(0, '.')
from woltka_b200.loaders import build_hierarchy
from woltka_b200.hierarchy import FlatTree
from baseline.reference_arm import find_reference

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
rng = np.random.default_rng(7)
d = tempfile.mkdtemp()
ranks = ['no rank', 'superkingdom', 'phylum', 'class', 'order', 'family', 'genus', 'species']
par = np.zeros(n, dtype=np.int64)
par[1:] = (rng.random(n - 1) * np.arange(1, n) * 0.9).astype(np.int64)
depth = np.zeros(n, dtype=np.int64)
for i in range(1, n):
    depth[i] = min(depth[par[i]] + 1, 7)
with open(os.path.join(d, 'nodes.dmp'), 'w') as f:
    f.write(''.join(f'{i + 1}\t|\t{par[i] + 1}\t|\t{ranks[depth[i]]}\t|\n' for i in range(n)))
with open(os.path.join(d, 'names.dmp'), 'w') as f:
    f.write(''.join(f'{i + 1}\t|\tTaxon {i + 1}\t|\t\t|\tscientific name\t|\n' for i in range(n)))
kw = dict(nodes_fps=[os.path.join(d, 'nodes.dmp')], names_fps=[os.path.join(d, 'names.dmp')])
res = {'nodes': n}
wf, where = find_reference()
def timed(f):
    t0 = time.perf_counter()
    with redirect_stdout(io.StringIO()):
        out = f()
    return time.perf_counter() - t0, out
if wf is not None:
    t, ref = timed(lambda: wf.build_hierarchy(**kw))
    t2, _ = timed(lambda: FlatTree.from_dicts(ref[0], ref[1], ref[3]))
    res['reference_build_hierarchy_s'] = t
    res['flatten_per_run_s'] = t2
t, ours = timed(lambda: build_hierarchy(**kw))
res['ours_parse_and_flatten_s'] = t
cache = os.path.join(d, 'cache')
t, _ = timed(lambda: build_hierarchy(cache_dir=cache, **kw))
res['ours_first_run_with_cache_write_s'] = t
t, hit = timed(lambda: build_hierarchy(cache_dir=cache, **kw))
res['ours_cache_hit_s'] = t
assert dict(hit[0]) == dict(ours[0]) and hit[0].flat_tree(hit[1], hit[3]) is not None
if wf is not None:
    assert (dict(ours[0]), ours[1], ours[2], ours[3]) == tuple(ref)
print(json.dumps(res))
