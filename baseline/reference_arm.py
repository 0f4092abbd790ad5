"""`bench.py --impl reference`: the reference's CPU path on the box's host cores.

When the unmodified reference is importable in this process — from
`baseline/_ref` (a driver-side install) or `/root/reference` (the build
container) — the arm times `woltka.workflow.classify()` itself
(/root/reference/woltka/workflow.py:162-353) on text the bench's generator
writes (SAM + nodes.dmp + taxid map, or SAM + gene coordinates):
  * one process (the reference is single-threaded), and
  * one process per host core, the samples dealt out, the profile dicts
    summed — its documented scale-out (doc/perform.md:70-92),
and reports the all-cores number as `value` (`kind: "reference"`).  The same
records classified by the CUDA path must give the same profile after
`round_profiles` when a GPU is present.

When it is not importable (the GPU box has no reference checkout) the arm
falls back to the C restatement oracle/woltka_oracle.c with OpenMP over
query-aligned ranges (`kind: "port"`) and says so.  This module and bench.py's
cpu_baseline leg are the only non-test code that executes anything under
oracle/.
"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_BIOM_STUB = '''"""Stand-in for biom-format (not installed): enough for
woltka.table / woltka.biom to import; BIOM output is not used here."""


class Table:
    def __init__(self, *a, **k):
        raise NotImplementedError('biom-format is not installed')


def load_table(*a, **k):
    raise NotImplementedError('biom-format is not installed')
'''


def find_reference():
    """Import the unmodified reference if it is on this box; returns the
    `woltka.workflow` module and where it came from, or (None, why)."""
    for cand in (os.path.join(ROOT, 'baseline', '_ref'), '/root/reference'):
        if not os.path.isdir(os.path.join(cand, 'woltka')):
            continue
        stub = tempfile.mkdtemp(prefix='biom_stub_')
        os.makedirs(os.path.join(stub, 'biom'))
        with open(os.path.join(stub, 'biom', '__init__.py'), 'w') as f:
            f.write(_BIOM_STUB)
        with open(os.path.join(stub, 'biom', 'util.py'), 'w') as f:
            f.write('def biom_open(*a, **k):\n    raise NotImplementedError\n')
        sys.dont_write_bytecode = True
        try:
            import biom  # noqa: F401
        except ImportError:
            sys.path.insert(0, stub)
        sys.path.insert(0, cand)
        try:
            import woltka.workflow as wf
            return wf, cand
        except Exception as err:           # missing numba / click ...
            sys.path.remove(cand)
            why = f'{cand}: {type(err).__name__}: {err}'
            continue
    return None, locals().get('why', 'no reference checkout on this box '
                              '(baseline/_ref, /root/reference)')


# ---- text the reference reads, from the bench's generator ----------------------
def write_cfg2_text(tmp, n_rec, n_samples, seed):
    """One SAM file per sample + nodes.dmp + taxid map; returns the paths and
    the integer columns they were written from."""
    from woltka_b200 import synth
    tax = synth.Taxonomy(seed=42)
    q, s, qs, nq = synth.gen_hits(n_rec, seed=seed, n_samples=n_samples)
    q, s, qs = q.numpy(), s.numpy(), qs.numpy()
    tax.write_nodes_dmp(os.path.join(tmp, 'nodes.dmp'))
    tax.write_taxid_map(os.path.join(tmp, 'taxid.map'))
    gid = [tax.genome_id(g) for g in range(tax.n_genomes)]
    files = {}
    rec_sample = qs[q]
    for si in range(n_samples):
        sel = np.flatnonzero(rec_sample == si)
        fp = os.path.join(tmp, f'S{si:02d}.sam')
        with open(fp, 'w') as f:
            f.write('@HD\tVN:1.0\tSO:unsorted\n')
            f.write(''.join(
                f'R{qi}\t0\t{gid[sj]}\t1\t42\t150M\t*\t0\t0\t*\t*\n'
                for qi, sj in zip(q[sel].tolist(), s[sel].tolist())))
        files[fp] = f'S{si:02d}'
    return files, (q, s, qs, nq), tax


def _ref_classify(wf, tmp, files, ranks, mode):
    """workflow.classify() of the reference on `files` (a dict path -> sample)."""
    import io
    from contextlib import redirect_stdout
    kw = dict(uniq=mode == 'uniq', above=mode == 'above',
              major=80 if mode == 'major' else None)
    with redirect_stdout(io.StringIO()):
        tree, rankdic, namedic, root = wf.build_hierarchy(
            map_fps=[os.path.join(tmp, 'taxid.map')],
            nodes_fps=[os.path.join(tmp, 'nodes.dmp')])
        mapper, chunk = wf.build_mapper()
        return wf.classify(mapper, files, fmt='sam', demux=False, tree=tree,
                           rankdic=rankdic, namedic=namedic, root=root,
                           ranks=ranks, chunk=chunk, **kw)


def _worker(job):
    cand, tmp, files, ranks, mode = job
    wf, _ = find_reference()
    t0 = time.perf_counter()
    data = _ref_classify(wf, tmp, files, ranks, mode)
    return data, time.perf_counter() - t0


def _sum_profiles(parts):
    out = {}
    for data in parts:
        for rank, samples in data.items():
            for sample, prof in samples.items():
                o = out.setdefault(rank, {}).setdefault(sample, {})
                for k, v in prof.items():
                    o[k] = o.get(k, 0) + v
    return out


def run_real(args, wf, where, metric, unit, config_of, threads, model):
    """cfg2 / cfg4 through the unmodified reference."""
    import multiprocessing as mp
    ranks = args.ranks.split(',')
    P = threads
    per = 250_000                      # records per sample file
    n_samples = max(P, args.samples)
    n = per * n_samples
    tmp = tempfile.mkdtemp(prefix='wk_ref_')
    files, cols, tax = write_cfg2_text(tmp, n, n_samples, 1002)
    fps = sorted(files)
    steps = max(1, min(args.steps, 2))
    # one process: the first 4 sample files
    one = {fp: files[fp] for fp in fps[:4]}
    _ref_classify(wf, tmp, {fps[0]: files[fps[0]]}, ranks, args.mode)  # warm-up (numba, caches)
    t0 = time.perf_counter()
    d1 = _ref_classify(wf, tmp, one, ranks, args.mode)
    dt1 = time.perf_counter() - t0
    n1 = sum(1 for fp in one for _ in open(fp)) - len(one)
    # P processes, samples dealt out, dicts summed
    jobs = [(where, tmp, {fp: files[fp] for fp in fps[i::P]}, ranks, args.mode)
            for i in range(P)]
    dts = []
    with mp.get_context('spawn').Pool(P) as pool:
        pool.map(_worker, jobs[:P])            # warm-up of every process
        for _ in range(steps):
            t0 = time.perf_counter()
            parts = pool.map(_worker, jobs)
            dts.append(time.perf_counter() - t0)
    merged = _sum_profiles([p[0] for p in parts])
    dt = sum(dts)
    value = n * steps / dt

    # the same records on the GPU (when there is one) must give the same profile
    parity = None
    try:
        import ctypes as C
        from woltka_b200 import _lib
        k = C.c_int(0)
        have_gpu = _lib.load().wk_device_count(C.byref(k)) == 0 and k.value > 0
    except Exception:
        have_gpu = False
    if have_gpu:
        import io
        from contextlib import redirect_stdout
        from woltka_b200 import workflow as ours
        kw = dict(uniq=args.mode == 'uniq', above=args.mode == 'above',
                  major=80 if args.mode == 'major' else None)
        with redirect_stdout(io.StringIO()):
            tree, rankdic, namedic, root = wf.build_hierarchy(
                map_fps=[os.path.join(tmp, 'taxid.map')],
                nodes_fps=[os.path.join(tmp, 'nodes.dmp')])
            got = ours.classify(ours.build_mapper()[0], files, fmt='sam',
                                demux=False, tree=tree, rankdic=rankdic,
                                namedic=namedic, root=root, ranks=ranks, **kw)
        wf.round_profiles(got)
        exp = {r: {s: dict(p) for s, p in v.items()} for r, v in merged.items()}
        wf.round_profiles(exp)
        parity = got == exp
        assert parity, 'GPU profile differs from the reference'
    for fp in list(files) + [os.path.join(tmp, 'nodes.dmp'),
                             os.path.join(tmp, 'taxid.map')]:
        os.unlink(fp)
    os.rmdir(tmp)
    return {
        'impl': 'reference', 'metric': metric, 'value': value, 'unit': unit,
        'n_gpus': args.gpus, 'steps': steps, 'warmup': 1,
        'ms_per_step': dt / steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'python str/dict',
        'data': 'synthetic',
        'config': config_of(args.workload, args.records, ranks, args.mode,
                            args.samples),
        'cpu_baseline': {
            'value': value, 'unit': unit, 'cores': P, 'kind': 'reference',
            'cpu_model': model, 'reference_from': where,
            'single_process_value': n1 / dt1,
            'sample': f'{n} records of the bench generator per step as '
                      f'{n_samples} SAM files of {per} records; the unmodified '
                      f'woltka.workflow.classify() in {P} processes with the '
                      f'sample files dealt out and the profiles summed '
                      f'(doc/perform.md:70-92); single process: {n1} records',
            'gpu_profile_equals_reference': parity},
        'e2e': {'value': value, 'unit': unit, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0}


# ---- the C port (no reference on this box) ----------------------------------------
def run_port(args, why, metric, unit, config_of, threads, model):
    from oracle import oracle as O
    from woltka_b200 import synth
    steps, warmup = args.steps, args.warmup
    note = f'the reference is not on this box ({why}): C restatement ' \
           f'oracle/woltka_oracle.c'
    if args.workload == 'cfg3':
        n = min(args.records, 1_000_000)
        coff, gb, ge = synth.gen_genes()
        rq, rc, rb, re_, rl, nq = synth.gen_reads(n, seed=1003)
        cols = [x.numpy() for x in (rc, rb, re_, rl)]
        steps, warmup, cores = min(steps, 5), 0, 1
        t0 = time.perf_counter()
        for _ in range(steps):
            O.ordinal_match(*cols, 0.8, coff, gb, ge)
        dt = time.perf_counter() - t0
        entries, sample = ['none'], f'{n} reads, sweep matcher only, 1 thread'
    elif args.workload == 'cfg5':
        import bench
        from woltka_b200._lib import KIND_RANK
        n = min(args.records, 5_000_000)
        q, s, qs, qt, nq, tab, n_ko = bench.make_cfg5(n, 1005, 'cpu')
        T = 1 + n_ko
        node_rank = np.zeros(T, dtype=np.int32)
        node_rank[0] = -1
        kw = dict(parent=np.zeros(T, dtype=np.int32), node_rank=node_rank,
                  root=0, sub_node=tab[0].astype(np.int32), sub_feat=None,
                  kinds=np.array([KIND_RANK], dtype=np.int32), target_rank=[0],
                  flags=0, n_samples=8, n_features=T, q_sample=qs.numpy(),
                  q_stratum=qt.numpy(), n_threads=threads)
        q, s = q.numpy(), s.numpy()
        steps, warmup, cores = max(1, min(steps, 3)), 1, threads
        O.classify(q, s, **kw)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.classify(q, s, **kw)
        dt = time.perf_counter() - t0
        entries = ['ko']
        sample = f'{n} records of the same generator per step, {threads} OpenMP threads'
    else:
        n = min(args.records, args.cpu_sample)
        case = synth.Case(synth.Taxonomy(seed=42))
        entries = args.ranks.split(',')
        flags = synth.MODES[args.mode]
        q, s, qs, nq = synth.gen_hits(n, seed=1002, n_samples=args.samples)
        q, s = q.numpy(), s.numpy()
        kinds, _, trk = case.tables(entries)
        kw = dict(parent=case.ft.parent, node_rank=case.ft.node_rank, root=0,
                  sub_node=case.sub_node, sub_feat=case.sub_feat, kinds=kinds,
                  target_rank=trk, flags=flags, major_th=0.8,
                  n_features=case.NF, n_threads=threads)
        if args.samples > 1:
            kw.update(n_samples=args.samples, q_sample=qs.numpy())
        cores = threads
        steps, warmup = min(steps, 10), min(warmup, 3)
        for _ in range(warmup):
            O.classify(q, s, **kw)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.classify(q, s, **kw)
        dt = time.perf_counter() - t0
        sample = f'{n} records of the same generator per step, {threads} OpenMP threads'
    val = n * steps / dt
    return {
        'impl': 'reference', 'metric': metric, 'value': val, 'unit': unit,
        'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup,
        'ms_per_step': dt / steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int32',
        'data': 'synthetic',
        'config': config_of(args.workload, args.records, entries, args.mode,
                            args.samples),
        'cpu_baseline': {'value': val, 'unit': unit, 'cores': cores,
                         'kind': 'port', 'cpu_model': model,
                         'sample': sample, 'note': note},
        'e2e': {'value': val, 'unit': unit, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0}


def run(args, metric, unit, config_of, threads, model):
    wf, where = find_reference()
    if wf is not None and args.workload in ('cfg2', 'cfg4'):
        return run_real(args, wf, where, metric, unit, config_of, threads, model)
    why = where if wf is None else 'cfg3 / cfg5 text generation is not wired to the reference arm'
    return run_port(args, why, metric, unit, config_of, threads, model)
