#!/usr/bin/env python
"""Benchmark of the classify hot path (BASELINE.json metric: alignment
records classified per second; achieved HBM GB/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg4|cfg5]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N
    python bench.py --impl reference        # CPU arm (the reference itself when
                                            # it is importable, else the C port)

A step is one pass of the hot path over one batch of synthetic records
(SURVEY.md §8d generator) and ends with the merged result on rank 0.  `value`
is timed with the int32 SoA columns already resident in HBM; `e2e` goes
through the reference-facing C-ABI call with pinned HOST buffers (H2D of the
columns and D2H of the result inside the timed region).  Every rank works on
its own batch (weak scaling: its own samples of one shared table); per step
the per-rank results are merged on rank 0: ONE NCCL reduce of the dense units
table, and for stratified plans the strata cells sent to rank 0 and added by
key (woltka_b200.distributed.merge_engine).

The default workload is cfg2 (BASELINE.json configs[1]); its line carries, as
`extra`, short runs of cfg3, cfg4 and cfg5 at the same number of GPUs (cfg4 /
cfg5 at 8 GPUs are BASELINE.json configs[3] / [4]: 1e9 records over 64 samples,
5e8 stratified records).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'alignment records classified per second'
UNIT = 'records/s'
# records per GPU: BASELINE.json sizes / 8 GPUs for the 8-GPU configs
DEFAULT_RECORDS = {'cfg2': 100_000_000, 'cfg3': 100_000_000,
                   'cfg4': 125_000_000, 'cfg5': 62_500_000}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg2',
                    choices=['cfg2', 'cfg3', 'cfg4', 'cfg5'])
    ap.add_argument('--records', type=int, default=0,
                    help='records per GPU (default: the config\'s size)')
    ap.add_argument('--mode', default='default',
                    choices=['default', 'major', 'uniq', 'above'])
    ap.add_argument('--ranks', default='genus')
    ap.add_argument('--cpu-sample', type=int, default=20_000_000)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-extra', action='store_true',
                    help='default workload: skip the cfg3 / cfg4 / cfg5 extras')
    ap.add_argument('--samples', type=int, default=1,
                    help='samples per GPU (per-query sample column if > 1)')
    ap.add_argument('--opt', action='append', default=[],
                    help='name=value knob of the context (wk_set_option)')
    args = ap.parse_args()
    args.explicit_records = args.records
    if not args.records:
        args.records = DEFAULT_RECORDS[args.workload]
    if args.workload == 'cfg4':
        # BASELINE.json configs[3]: phylum/genus/species with multi-hit LCA,
        # 64 samples sharded 8 per GPU (pass --mode major for its second run)
        args.ranks, args.samples = 'phylum,genus,species', 8
        if args.mode == 'default':
            args.mode = 'above'
    return args


def host_threads():
    """Host cores this process may use (torchrun exports OMP_NUM_THREADS=1,
    which is not what the CPU arm should be limited to)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_model():
    try:
        with open('/proc/cpuinfo') as f:
            for line in f:
                if line.startswith('model name'):
                    return line.split(':', 1)[1].strip()
    except OSError:
        pass
    return 'unknown'


def measured_traffic(kernel, records):
    """DRAM bytes per launch read from the committed `ncu --set full` capture
    (profiles/traffic.json: measured under ncu once, NOT in this run), valid
    for the same record count only."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            t = json.load(f)[kernel]
        return t['bytes'] if t['records'] == records else None
    except Exception:
        return None


def peaks():
    fp = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(fp) as f:
            return float(json.load(f)['hbm_gbs']), 'measured'
    except Exception:
        return 6650.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index),
                 f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                 '-lms', '20'], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['n/a']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except Exception:
                continue
            for nm, v in zip(names, r[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


class Ctx:
    """One rank of the bench: its GPU, its engine, the process group."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from woltka_b200.engine import Engine
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(self.local)
        self.dev = torch.device('cuda', self.local)
        self.near = None
        if self.world > 1:
            # ranks of one box: keep each rank's pinned host columns and copy
            # threads on its GPU's NUMA node
            from woltka_b200.distributed import bind_near_gpu
            if not os.environ.get('WK_NO_BIND'):
                self.near = bind_near_gpu(self.local)
            dist.init_process_group('nccl', device_id=self.dev)
        self.opts = [o.split('=', 1) for o in args.opt]
        self.hbm_peak, self.peak_src = peaks()
        self.torch, self.dist = torch, dist

    def engine(self):
        from woltka_b200.engine import Engine
        eng = Engine(self.local)
        eng.set_stream(self.torch.cuda.current_stream().cuda_stream)
        for name, value in self.opts:
            eng.set_option(name, int(value))
        return eng

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        t = self.torch.tensor([ms], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def event(self):
        return self.torch.cuda.Event(enable_timing=True)

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def timed_steps(ctx, steps, warmup, step, classify_only=None, clocks=False):
    """W warm-up steps, then K steps between a barrier + synchronize on both
    sides, device events, max over ranks.  Returns (ms total, ms per kernel
    part, launches delta is the caller's, clocks)."""
    for _ in range(warmup):
        step()
    ctx.barrier()
    sampler = ClockSampler(ctx.local) if clocks and ctx.rank == 0 else None
    if sampler:
        sampler.start()
    k_ev = [(ctx.event(), ctx.event()) for _ in range(steps)]
    t_beg, t_end = ctx.event(), ctx.event()
    t_beg.record()
    for i in range(steps):
        step(k_ev[i])
    t_end.record()
    ctx.barrier()
    ck = sampler.stop() if sampler else None
    ms = ctx.max_over_ranks(t_beg.elapsed_time(t_end))
    k_ms = float(np.mean([a.elapsed_time(b) for a, b in k_ev]))
    return ms, k_ms, ck


def collect(eng, n_samples, NF):
    """Engine results in the canonical form oracle.classify returns."""
    units = eng.fetch_counts()
    cell, strat, den = eng.fetch_overflow()
    NF1 = NF + 1
    overflow = []
    for c, t, d in zip(cell.tolist(), strat.tolist(), den.tolist()):
        es, f = divmod(c, NF1)
        overflow.append((es // n_samples, es % n_samples, t, f, d))
    overflow.sort()
    e, sm, st, f, u = eng.fetch_strata()
    return units, overflow, (e, sm, st, f, u)


def oracle_classify(case, entries, flags, q, s, threads, **kw):
    """The CPU oracle (checker / CPU baseline only)."""
    from oracle import oracle as O
    kinds, _, trk = case.tables(entries)
    t0 = time.perf_counter()
    out = O.classify(q, s, parent=case.ft.parent, node_rank=case.ft.node_rank,
                     root=0, sub_node=case.sub_node, sub_feat=case.sub_feat,
                     kinds=kinds, target_rank=trk, flags=flags, major_th=0.8,
                     n_features=case.NF, n_threads=threads, **kw)
    return out, time.perf_counter() - t0


def python_port_rate(case, entries, flags, q, s, m=200_000):
    """Records/s of the pure-Python restatement (oracle/pyport.py: str / set
    / dict like the reference) on the first m records of the batch."""
    from oracle import pyport
    tax = case.tax
    ids = tax.ids()
    tree = {ids[i]: ids[tax.parent[i]] for i in range(tax.T)}
    rankdic = {ids[i]: tax.rank_names[tax.node_rank[i]]
               for i in range(tax.T) if tax.node_rank[i] >= 0}
    m = min(m, len(q))
    while 0 < m < len(q) and q[m] == q[m - 1]:
        m += 1
    gid = [tax.genome_id(g) for g in range(tax.n_genomes)]
    qryque, subque, last = [], [], None
    for qi, si in zip(q[:m].tolist(), s[:m].tolist()):
        if qi != last:
            qryque.append(f'R{qi}')
            subque.append(set())
            last = qi
        subque[-1].add(gid[si])
    chunks = [(qryque[i:i + 1024], subque[i:i + 1024])
              for i in range(0, len(qryque), 1024)]
    kw = dict(uniq=bool(flags & 1), above=bool(flags & 2),
              major=0.8 if flags & 4 else None, unasgd=bool(flags & 8))
    t0 = time.perf_counter()
    pyport.classify_chunks(chunks, entries, tree, rankdic, ids[0],
                           sample='S', **kw)
    dt = time.perf_counter() - t0
    return {'value': m / dt, 'unit': UNIT, 'cores': 1,
            'sample': f'{m} records, oracle/pyport.py (pure Python, the '
                      f'reference\'s data structures)'}


def workload_config(workload, records, entries, mode=None, samples=1):
    if workload == 'cfg5':
        return {'workload': 'cfg5: stratified taxonomy x function: gene '
                            'subjects (10k genomes x 500 genes), rank ko '
                            'through a gene -> KO map (10k KOs, 60 % '
                            'annotated), counts keyed by (genus stratum, KO), '
                            '8 samples per GPU, strata cells merged on rank 0',
                'records_per_gpu': records, 'ranks': entries,
                'subjects': 5_000_000, 'kos': 10_000, 'genera': 3000,
                'samples_per_gpu': 8, 'l2': 'inputs larger than L2'}
    if workload == 'cfg3':
        return {'workload': 'cfg3: coord-match ordinal profile, synthetic '
                            'reads x 5M gene intervals over 1k contigs, '
                            'overlap 80, rank none',
                'records_per_gpu': records, 'genes': 5_000_000,
                'contigs': 1000,
                'l2': 'inputs larger than L2 (2 GB of columns per step)'}
    if workload == 'cfg4':
        return {'workload': 'cfg4: phylum/genus/species with multi-hit LCA '
                            '(--above), synthetic records, 8 samples per GPU',
                'records_per_gpu': records, 'ranks': entries,
                'mode': mode, 'samples_per_gpu': samples,
                'taxonomy_nodes': 21603, 'genomes': 10000,
                'l2': 'inputs larger than L2'}
    return {'workload': 'cfg2: genus-rank taxonomic classify, synthetic SAM '
                        'records x 10k-genome / 21,603-node taxonomy',
            'records_per_gpu': records, 'ranks': entries,
            'mode': mode, 'taxonomy_nodes': 21603, 'genomes': 10000,
            'l2': 'inputs larger than L2 (0.8 GB of columns per step)'}


def base_line(ctx, workload, records, entries, mode, samples, steps, warmup,
              ms, value):
    return {'metric': METRIC, 'value': value, 'unit': UNIT,
            'n_gpus': ctx.world, 'steps': steps, 'warmup': warmup,
            'ms_per_step': ms / steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int32',
            'data': 'synthetic',
            'config': workload_config(workload, records, entries, mode,
                                      samples)}


# ---- cfg2 / cfg4: classify over the taxonomy ---------------------------------
def text_e2e(ctx, case, entries, flags, q, s, n_samples, m=2_000_000):
    """Same plan fed from SAM TEXT in host memory (the form the reference
    reads): wk_parse_text + wk_classify_parsed on the first m records of the
    batch, checked against the column-fed result of the same records."""
    from woltka_b200.engine import Engine, pinned_empty
    m = min(m, len(q))
    while 0 < m < len(q) and q[m] == q[m - 1]:
        m += 1
    q, s = q[:m], s[:m]
    tax = case.tax
    gid = [tax.genome_id(g).encode() for g in range(tax.n_genomes)]
    tail = b'\t1\t42\t150M\t*\t0\t0\t' + b'A' * 50 + b'\t' + b'I' * 50 + b'\n'
    text = b''.join(b'r%d\t0\t%s%s' % (qi, gid[si], tail)
                    for qi, si in zip(q.tolist(), s.tolist()))
    nbytes = len(text)
    ptext = pinned_empty(nbytes, np.uint8)     # the file block, read into pinned memory
    ptext[:] = np.frombuffer(text, dtype=np.uint8)
    text = ptext
    kinds, tab, _ = case.tables(entries)
    eng = Engine(ctx.local)
    eng.set_tree(case.ft.parent, 0)
    eng.set_plan(kinds, flags, 0.8, n_samples, case.NF)
    # subjects get their index in order of appearance: parse once to learn it
    _, _, n_sub, _ = eng.parse_sam(text)
    names = eng.fetch_names(0, 0, n_sub)
    order = np.array([int(x[1:]) for x in names], dtype=np.int64)
    eng.set_subjects(np.ascontiguousarray(tab[:, order]),
                     np.ascontiguousarray(case.sub_node[order]))
    for _ in range(2):
        eng.reset_counts()
        eng.parse_sam(text)
        eng.classify_parsed(None, 0)
        got = eng.fetch_counts()
    t0 = time.perf_counter()
    K = 5
    for _ in range(K):
        eng.reset_counts()
        eng.parse_sam(text)
        eng.classify_parsed(None, 0)
        got = eng.fetch_counts()
    dt = (time.perf_counter() - t0) / K
    ref = Engine(ctx.local)
    ref.set_tree(case.ft.parent, 0)
    ref.set_plan(kinds, flags, 0.8, n_samples, case.NF)
    ref.set_subjects(tab, case.sub_node)
    ref.classify_chunk(q, s, None, None, 0)
    exp = ref.fetch_counts()
    ref.close()
    eng.close()
    return {'value': m / dt, 'unit': UNIT, 'records': int(m),
            'text_bytes': nbytes, 'ms': dt * 1e3,
            'matches_column_fed_result': bool(np.array_equal(got, exp)),
            'api': 'wk_parse_text(host SAM text) + wk_classify_parsed + '
                   'wk_fetch_counts (text in pinned host memory, H2D inside)'}


def bench_classify(ctx, workload, records, ranks, mode, samples, steps, warmup,
                   full, cpu_sample, no_e2e=False, no_cpu=False):
    """cfg2 (one sample per GPU) / cfg4 (8 contiguous samples per GPU)."""
    torch, dist = ctx.torch, ctx.dist
    from woltka_b200 import synth
    from woltka_b200.distributed import merge_engine
    from woltka_b200.engine import pinned_empty
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    n = records
    case = synth.Case(synth.Taxonomy(seed=42))
    entries = ranks.split(',')
    flags = synth.MODES[mode]
    seed = {'cfg2': 1002, 'cfg4': 1004}[workload] + rank
    q, s, qs, nq = synth.gen_hits(n, seed=seed, device=dev, n_samples=samples)
    kinds, tab, _ = case.tables(entries)
    # every rank owns `samples` sample columns of one shared table
    S_loc, S_all = samples, samples * world
    qs = (qs + rank * S_loc).contiguous() if samples > 1 else None
    qs_ptr = qs.data_ptr() if qs is not None else None
    smp = 0 if qs is not None else rank * S_loc
    eng = ctx.engine()
    eng.set_tree(case.ft.parent, 0)
    eng.set_plan(kinds, flags, 0.8, S_all, case.NF)
    eng.set_subjects(tab, case.sub_node)
    counts = eng.counts_tensor()

    def classify_dev():
        eng.classify_device(q.data_ptr(), s.data_ptr(), n, qs_ptr, None, nq,
                            smp)

    def step(ev=None):
        eng.reset_counts()
        if ev:
            ev[0].record()
        classify_dev()
        if ev:
            ev[1].record()
        if world > 1:
            # the merge of a job: ONE reduce of the dense units table.  The
            # sparse part (shares 1/d, d > 16) is empty for this generator
            # (at most 16 hits per query) - checked after the timed region.
            dist.reduce(counts, dst=0)

    l0 = eng.launch_count()
    for _ in range(warmup):
        step()
    l0 = eng.launch_count()
    ms, k_ms, clocks = timed_steps(ctx, steps, 0, step, clocks=full)
    launches = eng.launch_count() - l0
    value = n * world * steps / (ms * 1e-3)
    line = base_line(ctx, workload, n, entries, mode, samples, steps, warmup,
                     ms, value)
    kernel = eng.last_kernel()

    # ---- merged table on hardware: rank 0's table after the merge must be
    # the sum of the tables the ranks computed on their own
    parity_merged = None
    if world > 1:
        eng.reset_counts()
        classify_dev()
        local = eng.fetch_counts()
        n_ovf = len(eng.fetch_overflow()[0])
        eng.reset_counts()
        classify_dev()
        merge_engine(eng, dst=0, dense=True, strata=False)
        merged = eng.fetch_counts() if rank == 0 else None
        parts = [None] * world if rank == 0 else None
        dist.gather_object((local, n_ovf), parts, dst=0)
        if rank == 0:
            total = sum(p[0] for p in parts)
            parity_merged = bool(np.array_equal(total, merged)) and \
                sum(p[1] for p in parts) == len(eng.fetch_overflow()[0])
            assert parity_merged, 'merged table differs from the sum of the ranks'
            # the samples of the ranks are disjoint columns of one table
            assert all(int((p[0] != 0).any(axis=(0, 2)).sum()) <= S_loc
                       for p in parts)
    final_units = None
    if world == 1:
        eng.reset_counts()
        classify_dev()
        final_units = eng.fetch_counts()

    e2e = None
    if full and not no_e2e:
        hq, hs = pinned_empty(n), pinned_empty(n)
        hq[:] = q.cpu().numpy()
        hs[:] = s.cpu().numpy()
        hqs = None
        if qs is not None:
            hqs = pinned_empty(nq)
            hqs[:] = qs.cpu().numpy()
        e2e = {}
        for name, fn in (('soa', lambda: eng.classify_chunk(hq, hs, hqs, None, smp)),
                         ('packed', None)):
            if fn is None:
                if not hasattr(eng, 'classify_packed'):
                    continue
                packed = eng.pack_columns(hq, hs)
                fn = lambda: eng.classify_packed(packed, hqs, smp)  # noqa: E731
            for _ in range(2):
                eng.reset_counts()
                fn()
                res = eng.fetch_counts()
            ctx.barrier()
            t0 = time.perf_counter()
            b0, b1 = ctx.event(), ctx.event()
            b0.record()
            e2e_steps = max(3, min(steps, 10))
            for _ in range(e2e_steps):
                eng.reset_counts()
                fn()                            # H2D inside
                if world > 1:
                    dist.reduce(counts, dst=0)
                res = eng.fetch_counts()        # D2H of the count table
            b1.record()
            ctx.barrier()
            ems = ctx.max_over_ranks(max(b0.elapsed_time(b1),
                                         (time.perf_counter() - t0) * 1e3))
            if final_units is not None:
                assert np.array_equal(res, final_units), 'e2e != device path'
            if name == 'soa':
                h2d = int(2 * 4 * n + (4 * nq if hqs is not None else 0))
                api = 'wk_classify_chunk(host int32 SoA) + wk_fetch_counts'
            else:
                h2d = int(packed.nbytes + (4 * nq if hqs is not None else 0))
                api = ((f'wk_classify_packed_bits(host head bits + '
                        f'{packed.width}-bit subjects)' if packed.stream else
                        'wk_classify_packed(host head bits + uint16 subjects)')
                       + ' + wk_fetch_counts')
            e2e[name] = {'value': n * world * e2e_steps / (ems * 1e-3),
                         'unit': UNIT, 'h2d_bytes_per_step': h2d,
                         'd2h_bytes_per_step': int(res.nbytes),
                         'steps': e2e_steps, 'ms_per_step': ems / e2e_steps,
                         'api': api}
        # the headline is the wire format the host layer (woltka_b200.session)
        # sends: head bits + subjects in ceil(log2 V) bits; the int32 SoA entry point of the
        # north star is reported next to it
        head = dict(e2e['packed'] if 'packed' in e2e else e2e['soa'])
        if 'packed' in e2e:
            head['int32_soa'] = e2e['soa']
        e2e = head

    cpu = parity = None
    if rank == 0 and not no_cpu:
        threads = host_threads()
        m = min(n, cpu_sample)
        qh = q[:m + 64].cpu().numpy()        # cut the sample at a query boundary
        sh = s[:m + 64].cpu().numpy()
        while m < len(qh) and m > 0 and qh[m] == qh[m - 1]:
            m += 1
        qh, sh = qh[:m], sh[:m]
        qsh = qs.cpu().numpy() if qs is not None else None
        kw = dict(n_samples=S_all, q_sample=qsh, sample=smp)
        (eu, eo, _), dt = oracle_classify(case, entries, flags, qh, sh, threads,
                                          **kw)
        e2 = ctx.engine()
        e2.set_tree(case.ft.parent, 0)
        e2.set_plan(kinds, flags, 0.8, S_all, case.NF)
        e2.set_subjects(tab, case.sub_node)
        e2.classify_chunk(qh, sh, qsh, None, smp)
        gu, go, _ = collect(e2, S_all, case.NF)
        e2.close()
        parity = bool(np.array_equal(gu, eu)) and go == eo
        assert parity, 'GPU result differs from the oracle on the sample'
        if full:
            (_, _, _), dt1 = oracle_classify(case, entries, flags, qh[:m // 8],
                                             sh[:m // 8], 1, **kw)
            cpu = {'value': m / dt, 'unit': UNIT, 'cores': threads,
                   'kind': 'port',
                   'sample': f'{m} records of the timed batch, C restatement '
                             f'of the reference path (oracle/woltka_oracle.c), '
                             f'{threads} OpenMP threads',
                   'single_thread_value': (m // 8) / dt1,
                   'python_port': python_port_rate(case, entries, flags, qh, sh)}

    text = None
    if full and rank == 0 and world == 1 and not no_e2e and qs is None:
        text = text_e2e(ctx, case, entries, flags, q[:2_100_000].cpu().numpy(),
                        s[:2_100_000].cpu().numpy(), S_all)

    achieved = 8 * n / (k_ms * 1e-3) / 1e9
    line['roofline'] = {
        'bound': 'hbm', 'achieved': achieved, 'peak': ctx.hbm_peak,
        'unit': 'GB/s', 'frac': achieved / ctx.hbm_peak,
        'traffic': measured_traffic(
            kernel + ':' + ','.join(entries) + ':' + mode, n),
        'traffic_source': 'profiles/traffic.json (ncu capture of the same '
                          'launch, not measured in this run)',
        'kernel': kernel, 'kernel_ms': k_ms, 'peak_source': ctx.peak_src,
        'algorithmic_bytes_per_record': 8}
    line.update({'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': launches,
                 'clocks': clocks, 'parity_on_sample': parity})
    if world > 1:
        line['parity_merged'] = parity_merged
    if ctx.near:
        line['cpus_bound_per_rank'] = len(ctx.near)
    if text is not None:
        line['e2e_from_text'] = text
    eng.close()
    del q, s, qs
    torch.cuda.empty_cache()
    return line


# ---- cfg3: coord-match --------------------------------------------------------
def bench_cfg3(ctx, records, steps, warmup, full, no_e2e=False, no_cpu=False):
    torch, dist = ctx.torch, ctx.dist
    from woltka_b200 import synth
    from woltka_b200._lib import KIND_NONE_ID
    from woltka_b200.engine import pinned_empty
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    n = records
    coff, gb, ge = synth.gen_genes()
    G = len(gb)
    eng = ctx.engine()
    eng.set_plan(np.array([KIND_NONE_ID]), 0, 0.0, 1, G)
    eng.set_subjects(None, None, G)
    eng.ordinal_set_genes(coff, gb, ge, np.arange(G, dtype=np.int32))
    cols = synth.gen_reads(n, seed=1003 + rank, device=dev)
    rq, rc, rb, re_, rl, nq = cols
    ptrs = [x.data_ptr() for x in (rq, rc, rb, re_, rl)]
    counts = eng.counts_tensor()

    def step(ev=None):
        eng.reset_counts()
        if ev:
            ev[0].record()
        eng.ordinal_device(ptrs, n, 0.8)
        if ev:
            ev[1].record()
        if world > 1:
            dist.reduce(counts, dst=0)     # the ranks share one gene table

    for _ in range(warmup):
        step()
    l0 = eng.launch_count()
    ms, k_ms, clocks = timed_steps(ctx, steps, 0, step, clocks=full)
    launches = eng.launch_count() - l0
    value = n * world * steps / (ms * 1e-3)
    line = base_line(ctx, 'cfg3', n, ['none'], None, 1, steps, warmup, ms, value)

    parity_merged = None
    if world > 1:
        eng.reset_counts()
        eng.ordinal_device(ptrs, n, 0.8)
        local = eng.counts_tensor().clone()
        step()
        tot = local.clone()
        dist.reduce(tot, dst=0)
        if rank == 0:
            parity_merged = bool(torch.equal(tot, eng.counts_tensor()))
            assert parity_merged

    e2e = None
    if full and not no_e2e:
        host = []
        for x in (rq, rc, rb, re_, rl):
            h = pinned_empty(n)
            h[:] = x.cpu().numpy()
            host.append(h)
        eng.reset_counts()
        eng.ordinal_chunk(*host, 0.8)
        ctx.barrier()
        t0 = time.perf_counter()
        e2e_steps = 3
        for _ in range(e2e_steps):
            eng.reset_counts()
            eng.ordinal_chunk(*host, 0.8)
            if world > 1:
                dist.reduce(counts, dst=0)
            res = eng.fetch_counts()
        ctx.barrier()
        ems = ctx.max_over_ranks((time.perf_counter() - t0) * 1e3)
        e2e = {'value': n * world * e2e_steps / (ems * 1e-3), 'unit': UNIT,
               'h2d_bytes_per_step': int(5 * 4 * n),
               'd2h_bytes_per_step': int(res.nbytes), 'steps': e2e_steps,
               'ms_per_step': ems / e2e_steps,
               'api': 'wk_ordinal_chunk(host SoA) + wk_fetch_counts'}
        del host

    cpu = parity = None
    if rank == 0 and not no_cpu:
        from oracle import oracle as O
        m = min(n, 1_000_000 if full else 200_000)
        c4 = [x[:m].cpu().numpy() for x in (rc, rb, re_, rl)]
        t0 = time.perf_counter()
        er, eg = O.ordinal_match(*c4, 0.8, coff, gb, ge)
        dt = time.perf_counter() - t0
        e2 = ctx.engine()
        e2.ordinal_set_genes(coff, gb, ge, np.arange(G, dtype=np.int32))
        e2.ordinal_enable_pairs()
        e2.ordinal_chunk(rq[:m].cpu().numpy(), *c4, 0.8)
        r, g = e2.ordinal_pairs()
        e2.close()
        parity = bool(np.array_equal(r, er) and np.array_equal(g, eg))
        assert parity, 'GPU pairs differ from the oracle sweep on the sample'
        cpu = {'value': m / dt, 'unit': UNIT, 'cores': 1, 'kind': 'port',
               'sample': f'{m} reads of the timed batch, sweep matcher '
                         f'(ordinal.match_read_gene restated in C)'}

    alg_bytes = 20 * n + 8 * G
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    line['roofline'] = {
        'bound': 'hbm', 'achieved': achieved, 'peak': ctx.hbm_peak,
        'unit': 'GB/s', 'frac': achieved / ctx.hbm_peak,
        'traffic': measured_traffic('ordinal:cfg3', n),
        'traffic_source': 'profiles/traffic.json (ncu capture, not measured '
                          'in this run)',
        'kernel': (eng.last_kernel() if eng.last_kernel().startswith('ordinal_fused')
                   else 'ordinal_match_kernel+' + eng.last_kernel()),
        'kernel_ms': k_ms, 'peak_source': ctx.peak_src,
        'algorithmic_bytes_per_record': 20, 'algorithmic_bytes_per_gene': 8}
    line.update({'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': launches,
                 'clocks': clocks, 'parity_on_sample': parity})
    if world > 1:
        line['parity_merged'] = parity_merged
    eng.close()
    del rq, rc, rb, re_, rl, cols
    torch.cuda.empty_cache()
    return line


# ---- cfg5: stratified ----------------------------------------------------------
def make_cfg5(n, seed, device, n_genomes=10_000, genes_per=500, n_ko=10_000,
              n_genus=3000, n_samples=8, p=0.48, kmax=16):
    """the second pass of a stratified run.  Subjects are
    genes, the rank is 'ko' through a gene -> KO map (tree.read_map read as a
    two-level tree; 60 % of the genes annotated), every query carries the
    stratum its unique genus assignment of the first pass gave it (80 %
    assigned, classify.counter_strat skips the others)."""
    import torch
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    q_est = int(n / 1.8) + 1024
    u = torch.rand(q_est, generator=g, device=dev, dtype=torch.float64)
    k = torch.floor(torch.log1p(-u) / np.log(1.0 - p)).to(torch.int64) + 1
    k.clamp_(1, kmax)
    csum = torch.cumsum(k, 0)
    nq = int(torch.searchsorted(csum, torch.tensor(n, device=dev)).item()) + 1
    k = k[:nq].clone()
    k[-1] -= csum[nq - 1] - n
    q = torch.repeat_interleave(torch.arange(nq, device=dev, dtype=torch.int32), k)
    first = torch.randint(0, n_genomes, (nq,), generator=g, device=dev)
    genome = (first[q.long()] + torch.randint(0, 3, (n,), generator=g, device=dev)) % n_genomes
    s = (genome * genes_per + torch.randint(0, genes_per, (n,), generator=g,
                                            device=dev)).to(torch.int32)
    del genome
    # genera own contiguous blocks of genomes (like the taxonomy generator)
    genus_of = (torch.arange(n_genomes, device=dev) * n_genus // n_genomes)
    strat = torch.where(torch.rand(nq, generator=g, device=dev) < 0.8,
                        genus_of[first], torch.full_like(first, -1)).to(torch.int32)
    q_sample = (torch.arange(nq, device=dev, dtype=torch.int64) * n_samples //
                nq).to(torch.int32)
    gk = torch.Generator(device='cpu')
    gk.manual_seed(1005)
    V = n_genomes * genes_per
    ko = torch.randint(1, n_ko + 1, (V,), generator=gk)
    ko[torch.rand(V, generator=gk) >= 0.6] = -1
    return q, s, q_sample, strat, nq, ko.to(torch.int32).numpy()[None, :], n_ko


def cells_checksum(torch, keys, units):
    """(cells, total units, sum of mix(key) * units mod 2^64) of a strata
    table: additive under a merge by key, so the merged table of N ranks must
    carry the sum of their checksums."""
    if not keys.numel():
        return [0, 0, 0]
    h = keys * -7046029254386353131         # 0x9E3779B97F4A7C15 as int64
    h = h ^ (h >> 29)
    return [int(keys.numel()), int(units.sum().item()),
            int((h * units).sum().item())]


def bench_cfg5(ctx, records, steps, warmup, full, no_e2e=False, no_cpu=False):
    torch, dist = ctx.torch, ctx.dist
    from woltka_b200._lib import KIND_RANK
    from woltka_b200.distributed import merge_engine, reduce_scatter_strata
    from woltka_b200.engine import pinned_empty
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    n = records
    S_loc = 8
    S_all = S_loc * world
    q, s, qs, qt, nq, tab, n_ko = make_cfg5(n, 1005 + rank, dev, n_samples=S_loc)
    qs = (qs + rank * S_loc).contiguous()     # this rank's samples of the shared table
    T = 1 + n_ko                      # root + KOs; the genes are the subjects
    parent = np.zeros(T, dtype=np.int32)
    eng = ctx.engine()
    eng.set_tree(parent, 0)
    eng.set_plan(np.array([KIND_RANK], dtype=np.int32), 0, 0.0, S_all, T)
    eng.set_subjects(tab, None)
    ptr = (q.data_ptr(), s.data_ptr(), qs.data_ptr(), qt.data_ptr())

    def classify_dev():
        eng.classify_device(ptr[0], ptr[1], n, ptr[2], ptr[3], nq, 0)

    def step(ev=None):
        # one job: empty table -> classify -> strata cells of every rank
        # merged by key: one all-to-all, every rank ends up with the sum of
        # the cells it owns (reduce-scatter of the sparse table); the overflow
        # list goes to rank 0
        eng.reset_counts()
        if ev:
            ev[0].record()
        classify_dev()
        if ev:
            ev[1].record()
        if world > 1:
            reduce_scatter_strata(eng)
            merge_engine(eng, dst=0, dense=False, strata=False)

    for _ in range(max(warmup, 1)):
        step()
    l0 = eng.launch_count()
    ms, k_ms, clocks = timed_steps(ctx, steps, 0, step, clocks=full)
    launches = eng.launch_count() - l0
    value = n * world * steps / (ms * 1e-3)
    line = base_line(ctx, 'cfg5', n, ['ko'], None, S_loc, steps, warmup, ms,
                     value)

    # merged strata table: checksum of checksums (additive under merge by key)
    parity_merged = None
    if world > 1:
        eng.reset_counts()
        classify_dev()
        k_, u_ = eng.strata_export()
        mine = torch.tensor(cells_checksum(torch, k_, u_), device=dev,
                            dtype=torch.int64)
        step()
        tot = mine.clone()
        dist.reduce(tot, dst=0)               # int64 wrap-around = mod 2^64
        k_, u_ = eng.strata_export()          # the cells this rank owns now
        got = torch.tensor(cells_checksum(torch, k_, u_), device=dev,
                           dtype=torch.int64)
        dist.reduce(got, dst=0)
        if rank == 0:
            # samples are disjoint over the ranks, so even the cell counts add up
            parity_merged = got.tolist() == tot.tolist()
            assert parity_merged, ('merged strata table', got.tolist(),
                                   tot.tolist())
    eng.reset_counts()
    classify_dev()
    cells = int(eng.strata_export()[0].numel())

    e2e = None
    if full and not no_e2e:
        host = []
        for x, m in ((q, n), (s, n), (qs, nq), (qt, nq)):
            h = pinned_empty(m)
            h[:] = x.cpu().numpy()
            host.append(h)
        eng.reset_counts()
        eng.classify_chunk(*host)
        eng.fetch_strata()          # (the engine pins its result buffers once)
        ctx.barrier()
        t0 = time.perf_counter()
        e2e_steps = 3
        for _ in range(e2e_steps):
            eng.reset_counts()
            eng.classify_chunk(*host)
            if world > 1:
                merge_engine(eng, dst=0, dense=False, strata=True)
            nc = len(eng.fetch_strata()[0]) if rank == 0 else 0
        ctx.barrier()
        ems = ctx.max_over_ranks((time.perf_counter() - t0) * 1e3)
        e2e = {'value': n * world * e2e_steps / (ems * 1e-3), 'unit': UNIT,
               'h2d_bytes_per_step': int(8 * n + 8 * nq),
               'd2h_bytes_per_step': int(nc * 28), 'steps': e2e_steps,
               'ms_per_step': ems / e2e_steps,
               'api': 'wk_classify_chunk(host SoA + per-query sample and '
                      'stratum) + wk_fetch_strata'}
        del host

    cpu = parity = None
    if rank == 0 and not no_cpu:
        from oracle import oracle as O
        m = min(n, 5_000_000 if full else 1_000_000)
        qh = q[:m + 64].cpu().numpy()
        while m < len(qh) and qh[m] == qh[m - 1]:
            m += 1
        qh, sh = qh[:m], s[:m].cpu().numpy()
        mq = int(qh[-1]) + 1
        qsh, qth = qs[:mq].cpu().numpy(), qt[:mq].cpu().numpy()
        node_rank = np.zeros(T, dtype=np.int32)
        node_rank[0] = -1
        # the oracle walks the tree: gene -> KO node (or none) as sub_node
        sub_node = tab[0].astype(np.int32)
        threads = host_threads()
        t0 = time.perf_counter()
        exp = O.classify(qh, sh, parent=parent, node_rank=node_rank, root=0,
                         sub_node=sub_node, sub_feat=None,
                         kinds=np.array([KIND_RANK], dtype=np.int32),
                         target_rank=[0], flags=0, n_samples=S_all,
                         n_features=T, q_sample=qsh, q_stratum=qth,
                         n_threads=threads)
        dt = time.perf_counter() - t0
        e2 = ctx.engine()
        e2.set_tree(parent, 0)
        e2.set_plan(np.array([KIND_RANK], dtype=np.int32), 0, 0.0, S_all, T)
        e2.set_subjects(tab, None)
        e2.classify_chunk(qh, sh, qsh, qth, 0)
        gu, go, (e_, s_, t_, f_, u_) = collect(e2, S_all, T)
        e2.close()
        got = dict(zip(zip(e_.tolist(), s_.tolist(), t_.tolist(), f_.tolist()),
                       u_.tolist()))
        parity = bool(np.array_equal(gu, exp[0]) and go == exp[1] and
                      got == exp[2])
        assert parity, 'GPU strata cells differ from the oracle on the sample'
        cpu = {'value': m / dt, 'unit': UNIT, 'cores': threads, 'kind': 'port',
               'sample': f'{m} records of the timed batch, C restatement of '
                         f'the reference path (oracle/woltka_oracle.c), '
                         f'{threads} OpenMP threads'}

    achieved = 8 * n / (k_ms * 1e-3) / 1e9
    line['roofline'] = {
        'bound': 'hbm', 'achieved': achieved, 'peak': ctx.hbm_peak,
        'unit': 'GB/s', 'frac': achieved / ctx.hbm_peak,
        'traffic': measured_traffic('classify_strata_kernel:cfg5', n),
        'traffic_source': 'profiles/traffic.json (ncu capture, not measured '
                          'in this run)',
        'kernel': eng.last_kernel(), 'kernel_ms': k_ms,
        'peak_source': ctx.peak_src, 'algorithmic_bytes_per_record': 8}
    line.update({'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': launches,
                 'clocks': clocks, 'parity_on_sample': parity,
                 'strata_cells_per_gpu': cells})
    if world > 1:
        line['parity_merged'] = parity_merged
    eng.close()
    del q, s, qs, qt
    torch.cuda.empty_cache()
    return line


def brief(line):
    """What an `extra` entry keeps of a sub-run's line."""
    keep = ('value', 'unit', 'n_gpus', 'steps', 'ms_per_step', 'config',
            'roofline', 'parity_on_sample', 'parity_merged', 'gpu_launches',
            'strata_cells_per_gpu')
    return {k: line[k] for k in keep if k in line}


def run_ours(args):
    ctx = Ctx(args)
    wl, n = args.workload, args.records
    kw = dict(no_e2e=args.no_e2e, no_cpu=args.no_cpu)
    if wl == 'cfg3':
        line = bench_cfg3(ctx, n, args.steps, args.warmup, True, **kw)
    elif wl == 'cfg5':
        line = bench_cfg5(ctx, n, args.steps, args.warmup, True, **kw)
    else:
        line = bench_classify(ctx, wl, n, args.ranks, args.mode, args.samples,
                              args.steps, args.warmup, True, args.cpu_sample,
                              **kw)
        default = (wl == 'cfg2' and args.ranks == 'genus' and
                   args.mode == 'default' and args.samples == 1 and
                   not args.explicit_records)
        if default and not args.no_extra:
            # the other BASELINE.json configs at the same number of GPUs, short
            xs, xw = max(3, min(args.steps, 5)), 3
            extra = {}
            extra['cfg4'] = brief(bench_classify(
                ctx, 'cfg4', DEFAULT_RECORDS['cfg4'], 'phylum,genus,species',
                'above', 8, xs, xw, False, 2_000_000, no_cpu=args.no_cpu))
            extra['cfg4_major80'] = brief(bench_classify(
                ctx, 'cfg4', DEFAULT_RECORDS['cfg4'], 'phylum,genus,species',
                'major', 8, xs, xw, False, 2_000_000, no_cpu=args.no_cpu))
            extra['cfg5'] = brief(bench_cfg5(
                ctx, DEFAULT_RECORDS['cfg5'], xs, xw, False,
                no_cpu=args.no_cpu))
            extra['cfg3'] = brief(bench_cfg3(
                ctx, DEFAULT_RECORDS['cfg3'], xs, xw, False,
                no_cpu=args.no_cpu))
            line['extra'] = extra
    if ctx.rank == 0:
        print(json.dumps(line))
    ctx.close()


# ---- reference arm ---------------------------------------------------------------
def run_reference(args):
    """CPU arm on rank 0: the reference's own `woltka.workflow.classify()`
    (baseline/_ref or /root/reference, when importable) on the box's host
    cores - single process and one process per core with a dict merge, its
    documented scale-out (doc/perform.md:70-92) - else the C port."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from baseline import reference_arm
    print(json.dumps(reference_arm.run(args, METRIC, UNIT, workload_config,
                                       host_threads(), cpu_model())))


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
