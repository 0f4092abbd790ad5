#!/usr/bin/env python
"""Benchmark of the classify hot path (BASELINE.json metric: alignment
records classified per second; achieved HBM GB/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N
    python bench.py --impl reference        # CPU arm (oracle port, all cores)

A step is one pass of the hot path over one batch of synthetic records
(SURVEY.md §8d generator).  `value` is timed with the int32 SoA columns
already resident in HBM; `e2e` goes through the reference-facing C-ABI call
with pinned HOST buffers (H2D of the columns and D2H of the count table inside
the timed region).  Every rank works on its own batch (weak scaling); the
per-rank count tables are merged by one NCCL reduce to rank 0 per step.
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg2', choices=['cfg2', 'cfg3', 'cfg4', 'cfg5'])
    ap.add_argument('--records', type=int, default=100_000_000)
    ap.add_argument('--mode', default='default',
                    choices=['default', 'major', 'uniq', 'above'])
    ap.add_argument('--ranks', default='genus')
    ap.add_argument('--cpu-sample', type=int, default=20_000_000)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--samples', type=int, default=1,
                    help='samples per GPU (per-query sample column if > 1)')
    args = ap.parse_args()
    if args.workload == 'cfg4':
        # BASELINE.json configs[3]: phylum/genus/species with multi-hit LCA,
        # 64 samples sharded 8 per GPU
        args.ranks, args.mode, args.samples = 'phylum,genus,species', 'above', 8
    return args


def host_threads():
    """Host cores this process may use (torchrun exports OMP_NUM_THREADS=1,
    which is not what the CPU arm should be limited to)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_traffic(kernel, records):
    """DRAM bytes per launch from the committed `ncu --set full` capture
    (profiles/traffic.json), valid for the same record count only."""
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


# ---- workloads ---------------------------------------------------------------
def make_cfg2(args, device, seed):
    """genus-rank (or --ranks) classify over the 21,603-node taxonomy."""
    from tests import cases
    from woltka_b200 import synth
    case = cases.Case(synth.Taxonomy(seed=42))
    entries = args.ranks.split(',')
    flags = cases.MODES[args.mode]
    q, s, qs, nq = synth.gen_hits(args.records, seed=seed, device=device,
                                  n_samples=args.samples)
    args._q_sample = qs if args.samples > 1 else None
    return case, entries, flags, q, s, nq


def cpu_classify(case, entries, flags, q, s, threads, **kw):
    from tests import cases
    t0 = time.perf_counter()
    out = cases.run_oracle(case, entries, flags, 0.8, q, s, n_threads=threads,
                           **kw)
    return out, time.perf_counter() - t0


def python_port_rate(case, entries, flags, q, s, m=200_000):
    """Records/s of the pure-Python restatement (oracle/pyport.py: str / set
    / dict like the reference) on the first m records of the batch."""
    from oracle import pyport
    from tests import cases as C
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


def text_e2e(case, entries, flags, q, s, n_samples, m=2_000_000):
    """Same plan fed from SAM TEXT in host memory (the form the reference
    reads): wk_parse_text + wk_classify_parsed on the first m records of the
    batch, checked against the column-fed result of the same records."""
    from woltka_b200.engine import Engine
    from tests import cases as C
    m = min(m, len(q))
    while 0 < m < len(q) and q[m] == q[m - 1]:
        m += 1
    q, s = q[:m], s[:m]
    tax = case.tax
    gid = [tax.genome_id(g).encode() for g in range(tax.n_genomes)]
    tail = b'\t1\t42\t150M\t*\t0\t0\t' + b'A' * 50 + b'\t' + b'I' * 50 + b'\n'
    text = b''.join(b'r%d\t0\t%s%s' % (qi, gid[si], tail)
                    for qi, si in zip(q.tolist(), s.tolist()))
    from woltka_b200.engine import pinned_empty
    nbytes = len(text)
    ptext = pinned_empty(nbytes, np.uint8)     # the file block, read into pinned memory
    ptext[:] = np.frombuffer(text, dtype=np.uint8)
    text = ptext
    kinds, tab, _ = case.tables(entries)
    eng = Engine(0)
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
    ref = Engine(0)
    exp = C.run_engine(ref, case, entries, flags, 0.8, q, s, n_samples=n_samples)[0]
    ref.close()
    eng.close()
    return {'value': m / dt, 'unit': UNIT, 'records': int(m),
            'text_bytes': nbytes, 'ms': dt * 1e3,
            'matches_column_fed_result': bool(np.array_equal(got, exp)),
            'api': 'wk_parse_text(host SAM text) + wk_classify_parsed + '
                   'wk_fetch_counts (text in pinned host memory, H2D inside)'}


def run_reference(args):
    """CPU arm: the oracle port of the reference's path, all host threads."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import oracle as O
    threads = host_threads()
    n = min(args.records, args.cpu_sample)
    if args.workload == 'cfg3':
        return run_reference_cfg3(args, threads)
    if args.workload == 'cfg5':
        return run_reference_cfg5(args, threads)
    args_records = args.records
    args.records = n
    case, entries, flags, q, s, nq = make_cfg2(args, 'cpu', 1002)
    args.records = args_records
    q, s = q.numpy(), s.numpy()
    kw = {}
    if args._q_sample is not None:
        kw = dict(n_samples=args.samples, q_sample=args._q_sample.numpy())
    for _ in range(args.warmup):
        cpu_classify(case, entries, flags, q, s, threads, **kw)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_classify(case, entries, flags, q, s, threads, **kw)
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int32',
        'data': 'synthetic',
        'config': workload_config(args, entries),
        'cpu_baseline': {
            'value': val, 'unit': UNIT, 'cores': threads, 'kind': 'port',
            'sample': f'{n} records of the same generator per step '
                      f'(C restatement oracle/woltka_oracle.c, OpenMP)',
            'python_port': python_port_rate(case, entries, flags, q, s)},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def run_reference_cfg5(args, threads):
    """CPU arm of cfg5: the C port over a bounded sample (the strata table is
    a hash on the CPU too)."""
    from oracle import oracle as O
    from woltka_b200._lib import KIND_RANK
    n = min(args.records, 5_000_000)
    q, s, qs, qt, nq, tab, n_ko = make_cfg5(n, 1005, 'cpu')
    T = 1 + n_ko
    parent = np.zeros(T, dtype=np.int32)
    node_rank = np.zeros(T, dtype=np.int32)
    node_rank[0] = -1
    kw = dict(parent=parent, node_rank=node_rank, root=0,
              sub_node=tab[0].astype(np.int32), sub_feat=None,
              kinds=np.array([KIND_RANK], dtype=np.int32), target_rank=[0],
              flags=0, n_samples=8, n_features=T, q_sample=qs.numpy(),
              q_stratum=qt.numpy(), n_threads=threads)
    q, s = q.numpy(), s.numpy()
    steps = max(1, min(args.steps, 3))
    O.classify(q, s, **kw)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.classify(q, s, **kw)
    dt = time.perf_counter() - t0
    val = n * steps / dt
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': steps, 'warmup': 1,
        'ms_per_step': dt / steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int32',
        'data': 'synthetic', 'config': workload_config(args, ['ko']),
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': threads,
                         'kind': 'port',
                         'sample': f'{n} records of the same generator per '
                                   f'step (C restatement, OpenMP)'},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


def run_reference_cfg3(args, threads):
    from oracle import oracle as O
    from woltka_b200 import synth
    n = min(args.records, 1_000_000)
    coff, gb, ge = synth.gen_genes()
    rq, rc, rb, re_, rl, nq = synth.gen_reads(n, seed=1003)
    cols = [x.numpy() for x in (rc, rb, re_, rl)]
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.ordinal_match(*cols, 0.8, coff, gb, ge)
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': 0,
        'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int32',
        'data': 'synthetic', 'config': workload_config(args, ['none']),
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': 1,
                         'kind': 'port',
                         'sample': f'{n} reads, sweep matcher only'},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0}, 'gpu_launches': 0}))


def workload_config(args, entries):
    if args.workload == 'cfg5':
        return {'workload': 'cfg5: stratified taxonomy x function: gene '
                            'subjects (10k genomes x 500 genes), rank ko '
                            'through a gene -> KO map (10k KOs, 60 % '
                            'annotated), counts keyed by (genus stratum, KO), '
                            '8 samples per GPU',
                'records_per_gpu': args.records, 'ranks': entries,
                'subjects': 5_000_000, 'kos': 10_000, 'genera': 3000,
                'samples_per_gpu': 8, 'l2': 'inputs larger than L2'}
    if args.workload == 'cfg3':
        return {'workload': 'cfg3: coord-match ordinal profile, synthetic '
                            'reads x 5M gene intervals over 1k contigs, '
                            'overlap 80, rank none',
                'records_per_gpu': args.records, 'genes': 5_000_000,
                'contigs': 1000,
                'l2': 'inputs larger than L2 (2 GB of columns per step)'}
    if args.workload == 'cfg4':
        return {'workload': 'cfg4: phylum/genus/species with multi-hit LCA '
                            '(--above), synthetic records, 8 samples per GPU',
                'records_per_gpu': args.records, 'ranks': entries,
                'mode': args.mode, 'samples_per_gpu': args.samples,
                'taxonomy_nodes': 21603, 'genomes': 10000,
                'l2': 'inputs larger than L2'}
    return {'workload': 'cfg2: genus-rank taxonomic classify, synthetic SAM '
                        'records x 10k-genome / 21,603-node taxonomy',
            'records_per_gpu': args.records, 'ranks': entries,
            'mode': args.mode, 'taxonomy_nodes': 21603, 'genomes': 10000,
            'l2': 'inputs larger than L2 (0.8 GB of columns per step)'}


# ---- our arm -----------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from woltka_b200.engine import Engine, pinned_empty
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    near = None
    if world > 1:
        # ranks of one box: keep each rank's pinned host columns and copy
        # threads on its GPU's NUMA node (N = 1 keeps every core for the CPU arm)
        from woltka_b200.distributed import bind_near_gpu
        if not os.environ.get('WK_NO_BIND'):
            near = bind_near_gpu(local)
        dist.init_process_group('nccl', device_id=dev)
    args._near_cpus = len(near) if near else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng = Engine(local)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    n = args.records
    hbm_peak, peak_src = peaks()

    if args.workload == 'cfg3':
        return run_ours_cfg3(args, eng, dev, world, rank, barrier, hbm_peak,
                             peak_src)
    if args.workload == 'cfg5':
        return run_ours_cfg5(args, eng, dev, world, rank, barrier, hbm_peak,
                             peak_src)

    from tests import cases
    case, entries, flags, q, s, nq = make_cfg2(args, dev, 1002 + rank)
    kinds, tab, _ = case.tables(entries)
    # every rank owns `samples` sample columns of one shared table
    S_loc = args.samples
    S_all = S_loc * world
    qs = args._q_sample
    if qs is not None:
        qs = (qs + rank * S_loc).contiguous()
    qs_ptr = qs.data_ptr() if qs is not None else None
    smp = 0 if qs is not None else rank * S_loc
    eng.set_tree(case.ft.parent, 0)
    eng.set_plan(kinds, flags, 0.8, S_all, case.NF)
    eng.set_subjects(tab, case.sub_node)
    counts = eng.counts_tensor()
    bytes_per_rec = 8

    def classify_dev():
        eng.classify_device(q.data_ptr(), s.data_ptr(), n, qs_ptr, None, nq,
                            smp)

    def step():
        eng.reset_counts()
        classify_dev()
        if world > 1:
            dist.reduce(counts, dst=0)

    for _ in range(args.warmup):
        step()
    barrier()
    k_ev = [(torch.cuda.Event(enable_timing=True),
             torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = eng.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_beg = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_beg.record()
    for i in range(args.steps):
        eng.reset_counts()
        k_ev[i][0].record()
        classify_dev()
        k_ev[i][1].record()
        if world > 1:
            dist.reduce(counts, dst=0)
    t_end.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = t_beg.elapsed_time(t_end)
    k_ms = float(np.mean([a.elapsed_time(b) for a, b in k_ev]))
    launches = eng.launch_count() - l0
    tms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    value = n * world * args.steps / (ms * 1e-3)
    # parity of the timed configuration on a bounded sample + CPU baseline
    final_units = eng.fetch_counts() if world == 1 else None

    e2e = None
    if not args.no_e2e:
        hq = pinned_empty(n)
        hs = pinned_empty(n)
        hq[:] = q.cpu().numpy()
        hs[:] = s.cpu().numpy()
        hqs = None
        if qs is not None:
            hqs = pinned_empty(nq)
            hqs[:] = qs.cpu().numpy()
        for _ in range(2):
            eng.reset_counts()
            eng.classify_chunk(hq, hs, hqs, None, smp)
            res = eng.fetch_counts()
        barrier()
        t0 = time.perf_counter()
        b0 = torch.cuda.Event(enable_timing=True)
        b1 = torch.cuda.Event(enable_timing=True)
        b0.record()
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(e2e_steps):
            eng.reset_counts()
            eng.classify_chunk(hq, hs, hqs, None, smp)      # H2D inside
            if world > 1:
                dist.reduce(counts, dst=0)
            res = eng.fetch_counts()        # D2H of the count table
        b1.record()
        barrier()
        ems = max(b0.elapsed_time(b1), (time.perf_counter() - t0) * 1e3)
        tms = torch.tensor([ems], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ems = float(tms.item())
        e2e = {'value': n * world * e2e_steps / (ems * 1e-3), 'unit': UNIT,
               'h2d_bytes_per_step': int(2 * 4 * n + (4 * nq if hqs is not
                                                      None else 0)),
               'd2h_bytes_per_step': int(res.nbytes),
               'steps': e2e_steps, 'ms_per_step': ems / e2e_steps,
               'api': 'wk_classify_chunk(host SoA) + wk_fetch_counts'}
        if final_units is not None:
            assert np.array_equal(res, final_units), 'e2e != device path'

    cpu = None
    parity = None
    if rank == 0 and not args.no_cpu:
        from oracle import oracle as O
        threads = host_threads()
        m = min(n, args.cpu_sample)
        # cut the sample at a query boundary
        qh = q[:m + 64].cpu().numpy()
        sh = s[:m + 64].cpu().numpy()
        while m < len(qh) and m > 0 and qh[m] == qh[m - 1]:
            m += 1
        qh, sh = qh[:m], sh[:m]
        qsh = qs.cpu().numpy() if qs is not None else None
        kw = dict(n_samples=S_all, q_sample=qsh, sample=smp)
        (eu, eo, _), dt = cpu_classify(case, entries, flags, qh, sh, threads,
                                       **kw)
        (_, _, _), dt1 = cpu_classify(case, entries, flags, qh[:m // 8],
                                      sh[:m // 8], 1, **kw)
        eng.reset_counts()
        eng.classify_chunk(qh, sh, qsh, None, smp)
        gu, go, _ = cases.collect(eng, S_all, case.NF)
        parity = bool(np.array_equal(gu, eu)) and go == eo
        cpu = {'value': m / dt, 'unit': UNIT, 'cores': threads,
               'kind': 'port',
               'sample': f'{m} records of the timed batch, C restatement '
                         f'of the reference path (oracle/woltka_oracle.c), '
                         f'{threads} OpenMP threads',
               'single_thread_value': (m // 8) / dt1,
               'python_port': python_port_rate(case, entries, flags, qh, sh)}
        assert parity, 'GPU result differs from the oracle on the sample'

    text = None
    if rank == 0 and world == 1 and not args.no_e2e and qs is None:
        text = text_e2e(case, entries, flags, q[:2_100_000].cpu().numpy(),
                        s[:2_100_000].cpu().numpy(), S_all)

    if rank == 0:
        achieved = bytes_per_rec * n / (k_ms * 1e-3) / 1e9
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int32',
            'data': 'synthetic', 'config': workload_config(args, entries),
            'roofline': {'bound': 'hbm', 'achieved': achieved,
                         'peak': hbm_peak, 'unit': 'GB/s',
                         'frac': achieved / hbm_peak,
                         'traffic': measured_traffic(
                             eng.last_kernel() + ':' + ','.join(entries) +
                             ':' + args.mode, n),
                         'kernel': eng.last_kernel(),
                         'kernel_ms': k_ms, 'peak_source': peak_src,
                         'algorithmic_bytes_per_record': bytes_per_rec},
            'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': launches,
            'clocks': clocks, 'parity_on_sample': parity,
        }
        if args._near_cpus:
            line['cpus_bound_per_rank'] = args._near_cpus
        if text is not None:
            line['e2e_from_text'] = text
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_ours_cfg3(args, eng, dev, world, rank, barrier, hbm_peak, peak_src):
    import torch
    import torch.distributed as dist
    from woltka_b200 import synth
    from woltka_b200._lib import KIND_NONE_ID
    from woltka_b200.engine import pinned_empty
    n = args.records
    coff, gb, ge = synth.gen_genes()
    G = len(gb)
    eng.set_plan(np.array([KIND_NONE_ID]), 0, 0.0, 1, G)
    eng.set_subjects(None, None, G)
    eng.ordinal_set_genes(coff, gb, ge, np.arange(G, dtype=np.int32))
    cols = synth.gen_reads(n, seed=1003 + rank, device=dev)
    rq, rc, rb, re_, rl, nq = cols
    ptrs = [x.data_ptr() for x in (rq, rc, rb, re_, rl)]
    counts = eng.counts_tensor()

    def step():
        eng.reset_counts()
        eng.ordinal_device(ptrs, n, 0.8)
        if world > 1:
            dist.reduce(counts, dst=0)

    for _ in range(args.warmup):
        step()
    barrier()
    l0 = eng.launch_count()
    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()
    t_beg = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    k_ev = [(torch.cuda.Event(enable_timing=True),
             torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_beg.record()
    for i in range(args.steps):
        eng.reset_counts()
        k_ev[i][0].record()
        eng.ordinal_device(ptrs, n, 0.8)
        k_ev[i][1].record()
        if world > 1:
            dist.reduce(counts, dst=0)
    t_end.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = t_beg.elapsed_time(t_end)
    k_ms = float(np.mean([a.elapsed_time(b) for a, b in k_ev]))
    launches = eng.launch_count() - l0
    tms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    value = n * world * args.steps / (ms * 1e-3)

    e2e = None
    if not args.no_e2e:
        host = []
        for x in (rq, rc, rb, re_, rl):
            h = pinned_empty(n)
            h[:] = x.cpu().numpy()
            host.append(h)
        eng.reset_counts()
        eng.ordinal_chunk(*host, 0.8)
        barrier()
        t0 = time.perf_counter()
        e2e_steps = 3
        for _ in range(e2e_steps):
            eng.reset_counts()
            eng.ordinal_chunk(*host, 0.8)
            if world > 1:
                dist.reduce(counts, dst=0)
            res = eng.fetch_counts()
        barrier()
        ems = (time.perf_counter() - t0) * 1e3
        e2e = {'value': n * world * e2e_steps / (ems * 1e-3), 'unit': UNIT,
               'h2d_bytes_per_step': int(5 * 4 * n),
               'd2h_bytes_per_step': int(res.nbytes), 'steps': e2e_steps,
               'ms_per_step': ems / e2e_steps,
               'api': 'wk_ordinal_chunk(host SoA) + wk_fetch_counts'}

    cpu = None
    parity = None
    if rank == 0 and not args.no_cpu:
        from oracle import oracle as O
        m = min(n, 1_000_000)
        c4 = [x[:m].cpu().numpy() for x in (rc, rb, re_, rl)]
        t0 = time.perf_counter()
        er, eg = O.ordinal_match(*c4, 0.8, coff, gb, ge)
        dt = time.perf_counter() - t0
        eng.ordinal_enable_pairs()
        eng.reset_counts()
        eng.ordinal_chunk(rq[:m].cpu().numpy(), *c4, 0.8)
        r, g = eng.ordinal_pairs()
        parity = bool(np.array_equal(r, er) and np.array_equal(g, eg))
        cpu = {'value': m / dt, 'unit': UNIT, 'cores': 1, 'kind': 'port',
               'sample': f'{m} reads of the timed batch, sweep matcher '
                         f'(ordinal.match_read_gene restated in C)'}
        assert parity, 'GPU pairs differ from the oracle sweep on the sample'

    if rank == 0:
        alg_bytes = 20 * n + 8 * G
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        print(json.dumps({
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int32',
            'data': 'synthetic', 'config': workload_config(args, ['none']),
            'roofline': {'bound': 'hbm', 'achieved': achieved,
                         'peak': hbm_peak, 'unit': 'GB/s',
                         'frac': achieved / hbm_peak,
                         'traffic': measured_traffic('ordinal:cfg3', n),
                         'kernel': 'ordinal_match_kernel+' + eng.last_kernel(),
                         'kernel_ms': k_ms, 'peak_source': peak_src,
                         'algorithmic_bytes_per_record': 20,
                         'algorithmic_bytes_per_gene': 8},
            'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': launches,
            'clocks': clocks, 'parity_on_sample': parity}))
    if world > 1:
        dist.destroy_process_group()


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


def run_ours_cfg5(args, eng, dev, world, rank, barrier, hbm_peak, peak_src):
    import torch
    import torch.distributed as dist
    from woltka_b200._lib import KIND_RANK
    from woltka_b200.engine import pinned_empty
    n = args.records
    S_loc = 8
    q, s, qs, qt, nq, tab, n_ko = make_cfg5(n, 1005 + rank, dev, n_samples=S_loc)
    T = 1 + n_ko                      # root + KOs; the genes are the subjects
    parent = np.zeros(T, dtype=np.int32)
    eng.set_tree(parent, 0)
    # samples are sharded: every rank owns its own 8 sample columns, so the
    # (sample, genus, KO) cells of the ranks are disjoint - no collective
    eng.set_plan(np.array([KIND_RANK], dtype=np.int32), 0, 0.0, S_loc, T)
    eng.set_subjects(tab, None)
    ptr = (q.data_ptr(), s.data_ptr(), qs.data_ptr(), qt.data_ptr())

    def classify_dev():
        eng.classify_device(ptr[0], ptr[1], n, ptr[2], ptr[3], nq, 0)

    # the strata table keeps growing over the steps like over the chunks of a
    # run (same keys every step); it is not cleared inside the timed region
    for _ in range(max(args.warmup, 1)):
        classify_dev()
    barrier()
    l0 = eng.launch_count()
    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()
    t_beg = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_beg.record()
    for i in range(args.steps):
        classify_dev()
    t_end.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = t_beg.elapsed_time(t_end)
    launches = eng.launch_count() - l0
    tms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    k_ms = ms / args.steps
    value = n * world * args.steps / (ms * 1e-3)
    cells = len(eng.fetch_strata()[0])

    e2e = None
    if not args.no_e2e:
        host = []
        for x, m in ((q, n), (s, n), (qs, nq), (qt, nq)):
            h = pinned_empty(m)
            h[:] = x.cpu().numpy()
            host.append(h)
        eng.classify_chunk(*host)
        barrier()
        t0 = time.perf_counter()
        e2e_steps = 3
        for _ in range(e2e_steps):
            eng.classify_chunk(*host)
            nc = len(eng.fetch_strata()[0])
        barrier()
        ems = (time.perf_counter() - t0) * 1e3
        e2e = {'value': n * world * e2e_steps / (ems * 1e-3), 'unit': UNIT,
               'h2d_bytes_per_step': int(8 * n + 8 * nq),
               'd2h_bytes_per_step': int(nc * 28), 'steps': e2e_steps,
               'ms_per_step': ems / e2e_steps,
               'api': 'wk_classify_chunk(host SoA + per-query sample and '
                      'stratum) + wk_fetch_strata'}

    cpu = None
    parity = None
    if rank == 0 and not args.no_cpu:
        from oracle import oracle as O
        from woltka_b200.engine import Engine
        m = min(n, 5_000_000)
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
                         target_rank=[0], flags=0, n_samples=S_loc,
                         n_features=T, q_sample=qsh, q_stratum=qth,
                         n_threads=threads)
        dt = time.perf_counter() - t0
        from tests import cases
        e2 = Engine(torch.cuda.current_device())
        e2.set_tree(parent, 0)
        e2.set_plan(np.array([KIND_RANK], dtype=np.int32), 0, 0.0, S_loc, T)
        e2.set_subjects(tab, None)
        e2.classify_chunk(qh, sh, qsh, qth, 0)
        got = cases.collect(e2, S_loc, T)
        e2.close()
        parity = bool(np.array_equal(got[0], exp[0]) and got[1] == exp[1] and
                      got[2] == exp[2])
        cpu = {'value': m / dt, 'unit': UNIT, 'cores': threads, 'kind': 'port',
               'sample': f'{m} records of the timed batch, C restatement of '
                         f'the reference path (oracle/woltka_oracle.c), '
                         f'{threads} OpenMP threads'}
        assert parity, 'GPU strata cells differ from the oracle on the sample'

    if rank == 0:
        achieved = 8 * n / (k_ms * 1e-3) / 1e9
        print(json.dumps({
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int32',
            'data': 'synthetic', 'config': workload_config(args, ['ko']),
            'roofline': {'bound': 'hbm', 'achieved': achieved,
                         'peak': hbm_peak, 'unit': 'GB/s',
                         'frac': achieved / hbm_peak, 'traffic': None,
                         'kernel': eng.last_kernel(), 'kernel_ms': k_ms,
                         'peak_source': peak_src,
                         'algorithmic_bytes_per_record': 8},
            'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': launches,
            'clocks': clocks, 'parity_on_sample': parity,
            'strata_cells': cells}))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
