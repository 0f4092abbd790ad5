"""Multi-GPU plumbing: one process per GPU, records sharded at query
boundaries, ONE collective at the end.

The reference has no parallel mode; its documented scale-out is "split the
input, run independent jobs, `woltka merge` the tables"
(/root/reference/doc/perform.md:70-92; tools.merge_wf sums cells).  Counts are
additive over any partition of the queries, so the same holds here: each rank
classifies its shard into its own units table and the tables are summed by
one NCCL reduce to rank 0 (or all-reduce when every rank wants the merged
table) over NVLink (int64, exact).
"""
import numpy as np


def shard_bounds(qidx, world):
    """Cut [0, n) into `world` contiguous ranges that never split a query
    (all records of a query must be classified together, classify.py:81-127).
    Returns world+1 offsets."""
    qidx = np.asarray(qidx)
    n = len(qidx)
    cuts = [0]
    for r in range(1, world):
        c = max(n * r // world, cuts[-1])
        while 0 < c < n and qidx[c] == qidx[c - 1]:
            c += 1
        cuts.append(c)
    cuts.append(n)
    return cuts


def allreduce_counts(tensor):
    """Sum the per-rank units tables in place (no-op for a single rank)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and \
            dist.get_world_size() > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


def reduce_counts(tensor, dst=0):
    """Sum the per-rank units tables into rank `dst` (the single NCCL reduce
    that replaces `woltka merge`); no-op for a single rank."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and \
            dist.get_world_size() > 1:
        dist.reduce(tensor, dst=dst, op=dist.ReduceOp.SUM)
    return tensor


def _active():
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized() and \
        dist.get_world_size() > 1


def merge_engine(eng, dst=0, dense=True, strata=False):
    """Merge every rank's results into rank `dst`'s context — the device form
    of `woltka merge` (/root/reference/woltka/tools.py:153-208): ONE reduce
    of the dense units table, and the two sparse parts — the strata cells
    (classify.counter_strat keys) and the list of shares whose denominator
    does not divide UNITS — sent to `dst`, which adds them into its own table
    (reduce by key).  All ranks must share the plan and the index spaces.
    Collective: every rank calls it."""
    import torch
    import torch.distributed as dist
    if not _active():
        return
    rank, world = dist.get_rank(), dist.get_world_size()
    if dense:
        dist.reduce(eng.counts_tensor(), dst=dst, op=dist.ReduceOp.SUM)
    parts = [eng.overflow_export()]
    if strata:
        parts.append(eng.strata_export())
    dev = parts[0][0].device
    sizes = torch.tensor([p[0].numel() for p in parts], device=dev,
                         dtype=torch.int64)
    every = [torch.empty_like(sizes) for _ in range(world)]
    dist.all_gather(every, sizes)
    every = torch.stack(every).cpu().tolist()      # [rank][part]
    for pi, (a, b) in enumerate(parts):
        if rank == dst:
            if pi == 1:
                # every cell that is on its way may be a new one
                eng.strata_reserve(sum(every[src][pi] for src in range(world)
                                       if src != dst))
            for src in range(world):
                n = every[src][pi]
                if src == dst or not n:
                    continue
                ka = torch.empty(n, dtype=a.dtype, device=dev)
                kb = torch.empty(n, dtype=b.dtype, device=dev)
                dist.recv(ka, src=src)
                dist.recv(kb, src=src)
                torch.cuda.current_stream(dev).synchronize()
                if pi == 0:
                    eng.overflow_import(ka, kb, stratified=strata)
                else:
                    eng.strata_import(ka, kb)
        elif every[rank][pi]:
            dist.send(a.contiguous(), dst=dst)
            dist.send(b.contiguous(), dst=dst)


def reduce_scatter_strata(eng):
    """Merge the strata cells of all ranks BY KEY OWNERSHIP: rank r ends up
    with the sum of every rank's cells whose key hashes to r (a reduce-scatter
    of the sparse table; classify.counter_strat keys, summed like
    `woltka merge`, tools.py:153-208).  Each rank hashes 1/world of the cells
    instead of one rank hashing all of them: the cells go out with ONE
    all-to-all over NVLink and come back into the (emptied) local table.
    The merged table is the union of the ranks' tables — every rank fetches
    its share (`Engine.fetch_strata`) and the host concatenates them.
    Collective: every rank calls it.  Returns the number of cells received
    (an upper bound of the cells owned: equal keys of two ranks become one)."""
    import torch
    import torch.distributed as dist
    k, u = eng.strata_export()
    if not _active():
        return int(k.numel())
    world = dist.get_world_size()
    dev = k.device
    # owner of a key: a multiplicative hash, independent of any table size
    owner = ((k * -7046029254386353131) >> 40) % world       # int64 wraps
    srt = torch.sort(owner.to(torch.int16))
    k, u = k[srt.indices], u[srt.indices]    # (copies: k, u were views into
                                              # the engine's buffers)
    edges = torch.searchsorted(srt.values, torch.arange(
        world + 1, device=dev, dtype=torch.int16))
    counts = (edges[1:] - edges[:-1]).to(torch.int64)
    got = torch.empty_like(counts)
    dist.all_to_all_single(got, counts)
    n_in, n_out = got.cpu().tolist(), counts.cpu().tolist()
    rk = torch.empty(sum(n_in), dtype=k.dtype, device=dev)
    ru = torch.empty(sum(n_in), dtype=u.dtype, device=dev)
    dist.all_to_all_single(rk, k, n_in, n_out)
    dist.all_to_all_single(ru, u, n_in, n_out)
    torch.cuda.current_stream(dev).synchronize()
    eng.reset_strata()
    eng.strata_reserve(rk.numel())
    eng.strata_import(rk, ru)
    return int(rk.numel())


def merge_profiles(data, dst=0):
    """Object-level merge for the drop-in `classify()`: every rank interns its
    own strings, so the per-rank results travel as exact cells keyed by NAMES
    ({rank: {sample: {feature: Fraction}}}) and are summed on `dst`
    (util.sum_dict, util.py:78-94, applied across jobs like `woltka merge`).
    Returns the merged dict on `dst`, None elsewhere."""
    import torch.distributed as dist
    if not _active():
        return data
    rank, world = dist.get_rank(), dist.get_world_size()
    got = [None] * world if rank == dst else None
    dist.gather_object(data, got, dst=dst)
    if rank != dst:
        return None
    out = {}
    for part in got:
        for rk, samples in part.items():
            orow = out.setdefault(rk, {})
            for sample, prof in samples.items():
                o = orow.setdefault(sample, {})
                for key, val in prof.items():
                    o[key] = o.get(key, 0) + val
    return out


def bind_near_gpu(device):
    """Restrict this process to the CPUs NVML lists as local to CUDA device
    `device`, so that host buffers pinned afterwards (first touch) and the
    copy threads sit on the GPU's NUMA node — with one rank per GPU the ranks
    otherwise share one socket's memory and the inter-socket link.  Returns
    the CPU list, or None when NVML / the topology is not available (the
    binding is an optimisation, never a requirement)."""
    import os
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(device)
        bus = '%08x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id,
                                    pr.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        ncpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        near = {64 * w + b for w, m in enumerate(words) for b in range(64)
                if (int(m) >> b) & 1}
        cpus = sorted(near & os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None
