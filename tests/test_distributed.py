"""N > 1 host logic on CPU: two gloo ranks classify disjoint shards (oracle
behind the engine interface), one all-reduce merges the units tables; the
result must equal the single-process table."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from tests import cases
    from woltka_b200 import synth
    from woltka_b200.distributed import shard_bounds, allreduce_counts
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    tax = synth.Taxonomy(seed=7, level_sizes=[1, 2, 5, 12, 30, 60, 150, 400],
                         n_genomes=900)
    case = cases.Case(tax, n_extra=10, internal_subjects=20, seed=3)
    q, s = cases.random_hits(case, 30000, seed=5)
    cuts = shard_bounds(q, world)
    a, b = cuts[rank], cuts[rank + 1]
    units, ovf, _ = cases.run_oracle(case, ['genus', 'none'], 0, 0.8,
                                     q[a:b], s[a:b])
    t = torch.from_numpy(units.copy())
    allreduce_counts(t)
    if rank == 0:
        full, _, _ = cases.run_oracle(case, ['genus', 'none'], 0, 0.8, q, s)
        np.save(out, np.array([np.array_equal(t.numpy(), full),
                               int(t.numpy().sum() // 720720)]))
    dist.destroy_process_group()


def test_two_ranks_gloo(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / 'res.npy')
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = np.load(out)
    assert res[0] == 1 and res[1] > 0


def test_shard_bounds_never_split_a_query():
    from woltka_b200.distributed import shard_bounds
    q = np.repeat(np.arange(100), np.random.default_rng(0).integers(1, 9, 100))
    for world in (1, 2, 3, 8, 64):
        cuts = shard_bounds(q, world)
        assert cuts[0] == 0 and cuts[-1] == len(q) and len(cuts) == world + 1
        assert all(x <= y for x, y in zip(cuts, cuts[1:]))
        for c in cuts[1:-1]:
            assert c == 0 or c == len(q) or q[c] != q[c - 1]
    assert shard_bounds(np.zeros(50, dtype=int), 4) == [0, 50, 50, 50, 50]
    assert shard_bounds(np.zeros(0, dtype=int), 2) == [0, 0, 0]


def test_bind_near_gpu_without_a_gpu_is_a_no_op():
    """The NUMA binding of multi-rank runs is an optimisation: without NVML or
    a device it leaves the affinity alone and says so."""
    import os
    from woltka_b200.distributed import bind_near_gpu
    before = os.sched_getaffinity(0)
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    assert bind_near_gpu(0) is None
    assert os.sched_getaffinity(0) == before
