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


def _classify_worker(rank, world, port, out, names):
    """classify() under torch.distributed: the files of a golden case dealt out
    to the ranks (oracle-backed engines on CPU), exact profiles merged on
    rank 0."""
    sys.path.insert(0, ROOT)
    import io
    import pickle
    from contextlib import redirect_stdout
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from tests import test_golden as G
    res = {}
    for name in names:
        with redirect_stdout(io.StringIO()):
            got, exp_raw, exp_rounded = G.run_case(name, 'oracle')
        res[name] = got
    if rank == 0:
        with open(out, 'wb') as f:
            pickle.dump(res, f)
    else:
        # the other ranks hold empty profiles and do not write tables
        from woltka_b200.workflow import is_output_rank
        assert not is_output_rank()
        assert all(not prof for got in res.values() for prof in got.values())
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_classify_shards_files_over_ranks(tmp_path, world):
    """The drop-in classify() under torchrun: same profiles (values and
    sample order) as a single process — counts are additive over any
    partition of the files (workflow.py:1058, doc/perform.md:70-98).  Cases:
    one file per sample (fractional counts), multi-rank --above, stratified,
    demultiplexed (one file: one rank does it)."""
    import pickle
    import torch.multiprocessing as mp
    from tests import test_golden as G
    names = ['bowtie2_ogu', 'bt2sho_phylo', 'burst_genus_process',
             'blastn_species', 'synth_above', 'synth_default',
             'bt2sho_order_sizes']
    out = str(tmp_path / 'res.pkl')
    port = 29500 + (os.getpid() + world) % 2000
    mp.spawn(_classify_worker, args=(world, port, out, names), nprocs=world,
             join=True)
    with open(out, 'rb') as f:
        res = pickle.load(f)
    for name in names:
        import io
        from contextlib import redirect_stdout
        with redirect_stdout(io.StringIO()):
            single, exp_raw, exp_rounded = G.run_case(name, 'oracle')
        G.check(res[name], exp_raw, exp_rounded)
        if 'sizes' not in name:   # (size-weighted cells are floating-point sums)
            assert res[name] == single, name
        for rk in single:       # same sample order as one process
            assert list(res[name][rk]) == list(single[rk]), (name, rk)


def _nccl_merge_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from tests import cases
    from woltka_b200 import synth
    from woltka_b200.distributed import shard_bounds, merge_engine
    from woltka_b200.engine import Engine
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    tax = synth.Taxonomy(seed=7, level_sizes=[1, 2, 5, 12, 30, 60, 150, 400],
                         n_genomes=900)
    case = cases.Case(tax, n_extra=10, internal_subjects=20, seed=3)
    # long queries of 17 and 19 distinct subjects: shares that go to the
    # overflow list; strata on every query
    q, s = cases.random_hits(case, 40000, seed=5, long_every=997, long_len=17)
    k = np.bincount(q)
    pos = np.arange(len(q)) - (np.cumsum(k) - k)[q]
    long = k[q] == 17          # 17 distinct subjects: shares of 1/17
    s[long] = ((q[long] * 7 + pos[long]) % case.V).astype(np.int32)
    nq = int(q.max()) + 1
    rng = np.random.default_rng(1)
    q_sample = np.sort(rng.integers(0, 4, nq)).astype(np.int32)
    q_stratum = rng.integers(-1, 30, nq).astype(np.int32)
    cuts = shard_bounds(q, world)
    a, b = cuts[rank], cuts[rank + 1]
    ok = []
    for strata in (False, True):
        eng = Engine(rank)
        eng.set_stream(torch.cuda.current_stream().cuda_stream)
        kinds, tab, _ = case.tables(['genus', 'none'])
        eng.set_tree(case.ft.parent, 0)
        eng.set_plan(kinds, 0, 0.8, 4, case.NF)
        eng.set_subjects(tab, case.sub_node)
        eng.classify_chunk(q[a:b], s[a:b], q_sample,
                           q_stratum if strata else None, 0)
        merge_engine(eng, dst=0, dense=not strata, strata=strata)
        if rank == 0:
            got = cases.collect(eng, 4, case.NF)
            exp = cases.run_oracle(case, ['genus', 'none'], 0, 0.8, q, s,
                                   n_samples=4, q_sample=q_sample,
                                   q_stratum=q_stratum if strata else None)
            ok.append(bool(np.array_equal(got[0], exp[0])) and
                      got[1] == exp[1] and got[2] == exp[2] and
                      (len(got[1]) > 0) and (strata == bool(got[2])))
        eng.close()
    # strata cells merged by key ownership (reduce-scatter): the union of the
    # ranks' tables is the oracle's, no key on two ranks
    from woltka_b200.distributed import reduce_scatter_strata
    eng = Engine(rank)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    kinds, tab, _ = case.tables(['genus', 'none'])
    eng.set_tree(case.ft.parent, 0)
    eng.set_plan(kinds, 0, 0.8, 4, case.NF)
    eng.set_subjects(tab, case.sub_node)
    eng.classify_chunk(q[a:b], s[a:b], q_sample, q_stratum, 0)
    owned = reduce_scatter_strata(eng)
    merge_engine(eng, dst=0, dense=False, strata=False)     # the overflow list
    got = cases.collect(eng, 4, case.NF)
    parts = [None] * world if rank == 0 else None
    dist.gather_object((got[2], owned), parts, dst=0)
    if rank == 0:
        exp = cases.run_oracle(case, ['genus', 'none'], 0, 0.8, q, s,
                               n_samples=4, q_sample=q_sample,
                               q_stratum=q_stratum)
        union = {}
        for cells, n in parts:
            assert len(cells) <= n and not (set(cells) & set(union))
            union.update(cells)
        ok.append(union == exp[2] and got[1] == exp[1] and
                  all(len(c) > 0 for c, _ in parts))
    eng.close()
    if rank == 0:
        np.save(out, np.array(ok))
    dist.destroy_process_group()


@pytest.mark.gpu
def test_merge_engine_over_nccl(tmp_path):
    """Two GPUs, one rank each: dense units table (one reduce), overflow list
    and strata cells (sent to rank 0, added by key) against the oracle on the
    whole stream."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    import torch.multiprocessing as mp
    out = str(tmp_path / 'res.npy')
    port = 29500 + os.getpid() % 2000
    mp.spawn(_nccl_merge_worker, args=(2, port, out), nprocs=2, join=True)
    assert np.load(out).tolist() == [True, True, True]
