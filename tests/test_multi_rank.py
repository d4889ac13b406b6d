"""The N>1 path without GPUs: two gloo ranks run the sharding / aggregation logic bench.py uses (no data-path collective
exists to test: code blocks are independent)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from srslte_b200 import shard

    dist.init_process_group("gloo", rank=rank, world_size=world)
    r, w, l = shard.rank_info()
    assert (r, w, l) == (rank, world, rank)
    first, last = shard.shard_range(64 * 1000, r, w)
    ms = 10.0 + 5.0 * rank  # rank 1 is slower: the job time is the max
    thr, ms_max = shard.aggregate_throughput(float(last - first), ms, dist)
    cells = [c for c in range(64) if shard.cell_to_rank(c, w) == r]
    q.put((rank, first, last, thr, ms_max, cells, shard.shard_seed(7, r)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_aggregation():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, f0, l0, thr0, ms0, cells0, s0), (_, f1, l1, thr1, ms1, cells1, s1) = res
    assert (f0, l0, f1, l1) == (0, 32000, 32000, 64000)           # disjoint, complete
    assert ms0 == ms1 == 15.0                                       # max over ranks
    assert thr0 == thr1 == pytest.approx(64000 / 15e-3)              # all ranks' units / max time
    assert sorted(cells0 + cells1) == list(range(64)) and not set(cells0) & set(cells1)
    assert s0 != s1


def test_shard_range_properties():
    sys.path.insert(0, ROOT)
    from srslte_b200 import shard

    for n in (0, 1, 7, 64, 65536, 832000):
        for w in (1, 2, 4, 8):
            parts = [shard.shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
