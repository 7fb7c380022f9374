"""CPU, world_size 2 over gloo: the N > 1 path of bench.py (barrier + max-over-ranks timing, weak-scaling aggregate,
unit sharding, per-rank key spaces).  The data path itself has no collective."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lmcache_b200.dist_util import aggregate_gbps, max_over_ranks, rank_keys, shard_range
    dist.barrier()
    my_ms = 10.0 + 5.0 * rank                       # rank 1 is the slow one
    ms_max = max_over_ranks(my_ms)
    lo, hi = shard_range(7, rank, world)
    keys = rank_keys(["ab" * 32, "cd" * 32], "vllm", "m", world, rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, keys)
    dist.barrier()
    q.put((rank, ms_max, aggregate_gbps(1 << 30, ms_max, world), (lo, hi), gathered))
    dist.destroy_process_group()


def test_two_rank_timing_and_sharding():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [15.0, 15.0]                       # max over ranks, seen by every rank
    want = 2 * (1 << 30) / 15e-3 / 1e9
    assert all(abs(r[2] - want) < 1e-9 for r in res)                 # aggregate = world * per-rank bytes / max time
    assert [r[3] for r in res] == [(0, 4), (4, 7)]                   # 7 units -> 4 + 3, contiguous, disjoint
    k0, k1 = res[0][4]
    assert set(k0).isdisjoint(k1) and all(k.split("@")[3] == "0" for k in k0) and all(k.split("@")[3] == "1" for k in k1)


def test_single_process_identity():
    from lmcache_b200.dist_util import max_over_ranks, shard_range
    assert max_over_ranks(3.5) == 3.5
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
