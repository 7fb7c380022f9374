"""Helpers for the N > 1 measurement path (bench.py under torchrun).

The hot path shards by independent engines -- one process, one codec context and one pinned slab per GPU; keys
carry (world_size, worker_id) (lmcache/utils.py:12-31) -- so there is no data-path collective.  torch.distributed
is used only to line ranks up (barrier) and to take the max of their device timings."""
from __future__ import annotations

from typing import List, Tuple

import torch


def max_over_ranks(value: float, device: torch.device | str = "cpu") -> float:
    """Largest `value` across ranks (identity when torch.distributed is not initialised)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_gbps(bytes_per_rank: int, ms_max: float, world: int) -> float:
    """Whole-job throughput under weak scaling: every rank moved bytes_per_rank in (at most) ms_max."""
    return world * bytes_per_rank / (ms_max * 1e-3) / 1e9


def shard_range(n_units: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of n_units independent units (sequences / chunks) over ranks."""
    base, rem = divmod(n_units, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def rank_keys(hashes: List[str], fmt: str, model: str, world: int, rank: int) -> List[str]:
    """Wire keys of one rank's chunks: `fmt@model@world@rank@hash` -- disjoint across ranks by construction."""
    from lmcache_b200.utils import CacheEngineKey
    return [CacheEngineKey(fmt, model, world, rank, h).to_string() for h in hashes]
