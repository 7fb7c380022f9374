"""Key type and annotations shared across the package.

CacheEngineKey mirrors lmcache/utils.py:12-39 (same fields, same `fmt@model@ws@wid@hash`
string form -- the server/connector wire key, <= 150 chars per lmcache/protocol.py:4).
"""
from __future__ import annotations

import functools
from dataclasses import dataclass
from typing import Tuple

import torch

# nested tuple of per-layer (K, V) tensors
KVCache = Tuple[Tuple[torch.Tensor, torch.Tensor], ...]


@dataclass
class CacheEngineKey:
    fmt: str
    model_name: str
    world_size: int
    worker_id: int
    chunk_hash: str

    def __hash__(self):
        return hash((self.fmt, self.model_name, self.world_size, self.worker_id, self.chunk_hash))

    def to_string(self) -> str:
        return "@".join((self.fmt, self.model_name, str(self.world_size), str(self.worker_id), self.chunk_hash))

    @staticmethod
    def from_string(s: str) -> "CacheEngineKey":
        parts = s.split("@")
        if len(parts) != 5:
            raise ValueError(f"Invalid key string: {s}")
        return CacheEngineKey(parts[0], parts[1], int(parts[2]), int(parts[3]), parts[4])


def _lmcache_nvtx_annotate(func, domain: str = "lmcache"):
    """NVTX range around `func` (reference: lmcache/utils.py:42-60).  Uses torch's NVTX bindings when a
    CUDA runtime is present; otherwise a transparent wrapper."""

    @functools.wraps(func)
    def wrapper(*args, **kwargs):
        pushed = False
        if torch.cuda.is_available():
            try:
                torch.cuda.nvtx.range_push(f"{domain}:{func.__qualname__}")
                pushed = True
            except Exception:  # NVTX unavailable: annotation is best-effort only
                pushed = False
        try:
            return func(*args, **kwargs)
        finally:
            if pushed:
                torch.cuda.nvtx.range_pop()

    return wrapper
