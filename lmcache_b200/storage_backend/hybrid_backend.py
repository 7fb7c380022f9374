"""LMCHybridBackend -- local tier in front of a remote one (lmcache/storage_backend/hybrid_backend.py): writes go to
both, reads are served locally when possible and fall through to the remote tier (whose chunks are then kept locally).
The reference prefetches the whole remote store into the local tier at start-up (hybrid_backend.py:26-62: list() + one
get/put per key -- a full deserialise of everything the server holds); here the local tier fills on demand.

Both tiers keep their engine fast paths: a store hands the caller's KV view to each tier once, a retrieve asks the local
tier for the longest prefix it holds and the remote tier for the rest, all decoded / copied straight into the one
destination blob."""
from typing import Iterable, List, Optional, Tuple

import torch

from lmcache_b200.config import LMCacheEngineConfig, LMCacheEngineMetadata
from lmcache_b200.storage_backend.abstract_backend import LMCBackendInterface
from lmcache_b200.utils import CacheEngineKey


class LMCHybridBackend(LMCBackendInterface):

    def __init__(self, config: LMCacheEngineConfig, metadata: LMCacheEngineMetadata):
        super().__init__()
        from lmcache_b200.storage_backend import CreateStorageBackend
        local_cfg = LMCacheEngineConfig(config.chunk_size, config.local_device, None, None, False, config.save_decode_cache,
                                        config.local_serde)
        remote_cfg = LMCacheEngineConfig(config.chunk_size, None, config.remote_url, config.remote_serde,
                                         config.pipelined_backend, config.save_decode_cache, None)
        self.local_store = CreateStorageBackend(local_cfg, metadata)
        self.remote_store = CreateStorageBackend(remote_cfg, metadata)

    def contains(self, key: CacheEngineKey) -> bool:
        return self.local_store.contains(key) or self.remote_store.contains(key)

    def put(self, key: CacheEngineKey, kv_chunk: torch.Tensor, blocking: bool = True) -> None:
        self.local_store.put(key, kv_chunk, blocking=True)
        self.remote_store.put(key, kv_chunk, blocking)

    def get(self, key: CacheEngineKey) -> Optional[torch.Tensor]:
        val = self.local_store.get(key)
        if val is None:
            val = self.remote_store.get(key)
            if val is not None:
                self.local_store.put(key, val, blocking=True)
        return val

    def batched_put(self, keys_and_chunks: Iterable[Tuple[CacheEngineKey, torch.Tensor]], blocking=True) -> int:
        n = 0
        for key, chunk in keys_and_chunks:
            self.put(key, chunk, blocking=blocking)
            n += 1
        return n

    # ------------------------------------------------------------------ engine fast paths
    def supports_kv_view(self) -> bool:
        f, g = getattr(self.local_store, "supports_kv_view", None), getattr(self.remote_store, "supports_kv_view", None)
        return bool(f and f() and g and g())

    def put_kv_chunks(self, keys: List[CacheEngineKey], view, tok_begin: int, chunk_size: int, blocking: bool = True) -> int:
        self.local_store.put_kv_chunks(keys, view, tok_begin, chunk_size, blocking=True)
        return self.remote_store.put_kv_chunks(keys, view, tok_begin, chunk_size, blocking=blocking)

    def get_kv_into(self, keys: List[CacheEngineKey], dst, dst_tok0: int, chunk_size: int) -> int:
        n = self.local_store.get_kv_into(keys, dst, dst_tok0, chunk_size)
        if n < len(keys):
            n += self.remote_store.get_kv_into(keys[n:], dst, dst_tok0 + n * chunk_size, chunk_size)
        return n

    def peek_geometry(self, key: CacheEngineKey, fmt: str = "vllm"):
        for store in (self.local_store, self.remote_store):
            f = getattr(store, "peek_geometry", None)
            g = f(key, fmt) if f is not None else None
            if g is not None:
                return g
        return None

    def out_dtype(self):
        f = getattr(self.remote_store, "out_dtype", None) or \
            getattr(getattr(self.remote_store, "deserializer", None), "out_dtype", None)
        return f() if f is not None else None

    def close(self):
        self.local_store.close()
        self.remote_store.close()
