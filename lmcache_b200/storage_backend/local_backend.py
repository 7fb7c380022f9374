"""LMCLocalBackend -- KV chunks in local GPU memory or page-locked host memory.

Reference: lmcache/storage_backend/local_backend.py:28-153.  There the "cpu" tier is a pageable
`tensor.to("cpu")` (pinned branch disabled, `torch.cuda.synchronize()` per put :82-100), a queue plus a
worker thread for non-blocking puts, and `.to("cuda")` on get (:141-144).

Here the host tier is the mover of include/b200kv.h: every put is one `b200kv_copy_async` into page-locked
memory on a dedicated side stream, completion is a CUDA event (no device-wide synchronize, no worker thread:
a "non-blocking put" is simply a put whose event has not been waited on yet), and a get is an async upload
ordered on the caller's stream.
"""
import ctypes
import threading
from typing import Dict, Optional

import torch

from lmcache_b200 import _native as N
from lmcache_b200.config import LMCacheEngineConfig
from lmcache_b200.logging import init_logger
from lmcache_b200.storage_backend.abstract_backend import LMCBackendInterface
from lmcache_b200.utils import CacheEngineKey, _lmcache_nvtx_annotate

logger = init_logger(__name__)


class _HostEntry:
    __slots__ = ("host", "event", "src")

    def __init__(self, host: torch.Tensor, event: Optional[torch.cuda.Event], src):
        self.host = host      # page-locked copy
        self.event = event    # completion of the device->host copy (None: already complete)
        self.src = src        # keeps the source alive until the copy has run

    def wait(self):
        if self.event is not None:
            self.event.synchronize()
            self.event = None
            self.src = None


def _copy_async(dst: torch.Tensor, src: torch.Tensor, stream: torch.cuda.Stream) -> None:
    N.check(N.lib().b200kv_copy_async(ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(src.data_ptr()),
                                      src.numel() * src.element_size(), stream.cuda_stream), "copy_async")


class LMCLocalBackend(LMCBackendInterface):

    def __init__(self, config: LMCacheEngineConfig):
        super().__init__()
        N.require_cuda()
        self.chunk_size = config.chunk_size
        self.config = config
        self.device = config.local_device       # "cpu" | "cuda"
        self.dst_device = "cuda"                # like the reference (:53): gets land on the GPU
        self.dict: Dict[CacheEngineKey, object] = {}
        self.update_lock = threading.Lock()
        self._side: Optional[torch.cuda.Stream] = None
        self._inflight = []   # (event, pinned tensor) of uploads still reading host memory

    def _side_stream(self, device) -> torch.cuda.Stream:
        if self._side is None or self._side.device != device:
            self._side = torch.cuda.Stream(device=device)
        return self._side

    def contains(self, key: CacheEngineKey) -> bool:
        return key in self.dict

    @_lmcache_nvtx_annotate
    def put(self, key: CacheEngineKey, kv_chunk: torch.Tensor, blocking: bool = True) -> None:
        if self.device == "cuda":
            # reference: kv_chunk.to("cuda") -- a no-op for GPU chunks, an upload for host chunks
            val = kv_chunk if kv_chunk.is_cuda else kv_chunk.to("cuda", non_blocking=False)
            with self.update_lock:
                self.dict[key] = val
            return
        # host tier
        if not kv_chunk.is_cuda:
            entry = _HostEntry(kv_chunk.detach().clone().pin_memory(), None, None)
        else:
            src = kv_chunk if kv_chunk.is_contiguous() else kv_chunk.contiguous()
            host = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
            side = self._side_stream(src.device)
            side.wait_stream(torch.cuda.current_stream(src.device))   # producer kernels finished first
            _copy_async(host, src, side)
            ev = torch.cuda.Event()
            ev.record(side)
            entry = _HostEntry(host, ev, src)
            if blocking:
                entry.wait()
        with self.update_lock:
            self.dict[key] = entry

    @_lmcache_nvtx_annotate
    def get(self, key: CacheEngineKey) -> Optional[torch.Tensor]:
        val = self.dict.get(key, None)
        if val is None:
            return None
        if isinstance(val, _HostEntry):
            val.wait()
            out = torch.empty(val.host.shape, dtype=val.host.dtype, device=self.dst_device)
            stream = torch.cuda.current_stream(out.device)
            _copy_async(out, val.host, stream)   # ordered on the consumer's stream
            # the pinned block must outlive the DMA even if the key is overwritten meanwhile
            ev = torch.cuda.Event()
            ev.record(stream)
            self._inflight = [(e, h) for e, h in self._inflight if not e.query()]
            self._inflight.append((ev, val.host))
            return out
        return val.to(self.dst_device)

    def close(self):
        for val in list(self.dict.values()):
            if isinstance(val, _HostEntry):
                val.wait()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
