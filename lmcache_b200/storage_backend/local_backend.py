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

    # ------------------------------------------------------------------ engine fast paths
    def supports_kv_view(self) -> bool:
        return True

    def put_kv_chunks(self, keys, view, tok_begin: int, chunk_size: int, blocking: bool = True) -> int:
        """Store tokens [tok_begin, T) of `view` as len(keys) chunk blobs: ONE gather kernel (b200kv_pack_chunks)
        builds every chunk blob; for the host tier ONE device->host DMA moves them all into a page-locked slab."""
        fmt_hf = getattr(view, "fmt", "vllm") == "huggingface"
        n_tok = view.ntokens - tok_begin
        n_chunks = len(keys)
        assert n_chunks == (n_tok + chunk_size - 1) // chunk_size
        last = n_tok - (n_chunks - 1) * chunk_size
        per_tok = 2 * view.L * view.H * view.D
        stride = per_tok * chunk_size
        dev = torch.empty(n_chunks * stride, dtype=view.dtype, device=view.device)
        with torch.cuda.device(view.device):
            cur = torch.cuda.current_stream()
            N.check(N.lib().b200kv_pack_chunks(ctypes.byref(view.desc), tok_begin, n_chunks, chunk_size, last,
                                               1 if fmt_hf else 0, ctypes.c_void_p(dev.data_ptr()),
                                               stride * dev.element_size(), cur.cuda_stream), "pack_chunks")

            def shape(t):
                return (view.L, 2, view.H, t, view.D) if fmt_hf else (view.L, 2, t, view.H, view.D)

            if self.device == "cuda":
                vals = [dev[j * stride: j * stride + per_tok * (chunk_size if j < n_chunks - 1 else last)]
                        .view(shape(chunk_size if j < n_chunks - 1 else last)) for j in range(n_chunks)]
            else:
                host = torch.empty(n_chunks * stride, dtype=view.dtype, pin_memory=True)
                side = self._side_stream(view.device)
                side.wait_stream(cur)
                _copy_async(host, dev, side)
                ev = torch.cuda.Event()
                ev.record(side)
                vals = []
                for j in range(n_chunks):
                    t = chunk_size if j < n_chunks - 1 else last
                    vals.append(_HostEntry(host[j * stride: j * stride + per_tok * t].view(shape(t)), ev, dev))
                if blocking:
                    ev.synchronize()
                    for v in vals:
                        v.event, v.src = None, None
        with self.update_lock:
            for key, v in zip(keys, vals):
                self.dict[key] = v
        return n_chunks

    def get_kv_into(self, keys, dst, dst_tok0: int, chunk_size: int) -> int:
        """Copy consecutive chunks (until the first miss) straight into the destination blob view `dst` at token
        offsets dst_tok0 + i * chunk_size: strided 2-D copies (host tier: async uploads), no intermediate chunk tensors,
        no torch.cat."""
        fmt_hf = getattr(dst, "fmt", "vllm") == "huggingface"
        blob = dst.blob
        if blob is None:
            return self._get_kv_scatter(keys, dst, dst_tok0, chunk_size)
        n = 0
        with torch.cuda.device(blob.device):
            stream = torch.cuda.current_stream()
            for i, key in enumerate(keys):
                val = self.dict.get(key, None)
                if val is None:
                    break
                src = val.host if isinstance(val, _HostEntry) else val
                if isinstance(val, _HostEntry):
                    val.wait()
                elif not src.is_cuda:
                    src = src.cuda()
                t = src.shape[3] if fmt_hf else src.shape[2]
                tok = dst_tok0 + i * chunk_size
                if tok + t > dst.ntokens or src.dtype != blob.dtype:
                    break
                es = blob.element_size()
                if fmt_hf:      # rows = (l, kv, h): t*D contiguous elements each
                    rows, row_bytes = blob.shape[0] * 2 * blob.shape[2], t * blob.shape[4] * es
                    dst_pitch = blob.shape[3] * blob.shape[4] * es
                    dptr = blob.data_ptr() + tok * blob.shape[4] * es
                else:           # rows = (l, kv): t*H*D contiguous elements each
                    rows, row_bytes = blob.shape[0] * 2, t * blob.shape[3] * blob.shape[4] * es
                    dst_pitch = blob.shape[2] * blob.shape[3] * blob.shape[4] * es
                    dptr = blob.data_ptr() + tok * blob.shape[3] * blob.shape[4] * es
                N.check(N.lib().b200kv_copy2d_async(ctypes.c_void_p(dptr), dst_pitch, ctypes.c_void_p(src.data_ptr()),
                                                    row_bytes, row_bytes, rows, stream.cuda_stream), "copy2d")
                if isinstance(val, _HostEntry):
                    ev = torch.cuda.Event()
                    ev.record(stream)
                    self._inflight = [(e, h) for e, h in self._inflight if not e.query()]
                    self._inflight.append((ev, val.host))
                n += 1
        return n

    def _get_kv_scatter(self, keys, dst, dst_tok0: int, chunk_size: int) -> int:
        """get_kv_into for destinations that are not one blob (the engine's 2L tensors, or a paged KV cache with its
        slot mapping): each hit chunk is scattered by ONE b200kv_unpack_chunks launch (host tier: after one upload)."""
        n = 0
        hf = getattr(dst, "fmt", "vllm") == "huggingface"      # chunk blobs carry the engine's layout
        with torch.cuda.device(dst.device):
            stream = torch.cuda.current_stream()
            for i, key in enumerate(keys):
                val = self.dict.get(key, None)
                if val is None:
                    break
                if isinstance(val, _HostEntry):
                    val.wait()
                    src = val.host.to(dst.device, non_blocking=True)
                else:
                    src = val if val.is_cuda else val.cuda()
                t = src.shape[3] if hf else src.shape[2]
                tok = dst_tok0 + i * chunk_size
                if tok + t > dst.ntokens or src.dtype != dst.dtype:
                    break
                src = src.contiguous()
                N.check(N.lib().b200kv_unpack_chunks(ctypes.c_void_p(src.data_ptr()), src.numel() * src.element_size(), 1,
                                                     t, t, 1 if hf else 0, ctypes.byref(dst.desc), tok,
                                                     stream.cuda_stream), "unpack_chunks")
                src.record_stream(stream)
                n += 1
        return n

    def close(self):
        for val in list(self.dict.values()):
            if isinstance(val, _HostEntry):
                val.wait()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
