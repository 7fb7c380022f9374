"""Host-side driver of the CUDA codec: descriptors, buffers, and batched encode / decode calls.

This is plumbing above the C ABI (include/b200kv.h): PyTorch supplies device memory and streams,
libb200kv does all the work.  The serde plugins (storage_backend/serde/cachegen_*.py) and the
engine fast paths are thin layers over `CacheGenCodec`.
"""
from __future__ import annotations

import contextlib
import ctypes
import threading
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from lmcache_b200 import _native as N

_DTYPE_CODE = {torch.bfloat16: N.DT_BF16, torch.float16: N.DT_FP16}
_CODE_DTYPE = {v: k for k, v in _DTYPE_CODE.items()}


def _stream_ptr(stream: Optional[torch.cuda.Stream]) -> int:
    return (stream if stream is not None else torch.cuda.current_stream()).cuda_stream


class PinnedBuffer:
    """Page-locked, device-mapped host memory from b200kv_pinned_alloc (replaces the reference's
    pageable .to("cpu") staging, local_backend.py:82-100)."""

    def __init__(self, nbytes: int):
        N.require_cuda()
        self.nbytes = int(nbytes)
        p = ctypes.c_void_p()
        N.check(N.lib().b200kv_pinned_alloc(ctypes.byref(p), self.nbytes), "pinned_alloc")
        self.host_ptr = p.value
        d = ctypes.c_void_p()
        N.check(N.lib().b200kv_host_device_ptr(p, ctypes.byref(d)), "host_device_ptr")
        self.dev_ptr = d.value
        self._arr = (ctypes.c_uint8 * self.nbytes).from_address(self.host_ptr)

    def view(self, offset: int = 0, nbytes: Optional[int] = None) -> memoryview:
        n = self.nbytes - offset if nbytes is None else nbytes
        return memoryview(self._arr)[offset:offset + n]

    def close(self):
        if self.host_ptr:
            self._arr = None
            N.lib().b200kv_pinned_free(ctypes.c_void_p(self.host_ptr))
            self.host_ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class KvView:
    """A KV source / destination for the native library: either one strided blob tensor or the engine's
    tuple of 2L per-layer tensors (no stack / permute / contiguous copies).  Keeps the tensors alive."""

    def __init__(self, desc: N.KvDesc, keep, ntokens: int, device: torch.device, dtype: torch.dtype,
                 fmt: str = "vllm", blob: Optional[torch.Tensor] = None):
        self.desc = desc
        self._keep = keep
        self.ntokens = ntokens
        self.device = device
        self.dtype = dtype
        self.fmt = fmt
        self.blob = blob      # the single blob tensor behind this view, when there is one

    @property
    def L(self): return self.desc.L
    @property
    def H(self): return self.desc.H
    @property
    def D(self): return self.desc.D

    @staticmethod
    def _code(dtype: torch.dtype) -> int:
        if dtype not in _DTYPE_CODE:
            raise TypeError(f"KV dtype must be bfloat16 or float16, got {dtype}")
        return _DTYPE_CODE[dtype]

    @staticmethod
    def from_blob(blob: torch.Tensor, fmt: str) -> "KvView":
        """blob: [L,2,T,H,D] (vllm) or [L,2,H,T,D] (huggingface); any strides with a contiguous last dim."""
        if blob.dim() != 5 or blob.shape[1] != 2:
            raise ValueError(f"expected a [L,2,..] KV blob, got {tuple(blob.shape)}")
        if not blob.is_cuda:
            raise RuntimeError("KV blob must live on a CUDA device (no CPU fallback)")
        if blob.stride(4) != 1:
            blob = blob.contiguous()
        if fmt == "vllm":
            L, _, T, H, D = blob.shape
            sL, sKV, sT, sH, _ = blob.stride()
        elif fmt == "huggingface":
            L, _, H, T, D = blob.shape
            sL, sKV, sH, sT, _ = blob.stride()
        else:
            raise ValueError(f"Invalid format: {fmt}")
        d = N.KvDesc()
        d.base = blob.data_ptr()
        d.planes = None
        d.sL, d.sKV, d.sT, d.sH = sL, sKV, sT, sH
        d.L, d.H, d.D = L, H, D
        d.dtype = KvView._code(blob.dtype)
        return KvView(d, blob, T, blob.device, blob.dtype, fmt, blob)

    @staticmethod
    def from_tuple(kv: Sequence[Tuple[torch.Tensor, torch.Tensor]], fmt: str) -> "KvView":
        """kv: L pairs of [T,H,D] (vllm) / [H,T,D] (huggingface) tensors, as passed to LMCacheEngine.store."""
        L = len(kv)
        if L == 0:
            raise ValueError("Empty kv_tensors")
        ref = kv[0][0]
        if not ref.is_cuda:
            raise RuntimeError("KV tensors must live on a CUDA device (no CPU fallback)")
        keep = []
        ptrs = (ctypes.c_void_p * (2 * L))()
        for l, (k, v) in enumerate(kv):
            for kvi, t in ((0, k), (1, v)):
                if t.shape != ref.shape or t.dtype != ref.dtype or t.device != ref.device:
                    raise ValueError("all K/V tensors must share shape, dtype and device")
                if t.stride() != ref.stride() or t.stride(2) != 1:
                    t = t.contiguous()
                    if t.stride() != ref.stride():
                        return KvView.from_tuple(tuple((a.contiguous(), b.contiguous()) for a, b in kv), fmt)
                keep.append(t)
                ptrs[kvi * L + l] = t.data_ptr()
        if fmt == "vllm":
            T, H, D = ref.shape
            sT, sH, _ = ref.stride()
        elif fmt == "huggingface":
            H, T, D = ref.shape
            sH, sT, _ = ref.stride()
        else:
            raise ValueError(f"Invalid format: {fmt}")
        d = N.KvDesc()
        d.base = None
        d.planes = ctypes.cast(ptrs, ctypes.POINTER(ctypes.c_void_p))
        d.sL = d.sKV = 0
        d.sT, d.sH = sT, sH
        d.L, d.H, d.D = L, H, D
        d.dtype = KvView._code(ref.dtype)
        return KvView(d, (keep, ptrs), T, ref.device, ref.dtype, fmt)

    @staticmethod
    def from_paged(kv_caches: Sequence[Tuple[torch.Tensor, torch.Tensor]], slot_mapping: torch.Tensor) -> "KvView":
        """vLLM's paged KV cache used in place: `kv_caches` holds, per layer, the (key_cache, value_cache) pair shaped
        [num_blocks, block_size, H, D] (or already flattened [num_slots, H, D]); `slot_mapping` (int64, CUDA) gives the
        cache row of every token of the sequence (block * block_size + offset) -- what lmcache-vllm's
        lmcache_store_kv / lmcache_retrieve_kv gather and scatter with torch indexing (LLM_Engine.rst:91-109).
        The codec kernels read (encode) and write (decode) the rows directly; token i of the view is row
        slot_mapping[i]."""
        L = len(kv_caches)
        if L == 0:
            raise ValueError("Empty kv_caches")
        if slot_mapping.dtype != torch.int64 or not slot_mapping.is_cuda or slot_mapping.dim() != 1:
            raise ValueError("slot_mapping must be a 1-D int64 CUDA tensor")
        slot_mapping = slot_mapping.contiguous()
        ref = kv_caches[0][0]
        if not ref.is_cuda:
            raise RuntimeError("KV caches must live on a CUDA device (no CPU fallback)")
        keep = [slot_mapping]
        ptrs = (ctypes.c_void_p * (2 * L))()
        for l, (k, v) in enumerate(kv_caches):
            for kvi, t in ((0, k), (1, v)):
                if t.shape != ref.shape or t.dtype != ref.dtype or t.device != ref.device or t.stride() != ref.stride():
                    raise ValueError("all K/V caches must share shape, strides, dtype and device")
                if not t.is_contiguous():
                    raise ValueError("paged K/V caches must be contiguous (they are written in place)")
                keep.append(t)
                ptrs[kvi * L + l] = t.data_ptr()
        H, D = ref.shape[-2], ref.shape[-1]
        d = N.KvDesc()
        d.base = None
        d.planes = ctypes.cast(ptrs, ctypes.POINTER(ctypes.c_void_p))
        d.sL = d.sKV = 0
        d.sT, d.sH = H * D, D
        d.L, d.H, d.D = L, H, D
        d.dtype = KvView._code(ref.dtype)
        d.slot_map = slot_mapping.data_ptr()
        return KvView(d, (keep, ptrs), slot_mapping.numel(), ref.device, ref.dtype, "vllm")


def parse_header(buf) -> N.Header:
    """Validate and return the 64-byte header of a B2KV container (bytes / bytearray / memoryview)."""
    mv = memoryview(buf)
    if mv.nbytes < N.HEADER_BYTES:
        raise ValueError("buffer too small for a B2KV container")
    hd = N.Header.from_buffer_copy(bytes(mv[:N.HEADER_BYTES]))
    if hd.magic != N.MAGIC:
        raise ValueError("not a B2KV container (bad magic)")
    if hd.version not in (1, 2, 3):
        raise ValueError(f"unsupported B2KV version {hd.version}")
    if hd.total_bytes > mv.nbytes:
        raise ValueError("truncated B2KV container")
    if hd.status != 0:
        raise ValueError(f"B2KV container carries encoder error status {hd.status}")
    nb = None
    if hd.version == 3:
        if not 0 < hd.L <= N.MAX_PLANES // 2 or mv.nbytes < N.HEADER_BYTES + 2 * hd.L:
            raise ValueError("B2KV header carries an impossible shape")
        nb = list(bytes(mv[N.HEADER_BYTES:N.HEADER_BYTES + 2 * hd.L]))
    check_header(hd, nb)
    return hd


def container_layout_of(hd: "N.Header") -> "N.Layout":
    """Section offsets of a parsed container."""
    return N.container_layout(hd.L, hd.H, hd.D, hd.ntokens, int(hd.version) - 1)


def check_header(hd: "N.Header", nb: Optional[Sequence[int]] = None) -> None:
    """Structural checks that make a damaged blob a miss (ValueError) instead of bad device addresses: the section
    offsets follow from (L, H, D, ntokens, version), so total_bytes must be exactly fixed sections + payload.  A compact
    container (version 3) also carries its nb map -- the 2L bytes after the header, passed as `nb` and attached to the
    header as `hd.nb`: the symbols per plane its writer's bin table allowed."""
    if not (0 < hd.L <= N.MAX_PLANES // 2 and hd.H > 0 and hd.D > 0 and hd.ntokens > 0):
        raise ValueError("B2KV header carries an impossible shape")
    if hd.max_dtype not in (N.DT_BF16, N.DT_FP16):
        raise ValueError("B2KV header carries an unknown max_dtype")
    hd.nb = None
    if hd.version == 3:
        if nb is None or len(nb) != 2 * hd.L or any(v < 4 or v > 32 or v % 2 for v in nb):
            raise ValueError("B2KV v3 header: bad nb map")
        if hd.ntokens > N.GROUP_TOKENS:
            raise ValueError("B2KV v3 header: more than 256 tokens")
        hd.nb = [int(v) for v in nb]
    lo = container_layout_of(hd)
    if hd.ngroups != (hd.ntokens + N.GROUP_TOKENS - 1) // N.GROUP_TOKENS:
        raise ValueError("B2KV header: ngroups does not match ntokens")
    if hd.total_bytes != lo.off_payload + hd.payload_bytes:
        raise ValueError("B2KV header: total_bytes != fixed sections + payload_bytes (truncated or corrupt)")
    nstreams = 2 * hd.L * hd.H * hd.D * hd.ngroups
    per_stream = 1 if hd.version == 1 else 4          # rANS streams are >= 4 bytes, arithmetic-coder streams >= 1
    if hd.payload_bytes < per_stream * nstreams or hd.payload_bytes > lo.max_total_bytes:
        raise ValueError("B2KV header: payload_bytes impossible for this shape")


@dataclass
class EncodedBatch:
    """Device-resident result of one encode call: n containers at `stride` in `buf`."""
    buf: torch.Tensor            # uint8 device staging
    stride: int
    sizes: List[int]             # total bytes per container (host, valid after the call returns)
    max_dtype: int = 0           # dtype code of the stored row maxima (== input dtype)
    coder: int = N.CODER_RANS    # which entropy coder filled the payloads (container version - 1)

    def container(self, j: int) -> torch.Tensor:
        return self.buf[j * self.stride: j * self.stride + self.sizes[j]]


@dataclass
class EncodeTicket:
    """An encode in flight (CacheGenCodec.encode_async).  `wait()` blocks the calling host thread -- not the stream --
    until the kernels are done and returns the batch with its sizes."""
    buf: torch.Tensor
    stride: int
    n_chunks: int
    sizes_buf: "PinnedBuffer"
    event: torch.cuda.Event
    max_dtype: int
    coder: int
    keep: object = None          # keeps the source view (and through it the KV tensors) alive until the kernels ran
    codec: object = None         # told the measured bits per symbol (picks the next call's kernel variant)
    stats: tuple = (0, 0.0)      # (fixed bytes per container, symbols in this call)

    def wait(self) -> EncodedBatch:
        self.event.synchronize()
        self.keep = None
        sizes = list((ctypes.c_uint64 * self.n_chunks).from_address(self.sizes_buf.host_ptr))
        for j, s in enumerate(sizes):
            if s < N.HEADER_BYTES or s > self.stride:
                raise N.NativeError(f"encoder produced an invalid container size {s} for chunk {j}")
        if self.codec is not None and self.stats[1] > 0:
            self.codec._last_bits_per_symbol = 8.0 * (sum(sizes) - self.n_chunks * self.stats[0]) / self.stats[1]
            self.codec = None
        return EncodedBatch(self.buf, self.stride, [int(s) for s in sizes], self.max_dtype, self.coder)


class CacheGenCodec:
    """Batched CacheGen encode / decode on the current CUDA device.

    Thread model: one encoder and one decoder may run concurrently from different threads (the
    reference's put_worker / deserialize_worker, remote_backend.py:61-69,234-246); each direction owns
    its buffers and is guarded by its own lock.
    """

    def __init__(self, model_name: str, coder: Optional[str] = None):
        """coder: "rans_compact" (container version 3, the default: rANS payload + symbol counts instead of CDF rows;
        chunks of more than 256 tokens fall back to version 2), "rans" (version 2) or "ac" (version 1, the
        torchac-lineage arithmetic coder); the environment variable LMCACHE_B200_CODER overrides the default.
        Decoding accepts all three."""
        import os
        from lmcache_b200.storage_backend.serde.cachegen_basics import CacheGenConfig
        N.require_cuda()
        name = (coder or os.environ.get("LMCACHE_B200_CODER", "rans_compact")).lower()
        if name not in N.CODERS:
            raise ValueError(f"unknown coder {name!r} (expected one of {sorted(N.CODERS)})")
        self.coder = N.CODERS[name]
        self.config = CacheGenConfig.from_model_name(model_name)
        kb, vb = self.config.key_bins_list(), self.config.value_bins_list()
        self.nlayers = len(kb)
        self._kb = N.float_array(kb)
        self._vb = N.float_array(vb)
        self._nb = (N.nb_map(kb, vb, len(kb)))          # keys then values, all layers of the model
        self._enc_lock = threading.RLock()
        self._dec_lock = threading.Lock()
        self._enc_event: Optional[torch.cuda.Event] = None
        self._last_bits_per_symbol = 0.0        # payload bits per symbol of the most recent encode whose sizes were read
        self._enc_ws: Optional[torch.Tensor] = None
        self._dec_ws: Optional[torch.Tensor] = None
        self._enc_out: Optional[torch.Tensor] = None
        self._sizes: Optional[PinnedBuffer] = None
        self._dec_in: Optional[torch.Tensor] = None
        self._dec_event: Optional[torch.cuda.Event] = None
        self._dec_status: Optional[PinnedBuffer] = None   # uint32 per chunk of the last decode call (mapped host memory)
        self._dec_status_n = 0
        self._pin_lock = threading.Lock()
        self._pin_in_lock = threading.Lock()
        self._pin_out: Optional[PinnedBuffer] = None      # containers on their way out (encode_to_pinned)
        self._pin_in: Optional[PinnedBuffer] = None       # containers on their way in (pinned_staging)

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _grow(t: Optional[torch.Tensor], nbytes: int, device) -> torch.Tensor:
        if t is None or t.numel() < nbytes or t.device != device:
            t = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        return t

    def coder_for(self, chunk_tokens: int) -> int:
        """The container this codec writes for chunks of `chunk_tokens`: the compact one holds <= 256 tokens."""
        if self.coder == N.CODER_RANS_COMPACT and chunk_tokens > N.GROUP_TOKENS:
            return N.CODER_RANS
        return self.coder

    def layout(self, L: int, H: int, D: int, chunk_tokens: int) -> "N.Layout":
        return N.container_layout(L, H, D, chunk_tokens, self.coder_for(chunk_tokens))

    def accepts(self, hd: "N.Header") -> bool:
        """Can this codec decode the container?  A compact container must have been written with this model's bins."""
        if hd.version != 3:
            return True
        n = self.nlayers
        return hd.L <= n and hd.nb == self._nb[:hd.L] + self._nb[n:n + hd.L]

    def max_container_bytes(self, L: int, H: int, D: int, chunk_tokens: int) -> int:
        """Upper bound of a container of ANY version this codec can decode (what a receive slab must reserve when the
        writer may have been configured differently): version 2's sections are the largest."""
        lo = N.container_layout(L, H, D, chunk_tokens)
        if chunk_tokens <= N.GROUP_TOKENS:
            return (lo.fixed_bytes + 2 * L * H * D * (chunk_tokens + 4) + 16 + 15) & ~15
        return lo.max_total_bytes

    def out_stride(self, L: int, H: int, D: int, chunk_tokens: int) -> int:
        """Bytes reserved per container.  Chunks of <= 256 tokens are coded with their own empirical CDF,
        so a stream costs <= 8 bits/symbol (+ flush); larger chunks may reach 16 bits/symbol."""
        lo = self.layout(L, H, D, chunk_tokens)
        if chunk_tokens <= N.GROUP_TOKENS:
            hdr = N.HDR_MAX if self.coder_for(chunk_tokens) == N.CODER_RANS_COMPACT else 0
            return (lo.fixed_bytes + 2 * L * H * D * (chunk_tokens + 4 + hdr) + 16 + 15) & ~15
        return lo.max_total_bytes

    # ------------------------------------------------------------------ encode
    def encode_async(self, view: KvView, tok_begin: int, n_tokens: int, chunk_size: int,
                     stream: Optional[torch.cuda.Stream] = None, out: Optional[torch.Tensor] = None,
                     sizes: Optional[PinnedBuffer] = None) -> "EncodeTicket":
        """Enqueue the encode of tokens [tok_begin, tok_begin + n_tokens) of `view` as ceil(n_tokens / chunk_size)
        containers on `stream` and return at once: no host synchronisation.  The containers land in `out` (device,
        `out_stride` apart; the codec's own staging when None) and their sizes in `sizes` (mapped page-locked memory,
        8 bytes per chunk; the codec's own when None) -- both are valid once the ticket's event has completed.
        The KV is read in stream order, so the caller may reuse it for later work on the same stream (this is the
        snapshot a non-blocking store needs; reference cache_engine.py:274-275 materialises chunk copies instead)."""
        if n_tokens <= 0:
            raise ValueError("n_tokens must be positive")
        if view.L > self.nlayers:
            raise ValueError(f"KV has {view.L} layers but the bin table of this model has {self.nlayers}")
        n_chunks = (n_tokens + chunk_size - 1) // chunk_size
        last = n_tokens - (n_chunks - 1) * chunk_size
        stride = self.out_stride(view.L, view.H, view.D, chunk_size)
        lib = N.lib()
        with self._enc_lock, torch.cuda.device(view.device):
            tstream = stream if stream is not None else torch.cuda.current_stream()
            coder = self.coder_for(chunk_size)
            ws_bytes = lib.b200kv_encode_workspace_bytes(view.L, view.H, view.D, chunk_size, n_chunks, coder)
            own_out, own_sizes = out is None, sizes is None
            need_out = stride * n_chunks + N.READ_SLACK if own_out else 0
            # the workspace (and the codec's own staging) are shared by consecutive calls: order after the previous
            # encode on whatever stream it ran; never free a buffer a kernel may still be using
            if self._enc_event is not None:
                grow = (self._enc_ws is None or self._enc_ws.numel() < ws_bytes or self._enc_ws.device != view.device or
                        (own_out and (self._enc_out is None or self._enc_out.numel() < need_out)))
                if grow:
                    self._enc_event.synchronize()
                else:
                    tstream.wait_event(self._enc_event)
            self._enc_ws = self._grow(self._enc_ws, ws_bytes, view.device)
            if own_out:
                self._enc_out = self._grow(self._enc_out, need_out, view.device)
                out = self._enc_out
            elif out.numel() < stride * n_chunks:
                raise ValueError("encode output buffer too small")
            if own_sizes:
                if self._sizes is None or self._sizes.nbytes < 8 * n_chunks:
                    self._sizes = PinnedBuffer(max(4096, 8 * n_chunks))
                sizes = self._sizes
            elif sizes.nbytes < 8 * n_chunks:
                raise ValueError("sizes buffer too small")
            # KV statistics of one model are stable from call to call: the previous call's measured entropy picks the
            # encode kernel variant for this one (byte-identical output either way)
            # (the thresholds are in coder bits per symbol; a version-3 payload also holds ~0.4 bits of stream headers)
            b = self._last_bits_per_symbol - (0.4 if coder == N.CODER_RANS_COMPACT else 0.0)
            flags = coder | (N.ENCODE_HINT_HIGH_ENTROPY if b > 2.7 else 0) | (N.ENCODE_HINT_MID_ENTROPY if b > 1.2 else 0)
            N.check(lib.b200kv_encode_chunks(ctypes.byref(view.desc), tok_begin, n_chunks, chunk_size, last,
                                             self._kb, self._vb, flags, out.data_ptr(), stride, sizes.dev_ptr,
                                             self._enc_ws.data_ptr(), self._enc_ws.numel(), tstream.cuda_stream),
                    "encode_chunks")
            ev = torch.cuda.Event()
            ev.record(tstream)
            self._enc_event = ev
            fixed = self.layout(view.L, view.H, view.D, chunk_size).fixed_bytes
            return EncodeTicket(out, stride, n_chunks, sizes, ev, int(view.desc.dtype), coder, view, self,
                                (fixed, 2.0 * view.L * view.H * view.D * n_tokens))

    def encode(self, view: KvView, tok_begin: int, n_tokens: int, chunk_size: int,
               stream: Optional[torch.cuda.Stream] = None, out: Optional[torch.Tensor] = None) -> EncodedBatch:
        """encode_async + one event wait: blocks until the containers' sizes are known; payloads stay on the device.
        With out=None the batch aliases the codec's staging, which the next encode call overwrites."""
        return self.encode_async(view, tok_begin, n_tokens, chunk_size, stream, out).wait()

    def encode_to_host(self, view: KvView, tok_begin: int, n_tokens: int, chunk_size: int,
                       stream: Optional[torch.cuda.Stream] = None) -> List[bytes]:
        """encode + one device->host copy per container through the codec's page-locked slab, returned as immutable
        bytes (the Serializer.to_bytes contract, serde.py:12-27).  The encoder lock is held until the copies are done:
        the staging the batch aliases cannot be overwritten by a concurrent encode."""
        with self._enc_lock, self.encode_to_pinned(view, tok_begin, n_tokens, chunk_size, stream) as views:
            return [bytes(v) for v in views]

    @contextlib.contextmanager
    def encode_to_pinned(self, view: KvView, tok_begin: int, n_tokens: int, chunk_size: int,
                         stream: Optional[torch.cuda.Stream] = None):
        """encode + one device->host copy per container into the codec's page-locked slab (kept across calls, grown on
        demand); yields one writable memoryview per container.  The views -- e.g. handed to a socket send -- are valid
        inside the `with` block only: the slab is reused by the next call (serialised by a lock)."""
        with self._enc_lock, self._pin_lock:
            batch = self.encode(view, tok_begin, n_tokens, chunk_size, stream)
            total = sum((s + 15) & ~15 for s in batch.sizes)
            if self._pin_out is None or self._pin_out.nbytes < total:
                if self._pin_out is not None:
                    self._pin_out.close()
                self._pin_out = PinnedBuffer(max(total, 1) * 5 // 4)
            pin = self._pin_out
            lib = N.lib()
            sp = _stream_ptr(stream)
            offs, o = [], 0
            with torch.cuda.device(view.device):
                for j, s in enumerate(batch.sizes):
                    N.check(lib.b200kv_copy_async(pin.host_ptr + o, batch.buf.data_ptr() + j * batch.stride, s, sp), "copy")
                    offs.append(o)
                    o += (s + 15) & ~15
                N.check(lib.b200kv_stream_sync(sp), "stream_sync")
            views = [pin.view(offs[j], batch.sizes[j]) for j in range(len(offs))]
            for v in views:
                parse_header(v)        # raises on encoder error status
            try:
                yield views
            finally:
                del views

    @contextlib.contextmanager
    def pinned_staging(self, nbytes: int):
        """A page-locked receive slab of at least nbytes (kept across calls): a remote tier reads containers straight
        into it and decode() uploads from it with true asynchronous copies."""
        with self._pin_in_lock:
            if self._pin_in is None or self._pin_in.nbytes < nbytes:
                if self._pin_in is not None:
                    self._dec_sync()
                    self._pin_in.close()
                self._pin_in = PinnedBuffer(max(nbytes, 1) * 5 // 4)
            try:
                yield self._pin_in
            finally:
                self._dec_sync()       # the uploads out of the slab must finish before the next user overwrites it

    def _dec_sync(self) -> None:
        if self._dec_event is not None:
            self._dec_event.synchronize()

    # ------------------------------------------------------------------ decode
    def _order_decode(self, tstream, need_in: int, need_ws: int) -> None:
        """staging / workspace are reused across calls: order after the previous decode and never free a
        buffer a kernel may still be reading."""
        if self._dec_event is None:
            return
        if ((self._dec_in is not None and self._dec_in.numel() < need_in) or
                (self._dec_ws is not None and self._dec_ws.numel() < need_ws)):
            self._dec_event.synchronize()
        else:
            tstream.wait_event(self._dec_event)

    def decode_raw(self, base_ptr: int, buf_bytes: int, offsets: Sequence[int], totals: Sequence[int],
                   ntokens: Sequence[int], dst: KvView, dst_tok: Sequence[int], max_dtype: int, coder: int,
                   stream: Optional[torch.cuda.Stream] = None, _locked: bool = False) -> None:
        """Decode containers that already sit in device memory at base_ptr + offsets[j] (asynchronous).  `buf_bytes` is
        the size of the buffer behind base_ptr: it must extend N.READ_SLACK bytes past every container (checked by the
        library); totals[j] = header.total_bytes."""
        n = len(offsets)
        if n == 0:
            return
        lib = N.lib()
        tmax = max(ntokens)

        def run():
            tstream = stream if stream is not None else torch.cuda.current_stream()
            ws_bytes = lib.b200kv_decode_workspace_bytes(dst.L, dst.H, dst.D, tmax, n)
            if not _locked:
                self._order_decode(tstream, 0, ws_bytes)
            self._dec_ws = self._grow(self._dec_ws, ws_bytes, dst.device)
            if self._dec_status is None or self._dec_status.nbytes < 4 * n:
                if self._dec_status is not None:
                    self._dec_sync()
                self._dec_status = PinnedBuffer(max(4096, 8 * n))
            self._dec_status_n = n
            N.check(lib.b200kv_decode_chunks(base_ptr, int(buf_bytes), N.i64_array(list(offsets)),
                                             N.i64_array(list(totals)), N.i32_array(list(ntokens)),
                                             N.i64_array(list(dst_tok)), n, int(max_dtype), int(coder),
                                             ctypes.byref(dst.desc), self._kb, self._vb, self._dec_status.dev_ptr,
                                             self._dec_ws.data_ptr(), self._dec_ws.numel(), tstream.cuda_stream),
                    "decode_chunks")
            if self._dec_event is None:
                self._dec_event = torch.cuda.Event()
            self._dec_event.record(tstream)

        if _locked:
            run()
        else:
            with self._dec_lock, torch.cuda.device(dst.device):
                run()

    def decode_status(self) -> List[int]:
        """Wait for the most recent decode call and return its per-chunk status words (0 = clean; bit 0: a rANS stream
        did not return to its initial state, bit 1: stream offsets beyond the payload).  A nonzero word means the
        container's bytes were damaged after its header was written: treat the chunk as a miss."""
        with self._dec_lock:
            if self._dec_status is None or self._dec_event is None:
                return []
            self._dec_event.synchronize()
            return list((ctypes.c_uint32 * self._dec_status_n).from_address(self._dec_status.host_ptr))

    def decode_device_batch(self, batch: EncodedBatch, ntokens: Sequence[int], dst: KvView, dst_tok: Sequence[int],
                            stream: Optional[torch.cuda.Stream] = None) -> None:
        """Decode an EncodedBatch straight from its device staging buffer (no host hop, no header reads)."""
        self.decode_raw(batch.buf.data_ptr(), batch.buf.numel(), [j * batch.stride for j in range(len(batch.sizes))],
                        batch.sizes, ntokens, dst, dst_tok, batch.max_dtype, batch.coder, stream)

    def decode(self, containers: Sequence[Union[bytes, bytearray, memoryview, torch.Tensor]], dst: KvView,
               dst_tok: Sequence[int], stream: Optional[torch.cuda.Stream] = None) -> None:
        """Decode containers into `dst` at token offsets `dst_tok` (asynchronous on `stream`).
        Host containers are uploaded first; a single 16-byte-aligned device tensor is used in place."""
        n = len(containers)
        if n == 0:
            return
        lib = N.lib()
        heads = []
        for c in containers:
            if isinstance(c, torch.Tensor):
                hb = c[:N.HEADER_BYTES + N.MAX_PLANES].cpu().numpy().tobytes()
                hd = N.Header.from_buffer_copy(hb[:N.HEADER_BYTES])
                if hd.magic != N.MAGIC or hd.version not in (1, 2, 3) or hd.status != 0 or hd.total_bytes > c.numel():
                    raise ValueError("bad B2KV container tensor")
                check_header(hd, list(hb[N.HEADER_BYTES:N.HEADER_BYTES + 2 * hd.L]) if hd.version == 3 else None)
            else:
                hd = parse_header(c)
            if not self.accepts(hd):
                raise ValueError("compact container written with another model's bins")
            if (hd.L, hd.H, hd.D) != (dst.L, dst.H, dst.D):
                raise ValueError(f"container shape L/H/D={hd.L}/{hd.H}/{hd.D} does not match destination "
                                 f"{dst.L}/{dst.H}/{dst.D}")
            heads.append(hd)
        max_dtype = heads[0].max_dtype
        if any(h.max_dtype != max_dtype or h.version != heads[0].version for h in heads):
            raise ValueError("containers of one decode call must share max_dtype and container version")
        coder = int(heads[0].version) - 1
        totals = [int(h.total_bytes) for h in heads]
        ntoks = [int(h.ntokens) for h in heads]
        tmax = max(ntoks)
        for tok, nt in zip(dst_tok, ntoks):
            if tok < 0 or tok + nt > dst.ntokens:
                raise ValueError(f"container of {nt} tokens at offset {tok} does not fit a {dst.ntokens}-token destination")
        with self._dec_lock, torch.cuda.device(dst.device):
            tstream = stream if stream is not None else torch.cuda.current_stream()
            sp = tstream.cuda_stream
            need_in = sum((int(h.total_bytes) + 15) & ~15 for h in heads) + N.READ_SLACK
            self._order_decode(tstream, need_in, lib.b200kv_decode_workspace_bytes(dst.L, dst.H, dst.D, tmax, n))
            if n == 1 and isinstance(containers[0], torch.Tensor) and containers[0].is_cuda \
                    and containers[0].data_ptr() % 16 == 0 and containers[0].numel() >= totals[0] + N.READ_SLACK:
                keep_dev = containers[0]
                self.decode_raw(keep_dev.data_ptr(), keep_dev.numel(), [0], totals, ntoks, dst, dst_tok, max_dtype, coder,
                                tstream, _locked=True)
                return
            self._dec_in = self._grow(self._dec_in, need_in, dst.device)
            base_ptr = self._dec_in.data_ptr()
            offsets, o = [], 0
            for c, h in zip(containers, heads):
                nb = int(h.total_bytes)
                if isinstance(c, torch.Tensor):
                    keep = c
                    src_ptr = c.data_ptr()
                else:
                    keep = np.frombuffer(c, dtype=np.uint8, count=nb)   # zero-copy view of bytes/bytearray/memoryview
                    src_ptr = keep.ctypes.data
                # pageable sources are staged by the driver before the call returns
                N.check(lib.b200kv_copy_async(base_ptr + o, src_ptr, nb, sp), "copy")
                del keep
                offsets.append(o)
                o += (nb + 15) & ~15
            self.decode_raw(base_ptr, self._dec_in.numel(), offsets, totals, ntoks, dst, dst_tok, max_dtype, coder, tstream,
                            _locked=True)
