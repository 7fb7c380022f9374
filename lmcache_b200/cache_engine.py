"""LMCacheEngine -- store()/retrieve() with the reference's signatures and semantics
(lmcache/cache_engine.py:16-436), rebuilt around the CUDA hot path:

  * chunk hashes: one b200kv_sha256_chain launch on the token ids (no per-chunk tokens.cpu() sync,
    cache_engine.py:58-96); digests are bit-identical to hashlib's.
  * blob pack + chunking: one b200kv_pack_chunks gather from the caller's 2L tensors straight into per-chunk
    blobs (replaces 3x torch.stack + permute + split + .contiguous() per chunk, :98-161).
  * retrieve assembles chunks into one preallocated blob (replaces torch.cat, :362-368).
Prefix-match / mask semantics are unchanged: chunks are matched front to back through `contains`, the first
miss ends the match, everything after the first missing chunk is (re)stored (:183-208).
"""
from __future__ import annotations

import ctypes
import os
import threading
import time
from typing import Dict, Iterable, List, Optional, Tuple, Union

import torch

from lmcache_b200 import _native as N
from lmcache_b200.codec import KvView, PinnedBuffer
from lmcache_b200.config import LMCacheEngineConfig, LMCacheEngineMetadata
from lmcache_b200.logging import init_logger
from lmcache_b200.storage_backend import CreateStorageBackend
from lmcache_b200.utils import CacheEngineKey, KVCache, _lmcache_nvtx_annotate

logger = init_logger(__name__)


class LazySeq:
    """A read-only sequence whose items are computed on access: fn(base[i]).  Slices stay lazy.  The engine passes its
    chunk keys around in this form: item i only exists once the hash chain has reached chunk i (38 us per chunk), and a
    consumer that walks the sequence front to back -- look up / fetch / encode wave by wave -- overlaps with the chain
    instead of waiting for its end."""

    def __init__(self, fn, base, start: int = 0, stop: Optional[int] = None):
        self._fn, self._base = fn, base
        self._start = start
        self._stop = len(base) if stop is None else stop

    def __len__(self) -> int:
        return self._stop - self._start

    def __getitem__(self, i):
        if isinstance(i, slice):
            a, b, step = i.indices(len(self))
            if step != 1:
                return [self[k] for k in range(a, b, step)]
            return LazySeq(self._fn, self._base, self._start + a, self._start + max(a, b))
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        v = self._base[self._start + i]
        return v if self._fn is None else self._fn(v)

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


class _HashRun:
    """One hash-chain launch: digests and their ready words land in mapped page-locked memory; item i blocks (a short
    spin on the ready word) until the kernel has produced digest i."""
    _pool: List[PinnedBuffer] = []
    _pool_lock = threading.Lock()
    _epoch = 0
    _streams: dict = {}

    def __init__(self, dev_tokens: torch.Tensor, chunk_size: int, offs: List[int], nchunks: int):
        self.n = nchunks
        need = 36 * nchunks                                  # 32-byte digests, then one ready word each
        with _HashRun._pool_lock:
            _HashRun._epoch = (_HashRun._epoch % 0x7fffffff) + 1
            self.epoch = _HashRun._epoch
            fit = [b for b in _HashRun._pool if b.nbytes >= need]
            if fit:
                self.buf = min(fit, key=lambda b: b.nbytes)
                _HashRun._pool.remove(self.buf)
            else:
                self.buf = PinnedBuffer(max(1 << 14, 2 * need))     # cudaHostAlloc zero-fills: no stale epoch inside
        self.cap = self.buf.nbytes // 36
        self._tokens = dev_tokens                            # alive until the kernel has run
        self._hex: List[Optional[str]] = [None] * nchunks
        self._flags = (ctypes.c_uint32 * self.cap).from_address(self.buf.host_ptr + 32 * self.cap)
        dev = dev_tokens.device
        with torch.cuda.device(dev):
            # its own stream: the chain is one warp on one SM; the caller's encode / decode kernels must not queue behind it
            side = _HashRun._streams.get(dev.index)
            if side is None:
                side = _HashRun._streams[dev.index] = torch.cuda.Stream(device=dev)
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream())
            side.wait_event(ready)
            N.check(N.lib().b200kv_sha256_chain_ready(ctypes.c_void_p(dev_tokens.data_ptr()), dev_tokens.element_size(),
                                                      N.i64_array(offs), len(offs) - 1, chunk_size,
                                                      ctypes.c_void_p(self.buf.dev_ptr),
                                                      ctypes.c_void_p(self.buf.dev_ptr + 32 * self.cap), self.epoch,
                                                      side.cuda_stream), "sha256_chain")
            self.done = torch.cuda.Event()
            self.done.record(side)
            dev_tokens.record_stream(side)

    def __len__(self) -> int:
        return self.n

    def __getitem__(self, i: int) -> str:
        h = self._hex[i]
        if h is None:
            spins = 0
            while self._flags[i] != self.epoch:
                spins += 1
                if spins % 4096 == 0 and self.done.query():
                    # the kernel is over: either the word is there now, or the launch failed
                    if self._flags[i] != self.epoch:
                        self.done.synchronize()
                        raise N.NativeError("hash chain kernel finished without producing every digest")
            h = self._hex[i] = bytes(self.buf.view(32 * i, 32)).hex()
        return h

    def __del__(self):
        try:
            if not self.done.query():
                self.done.synchronize()                      # the kernel still writes into the buffer
            with _HashRun._pool_lock:
                if len(_HashRun._pool) < 8:
                    _HashRun._pool.append(self.buf)
                    self.buf = None
            if self.buf is not None:
                self.buf.close()
        except Exception:       # noqa: BLE001 -- interpreter shutdown
            pass


def sha256_prefix_chain_lazy(tokens: torch.Tensor, chunk_size: int, seq_offsets: Optional[List[int]] = None) -> LazySeq:
    """sha256_prefix_chain without waiting for the end of the chain: returns at once with a lazy sequence of hex digests;
    digest i becomes available ~38 us x (i + 1) after the launch (one 256-token chunk of int64 ids)."""
    N.require_cuda()
    if tokens.dim() != 1:
        raise ValueError(f"Invalid shape of tokens: {tokens.shape}")
    n = tokens.shape[0]
    offs = [0, n] if seq_offsets is None else list(seq_offsets)
    n_seq = len(offs) - 1
    nchunks = sum((offs[i + 1] - offs[i] + chunk_size - 1) // chunk_size for i in range(n_seq))
    if nchunks == 0:
        return LazySeq(None, [])
    dev_tokens = tokens if tokens.is_cuda else tokens.to("cuda", non_blocking=False)
    return LazySeq(None, _HashRun(dev_tokens.contiguous(), chunk_size, offs, nchunks))


def sha256_prefix_chain(tokens: torch.Tensor, chunk_size: int, seq_offsets: Optional[List[int]] = None) -> List[str]:
    """Hex digests h_i = sha256(hex(h_{i-1}) || bytes(chunk_i)) for every chunk of every sequence
    (cache_engine.py:58-96), computed on the GPU.  tokens: 1-D integer tensor on any device; the bytes hashed
    are the tensor's native little-endian dtype, as in the reference.  seq_offsets: token boundaries of
    independent sequences (default: one sequence)."""
    return list(sha256_prefix_chain_lazy(tokens, chunk_size, seq_offsets))


class LMCacheEngine:

    def __init__(self, config: LMCacheEngineConfig, metadata: LMCacheEngineMetadata):
        self.config = config
        self.metadata = metadata
        self.chunk_size = config.chunk_size
        self.save_decode_cache = config.save_decode_cache
        self.engine_ = CreateStorageBackend(config, metadata)
        logger.debug(f"Current storage backend type {type(self.engine_)}")

    # ------------------------------------------------------------------ keys / hashes
    def _make_key(self, chunk_hash: str, fmt: str) -> CacheEngineKey:
        return CacheEngineKey(fmt, self.metadata.model_name, self.metadata.world_size, self.metadata.worker_id,
                              chunk_hash)

    def _num_tokens_in_kv(self, kv_tensors: Union[KVCache, torch.Tensor], fmt: str) -> int:
        if fmt == "huggingface":
            return kv_tensors[0][0].shape[1]
        elif fmt == "vllm":
            return kv_tensors[0][0].shape[0]
        raise ValueError(f"Invalid format: {fmt}")

    def _get_init_hash(self) -> str:
        return ""

    def _prefix_hash(self, tokens: torch.Tensor, num_skip_chunk: Optional[int] = 0):
        """All chunk digests of `tokens` (the whole chain is hashed, then the first num_skip_chunk digests are
        dropped, like cache_engine.py:86-96).  A lazy sequence: digest i is there once the chain has reached chunk i."""
        hashes = sha256_prefix_chain_lazy(tokens, self.chunk_size)[num_skip_chunk or 0:]
        if os.environ.get("LMCACHE_B200_EAGER_KEYS") == "1":      # measurement knob: wait for the whole chain first
            return list(hashes)
        return hashes

    def _keys_of(self, chunk_hashes, fmt: str) -> LazySeq:
        return LazySeq(lambda h: self._make_key(h, fmt), chunk_hashes)

    # ------------------------------------------------------------------ blob helpers
    def _chunk_shape(self, view: KvView, t: int, fmt: str) -> Tuple[int, ...]:
        return (view.L, 2, t, view.H, view.D) if fmt == "vllm" else (view.L, 2, view.H, t, view.D)

    def _pack_chunks(self, view: KvView, tok_begin: int, fmt: str) -> List[torch.Tensor]:
        """Gather tokens [tok_begin, T) of the kv tuple into contiguous per-chunk blobs with one kernel."""
        n_tok = view.ntokens - tok_begin
        if n_tok <= 0:
            return []
        cs = self.chunk_size
        n_chunks = (n_tok + cs - 1) // cs
        last = n_tok - (n_chunks - 1) * cs
        per_tok = 2 * view.L * view.H * view.D          # halfs per token over all planes
        stride_elems = per_tok * cs
        buf = torch.empty(n_chunks * stride_elems, dtype=view.dtype, device=view.device)
        with torch.cuda.device(view.device):
            N.check(N.lib().b200kv_pack_chunks(ctypes.byref(view.desc), tok_begin, n_chunks, cs, last,
                                               1 if fmt == "huggingface" else 0, ctypes.c_void_p(buf.data_ptr()),
                                               stride_elems * buf.element_size(),
                                               torch.cuda.current_stream().cuda_stream), "pack_chunks")
        chunks = []
        for j in range(n_chunks):
            t = cs if j < n_chunks - 1 else last
            chunks.append(buf[j * stride_elems: j * stride_elems + per_tok * t].view(self._chunk_shape(view, t, fmt)))
        return chunks

    def _pack_chunks_torch(self, kv: KVCache, tok_begin: int, fmt: str) -> List[torch.Tensor]:
        """_tuple_kv_to_blob + _slice_kv_at with torch ops (cache_engine.py:98-161), for dtypes the kernels do not move"""
        k = torch.stack([x[0] for x in kv])
        v = torch.stack([x[1] for x in kv])
        blob = torch.stack((k, v)).permute(1, 0, 2, 3, 4)
        tdim = 2 if fmt == "vllm" else 3
        blob = blob.narrow(tdim, tok_begin, blob.shape[tdim] - tok_begin)
        return [c.contiguous() for c in torch.split(blob, self.chunk_size, dim=tdim)]

    @staticmethod
    def _as_cuda_kv(kv_tensors_raw: KVCache) -> KVCache:
        if kv_tensors_raw[0][0].is_cuda:
            return kv_tensors_raw
        return tuple((k.cuda(), v.cuda()) for k, v in kv_tensors_raw)

    def _blob_to_tuple_kv(self, blob: torch.Tensor) -> KVCache:
        return tuple((layer[0], layer[1]) for layer in torch.unbind(blob, dim=0))

    # ------------------------------------------------------------------ store
    @_lmcache_nvtx_annotate
    @torch.no_grad()
    def store(self, tokens: torch.Tensor, kv_tensors_raw: KVCache, skip_existing=True, blocking=True) -> None:
        """Store the KV cache of `tokens`.  kv_tensors_raw: nested tuple of per-layer (K, V), each
        [num_tokens, num_heads, head_size] (vllm) or [num_heads, num_tokens, head_size] (huggingface),
        without a batch dimension."""
        start_time = time.perf_counter()
        fmt = self.metadata.fmt
        assert len(tokens.shape) == 1, f"Invalid shape of tokens: {tokens.shape}"
        assert len(kv_tensors_raw) > 0, "Empty kv_tensors"
        assert len(tokens) == self._num_tokens_in_kv(kv_tensors_raw, fmt), \
            "Number of tokens in the kv cache does not match the input tokens"

        chunk_hashes = self._prefix_hash(tokens)
        start_chunk_idx = 0
        if skip_existing:
            # prefix match: first chunk whose key is absent; everything from there on is stored
            start_chunk_idx = len(chunk_hashes)
            for i, h in enumerate(chunk_hashes):
                if not self.engine_.contains(self._make_key(h, fmt)):
                    start_chunk_idx = i
                    break
        n_chunks = 0
        if start_chunk_idx < len(chunk_hashes):
            keys = self._keys_of(chunk_hashes[start_chunk_idx:], fmt)
            kv_cuda = self._as_cuda_kv(kv_tensors_raw)
            if kv_cuda[0][0].dtype not in (torch.bfloat16, torch.float16):
                # the native pack / codec kernels move 16-bit KV; any other dtype (the reference's local and torch-serde
                # paths accept every dtype) takes the reference's own blob ops on the GPU (cache_engine.py:98-161)
                chunks = self._pack_chunks_torch(kv_cuda, start_chunk_idx * self.chunk_size, fmt)
                end_make_chunks = time.perf_counter()
                n_chunks = self.engine_.batched_put(zip(keys, chunks), blocking=blocking)
                logger.info(f"Stored/updated {n_chunks} chunks, total time {time.perf_counter() - start_time:.2f}s, "
                            f"make chunks time {end_make_chunks - start_time:.2f}s")
                return
            view = KvView.from_tuple(kv_cuda, fmt)
            self._geom = (view.L, view.H, view.D, view.dtype)
            if self._fast_path():
                # B200-native path: the backend consumes the caller's 2L tensors directly (batched encode / one gather)
                end_make_chunks = time.perf_counter()
                n_chunks = self.engine_.put_kv_chunks(keys, view, start_chunk_idx * self.chunk_size, self.chunk_size,
                                                      blocking=blocking)
            else:
                chunks = self._pack_chunks(view, start_chunk_idx * self.chunk_size, fmt)
                end_make_chunks = time.perf_counter()
                n_chunks = self.engine_.batched_put(zip(keys, chunks), blocking=blocking)
        else:
            end_make_chunks = time.perf_counter()
        end_time = time.perf_counter()
        logger.info(f"Stored/updated {n_chunks} chunks, total time {end_time - start_time:.2f}s, "
                    f"make chunks time {end_make_chunks - start_time:.2f}s")

    # ------------------------------------------------------------------ retrieve
    @_lmcache_nvtx_annotate
    @torch.no_grad()
    def retrieve(self, tokens: torch.Tensor, mask: Optional[torch.Tensor] = None) -> Tuple[KVCache, torch.Tensor]:
        """Retrieve the longest cached prefix of `tokens` (optionally only the suffix selected by a boolean
        suffix `mask`).  Returns (kv tuple -- empty tuple on a total miss, ret_mask marking retrieved tokens)."""
        num_skip_chunk = 0
        num_skip_tok = 0
        ret_mask = torch.ones_like(tokens, dtype=torch.bool)
        if mask is not None:
            num_skip_tok = int(len(mask) - torch.sum(mask))
            num_skip_chunk = num_skip_tok // self.chunk_size
        ret_mask[:num_skip_tok] = False

        st = time.perf_counter()
        fmt = self.metadata.fmt
        if fmt not in ("vllm", "huggingface"):
            raise ValueError(f"Invalid format: {fmt}")
        chunk_hashes = self._prefix_hash(tokens, num_skip_chunk)
        if self._fast_path() and len(chunk_hashes) > 0 and not getattr(self, "_wide_dtype", False):
            try:
                return self._retrieve_into_blob(tokens, chunk_hashes, num_skip_tok, num_skip_chunk, ret_mask, fmt, st)
            except TypeError:
                self._wide_dtype = True      # chunks of a dtype the kernels do not move: per-chunk path from now on
        retrieved: List[torch.Tensor] = []
        for chunk in self.engine_.batched_get(self._make_key(h, fmt) for h in chunk_hashes):
            if chunk is None:
                break
            retrieved.append(chunk)
        if len(retrieved) == 0:
            logger.info("Retrieved 0 chunks")
            ret_mask[:] = False
            return (), ret_mask

        # assemble into one blob; drop the extra leading tokens of the first chunk (suffix mask)
        tdim = 2 if fmt == "vllm" else 3
        extra = num_skip_tok - num_skip_chunk * self.chunk_size
        sizes = [c.shape[tdim] for c in retrieved]
        total = sum(sizes) - extra
        first = retrieved[0]
        shape = list(first.shape)
        shape[tdim] = total
        blob = torch.empty(shape, dtype=first.dtype, device=first.device)
        pos = 0
        for i, c in enumerate(retrieved):
            src = c.narrow(tdim, extra, sizes[i] - extra) if i == 0 else c
            n = src.shape[tdim]
            blob.narrow(tdim, pos, n).copy_(src)
            pos += n
        ret = self._blob_to_tuple_kv(blob)
        retrieved_token_count = total
        logger.info(f"Retrieved {len(retrieved)} chunks ({retrieved_token_count} tokens in total) -- "
                    f"elapsed time {time.perf_counter() - st}")
        ret_mask[num_skip_tok + retrieved_token_count:] = False
        return ret, ret_mask

    # ------------------------------------------------------------------ paged KV caches, in place
    @_lmcache_nvtx_annotate
    @torch.no_grad()
    def store_paged(self, tokens: torch.Tensor, kv_caches, slot_mapping: torch.Tensor, skip_existing=True,
                    blocking=True) -> None:
        """store() for a vLLM paged KV cache: token i's K/V live in row slot_mapping[i] of every layer's
        (key_cache, value_cache) [num_blocks, block_size, H, D].  What lmcache-vllm's lmcache_store_kv does with a
        torch gather per layer + store() (LLM_Engine.rst:91-99); here the backend's kernels read the cache rows
        directly (cachegen: quantise + code from the rows; local tiers: one gather straight into the chunk blobs)."""
        if self.metadata.fmt != "vllm":
            raise ValueError(f"paged KV caches use the vllm layout, engine fmt is {self.metadata.fmt}")
        assert len(tokens.shape) == 1, f"Invalid shape of tokens: {tokens.shape}"
        assert len(kv_caches) > 0, "Empty kv_caches"
        assert len(tokens) == slot_mapping.numel(), "Number of slots does not match the input tokens"
        if not self._fast_path():
            flat = [(k.reshape(-1, k.shape[-2], k.shape[-1]), v.reshape(-1, v.shape[-2], v.shape[-1])) for k, v in kv_caches]
            idx = slot_mapping.to(flat[0][0].device)
            return self.store(tokens, tuple((k[idx], v[idx]) for k, v in flat), skip_existing, blocking)
        fmt = "vllm"
        chunk_hashes = self._prefix_hash(tokens)
        start_chunk_idx = 0
        if skip_existing:
            start_chunk_idx = len(chunk_hashes)
            for i, h in enumerate(chunk_hashes):
                if not self.engine_.contains(self._make_key(h, fmt)):
                    start_chunk_idx = i
                    break
        if start_chunk_idx < len(chunk_hashes):
            view = KvView.from_paged(kv_caches, slot_mapping.cuda())
            self._geom = (view.L, view.H, view.D, view.dtype)
            keys = self._keys_of(chunk_hashes[start_chunk_idx:], fmt)
            self.engine_.put_kv_chunks(keys, view, start_chunk_idx * self.chunk_size, self.chunk_size, blocking=blocking)

    @_lmcache_nvtx_annotate
    @torch.no_grad()
    def retrieve_paged(self, tokens: torch.Tensor, kv_caches, slot_mapping: torch.Tensor,
                       mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """retrieve() straight into a paged KV cache: the longest cached prefix of `tokens` (optionally only the
        suffix selected by `mask`) is written to rows slot_mapping[i]; returns ret_mask as retrieve() does.  Rows of
        tokens that are not retrieved -- misses and the masked-off prefix -- are left untouched."""
        if self.metadata.fmt != "vllm":
            raise ValueError(f"paged KV caches use the vllm layout, engine fmt is {self.metadata.fmt}")
        assert len(tokens) == slot_mapping.numel(), "Number of slots does not match the input tokens"
        flat = [(k.reshape(-1, k.shape[-2], k.shape[-1]), v.reshape(-1, v.shape[-2], v.shape[-1])) for k, v in kv_caches]
        dev = flat[0][0].device
        slots = slot_mapping.to(dev)
        if not self._fast_path():
            kv, ret_mask = self.retrieve(tokens, mask)
            if len(kv) > 0:
                idx = slots[ret_mask.to(dev)]
                for (kc, vc), (k, v) in zip(flat, kv):
                    kc[idx] = k.to(kc.dtype)
                    vc[idx] = v.to(vc.dtype)
            return ret_mask
        cs = self.chunk_size
        num_skip_tok = int(len(mask) - torch.sum(mask)) if mask is not None else 0
        num_skip_chunk = num_skip_tok // cs
        extra = num_skip_tok - num_skip_chunk * cs
        ret_mask = torch.ones_like(tokens, dtype=torch.bool)
        ret_mask[:num_skip_tok] = False
        keys = self._keys_of(self._prefix_hash(tokens, num_skip_chunk), "vllm")
        base = num_skip_chunk * cs
        view = KvView.from_paged(kv_caches, slots[base:])
        got_chunks, first = 0, 0
        if extra > 0 and keys:
            # the first chunk straddles the mask: decode it next to the cache and scatter only its unmasked tail
            t0 = min(cs, len(tokens) - base)
            tmp = torch.empty((view.L, 2, t0, view.H, view.D), dtype=view.dtype, device=dev)
            if self.engine_.get_kv_into(keys[:1], KvView.from_blob(tmp, "vllm"), 0, cs) == 0:
                ret_mask[:] = False
                return ret_mask
            idx = slots[base + extra: base + t0]
            for l, (kc, vc) in enumerate(flat):
                kc[idx] = tmp[l, 0, extra:]
                vc[idx] = tmp[l, 1, extra:]
            got_chunks, first = 1, 1
        if len(keys) > first:
            got_chunks += self.engine_.get_kv_into(keys[first:], view, first * cs, cs)
        got = min(base + got_chunks * cs, len(tokens))
        if got <= num_skip_tok:
            ret_mask[:] = False
        else:
            ret_mask[got:] = False
        return ret_mask

    # ------------------------------------------------------------------ B200-native fast paths
    def _fast_path(self) -> bool:
        f = getattr(self.engine_, "supports_kv_view", None)
        return bool(f and f())

    def _kv_geometry(self):
        """(L, H, D, dtype) of this engine's chunks, learnt from the first store / a probe get."""
        return getattr(self, "_geom", None)

    def _retrieve_into_blob(self, tokens, chunk_hashes, num_skip_tok, num_skip_chunk, ret_mask, fmt, st):
        """retrieve() without per-chunk tensors or torch.cat: the backend decodes / uploads every hit chunk straight
        into one preallocated blob; the suffix-mask trim of the first chunk is a view offset, not a copy."""
        keys = self._keys_of(chunk_hashes, fmt)
        geom = self._kv_geometry()
        if geom is None:
            # shapes unknown (nothing stored through this engine yet -- the normal case for a retrieve-only replica):
            # read them from the first chunk's container header / stored blob; only backends without that door pay
            # for a full get of chunk 0
            peek = getattr(self.engine_, "peek_geometry", None)
            if peek is not None:
                geom = peek(keys[0], fmt)
            else:
                first = self.engine_.get(keys[0])
                if first is not None:
                    geom = ((first.shape[0], first.shape[3], first.shape[4]) if fmt == "vllm" else
                            (first.shape[0], first.shape[2], first.shape[4])) + (first.dtype,)
            if geom is None:
                logger.info("Retrieved 0 chunks")
                ret_mask[:] = False
                return (), ret_mask
            self._geom = geom
        L, H, D, dtype = geom
        od = getattr(self.engine_, "out_dtype", None) or getattr(getattr(self.engine_, "deserializer", None), "out_dtype", None)
        if od is not None and od() is not None:
            dtype = od()
        n_tok_max = len(tokens) - num_skip_chunk * self.chunk_size
        shape = (L, 2, n_tok_max, H, D) if fmt == "vllm" else (L, 2, H, n_tok_max, D)
        device = torch.device("cuda", torch.cuda.current_device())
        blob = torch.empty(shape, dtype=dtype, device=device)
        n = self.engine_.get_kv_into(keys, KvView.from_blob(blob, fmt), 0, self.chunk_size)
        if n == 0:
            logger.info("Retrieved 0 chunks")
            ret_mask[:] = False
            return (), ret_mask
        tdim = 2 if fmt == "vllm" else 3
        got = min(n * self.chunk_size, n_tok_max)              # the last hit chunk may be the ragged tail
        extra = num_skip_tok - num_skip_chunk * self.chunk_size
        ret = self._blob_to_tuple_kv(blob.narrow(tdim, extra, got - extra))
        retrieved_token_count = got - extra
        logger.info(f"Retrieved {n} chunks ({retrieved_token_count} tokens in total) -- "
                    f"elapsed time {time.perf_counter() - st}")
        ret_mask[num_skip_tok + retrieved_token_count:] = False
        return ret, ret_mask

    def close(self):
        self.engine_.close()


class LMCacheEngineBuilder:
    """Process-wide engine registry (cache_engine.py:387-436)."""
    _instances: Dict[str, LMCacheEngine] = {}
    _cfgs: Dict[str, LMCacheEngineConfig] = {}
    _metadatas: Dict[str, LMCacheEngineMetadata] = {}

    @classmethod
    def get_or_create(cls, instance_id: str, config: LMCacheEngineConfig,
                      metadata: LMCacheEngineMetadata) -> LMCacheEngine:
        if instance_id not in cls._instances:
            engine = LMCacheEngine(config, metadata)
            cls._instances[instance_id] = engine
            cls._cfgs[instance_id] = config
            cls._metadatas[instance_id] = metadata
            return engine
        if cls._cfgs[instance_id] != config or cls._metadatas[instance_id] != metadata:
            raise ValueError(f"Instance {instance_id} already exists with a different configuration or metadata.")
        return cls._instances[instance_id]

    @classmethod
    def get(cls, instance_id: str) -> Optional[LMCacheEngine]:
        return cls._instances.get(instance_id)

    @classmethod
    def destroy(cls, instance_id: str) -> None:
        if instance_id in cls._instances:
            cls._instances[instance_id].close()
            cls._instances.pop(instance_id, None)
            cls._cfgs.pop(instance_id, None)
            cls._metadatas.pop(instance_id, None)
