"""CacheGen configuration and wire-container views.

Mirrors lmcache/storage_backend/serde/cachegen_basics.py:
  * CACHEGEN_GPU_MAX_TOKENS_PER_CHUNK (:13), CacheGenConfig.from_model_name (:16-78): identical bin table.
  * CacheGenGPUBytestream / CacheGenGPUEncoderOutput (:109-142): same field names, but `to_bytes` /
    `from_bytes` speak the flat "B2KV" container (include/b200kv.h) that the encode kernel writes on the
    device, instead of pickling CUDA tensors.  `from_bytes` gives the same object a reference consumer
    would unpickle (test_serde.py:60-62 reads .num_heads / .head_size).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List

import numpy as np
import torch

from lmcache_b200 import _native as N

CACHEGEN_GPU_MAX_TOKENS_PER_CHUNK = N.GROUP_TOKENS

_FAMILY_LAYERS = {
    "mistralai/Mistral-7B-Instruct-v0.2": 32,
    "lmsys/longchat-7b-16k": 32,
    "Qwen/Qwen-7B": 32,
    "meta-llama/Llama-3.1-8B-Instruct": 32,
    "THUDM/glm-4-9b-chat": 40,
}


@dataclass
class CacheGenConfig:
    key_first_layers: int
    key_second_layers: int
    key_third_layers: int
    key_first_bins: int
    key_second_bins: int
    key_third_bins: int
    value_first_layers: int
    value_first_bins: int
    value_second_bins: int

    def __getitem__(self, key: str) -> int:
        return getattr(self, key)

    @staticmethod
    def from_model_name(model_name: str) -> "CacheGenConfig":
        if model_name not in _FAMILY_LAYERS:
            raise ValueError(f"Model {model_name} is not supported")
        return CacheGenConfig(
            key_first_layers=10, key_second_layers=20, key_third_layers=_FAMILY_LAYERS[model_name],
            key_first_bins=32, key_second_bins=16, key_third_bins=16,
            value_first_layers=2, value_first_bins=32, value_second_bins=16)

    def key_bins_list(self) -> List[float]:
        """make_key_bins (cachegen_encoder.py:339-344): per-layer fp32 bin counts."""
        bins = [float(self.key_third_bins)] * self.key_third_layers
        for i in range(min(self.key_second_layers, self.key_third_layers)):
            bins[i] = float(self.key_second_bins)
        for i in range(min(self.key_first_layers, self.key_third_layers)):
            bins[i] = float(self.key_first_bins)
        return bins

    def value_bins_list(self) -> List[float]:
        """make_value_bins (cachegen_encoder.py:346-350)."""
        bins = [float(self.value_second_bins)] * self.key_third_layers
        for i in range(min(self.value_first_layers, self.key_third_layers)):
            bins[i] = float(self.value_first_bins)
        return bins


@dataclass
class CacheGenGPUBytestream:
    bytestream: torch.Tensor          # uint8 [N]
    bytestream_lengths: torch.Tensor  # int32 [2L, C]
    ntokens: int

    def __getitem__(self, key: str):
        return getattr(self, key)


_HALF = {N.DT_BF16: torch.bfloat16, N.DT_FP16: torch.float16}


@dataclass
class CacheGenGPUEncoderOutput:
    data_chunks: List[CacheGenGPUBytestream]
    cdf: torch.Tensor                 # int16 [2L, C, 33]
    max_tensors_key: torch.Tensor     # half [L, t, 1]
    max_tensors_value: torch.Tensor   # half [L, t, 1]
    num_heads: int
    head_size: int
    coder: int = N.CODER_RANS         # entropy coder of the bytestreams = container version - 1 (not in the reference)

    def __getitem__(self, key: str):
        return getattr(self, key)

    @staticmethod
    def from_bytes(bs) -> "CacheGenGPUEncoderOutput":
        """Parse a B2KV container into host tensors (zero-copy views where possible)."""
        from lmcache_b200.codec import parse_header
        hd = parse_header(bs)
        L, H, D, t, G = hd.L, hd.H, hd.D, hd.ntokens, hd.ngroups
        C = H * D
        lo = N.container_layout(L, H, D, t)
        raw = np.frombuffer(bs, dtype=np.uint8, count=int(hd.total_bytes))

        def section(off, count, dtype):
            return torch.from_numpy(raw[off:off + count * np.dtype(dtype).itemsize].view(dtype).copy())

        cdf = section(lo.off_cdf, 2 * L * C * N.LP, np.int16).reshape(2 * L, C, N.LP)
        maxes = section(lo.off_maxes, 2 * L * t, np.int16).view(_HALF[hd.max_dtype]).reshape(2, L, t, 1)
        lengths = section(lo.off_lengths, G * 2 * L * C, np.int32).reshape(G, 2 * L, C)
        payload = raw[lo.off_payload: lo.off_payload + int(hd.payload_bytes)]
        chunks, pos = [], 0
        for g in range(G):
            nb = int(lengths[g].sum())
            gt = min(N.GROUP_TOKENS, t - g * N.GROUP_TOKENS)
            chunks.append(CacheGenGPUBytestream(torch.from_numpy(payload[pos:pos + nb].copy()), lengths[g], gt))
            pos += nb
        return CacheGenGPUEncoderOutput(chunks, cdf, maxes[0], maxes[1], H, D, int(hd.version) - 1)

    def to_bytes(self) -> bytes:
        """Re-assemble the flat container from the logical fields (host side; used by tests / tools)."""
        L = self.max_tensors_key.shape[0]
        t = self.max_tensors_key.shape[1]
        H, D = self.num_heads, self.head_size
        C = H * D
        lo = N.container_layout(L, H, D, t)
        payload = b"".join(c.bytestream.cpu().numpy().tobytes() for c in self.data_chunks)
        total = lo.off_payload + len(payload)
        buf = bytearray(total)
        hd = N.Header.from_buffer(buf)
        hd.magic, hd.version = N.MAGIC, int(self.coder) + 1
        hd.L, hd.H, hd.D, hd.ntokens, hd.ngroups = L, H, D, t, len(self.data_chunks)
        hd.max_dtype = N.DT_BF16 if self.max_tensors_key.dtype == torch.bfloat16 else N.DT_FP16
        hd.payload_bytes, hd.total_bytes, hd.status = len(payload), total, 0
        del hd

        def put(off, tensor):
            b = tensor.contiguous().cpu().view(torch.uint8).numpy().tobytes()
            buf[off:off + len(b)] = b

        put(lo.off_cdf, self.cdf.reshape(2 * L, C, N.LP))
        put(lo.off_maxes, torch.stack([self.max_tensors_key.reshape(L, t), self.max_tensors_value.reshape(L, t)]))
        put(lo.off_lengths, torch.stack([c.bytestream_lengths.reshape(2 * L, C) for c in self.data_chunks]))
        buf[lo.off_payload:] = payload
        return bytes(buf)


# the legacy (CPU coder) container name the reference test imports (test_serde.py:5,60)
CacheGenEncoderOutput = CacheGenGPUEncoderOutput
