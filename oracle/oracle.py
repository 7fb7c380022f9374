"""numpy/ctypes front-end of the CPU oracle (oracle/cachegen_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (lmcache_b200/) never imports it.

Every function cites the reference file:line it restates; see the C file for details.
Parity status: quantise / dequantise / CDF / hash are pinned against golden vectors made
from the reference's own functions (tests/golden/make_golden.py); the arithmetic-coder
bitstream is "parity unpinned" (torchac_cuda wheel absent) and is pinned only to the
published torchac algorithm restated in SURVEY.md Appendix A.  The rANS coder and the
version-3 stream framing (v3_pack / v3_unpack: every stream carries its symbol histogram,
from which cdf_from_counts rebuilds the reference's CDF tensor -- that function IS pinned
to the reference-made CDF goldens) are this build's own wire format, restated here
independently of the kernels and of the product's host shim.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

LP = 33  # int(bins.max()) + 1, cachegen_encoder.py:287-289
GROUP = 256  # CACHEGEN_GPU_MAX_TOKENS_PER_CHUNK, cachegen_basics.py:13

DT_BF16 = 0
DT_FP16 = 1


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "cachegen_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        i64, i32, vp = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p
        L.oracle_quantize.argtypes = [vp, i32, i32, i32, i32, i64, i64, i64, vp, vp, vp, vp]
        L.oracle_quantize.restype = None
        L.oracle_cdf.argtypes = [vp, i32, i32, i32, vp]
        L.oracle_cdf.restype = None
        L.oracle_counts.argtypes = [vp, i32, i32, i32, vp]
        L.oracle_counts.restype = None
        L.oracle_cdf_from_counts.argtypes = [vp, i64, i32, vp]
        L.oracle_cdf_from_counts.restype = None
        L.oracle_v3_pack.argtypes = [vp, vp, i32, i32, vp, vp, vp, i64, vp]
        L.oracle_v3_pack.restype = i64
        L.oracle_v3_unpack.argtypes = [vp, i64, vp, vp, i32, i32, i32, vp, vp, vp]
        L.oracle_v3_unpack.restype = i64
        L.oracle_encode_group.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, i64, vp]
        L.oracle_encode_group.restype = i64
        L.oracle_decode_group.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp]
        L.oracle_decode_group.restype = None
        L.oracle_encode_group_rans.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, i64, vp]
        L.oracle_encode_group_rans.restype = i64
        L.oracle_decode_group_rans.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp]
        L.oracle_decode_group_rans.restype = i64
        L.oracle_dequantize.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, i32, vp, i64, i64, i64]
        L.oracle_dequantize.restype = None
        L.oracle_sha256_chain.argtypes = [vp, i64, i32, i32, vp]
        L.oracle_sha256_chain.restype = i32
        L.oracle_set_threads.argtypes = [i32]
        L.oracle_set_threads.restype = i32
        _lib = L
    return _lib


def set_threads(n: int) -> int:
    """Use n OpenMP threads (torchrun forces OMP_NUM_THREADS=1); returns the thread count in effect."""
    return int(lib().oracle_set_threads(int(n)))


def _p(a: np.ndarray) -> ctypes.c_void_p:
    return ctypes.c_void_p(a.ctypes.data)


# ----------------------------------------------------------------------------- bins (a5)
_FAMILIES = {
    # cachegen_basics.py:33-78
    "mistralai/Mistral-7B-Instruct-v0.2": 32,
    "lmsys/longchat-7b-16k": 32,
    "Qwen/Qwen-7B": 32,
    "meta-llama/Llama-3.1-8B-Instruct": 32,
    "THUDM/glm-4-9b-chat": 40,
}


def make_bins(model_name: str) -> Tuple[np.ndarray, np.ndarray]:
    """make_key_bins / make_value_bins (cachegen_encoder.py:339-350): fp32 [nlayers]."""
    if model_name not in _FAMILIES:
        raise ValueError(f"Model {model_name} is not supported")
    n = _FAMILIES[model_name]
    kb = np.full(n, 16.0, np.float32)
    kb[:20] = 16.0
    kb[:10] = 32.0
    vb = np.full(n, 16.0, np.float32)
    vb[:2] = 32.0
    return kb, vb


# ----------------------------------------------------------------------------- codec
def quantize(x_bits: np.ndarray, dtype: int, key_bins: np.ndarray, value_bins: np.ndarray):
    """x_bits: uint16 [L, 2, t, C] (bit pattern of bf16/fp16) -> (sym int8 [2L,t,C], maxes u16 [2,L,t]).
    cachegen_encoder.py:40-61,76-91,282-285."""
    x_bits = np.ascontiguousarray(x_bits, dtype=np.uint16)
    L, two, t, C = x_bits.shape
    assert two == 2
    sym = np.empty((2 * L, t, C), np.int8)
    maxes = np.empty((2, L, t), np.uint16)
    kb = np.ascontiguousarray(key_bins[:L], np.float32)
    vb = np.ascontiguousarray(value_bins[:L], np.float32)
    lib().oracle_quantize(_p(x_bits), dtype, L, t, C, 2 * t * C, t * C, C, _p(kb), _p(vb), _p(sym), _p(maxes))
    return sym, maxes


def cdf(sym: np.ndarray) -> np.ndarray:
    """sym int8 [NL,t,C] -> int16 [NL,C,33].  cachegen_encoder.py:95-126,185-196."""
    sym = np.ascontiguousarray(sym, np.int8)
    NL, t, C = sym.shape
    out = np.empty((NL, C, LP), np.int16)
    lib().oracle_cdf(_p(sym), NL, t, C, _p(out))
    return out


def counts(sym: np.ndarray) -> np.ndarray:
    """sym int8 [NL,t,C] -> uint32 [NL,C,33]: the histogram calculate_cdf normalises (cachegen_encoder.py:185-196)."""
    sym = np.ascontiguousarray(sym, np.int8)
    NL, t, C = sym.shape
    out = np.empty((NL, C, LP), np.uint32)
    lib().oracle_counts(_p(sym), NL, t, C, _p(out))
    return out


def cdf_from_counts(cnt: np.ndarray, t: int) -> np.ndarray:
    """uint32 [..., 33] histogram + token count -> int16 [..., 33] CDF: the second half of oracle_cdf on its own (what
    the reader of a version-3 container evaluates)."""
    cnt = np.ascontiguousarray(cnt, np.uint32)
    out = np.empty(cnt.shape, np.int16)
    lib().oracle_cdf_from_counts(_p(cnt), cnt.size // LP, int(t), _p(out))
    return out


CODER_AC = 0     # B2KV container version 1: torchac-lineage arithmetic coder
CODER_RANS = 1   # B2KV container version 2: rANS, 32-bit state / 16-bit renormalisation, same CDF section
CODER_RANS_COMPACT = 2   # version 3: the same rANS streams, each preceded by its own histogram (mask + count bytes, the
                         # last count implied); no CDF section; stream lengths stored as bytes / 2 in a u8


def nb_map(key_bins, value_bins, L: int) -> List[int]:
    """symbols a stream of every plane (keys, then values) can emit = a version-3 container's nb map: 2 * (bins // 2)"""
    return [2 * (int(b) // 2) for b in list(key_bins)[:L]] + [2 * (int(b) // 2) for b in list(value_bins)[:L]]


def v3_pack(cnt: np.ndarray, nb: List[int], rans_lengths: np.ndarray, rans: np.ndarray):
    """version-3 payload of one <= 256-token chunk: (payload u8 [N], half_lengths u8 [NL,C]) from the histogram
    uint32 [NL,C,33], the nb map and the group's rANS streams (lengths int32 [NL,C] + concatenated bytes)."""
    cnt = np.ascontiguousarray(cnt, np.uint32)
    NL, C, _ = cnt.shape
    nb_a = np.ascontiguousarray(nb, np.int32)
    ln = np.ascontiguousarray(rans_lengths, np.int32).reshape(NL, C)
    rans = np.ascontiguousarray(rans, np.uint8)
    cap = int(rans.size + NL * C * 40)
    out = np.empty(cap, np.uint8)
    half = np.empty((NL, C), np.uint8)
    n = lib().oracle_v3_pack(_p(cnt), _p(nb_a), NL, C, _p(ln), _p(rans), _p(out), cap, _p(half))
    assert n >= 0
    return out[:n].copy(), half


def v3_unpack(payload: np.ndarray, half: np.ndarray, nb: List[int], t: int):
    """inverse of v3_pack: (counts uint32 [NL,C,33], rans_lengths int32 [NL,C], rans u8 [M])"""
    payload = np.ascontiguousarray(payload, np.uint8)
    half = np.ascontiguousarray(half, np.uint8)
    NL, C = half.shape
    nb_a = np.ascontiguousarray(nb, np.int32)
    cnt = np.empty((NL, C, LP), np.uint32)
    ln = np.empty((NL, C), np.int32)
    rans = np.empty(max(payload.size, 1), np.uint8)
    n = lib().oracle_v3_unpack(_p(payload), payload.size, _p(half), _p(nb_a), NL, C, int(t), _p(cnt), _p(ln), _p(rans))
    assert n >= 0, "malformed version-3 stream header"
    return cnt, ln, rans[:n].copy()


def encode_group(cdf_i16: np.ndarray, sym: np.ndarray, tok0: int, g: int, coder: int = CODER_AC):
    """One <=256-token group -> (bytestream u8 [N], lengths i32 [NL,C]).
    cachegen_encoder.py:225-262,301-316."""
    sym = np.ascontiguousarray(sym, np.int8)
    cdf_i16 = np.ascontiguousarray(cdf_i16, np.int16)
    NL, t, C = sym.shape
    cap = NL * C * (2 * g + 8)
    out = np.empty(cap, np.uint8)
    lengths = np.empty((NL, C), np.int32)
    fn = lib().oracle_encode_group if coder == CODER_AC else lib().oracle_encode_group_rans
    n = fn(_p(cdf_i16), _p(sym), NL, t, tok0, g, C, _p(out), cap, _p(lengths))
    assert n >= 0
    return out[:n].copy(), lengths


def decode_group(cdf_i16: np.ndarray, bytestream: np.ndarray, lengths: np.ndarray, out_sym: np.ndarray, tok0: int,
                 g: int, coder: int = CODER_AC) -> None:
    """Inverse of encode_group, writes out_sym[:, tok0:tok0+g, :] (uint8 [NL,t,C]).
    cachegen_decoder.py:52-66,94-104."""
    NL, t, C = out_sym.shape
    assert out_sym.dtype == np.uint8 and out_sym.flags.c_contiguous
    bs = np.ascontiguousarray(bytestream, np.uint8)
    ln = np.ascontiguousarray(lengths, np.int32)
    cdf_i16 = np.ascontiguousarray(cdf_i16, np.int16)
    if coder != CODER_AC:
        bad = lib().oracle_decode_group_rans(_p(cdf_i16), _p(bs), _p(ln), NL, t, tok0, g, C, _p(out_sym))
        assert bad == 0, f"{bad} rANS streams did not return to the initial state"
    else:
        lib().oracle_decode_group(_p(cdf_i16), _p(bs), _p(ln), NL, t, tok0, g, C, _p(out_sym))


def dequantize(sym_u8: np.ndarray, maxes: np.ndarray, max_dtype: int, key_bins, value_bins, out_dtype: int) -> np.ndarray:
    """sym uint8 [2L,t,C] + maxes u16 [2,L,t] -> blob bits u16 [L,2,t,C].
    cachegen_decoder.py:24-35,177-200."""
    sym_u8 = np.ascontiguousarray(sym_u8, np.uint8)
    NL, t, C = sym_u8.shape
    L = NL // 2
    maxes = np.ascontiguousarray(maxes, np.uint16)
    out = np.empty((L, 2, t, C), np.uint16)
    kb = np.ascontiguousarray(key_bins[:L], np.float32)
    vb = np.ascontiguousarray(value_bins[:L], np.float32)
    lib().oracle_dequantize(_p(sym_u8), _p(maxes), max_dtype, L, t, C, _p(kb), _p(vb), out_dtype, _p(out), 2 * t * C,
                            t * C, C)
    return out


def encode_chunk(x_bits: np.ndarray, dtype: int, key_bins, value_bins, coder: int = CODER_AC):
    """Full encode_function (cachegen_encoder.py:266-325) on one chunk [L,2,t,C]:
    returns dict(cdf, maxes, groups=[(bytestream, lengths, ntokens)], sym, counts).  coder = CODER_RANS_COMPACT codes
    the same streams as CODER_RANS; the dict's `counts` is what such a container's stream headers carry in place of
    `cdf` (v3_pack builds its payload)."""
    sym, maxes = quantize(x_bits, dtype, key_bins, value_bins)
    c = cdf(sym)
    t = sym.shape[1]
    groups = []
    for tok0 in range(0, t, GROUP):
        g = min(GROUP, t - tok0)
        bs, ln = encode_group(c, sym, tok0, g, coder)
        groups.append((bs, ln, g))
    out = dict(cdf=c, maxes=maxes, groups=groups, sym=sym, coder=coder)
    if coder == CODER_RANS_COMPACT:          # only a version-3 container stores the histogram (keeps the timed CPU arm lean)
        out["counts"] = counts(sym)
    return out


def decode_chunk(enc: dict, max_dtype: int, key_bins, value_bins, out_dtype: int) -> np.ndarray:
    """decode_function_gpu + do_dequantize + assembly (cachegen_decoder.py:70-106,143-202)."""
    c = enc["cdf"]
    NL, C, _ = c.shape
    t = sum(g for _, _, g in enc["groups"])
    sym = np.zeros((NL, t, C), np.uint8)
    tok0 = 0
    for bs, ln, g in enc["groups"]:
        decode_group(c, bs, ln, sym, tok0, g, enc.get("coder", CODER_AC))
        tok0 += g
    return dequantize(sym, enc["maxes"], max_dtype, key_bins, value_bins, out_dtype)


# ----------------------------------------------------------------------------- hash (a1)
def sha256_chain(tokens: np.ndarray, chunk_size: int) -> List[str]:
    """_prefix_hash(_chunk_tokens(tokens)) (cache_engine.py:58-96) -> hex digests."""
    tokens = np.ascontiguousarray(tokens)
    n = tokens.shape[0]
    nchunks = (n + chunk_size - 1) // chunk_size
    out = np.empty((max(nchunks, 1), 32), np.uint8)
    got = lib().oracle_sha256_chain(_p(tokens), n, tokens.dtype.itemsize, chunk_size, _p(out))
    assert got == nchunks
    return [bytes(out[i]).hex() for i in range(nchunks)]


# ----------------------------------------------------------------------------- helpers
def f32_to_bf16_bits(a: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even float32 -> bfloat16 bit pattern (uint16)."""
    u = np.ascontiguousarray(a, np.float32).view(np.uint32)
    nan = (u & 0x7FFFFFFF) > 0x7F800000
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    r[nan] = ((u[nan] >> 16) | 0x40).astype(np.uint16)
    return r


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(b, np.uint16).astype(np.uint32) << 16).view(np.float32)


def synth_kv_bits(L: int, t: int, C: int, seed: int) -> np.ndarray:
    """SURVEY.md 8d synthetic KV: x = N(0,1) * sigma[l,kv,c], sigma ~ LogNormal(0,0.5) clipped to
    [0.1, 8], 1% outlier channels x10; cast to bf16.  Returns uint16 bits [L,2,t,C]."""
    rng = np.random.default_rng(seed)
    sigma = np.clip(rng.lognormal(0.0, 0.5, size=(L, 2, 1, C)), 0.1, 8.0).astype(np.float32)
    outl = rng.random(size=(L, 2, 1, C)) < 0.01
    sigma = np.where(outl, sigma * 10.0, sigma).astype(np.float32)
    x = rng.standard_normal(size=(L, 2, t, C), dtype=np.float32) * sigma
    return f32_to_bf16_bits(x)
