"""ctypes binding of libb200kv.so (include/b200kv.h) -- the only door to the CUDA hot path.

There is deliberately no fallback: if the shared library is missing, or no CUDA device is
visible, every compute entry point raises.  `lib()` only needs the .so; `require_cuda()`
additionally needs a device.
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200kv.so")

DT_BF16 = 0
DT_FP16 = 1
CODER_AC = 0       # container version 1: arithmetic coder
CODER_RANS = 1     # container version 2: rANS
CODER_RANS_COMPACT = 2   # container version 3: rANS streams that carry their own histogram (no CDF section), one-byte lengths
ENCODE_HINT_MID_ENTROPY = 0x200    # B200KV_ENCODE_HINT_MID_ENTROPY
HDR_MAX = 36             # longest version-3 stream header (4 mask bytes + 31 counts + 1 pad)
CODERS = {"ac": CODER_AC, "rans": CODER_RANS, "rans_compact": CODER_RANS_COMPACT}
ENCODE_HINT_HIGH_ENTROPY = 0x100   # B200KV_ENCODE_HINT_HIGH_ENTROPY
LP = 33
GROUP_TOKENS = 256
MAX_PLANES = 128
MAGIC = 0x564B3242
HEADER_BYTES = 64
READ_SLACK = 640     # B200KV_READ_SLACK

c_i32, c_i64, c_vp, c_u64 = ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p, ctypes.c_uint64


class NativeError(RuntimeError):
    """Nonzero return code from libb200kv (message from b200kv_last_error())."""


class KvDesc(ctypes.Structure):
    """struct b200kv_kv_desc"""
    _fields_ = [
        ("base", c_vp),
        ("planes", ctypes.POINTER(c_vp)),
        ("sL", c_i64), ("sKV", c_i64), ("sT", c_i64), ("sH", c_i64),
        ("L", c_i32), ("H", c_i32), ("D", c_i32),
        ("dtype", c_i32),
        ("slot_map", c_vp),     # device int64[ntokens] or NULL (paged KV)
    ]


class Header(ctypes.Structure):
    """struct b200kv_header (64 bytes)"""
    _fields_ = [
        ("magic", ctypes.c_uint32), ("version", ctypes.c_uint32),
        ("L", ctypes.c_uint32), ("H", ctypes.c_uint32), ("D", ctypes.c_uint32),
        ("ntokens", ctypes.c_uint32), ("ngroups", ctypes.c_uint32),
        ("max_dtype", ctypes.c_uint32),
        ("payload_bytes", c_u64), ("total_bytes", c_u64),
        ("status", ctypes.c_uint32), ("reserved", ctypes.c_uint32 * 3),
    ]


class Layout(ctypes.Structure):
    """struct b200kv_layout"""
    _fields_ = [("off_cdf", c_i64), ("off_maxes", c_i64), ("off_lengths", c_i64), ("off_payload", c_i64),
                ("fixed_bytes", c_i64), ("max_total_bytes", c_i64)]


assert ctypes.sizeof(Header) == HEADER_BYTES

# name -> (restype, argtypes); every symbol include/b200kv.h declares
SIGNATURES = {
    "b200kv_version": (c_i32, []),
    "b200kv_last_error": (ctypes.c_char_p, []),
    "b200kv_device_count": (c_i32, []),
    "b200kv_container_layout": (c_i32, [c_i32, c_i32, c_i32, c_i32, ctypes.POINTER(Layout)]),
    "b200kv_container_layout_v": (c_i32, [c_i32, c_i32, c_i32, c_i32, c_i32, ctypes.POINTER(Layout)]),
    "b200kv_encode_workspace_bytes": (c_i64, [c_i32, c_i32, c_i32, c_i32, c_i32, c_i32]),
    "b200kv_decode_workspace_bytes": (c_i64, [c_i32, c_i32, c_i32, c_i32, c_i32]),
    "b200kv_encode_chunks": (c_i32, [ctypes.POINTER(KvDesc), c_i64, c_i32, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp, c_i64,
                                      c_vp, c_vp, c_i64, c_vp]),
    "b200kv_decode_chunks": (c_i32, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, ctypes.POINTER(KvDesc),
                                      c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "b200kv_sha256_chain": (c_i32, [c_vp, c_i32, c_vp, c_i32, c_i32, c_vp, c_vp]),
    "b200kv_sha256_chain_ready": (c_i32, [c_vp, c_i32, c_vp, c_i32, c_i32, c_vp, c_vp, ctypes.c_uint32, c_vp]),
    "b200kv_pack_chunks": (c_i32, [ctypes.POINTER(KvDesc), c_i64, c_i32, c_i32, c_i32, c_i32, c_vp, c_i64, c_vp]),
    "b200kv_unpack_chunks": (c_i32, [c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, ctypes.POINTER(KvDesc), c_i64, c_vp]),
    "b200kv_pinned_alloc": (c_i32, [ctypes.POINTER(c_vp), c_i64]),
    "b200kv_pinned_free": (c_i32, [c_vp]),
    "b200kv_host_device_ptr": (c_i32, [c_vp, ctypes.POINTER(c_vp)]),
    "b200kv_copy_async": (c_i32, [c_vp, c_vp, c_i64, c_vp]),
    "b200kv_copy2d_async": (c_i32, [c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp]),
    "b200kv_stream_create": (c_i32, [ctypes.POINTER(c_vp)]),
    "b200kv_stream_destroy": (c_i32, [c_vp]),
    "b200kv_stream_sync": (c_i32, [c_vp]),
    "b200kv_event_create": (c_i32, [ctypes.POINTER(c_vp)]),
    "b200kv_event_destroy": (c_i32, [c_vp]),
    "b200kv_event_record": (c_i32, [c_vp, c_vp]),
    "b200kv_event_query": (c_i32, [c_vp]),
    "b200kv_event_sync": (c_i32, [c_vp]),
    "b200kv_stream_wait_event": (c_i32, [c_vp, c_vp]),
    "b200kv_event_elapsed_ms": (c_i32, [c_vp, c_vp, ctypes.POINTER(ctypes.c_float)]),
    "b200kv_profile_enable": (c_i32, [c_i32]),
    "b200kv_profile_last": (c_i32, [ctypes.POINTER(ctypes.c_float), c_i32]),
    "b200kv_lm_server_start": (c_i32, [ctypes.c_char_p, c_i32, ctypes.POINTER(c_vp)]),
    "b200kv_lm_server_port": (c_i32, [c_vp]),
    "b200kv_lm_server_num_keys": (c_i64, [c_vp]),
    "b200kv_lm_server_stop": (c_i32, [c_vp]),
    "b200kv_lm_connect": (c_i32, [ctypes.c_char_p, c_i32, ctypes.POINTER(c_vp)]),
    "b200kv_lm_close": (c_i32, [c_vp]),
    "b200kv_lm_put": (c_i32, [c_vp, ctypes.c_char_p, c_vp, c_i64]),
    "b200kv_lm_exists": (c_i32, [c_vp, ctypes.c_char_p]),
    "b200kv_lm_get_begin": (c_i64, [c_vp, ctypes.c_char_p]),
    "b200kv_lm_list_begin": (c_i64, [c_vp]),
    "b200kv_lm_read": (c_i32, [c_vp, c_vp, c_i64]),
}
PROFILE_SLOTS = ("absmax", "cdf", "encode", "compact", "tile_sum", "tile_scan", "decode")

_lib: Optional[ctypes.CDLL] = None
_lock = threading.Lock()


def lib() -> ctypes.CDLL:
    """Load libb200kv.so (built in-tree by __graft_entry__.build()).  Raises if absent."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: the CUDA extension is not built "
                        f"(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
                L = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(L, name)      # AttributeError if the .so lacks a declared symbol
                    fn.restype = res
                    fn.argtypes = args
                _lib = L
    return _lib


def last_error() -> str:
    msg = lib().b200kv_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str = "") -> int:
    if rc < 0:
        raise NativeError(f"libb200kv {what} failed (rc={rc}): {last_error()}")
    return rc


_cuda_ok: Optional[bool] = None


def require_cuda() -> None:
    """Fail loudly when the hot path cannot run (no device / no driver)."""
    global _cuda_ok
    if _cuda_ok is None:
        n = lib().b200kv_device_count()
        _cuda_ok = n > 0
        if not _cuda_ok:
            _cuda_ok = None
            raise RuntimeError(f"lmcache_b200 needs a CUDA device (sm_100a); none usable: {last_error() or 'count=0'}. "
                               f"There is no CPU fallback.")


def container_layout(L: int, H: int, D: int, ntokens: int, coder: int = CODER_RANS) -> Layout:
    """Section offsets of the container `coder` produces (versions 1 and 2 share a layout)."""
    lo = Layout()
    check(lib().b200kv_container_layout_v(L, H, D, ntokens, coder, ctypes.byref(lo)), "container_layout")
    return lo


def nb_map(key_bins, value_bins, L: int) -> list:
    """symbols a stream of every plane (keys then values) can emit = the nb map of a compact container: 2 * (bins // 2)"""
    return [2 * (int(b) // 2) for b in list(key_bins)[:L]] + [2 * (int(b) // 2) for b in list(value_bins)[:L]]


def float_array(vals: Sequence[float]):
    return (ctypes.c_float * len(vals))(*[float(v) for v in vals])


def i64_array(vals: Sequence[int]):
    return (c_i64 * len(vals))(*[int(v) for v in vals])


def i32_array(vals: Sequence[int]):
    return (c_i32 * len(vals))(*[int(v) for v in vals])
