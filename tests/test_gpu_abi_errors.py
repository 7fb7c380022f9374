"""GPU: the C ABI rejects bad arguments with a negative return code and a message (never a crash), and encoder-side
capacity problems surface as a header status instead of memory corruption."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _desc(t, L, H, D, dtype=0):
    from lmcache_b200 import _native as N
    d = N.KvDesc()
    d.base = t.data_ptr()
    d.planes = None
    d.sL, d.sKV, d.sT, d.sH = t.stride(0), t.stride(1), t.stride(2), t.stride(3)
    d.L, d.H, d.D, d.dtype = L, H, D, dtype
    return d


def test_encode_argument_validation():
    from lmcache_b200 import _native as N
    lib = N.lib()
    L, H, D, t = 2, 1, 128, 16
    kv = torch.randn(L, 2, t, H, D, device="cuda").to(torch.bfloat16)
    d = _desc(kv, L, H, D)
    bins = N.float_array([32.0] * L)
    lo = N.container_layout(L, H, D, t)
    out = torch.empty(lo.max_total_bytes, dtype=torch.uint8, device="cuda")
    ws = torch.empty(max(lib.b200kv_encode_workspace_bytes(L, H, D, t, 1, c) for c in range(3)), dtype=torch.uint8, device="cuda")
    sizes = torch.zeros(1, dtype=torch.int64, device="cuda")
    sp = torch.cuda.current_stream().cuda_stream

    def call(**kw):
        a = dict(kv=ctypes.byref(d), tok=0, n=1, ct=t, last=t, kb=bins, vb=bins, coder=N.CODER_RANS, out=out.data_ptr(),
                 stride=lo.max_total_bytes,
                 sizes=sizes.data_ptr(), ws=ws.data_ptr(), wsb=ws.numel())
        a.update(kw)
        return lib.b200kv_encode_chunks(a["kv"], a["tok"], a["n"], a["ct"], a["last"], a["kb"], a["vb"], a["coder"], a["out"],
                                        a["stride"], a["sizes"], a["ws"], a["wsb"], sp)

    for coder in (N.CODER_AC, N.CODER_RANS, N.CODER_RANS_COMPACT):
        assert call(coder=coder) == 0
        torch.cuda.synchronize()
        assert int(sizes[0]) > N.container_layout(L, H, D, t, coder).fixed_bytes
        assert N.Header.from_buffer_copy(out[:64].cpu().numpy().tobytes()).version == coder + 1
    for bad in (dict(n=0), dict(ct=0), dict(last=t + 1), dict(out=None), dict(out=out.data_ptr() + 1), dict(wsb=16),
                dict(kb=None), dict(tok=-1), dict(stride=64), dict(coder=3), dict(coder=-1),
                dict(coder=N.CODER_RANS_COMPACT, ct=257, last=257)):      # the compact container holds <= 256 tokens
        rc = call(**bad)
        assert rc < 0 and len(N.last_error()) > 0, bad
    bad_desc = _desc(kv, L, H, D, dtype=7)
    assert call(kv=ctypes.byref(bad_desc)) < 0
    bad_bins = N.float_array([2.0] * L)          # bins // 2 - 1 = 0: not a valid quantiser
    assert call(kb=bad_bins) < 0


@pytest.mark.parametrize("coder", [0, 1, 2])
def test_slot_too_small_sets_status_not_corruption(coder):
    """A payload that does not fit its slot must not be written past it; the header carries a nonzero status."""
    from lmcache_b200 import _native as N
    from lmcache_b200.codec import parse_header
    lib = N.lib()
    L, H, D, t = 2, 1, 128, 256
    kv = torch.rand(L, 2, t, H, D, device="cuda").to(torch.bfloat16)       # ~5 bits/symbol: a real payload
    d = _desc(kv, L, H, D)
    bins = N.float_array([32.0] * L)
    lo = N.container_layout(L, H, D, t, coder)
    stride = lo.fixed_bytes + 256                                           # far too small for the payload
    guard = 4096
    out = torch.full((stride + guard,), 0xAB, dtype=torch.uint8, device="cuda")
    ws = torch.empty(lib.b200kv_encode_workspace_bytes(L, H, D, t, 1, coder), dtype=torch.uint8, device="cuda")
    sizes = torch.zeros(1, dtype=torch.int64, device="cuda")
    rc = lib.b200kv_encode_chunks(ctypes.byref(d), 0, 1, t, t, bins, bins, coder, out.data_ptr(), stride, sizes.data_ptr(),
                                  ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    assert bool((out[stride:] == 0xAB).all()), "wrote past the slot"
    hd = N.Header.from_buffer_copy(out[:64].cpu().numpy().tobytes())
    assert hd.status != 0
    with pytest.raises(ValueError):
        parse_header(out[:stride].cpu().numpy().tobytes())


def test_decode_and_misc_validation():
    from lmcache_b200 import _native as N
    lib = N.lib()
    L, H, D, t = 2, 1, 128, 16
    dst = torch.empty(L, 2, t, H, D, dtype=torch.bfloat16, device="cuda")
    d = _desc(dst, L, H, D)
    bins = N.float_array([32.0] * L)
    buf = torch.zeros(1 << 16, dtype=torch.uint8, device="cuda")
    ws = torch.empty(max(1, lib.b200kv_decode_workspace_bytes(L, H, D, t, 1)), dtype=torch.uint8, device="cuda")
    sp = torch.cuda.current_stream().cuda_stream
    total = N.container_layout(L, H, D, t).fixed_bytes + 4 * 2 * L * H * D
    ok = (buf.data_ptr(), buf.numel(), N.i64_array([0]), N.i64_array([total]), N.i32_array([t]), N.i64_array([0]), 1, 0,
          N.CODER_RANS, ctypes.byref(d), bins, bins, None, ws.data_ptr(), ws.numel(), sp)
    # misaligned offset, zero tokens, zero chunks, bad max_dtype, bad coder, tiny workspace, NULL buffer, a container
    # shorter than its fixed sections, a buffer without the read slack
    for i, bad in [(2, N.i64_array([8])), (4, N.i32_array([0])), (6, 0), (7, 9), (8, 5), (14, 8), (0, None),
                   (3, N.i64_array([100])), (1, total + 16)]:
        a = list(ok)
        a[i] = bad
        assert lib.b200kv_decode_chunks(*a) < 0 and N.last_error(), i
    # a compact container cannot hold more than one 256-token group
    a = list(ok)
    a[4], a[8] = N.i32_array([257]), N.CODER_RANS_COMPACT
    assert lib.b200kv_decode_chunks(*a) < 0 and N.last_error()
    with pytest.raises(N.NativeError):
        N.container_layout(L, H, D, 257, N.CODER_RANS_COMPACT)
    assert lib.b200kv_sha256_chain(buf.data_ptr(), 3, N.i64_array([0, 4]), 1, 4, buf.data_ptr(), sp) < 0     # elem_size
    assert lib.b200kv_sha256_chain(buf.data_ptr(), 8, N.i64_array([4, 0]), 1, 4, buf.data_ptr(), sp) < 0     # decreasing offsets
    assert lib.b200kv_sha256_chain(buf.data_ptr(), 8, N.i64_array([0, 0]), 1, 4, buf.data_ptr(), sp) == 0    # empty: no-op
    assert lib.b200kv_pinned_alloc(None, 16) < 0
    assert lib.b200kv_copy_async(None, buf.data_ptr(), 16, sp) < 0
    assert N.container_layout(1, 1, 1, 1).off_cdf == 64
    assert lib.b200kv_encode_workspace_bytes(0, 1, 1, 1, 1, 0) < 0
    assert lib.b200kv_encode_workspace_bytes(1, 1, 1, 1, 1, 7) < 0
