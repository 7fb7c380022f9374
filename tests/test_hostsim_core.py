"""CPU: the product's device arithmetic (lmcache_b200/csrc/ac_core.cuh, compiled for the host by
tests/hostsim) against the oracle -- same functions the CUDA kernels call."""
import ctypes
import os

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def sim():
    S = ctypes.CDLL(os.path.join(HERE, "hostsim", "libhostsim.so"))
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    S.sim_encode_stream.restype = i64
    S.sim_encode_stream.argtypes = [vp, vp, i64, i32, vp, i64]
    S.sim_decode_stream.argtypes = [vp, vp, i64, i32, vp, i64]
    S.sim_encode_stream2.restype = i64
    S.sim_encode_stream2.argtypes = [vp, vp, i64, i32, vp, i64]
    S.sim_decode_stream2.argtypes = [vp, vp, i64, i32, vp, i64, i32, i32]
    S.sim_cdf.argtypes = [vp, i32, vp]
    S.sim_cdf_skip.argtypes = [vp, i32, ctypes.c_uint32, vp]
    S.sim_quant_row.argtypes = [vp, i32, i32, ctypes.c_uint16, ctypes.c_float, vp]
    S.sim_dequant_row.argtypes = [vp, i32, ctypes.c_uint16, i32, ctypes.c_float, i32, vp]
    S.sim_half_to_float.argtypes = [vp, i32, i32, vp]
    S.sim_float_to_half.argtypes = [vp, i32, i32, vp]
    S.sim_layout.argtypes = [i32, i32, i32, i32, vp]
    S.sim_hdr_write.argtypes = [vp, i32, vp]
    S.sim_hdr_len.argtypes = [ctypes.c_uint32, i32]
    return S


def P(a):
    return ctypes.c_void_p(a.ctypes.data)


def _symbols(rng, kind, shape):
    if kind == "peaked":
        return np.clip(np.rint(rng.normal(15, 1.5, size=shape)), 0, 30).astype(np.int8)
    if kind == "uniform":
        return rng.integers(0, 31, size=shape).astype(np.int8)
    if kind == "rare":
        return np.where(rng.random(shape) < 0.02, rng.integers(0, 31, shape), 7).astype(np.int8)
    if kind == "mid":   # two symbols straddling the midpoint: long E3 (pending) runs
        return np.where(rng.random(shape) < 0.5, 14, 15).astype(np.int8)
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["peaked", "uniform", "rare", "mid"])
@pytest.mark.parametrize("t", [1, 2, 3, 7, 100, 236, 256])
def test_coder_bit_exact_vs_oracle(sim, kind, t):
    rng = np.random.default_rng(hash((kind, t)) & 0xffff)
    NL, C = 2, 6
    sym = _symbols(rng, kind, (NL, t, C))
    cdf = O.cdf(sym)
    bs, ln = O.encode_group(cdf, sym, 0, t)
    off = 0
    for nl in range(NL):
        for c in range(C):
            col = np.ascontiguousarray(sym[nl, :, c])
            cd = np.ascontiguousarray(cdf[nl, c]).view(np.uint16)
            out = np.zeros(2 * t + 16, np.uint8)
            n = sim.sim_encode_stream(P(cd), P(col), 1, t, P(out), out.size)
            ref = np.ascontiguousarray(bs[off:off + ln[nl, c]])
            assert n == ln[nl, c] and np.array_equal(out[:n], ref)
            dec = np.zeros(t, np.uint8)
            sim.sim_decode_stream(P(cd), P(ref), ref.size, t, P(dec), 1)
            assert np.array_equal(dec, col.view(np.uint8))
            off += ln[nl, c]


def _check_v2(sim, cdf, sym, tok0, g):
    """production (branch-light) encoder / decoder of ac_core.cuh vs the bit-by-bit oracle, incl. a stream that
    starts mid-word (foreign bytes before it) and garbage after its end"""
    NL, _, C = sym.shape
    bs, ln = O.encode_group(cdf, sym, tok0, g)
    off = 0
    for nl in range(NL):
        for c in range(C):
            col = np.ascontiguousarray(sym[nl, tok0:tok0 + g, c])
            cd = np.ascontiguousarray(cdf[nl, c]).view(np.uint16)
            out = np.zeros(2 * g + 64, np.uint8)
            n = sim.sim_encode_stream2(P(cd), P(col), 1, g, P(out), out.size)
            ref = np.ascontiguousarray(bs[off:off + ln[nl, c]])
            assert n == ln[nl, c] and np.array_equal(out[:n], ref)
            for skip in (0, 1, 2, 3):
                buf = np.concatenate([np.full(skip, 0x5A, np.uint8), ref])
                dec = np.zeros(g, np.uint8)
                sim.sim_decode_stream2(P(cd), ctypes.c_void_p(buf.ctypes.data + skip), ref.size, g, P(dec), 1, skip, 5)
                assert np.array_equal(dec, col.view(np.uint8)), (nl, c, skip)
                if col.max() <= 15:      # 16-bin planes: the 4-step search covers every symbol they can hold
                    sim.sim_decode_stream2(P(cd), ctypes.c_void_p(buf.ctypes.data + skip), ref.size, g, P(dec), 1, skip, 4)
                    assert np.array_equal(dec, col.view(np.uint8)), (nl, c, skip, "4-step")
            off += ln[nl, c]


@pytest.mark.parametrize("kind", ["peaked", "uniform", "rare", "mid"])
@pytest.mark.parametrize("t", [1, 2, 3, 5, 16, 100, 236, 256])
def test_production_coder_bit_exact_vs_oracle(sim, kind, t):
    rng = np.random.default_rng(hash((kind, t, 2)) & 0xffff)
    sym = _symbols(rng, kind, (2, t, 5))
    _check_v2(sim, O.cdf(sym), sym, 0, t)


def test_production_coder_foreign_cdf(sim):
    rng = np.random.default_rng(17)
    base = np.full((1, 8192, 6), 15, np.int8)
    base[:, :256, :] = rng.integers(0, 31, size=(1, 256, 6))                      # ~10 bits / symbol
    base[:, 256:512, :] = np.where(rng.random((1, 256, 6)) < 0.5, 14, 15)         # straddles the midpoint
    base[:, 512:768, :] = np.where(rng.random((1, 256, 6)) < 0.03, 3, 15)
    cdf = O.cdf(base)
    for tok0 in (0, 256, 512, 4096):
        _check_v2(sim, cdf, base, tok0, 256)


def test_production_coder_long_pending_run(sim):
    """hand-made CDF whose two symbols meet exactly at the midpoint: E3 runs far longer than one 32-bit word"""
    cdf = np.zeros((1, 1, 33), np.uint16)
    cdf[0, 0, 1:] = 32768 + np.arange(32)            # symbol 0: [0, 32768), symbol 1: [32768, 32769), ...
    cdf[0, 0, 32] = 0
    sym = np.zeros((1, 256, 1), np.int8)
    sym[0, ::2, 0] = 1                               # 1,0,1,0,... hugs the midpoint from above / below
    sym[0, 200:, 0] = 0
    _check_v2(sim, cdf.view(np.int16), sym, 0, 256)
    sym2 = np.zeros((1, 256, 1), np.int8)
    sym2[0, 0, 0] = 1
    _check_v2(sim, cdf.view(np.int16), sym2, 0, 256)
    # middle symbol [0.25, 0.75): every occurrence is one E3 step -> pending grows by one per symbol
    cdf3 = np.zeros((1, 1, 33), np.uint16)
    cdf3[0, 0, 1], cdf3[0, 0, 2] = 16384, 49152
    cdf3[0, 0, 3:32] = 65000 + np.arange(29)
    for pattern in ([1] * 60 + [0] + [1] * 100 + [2] + [1] * 33 + [0] + [1] * 60, [1] * 256, [1] * 255 + [2]):
        sym3 = np.array(pattern, np.int8).reshape(1, 256, 1)
        _check_v2(sim, cdf3.view(np.int16), sym3, 0, 256)


def test_production_coder_random_carry_stress(sim):
    """random few-symbol CDFs whose boundaries sit on / next to the midpoint and quarter points: the pending run
    (= carry ripple in the production coder) crosses flushed 32-bit words in both directions"""
    rng = np.random.default_rng(2024)
    anchors = np.array([16384, 32768, 49152, 8192, 24576, 40960, 57344])
    for case in range(120):
        nsym = int(rng.integers(2, 6))
        cuts = np.unique(np.clip(rng.choice(anchors, nsym - 1) + rng.integers(-2, 3, nsym - 1), 1, 65000))
        nsym = len(cuts) + 1
        cdf = np.zeros((1, 1, 33), np.uint16)
        cdf[0, 0, 1:nsym] = cuts
        cdf[0, 0, nsym:32] = 65100 + np.arange(32 - nsym)          # unused symbols keep width 1
        g = int(rng.integers(1, 257))
        if case % 3 == 0:       # long runs of one symbol, occasionally broken
            sym = np.full(g, rng.integers(0, nsym), np.int64)
            brk = rng.random(g) < 0.03
            sym[brk] = rng.integers(0, nsym, brk.sum())
        else:
            sym = rng.integers(0, nsym, g)
        _check_v2(sim, cdf.view(np.int16), sym.astype(np.int8).reshape(1, g, 1), 0, g)


def test_coder_foreign_cdf_expensive_symbols(sim):
    """Group coded with a CDF that is NOT its own histogram (chunk > 256 tokens): symbols may cost up to 16
    bits and pending runs get long; encoder and decoder must still match the bit-by-bit oracle."""
    rng = np.random.default_rng(99)
    t_total, g, NL, C = 8192, 256, 1, 8
    base = np.full((NL, t_total, C), 15, np.int8)
    base[:, :g, :] = rng.integers(0, 31, size=(NL, g, C))      # first group uses symbols that are rare chunk-wide
    cdf = O.cdf(base)
    bs, ln = O.encode_group(cdf, base, 0, g)
    assert ln.max() > g                                         # > 8 bits / symbol really happens
    off = 0
    for c in range(C):
        col = np.ascontiguousarray(base[0, :g, c])
        cd = np.ascontiguousarray(cdf[0, c]).view(np.uint16)
        out = np.zeros(2 * g + 16, np.uint8)
        n = sim.sim_encode_stream(P(cd), P(col), 1, g, P(out), out.size)
        ref = np.ascontiguousarray(bs[off:off + ln[0, c]])
        assert n == ln[0, c] and np.array_equal(out[:n], ref)
        dec = np.zeros(g, np.uint8)
        sim.sim_decode_stream(P(cd), P(ref), ref.size, g, P(dec), 1)
        assert np.array_equal(dec, col.view(np.uint8))
        off += ln[0, c]


def test_cdf_bit_exact_vs_golden(sim, golden, golden_names):
    for n in golden_names:
        sym, cdf = golden[f"{n}/sym"], golden[f"{n}/cdf"]
        NL, t, C = sym.shape
        for nl in range(0, NL, 5):
            for c in range(0, C, 7):
                counts = np.bincount(sym[nl, :, c], minlength=33)[:33].astype(np.uint32)
                out = np.zeros(33, np.uint16)
                sim.sim_cdf(P(counts), t, P(out))
                assert np.array_equal(out.view(np.int16), cdf[nl, c]), (n, nl, c)
                # the value / absorb form the kernels use, skipping every symbol with a zero count (the kernels skip the
                # ones no lane of a warp uses) and skipping none: both equal the reference-made CDF
                for skip in (int(sum(1 << i for i in range(32) if counts[i] == 0)), 0):
                    out2 = np.zeros(33, np.uint16)
                    sim.sim_cdf_skip(P(counts), t, skip, P(out2))
                    assert np.array_equal(out2.view(np.int16), cdf[nl, c]), (n, nl, c, skip)


def test_quant_dequant_bit_exact_vs_golden(sim, golden, golden_names):
    kb, vb = golden["key_bins"], golden["value_bins"]
    for n in golden_names:
        x = golden[f"{n}/x"]
        dt = int(golden[f"{n}/dtype"][0])
        L, _, t, H, D = x.shape
        C = H * D
        x = x.reshape(L, 2, t, C)
        sym = golden[f"{n}/sym"]
        mk, mv = golden[f"{n}/max_k"].reshape(L, t), golden[f"{n}/max_v"].reshape(L, t)
        deq = golden[f"{n}/deq_vllm_bf16"].reshape(L, 2, t, C)
        for l in range(L):
            for kv in range(2):
                bins = (vb if kv else kb)[l]
                maxq = float(bins // 2 - 1)
                for tok in range(0, t, max(1, t // 5)):
                    mb = int((mv if kv else mk)[l, tok])
                    row = np.ascontiguousarray(x[l, kv, tok])
                    out = np.zeros(C, np.uint8)
                    sim.sim_quant_row(P(row), dt, C, mb, maxq, P(out))
                    assert np.array_equal(out, sym[kv * L + l, tok].view(np.uint8)), (n, l, kv, tok)
                    dq = np.zeros(C, np.uint16)
                    sim.sim_dequant_row(P(out), C, mb, dt, maxq, 0, P(dq))
                    assert np.array_equal(dq, deq[l, kv, tok]), (n, l, kv, tok)


def test_half_conversions_exhaustive(sim):
    allh = np.arange(65536, dtype=np.uint16)
    f = np.zeros(65536, np.float32)
    sim.sim_half_to_float(P(allh), 65536, 1, P(f))
    ref = allh.view(np.float16).astype(np.float32)
    assert np.array_equal(f.view(np.uint32), ref.view(np.uint32))
    sim.sim_half_to_float(P(allh), 65536, 0, P(f))
    assert np.array_equal(f.view(np.uint32), allh.astype(np.uint32) << 16)
    # float -> half, RNE, over a dense sample incl. ties, subnormals, overflow
    rng = np.random.default_rng(3)
    vals = np.concatenate([
        rng.standard_normal(200000).astype(np.float32) * np.float32(10.0) ** rng.integers(-9, 6, 200000).astype(np.float32),
        ref[np.isfinite(ref)], np.array([65504.0, 65519.9, 65520.0, 1e9, -1e9, 2.0 ** -24, 2.0 ** -25, 1.5 * 2.0 ** -25,
                                          np.inf, -np.inf, 0.0, -0.0], np.float32)])
    mids = (ref[:-1].astype(np.float64) + ref[1:].astype(np.float64)) / 2   # exact ties between neighbours
    vals = np.concatenate([vals, mids[np.isfinite(mids)].astype(np.float32)]).astype(np.float32)
    out = np.zeros(vals.size, np.uint16)
    sim.sim_float_to_half(P(vals), vals.size, 1, P(out))
    with np.errstate(over="ignore"):
        assert np.array_equal(out, vals.astype(np.float16).view(np.uint16))
    sim.sim_float_to_half(P(vals), vals.size, 0, P(out))
    assert np.array_equal(out, O.f32_to_bf16_bits(vals))


def test_layout_matches_c_abi(sim):
    from lmcache_b200 import _native as N
    for (L, H, D, t) in [(32, 32, 128, 256), (12, 1, 16, 300), (40, 2, 128, 17), (1, 1, 1, 1)]:
        for coder in (N.CODER_RANS, N.CODER_RANS_COMPACT):
            if coder == N.CODER_RANS_COMPACT and t > 256:
                continue
            lo = N.container_layout(L, H, D, t, coder)
            out = np.zeros(5, np.int64)
            sim.sim_layout(L, H * D, t, 1 if coder == N.CODER_RANS_COMPACT else 0, P(out))
            assert (lo.off_cdf, lo.off_maxes, lo.off_lengths, lo.off_payload) == tuple(out[:4])
            assert lo.off_payload % 16 == 0 and lo.max_total_bytes >= lo.off_payload


def test_stream_header_spec_matches_oracle(sim):
    """version-3 stream header: the shared spec functions of ac_core.cuh (hdr_write_host, hdr_len -- the latter is what
    the compaction and decode kernels evaluate) against the oracle's packer, on random histograms incl. a lone symbol
    with 256 tokens, every nb, odd and even header lengths."""
    from oracle import oracle as O
    rng = np.random.default_rng(5)
    for nb in (4, 6, 8, 10, 16, 30, 32):
        for trial in range(40):
            t = int(rng.integers(1, 257))
            k = int(rng.integers(1, min(nb - 1, t) + 1))                   # symbols that occur (<= nb - 1 by construction)
            syms = rng.choice(nb - 1, size=k, replace=False)
            cnt = np.zeros(33, np.uint32)
            cnt[syms] = 1
            for _ in range(t - k):
                cnt[rng.choice(syms)] += 1
            if trial == 0:
                cnt[:] = 0
                cnt[nb - 2] = 256                                           # a lone symbol: count implied
            out = np.zeros(48, np.uint8)
            n = sim.sim_hdr_write(P(cnt), nb, P(out))
            mask = int(sum(1 << i for i in range(nb) if cnt[i]))
            assert n == sim.sim_hdr_len(mask, nb) and n % 2 == 0 and n <= 36
            pl, half = O.v3_pack(cnt.reshape(1, 1, 33), [nb], np.array([[4]], np.int32), np.zeros(4, np.uint8))
            assert bytes(out[:n]) == pl[:n].tobytes() and int(half[0, 0]) * 2 == n + 4
            back, ln, _ = O.v3_unpack(pl, half, [nb], int(cnt.sum()))
            assert np.array_equal(back[0, 0], cnt)
