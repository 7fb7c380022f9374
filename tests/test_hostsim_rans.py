"""CPU: the rANS coder of B2KV container version 2 -- the product's arithmetic (lmcache_b200/csrc/ac_core.cuh compiled
for the host by tests/hostsim) against the oracle's independent restatement (oracle/cachegen_oracle.c), plus the
properties that pin a coder whose bitstream the reference does not define: encode -> decode identity on
symbols, code length within a stated bound of the ideal, and the free end-of-stream integrity check."""
import ctypes
import math
import os

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def sim():
    S = ctypes.CDLL(os.path.join(HERE, "hostsim", "libhostsim.so"))
    vp, i64, i32, u32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_uint32
    S.sim_rans_encode_stream.restype = i64
    S.sim_rans_encode_stream.argtypes = [vp, vp, i64, i32, vp, i64]
    S.sim_rans_decode_stream.restype = u32
    S.sim_rans_decode_stream.argtypes = [vp, vp, i64, i32, vp, i64, i32, i32]
    S.sim_rans_divmod.restype = u32
    S.sim_rans_divmod.argtypes = [u32, u32, vp]
    return S


def P(a):
    return ctypes.c_void_p(a.ctypes.data)


def _symbols(rng, kind, shape):
    if kind == "peaked":
        return np.clip(np.rint(rng.normal(15, 1.5, size=shape)), 0, 30).astype(np.int8)
    if kind == "sharp":     # the bench distribution: ~0.5 bits / symbol
        return np.clip(np.rint(rng.normal(7, 0.28, size=shape)), 0, 14).astype(np.int8)
    if kind == "uniform":
        return rng.integers(0, 31, size=shape).astype(np.int8)
    if kind == "uniform15":
        return rng.integers(0, 15, size=shape).astype(np.int8)
    if kind == "rare":
        return np.where(rng.random(shape) < 0.02, rng.integers(0, 31, shape), 7).astype(np.int8)
    if kind == "const":
        return np.full(shape, 9, np.int8)
    raise ValueError(kind)


def _check(sim, cdf, sym, tok0, g):
    """product encoder == oracle encoder byte for byte; product decoder inverts it from both halfword phases with
    garbage on either side of the stream; final state back at 2^16"""
    NL, _, C = sym.shape
    bs, ln = O.encode_group(cdf, sym, tok0, g, O.CODER_RANS)
    back = np.zeros((NL, sym.shape[1], C), np.uint8)
    O.decode_group(cdf, bs, ln, back, tok0, g, O.CODER_RANS)          # oracle decodes its own stream
    assert np.array_equal(back[:, tok0:tok0 + g], sym[:, tok0:tok0 + g].view(np.uint8))
    off = 0
    for nl in range(NL):
        for c in range(C):
            col = np.ascontiguousarray(sym[nl, tok0:tok0 + g, c])
            cd = np.ascontiguousarray(cdf[nl, c]).view(np.uint16)
            out = np.zeros(2 * g + 64, np.uint8)
            n = sim.sim_rans_encode_stream(P(cd), P(col), 1, g, P(out), out.size)
            ref = np.ascontiguousarray(bs[off:off + ln[nl, c]])
            assert n == ln[nl, c] and n % 2 == 0 and n >= 4 and np.array_equal(out[:n], ref)
            for odd in (0, 1):
                buf = np.concatenate([np.full(2 * odd, 0x5A, np.uint8), ref])
                dec = np.zeros(g, np.uint8)
                xf = sim.sim_rans_decode_stream(P(cd), ctypes.c_void_p(buf.ctypes.data + 2 * odd), ref.size, g, P(dec), 1,
                                                odd, 5)
                assert np.array_equal(dec, col.view(np.uint8)) and xf == 1 << 16, (nl, c, odd)
                if col.max() <= 15:
                    xf = sim.sim_rans_decode_stream(P(cd), ctypes.c_void_p(buf.ctypes.data + 2 * odd), ref.size, g,
                                                    P(dec), 1, odd, 4)
                    assert np.array_equal(dec, col.view(np.uint8)) and xf == 1 << 16, (nl, c, odd, "4-step")
            off += ln[nl, c]
    return bs, ln


@pytest.mark.parametrize("kind", ["peaked", "sharp", "uniform", "uniform15", "rare", "const"])
@pytest.mark.parametrize("t", [1, 2, 3, 7, 100, 236, 256])
def test_rans_product_core_vs_oracle(sim, kind, t):
    rng = np.random.default_rng(hash((kind, t, 7)) & 0xffff)
    sym = _symbols(rng, kind, (2, t, 5))
    _check(sim, O.cdf(sym), sym, 0, t)


def test_rans_foreign_cdf_groups(sim):
    """chunk-wide CDF (chunk > 256 tokens): a group may hold symbols that are rare chunk-wide, up to 16 bits each"""
    rng = np.random.default_rng(17)
    base = np.full((1, 8192, 6), 15, np.int8)
    base[:, :256, :] = rng.integers(0, 31, size=(1, 256, 6))
    base[:, 512:768, :] = np.where(rng.random((1, 256, 6)) < 0.03, 3, 15)
    cdf = O.cdf(base)
    for tok0 in (0, 256, 512, 4096):
        bs, ln = _check(sim, cdf, base, tok0, 256)
        assert ln.max() <= 4 + 2 * 256          # never more than one halfword per symbol


def test_rans_hand_made_cdfs(sim):
    """extreme frequency tables: one symbol owning almost everything, width-1 symbols, symbol 31 (upper bound 2^16)"""
    rng = np.random.default_rng(5)
    for case in range(60):
        nsym = int(rng.integers(1, 32))
        cuts = np.sort(rng.choice(np.arange(1, 65400), nsym - 1, replace=False)) if nsym > 1 else np.array([], int)
        cdf = np.zeros((1, 1, 33), np.uint16)
        cdf[0, 0, 1:nsym] = cuts
        cdf[0, 0, nsym:32] = 65450 + np.arange(32 - nsym)      # the remaining symbols keep width 1 .. the last one the rest
        g = int(rng.integers(1, 257))
        hi = 32 if case % 4 == 0 else nsym                       # every 4th case also codes the width-1 tail incl. symbol 31
        sym = rng.integers(0, hi, g).astype(np.int8).reshape(1, g, 1)
        _check(sim, cdf.view(np.int16), sym, 0, g)


def test_rans_code_length_and_row_bound(sim):
    """own-CDF streams of 256 symbols: length within [ideal, ideal + 32 + 16) bits, ideal = sum log2(65536 / freq) (a step
    can overshoot its ideal growth by < 1 bit and undershoot likewise; over a stream the deviations cancel to a fraction
    of a halfword), and the halfword count stays far inside the encoder's row capacity of 96 (DESIGN.md 3.7 proves
    <= 95 for ANY symbol order: ideal <= 1268.5 bits, < 1 bit of overshoot per step)"""
    rng = np.random.default_rng(11)
    worst = 0
    for kind in ["uniform", "uniform15", "peaked", "sharp", "rare"]:
        sym = _symbols(rng, kind, (1, 256, 64))
        cdf = O.cdf(sym)
        bs, ln = O.encode_group(cdf, sym, 0, 256, O.CODER_RANS)
        u = cdf.view(np.uint16).astype(np.int64)
        u[..., 32] = 65536
        for c in range(64):
            f = u[0, c, sym[0, :, c].astype(int) + 1] - u[0, c, sym[0, :, c].astype(int)]
            ideal = float(np.log2(65536.0 / f).sum())
            assert 8 * ln[0, c] < ideal + 32.0 + 16.0, (kind, c)
            assert 8 * ln[0, c] >= ideal - 1e-6                     # never below the information content
        worst = max(worst, int(ln.max()))
    # adversarial: 31 symbols as evenly as 256 tokens allow, in every rotation
    for rot in range(31):
        col = ((np.arange(256) + rot) % 31).astype(np.int8).reshape(1, 256, 1)
        cdf = O.cdf(col)
        _, ln = O.encode_group(cdf, col, 0, 256, O.CODER_RANS)
        worst = max(worst, int(ln.max()))
    assert worst <= 4 + 2 * 80, worst


def test_rans_overhead_vs_arithmetic_coder():
    """the price of the cheaper coder: the 32-bit final state instead of ~2 termination bits"""
    rng = np.random.default_rng(3)
    for kind, lo, hi in [("sharp", 1.5, 3.5), ("uniform15", 1.5, 3.5)]:
        sym = _symbols(rng, kind, (2, 256, 256))
        cdf = O.cdf(sym)
        _, ln_ac = O.encode_group(cdf, sym, 0, 256, O.CODER_AC)
        _, ln_r = O.encode_group(cdf, sym, 0, 256, O.CODER_RANS)
        extra = (ln_r.astype(np.int64) - ln_ac).mean()
        assert lo <= extra <= hi, (kind, extra)


def test_rans_divmod_device_formula_bounds(sim):
    """host build of rans_divmod is plain division; this checks the device estimate's error budget numerically: the
    biased reciprocal estimate is q or q - 1 for every (x, f) it can meet (x < f << 16)"""
    rng = np.random.default_rng(9)
    f = np.concatenate([rng.integers(1, 65536, 200000), np.array([1, 2, 3, 65535, 65505, 32768, 32769])]).astype(np.uint64)
    x = (rng.random(f.size) * (f.astype(np.float64) * 65536.0)).astype(np.uint64)
    x = np.minimum(np.maximum(x, 65536), f * 65536 - 1)
    x = np.where(f * 65536 - 1 >= 65536, x, f * 65536 - 1)
    # float32 model of the device code: fx = RZ(float(x)), rc within 1 ulp of 1/f (worst case on both sides), products RN
    fx = x.astype(np.float64)
    rn = fx.astype(np.float32)
    fx32 = np.where(rn.astype(np.float64) <= fx, rn, np.nextafter(rn, np.float32(0))).astype(np.float32)
    rc = (1.0 / f.astype(np.float32)).astype(np.float32)
    for bump in (-1, 0, 1):
        rcb = rc if bump == 0 else np.nextafter(rc, np.float32(np.inf if bump > 0 else 0)).astype(np.float32)
        est = np.floor((fx32 * (rcb * np.float32(0.99999952316284179688)).astype(np.float32)).astype(np.float32)).astype(np.int64)
        q = (x // f).astype(np.int64)
        assert ((est == q) | (est == q - 1)).all()
    r = np.zeros(1, np.uint32)
    for xi, fi in [(65536, 1), (2 ** 32 - 1, 65535), (123456789, 4321)]:
        got = sim.sim_rans_divmod(xi, fi, P(r))
        assert got == xi // fi and int(r[0]) == xi % fi


def test_safe_factor_keeps_every_symbol(sim):
    """quant_factor_safe (the fused encode kernel's clamp-free pass 1): an infinite factor becomes NaN, every other
    factor is untouched, and the symbols equal the checked quantiser's for ordinary and exotic rows alike."""
    S = ctypes.CDLL(os.path.join(HERE, "hostsim", "libhostsim.so"))
    S.sim_quant_row.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint16, ctypes.c_float, ctypes.c_void_p]
    S.sim_quant_row_safe.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint16, ctypes.c_float,
                                     ctypes.c_void_p, ctypes.c_void_p]
    rng = np.random.default_rng(21)
    C = 64
    rows = []
    base = O.f32_to_bf16_bits(rng.standard_normal(C).astype(np.float32))
    rows.append((base.copy(), 0))
    z = np.zeros(C, np.uint16); rows.append((z, 0))                                   # all-zero row: max 0 -> factor inf
    tiny = np.full(C, 0x0001, np.uint16); tiny[::2] = 0x8001; rows.append((tiny, 0))  # bf16 subnormals: MAX / max overflows
    inf = base.copy(); inf[3] = 0x7f80; inf[9] = 0xff80; rows.append((inf, 0))        # +-inf elements: max inf -> factor 0
    nan = base.copy(); nan[5] = 0x7fc1; rows.append((nan, 0))                         # NaN element: max NaN
    h = rng.standard_normal(C).astype(np.float16).view(np.uint16); rows.append((h, 1))
    hz = np.zeros(C, np.uint16); hz[1] = 0x0001; rows.append((hz, 1))                 # fp16 subnormal maximum (finite factor)
    for x, dt in rows:
        a = (x & 0x7fff).astype(np.uint16)
        f = (a.astype(np.uint32) << 16).view(np.float32) if dt == 0 else a.view(np.float16).astype(np.float32)
        mb = int(a[np.argmax(np.where(np.isnan(f), np.inf, f))]) if not np.isnan(f).any() else int(a[np.isnan(f)][0])
        for maxq in (7.0, 15.0):
            want = np.zeros(C, np.uint8)
            S.sim_quant_row(P(x), dt, C, mb, maxq, P(want))
            got = np.zeros(C, np.uint8)
            fb = np.zeros(1, np.uint32)
            S.sim_quant_row_safe(P(x), dt, C, mb, maxq, P(got), P(fb))
            assert np.array_equal(got, want)
            fm = np.float32(maxq) / ((np.array([mb], np.uint32) << 16).view(np.float32)[0] if dt == 0
                                     else np.array([mb], np.uint16).view(np.float16)[0].astype(np.float32))
            if np.isinf(fm):
                assert np.isnan(fb.view(np.float32)[0])
            elif not np.isnan(fm):
                assert fb[0] == np.array([fm], np.float32).view(np.uint32)[0]
