"""CPU: the oracle (oracle/cachegen_oracle.c) against golden vectors generated from the reference's own
functions (tests/golden/make_golden.py).  This is what pins the oracle."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ["bf16_t1", "bf16_t7_L32", "bf16_t40", "bf16_t236", "bf16_t256", "bf16_t300", "bf16_uniform_t16", "fp16_t40",
         "fp16_uniform_t128"]


def _case(golden, n):
    x = golden[f"{n}/x"]
    dt = int(golden[f"{n}/dtype"][0])
    L, _, t, H, D = x.shape
    return x.reshape(L, 2, t, H * D), dt, (L, t, H, D)


@pytest.mark.parametrize("name", CASES)
def test_quantize_matches_reference(golden, name):
    x, dt, (L, t, H, D) = _case(golden, name)
    sym, maxes = O.quantize(x, dt, golden["key_bins"], golden["value_bins"])
    assert np.array_equal(sym, golden[f"{name}/sym"])
    assert np.array_equal(maxes[0], golden[f"{name}/max_k"].reshape(L, t))
    assert np.array_equal(maxes[1], golden[f"{name}/max_v"].reshape(L, t))


@pytest.mark.parametrize("name", CASES)
def test_cdf_matches_in_tree_spec(golden, name):
    assert np.array_equal(O.cdf(golden[f"{name}/sym"]), golden[f"{name}/cdf"])


@pytest.mark.parametrize("name", CASES)
def test_dequantize_matches_reference(golden, name):
    x, dt, (L, t, H, D) = _case(golden, name)
    sym = golden[f"{name}/sym"].view(np.uint8)
    maxes = np.stack([golden[f"{name}/max_k"].reshape(L, t), golden[f"{name}/max_v"].reshape(L, t)])
    kb, vb = golden["key_bins"], golden["value_bins"]
    bf = O.dequantize(sym, maxes, dt, kb, vb, O.DT_BF16).reshape(L, 2, t, H, D)
    assert np.array_equal(bf, golden[f"{name}/deq_vllm_bf16"])
    hf = O.dequantize(sym, maxes, dt, kb, vb, O.DT_FP16).reshape(L, 2, t, H, D).transpose(0, 1, 3, 2, 4)
    assert np.array_equal(hf, golden[f"{name}/deq_hf_fp16"])


@pytest.mark.parametrize("name", CASES)
def test_coder_roundtrip_and_entropy(golden, name):
    """Bitstream parity is unpinned (torchac_cuda absent): the coder must be the identity on symbols and
    within ~2 bits/stream + model slack of the empirical entropy under its own CDF."""
    sym = golden[f"{name}/sym"]
    cdf = golden[f"{name}/cdf"]
    NL, t, C = sym.shape
    out = np.zeros((NL, t, C), np.uint8)
    total_bits = 0
    for tok0 in range(0, t, O.GROUP):
        g = min(O.GROUP, t - tok0)
        bs, ln = O.encode_group(cdf, sym, tok0, g)
        assert ln.sum() == bs.size
        O.decode_group(cdf, bs, ln, out, tok0, g)
        total_bits += 8 * bs.size
    assert np.array_equal(out, sym.view(np.uint8))
    # ideal code length under the transmitted CDF
    cu = cdf.view(np.uint16).astype(np.int64)
    width = np.diff(np.concatenate([cu[..., :32], np.full(cu.shape[:-1] + (1,), 65536)], axis=-1), axis=-1)
    w = np.take_along_axis(width[:, None, :, :].repeat(t, 1), sym.astype(np.int64)[..., None], axis=-1)[..., 0]
    ideal = float(np.sum(16.0 - np.log2(w)))
    ngroups = (t + O.GROUP - 1) // O.GROUP
    slack = NL * C * ngroups * (2 + 8)   # 2 termination bits + byte padding per stream
    assert total_bits <= ideal + slack


def test_hash_chain_matches_reference():
    h = json.load(open(os.path.join(HERE, "golden", "golden_hash.json")))
    rng = np.random.default_rng(1234)
    for c in h["cases"]:
        if c["label"].startswith("arange"):
            toks = np.arange(c["n"], dtype=c["dtype"])
        else:
            toks = rng.integers(0, 32000, c["n"], dtype=np.int64)
        assert O.sha256_chain(toks, c["chunk_size"]) == c["hashes"], c["label"]


def test_hash_known_answers():
    """SURVEY.md 8c known answers produced by the reference _prefix_hash."""
    h = O.sha256_chain(np.arange(600, dtype=np.int64), 256)
    assert h[0] == "bbd330b12e8159e117376ef24fa106413bc9fc18032a0d43e95c5dae5e47953f"
    assert h[1] == "da67b0aaefba655d2edadd2cc5d11cd4564db9059ccf6264056d62d170b11ff5"
    assert h[2] == "02fe4699223cef67e6c17c405b9795d728114ed08533bd071c0787aafa7c34fc"
    assert O.sha256_chain(np.arange(600, dtype=np.int32), 256)[0] == \
        "8808405eec6fbe306fe3369f88daed79dd5613ddbb5e801f632b01d6218c5f08"


def test_sha256_against_hashlib():
    import hashlib
    rng = np.random.default_rng(5)
    for n in [1, 7, 55, 56, 63, 64, 65, 119, 120, 1000]:
        toks = rng.integers(0, 256, n, dtype=np.uint8)
        assert O.sha256_chain(toks, 1 << 20)[0] == hashlib.sha256(toks.tobytes()).hexdigest()


def test_ref_torch_helper_pinned_to_goldens(golden, golden_names):
    """tests/ref_torch.py (the torch restatement the full-size GPU parity tests and bench.py's spot check lean on) against
    the vectors generated from the reference's own torch_quant_vectorized / do_dequantize (tests/golden/make_golden.py)."""
    import sys
    import os
    import torch
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import ref_torch
    kb, vb = torch.from_numpy(golden["key_bins"]), torch.from_numpy(golden["value_bins"])
    for n in golden_names:
        x = golden[f"{n}/x"]
        dt = torch.bfloat16 if int(golden[f"{n}/dtype"][0]) == 0 else torch.float16
        blob = torch.from_numpy(x.view(np.int16).copy()).view(dt)
        L, _, t, H, D = blob.shape
        sym, mk, mv = ref_torch.quantize(blob, kb, vb)
        assert np.array_equal(sym.numpy(), golden[f"{n}/sym"]), n
        assert np.array_equal(mk.view(torch.int16).numpy().view(np.uint16).reshape(L, t), golden[f"{n}/max_k"].reshape(L, t)), n
        assert np.array_equal(mv.view(torch.int16).numpy().view(np.uint16).reshape(L, t), golden[f"{n}/max_v"].reshape(L, t)), n
        for fmt, key in (("vllm", "deq_vllm_bf16"), ("huggingface", "deq_hf_fp16")):
            out = ref_torch.roundtrip(blob, kb, vb, fmt)
            got = out.contiguous().view(torch.int16).numpy().view(np.uint16)
            want = golden[f"{n}/{key}"]
            fa = got.astype(np.uint32) << 16 if fmt == "vllm" else None
            same = got == want.reshape(got.shape)
            if not same.all():                      # NaN payloads may differ; nothing else may
                g = (got.astype(np.uint32) << 16).view(np.float32) if fmt == "vllm" else got.view(np.float16)
                w = (want.reshape(got.shape).astype(np.uint32) << 16).view(np.float32) if fmt == "vllm" else want.reshape(got.shape).view(np.float16)
                assert (same | (np.isnan(g) & np.isnan(w))).all(), (n, fmt)
