"""CPU: the C-ABI library loads and exports every symbol include/b200kv.h declares; host-side logic
(config, keys, protocol, factories, container views) behaves like the reference's."""
import os
import re

import pytest
import torch

from lmcache_b200 import _native as N
from lmcache_b200.config import GlobalConfig, LMCacheEngineConfig, LMCacheEngineMetadata
from lmcache_b200.protocol import ClientMetaMessage, Constants, ServerMetaMessage
from lmcache_b200.utils import CacheEngineKey

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported_and_bound():
    hdr = open(os.path.join(ROOT, "include", "b200kv.h")).read()
    declared = set(re.findall(r"\b(b200kv_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    L = N.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in b200kv.h but not exported"
    assert declared == set(N.SIGNATURES), declared ^ set(N.SIGNATURES)
    assert L.b200kv_version() == 4


def test_layout_arithmetic():
    lo = N.container_layout(32, 32, 128, 256)
    assert lo.off_cdf == 64
    assert lo.off_maxes == 64 + 64 * 4096 * 33 * 2
    assert lo.off_lengths == lo.off_maxes + 64 * 256 * 2
    assert lo.off_payload == lo.off_lengths + 64 * 4096 * 4
    with pytest.raises(N.NativeError):
        N.container_layout(0, 1, 1, 1)
    # compact container (version 3): nb map, maxes, u8 half-lengths; the histograms travel inside the streams
    lo = N.container_layout(32, 32, 128, 256, N.CODER_RANS_COMPACT)
    assert lo.off_cdf == 64
    assert lo.off_maxes == 64 + 64
    assert lo.off_lengths == lo.off_maxes + 64 * 256 * 2
    assert lo.off_payload == lo.fixed_bytes == lo.off_lengths + 64 * 4096
    assert lo.fixed_bytes < 0.02 * N.container_layout(32, 32, 128, 256).fixed_bytes
    assert N.container_layout(32, 32, 128, 256, N.CODER_AC).fixed_bytes == N.container_layout(32, 32, 128, 256).fixed_bytes
    with pytest.raises(N.NativeError):
        N.container_layout(32, 32, 128, 257, N.CODER_RANS_COMPACT)       # one <= 256-token group only
    with pytest.raises(N.NativeError):
        N.container_layout(2, 1, 8, 16, 3)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA device"):
        N.require_cuda()
    from lmcache_b200.codec import CacheGenCodec
    with pytest.raises(RuntimeError):
        CacheGenCodec("lmsys/longchat-7b-16k")
    from lmcache_b200.cache_engine import sha256_prefix_chain
    with pytest.raises(RuntimeError):
        sha256_prefix_chain(torch.arange(10), 4)


def test_key_string_roundtrip():
    k = CacheEngineKey("vllm", "mistralai/Mistral-7B-Instruct-v0.2", 3, 123, "ab" * 32)
    s = k.to_string()
    assert s == "vllm@mistralai/Mistral-7B-Instruct-v0.2@3@123@" + "ab" * 32
    assert CacheEngineKey.from_string(s) == k and hash(CacheEngineKey.from_string(s)) == hash(k)
    with pytest.raises(ValueError):
        CacheEngineKey.from_string("a@b@c")
    # SURVEY 8c known answer
    assert CacheEngineKey("vllm", "m", 1, 0, "bbd330b1").to_string() == "vllm@m@1@0@bbd330b1"


def test_protocol_headers():
    m = ClientMetaMessage(Constants.CLIENT_PUT, "some/key@1", 12345)
    s = m.serialize()
    assert len(s) == ClientMetaMessage.packlength() == 158
    assert ClientMetaMessage.deserialize(s) == m
    r = ServerMetaMessage(Constants.SERVER_SUCCESS, 77)
    assert len(r.serialize()) == ServerMetaMessage.packlength() == 8
    assert ServerMetaMessage.deserialize(r.serialize()) == r


def test_config_constructors(tmp_path):
    c = LMCacheEngineConfig.from_legacy(chunk_size=128, backend="cpu")
    assert (c.local_device, c.remote_url, c.chunk_size) == ("cpu", None, 128)
    c = LMCacheEngineConfig.from_legacy(backend="lm://localhost:65000", remote_serde="cachegen")
    assert (c.local_device, c.remote_url, c.remote_serde) == (None, "lm://localhost:65000", "cachegen")
    c = LMCacheEngineConfig.from_legacy(backend="file://local_disk/")
    assert c.local_device == "local_disk/"
    d = LMCacheEngineConfig.from_defaults()
    assert (d.chunk_size, d.local_device, d.remote_serde) == (256, "cuda", "torch")
    p = tmp_path / "c.yaml"
    p.write_text("chunk_size: 64\nlocal_device: cpu\nremote_url: lm://h:1\nremote_serde: cachegen\n")
    f = LMCacheEngineConfig.from_file(str(p))
    assert (f.chunk_size, f.local_device, f.remote_url, f.remote_serde) == (64, "cpu", "lm://h:1", "cachegen")
    p.write_text("local_device: tpu\n")
    with pytest.raises(ValueError):
        LMCacheEngineConfig.from_file(str(p))
    p.write_text("remote_url: nonsense\n")
    with pytest.raises(ValueError):
        LMCacheEngineConfig.from_file(str(p))
    assert GlobalConfig.is_debug() in (True, False)


def test_bins_table_matches_reference_goldens(golden):
    from lmcache_b200.storage_backend.serde.cachegen_basics import CacheGenConfig
    cfg = CacheGenConfig.from_model_name("lmsys/longchat-7b-16k")
    assert cfg.key_bins_list() == golden["key_bins"].tolist()
    assert cfg.value_bins_list() == golden["value_bins"].tolist()
    assert len(CacheGenConfig.from_model_name("THUDM/glm-4-9b-chat").key_bins_list()) == 40
    with pytest.raises(ValueError):
        CacheGenConfig.from_model_name("test_model")


def test_factories_error_behaviour():
    from lmcache_b200.storage_backend import CreateStorageBackend
    from lmcache_b200.storage_backend.serde import CreateSerde
    meta = LMCacheEngineMetadata("test_model", 1, 0, "vllm", "half")
    with pytest.raises(ValueError):
        CreateStorageBackend(LMCacheEngineConfig(256, None, None, "torch", False, False), meta)
    with pytest.raises(ValueError):
        CreateSerde("nonsense", LMCacheEngineConfig.from_defaults(), meta)
    s, d = CreateSerde("torch", LMCacheEngineConfig.from_defaults(), meta)
    t = torch.arange(24, dtype=torch.bfloat16).reshape(2, 3, 4)
    assert torch.equal(d.from_bytes(s.to_bytes(t)), t)        # config 1 plumbing: lossless on CPU tensors
    assert torch.equal(d.from_bytes(bytearray(s.to_bytes(t))), t)


def test_container_view_roundtrip_from_oracle_stream(golden):
    """A B2KV container assembled on the host from oracle output parses back to the same logical fields
    (the shape a reference consumer would unpickle, test_serde.py:60-62)."""
    import numpy as np
    from lmcache_b200.storage_backend.serde.cachegen_basics import (CacheGenGPUBytestream, CacheGenGPUEncoderOutput)
    from oracle import oracle as O
    x = golden["bf16_t300/x"]
    L, _, t, H, D = x.shape
    enc = O.encode_chunk(x.reshape(L, 2, t, H * D), 0, golden["key_bins"], golden["value_bins"])
    mk = torch.from_numpy(enc["maxes"][0].view(np.int16)).view(torch.bfloat16).reshape(L, t, 1)
    mv = torch.from_numpy(enc["maxes"][1].view(np.int16)).view(torch.bfloat16).reshape(L, t, 1)
    obj = CacheGenGPUEncoderOutput(
        [CacheGenGPUBytestream(torch.from_numpy(b), torch.from_numpy(ln), g) for b, ln, g in enc["groups"]],
        torch.from_numpy(enc["cdf"]), mk, mv, H, D)
    bs = obj.to_bytes()
    back = CacheGenGPUEncoderOutput.from_bytes(bs)
    assert back.num_heads == H and back.head_size == D and len(back.data_chunks) == 2
    assert torch.equal(back.cdf, obj.cdf) and torch.equal(back.max_tensors_key, mk)
    for a, b in zip(back.data_chunks, obj.data_chunks):
        assert a.ntokens == b.ntokens and torch.equal(a.bytestream, b.bytestream)
        assert torch.equal(a.bytestream_lengths, b.bytestream_lengths)
    with pytest.raises(ValueError):
        CacheGenGPUEncoderOutput.from_bytes(b"\0" * 100)


@pytest.mark.parametrize("name", ["bf16_t1", "bf16_t40", "bf16_t236", "bf16_t256", "bf16_uniform_t16", "fp16_uniform_t128"])
def test_compact_container_rebuilds_the_reference_cdf(golden, name):
    """A version-3 container has no CDF section: every stream carries its symbol histogram.  Assembled on the host from
    oracle output, it parses back to the CDF tensor THE REFERENCE'S OWN calculate_cdf spec produced (goldens), to the
    same lengths and bytestreams; the product's vectorised packer and the oracle's plain C one agree byte for byte."""
    import numpy as np
    from lmcache_b200.codec import parse_header
    from lmcache_b200.storage_backend.serde.cachegen_basics import (CacheGenGPUBytestream, CacheGenGPUEncoderOutput,
                                                                     cdf_from_counts)
    from oracle import oracle as O
    x = golden[f"{name}/x"]
    dt = int(golden[f"{name}/dtype"][0])
    L, _, t, H, D = x.shape
    kb, vb = golden["key_bins"], golden["value_bins"]
    enc = O.encode_chunk(x.reshape(L, 2, t, H * D), dt, kb, vb, O.CODER_RANS_COMPACT)
    assert np.array_equal(cdf_from_counts(enc["counts"], t), golden[f"{name}/cdf"])       # product host helper vs reference
    assert np.array_equal(O.cdf_from_counts(enc["counts"], t), golden[f"{name}/cdf"])     # oracle vs reference
    half = torch.bfloat16 if dt == 0 else torch.float16
    mk = torch.from_numpy(enc["maxes"][0].view(np.int16)).view(half).reshape(L, t, 1)
    mv = torch.from_numpy(enc["maxes"][1].view(np.int16)).view(half).reshape(L, t, 1)
    nb = O.nb_map(kb, vb, L)
    obj = CacheGenGPUEncoderOutput(
        [CacheGenGPUBytestream(torch.from_numpy(b), torch.from_numpy(ln), g) for b, ln, g in enc["groups"]],
        torch.from_numpy(enc["cdf"]), mk, mv, H, D, N.CODER_RANS_COMPACT, torch.from_numpy(enc["counts"].astype(np.int32)), nb)
    bs = obj.to_bytes()
    hd = parse_header(bs)
    assert hd.version == 3 and hd.nb == nb
    lo = N.container_layout(L, H, D, t, N.CODER_RANS_COMPACT)
    (b0, ln0, _), = enc["groups"]
    pl, half = O.v3_pack(enc["counts"], nb, ln0, b0)                                       # oracle's packer
    assert bs[lo.off_payload:] == pl.tobytes()
    assert bs[lo.off_lengths:lo.off_lengths + half.size] == half.tobytes()
    cnt2, ln2, r2 = O.v3_unpack(np.frombuffer(bs[lo.off_payload:], np.uint8), half, nb, t)  # ... and its parser
    assert np.array_equal(cnt2, enc["counts"]) and np.array_equal(ln2, ln0) and np.array_equal(r2, b0)
    v2 = CacheGenGPUEncoderOutput(obj.data_chunks, obj.cdf, mk, mv, H, D, N.CODER_RANS).to_bytes()
    assert len(bs) < len(v2)
    back = CacheGenGPUEncoderOutput.from_bytes(bs)
    assert back.coder == N.CODER_RANS_COMPACT
    assert np.array_equal(back.cdf.numpy(), golden[f"{name}/cdf"])
    assert torch.equal(back.max_tensors_key, mk) and torch.equal(back.max_tensors_value, mv)
    for a, b in zip(back.data_chunks, obj.data_chunks):
        assert a.ntokens == b.ntokens and torch.equal(a.bytestream, b.bytestream)
        assert torch.equal(a.bytestream_lengths, b.bytestream_lengths.to(torch.int32))
    # a damaged nb map, a truncated blob, a damaged stream header: ValueError (a miss), never a wrong layout
    bad = bytearray(bs)
    bad[64] = 33
    with pytest.raises(ValueError):
        parse_header(bytes(bad))
    with pytest.raises(ValueError):
        parse_header(bs[:-1])
    bad = bytearray(bs)
    bad[lo.off_payload:lo.off_payload + (nb[0] + 7) // 8] = bytes((nb[0] + 7) // 8)      # first stream: empty symbol mask
    with pytest.raises(ValueError):
        CacheGenGPUEncoderOutput.from_bytes(bytes(bad))


def test_compact_container_count_of_256():
    """all 256 tokens of a stream on one symbol: the count (256) does not fit a byte -- it is the implied one"""
    import numpy as np
    from lmcache_b200.storage_backend.serde.cachegen_basics import (CacheGenGPUBytestream, CacheGenGPUEncoderOutput)
    from oracle import oracle as O
    L, t, H, D = 2, 256, 1, 8
    kb = vb = np.array([32.0, 16.0], np.float32)
    x = np.zeros((L, 2, t, H * D), np.float32)
    x[..., 0] = 1.0                      # channel 0 pins every row maximum: constant symbols everywhere
    x[0, 0, ::2, 3] = -0.5               # ... except one stream with two symbols
    enc = O.encode_chunk(O.f32_to_bf16_bits(x), 0, kb, vb, O.CODER_RANS_COMPACT)
    assert enc["counts"].max() == 256
    mk = torch.from_numpy(enc["maxes"][0].view(np.int16)).view(torch.bfloat16).reshape(L, t, 1)
    mv = torch.from_numpy(enc["maxes"][1].view(np.int16)).view(torch.bfloat16).reshape(L, t, 1)
    obj = CacheGenGPUEncoderOutput(
        [CacheGenGPUBytestream(torch.from_numpy(b), torch.from_numpy(ln), g) for b, ln, g in enc["groups"]],
        torch.from_numpy(enc["cdf"]), mk, mv, H, D, N.CODER_RANS_COMPACT, torch.from_numpy(enc["counts"].astype(np.int32)),
        O.nb_map(kb, vb, L))
    back = CacheGenGPUEncoderOutput.from_bytes(obj.to_bytes())
    assert np.array_equal(back.counts.numpy(), enc["counts"].astype(np.int32))
    assert np.array_equal(back.cdf.numpy(), enc["cdf"])


def test_lazy_sequence_semantics():
    """cache_engine.LazySeq: the engine's chunk keys as a read-only sequence whose items are computed on access; slices stay
    lazy, iteration touches items front to back only as far as the consumer goes."""
    from lmcache_b200.cache_engine import LazySeq
    touched = []

    class Base:
        def __len__(self):
            return 10

        def __getitem__(self, i):
            touched.append(i)
            return i * i

    s = LazySeq(lambda v: v + 1, Base())
    assert len(s) == 10 and touched == []
    assert s[3] == 10 and s[-1] == 82 and touched == [3, 9]
    tail = s[4:]
    assert isinstance(tail, LazySeq) and len(tail) == 6 and touched == [3, 9]
    assert tail[0] == 17 and tail[1:3][1] == 37
    touched.clear()
    for i, v in enumerate(s):
        if i == 2:
            break
    assert touched == [0, 1, 2]
    assert list(s[8:]) == [65, 82] and list(s[5:5]) == [] and len(s[20:]) == 0
    with pytest.raises(IndexError):
        s[10]
    assert list(LazySeq(None, [])) == []


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_v3_stream_framing_product_vs_oracle_random(seed):
    """The host shim's vectorised version-3 stream packer / parser against the oracle's plain C ones on random histograms
    of every density (1 .. nb - 1 symbols in use, incl. lone symbols with 256 tokens), mixed plane widths, random rANS
    bytes: both directions byte-identical."""
    import numpy as np
    from lmcache_b200.storage_backend.serde.cachegen_basics import _v3_build_streams, _v3_parse_streams
    from oracle import oracle as O
    rng = np.random.default_rng(seed)
    nb = [32, 16, 8, 30, 16, 32]
    C, t = 37, int(rng.integers(1, 257))
    NL = len(nb)
    cnt = np.zeros((NL, C, 33), np.uint32)
    for nl in range(NL):
        for c in range(C):
            k = int(rng.integers(1, min(nb[nl] - 1, t) + 1))
            syms = rng.choice(nb[nl] - 1, size=k, replace=False)
            cnt[nl, c, syms] = 1
            extra = rng.multinomial(t - k, np.ones(k) / k)
            cnt[nl, c, syms] += extra.astype(np.uint32)
    assert np.all(cnt.sum(axis=2) == t)
    rlen = (4 + 2 * rng.integers(0, 40, size=(NL, C))).astype(np.int32)
    rans = rng.integers(0, 256, size=int(rlen.sum()), dtype=np.uint8)
    pl_o, half_o = O.v3_pack(cnt, nb, rlen, rans)
    pl_p, half_p = _v3_build_streams(cnt.reshape(NL * C, 33).astype(np.int32), nb, C, rlen.reshape(-1), rans)
    assert pl_p.tobytes() == pl_o.tobytes() and half_p.tobytes() == half_o.reshape(-1).tobytes()
    c2, l2, r2 = _v3_parse_streams(pl_o, half_o.reshape(-1), nb, C, t)
    c3, l3, r3 = O.v3_unpack(pl_o, half_o, nb, t)
    assert np.array_equal(c2.reshape(NL, C, 33), cnt.astype(np.int32)) and np.array_equal(c3, cnt)
    assert np.array_equal(l2.reshape(NL, C), rlen) and np.array_equal(l3, rlen)
    assert r2.tobytes() == rans.tobytes() == r3.tobytes()
