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
    assert L.b200kv_version() == 3


def test_layout_arithmetic():
    lo = N.container_layout(32, 32, 128, 256)
    assert lo.off_cdf == 64
    assert lo.off_maxes == 64 + 64 * 4096 * 33 * 2
    assert lo.off_lengths == lo.off_maxes + 64 * 256 * 2
    assert lo.off_payload == lo.off_lengths + 64 * 4096 * 4
    with pytest.raises(N.NativeError):
        N.container_layout(0, 1, 1, 1)


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
