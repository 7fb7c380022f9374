"""GPU: the compressed page-locked host tier (local_device="cpu", local_serde="cachegen") behind
LMCacheEngine.store()/retrieve() -- BASELINE configs[2]'s offload + reload path.  CacheGen is lossy by design, so the
bar is the reference's own decode: every retrieved token must equal bf16/fp16( do_dequantize( torch_quant_vectorized(x) ) )
of the stored chunk bit for bit (tests/ref_torch.py restates that op chain; the C oracle is the second witness)."""
import threading

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
MODEL = "lmsys/longchat-7b-16k"


def _meta(fmt="vllm"):
    from lmcache_b200.config import LMCacheEngineMetadata
    return LMCacheEngineMetadata(MODEL, 1, 0, fmt, "bfloat16")


def _cfg(chunk_size=256):
    from lmcache_b200.config import LMCacheEngineConfig
    return LMCacheEngineConfig.from_legacy(chunk_size=chunk_size, backend="cpu", local_serde="cachegen")


def _kv(T, fmt, L=6, H=2, D=128, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    shape = (T, H, D) if fmt == "vllm" else (H, T, D)
    dt = torch.bfloat16 if fmt == "vllm" else torch.float16
    return tuple((torch.randn(shape, device="cuda", generator=g).to(dt), torch.randn(shape, device="cuda", generator=g).to(dt))
                 for _ in range(L))


def _want(kv, fmt, cs, T):
    """reference decode of every chunk: quantise + dequantise + cast, chunk by chunk (per-chunk scales and CDFs);
    returns the blob in the engine's layout ([L,2,T,H,D] vllm / [L,2,H,T,D] huggingface)"""
    import ref_torch
    kb, vb = (torch.tensor(b) for b in O.make_bins(MODEL))
    blob = torch.stack((torch.stack([k for k, _ in kv]), torch.stack([v for _, v in kv]))).permute(1, 0, 2, 3, 4)
    if fmt == "huggingface":
        blob = blob.permute(0, 1, 3, 2, 4)               # -> [L,2,T,H,D], the layout the reference quantises in
    outs = [ref_torch.roundtrip(c.contiguous(), kb, vb, fmt) for c in torch.split(blob[:, :, :T], cs, dim=2)]
    return torch.cat(outs, dim=2 if fmt == "vllm" else 3)


def _blob_of(ret):
    return torch.stack((torch.stack([k for k, _ in ret]), torch.stack([v for _, v in ret]))).permute(1, 0, 2, 3, 4)


@pytest.mark.parametrize("fmt", ["vllm", "huggingface"])
@pytest.mark.parametrize("blocking", [True, False])
@pytest.mark.parametrize("cs", [256, 100])
def test_compressed_tier_store_retrieve_matches_reference_decode(fmt, blocking, cs, autorelease):
    from lmcache_b200.cache_engine import LMCacheEngine
    from lmcache_b200.storage_backend.local_backend import LMCLocalCompressedBackend
    T = 2300                                             # 9 chunks of 256 (3 waves of 4) + a ragged tail
    tokens = torch.randint(0, 32000, (T,), device="cuda")
    kv = _kv(T, fmt)
    engine = autorelease(LMCacheEngine(_cfg(cs), _meta(fmt)))
    assert isinstance(engine.engine_, LMCLocalCompressedBackend)
    r0, m0 = engine.retrieve(tokens)
    assert len(r0) == 0 and int(m0.sum()) == 0
    engine.store(tokens, kv, blocking=blocking)
    ret, mask = engine.retrieve(tokens)
    torch.cuda.synchronize()
    assert int(mask.sum()) == T and engine.engine_.codec.decode_status() == [0] * len(engine.engine_.codec.decode_status())
    want = _want(kv, fmt, cs, T)
    got = _blob_of(ret)
    assert got.dtype == (torch.bfloat16 if fmt == "vllm" else torch.float16)
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))
    # host memory holds containers, not raw KV
    raw = sum(k.numel() * 2 * 2 for k, _ in kv)
    assert 0 < engine.engine_.host_bytes() < 0.75 * raw


def test_compressed_tier_prefix_mask_and_retrieve_only_replica(autorelease):
    """reference prefix / suffix-mask semantics (tests/test_cache_engine.py:108-254) on the compressed tier, and a second
    engine object that never stored anything (geometry comes from the container header, no chunk is decoded twice)"""
    from lmcache_b200.cache_engine import LMCacheEngine
    cs, T = 256, 1000
    tokens = torch.randint(0, 32000, (T,), device="cuda")
    kv = _kv(T, "vllm", seed=3)
    engine = autorelease(LMCacheEngine(_cfg(cs), _meta()))
    engine.store(tokens, kv)
    longer = torch.cat([tokens, torch.randint(0, 32000, (500,), device="cuda")])
    ret, mask = engine.retrieve(longer)                  # the 232-token tail chunk was stored under the 1000-token chain
    assert int(mask.sum()) == 768 and ret[0][0].shape[0] == 768
    want = _want(kv, "vllm", cs, 768)
    assert torch.equal(_blob_of(ret).view(torch.int16), want.view(torch.int16))
    m = torch.ones(T, dtype=torch.bool)
    m[:300] = False
    ret, mask = engine.retrieve(tokens, m)
    assert int(mask.sum()) == 700 and ret[0][0].shape[0] == 700 and int(mask.nonzero()[0]) == 300
    full = _want(kv, "vllm", cs, T)
    assert torch.equal(_blob_of(ret).view(torch.int16), full[:, :, 300:].contiguous().view(torch.int16))
    # a replica that only retrieves: shares the backend object, owns no geometry yet
    replica = LMCacheEngine.__new__(LMCacheEngine)
    replica.__dict__.update(engine.__dict__)
    replica.__dict__.pop("_geom", None)
    calls = []
    orig = engine.engine_.get
    engine.engine_.get = lambda k: calls.append(k) or orig(k)
    ret, mask = replica.retrieve(tokens)
    assert int(mask.sum()) == T and calls == []
    assert torch.equal(_blob_of(ret).view(torch.int16), full.view(torch.int16))
    r4, m4 = engine.retrieve(torch.randint(0, 32000, (300,), device="cuda"))
    assert len(r4) == 0 and int(m4.sum()) == 0


def test_compressed_tier_overwrite_reuses_slab_and_generic_put_get(autorelease):
    from lmcache_b200.cache_engine import LMCacheEngine
    from lmcache_b200.utils import CacheEngineKey
    cs, T = 256, 1024
    tokens = torch.randint(0, 32000, (T,), device="cuda")
    engine = autorelease(LMCacheEngine(_cfg(cs), _meta()))
    be = engine.engine_
    for it in range(4):
        kv = _kv(T, "vllm", seed=10 + it)
        engine.store(tokens, kv, skip_existing=False)
        used = be.host_bytes()
        if it:
            assert used < 1.3 * first              # the old containers were freed, not piled up
        else:
            first = used
    ret, mask = engine.retrieve(tokens)
    assert torch.equal(_blob_of(ret).view(torch.int16), _want(kv, "vllm", cs, T).view(torch.int16))
    assert be.slab.stats()[0] == 1                  # one page-locked segment serves everything
    # the plain backend interface: put / contains / get of one chunk blob (what a third-party engine would call)
    blob = torch.randn(6, 2, 200, 2, 128, device="cuda").to(torch.bfloat16)
    key = CacheEngineKey("vllm", MODEL, 1, 0, "deadbeef")
    assert not be.contains(key) and be.get(key) is None
    be.put(key, blob, blocking=False)
    assert be.contains(key)
    got = be.get(key)
    import ref_torch
    kb, vb = (torch.tensor(b) for b in O.make_bins(MODEL))
    assert torch.equal(got.view(torch.int16), ref_torch.roundtrip(blob, kb, vb, "vllm").view(torch.int16))


def test_compressed_tier_nonblocking_store_is_stream_ordered_snapshot(autorelease):
    """store(blocking=False) must capture the KV as it is when store() returns (reference: chunks are materialised
    before enqueuing, cache_engine.py:274-275): overwriting the tensors right afterwards on the same stream must not
    leak into the stored chunks."""
    from lmcache_b200.cache_engine import LMCacheEngine
    cs, T = 256, 3000
    tokens = torch.randint(0, 32000, (T,), device="cuda")
    kv = _kv(T, "vllm", seed=5)
    want = _want(kv, "vllm", cs, T)
    engine = autorelease(LMCacheEngine(_cfg(cs), _meta()))
    engine.store(tokens, kv, blocking=False)
    for k, v in kv:                                  # the caller recycles its buffers immediately
        k.zero_()
        v.fill_(7.0)
    ret, mask = engine.retrieve(tokens)
    assert int(mask.sum()) == T
    assert torch.equal(_blob_of(ret).view(torch.int16), want.view(torch.int16))


def test_compressed_tier_concurrent_store_and_retrieve(autorelease):
    """one thread stores new sequences while another retrieves an old one (the reference's put_worker / caller split)"""
    from lmcache_b200.cache_engine import LMCacheEngine
    cs, T = 256, 1536
    engine = autorelease(LMCacheEngine(_cfg(cs), _meta()))
    tok0 = torch.randint(0, 32000, (T,), device="cuda")
    kv0 = _kv(T, "vllm", seed=1)
    engine.store(tok0, kv0)
    want0 = _want(kv0, "vllm", cs, T)
    errs = []

    def writer():
        try:
            torch.cuda.set_device(0)
            for i in range(4):
                engine.store(torch.randint(0, 32000, (T,), device="cuda"), _kv(T, "vllm", seed=100 + i), blocking=(i % 2 == 0))
        except Exception as e:      # noqa: BLE001
            errs.append(e)

    th = threading.Thread(target=writer)
    th.start()
    for _ in range(6):
        ret, mask = engine.retrieve(tok0)
        assert int(mask.sum()) == T
        assert torch.equal(_blob_of(ret).view(torch.int16), want0.view(torch.int16))
    th.join()
    assert not errs


def test_disk_tier_store_restart_retrieve(tmp_path, autorelease):
    """local_device = file://<dir>/ : chunks are B2KV container files; a second engine started on the same directory (a
    restart) rebuilds the index from the file headers and serves the chunks; damaged and foreign files are ignored."""
    import os

    from lmcache_b200.cache_engine import LMCacheEngine
    from lmcache_b200.config import LMCacheEngineConfig
    from lmcache_b200.storage_backend.local_backend import LMCLocalDiskBackend
    cs, T = 256, 1400
    d = str(tmp_path / "kvdisk") + "/"
    cfg = LMCacheEngineConfig.from_legacy(chunk_size=cs, backend="file://" + d)
    assert cfg.local_device == d
    tokens = torch.randint(0, 32000, (T,), device="cuda")
    kv = _kv(T, "vllm", seed=8)
    want = _want(kv, "vllm", cs, T)
    engine = LMCacheEngine(cfg, _meta())
    assert isinstance(engine.engine_, LMCLocalDiskBackend)
    engine.store(tokens, kv, blocking=False)
    ret, mask = engine.retrieve(tokens)                       # read-your-writes: waits for the files
    assert int(mask.sum()) == T
    assert torch.equal(_blob_of(ret).view(torch.int16), want.view(torch.int16))
    files = sorted(f for f in os.listdir(d) if f.endswith(".b2kv"))
    assert len(files) == 6 and not [f for f in os.listdir(d) if f.endswith(".tmp")]
    raw = sum(k.numel() * 2 * 2 for k, _ in kv)
    assert sum(os.path.getsize(d + f) for f in files) < 0.75 * raw          # containers, not raw blobs
    # restart: damage the file of chunk 3 (truncate), drop a foreign file in
    victim = engine.engine_._key_to_path(engine._make_key(engine._prefix_hash(tokens)[3], "vllm"))
    assert os.path.basename(victim) in files
    engine.close()
    with open(victim, "r+b") as f:
        f.truncate(os.path.getsize(victim) // 2)
    open(d + "notes.b2kv", "wb").write(b"not a container")
    engine2 = autorelease(LMCacheEngine(cfg, _meta()))
    assert len(engine2.engine_.dict) == 5
    ret, mask = engine2.retrieve(tokens)
    torch.cuda.synchronize()
    # chunks are matched front to back: everything before the damaged chunk comes back, nothing after it
    hashes = engine2._prefix_hash(tokens)
    bad = [i for i, h in enumerate(hashes) if engine2.engine_._key_to_path(engine2._make_key(h, "vllm")) == victim][0]
    assert int(mask.sum()) == bad * cs
    assert torch.equal(_blob_of(ret).view(torch.int16), want[:, :, :bad * cs].contiguous().view(torch.int16))
    # storing again repairs the cache (skip_existing stops at the damaged chunk and rewrites from there)
    engine2.store(tokens, kv)
    ret, mask = engine2.retrieve(tokens)
    assert int(mask.sum()) == T
    assert torch.equal(_blob_of(ret).view(torch.int16), want.view(torch.int16))
