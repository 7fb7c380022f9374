#!/usr/bin/env python
"""Secondary measurements on the B200 box (not the headline bench): SHA-256 prefix-hash kernel, the host mover
(LMCLocalBackend cpu tier through the engine, BASELINE configs[2] shape scaled to fit a quick run), and the
engine-level store/retrieve through the serde plugin boundary (lm:// + cachegen).  Prints one JSON object.

    python profiles/extra_bench.py > gpurun_out/extra.json
"""
import hashlib
import json
import os
import socket
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

ge.build_cuda()
from lmcache_b200.cache_engine import LMCacheEngine, sha256_prefix_chain  # noqa: E402
from lmcache_b200.config import LMCacheEngineConfig, LMCacheEngineMetadata  # noqa: E402

torch.cuda.set_device(0)
out = {}


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts), sum(ts) / len(ts)


# ---------------- hash chain (a1)
g = torch.Generator().manual_seed(5)
for label, n, seqs in [("8192_tokens_1_chain", 8192, None), ("65536_tokens_1_chain", 65536, None),
                       ("16x4096_tokens_16_chains", 16 * 4096, [i * 4096 for i in range(17)])]:
    toks = torch.randint(0, 32000, (n,), generator=g, dtype=torch.int64)
    dtoks = toks.cuda()
    best, mean = timeit(lambda: sha256_prefix_chain(dtoks, 256, seqs))

    def cpu_ref():
        offs = seqs or [0, n]
        outp = []
        for i in range(len(offs) - 1):
            pre = ""
            for a in range(offs[i], offs[i + 1], 256):
                pre = hashlib.sha256(pre.encode("ascii") + toks[a:min(a + 256, offs[i + 1])].numpy().tobytes()).hexdigest()
                outp.append(pre)
        return outp
    t0 = time.perf_counter()
    want = cpu_ref()
    tcpu = time.perf_counter() - t0
    assert sha256_prefix_chain(dtoks, 256, seqs) == want
    out["hash_" + label] = {"gpu_us_best": round(best * 1e6, 1), "gpu_us_mean": round(mean * 1e6, 1),
                            "cpu_hashlib_us": round(tcpu * 1e6, 1), "chunks": len(want)}

if os.environ.get('EXTRA_ONLY') == 'hash':
    print(json.dumps(out, indent=1))
    sys.exit(0)

# ---------------- paged KV cache in place vs gather + dense, 8192 tokens x 32L x 32H x 128D = 4 GiB
def paged_section():
    from lmcache_b200.codec import CacheGenCodec, KvView
    L, H, D, T, bs = 32, 32, 128, 8192, 16
    raw = 2 * L * T * H * D * 2
    nrows = 2 * T
    gg = torch.Generator().manual_seed(11)
    slots = torch.randperm(nrows, generator=gg)[:T].cuda()
    sigma = torch.rand(1, H, D, device="cuda") * 3 + 0.2
    caches = [((torch.randn(nrows // bs, bs, H, D, device="cuda") * sigma).to(torch.bfloat16),
               (torch.randn(nrows // bs, bs, H, D, device="cuda") * sigma).to(torch.bfloat16)) for _ in range(L)]
    codec = CacheGenCodec("lmsys/longchat-7b-16k")
    pview = KvView.from_paged(caches, slots)
    res = {}

    def gather():
        return tuple((k.view(-1, H, D)[slots], v.view(-1, H, D)[slots]) for k, v in caches)
    g_best, _ = timeit(gather, n=3, warm=1)
    dense = gather()
    dview = KvView.from_tuple(dense, "vllm")
    stride = codec.out_stride(L, H, D, 256)
    outbuf = torch.empty(stride * (T // 256), dtype=torch.uint8, device="cuda")
    ep, _ = timeit(lambda: codec.encode(pview, 0, T, 256, out=outbuf), n=3, warm=1)
    bp = codec.encode(pview, 0, T, 256, out=outbuf)
    sizes_p = list(bp.sizes)
    ref = outbuf.clone()
    ed, _ = timeit(lambda: codec.encode(dview, 0, T, 256, out=outbuf), n=3, warm=1)
    bd = codec.encode(dview, 0, T, 256, out=outbuf)
    torch.cuda.synchronize()
    same = sizes_p == list(bd.sizes) and all(
        torch.equal(ref[j * stride + 64: j * stride + sizes_p[j]], outbuf[j * stride + 64: j * stride + sizes_p[j]])
        for j in range(len(sizes_p)))
    nt = [256] * (T // 256)
    toks_off = [j * 256 for j in range(T // 256)]
    dp, _ = timeit(lambda: codec.decode_device_batch(bd, nt, pview, toks_off), n=3, warm=1)
    dblob = torch.empty((L, 2, T, H, D), dtype=torch.bfloat16, device="cuda")
    dd, _ = timeit(lambda: codec.decode_device_batch(bd, nt, KvView.from_blob(dblob, "vllm"), toks_off), n=3, warm=1)

    def scatter():
        for l, (k, v) in enumerate(caches):
            k.view(-1, H, D)[slots] = dblob[l, 0]
            v.view(-1, H, D)[slots] = dblob[l, 1]
    s_best, _ = timeit(scatter, n=3, warm=1)
    res = {"raw_bytes": raw, "containers_identical_to_dense": bool(same),
           "encode_paged_ms": round(ep * 1e3, 2), "encode_dense_ms": round(ed * 1e3, 2), "torch_gather_ms": round(g_best * 1e3, 2),
           "decode_paged_ms": round(dp * 1e3, 2), "decode_dense_ms": round(dd * 1e3, 2), "torch_scatter_ms": round(s_best * 1e3, 2),
           "store_side_GBps_paged": round(raw / ep / 1e9, 1), "store_side_GBps_gather_plus_dense": round(raw / (g_best + ed) / 1e9, 1),
           "load_side_GBps_paged": round(raw / dp / 1e9, 1), "load_side_GBps_dense_plus_scatter": round(raw / (dd + s_best) / 1e9, 1),
           "note": "wall clock incl. launch + sync per call; gather/scatter = per-layer torch indexing as lmcache-vllm does"}
    return res


if os.environ.get('EXTRA_ONLY') in (None, 'paged'):
    out["paged_kv_in_place"] = paged_section()
    torch.cuda.empty_cache()
if os.environ.get('EXTRA_ONLY') == 'paged':
    print(json.dumps(out, indent=1))
    sys.exit(0)

# ---------------- raw lm:// connector throughput on loopback: native (csrc/lmnet.cu) and pure-Python, all pairings
def lm_raw_section():
    import ctypes
    import threading
    from lmcache_b200 import _native as N
    from lmcache_b200.server.__main__ import LMCacheServer
    from lmcache_b200.storage_backend.connector.lm_connector import LMCServerConnector
    from lmcache_b200.storage_backend.connector.native_connector import LMCNativeConnector
    lib = N.lib()
    h = ctypes.c_void_p()
    N.check(lib.b200kv_lm_server_start(b"127.0.0.1", 0, ctypes.byref(h)))
    nport = lib.b200kv_lm_server_port(h)
    srv = LMCacheServer("127.0.0.1", 0)
    pport = srv.sock.getsockname()[1]
    threading.Thread(target=srv.run, daemon=True).start()
    blob = bytes(bytearray(os.urandom(1 << 20)) * 22)           # one 22 MiB container, about a 256-token chunk
    res = {}
    for sname, port in (("native_server", nport), ("python_server", pport)):
        for cname, C in (("native_client", LMCNativeConnector), ("python_client", LMCServerConnector)):
            c = C("127.0.0.1", port)
            n, best_put, best_get = 16, 1e9, 1e9
            for rep in range(3):
                t0 = time.perf_counter()
                for i in range(n):
                    c.set(f"k{rep}_{i}", blob)
                while not c.exists(f"k{rep}_{n - 1}"):
                    pass
                t1 = time.perf_counter()
                for i in range(n):
                    assert len(c.get(f"k{rep}_{i}")) == len(blob)
                t2 = time.perf_counter()
                best_put, best_get = min(best_put, t1 - t0), min(best_get, t2 - t1)
            res[f"{sname}+{cname}"] = {"put_GBps": round(n * len(blob) / best_put / 1e9, 2),
                                       "get_GBps": round(n * len(blob) / best_get / 1e9, 2)}
            c.close()
    lib.b200kv_lm_server_stop(h)
    srv.sock.close()
    res["note"] = "16 x 22 MiB values over one loopback TCP connection, best of 3; bytes on the wire"
    return res


if os.environ.get('EXTRA_ONLY') in (None, 'lm'):
    out["lm_connector_raw"] = lm_raw_section()

# ---------------- host mover through the engine (local cpu tier), 8192 tokens x 32L x 32H x 128D = 4 GiB
L, H, D, T = 32, 32, 128, 8192
kv = tuple((torch.randn(T, H, D, device="cuda").to(torch.bfloat16), torch.randn(T, H, D, device="cuda").to(torch.bfloat16))
           for _ in range(L))
raw = 2 * L * T * H * D * 2
toks = torch.randint(0, 32000, (T,), generator=g, dtype=torch.int64).cuda()
for backend in (("cpu", "cuda") if os.environ.get('EXTRA_ONLY') is None else ()):
    meta = LMCacheEngineMetadata("test_model", 1, 0, "vllm", "bfloat16")
    eng = LMCacheEngine(LMCacheEngineConfig.from_legacy(chunk_size=256, backend=backend), meta)
    t0 = time.perf_counter()
    eng.store(toks, kv, skip_existing=False)
    torch.cuda.synchronize()
    t_first = time.perf_counter() - t0
    st_best, _ = timeit(lambda: eng.store(toks, kv, skip_existing=False), n=3, warm=1)
    rt_best, _ = timeit(lambda: eng.retrieve(toks), n=3, warm=1)
    r, m = eng.retrieve(toks)
    assert int(m.sum()) == T and torch.equal(r[3][1], kv[3][1])
    out[f"engine_local_{backend}"] = {"store_GBps": round(raw / st_best / 1e9, 1), "retrieve_GBps": round(raw / rt_best / 1e9, 1),
                                      "first_store_s": round(t_first, 3), "raw_bytes": raw}
    eng.close()
    del eng

# ---------------- engine over lm:// + cachegen (serde plugin boundary), 2048 tokens
s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
env = dict(os.environ, PYTHONPATH=ROOT)
srv = subprocess.Popen([sys.executable, "-m", "lmcache_b200.server", "127.0.0.1", str(port)], env=env)
for _ in range(100):
    try:
        socket.create_connection(("127.0.0.1", port), timeout=0.2).close(); break
    except OSError:
        time.sleep(0.1)
try:
    T2 = 2048
    kv2 = tuple((k[:T2].contiguous(), v[:T2].contiguous()) for k, v in kv)
    toks2 = toks[:T2]
    raw2 = 2 * L * T2 * H * D * 2
    for fast in (True, False):
        meta = LMCacheEngineMetadata("lmsys/longchat-7b-16k", 1, 0 if fast else 1, "vllm", "bfloat16")
        eng = LMCacheEngine(LMCacheEngineConfig(256, None, f"lm://127.0.0.1:{port}", "cachegen", False, False), meta)
        if not fast:
            eng.engine_.supports_kv_view = lambda: False
        st_best, _ = timeit(lambda: eng.store(toks2, kv2, skip_existing=False), n=3, warm=1)
        rt_best, _ = timeit(lambda: eng.retrieve(toks2), n=3, warm=1)
        out["engine_lm_cachegen_" + ("fast_path" if fast else "generic_path")] = {
            "store_GBps": round(raw2 / st_best / 1e9, 2), "retrieve_GBps": round(raw2 / rt_best / 1e9, 2), "raw_bytes": raw2,
            "note": "lm:// (Python-socket client, pinned-slab zero copy) + native server process on loopback; GB/s of raw KV"}
        if fast:
            from lmcache_b200.codec import KvView
            be = eng.engine_
            view = KvView.from_tuple(kv2, "vllm")
            t0 = time.perf_counter(); hs = sha256_prefix_chain(toks2, 256); t1 = time.perf_counter()
            for _ in range(2):
                ta = time.perf_counter()
                with be.serializer.view_to_pinned_batch(view, 256, 0, T2) as blobs:
                    tb = time.perf_counter()
                    nbytes = sum(len(b) for b in blobs)
                    for i, mv in enumerate(blobs):
                        be.connection.set(f"brk{i}", mv)
                    while not be.connection.exists(f"brk{len(blobs) - 1}"):
                        pass
                    tc = time.perf_counter()
            tg = time.perf_counter()
            got = [be.connection.get(f"brk{i}") for i in range(len(blobs))]
            th = time.perf_counter()
            out["engine_lm_cachegen_store_breakdown"] = {
                "hash_ms": round((t1 - t0) * 1e3, 2), "encode_to_pinned_ms": round((tb - ta) * 1e3, 2),
                "send_ms": round((tc - tb) * 1e3, 2), "wire_bytes": nbytes,
                "send_GBps_wire": round(nbytes / (tc - tb) / 1e9, 2), "get_ms": round((th - tg) * 1e3, 2),
                "get_GBps_wire": round(nbytes / (th - tg) / 1e9, 2)}
        eng.close()
finally:
    srv.terminate(); srv.wait()

print(json.dumps(out, indent=1))
