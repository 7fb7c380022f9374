"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
import numpy as np, torch
from oracle import oracle as O
from lmcache_b200.codec import CacheGenCodec, KvView
from lmcache_b200.cache_engine import sha256_prefix_chain
MODEL = "lmsys/longchat-7b-16k"
codec = CacheGenCodec(MODEL)
def bits_to_t(bits): return torch.from_numpy(bits.view(np.int16).copy()).view(torch.bfloat16)
ok = True
for (L, H, D, T, cs) in [(2, 2, 72, 300, 256), (2, 1, 128, 520, 512), (1, 3, 40, 77, 64)]:
    C = H * D
    bits = O.synth_kv_bits(L, T, C, seed=T)
    kv = bits_to_t(bits).reshape(L, 2, T, H, D).cuda()
    raws = codec.encode_to_host(KvView.from_blob(kv, "vllm"), 0, T, cs)
    out = torch.zeros_like(kv)
    codec.decode(raws, KvView.from_blob(out, "vllm"), [j * cs for j in range(len(raws))])
    torch.cuda.synchronize()
    kb, vb = O.make_bins(MODEL)
    want = np.concatenate([O.decode_chunk(O.encode_chunk(bits[:, :, j*cs:min(T,(j+1)*cs)], 0, kb, vb), 0, kb, vb, 0) for j in range(len(raws))], axis=2)
    got = out.contiguous().cpu().view(torch.int16).numpy().view(np.uint16).reshape(L, 2, T, C)
    ok &= bool(np.array_equal(got, want))
    # paged
    slots = torch.randperm(T + 64)[:T].cuda()
    caches = []
    for l in range(L):
        k = torch.zeros((T + 64, H, D), dtype=torch.bfloat16, device="cuda"); v = torch.zeros_like(k)
        k[slots] = kv[l, 0]; v[slots] = kv[l, 1]; caches.append((k, v))
    raws2 = codec.encode_to_host(KvView.from_paged(caches, slots), 0, T, cs)
    ok &= all(bytes(a)[64:] == bytes(b)[64:] for a, b in zip(raws, raws2))
    codec.decode(raws2, KvView.from_paged(caches, slots), [j * cs for j in range(len(raws2))])
    torch.cuda.synchronize()
toks = torch.randint(0, 32000, (1000,), dtype=torch.int64).cuda()
ok &= sha256_prefix_chain(toks, 256) == O.sha256_chain(toks.cpu().numpy(), 256)
print("SANITIZE_RUN_OK" if ok else "SANITIZE_RUN_MISMATCH")
