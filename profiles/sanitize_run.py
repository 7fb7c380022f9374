"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
all three container versions, the register-staged and the TMA-staged encoder, both decoder table layouts and both
version-3 header readers (short headers: peaked data; long ones: uniform data), chunks > 256 tokens (split kernels), paged
KV, the hash chain (cp.async ring) with several chains in one warp, pack / unpack through the engine's raw host tier.

    compute-sanitizer --tool memcheck python profiles/sanitize_run.py        (on the GPU box)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
import numpy as np  # noqa: E402
import torch  # noqa: E402

from lmcache_b200.cache_engine import sha256_prefix_chain  # noqa: E402
from lmcache_b200.codec import CacheGenCodec, KvView  # noqa: E402
from oracle import oracle as O  # noqa: E402

MODEL = "lmsys/longchat-7b-16k"
kb, vb = O.make_bins(MODEL)


def bits_to_t(bits):
    return torch.from_numpy(bits.view(np.int16).copy()).view(torch.bfloat16)


ok = True
for coder in ("rans_compact", "rans", "ac"):
    codec = CacheGenCodec(MODEL, coder=coder)
    for (L, H, D, T, cs, kind) in [(2, 2, 72, 300, 256, "peaked"), (2, 1, 128, 520, 512, "peaked"), (1, 3, 40, 77, 64, "peaked"),
                                   (4, 2, 128, 300, 256, "uniform")]:
        C = H * D
        if kind == "peaked":
            bits = O.synth_kv_bits(L, T, C, seed=T)
        else:
            bits = O.f32_to_bf16_bits(np.random.default_rng(T).uniform(-1, 1, size=(L, 2, T, C)).astype(np.float32))
        kv = bits_to_t(bits).reshape(L, 2, T, H, D).cuda()
        want = np.concatenate([O.decode_chunk(O.encode_chunk(bits[:, :, j * cs:min(T, (j + 1) * cs)], 0, kb, vb), 0, kb, vb, 0)
                               for j in range((T + cs - 1) // cs)], axis=2)
        for enc_path in ("legacy", "tma"):
            os.environ["B200KV_ENCODE_PATH"] = enc_path
            raws = codec.encode_to_host(KvView.from_blob(kv, "vllm"), 0, T, cs)
            for table in ("rows", "transposed"):
                os.environ["B200KV_DECODE_TABLE"] = table
                out = torch.zeros_like(kv)
                codec.decode(raws, KvView.from_blob(out, "vllm"), [j * cs for j in range(len(raws))])
                torch.cuda.synchronize()
                got = out.contiguous().cpu().view(torch.int16).numpy().view(np.uint16).reshape(L, 2, T, C)
                ok &= bool(np.array_equal(got, want))
                ok &= all(w == 0 for w in codec.decode_status())
        # damaged containers (lengths, stream headers / CDF rows, payload): must stay inside the buffer under memcheck
        from lmcache_b200.codec import container_layout_of, parse_header
        rng = np.random.default_rng(T)
        for raw in raws[:1]:
            lo = container_layout_of(parse_header(raw))
            for a, b in ((lo.off_lengths, lo.off_payload), (lo.off_payload, len(raw)), (lo.off_cdf, lo.off_maxes)):
                bad = bytearray(raw)
                for pos in rng.integers(a, b, size=256):
                    bad[pos] = int(rng.integers(0, 256))
                try:
                    codec.decode([bytes(bad)], KvView.from_blob(torch.zeros_like(kv[:, :, :parse_header(raw).ntokens]), "vllm"), [0])
                    torch.cuda.synchronize()
                except ValueError:
                    pass
        # paged
        slots = torch.randperm(T + 64)[:T].cuda()
        caches = []
        for l in range(L):
            k = torch.zeros((T + 64, H, D), dtype=torch.bfloat16, device="cuda")
            v = torch.zeros_like(k)
            k[slots] = kv[l, 0]
            v[slots] = kv[l, 1]
            caches.append((k, v))
        raws2 = codec.encode_to_host(KvView.from_paged(caches, slots), 0, T, cs)
        ok &= all(bytes(a)[64:] == bytes(b)[64:] for a, b in zip(raws, raws2))
        codec.decode(raws2, KvView.from_paged(caches, slots), [j * cs for j in range(len(raws2))])
        torch.cuda.synchronize()
os.environ.pop("B200KV_ENCODE_PATH", None)
os.environ.pop("B200KV_DECODE_TABLE", None)
toks = torch.randint(0, 32000, (1000,), dtype=torch.int64).cuda()
ok &= sha256_prefix_chain(toks, 256) == O.sha256_chain(toks.cpu().numpy(), 256)
# several chains in one launch (16 lanes of the warp share the cp.async ring)
lens = [300 + 17 * i for i in range(16)]
offs = [0]
for n in lens:
    offs.append(offs[-1] + n)
allt = torch.randint(0, 32000, (offs[-1],), dtype=torch.int64).cuda()
got = sha256_prefix_chain(allt, 256, offs)
want_h = []
for i in range(16):
    want_h += O.sha256_chain(allt[offs[i]:offs[i + 1]].cpu().numpy(), 256)
ok &= got == want_h
print("SANITIZE_RUN_OK" if ok else "SANITIZE_RUN_MISMATCH")
