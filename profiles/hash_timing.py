"""Device time of the token hash chain (b200kv_sha256_chain through the engine's helper): 8192 / 65536 tokens in one chain,
16 chains x 4096 tokens in one launch.  Usage: python profiles/hash_timing.py   (on the GPU box)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from lmcache_b200.cache_engine import sha256_prefix_chain  # noqa: E402


def main():
    out = {}
    for n in (8192, 65536):
        t = torch.randint(0, 32000, (n,), device="cuda")
        for _ in range(3):
            sha256_prefix_chain(t, 256)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            sha256_prefix_chain(t, 256)
        torch.cuda.synchronize()
        out[f"one_chain_{n}_tokens_ms"] = round((time.perf_counter() - t0) / 10 * 1e3, 3)
    print(os.environ.get("B200KV_SHA_ADDS", "alu(default)"), out)


if __name__ == "__main__":
    main()
