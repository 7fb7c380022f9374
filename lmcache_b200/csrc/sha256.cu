// sha256.cu -- chunked token-id prefix hash on the GPU.
//
// Replaces LMCacheEngine._hash / _chunk_tokens / _prefix_hash (lmcache/cache_engine.py:58-96):
//     h_i = sha256( ascii_hex(h_{i-1}) || bytes(tokens[i*cs:(i+1)*cs]) ).hexdigest(),   h_{-1} = ""
// which costs the reference one tokens.cpu() device->host sync per chunk.  Here the whole chain runs in one
// launch on the caller's stream and only the 32-byte digests cross to the host (one copy, no per-chunk sync).
//
// A SHA-256 chain is serial by construction (Merkle-Damgard: every 64-byte block needs the previous
// state, and chunk i's first block is the previous digest), so one chain is latency-bound.  The kernel
// therefore splits the work per chain across one warp: all 32 lanes expand message schedules
// (W[16..63] + K, the state-independent ~45% of the work) for 32 blocks at a time into shared memory, then
// lane 0 runs the 64 serial rounds per block.  Independent sequences (n_seq chains: batched requests,
// RAG chunk mixes) run on different warps/SMs concurrently.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace b200kv {

__constant__ uint32_t kK[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

__device__ __forceinline__ uint32_t rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }

// Message byte `i` of the chunk message = prefix (64 hex chars, or empty for the first chunk) || token bytes ||
// 0x80 || zeros || 64-bit big-endian bit length.
struct ChunkMsg {
    const uint8_t* tok;     // token bytes of this chunk
    uint32_t ntok_bytes;
    uint32_t plen;          // 0 or 64
    const uint8_t* prefix;  // 64 hex chars in shared memory
    uint32_t total;         // plen + ntok_bytes
    uint32_t nblocks;       // padded block count
    __device__ __forceinline__ uint32_t byte_at(uint32_t i) const {
        if (i < plen) return prefix[i];
        if (i < total) return tok[i - plen];
        if (i == total) return 0x80u;
        const uint32_t end = nblocks * 64u;
        if (i >= end - 8u) {
            const uint64_t bits = (uint64_t)total * 8ull;
            return (uint32_t)(bits >> (8u * (end - 1u - i))) & 0xffu;
        }
        return 0u;
    }
};

constexpr int kWarpsPerCta = 4;

__global__ void __launch_bounds__(32 * kWarpsPerCta) sha256_chain_kernel(const uint8_t* tokens, int elem_size,
                                                                         const int64_t* seq_offsets, int n_seq,
                                                                         int chunk_size, uint8_t* digests) {
    __shared__ uint32_t s_wk[kWarpsPerCta][64][32];   // (W + K)[round][block] for 32 blocks per warp (32 KiB total)
    __shared__ uint8_t s_prefix[kWarpsPerCta][64];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int seq = blockIdx.x * kWarpsPerCta + wid;
    if (seq >= n_seq) return;
    const int64_t t0 = seq_offsets[seq], t1 = seq_offsets[seq + 1];
    // digest slot of this sequence = sum of chunk counts of the previous ones
    int64_t slot = 0;
    for (int s = 0; s < seq; ++s) slot += (seq_offsets[s + 1] - seq_offsets[s] + chunk_size - 1) / chunk_size;
    uint32_t (*wk)[32] = s_wk[wid];
    uint8_t* prefix = s_prefix[wid];

    bool first = true;
    for (int64_t tb = t0; tb < t1; tb += chunk_size, ++slot) {
        const int64_t cnt = (t1 - tb) < chunk_size ? (t1 - tb) : chunk_size;
        ChunkMsg msg;
        msg.tok = tokens + tb * elem_size;
        msg.ntok_bytes = (uint32_t)(cnt * elem_size);
        msg.plen = first ? 0u : 64u;
        msg.prefix = prefix;
        msg.total = msg.plen + msg.ntok_bytes;
        msg.nblocks = (msg.total + 9u + 63u) / 64u;
        const bool word_ok = (reinterpret_cast<uintptr_t>(msg.tok) & 3u) == 0u;

        uint32_t st[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au,
                          0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
        for (uint32_t b0 = 0; b0 < msg.nblocks; b0 += 32u) {
            const uint32_t nb = min(32u, msg.nblocks - b0);
            // ---- parallel: lane L expands block b0 + L
            if ((uint32_t)lane < nb) {
                uint32_t w[64];
                const uint32_t base = (b0 + lane) * 64u;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const uint32_t o = base + 4u * i;
                    if (word_ok && o >= msg.plen && o + 4u <= msg.total)   // whole word inside the token bytes
                        w[i] = __byte_perm(__ldg(reinterpret_cast<const uint32_t*>(msg.tok + (o - msg.plen))), 0u, 0x0123);
                    else
                        w[i] = (msg.byte_at(o) << 24) | (msg.byte_at(o + 1) << 16) | (msg.byte_at(o + 2) << 8) |
                               msg.byte_at(o + 3);
                }
#pragma unroll
                for (int i = 16; i < 64; ++i) {
                    const uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
                    const uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
                    w[i] = w[i - 16] + s0 + w[i - 7] + s1;
                }
#pragma unroll
                for (int i = 0; i < 64; ++i) wk[i][lane] = w[i] + kK[i];   // conflict-free: lanes -> banks
            }
            __syncwarp();
            // ---- serial: lane 0 runs the compression rounds
            if (lane == 0) {
                for (uint32_t b = 0; b < nb; ++b) {
                    uint32_t a = st[0], bb = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll 16
                    for (int i = 0; i < 64; ++i) {
                        const uint32_t t1_ = h + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + wk[i][b];
                        const uint32_t t2_ = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & bb) ^ (a & c) ^ (bb & c));
                        h = g; g = f; f = e; e = d + t1_; d = c; c = bb; bb = a; a = t1_ + t2_;
                    }
                    st[0] += a; st[1] += bb; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
                }
            }
            __syncwarp();
        }
        // ---- publish digest; hex of it is the next chunk's prefix
        if (lane == 0) {
            uint8_t* out = digests + slot * 32;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t byte = (st[i] >> (24 - 8 * k)) & 0xffu;
                    out[4 * i + k] = (uint8_t)byte;
                    const uint32_t hi = byte >> 4, lo = byte & 15u;
                    prefix[8 * i + 2 * k] = (uint8_t)(hi < 10u ? '0' + hi : 'a' + hi - 10u);
                    prefix[8 * i + 2 * k + 1] = (uint8_t)(lo < 10u ? '0' + lo : 'a' + lo - 10u);
                }
            }
        }
        __syncwarp();
        first = false;
    }
}

}  // namespace b200kv

using namespace b200kv;

extern "C" int b200kv_sha256_chain(const void* tokens, int32_t elem_size, const int64_t* seq_offsets, int32_t n_seq,
                                   int32_t chunk_size, void* digests, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    B2_REQUIRE(seq_offsets != nullptr && n_seq > 0, "bad sequence table");
    B2_REQUIRE(elem_size == 1 || elem_size == 2 || elem_size == 4 || elem_size == 8, "elem_size must be 1/2/4/8");
    B2_REQUIRE(chunk_size > 0, "chunk_size must be positive");
    B2_REQUIRE((int64_t)chunk_size * elem_size < (1ll << 28), "chunk too large");
    int64_t nchunks = 0;
    for (int s = 0; s < n_seq; ++s) {
        B2_REQUIRE(seq_offsets[s + 1] >= seq_offsets[s], "seq_offsets must be non-decreasing");
        nchunks += (seq_offsets[s + 1] - seq_offsets[s] + chunk_size - 1) / chunk_size;
    }
    if (nchunks == 0) return 0;
    B2_REQUIRE(tokens != nullptr && digests != nullptr, "NULL tokens / digests");
    // sequence table -> device (small; staged by the driver before return)
    int64_t* d_off = nullptr;
    B2_CHECK_CUDA(cudaMallocAsync(&d_off, sizeof(int64_t) * (size_t)(n_seq + 1), stream));
    cudaError_t e = cudaMemcpyAsync(d_off, seq_offsets, sizeof(int64_t) * (size_t)(n_seq + 1), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) {
        const int blocks = (n_seq + kWarpsPerCta - 1) / kWarpsPerCta;
        sha256_chain_kernel<<<blocks, 32 * kWarpsPerCta, 0, stream>>>(static_cast<const uint8_t*>(tokens), elem_size, d_off,
                                                                        n_seq, chunk_size, static_cast<uint8_t*>(digests));
        e = cudaGetLastError();
    }
    cudaFreeAsync(d_off, stream);
    B2_CHECK_CUDA(e);
    return 0;
}
