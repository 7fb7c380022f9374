// codec.cu -- CacheGen encode / decode kernels for sm_100a and their C-ABI entry points.
//
// Replaces (reference paths relative to the LMCache v0.1.2 tree):
//   encode: lmcache/storage_backend/serde/cachegen_encoder.py:266-325 (encode_function) incl. the three
//           torchac_cuda calls, collect_bytes and the pickle container (cachegen_basics.py:131-136)
//   decode: lmcache/storage_backend/serde/cachegen_decoder.py:52-106,143-202
//
// Three container versions share these kernels (include/b200kv.h): 1 = arithmetic coder, 2 = rANS, both with the
// reference's CDF tensor as a section; 3 (default) = rANS streams that carry their own symbol histogram, from which the
// decoder rebuilds the CDF (ac_core.cuh: stream header).
//
// Thread mapping: one entropy-coder stream = one (plane nl, channel c) = one thread; a CTA owns a
// tile of CT consecutive channels of one plane (and one <=256-token group).  Global KV reads/writes are
// then naturally coalesced along the channel dimension (a warp touches 64 contiguous bytes per token)
// and the tile's byte streams are contiguous in the container.  The coder writes each stream to a temp
// row in global memory; stream compaction (collect_bytes in the reference) is a separate scan + gather
// pass (enc_scan_kernel, compact_kernel) that stages a tile's byte range in shared memory and writes it
// with 16-byte stores, so no CTA ever waits on another one.  DESIGN.md section 3 has the per-kernel story.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "ac_core.cuh"
#include "common.cuh"

namespace b200kv {

constexpr int CT = 128;            // streams (threads) per tile
constexpr int SPW = 6;             // 5-bit symbols packed per 32-bit word in shared memory
constexpr int SYMW = 43;           // words per symbol row: ceil(256 / 6); odd -> conflict-free columns
constexpr int PAIRW = 33;          // split mode: words per CDF-pair row (32 symbols + cdf[32]); odd
constexpr int TEMPW_FUSED = 40;    // words per stream in the tile's temp rows: own-CDF streams of <= 256 symbols
                                   //   cost <= 256*log2(31) + 2 bits = 159 bytes
constexpr int TEMPW_SPLIT = 132;   // foreign CDF (chunk > 256 tokens): <= 16 bits / symbol + termination; 16-byte rows
constexpr int TEMPW_FUSED_RANS = 48;   // rANS, own-CDF streams: <= 95 renormalisation halfwords for ANY symbol order
                                       //   (ideal <= 1268.5 bits, < 1 bit of overshoot per step); the 32-bit
                                       //   final state goes to its own array.  Split mode: <= 1 halfword per symbol = 128 words
constexpr int CODER_AC = 0, CODER_RANS = 1;   // payload coder; B2KV container version = coder + 1 ...
constexpr int CODER_RANS_COMPACT = 2;         // ... 3 = rANS payload + compact side information (ac_core.cuh, make_layout)
constexpr int kHdrRowWords = 8;               // version 3: header words 2..8 of a long stream header, in front of the rANS part
constexpr int TEMPW_FUSED_RANS_HDR = TEMPW_FUSED_RANS + kHdrRowWords;

struct EncParams {
    PlaneTable pt;
    int64_t sT, sH, tok_begin;
    const int64_t* slot_map;         // paged KV: token i of the call lives in row slot_map[i] of every plane; NULL = row i
    int32_t L, H, D, C, dtype;
    int32_t n_chunks, chunk_tokens, last_chunk_tokens, tpp;   // tpp = tiles per plane
    int32_t tiles_full, tempw;                                 // tiles per full chunk; words per temp row
    int32_t stage_bytes;                                       // compact_kernel: bytes of shared-memory stage per CTA
    int32_t coder;                                             // CODER_AC | CODER_RANS (the payload coder)
    int32_t compact;                                           // 1 = container version 3 (histogram in the stream, u8 half-lengths)
    uint8_t* out;
    int64_t out_stride;
    uint64_t* sizes_out;
    uint32_t* temp;                  // [n_tiles][CT][tempw] coder output before compaction
    uint32_t* rstate;                // rANS: [n_tiles][CT] final coder states (the first 4 bytes of every stream); version 3:
                                     //   [n_tiles][CT] records of 4 words {header word 0, header word 1, state, header bytes}
    uint32_t* tile_tot;              // [n_chunks][tiles_full] bytes per tile, then exclusive prefix (in place)
    unsigned long long* totals;      // [n_chunks] payload bytes
    unsigned int* err;               // [n_chunks]
};

// section offsets of a container of this call (encode and decode parameter blocks alike)
template <class Prm>
__device__ __forceinline__ Layout layout_of(const Prm& P, int t) {
    return make_layout(P.L, P.C, t, P.compact);
}

// stream lengths section: int32 bytes (versions 1, 2) or u8 bytes / 2 (version 3: header + rANS stream, even, <= 230 bytes)
__device__ __forceinline__ uint32_t load_len(const uint8_t* sec, int64_t idx, bool compact) {
    return compact ? 2u * (uint32_t)sec[idx] : (uint32_t)reinterpret_cast<const int32_t*>(sec)[idx];
}
__device__ __forceinline__ void store_len(uint8_t* sec, int64_t idx, uint32_t len, bool compact) {
    if (compact) sec[idx] = (uint8_t)(len >> 1);
    else reinterpret_cast<int32_t*>(sec)[idx] = (int32_t)len;
}

// Version 3: build the stream's header (ac_core.cuh) -- which symbols occur and how often.  The bytes are assembled in
// a register, a word at a time: words 0 and 1 (mask + the first counts: all there is for streams with few symbols) are
// returned and go to the tile's side array (one coalesced 16-byte record per stream, next to the rANS state), words 2..8
// -- streams with many symbols only -- go to the front of the stream's temp row.  Returns the header length (even,
// <= kHdrMax).  cnt[i] is 0 for i >= nb by construction.
__device__ __forceinline__ uint32_t build_stream_header(const uint32_t (&cnt)[32], uint32_t mask, uint32_t wany, int nb,
                                                        uint32_t* rowfront, uint32_t& w0, uint32_t& w1) {
    // mask: bit i <=> cnt[i] != 0; wany: the OR of the masks of the warp's lanes (a symbol nobody uses costs one test)
    const uint32_t top = 0x80000000u >> __clz((int)mask);        // the last set bit: its count is implied
    const uint32_t st = mask & ~top;                             // symbols whose count is stored
    const uint32_t mb = (uint32_t)hdr_mask_bytes(nb);
    w0 = 0u;
    w1 = 0u;
    uint32_t widx = 0u, acc = mask, sh = 8u * mb;                // sh = 8 * bytes held in acc
    auto flush = [&]() {
        if (widx == 0u) w0 = acc;
        else if (widx == 1u) w1 = acc;
        else rowfront[widx - 2u] = acc;
        ++widx;
        acc = 0u;
        sh = 0u;
    };
    if (sh == 32u) flush();
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        if (i < nb && ((wany >> i) & 1u) && ((st >> i) & 1u)) {
            acc |= cnt[i] << sh;
            sh += 8u;
            if (sh == 32u) flush();
        }
    }
    uint32_t hlen = mb + (uint32_t)__popc(st);
    hlen += hlen & 1u;
    if (sh != 0u) flush();
    return hlen;
}

// the mask of a stream header at p (2-byte aligned), restricted to the plane's nb symbols
__device__ __forceinline__ uint32_t read_header_mask(const uint8_t* p, int nb) {
    uint32_t m = *reinterpret_cast<const uint16_t*>(p);
    if (nb > 16) m |= (uint32_t)*reinterpret_cast<const uint16_t*>(p + 2) << 16;
    if (nb <= 8) m &= 0xffu;
    return nb >= 32 ? m : m & ((1u << nb) - 1u);
}

__device__ __forceinline__ int chunk_tokens_of(const EncParams& P, int j) {
    return j == P.n_chunks - 1 ? P.last_chunk_tokens : P.chunk_tokens;
}

// Row of a token inside its plane.  PAGED: the caller's slot mapping (vLLM's paged KV cache: row = block * block_size
// + offset); every lane of a warp asks for the same token, so the lookup is one broadcast load that hits L1.
template <bool PAGED>
__device__ __forceinline__ int64_t tok_row(const int64_t* slot_map, int64_t tok) {
    if constexpr (PAGED) return __ldg(slot_map + tok);
    else return tok;
}

// ------------------------------------------------------------------------------------------ absmax
// max1 = amax(|x|, channels) per (plane, token), kept in the input half dtype
// (cachegen_encoder.py:54-55).  |x| ordering == integer ordering of (bits & 0x7fff); a NaN in the row
// wins (pattern above inf), like torch.amax.  One warp per row, 128-bit loads when alignment allows.
// Round 2 tried to hide this kernel -- the only HBM-bound one of the path -- under the instruction-bound coding kernels of
// the previous wave (second stream, 128-thread blocks with <= 32 registers so that a block fits next to them): no gain
// (8.31 -> 8.28 ms per step).  Both coder kernels fill the SM's shared memory with their own CTAs (7 x 32.5 KB, 12 x 18.6
// KB incl. the 1 KB the system reserves per CTA), so not even a block without shared memory finds room; the kernels only
// overlap at their tails.  Measured, rejected.
template <bool VEC, bool PAGED>
__global__ void __launch_bounds__(256) absmax_kernel(EncParams P, int64_t total_tokens) {
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t nrows = (int64_t)2 * P.L * total_tokens;
    if (warp >= nrows) return;
    const int nl = (int)(warp / total_tokens);
    const int64_t T = warp % total_tokens;
    const uint16_t* row = P.pt.p[nl] + tok_row<PAGED>(P.slot_map, P.tok_begin + T) * P.sT;
    uint32_t m = 0;
    if (VEC) {
        const int vec_per_head = P.D >> 3;
        const int nvec = P.H * vec_per_head;
        for (int v = lane; v < nvec; v += 32) {
            const int h = v / vec_per_head, dv = v - h * vec_per_head;
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(row + (int64_t)h * P.sH + dv * 8));
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t a = w[k] & 0x7fff7fffu;
                m = max(m, max(a & 0xffffu, a >> 16));
            }
        }
    } else {
        for (int c = lane; c < P.C; c += 32) {
            const int h = c / P.D, d = c - h * P.D;
            m = max(m, (uint32_t)(__ldg(row + (int64_t)h * P.sH + d) & 0x7fffu));
        }
    }
    m = __reduce_max_sync(0xffffffffu, m);
    if (lane == 0) {
        const int j = (int)(T / P.chunk_tokens);
        const int tj = chunk_tokens_of(P, j);
        const Layout lo = layout_of(P, tj);
        uint16_t* maxes = reinterpret_cast<uint16_t*>(P.out + (int64_t)j * P.out_stride + lo.off_maxes);
        maxes[(int64_t)nl * tj + (T - (int64_t)j * P.chunk_tokens)] = (uint16_t)m;
    }
}

// ------------------------------------------------------------------------------------------ helpers
// block-wide exclusive scan of one uint32 per thread (CT threads); returns exclusive prefix, total in *total
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* s_warp, uint32_t* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < CT / 32; ++w) {
        const uint32_t s = s_warp[w];
        if (w < wid) base += s;
        tot += s;
    }
    *total = tot;
    return base + inc - v;
}

// tile -> (chunk j, group g, plane nl, channel tile ct); tiles of a chunk are ordered (g, nl, ct), which is the
// order of the streams in the container payload.  Returns false for tiles beyond the (ragged) last chunk.
struct TileId { int j, g, nl, ct, t, tok0, gt, tile_in_chunk; };
__device__ __forceinline__ bool decode_tile(const EncParams& P, uint32_t tile, TileId* id) {
    const uint32_t per_group = 2u * P.L * P.tpp;
    const uint32_t j = tile / (uint32_t)P.tiles_full;
    const uint32_t rem = tile - j * (uint32_t)P.tiles_full;
    id->j = (int)j;
    id->tile_in_chunk = (int)rem;
    id->t = chunk_tokens_of(P, (int)j);
    id->g = (int)(rem / per_group);
    if (id->g * kGroup >= id->t) return false;
    const uint32_t rem2 = rem - (uint32_t)id->g * per_group;
    id->nl = (int)(rem2 / P.tpp);
    id->ct = (int)(rem2 - (uint32_t)id->nl * P.tpp);
    id->tok0 = id->g * kGroup;
    id->gt = min(kGroup, id->t - id->tok0);
    return true;
}

// Walk one stream's gt tokens in register double-buffered batches of BT loads (the loads of batch b+1 are in flight
// while batch b is consumed) and hand every quantised symbol to `f(symbol)`, in token order.  Used where symbols are
// consumed on the fly (chunk-wide histogram, coding against a chunk-wide CDF).
template <int DT, bool PAGED, int BT, class F>
__device__ __forceinline__ void for_each_symbol(const uint16_t* cbase, const int64_t* slot_map, int64_t tokabs, int64_t s1,
                                                int gt, const float* fac, float maxq, F&& f) {
    const int nbatch = (gt + BT - 1) / BT;
    uint16_t xa[BT], xb[BT];
    auto load = [&](uint16_t (&x)[BT], int b) {
        const int tk = b * BT;
#pragma unroll
        for (int k = 0; k < BT; ++k)
            x[k] = tk + k < gt ? __ldg(cbase + tok_row<PAGED>(slot_map, tokabs + tk + k) * s1) : (uint16_t)0;
    };
    auto use = [&](const uint16_t (&x)[BT], int b) {
        const int tk = b * BT;
        uint32_t q[BT];
#pragma unroll
        for (int k = 0; k < BT; ++k) q[k] = quant_symbol(half_to_float(x[k], DT), fac[min(tk + k, kGroup - 1)], maxq);
#pragma unroll
        for (int k = 0; k < BT; ++k)
            if (tk + k < gt) f(q[k]);
    };
    load(xa, 0);
    for (int b = 0; b < nbatch; b += 2) {
        if (b + 1 < nbatch) load(xb, b + 1);
        use(xa, b);
        if (b + 1 < nbatch) {
            if (b + 2 < nbatch) load(xa, b + 2);
            use(xb, b + 1);
        }
    }
}

// Same walk from the last token to the first (rANS codes a stream back to front so that the decoder runs forwards).
template <int DT, bool PAGED, int BT, class F>
__device__ __forceinline__ void for_each_symbol_rev(const uint16_t* cbase, const int64_t* slot_map, int64_t tokabs, int64_t s1,
                                                    int gt, const float* fac, float maxq, F&& f) {
    const int nbatch = (gt + BT - 1) / BT;
    uint16_t xa[BT], xb[BT];
    auto load = [&](uint16_t (&x)[BT], int b) {
        const int tk = b * BT;
#pragma unroll
        for (int k = 0; k < BT; ++k)
            x[k] = tk + k < gt ? __ldg(cbase + tok_row<PAGED>(slot_map, tokabs + tk + k) * s1) : (uint16_t)0;
    };
    auto use = [&](const uint16_t (&x)[BT], int b) {
        const int tk = b * BT;
        uint32_t q[BT];
#pragma unroll
        for (int k = 0; k < BT; ++k) q[k] = quant_symbol(half_to_float(x[k], DT), fac[min(tk + k, kGroup - 1)], maxq);
#pragma unroll
        for (int k = BT - 1; k >= 0; --k)
            if (tk + k < gt) f(q[k]);
    };
    load(xa, nbatch - 1);
    for (int b = nbatch - 1; b >= 0; b -= 2) {
        if (b >= 1) load(xb, b - 1);
        use(xa, b);
        if (b >= 1) {
            if (b >= 2) load(xa, b - 2);
            use(xb, b - 1);
        }
    }
}

// rANS encoder step as the kernels run it (same arithmetic as rans_enc_symbol in ac_core.cuh, laid out for issue slots):
//   * the state shrinks by one halfword when it must; halfword number k (counted downwards: nk = -k) lands at
//     row_end + 2 * nk - 2, so the row's tail holds the halfwords in decode order.  The address is one unpredicated
//     IMAD.WIDE, only the 16-bit store and the count are predicated;
//   * q = x / f by one reciprocal biased low (estimate is q or q - 1) and one fix-up;
//   * x' = (q << 16) + (x - q f) + start  ==  x + start + q * (65536 - f).
__device__ __forceinline__ void rans_put(uint32_t& x, int32_t& nk, const uint16_t* row_end, uint32_t start, uint32_t freq) {
    const uint32_t xh = x >> 16;
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 ad;\n\t"
        "setp.ge.u32 p, %2, %3;\n\t"
        "mad.wide.s32 ad, %1, 2, %4;\n\t"
        "@p st.global.u16 [ad+-2], %0;\n\t"
        "@p add.s32 %1, %1, -1;\n\t"
        "selp.b32 %0, %2, %0, p;\n\t"
        "}"
        : "+r"(x), "+r"(nk) : "r"(xh), "r"(freq), "l"(row_end) : "memory");
    float rc;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(__uint2float_rn(freq)));
    const uint32_t q = __float2uint_rz(__uint2float_rz(x) * (rc * 0.99999952316284179688f));   // floor(x / f) or one less
    // With m = 65536 - f:  a = q m + x;  x - q f = a - (q << 16);  x' = a + start, plus m when the estimate was one short.
    // Six integer instructions (m, a, r, compare, add, predicated add) where "negate f, r, compare, q + 1, select, x +
    // start, m, multiply-add" took eight (ncu, round 2).
    const uint32_t m = 65536u - freq;
    const uint32_t a = q * m + x;
    const uint32_t r = a - (q << 16);
    x = a + start;
    if (r >= freq) x += m;
}

// ------------------------------------------------------------------------------------------ encode
// One tile = CT consecutive channels of one plane and one <= 256-token group; one thread = one coder stream.
// FUSED (chunk <= 256 tokens): quantise -> 5-bit symbols in shared memory + thread-private histogram -> CDF (the
//   33-entry rows are contiguous in smem and in the container: one coalesced copy) -> arithmetic coding.  KV is read
//   from HBM exactly once here (plus once by absmax).
// !FUSED (chunk > 256 tokens): the chunk-wide CDF was produced by cdf_kernel; one tile codes one group, quantising
//   on the fly.
// Coder output goes to the tile's temp rows in global memory (sparse 32-bit stores, merged in L2); stream lengths go to
// the container; the tile's byte total goes to tile_tot.  Compaction into the contiguous payload (collect_bytes in the
// reference) is done by scan_kernel + compact_kernel afterwards, so no CTA ever waits on another one.
template <bool FUSED, int DT, bool PAGED, int CODER>
__global__ void __launch_bounds__(CT, FUSED ? 7 : 4) encode_kernel(EncParams P) {
    extern __shared__ __align__(16) uint32_t smem[];
    // FUSED : symbol rows u32[CT][SYMW] | cdf rows u16[CT][33] (also the histogram) | fac[256] (later fl32(n/t)[257])
    // !FUSED: pair rows u32[CT][33]                                                 | fac[256]
    constexpr int ROWS_W = FUSED ? CT * SYMW : 0;
    constexpr int TAB_W = FUSED ? (CT * kLp * 2 + 3) / 4 : CT * PAIRW;
    uint32_t* rows = smem;
    uint32_t* tab = smem + ROWS_W;
    float* fac = reinterpret_cast<float*>(smem + ((ROWS_W + TAB_W + 3) & ~3));
    __shared__ uint32_t s_warp[CT / 32];

    const int tid = threadIdx.x;
    TileId id;
    if (!decode_tile(P, blockIdx.x, &id)) return;
    const int NL = 2 * P.L;
    const int j = id.j, nl = id.nl, ct = id.ct, t = id.t, gt = id.gt;
    const int c = ct * CT + tid;
    const bool active = c < P.C;
    const int ncols = min(CT, P.C - ct * CT);

    uint8_t* cont = P.out + (int64_t)j * P.out_stride;
    const Layout lo = layout_of(P, t);
    const uint16_t* maxes = reinterpret_cast<const uint16_t*>(cont + lo.off_maxes) + (int64_t)nl * t + id.tok0;
    const float maxq = P.pt.maxq[nl];
    const int64_t s1 = P.sT;
    // FUSED: "safe" factors (an infinite factor becomes NaN: same symbols, and pass 1 can skip the range check)
    for (int i = tid; i < gt; i += CT)
        fac[i] = FUSED ? quant_factor_safe(maxq, half_to_float(maxes[i], DT)) : quant_factor(maxq, half_to_float(maxes[i], DT));
    if (FUSED) for (int i = gt + tid; i < kGroup + 8; i += CT) fac[i] = 0.0f;      // padded slots of the last batch

    const int h = active ? c / P.D : 0;
    // first token of this tile in the call's token numbering; cbase = this stream's channel in row 0 of the plane
    const int64_t tokabs = P.tok_begin + (int64_t)j * P.chunk_tokens + id.tok0;
    const uint16_t* cbase = P.pt.p[nl] + (int64_t)h * P.sH + (active ? c - h * P.D : 0);
    const uint16_t* src = cbase + (PAGED ? 0 : tokabs * P.sT);          // !PAGED: token tk is at src + tk * sT
    uint32_t* trow = P.temp + ((int64_t)blockIdx.x * CT + tid) * P.tempw;
    uint32_t cap = (uint32_t)P.tempw;
    // keep the row pointer and the capacity as plain register values: otherwise every flush re-derives the address
    // from (tile, tid, tempw, base) with six extra instructions
    asm volatile("" : "+l"(trow), "+r"(cap));
    uint32_t len = 0u, hlen = 0u;

    if (FUSED) {
        uint32_t* myrow = rows + tid * SYMW;
        uint16_t* cdfr = reinterpret_cast<uint16_t*>(tab);
        uint16_t* crow = cdfr + tid * kLp;
        uint16_t* hist = crow;
#pragma unroll
        for (int i = 0; i < kLp; ++i) crow[i] = 0;
        // pull the whole tile (gt token rows x 256 B) into L2 up front: one prefetch per 128-byte line, spread over
        // the CTA, so pass 1's loads pay L2 latency instead of a DRAM round trip per batch
        if (!PAGED) {
            const int lines_per_tok = (ncols * 2 + 127) >> 7;
            const uint16_t* tile0 = P.pt.p[nl] + (P.tok_begin + (int64_t)j * P.chunk_tokens + id.tok0) * P.sT +
                                    (int64_t)((ct * CT) / P.D) * P.sH + ((ct * CT) % P.D);
            if ((ct * CT) / P.D == (ct * CT + ncols - 1) / P.D)    // tile lies within one head row: contiguous per token
                for (int i = tid; i < gt * lines_per_tok; i += CT) {
                    const int tok = i / lines_per_tok, ln = i - tok * lines_per_tok;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(tile0 + (int64_t)tok * s1 + ln * 64));
                }
        }
        __syncthreads();   // fac ready
        if (active) {
            // ---- pass 1: batches of 12 tokens (two packed words), register double-buffered: the loads of batch b+1
            // are in flight while batch b is quantised, so a warp never sits out a full L2 round trip per batch.
            // Slots past the group's end (last word / last batch) read as x = 0 with factor 0 and are quantised like
            // the rest; their known symbol is taken out of the histogram afterwards, and pass 2 never reads them.
            constexpr int NB = 2, BT = NB * SPW;
            const int nbatch = (gt + BT - 1) / BT;
            uint16_t xa[BT], xb[BT];
            const uint32_t s1u = (uint32_t)s1;
            const uint32_t two = 2u * (uint32_t)min(P.n_chunks, 1);      // 2, but opaque to the compiler
            auto load = [&](uint16_t (&x)[BT], int b) {
                const int tk = b * BT;
                if constexpr (PAGED) {
                    const int64_t* sm = P.slot_map + tokabs + tk;
#pragma unroll
                    for (int k = 0; k < BT; ++k) x[k] = tk + k < gt ? __ldg(cbase + __ldg(sm + k) * s1) : (uint16_t)0;
                } else {
                    // token tk + k is one row further than token tk + k - 1: ONE IMAD.WIDE.U32 per address (row
                    // pitch x 2, the 2 from a register ptxas cannot fold, added to the previous address) instead of the
                    // 64-bit add chains the compiler builds from the pointer arithmetic (4.6 -> 2 instructions per load,
                    // ncu round 2).  The chain of twelve is off the critical path: the batch is a prefetch.
                    const uint16_t* q = src + (int64_t)tk * s1;
                    if (tk + BT <= gt) {
#pragma unroll
                        for (int k = 0; k < BT; ++k) {
                            x[k] = __ldg(q);
                            asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(q) : "r"(s1u), "r"(two));
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < BT; ++k) {
                            x[k] = tk + k < gt ? __ldg(q) : (uint16_t)0;
                            asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(q) : "r"(s1u), "r"(two));
                        }
                    }
                }
            };
            auto quantise = [&](const uint16_t (&x)[BT], int b) {
                const int tk = b * BT;
                // all symbols of the batch first (independent chains), then the histogram read-modify-writes: those
                // may alias each other, so interleaving them with the arithmetic would serialise the whole batch
                uint32_t q[BT];
#pragma unroll
                for (int k = 0; k < BT; ++k) q[k] = quant_symbol_nc(half_to_float(x[k], DT), fac[tk + k], maxq);
#pragma unroll
                for (int half = 0; half < NB; ++half) {
                    if (tk + half * SPW < gt) {
                        uint32_t word = 0u;
#pragma unroll
                        for (int k = 0; k < SPW; ++k) {
                            hist[q[half * SPW + k]] += 1;               // symbols are <= 30 by construction
                            word |= q[half * SPW + k] << (5 * k);
                        }
                        myrow[b * NB + half] = word;
                    }
                }
            };
            load(xa, 0);
            for (int b = 0; b < nbatch; b += 2) {
                if (b + 1 < nbatch) load(xb, b + 1);
                quantise(xa, b);
                if (b + 1 < nbatch) {
                    if (b + 2 < nbatch) load(xa, b + 2);
                    quantise(xb, b + 1);
                }
            }
            const int pad = (gt + SPW - 1) / SPW * SPW - gt;            // padded slots in the last word
            if (pad) hist[quant_symbol(0.0f, 0.0f, maxq)] -= (uint16_t)pad;
        }
        __syncthreads();   // every thread is done with fac: reuse it as the fl32(n / t) table
        {
            const float tf = (float)t;
            for (int n = tid; n <= t; n += CT) fac[n] = fdiv((float)n, tf);
        }
        __syncthreads();
        if (active) {
            // ---- CDF from the thread's own histogram, written over it (counts -> registers first)
            uint32_t cnt[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) cnt[i] = hist[i];
            // symbols no lane of the warp uses are skipped (their cdf entries follow from the previous one)
            uint32_t mask = 0u;
#pragma unroll
            for (int i = 0; i < 32; ++i) mask |= (cnt[i] != 0u ? 1u : 0u) << i;
            const uint32_t wany = __reduce_or_sync(__activemask(), mask);
            CdfAccum2 acc;
            acc.init();
#pragma unroll
            for (uint32_t i = 0; i < 32u; ++i) {
                crow[i] = (uint16_t)acc.value(i);
                if ((wany >> i) & 1u) acc.absorb(fac[cnt[i]]);
            }
            crow[32] = (uint16_t)acc.value(32u);
            // version 3 keeps the histogram instead of the CDF row (the CDF is a function of it): the stream's header
            if (P.compact) {
                uint32_t w0, w1;
                hlen = build_stream_header(cnt, mask, wany, 2 * ((int)maxq + 1), trow, w0, w1);
                reinterpret_cast<uint2*>(P.rstate)[((int64_t)blockIdx.x * CT + tid) * 2] = make_uint2(w0, w1);
            }
        }
        __syncthreads();
        if (!P.compact) {   // the tile's 33-entry rows are contiguous in smem and in the container: straight coalesced copy
            uint16_t* dstc = reinterpret_cast<uint16_t*>(cont + lo.off_cdf) + ((int64_t)nl * P.C + ct * CT) * kLp;
            for (int e = tid; e < ncols * kLp; e += CT) dstc[e] = cdfr[e];
        }
        // ---- pass 2: entropy-code the stream
        if (active && CODER == CODER_RANS) {
            // rANS: last token first; halfwords land at a descending pointer, so the row's tail is the stream in
            // decode order.  Row capacity (96 halfwords) cannot be exceeded: no clamp, no flag.
            uint32_t x = kRansLow;
            const uint16_t* const wend = reinterpret_cast<const uint16_t*>(trow) + 2 * P.tempw;     // the row's end
            int32_t nk = 0;                                                // minus the number of halfwords pushed
            const char* const cb = reinterpret_cast<const char*>(crow);
            // symbol s -> byte offset 2 s of its CDF entry: ((word >> 5 k) & 31) * 2 as one shift + one mask
            auto code = [&](uint32_t word, int k) {
                const uint32_t o = (k == 0 ? word << 1 : word >> (5 * k - 1)) & 62u;     // s <= 30: entry s + 1 is real
                const uint32_t c_lo = *reinterpret_cast<const uint16_t*>(cb + o);
                const uint32_t c_hi = *reinterpret_cast<const uint16_t*>(cb + o + 2);
                rans_put(x, nk, wend, c_lo, c_hi - c_lo);
            };
            int w = (gt - 1) / SPW;
            {
                const uint32_t word = myrow[w];
                for (int k = gt - 1 - w * SPW; k >= 0; --k) code(word, k);
            }
            for (--w; w >= 0; --w) {
                const uint32_t word = myrow[w];
#pragma unroll
                for (int k = SPW - 1; k >= 0; --k) code(word, k);
            }
            if (P.compact) reinterpret_cast<uint2*>(P.rstate)[((int64_t)blockIdx.x * CT + tid) * 2 + 1] = make_uint2(x, hlen);
            else P.rstate[(int64_t)blockIdx.x * CT + tid] = x;
            len = hlen + 4u - 2u * (uint32_t)nk;
        } else if (active) {
            EncState2 st;
            st.init();
            int tk = 0;
            for (int w = 0; tk < gt; ++w) {
                const uint32_t word = myrow[w];
                if (tk + SPW <= gt) {
#pragma unroll
                    for (int k = 0; k < SPW; ++k) {
                        const uint32_t sidx = (word >> (5 * k)) & 31u;       // <= 30: crow[sidx + 1] is a real entry
                        const uint32_t c_lo = crow[sidx];
                        enc_symbol2(st, c_lo, (uint32_t)crow[sidx + 1u] - c_lo, trow, cap);
                    }
                    tk += SPW;
                } else {
                    for (int k = 0; tk < gt; ++k, ++tk) {
                        const uint32_t sidx = (word >> (5 * k)) & 31u;
                        const uint32_t c_lo = crow[sidx];
                        enc_symbol2(st, c_lo, (uint32_t)crow[sidx + 1u] - c_lo, trow, cap);
                    }
                }
            }
            len = enc_finish2(st, trow, cap);
            if (st.w > cap) atomicOr(&P.err[j], 1u);   // size bound violated (cannot happen; stores were clamped)
        }
    } else {
        // CDF of the whole chunk was written by cdf_kernel: load it and build the (c_lo | width << 16) rows
        uint32_t* prow = tab + tid * PAIRW;
        const uint16_t* cdf_src =
            reinterpret_cast<const uint16_t*>(cont + lo.off_cdf) + ((int64_t)nl * P.C + ct * CT) * kLp;
        for (int e = tid; e < ncols * kLp; e += CT) tab[e] = cdf_src[e];
        __syncthreads();   // also: fac ready
        if (active) {
            uint32_t cv[kLp];
#pragma unroll
            for (int i = 0; i < kLp; ++i) cv[i] = prow[i] & 0xffffu;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const uint32_t hi = (i == 31) ? 0x10000u : cv[i + 1];
                prow[i] = cv[i] | ((hi - cv[i]) << 16);
            }
            if (CODER == CODER_RANS) {
                // at most one halfword per symbol: the 264-halfword row cannot overflow
                uint32_t x = kRansLow;
                const uint16_t* const wend = reinterpret_cast<const uint16_t*>(trow) + 2 * TEMPW_SPLIT;
                int32_t nk = 0;
                for_each_symbol_rev<DT, PAGED, 4>(cbase, P.slot_map, tokabs, s1, gt, fac, maxq, [&](uint32_t q) {
                    const uint32_t pr = prow[q];
                    rans_put(x, nk, wend, pr & 0xffffu, pr >> 16);
                });
                P.rstate[(int64_t)blockIdx.x * CT + tid] = x;
                len = 4u - 2u * (uint32_t)nk;
            } else {
                EncState2 st;
                st.init();
                for_each_symbol<DT, PAGED, 4>(cbase, P.slot_map, tokabs, s1, gt, fac, maxq, [&](uint32_t q) {
                    const uint32_t pr = prow[q];
                    enc_symbol2(st, pr & 0xffffu, pr >> 16, trow, cap);
                });
                len = enc_finish2(st, trow, cap);
                if (st.w > cap) atomicOr(&P.err[j], 1u);
            }
        }
    }

    // ---- stream lengths to the container, tile total for the compaction scan
    if (active) store_len(cont + lo.off_lengths, ((int64_t)id.g * NL + nl) * P.C + c, len, P.compact != 0);
    uint32_t tile_total;
    (void)block_excl_scan(len, s_warp, &tile_total);
    if (tid == 0) P.tile_tot[(int64_t)j * P.tiles_full + id.tile_in_chunk] = tile_total;
}

// ------------------------------------------------------------------------------------------ encode, TMA-staged
// The fused rANS encoder again, rebuilt around three facts the profiles of the kernel above showed (profiles/r2b_*):
// it is bound by instruction issue (85 % at 0.6 bits/symbol) and, at high entropy, by shared-memory bank conflicts (9.9
// extra wavefronts per warp-symbol at 4.1 bits); its per-thread LDG.U16 loads with their address arithmetic cost 4.4
// instructions per symbol; and packing symbols into shared-memory rows only to unpack them again costs another 4.
//   * Input tiles arrive through the TMA unit: `cp.async.bulk` copies of the tile's token rows (128 channels = 256
//     contiguous bytes each; any row pitch, paged slots included) into a 3-stage ring of 16-token boxes, completion on
//     mbarriers.  A symbol's load is then one LDS.U16 at an immediate offset.
//   * No symbol rows: pass 2 streams the tile through the ring a second time (last box first: rANS codes a stream back
//     to front) and re-quantises -- 4 instructions, fewer than store + load + unpack, and 22 KB of shared memory less.
//     The second read is served by L2 or DRAM; the kernel sits at a fifth of the HBM roofline, the bytes are there.
//   * One TRANSPOSED table [32 entries][128 streams] of 32-bit words serves as histogram (pass 1: one red.shared.add
//     per symbol) and then as the coder's (start | freq << 16) table (pass 2: one LDS per symbol): a lane never leaves
//     its own bank, so neither pass has bank conflicts at any entropy; the columns are thread-private, so the passes
//     need no CTA-wide barrier between them.
// Eligibility (host side): rANS, chunk <= 256 tokens, tiles of exactly 128 channels that are contiguous in every
// token row, 16-byte aligned rows.  Everything else takes encode_kernel above.
constexpr int BOXT = 16;                 // tokens per ring box: 16 x 256 B = 4 KB
constexpr int RSTAGES = 3;
constexpr int kEncTmaSmem = 32 * CT * 4 + (kGroup + 16) * 4 + (kGroup + 8) * 4 + RSTAGES * BOXT * CT * 2 + 64;

__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t a) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tLAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra LAB_DONE;\n\tbra LAB_WAIT;\n\tLAB_DONE:\n\t}" ::"r"(a), "r"(parity), "r"(20000u) : "memory");
}
// one row of a box: global -> shared through the TMA unit, completion (bytes) on the stage's mbarrier
__device__ __forceinline__ void bulk_row_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

template <int DT, bool PAGED>
__global__ void __launch_bounds__(CT, 7) encode_tma_kernel(EncParams P) {
    extern __shared__ __align__(128) uint8_t smem8[];
    uint32_t* tbl = reinterpret_cast<uint32_t*>(smem8);                                   // [32][CT]: counts, then (start | freq << 16)
    float* fac = reinterpret_cast<float*>(smem8 + 32 * CT * 4);                          // [kGroup + 16]
    float* ntab = fac + (kGroup + 16);                                                    // fl32(n / t), n = 0..t
    uint8_t* ring = smem8 + 32 * CT * 4 + (kGroup + 16) * 4 + (kGroup + 8) * 4;          // RSTAGES boxes of BOXT x 256 B
    const uint32_t ring_a = (uint32_t)__cvta_generic_to_shared(ring);
    const uint32_t bar_a = ring_a + RSTAGES * BOXT * CT * 2;                              // full[RSTAGES], empty[RSTAGES]
    __shared__ uint32_t s_warp[CT / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    TileId id;
    if (!decode_tile(P, blockIdx.x, &id)) return;
    const int NL = 2 * P.L;
    const int j = id.j, nl = id.nl, ct = id.ct, t = id.t, gt = id.gt;
    const int c = ct * CT + tid;
    uint8_t* cont = P.out + (int64_t)j * P.out_stride;
    const Layout lo = layout_of(P, t);
    const uint16_t* maxes = reinterpret_cast<const uint16_t*>(cont + lo.off_maxes) + (int64_t)nl * t + id.tok0;
    const float maxq = P.pt.maxq[nl];
    const int64_t tokabs = P.tok_begin + (int64_t)j * P.chunk_tokens + id.tok0;
    // channel 0 of this tile in row 0 of the plane; the tile's 128 channels are contiguous in every row (host checked)
    const uint16_t* tile0 = P.pt.p[nl] + (int64_t)((ct * CT) / P.D) * P.sH + ((ct * CT) % P.D);
    const int NB = (gt + BOXT - 1) / BOXT;               // boxes per pass
    const int NQ = 2 * NB;                               // pass 1 forwards, pass 2 backwards

    // ---- prologue: factors, n/t table, zeroed counters, barriers
    for (int i = tid; i < kGroup + 16; i += CT)
        fac[i] = i < gt ? quant_factor_safe(maxq, half_to_float(maxes[i], DT)) : 0.0f;
    {
        const float tf = (float)t;
        for (int n = tid; n <= t; n += CT) ntab[n] = fdiv((float)n, tf);
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) tbl[i * CT + tid] = 0u;
    if (tid == 0) {
        for (int s = 0; s < RSTAGES; ++s) {
            mbar_init(bar_a + 8 * s, 1);                          // full: the producer's expect_tx arrival (+ the bytes)
            mbar_init(bar_a + 8 * (RSTAGES + s), CT / 32);        // empty: one arrival per warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // box q of the 2 NB box sequence: pass 1 walks boxes 0..NB-1, pass 2 walks NB-1..0
    auto issue = [&](int q) {                                      // warp 0, all lanes
        const int b = q < NB ? q : NQ - 1 - q;
        const int rows = min(BOXT, gt - b * BOXT);
        const int stage = q % RSTAGES;
        const uint32_t full = bar_a + 8 * stage;
        if (lane == 0) mbar_expect_tx(full, (uint32_t)rows * CT * 2);
        __syncwarp();
        if (lane < rows) {
            const int64_t row = tok_row<PAGED>(P.slot_map, tokabs + b * BOXT + lane);
            bulk_row_g2s(ring_a + (uint32_t)(stage * BOXT + lane) * CT * 2, tile0 + row * P.sT, CT * 2, full);
        }
    };
    if (warp == 0)
        for (int q = 0; q < min(RSTAGES, NQ); ++q) issue(q);

    uint32_t* const mycol = tbl + tid;
    const uint32_t mycol_a = (uint32_t)__cvta_generic_to_shared(mycol);
    auto consume_done = [&](int q) {                               // every thread, after its last read of box q
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_a + 8 * (RSTAGES + q % RSTAGES));
        if (warp == 0 && q + RSTAGES < NQ) {                       // the stage is free once all four warps have arrived
            mbar_wait(bar_a + 8 * (RSTAGES + q % RSTAGES), (uint32_t)(q / RSTAGES) & 1u);
            issue(q + RSTAGES);
        }
    };
    auto load_box = [&](int q, uint16_t (&x)[BOXT]) {
        const int stage = q % RSTAGES;
        mbar_wait(bar_a + 8 * stage, (uint32_t)(q / RSTAGES) & 1u);
        const uint16_t* colp = reinterpret_cast<const uint16_t*>(ring) + stage * (BOXT * CT) + tid;
#pragma unroll
        for (int k = 0; k < BOXT; ++k) x[k] = colp[k * CT];        // LDS.U16 at immediate offsets
    };

    // ---- pass 1: histogram (one shared-memory reduction per symbol, conflict-free by layout)
    auto count = [&](uint16_t xv, float f) {
        const uint32_t sym = quant_symbol_nc(half_to_float(xv, DT), f, maxq);
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(mycol_a + sym * (CT * 4)) : "memory");
    };
    for (int q = 0; q < NB; ++q) {
        uint16_t x[BOXT];
        load_box(q, x);
        const int tk0 = q * BOXT;
        const int rows = min(BOXT, gt - tk0);
        if (rows == BOXT) {                                         // all boxes but (at most) the last: no per-token test
#pragma unroll
            for (int k = 0; k < BOXT; ++k) count(x[k], fac[tk0 + k]);
        } else {
#pragma unroll
            for (int k = 0; k < BOXT; ++k)
                if (k < rows) count(x[k], fac[tk0 + k]);
        }
        consume_done(q);
    }

    // ---- CDF from the thread's own column, written back as (start | freq << 16) over the counts
    uint32_t c32, hlen = 0u;
    uint32_t* trow = P.temp + ((int64_t)blockIdx.x * CT + tid) * P.tempw;
    {
        uint32_t cnt[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) cnt[i] = mycol[i * CT];
        uint32_t mask = 0u;
#pragma unroll
        for (int i = 0; i < 32; ++i) mask |= (cnt[i] != 0u ? 1u : 0u) << i;
        // this kernel is picked for high-entropy data, where every symbol is in use somewhere in the warp: no skipping
        // of unused symbols here (the tests cost more than they save: 5.23 -> 5.48 ms at 4.1 bits per symbol, measured)
        const uint32_t wany = 0xffffffffu;
        if (P.compact) {
            uint32_t w0, w1;
            hlen = build_stream_header(cnt, mask, wany, 2 * ((int)maxq + 1), trow, w0, w1);
            reinterpret_cast<uint2*>(P.rstate)[((int64_t)blockIdx.x * CT + tid) * 2] = make_uint2(w0, w1);
        }
        CdfAccum2 acc;
        acc.init();
        uint32_t c0 = 0u;
#pragma unroll
        for (uint32_t i = 0; i < 31u; ++i) {
            if ((wany >> i) & 1u) acc.absorb(ntab[cnt[i]]);
            const uint32_t c1 = acc.value(i + 1u);
            mycol[i * CT] = c0 | ((c1 - c0) << 16);                 // symbols are <= 30: entry 31 is never coded
            c0 = c1;
        }
        if ((wany >> 31) & 1u) acc.absorb(ntab[cnt[31]]);
        c32 = acc.value(32u);
        mycol[31 * CT] = c0 | (c32 << 16);                          // keeps cdf[31] and cdf[32] for the container's CDF row
    }

    // ---- pass 2: rANS, last token first
    uint32_t x_state = kRansLow;
    const uint16_t* const wend = reinterpret_cast<const uint16_t*>(trow) + 2 * P.tempw;             // the row's end
    int32_t nk = 0;
    auto code = [&](uint16_t xv, float f) {
        // the same symbol as pass 1 without the XU pipe (F2I there, reciprocal + F2I in rans_put here): adding
        // 1.5 * 2^23 rounds v to an integer half-to-even exactly like F2I.RN and leaves it in the low mantissa bits;
        // fmaxf turns the NaN of a "safe factor" row into the bias itself, i.e. symbol 0, as F2I does
        const float v = fadd(fmul(half_to_float(xv, DT), f), maxq);
        const uint32_t sym = __float_as_uint(fmaxf(__fadd_rn(v, 12582912.0f), 12582912.0f)) & 31u;
        uint32_t pk;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(pk) : "r"(mycol_a + sym * (CT * 4)));
        rans_put(x_state, nk, wend, pk & 0xffffu, pk >> 16);
    };
    for (int q = NB; q < NQ; ++q) {
        uint16_t x[BOXT];
        load_box(q, x);
        const int b = NQ - 1 - q;
        const int tk0 = b * BOXT;
        const int rows = min(BOXT, gt - tk0);
        if (rows == BOXT) {
#pragma unroll
            for (int k = BOXT - 1; k >= 0; --k) code(x[k], fac[tk0 + k]);
        } else {
#pragma unroll
            for (int k = BOXT - 1; k >= 0; --k)
                if (k < rows) code(x[k], fac[tk0 + k]);
        }
        consume_done(q);
    }
    if (P.compact) reinterpret_cast<uint2*>(P.rstate)[((int64_t)blockIdx.x * CT + tid) * 2 + 1] = make_uint2(x_state, hlen);
    else P.rstate[(int64_t)blockIdx.x * CT + tid] = x_state;
    const uint32_t len = hlen + 4u - 2u * (uint32_t)nk;

    // ---- stream lengths, tile total, CDF rows (staged through the now idle ring: stream-major u16[33] rows,
    //      contiguous in the container -> one coalesced copy)
    store_len(cont + lo.off_lengths, ((int64_t)id.g * NL + nl) * P.C + c, len, P.compact != 0);
    uint32_t tile_total;
    (void)block_excl_scan(len, s_warp, &tile_total);               // has a __syncthreads: every thread is done with the ring
    if (tid == 0) P.tile_tot[(int64_t)j * P.tiles_full + id.tile_in_chunk] = tile_total;
    if (id.g == 0 && !P.compact) {                                  // the CDF belongs to the chunk; its first group writes it
        uint16_t* stg = reinterpret_cast<uint16_t*>(ring);
#pragma unroll
        for (int i = 0; i < 32; ++i) stg[tid * kLp + i] = (uint16_t)mycol[i * CT];
        stg[tid * kLp + 32] = (uint16_t)c32;
        __syncthreads();
        uint16_t* dstc = reinterpret_cast<uint16_t*>(cont + lo.off_cdf) + ((int64_t)nl * P.C + ct * CT) * kLp;
        for (int e = tid; e < CT * kLp; e += CT) dstc[e] = stg[e];
    }
}

// ------------------------------------------------------------------------------------------ cdf (chunks > 256 tokens)
template <int DT, bool PAGED>
__global__ void __launch_bounds__(CT) cdf_kernel(EncParams P) {
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t* cnts = smem;                                        // CT * PAIRW counters
    float* fac = reinterpret_cast<float*>(cnts + CT * PAIRW);     // kGroup
    const int tid = threadIdx.x;
    const int NL = 2 * P.L;
    const uint32_t per_chunk = (uint32_t)NL * P.tpp;
    const uint32_t j = blockIdx.x / per_chunk;
    const uint32_t rem = blockIdx.x - j * per_chunk;
    const int nl = (int)(rem / P.tpp);
    const int ct = (int)(rem - (uint32_t)nl * P.tpp);
    const int t = chunk_tokens_of(P, (int)j);
    const int c = ct * CT + tid;
    const bool active = c < P.C;
    const int ncols = min(CT, P.C - ct * CT);
    uint8_t* cont = P.out + (int64_t)j * P.out_stride;
    const Layout lo = layout_of(P, t);
    const uint16_t* maxes = reinterpret_cast<const uint16_t*>(cont + lo.off_maxes) + (int64_t)nl * t;
    const float maxq = P.pt.maxq[nl];
    uint32_t* prow = cnts + tid * PAIRW;
#pragma unroll
    for (int i = 0; i < PAIRW; ++i) prow[i] = 0u;
    const int h = active ? c / P.D : 0;
    const int64_t tokabs = P.tok_begin + (int64_t)j * P.chunk_tokens;
    const uint16_t* cbase = P.pt.p[nl] + (int64_t)h * P.sH + (active ? c - h * P.D : 0);
    for (int tok0 = 0; tok0 < t; tok0 += kGroup) {
        const int gt = min(kGroup, t - tok0);
        __syncthreads();
        for (int i = tid; i < gt; i += CT) fac[i] = quant_factor(maxq, half_to_float(maxes[tok0 + i], DT));
        __syncthreads();
        if (active)
            for_each_symbol<DT, PAGED, 12>(cbase, P.slot_map, tokabs + tok0, P.sT, gt, fac, maxq,
                                           [&](uint32_t q) { prow[q] += 1u; });
    }
    if (active) {   // counts -> CDF values, written back over the counter row
        uint32_t cnt[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) cnt[i] = prow[i];
        CdfAccum acc;
        acc.init(t);
#pragma unroll
        for (uint32_t i = 0; i < 32u; ++i) prow[i] = acc.next(i, cnt[i]);
        prow[32] = acc.next(32u, 0u);
    }
    __syncthreads();
    uint16_t* dstc = reinterpret_cast<uint16_t*>(cont + lo.off_cdf) + ((int64_t)nl * P.C + ct * CT) * kLp;
    for (int e = tid; e < ncols * kLp; e += CT) dstc[e] = (uint16_t)cnts[e];   // rows are PAIRW == kLp words: e maps 1:1
}

// ------------------------------------------------------------------------------------------ compaction
// exclusive prefix over a chunk's tile totals (in place) + the chunk's payload size; one CTA per chunk
__global__ void __launch_bounds__(1024) enc_scan_kernel(EncParams P) {
    __shared__ unsigned long long s_w[32];
    __shared__ unsigned long long s_carry;
    const int j = blockIdx.x;
    const int t = chunk_tokens_of(P, j);
    const int ntiles = ((t + kGroup - 1) / kGroup) * 2 * P.L * P.tpp;
    uint32_t* tb = P.tile_tot + (int64_t)j * P.tiles_full;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0ull;
    __syncthreads();
    for (int base = 0; base < ntiles; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned long long v = i < ntiles ? tb[i] : 0ull;
        unsigned long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        if (lane == 31) s_w[wid] = inc;
        __syncthreads();
        unsigned long long wbase = 0ull, tot = 0ull;
        for (int w = 0; w < 32; ++w) {
            const unsigned long long s = s_w[w];
            if (w < wid) wbase += s;
            tot += s;
        }
        const unsigned long long carry = s_carry;
        const unsigned long long excl = carry + wbase + inc - v;
        if (i < ntiles) {
            if (excl + v >= (1ull << 32)) atomicOr(&P.err[j], 8u);     // payload offsets are 32-bit per chunk
            tb[i] = (uint32_t)excl;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) P.totals[j] = s_carry;
}

// one stream: temp row (MSB-first native words, 16-byte aligned) -> len bytes at d.  16-byte loads, all of a row's loads
// (ten at a time for the long split-mode rows) in flight before the first use -- a word-at-a-time loop left one load in
// flight per thread and the kernel waiting on L2/DRAM latency.
__device__ __forceinline__ void copy_row(uint8_t* d, const uint32_t* srcw, uint32_t len, bool short_row) {
    auto put_word = [&](uint32_t w, uint32_t v) {
        const uint32_t nb = min(4u, len - 4u * w);
#pragma unroll
        for (uint32_t b = 0; b < 4u; ++b)
            if (b < nb) d[4u * w + b] = (uint8_t)(v >> (24u - 8u * b));
    };
    auto put_vec = [&](uint32_t qq, const uint4& v) {
        put_word(4u * qq, v.x);
        if (16u * qq + 4u < len) put_word(4u * qq + 1u, v.y);
        if (16u * qq + 8u < len) put_word(4u * qq + 2u, v.z);
        if (16u * qq + 12u < len) put_word(4u * qq + 3u, v.w);
    };
    constexpr int NV = TEMPW_FUSED / 4;
    static_assert(TEMPW_SPLIT % 4 == 0 && TEMPW_FUSED % 4 == 0, "temp rows must be 16-byte multiples");
    if (short_row) {                                       // fused mode: the whole row is <= NV vectors
        uint4 v[NV];
#pragma unroll
        for (int q = 0; q < NV; ++q)
            if (16u * q < len) v[q] = __ldg(reinterpret_cast<const uint4*>(srcw) + q);
#pragma unroll
        for (int q = 0; q < NV; ++q)
            if (16u * q < len) put_vec((uint32_t)q, v[q]);
    } else {                                               // split mode: rows up to 33 vectors, four in flight
        constexpr int NL = 4;
        for (uint32_t q0 = 0; 16u * q0 < len; q0 += NL) {
            uint4 v[NL];
#pragma unroll
            for (int q = 0; q < NL; ++q)
                if (16u * (q0 + q) < len) v[q] = __ldg(reinterpret_cast<const uint4*>(srcw) + q0 + q);
#pragma unroll
            for (int q = 0; q < NL; ++q)
                if (16u * (q0 + q) < len) put_vec(q0 + q, v[q]);
        }
    }
}

// rANS stream: the 32-bit final state (little-endian), then the tail of the temp row -- `nb` bytes of halfwords that the
// coder stored at descending addresses, already in decode order.  d is 2-byte aligned (every stream length is even).
// 16-byte loads, four in flight; typical streams (a few dozen bytes) take one round.
__device__ __forceinline__ void copy_row_rans(uint8_t* d, const uint32_t* row, uint32_t rowbytes, uint32_t nb, uint32_t state) {
    uint16_t* dh = reinterpret_cast<uint16_t*>(d);
    dh[0] = (uint16_t)state;
    dh[1] = (uint16_t)(state >> 16);
    const uint32_t h0 = (rowbytes - nb) >> 1;                  // first halfword of the stream inside the row
    const uint32_t hend = rowbytes >> 1;
    constexpr int NLV = 4;
    for (uint32_t q0 = h0 >> 3; 8u * q0 < hend; q0 += NLV) {
        uint4 v[NLV];
#pragma unroll
        for (int q = 0; q < NLV; ++q)
            if (8u * (q0 + q) < hend) v[q] = __ldg(reinterpret_cast<const uint4*>(row) + q0 + q);
#pragma unroll
        for (int q = 0; q < NLV; ++q) {
            if (8u * (q0 + q) < hend) {
                const uint32_t w[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
                for (uint32_t i = 0; i < 8u; ++i) {
                    const uint32_t hw = 8u * (q0 + q) + i;
                    if (hw >= h0) dh[2u + hw - h0] = (uint16_t)(w[i >> 1] >> (16u * (i & 1u)));
                }
            }
        }
    }
}

// move each tile's streams from its temp rows to their final, contiguous place in the payload
// (collect_bytes, cachegen_encoder.py:225-238).  Each thread copies its own stream into a shared-memory image of the
// tile's byte range (placed at the destination's 16-byte phase), then the CTA writes that range with 16-byte vector
// stores: the payload is written as full sectors no matter how short the individual streams are.
__global__ void __launch_bounds__(CT, 12) compact_kernel(EncParams P) {
    extern __shared__ __align__(16) uint8_t stage[];      // 16 + CT * tempw * 4 bytes
    __shared__ uint32_t s_warp[CT / 32];
    const int tid = threadIdx.x;
    TileId id;
    if (!decode_tile(P, blockIdx.x, &id)) return;
    const int NL = 2 * P.L;
    const int c = id.ct * CT + tid;
    uint8_t* cont = P.out + (int64_t)id.j * P.out_stride;
    const Layout lo = layout_of(P, id.t);
    const bool rans = P.coder == CODER_RANS;
    const uint32_t rowbytes = (uint32_t)P.tempw * 4u;
    // no stream is longer than its temp row (+ the state, + header words 0 and 1, which live in the side array)
    const uint32_t len = c < P.C ? min(load_len(cont + lo.off_lengths, ((int64_t)id.g * NL + id.nl) * P.C + c, P.compact != 0),
                                       rowbytes + (rans ? 4u : 0u) + (P.compact ? 8u : 0u)) : 0u;
    uint32_t tile_total;
    const uint32_t my_off = block_excl_scan(len, s_warp, &tile_total);
    const uint64_t base = P.tile_tot[(int64_t)id.j * P.tiles_full + id.tile_in_chunk];
    const int64_t room = P.out_stride - lo.off_payload;
    if ((int64_t)(base + tile_total) > room) {          // never write past the slot the caller gave us
        if (tid == 0) atomicOr(&P.err[id.j], 4u);
        return;
    }
    uint8_t* dst = cont + lo.off_payload + base;
    const uint32_t phase = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u);
    // the image of the tile's byte range must fit the shared-memory stage the launch provided; a tile coded against a
    // foreign CDF may (rarely) exceed it, then every thread writes its own stream straight to the payload
    const bool staged = phase + tile_total <= (uint32_t)P.stage_bytes;
    if (len && rans) {
        const uint32_t* srcw = P.temp + ((int64_t)blockIdx.x * CT + tid) * P.tempw;
        uint32_t state, hl = 0u;
        const uint32_t* tail = srcw;                               // the row's rANS part (halfwords right-aligned in it)
        uint32_t tailbytes = rowbytes;
        if (P.compact) {
            // version 3: header words 0, 1, the state and the header length come from the tile's side array (one coalesced
            // 16-byte load), header words 2..8 -- long headers only -- from the front of the row
            const uint4 rec = __ldg(reinterpret_cast<const uint4*>(P.rstate) + (int64_t)blockIdx.x * CT + tid);
            state = rec.z;
            hl = min(min(rec.w, (uint32_t)kHdrMax), max(len, 4u) - 4u);
            uint32_t hw[9] = {rec.x, rec.y, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            if (hl > 8u) {
                const uint4 a = __ldg(reinterpret_cast<const uint4*>(srcw)), b = __ldg(reinterpret_cast<const uint4*>(srcw) + 1);
                hw[2] = a.x; hw[3] = a.y; hw[4] = a.z; hw[5] = a.w; hw[6] = b.x; hw[7] = b.y; hw[8] = b.z;
            }
            uint16_t* dh = reinterpret_cast<uint16_t*>(staged ? stage + phase + my_off : dst + my_off);
#pragma unroll
            for (uint32_t k = 0; k < kHdrMax / 2; ++k)
                if (2u * k < hl) dh[k] = (uint16_t)(hw[k >> 1] >> (16u * (k & 1u)));
            tail = srcw + kHdrRowWords;
            tailbytes = rowbytes - 4u * kHdrRowWords;
        } else {
            state = P.rstate[(int64_t)blockIdx.x * CT + tid];
        }
        if (staged) copy_row_rans(stage + phase + my_off + hl, tail, tailbytes, max(len, 4u + hl) - 4u - hl, state);
        else copy_row_rans(dst + my_off + hl, tail, tailbytes, max(len, 4u + hl) - 4u - hl, state);
    } else if (len) {
        const uint32_t* srcw = P.temp + ((int64_t)blockIdx.x * CT + tid) * P.tempw;
        // two instantiations, so that the staged one compiles to shared-memory stores and not to generic ones
        if (staged) {
            copy_row(stage + phase + my_off, srcw, len, P.tempw == TEMPW_FUSED);
        } else {                                             // rare: plain word loop, keeps the kernel's registers low
            uint8_t* d = dst + my_off;
            for (uint32_t w = 0; 4u * w < len; ++w) {
                const uint32_t v = __ldg(srcw + w);
                for (uint32_t b = 0; b < 4u && 4u * w + b < len; ++b) d[4u * w + b] = (uint8_t)(v >> (24u - 8u * b));
            }
        }
    }
    if (!staged) return;                                     // uniform per CTA
    __syncthreads();
    // [phase, phase + tile_total) of `stage` -> dst - phase + same offsets; vector body, byte head / tail
    const uint32_t lo_b = phase, hi_b = phase + tile_total;
    const uint32_t body0 = min(hi_b, (lo_b + 15u) & ~15u), body1 = max(body0, hi_b & ~15u);
    uint8_t* dbase = dst - phase;
    for (uint32_t i = lo_b + tid; i < body0; i += CT) dbase[i] = stage[i];
    for (uint32_t i = body0 + 16u * tid; i < body1; i += 16u * CT)
        *reinterpret_cast<uint4*>(dbase + i) = *reinterpret_cast<const uint4*>(stage + i);
    for (uint32_t i = body1 + tid; i < hi_b; i += CT) dbase[i] = stage[i];
}

// ------------------------------------------------------------------------------------------ finalize
__global__ void finalize_kernel(EncParams P) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.n_chunks) return;
    const int t = chunk_tokens_of(P, j);
    const Layout lo = layout_of(P, t);
    b200kv_header* hd = reinterpret_cast<b200kv_header*>(P.out + (int64_t)j * P.out_stride);
    hd->magic = B200KV_MAGIC;
    hd->version = P.compact ? 3u : (uint32_t)P.coder + 1u;    // 1: arithmetic coder, 2: rANS, 3: rANS + compact sections
    hd->L = P.L; hd->H = P.H; hd->D = P.D;
    hd->ntokens = t;
    hd->ngroups = lo.ngroups;
    hd->max_dtype = P.dtype;
    hd->payload_bytes = P.totals[j];
    hd->total_bytes = (uint64_t)lo.off_payload + P.totals[j];
    hd->status = P.err[j];
    hd->reserved[0] = hd->reserved[1] = hd->reserved[2] = 0u;
    if (P.sizes_out) P.sizes_out[j] = hd->status ? 0ull : hd->total_bytes;     // 0 = this chunk failed (see header.status)
    if (P.compact) {                               // counts per stream of every plane: makes the container self-describing
        uint8_t* nbmap = reinterpret_cast<uint8_t*>(hd) + lo.off_cdf;
        const int NL = 2 * P.L;
        for (int nl = 0; nl < (int)align16(NL); ++nl) nbmap[nl] = nl < NL ? (uint8_t)(2 * ((int)P.pt.maxq[nl] + 1)) : (uint8_t)0;
    }
}

// ------------------------------------------------------------------------------------------ decode
struct DecChunk {
    const uint8_t* base;
    int64_t dst_tok;
    int32_t t, ngroups;
    uint32_t payload_bytes;      // from the (host-validated) header: stream offsets are clamped to it
    uint32_t pad;
};

struct DecParams {
    PlaneTable pt;               // destination planes; maxq = C_l = bins // 2 - 1
    int64_t sT, sH;
    const int64_t* slot_map;     // paged destination: token i lives in row slot_map[i]; NULL = row i
    int32_t L, H, D, C, out_dtype, max_dtype, n_chunks, tpp, tiles_max;
    int32_t compact;             // containers are version 3
    const DecChunk* chunks;      // device
    unsigned long long* tile_base;   // [n_chunks][tiles_max]: tile sums, then exclusive prefix
    uint32_t* status;            // [n_chunks] or NULL: bit 0 = a rANS stream did not return to its initial state,
                                 //   bit 1 = stream offsets beyond the payload (corrupt lengths section)
};

// tile sums of the stream lengths: one warp per tile
__global__ void __launch_bounds__(128) tile_sum_kernel(DecParams P) {
    const int j = blockIdx.y;
    const int tile = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const DecChunk dc = P.chunks[j];
    const int NL = 2 * P.L;
    const int ntiles = dc.ngroups * NL * P.tpp;
    if (tile >= ntiles) return;
    const int plane_row = tile / P.tpp;          // g * NL + nl
    const int ct = tile - plane_row * P.tpp;
    const Layout lo = layout_of(P, dc.t);
    const int c0 = ct * CT, c1 = min(P.C, c0 + CT);
    uint32_t s = 0;
    for (int c = c0 + lane; c < c1; c += 32) s += load_len(dc.base + lo.off_lengths, (int64_t)plane_row * P.C + c, P.compact != 0);
    s = __reduce_add_sync(0xffffffffu, s);
    if (lane == 0) P.tile_base[(int64_t)j * P.tiles_max + tile] = s;
}

// exclusive prefix over a chunk's tile sums (in place); one CTA per chunk
__global__ void __launch_bounds__(1024) tile_scan_kernel(DecParams P) {
    __shared__ unsigned long long s_w[32];
    __shared__ unsigned long long s_carry;
    const int j = blockIdx.x;
    const DecChunk dc = P.chunks[j];
    const int ntiles = dc.ngroups * 2 * P.L * P.tpp;
    unsigned long long* tb = P.tile_base + (int64_t)j * P.tiles_max;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0ull;
    __syncthreads();
    for (int base = 0; base < ntiles; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned long long v = i < ntiles ? tb[i] : 0ull;
        unsigned long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        if (lane == 31) s_w[wid] = inc;
        __syncthreads();
        unsigned long long wbase = 0ull, tot = 0ull;
        for (int w = 0; w < 32; ++w) {
            const unsigned long long s = s_w[w];
            if (w < wid) wbase += s;
            tot += s;
        }
        const unsigned long long carry = s_carry;
        if (i < ntiles) tb[i] = carry + wbase + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + tot;
        __syncthreads();
    }
}

// Aligned big-endian word reader over the stream's bytes in global memory with a one-word look-ahead: the
// load for word i+1 is issued when word i is consumed, so its L2/L1 latency overlaps ~8+ symbols of decoding.
// Each lane walks its own stream; a 32-byte sector serves 8 consecutive refills from L1.
struct WordSrc {
    const uint32_t* base;   // the container, as words
    uint32_t idx;        // next word to load: a 32-bit index keeps the refill to one IMAD.WIDE + one add
    uint32_t ahead;
    __device__ __forceinline__ void prime() { ahead = __ldg(base + idx); ++idx; }
    __device__ __forceinline__ uint32_t next_be() {
        const uint32_t w = ahead;
        ahead = __ldg(base + idx);
        ++idx;
        return __byte_perm(w, 0u, 0x0123);
    }
};
// Both readers may run past the end of their stream (a decoder consumes at most 18 bits (arithmetic coder) / one
// halfword (rANS) per symbol, whatever the bytes say): at most B200KV_READ_SLACK bytes past the stream's start, which
// b200kv_decode_chunks checks against the size of the caller's buffer.  A corrupt lengths section therefore cannot make
// a kernel read outside that buffer: stream starts are clamped to the payload, reads are bounded from there.

// rANS: aligned little-endian words off a running pointer; the look-ahead lives in the decoder state (RansDec::nxt)
struct LeWordSrc {
    const uint32_t* p;
    __device__ __forceinline__ uint32_t next_le() { return __ldg(p++); }
};

__device__ __forceinline__ uint16_t out_half(float v, int dt) {
    // hardware RNE converts (NaN payloads are canonicalised; every finite / inf value matches torch's cast)
    return dt ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
// typed store of the converted value (the 16-bit result goes straight from F2FP to STG.U16)
template <int OUT_DT>
__device__ __forceinline__ void store_half(uint16_t* p, float v) {
    if constexpr (OUT_DT) *reinterpret_cast<__half*>(p) = __float2half_rn(v);
    else *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16_rn(v);
}

// value = lut * row_max (one rounded multiply, cachegen_decoder.py:31-35), converted RNE and stored as 16 bits
// (F2FP + STG.U16, no register merge in between)
template <int OUT_DT>
__device__ __forceinline__ void store_dequant(uint16_t* base, uint32_t off, float lutv, float row_max, uint32_t two) {
    const float v = __fmul_rn(lutv, row_max);
    // address = base + 2 * off as ONE IMAD.WIDE (`two` = 2, opaque to ptxas, keeps it off the ALU pipe); the converted value stays in the low half of a 32-bit register
    // (F2FP.PACK_AB with a zero upper half) and STG.U16 stores that half: no 16-bit register shuffling
    if constexpr (OUT_DT)
        asm volatile("{\n\t.reg .b64 ad;\n\t.reg .b32 r;\n\t.reg .b16 lo, hi;\n\t"
                     "mad.wide.u32 ad, %1, %3, %0;\n\t"
                     "cvt.rn.f16x2.f32 r, 0f00000000, %2;\n\t"
                     "mov.b32 {lo, hi}, r;\n\t"
                     "st.global.b16 [ad], lo;\n\t}" ::"l"(base), "r"(off), "f"(v), "r"(two) : "memory");
    else
        asm volatile("{\n\t.reg .b64 ad;\n\t.reg .b32 r;\n\t.reg .b16 lo, hi;\n\t"
                     "mad.wide.u32 ad, %1, %3, %0;\n\t"
                     "cvt.rn.bf16x2.f32 r, 0f00000000, %2;\n\t"
                     "mov.b32 {lo, hi}, r;\n\t"
                     "st.global.b16 [ad], lo;\n\t}" ::"l"(base), "r"(off), "f"(v), "r"(two) : "memory");
}

// per-thread decode loop: one stream, gt symbols, straight to the destination layout.
// PAGED: dst is the stream's channel in row 0 of the plane and `slots` points at the group's first slot-map entry.
template <int OUT_DT, int NSTEPS, bool PAGED>
__device__ __forceinline__ void decode_stream(const uint8_t* cont, uint32_t my_off, const uint32_t* erow,
                                              const float* lut, const float* mx, uint16_t* dst, uint32_t sT, int gt,
                                              const int64_t* slots) {
    const uint32_t skip = my_off & 3u;            // containers are 16-byte aligned
    WordSrc src{reinterpret_cast<const uint32_t*>(cont), my_off >> 2, 0u};
    src.prime();
    DecState2 st;
    dec_init2(st, src, skip);
    uint16_t* d = dst;                                      // running pointer: one 64-bit add per token
    // dec_symbol2 returns 4 * symbol = the byte offset into the fp32 LUT
    auto lut_at = [&](uint32_t s4) { return *reinterpret_cast<const float*>(reinterpret_cast<const char*>(lut) + s4); };
    if constexpr (PAGED) {
        for (int i = 0; i < gt - 1; ++i)
            store_half<OUT_DT>(dst + __ldg(slots + i) * (int64_t)sT,
                               dequant_value(lut_at(dec_symbol2<NSTEPS>(st, src, erow, false)), mx[i]));
        store_half<OUT_DT>(dst + __ldg(slots + gt - 1) * (int64_t)sT,
                           dequant_value(lut_at(dec_symbol2<NSTEPS>(st, src, erow, true)), mx[gt - 1]));
    } else {
        for (int i = 0; i < gt - 1; ++i, d += sT)
            store_half<OUT_DT>(d, dequant_value(lut_at(dec_symbol2<NSTEPS>(st, src, erow, false)), mx[i]));
        store_half<OUT_DT>(d, dequant_value(lut_at(dec_symbol2<NSTEPS>(st, src, erow, true)), mx[gt - 1]));
    }
}

// one lower-bound step through a stream's table whose entries are PITCH bytes apart: a += ROWS entries iff the entry ROWS
// further is <= key.  The add is an IMAD with an opaque multiplier (FMA pipe; see rans_decode_stream).
template <uint32_t ROWS, uint32_t PITCH>
__device__ __forceinline__ void rans_search_step(uint32_t& a, uint32_t key, uint32_t one) {
    uint32_t ev;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(ev) : "r"(a), "n"(ROWS * PITCH));
    asm("{\n\t.reg .pred p;\n\tsetp.le.u32 p, %1, %2;\n\t@p mad.lo.u32 %0, %3, %4, %0;\n\t}"
        : "+r"(a) : "r"(ev), "r"(key), "r"(one), "n"(ROWS * PITCH));
}

// rANS decode loop (container version 2): one stream, gt symbols, straight to the destination layout.
// Per symbol: key = (x << 16) | 0xffff; lower-bound search over the stream's packed table pk[i] = (cdf[i] << 16) | freq(i)
// -- the two top levels sit in registers, the rest are LDS off a running shared-memory address --; the winning entry
// carries start and freq, so the state update is one multiply-add; at most one 16-bit renormalisation (a PRMT out of the
// two-word window, a predicated aligned load when the window moves on).  Returns the final state (2^16 when intact).
template <int OUT_DT, int NSTEPS, bool PAGED, bool TR>
__device__ __forceinline__ uint32_t rans_decode_stream(const uint8_t* cont, uint32_t my_off, const uint32_t* pk,
                                                       const float* lut, const float* mx, uint16_t* dst, uint32_t sT, int gt,
                                                       const int64_t* slots, uint32_t one) {
    // stream words are addressed as container base + 32-bit word index: the address of the next word is one IMAD.WIDE
    // (FMA pipe, not predicated), only the load and the index increment are predicated
    const uint32_t* const wbase = reinterpret_cast<const uint32_t*>(cont);
    uint32_t idx = my_off >> 2;
    struct Src {
        const uint32_t* b; uint32_t& i;
        __device__ __forceinline__ uint32_t next_le() { return __ldg(b + i++); }
    } src{wbase, idx};
    RansDec st;
    rans_dec_init(st, src, (my_off >> 1) & 1u);
    constexpr uint32_t H = 1u << (NSTEPS - 1);
    // `pk` points at this stream's entry 0.  Two table layouts (decode_kernel builds either):
    //   TR = false  rows of 33 words per stream (odd pitch: lanes that read the SAME entry never collide; lanes that read
    //               different entries sometimes do -- 2.7 extra wavefronts per warp-symbol at 0.6 bits/symbol, 7.4 at 4.1)
    //   TR = true   transposed, entry i of every stream in one 512-byte row: a lane never leaves its own bank, at the
    //               price of two more instructions per symbol for the LUT address and a two-step table build
    // The host picks by the containers' measured bits per symbol (b200kv_decode_chunks).
    constexpr uint32_t kPitch = TR ? CT * 4u : 4u;
    constexpr uint32_t kIdx = TR ? CT : 1u;
    const uint32_t a0 = (uint32_t)__cvta_generic_to_shared(pk);
    const uint32_t a0h = a0 + kPitch * H;
    const uint32_t r_mid = pk[kIdx * H], r_lo = pk[kIdx * (H / 2)], r_hi = pk[kIdx * (H + H / 2)];
    const uint32_t lut_a = (uint32_t)__cvta_generic_to_shared(lut);            // TR: lut[s] = lut_a + ((a - a0) >> 7)
    const uint32_t lut_rel = lut_a - a0;                                       // !TR: lut[s] = a + lut_rel
    uint32_t off = 0u;                                                         // element offset of the current token row
    // The integer ALU pipe (ISETP / SEL / PRMT / LOP3, one warp instruction per 2 cycles) is what bounds this loop, the
    // FMA pipe idles: additions and shifts are therefore written as IMADs whose multiplier ptxas cannot fold (`one` is
    // 1 but comes from a kernel parameter), which pins them to the FMA pipe.
    const uint32_t c64k = one << 16, mone = 0u - one, two = one + one, c25 = one << 25;
    auto step = [&](float row_max, int i) {
        uint32_t key, xh;
        asm("mad.lo.u32 %0, %1, %2, 65535;" : "=r"(key) : "r"(st.x), "r"(c64k));      // (x << 16) | 0xffff
        asm("mul.hi.u32 %0, %1, %2;" : "=r"(xh) : "r"(st.x), "r"(c64k));             // x >> 16
        const bool p1 = r_mid <= key;
        uint32_t a = p1 ? a0h : a0;
        const uint32_t m = p1 ? r_hi : r_lo;
        asm("{\n\t.reg .pred p;\n\tsetp.le.u32 p, %1, %2;\n\t@p mad.lo.u32 %0, %3, %4, %0;\n\t}"
            : "+r"(a) : "r"(m), "r"(key), "r"(one), "n"(kPitch * (H / 2)));
        if constexpr (H >= 16) rans_search_step<4, kPitch>(a, key, one);         // entries H/4 .. 1 further on
        rans_search_step<2, kPitch>(a, key, one);
        rans_search_step<1, kPitch>(a, key, one);
        uint32_t e, la;
        float lv;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(a));
        if constexpr (TR) asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(la) : "r"(a - a0), "r"(c25), "r"(lut_a));   // lut_a + 4 * symbol
        else la = a + lut_rel;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(lv) : "r"(la));
        // x = freq * (x >> 16) + slot - start
        uint32_t dl;
        asm("mul.hi.u32 %0, %1, %2;" : "=r"(dl) : "r"(key - e), "r"(c64k));
        st.x = (e & 0xffffu) * xh + dl;
        // renormalisation, branch-free (a warp takes this path on most symbols, so a branch would run for all lanes
        // anyway): p = x < 2^16 -> pull the next halfword out of the window; q = p and the window's upper half was
        // taken -> the window moves on (cur = nxt, nxt = next aligned word)
        const uint32_t* const wad = wbase + idx;
        asm volatile(
            "{\n\t.reg .pred p, q;\n\t"
            "setp.lt.u32 p, %0, 65536;\n\t"
            "setp.eq.and.u32 q, %3, 0x1076, p;\n\t"
            "@p prmt.b32 %0, %0, %1, %3;\n\t"
            "@p mad.lo.u32 %3, %3, %7, 0x20ca;\n\t"     // 0x1054 <-> 0x1076: sel = 0x20ca - sel
            "@q mov.b32 %1, %2;\n\t"
            "@q ld.global.nc.u32 %2, [%5];\n\t"
            "@q mad.lo.u32 %4, %6, %6, %4;\n\t"         // idx += 1
            "}"
            : "+r"(st.x), "+r"(st.cur), "+r"(st.nxt), "+r"(st.sel), "+r"(idx)
            : "l"(wad), "r"(one), "r"(mone));
        if constexpr (PAGED) {
            store_dequant<OUT_DT>(dst + __ldg(slots + i) * (int64_t)sT, 0u, lv, row_max, two);
        } else {
            store_dequant<OUT_DT>(dst, off, lv, row_max, two);
            asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(off) : "r"(one), "r"(sT));
        }
    };
    int i = 0;
#pragma unroll 1
    for (; i + 4 <= gt; i += 4) {                                               // 4 symbols per trip: the body is ~150
        const float4 m4 = *reinterpret_cast<const float4*>(mx + i);            // instructions, it must stay in the
        step(m4.x, i); step(m4.y, i + 1); step(m4.z, i + 2); step(m4.w, i + 3);   // instruction cache next to its twin
    }
#pragma unroll 1
    for (; i < gt; ++i) step(mx[i], i);
    return st.x;
}

// One tile = CT streams of one (chunk, group, plane).  Only the per-stream tables (33 words per stream, built from the
// 66-byte CDF rows which are contiguous in the container: one coalesced read), the row maxima and a 32-entry
// dequantisation LUT live in shared memory (~18 KB per CTA), so many CTAs stay resident and hide the serial latency of
// each stream's coder.  Symbols are dequantised and stored straight into the destination layout (no uint8 / fp32
// intermediates in HBM).  CODER selects the payload format (container version 1: arithmetic coder, 2: rANS).
template <int OUT_DT, bool PAGED, int CODER, bool TR>
__global__ void __launch_bounds__(CT, 12) decode_kernel(DecParams P) {
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t* tab = smem;                                                            // CT * 33 words (rows of 33, odd)
    float* mx = reinterpret_cast<float*>(smem + CT * kLp);                           // kGroup
    float* lut = mx + kGroup;                                                        // 32
    __shared__ uint32_t s_warp[CT / 32];

    const int tid = threadIdx.x;
    const int j = blockIdx.y;
    const DecChunk dc = P.chunks[j];
    const int NL = 2 * P.L;
    const int per_group = NL * P.tpp;
    const int tile = blockIdx.x;
    const int g = tile / per_group;
    if (g >= dc.ngroups) return;
    const int rem = tile - g * per_group;
    const int nl = rem / P.tpp;
    const int ct = rem - nl * P.tpp;
    const int tok0 = g * kGroup;
    const int gt = min(kGroup, dc.t - tok0);
    const int c = ct * CT + tid;
    const bool active = c < P.C;
    const int ncols = min(CT, P.C - ct * CT);
    const Layout lo = layout_of(P, dc.t);

    const uint32_t len = active ? load_len(dc.base + lo.off_lengths, ((int64_t)g * NL + nl) * P.C + c, P.compact != 0) : 0u;
    uint32_t tile_total;
    const uint32_t my_rel = block_excl_scan(len, s_warp, &tile_total);
    // the stream's byte offset inside the container; a corrupt lengths section cannot push it outside the payload
    const unsigned long long want = P.tile_base[(int64_t)j * P.tiles_max + tile] + my_rel;
    const bool beyond = want + len > (unsigned long long)dc.payload_bytes;
    const uint32_t my_off = (uint32_t)lo.off_payload + (uint32_t)min(want, (unsigned long long)dc.payload_bytes);

    // stage the per-stream tables (one contiguous run of ncols * 33 halfwords in the container), row maxima, LUT
    const uint16_t* cdf_src = reinterpret_cast<const uint16_t*>(dc.base + lo.off_cdf) + ((int64_t)nl * P.C + ct * CT) * kLp;
    const uint16_t* maxes = reinterpret_cast<const uint16_t*>(dc.base + lo.off_maxes) + (int64_t)nl * dc.t + tok0;
    const float cq = P.pt.maxq[nl];
    bool built = false;
    uint32_t hl = 0u;                        // version 3: bytes of stream header in front of the rANS state
    if constexpr (CODER == CODER_RANS) {
        if (P.compact) {
            // Container version 3: the stream starts with its symbol histogram; the CDF is a function of it (CdfAccum,
            // the same arithmetic as the encoder), so every thread rebuilds its own table -- row-major or transposed,
            // both conflict-free for thread-private writes.  fl32(n / t) comes from a table that borrows the row-maxima
            // area (mx[0..255] + lut[0] = 257 floats) until the table is built.
            float* pn = mx;
            const float tf = (float)dc.t;
            for (int n = tid; n <= kGroup; n += CT) pn[n] = fdiv((float)n, tf);
            __syncthreads();
            // the stream's header: mask of the symbols that occur, then their counts (the last one is implied).
            // Two readers, chosen per warp: when most lanes' headers are short (<= 8 bytes: few symbols per stream, the
            // streams themselves are short and neighbours share cache lines) each count is one byte load at a position
            // that depends on the mask alone -- branch-free; when a quarter of the lanes or more have long headers, the
            // header is pulled into registers with aligned word loads, only as many as it is long, and consumed a byte
            // at a time (scattered byte loads would cost a cache-line access each).  Measured, whole kernel: 3.29 / 4.64
            // ms (byte loads only) vs 3.40 / 3.95 ms (registers only) at 0.6 / 4.1 payload bits per symbol.
            const int nb = 2 * ((int)cq + 1);
            const uint32_t mbytes = (uint32_t)hdr_mask_bytes(nb);
            const uint8_t* sp = dc.base + my_off;
            const uint32_t al = (uint32_t)(reinterpret_cast<uintptr_t>(sp) & 3u);          // 0 or 2
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(sp - al);
            const uint32_t fs = 8u * al;
            uint32_t x_prev = __ldg(wp), x_next = __ldg(wp + 1);
            const uint32_t hb0 = __funnelshift_r(x_prev, x_next, fs);
            uint32_t mask = hb0;
            if (nb <= 8) mask &= 0xffu;
            else if (nb <= 16) mask &= 0xffffu;
            if (nb < 32) mask &= (1u << nb) - 1u;
            hl = hdr_len(mask, nb);
            const uint32_t top = 0x80000000u >> __clz((int)mask);               // the last set bit: its count is implied
            const bool all_short = __popc(__ballot_sync(0xffffffffu, active && hl > 8u)) < 8;
            const uint32_t wany = __reduce_or_sync(0xffffffffu, active ? mask : 0u);
            if (active) {
                uint32_t sum = 0u;
                uint32_t hb[9];
                uint32_t widx = mbytes >> 2, inw = mbytes & 3u, cur = 0u;        // register reader: word, byte in word
                const uint8_t* const cb = sp + mbytes;
                if (!all_short) {
                    hb[0] = hb0;
#pragma unroll
                    for (int k = 1; k < 9; ++k) {
                        hb[k] = 0u;
                        if ((uint32_t)(4 * k) < hl && (k < 5 || nb > 16)) {      // 16-symbol planes: <= 18 bytes
                            x_prev = x_next;
                            x_next = __ldg(wp + k + 1);
                            hb[k] = __funnelshift_r(x_prev, x_next, fs);
                        }
                    }
                    cur = (widx == 0u ? hb[0] : hb[1]) >> (8u * inw);
                }
                auto next_byte = [&]() -> uint32_t {
                    const uint32_t v = cur & 255u;
                    cur >>= 8;
                    if (++inw == 4u) {
                        inw = 0u;
                        ++widx;
                        cur = hb[1];
#pragma unroll
                        for (int k = 2; k < 9; ++k) cur = widx == (uint32_t)k ? hb[k] : cur;
                    }
                    return v;
                };
                auto count = [&](int i) -> uint32_t {                            // called once per i, ascending
                    if (i >= nb) return 0u;
                    const uint32_t bit = 1u << (i & 31);
                    uint32_t n = 0u;
                    if (all_short) {                                             // uniform per warp
                        const uint32_t below = i == 0 ? 0u : mask & (0xffffffffu >> (32 - (i & 31)));
                        const uint32_t v = __ldg(cb + __popc(below));
                        n = (mask & bit) ? v : 0u;
                        n = (top & bit) ? (uint32_t)dc.t - sum : n;
                        sum += n;
                    } else if (mask & bit) {
                        n = (top & bit) ? (uint32_t)dc.t - sum : next_byte();
                        sum += n;
                    }
                    return min(n, (uint32_t)kGroup);                             // a damaged header cannot index past pn[256]
                };
                CdfAccum2 acc;
                acc.init();
                uint32_t c0 = 0u;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (i < nb) {
                        if ((wany >> i) & 1u) acc.absorb(pn[count(i)]);          // uniform per warp: symbols nobody uses
                        uint32_t c1 = acc.value((uint32_t)i + 1u);
                        if (i == 31) c1 = 0x10000u;                              // cdf[32] wraps to 0 in 16 bits and means 65536
                        const uint32_t e = rans_table_entry(c0, c1);
                        if (TR) tab[i * CT + tid] = e;
                        else tab[tid * kLp + i] = e;
                        c0 = c1;
                    }
                }
            }
            __syncthreads();                                                     // pn is dead: maxima and LUT take its place
            for (int i = tid; i < gt; i += CT) mx[i] = half_to_float(maxes[i], P.max_dtype);
            if (tid < 32) lut[tid] = dequant_lut((uint32_t)tid, cq);
            built = true;
        }
    }
    if (!built) {
        if constexpr (CODER == CODER_RANS && TR) {
            // TRANSPOSED table: entry i of stream tid at tab[i * CT + tid], entry = (cdf[i] << 16) | freq(i).  Built in two
            // steps through a staging copy of the raw CDF rows that lives in the table's own upper half (bytes 8448..16895):
            // coalesced global -> staging; every thread turns ITS row (33 halfwords, stride 33: conflict-free) into column
            // entries 0..15 (bytes 0..8191, clear of the staging); entries 16..31 -- needed by 32-bin planes only -- go
            // through registers so that they may overwrite the staging once everybody has read it.
            uint16_t* stg = reinterpret_cast<uint16_t*>(smem) + (CT * kLp);            // second half of the CT*kLp words
            for (int e = tid; e < ncols * kLp; e += CT) stg[e] = __ldg(cdf_src + e);
            for (int i = tid; i < gt; i += CT) mx[i] = half_to_float(maxes[i], P.max_dtype);
            if (tid < 32) lut[tid] = dequant_lut((uint32_t)tid, cq);
            __syncthreads();
            const uint16_t* my = stg + tid * kLp;
            if (active) {
                uint32_t c0 = my[0];
    #pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const uint32_t c1 = my[i + 1];
                    tab[i * CT + tid] = rans_table_entry(c0, c1);
                    c0 = c1;
                }
            }
            if (cq > 7.0f) {                                                            // uniform per CTA
                uint32_t cv[17];
                if (active) {
    #pragma unroll
                    for (int i = 0; i < 16; ++i) cv[i] = my[16 + i];
                    cv[16] = 0x10000u;                                                   // cdf[32] is stored as 0 and means 65536
                }
                __syncthreads();
                if (active) {
    #pragma unroll
                    for (int i = 0; i < 16; ++i) tab[(16 + i) * CT + tid] = rans_table_entry(cv[i], cv[i + 1]);
                }
            }
        } else {
            for (int e = tid; e < ncols * kLp; e += CT) {
                const uint32_t i = (uint32_t)e % (uint32_t)kLp;
                const uint32_t c0 = __ldg(cdf_src + e);
                if constexpr (CODER == CODER_RANS) {
                    // (cdf[i] << 16) | freq(i); cdf[32] is stored as 0 and stands for 65536; entry 32 is never searched
                    const uint32_t c1 = i < 31u ? (uint32_t)__ldg(cdf_src + e + 1) : 0x10000u;
                    tab[e] = i < 32u ? rans_table_entry(c0, c1) : 0xFFFFFFFFu;
                } else {
                    tab[e] = dec_table_entry(i, c0);
                }
            }
            for (int i = tid; i < gt; i += CT) mx[i] = half_to_float(maxes[i], P.max_dtype);
            if (tid < 32) lut[tid] = dequant_lut((uint32_t)tid, cq);
        }
    }
    __syncthreads();

    if (!active) return;
    const uint32_t* erow = tab + tid * kLp;
    const int h = c / P.D;
    uint16_t* dst = const_cast<uint16_t*>(P.pt.p[nl]) + (PAGED ? 0 : (dc.dst_tok + tok0) * P.sT) + (int64_t)h * P.sH +
                    (c - h * P.D);
    const int64_t* slots = PAGED ? P.slot_map + dc.dst_tok + tok0 : nullptr;
    uint32_t bad = beyond ? 2u : 0u;
    if constexpr (CODER == CODER_RANS) {
        uint32_t xf;
        const uint32_t one = min((uint32_t)P.n_chunks, 1u);      // 1, but opaque to the compiler (see rans_decode_stream)
        const uint32_t* pk = TR ? tab + tid : tab + tid * kLp;
        if (cq <= 7.0f) xf = rans_decode_stream<OUT_DT, 4, PAGED, TR>(dc.base, my_off + hl, pk, lut, mx, dst, (uint32_t)P.sT, gt, slots, one);
        else xf = rans_decode_stream<OUT_DT, 5, PAGED, TR>(dc.base, my_off + hl, pk, lut, mx, dst, (uint32_t)P.sT, gt, slots, one);
        bad |= xf != kRansLow ? 1u : 0u;
    } else {
        if (cq <= 7.0f) decode_stream<OUT_DT, 4, PAGED>(dc.base, my_off, erow, lut, mx, dst, (uint32_t)P.sT, gt, slots);   // <= 16 bins: symbols 0..14
        else decode_stream<OUT_DT, 5, PAGED>(dc.base, my_off, erow, lut, mx, dst, (uint32_t)P.sT, gt, slots);
    }
    if (bad != 0u && P.status != nullptr) atomicOr(&P.status[j], bad);
}

// ------------------------------------------------------------------------------------------ host side
int make_plane_table(const b200kv_kv_desc* kv, const float* key_bins, const float* value_bins, PlaneTable* out) {
    B2_REQUIRE(kv != nullptr, "kv descriptor is NULL");
    B2_REQUIRE(kv->L > 0 && 2 * kv->L <= B200KV_MAX_PLANES, "L out of range");
    B2_REQUIRE(kv->H > 0 && kv->D > 0, "H/D must be positive");
    B2_REQUIRE(kv->dtype == B200KV_DT_BF16 || kv->dtype == B200KV_DT_FP16, "dtype must be bf16 or fp16");
    B2_REQUIRE(kv->planes != nullptr || kv->base != nullptr, "no KV pointer");
    for (int kvi = 0; kvi < 2; ++kvi)
        for (int l = 0; l < kv->L; ++l) {
            const int nl = kvi * kv->L + l;
            const uint16_t* p = kv->planes ? static_cast<const uint16_t*>(kv->planes[nl])
                                           : static_cast<const uint16_t*>(kv->base) + l * kv->sL + kvi * kv->sKV;
            B2_REQUIRE(p != nullptr, "NULL plane pointer");
            out->p[nl] = p;
            const float bins = kvi ? value_bins[l] : key_bins[l];
            out->maxq[nl] = floorf(bins / 2.0f) - 1.0f;       // bins // 2 - 1  (cachegen_encoder.py:53)
            B2_REQUIRE(out->maxq[nl] >= 1.0f && out->maxq[nl] <= 15.0f, "bins must be in [4, 32]");
        }
    return 0;
}

static int tiles_per_plane(int C) { return (C + CT - 1) / CT; }

// ---- optional per-kernel timing (bench.py's roofline leg): events around each launch of the last call
enum { kProfAbsmax = 0, kProfCdf, kProfEncode, kProfFinalize, kProfTileSum, kProfTileScan, kProfDecode, kProfCount };
static bool g_prof_on = false;
static cudaEvent_t g_prof_ev[kProfCount][2];
static bool g_prof_have[kProfCount];
static bool g_prof_init = false;

struct ProfScope {
    int slot;
    cudaStream_t stream;
    ProfScope(int slot_, cudaStream_t s) : slot(slot_), stream(s) {
        if (!g_prof_on) return;
        if (!g_prof_init) {
            for (int i = 0; i < kProfCount; ++i) { cudaEventCreate(&g_prof_ev[i][0]); cudaEventCreate(&g_prof_ev[i][1]); }
            g_prof_init = true;
        }
        cudaEventRecord(g_prof_ev[slot][0], stream);
    }
    ~ProfScope() {
        if (!g_prof_on) return;
        cudaEventRecord(g_prof_ev[slot][1], stream);
        g_prof_have[slot] = true;
    }
};

static int enc_tempw(bool fused, int coder, bool compact = false) {
    if (compact) return TEMPW_FUSED_RANS_HDR;
    return fused ? (coder == CODER_RANS ? TEMPW_FUSED_RANS : TEMPW_FUSED) : TEMPW_SPLIT;
}

static size_t enc_ws_layout(int64_t n_tiles_alloc, int n_chunks, int tempw, int coder, bool compact, size_t* off_tot,
                            size_t* off_totals, size_t* off_err, size_t* off_state, size_t* off_temp) {
    size_t o = 0;
    *off_tot = o;    o += (size_t)n_tiles_alloc * 4;  o = (o + 255) & ~(size_t)255;
    *off_totals = o; o += (size_t)n_chunks * 8;       o = (o + 255) & ~(size_t)255;
    *off_err = o;    o += (size_t)n_chunks * 4;       o = (o + 255) & ~(size_t)255;
    *off_state = o;  o += coder == CODER_RANS ? (size_t)n_tiles_alloc * CT * (compact ? 16 : 4) : 0;  o = (o + 255) & ~(size_t)255;
    *off_temp = o;   o += (size_t)n_tiles_alloc * CT * (size_t)tempw * 4;
    return (o + 255) & ~(size_t)255;
}

static size_t dec_ws_layout(int64_t tiles_max, int n_chunks, size_t* off_tb) {
    size_t o = ((size_t)n_chunks * sizeof(DecChunk) + 255) & ~(size_t)255;
    *off_tb = o;
    o += (size_t)n_chunks * (size_t)tiles_max * 8;
    return (o + 255) & ~(size_t)255;
}

}  // namespace b200kv

using namespace b200kv;

extern "C" {

int b200kv_container_layout(int32_t L, int32_t H, int32_t D, int32_t ntokens, b200kv_layout* out) {
    return b200kv_container_layout_v(L, H, D, ntokens, CODER_RANS, out);
}

int b200kv_container_layout_v(int32_t L, int32_t H, int32_t D, int32_t ntokens, int32_t coder, b200kv_layout* out) {
    B2_REQUIRE(out != nullptr && L > 0 && H > 0 && D > 0 && ntokens > 0, "bad shape");
    B2_REQUIRE(coder >= CODER_AC && coder <= CODER_RANS_COMPACT, "unknown coder");
    const int compact = coder == CODER_RANS_COMPACT ? 1 : 0;
    B2_REQUIRE(!compact || ntokens <= kGroup, "the compact container holds chunks of at most 256 tokens");
    const Layout lo = make_layout(L, H * D, ntokens, compact);
    out->off_cdf = lo.off_cdf;
    out->off_maxes = lo.off_maxes;
    out->off_lengths = lo.off_lengths;
    out->off_payload = lo.off_payload;
    out->fixed_bytes = lo.off_payload;
    // per stream per group: <= 16 bits per symbol (CDF width >= 1/65536) + termination -- 2 flush bits + pad for the
    // arithmetic coder, the 32-bit final state for rANS
    const int64_t streams = 2 * (int64_t)L * H * D;
    out->max_total_bytes = align16(lo.off_payload + streams * (2 * (int64_t)ntokens + 4 * (int64_t)lo.ngroups +
                                                              (compact ? kHdrMax : 0)) + 16);
    return 0;
}

int64_t b200kv_encode_workspace_bytes(int32_t L, int32_t H, int32_t D, int32_t chunk_tokens, int32_t n_chunks,
                                      int32_t coder) {
    if (L <= 0 || H <= 0 || D <= 0 || chunk_tokens <= 0 || n_chunks <= 0) return -2;
    coder &= 0xff;
    const bool compact = coder == CODER_RANS_COMPACT;          // same kernels; rows hold the stream header too
    if (compact) coder = CODER_RANS;
    if (coder != CODER_AC && coder != CODER_RANS) return -2;
    if (compact && chunk_tokens > kGroup) return -2;
    const int64_t G = (chunk_tokens + kGroup - 1) / kGroup;
    const int64_t n_tiles = (int64_t)n_chunks * G * 2 * L * tiles_per_plane(H * D);
    size_t a, b, c, d, e;
    return (int64_t)enc_ws_layout(n_tiles, n_chunks, enc_tempw(chunk_tokens <= kGroup, coder, compact), coder, compact, &a, &b, &c, &d, &e);
}

int64_t b200kv_decode_workspace_bytes(int32_t L, int32_t H, int32_t D, int32_t chunk_tokens, int32_t n_chunks) {
    if (L <= 0 || H <= 0 || D <= 0 || chunk_tokens <= 0 || n_chunks <= 0) return -2;
    const int64_t G = (chunk_tokens + kGroup - 1) / kGroup;
    const int64_t tiles_max = G * 2 * L * tiles_per_plane(H * D);
    size_t a;
    return (int64_t)dec_ws_layout(tiles_max, n_chunks, &a);
}

int b200kv_encode_chunks(const b200kv_kv_desc* kv, int64_t tok_begin, int32_t n_chunks, int32_t chunk_tokens,
                         int32_t last_chunk_tokens, const float* key_bins, const float* value_bins, int32_t coder,
                         void* out, int64_t out_stride, uint64_t* sizes_out, void* workspace, int64_t workspace_bytes,
                         void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    EncParams P;
    B2_REQUIRE(key_bins && value_bins, "bins are NULL");
    const bool hint_tma = (coder & B200KV_ENCODE_HINT_HIGH_ENTROPY) != 0;
    const bool hint_mid = (coder & B200KV_ENCODE_HINT_MID_ENTROPY) != 0;
    coder &= 0xff;
    B2_REQUIRE(coder >= CODER_AC && coder <= CODER_RANS_COMPACT, "coder must be one of B200KV_CODER_*");
    P.compact = coder == CODER_RANS_COMPACT ? 1 : 0;
    if (P.compact) coder = CODER_RANS;                          // version 3 = rANS payload + compact side information
    P.coder = coder;
    if (int rc = make_plane_table(kv, key_bins, value_bins, &P.pt)) return rc;
    B2_REQUIRE(n_chunks > 0 && chunk_tokens > 0, "n_chunks / chunk_tokens must be positive");
    B2_REQUIRE(!P.compact || chunk_tokens <= kGroup, "the compact container (B200KV_CODER_RANS_COMPACT) holds chunks of at most 256 tokens");
    B2_REQUIRE(last_chunk_tokens > 0 && last_chunk_tokens <= chunk_tokens, "last_chunk_tokens out of range");
    B2_REQUIRE(out != nullptr && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (out_stride & 15) == 0,
               "out / out_stride must be 16-byte aligned");
    B2_REQUIRE(tok_begin >= 0, "tok_begin must be >= 0");
    P.sT = kv->sT; P.sH = kv->sH; P.tok_begin = tok_begin;
    P.slot_map = kv->slot_map;
    const bool paged = kv->slot_map != nullptr;
    P.L = kv->L; P.H = kv->H; P.D = kv->D; P.C = kv->H * kv->D; P.dtype = kv->dtype;
    P.n_chunks = n_chunks; P.chunk_tokens = chunk_tokens; P.last_chunk_tokens = last_chunk_tokens;
    P.tpp = tiles_per_plane(P.C);
    P.out = static_cast<uint8_t*>(out);
    P.out_stride = out_stride;
    P.sizes_out = sizes_out;
    const Layout lo = make_layout(P.L, P.C, chunk_tokens, P.compact);
    B2_REQUIRE(out_stride >= lo.off_payload + 16, "out_stride smaller than the fixed container sections");

    const int64_t G = lo.ngroups;
    const int64_t per_group = 2 * (int64_t)P.L * P.tpp;
    const int64_t tiles_full = G * per_group;
    const int64_t n_tiles = (int64_t)n_chunks * tiles_full;     // tiles beyond a ragged last chunk exit at once
    const bool fused = chunk_tokens <= kGroup;
    P.tiles_full = (int32_t)tiles_full;
    P.tempw = enc_tempw(fused, coder, P.compact != 0);
    size_t off_tot, off_totals, off_err, off_state, off_temp;
    const size_t need = enc_ws_layout(n_tiles, n_chunks, P.tempw, coder, P.compact != 0, &off_tot, &off_totals, &off_err, &off_state,
                                      &off_temp);
    B2_REQUIRE(workspace != nullptr && workspace_bytes >= (int64_t)need, "workspace too small");
    B2_REQUIRE(n_tiles < (1ll << 31) && tiles_full < (1ll << 31), "too many tiles in one call");
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    P.tile_tot = reinterpret_cast<uint32_t*>(ws + off_tot);
    P.totals = reinterpret_cast<unsigned long long*>(ws + off_totals);
    P.err = reinterpret_cast<unsigned int*>(ws + off_err);
    P.temp = reinterpret_cast<uint32_t*>(ws + off_temp);
    P.rstate = reinterpret_cast<uint32_t*>(ws + off_state);
    B2_CHECK_CUDA(cudaMemsetAsync(ws, 0, off_state, stream));    // counters only; states and temp rows need no init

    // 1) per-(plane, token) absmax -> maxes sections
    const int64_t total_tokens = (int64_t)(n_chunks - 1) * chunk_tokens + last_chunk_tokens;
    for (int i = 0; i <= kProfFinalize; ++i) g_prof_have[i] = false;
    {
        bool vec = (kv->D % 8 == 0) && (kv->sT % 8 == 0) && (kv->sH % 8 == 0);
        for (int nl = 0; nl < 2 * P.L && vec; ++nl) vec = (reinterpret_cast<uintptr_t>(P.pt.p[nl]) & 15) == 0;
        const int64_t rows = 2 * (int64_t)P.L * total_tokens;
        const int64_t blocks = (rows + 7) / 8;
        B2_REQUIRE(blocks < (1ll << 31), "too many rows in one call");
        ProfScope prof(kProfAbsmax, stream);
        if (vec && !paged) absmax_kernel<true, false><<<(unsigned)blocks, 256, 0, stream>>>(P, total_tokens);
        else if (vec) absmax_kernel<true, true><<<(unsigned)blocks, 256, 0, stream>>>(P, total_tokens);
        else if (!paged) absmax_kernel<false, false><<<(unsigned)blocks, 256, 0, stream>>>(P, total_tokens);
        else absmax_kernel<false, true><<<(unsigned)blocks, 256, 0, stream>>>(P, total_tokens);
        B2_CHECK_CUDA(cudaGetLastError());
    }
    // 2) encode (streams -> temp rows, lengths, tile totals)
    const size_t smem_fused = (size_t)(((CT * SYMW + (CT * kLp * 2 + 3) / 4 + 3) & ~3) + kGroup + 8) * 4;
    const size_t smem_split = (size_t)(((CT * PAIRW + 3) & ~3) + kGroup + 4) * 4;
    const size_t smem_cdf = (size_t)(CT * PAIRW + kGroup) * 4;
#define B2_LAUNCH_ENC1(FUSED, DT, PAGED, CODER, SMEM)                                                      \
    do {                                                                                                   \
        B2_CHECK_CUDA(cudaFuncSetAttribute(encode_kernel<FUSED, DT, PAGED, CODER>,                         \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM)));     \
        encode_kernel<FUSED, DT, PAGED, CODER><<<(unsigned)n_tiles, CT, (SMEM), stream>>>(P);              \
    } while (0)
#define B2_LAUNCH_ENC(FUSED, DT, PAGED, SMEM)                                                              \
    do {                                                                                                   \
        if (coder == CODER_RANS) B2_LAUNCH_ENC1(FUSED, DT, PAGED, CODER_RANS, SMEM);                       \
        else B2_LAUNCH_ENC1(FUSED, DT, PAGED, CODER_AC, SMEM);                                             \
    } while (0)
#define B2_LAUNCH_ENC2(FUSED, SMEM)                                                                        \
    do {                                                                                                   \
        if (P.dtype == B200KV_DT_BF16) { if (paged) B2_LAUNCH_ENC(FUSED, 0, true, SMEM); else B2_LAUNCH_ENC(FUSED, 0, false, SMEM); } \
        else { if (paged) B2_LAUNCH_ENC(FUSED, 1, true, SMEM); else B2_LAUNCH_ENC(FUSED, 1, false, SMEM); } \
    } while (0)
    // TMA-staged kernel (encode_tma_kernel): rANS, fused mode, tiles of exactly CT channels that are contiguous in every
    // token row and 16-byte aligned.  B200KV_ENCODE_PATH=legacy forces the kernel above (A/B measurements).
    // Measured (profiles/r2_encode_variants.json): the TMA-staged kernel is flat in the data's entropy (4.4 .. 5.2 ms per
    // 8192-token block), the register-staged one is faster below ~2.7 payload bits per symbol (3.75 ms at 0.6) and slower
    // above (6.2 ms at 4.1): the caller says which regime it expects (B200KV_ENCODE_HINT_HIGH_ENTROPY, e.g. from the
    // sizes of the previous call).  B200KV_ENCODE_PATH=legacy|tma overrides (A/B measurements).
    bool tma = hint_tma;
    if (const char* e = getenv("B200KV_ENCODE_PATH")) tma = e[0] == 't';
    tma = tma && fused && coder == CODER_RANS && P.C % CT == 0 && (kv->sH == kv->D || kv->D % CT == 0) &&
          kv->sT % 8 == 0 && kv->sH % 8 == 0;
    for (int nl = 0; nl < 2 * P.L && tma; ++nl) tma = (reinterpret_cast<uintptr_t>(P.pt.p[nl]) & 15) == 0;
    if (tma) {
        ProfScope prof(kProfEncode, stream);
#define B2_LAUNCH_TMA(DT, PAGED)                                                                                       \
    do {                                                                                                               \
        B2_CHECK_CUDA(cudaFuncSetAttribute(encode_tma_kernel<DT, PAGED>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                           kEncTmaSmem));                                                              \
        encode_tma_kernel<DT, PAGED><<<(unsigned)n_tiles, CT, kEncTmaSmem, stream>>>(P);                               \
    } while (0)
        if (P.dtype == B200KV_DT_BF16) { if (paged) B2_LAUNCH_TMA(0, true); else B2_LAUNCH_TMA(0, false); }
        else { if (paged) B2_LAUNCH_TMA(1, true); else B2_LAUNCH_TMA(1, false); }
#undef B2_LAUNCH_TMA
    } else if (fused) {
        ProfScope prof(kProfEncode, stream);
        B2_LAUNCH_ENC2(true, smem_fused);
    } else {
        const unsigned cdf_blocks = (unsigned)((int64_t)n_chunks * per_group);
        {
            ProfScope prof(kProfCdf, stream);
            if (P.dtype == B200KV_DT_BF16) {
                if (paged) cdf_kernel<0, true><<<cdf_blocks, CT, smem_cdf, stream>>>(P);
                else cdf_kernel<0, false><<<cdf_blocks, CT, smem_cdf, stream>>>(P);
            } else {
                if (paged) cdf_kernel<1, true><<<cdf_blocks, CT, smem_cdf, stream>>>(P);
                else cdf_kernel<1, false><<<cdf_blocks, CT, smem_cdf, stream>>>(P);
            }
        }
        B2_CHECK_CUDA(cudaGetLastError());
        ProfScope prof(kProfEncode, stream);
        B2_LAUNCH_ENC2(false, smem_split);
    }
#undef B2_LAUNCH_ENC2
#undef B2_LAUNCH_ENC
#undef B2_LAUNCH_ENC1
    B2_CHECK_CUDA(cudaGetLastError());
    // 3) compaction (collect_bytes) + headers + sizes
    {
        ProfScope prof(kProfFinalize, stream);
        // stage = the fused mode's worst case (128 x 160 B + alignment phase); in split mode a tile can in principle
        // reach 128 x 528 B, but sizing the stage for that would leave 3 CTAs per SM for streams that are typically
        // a few dozen bytes long -- oversized tiles take the direct path inside the kernel
        P.stage_bytes = coder == CODER_RANS ? CT * (TEMPW_FUSED_RANS * 4 + 4) + 32 : CT * TEMPW_FUSED * 4 + 32;
        // The kernel waits on sparse row reads: more resident CTAs hide more of that latency.  Without an entropy hint a
        // tile's streams total a few KB, so a 12 KB stage (12 CTAs per SM instead of 9: 0.41 -> 0.31 ms per block) covers
        // them; the rare larger tile takes the kernel's direct path.  B200KV_COMPACT_STAGE=<bytes> overrides (knob).
        if (coder == CODER_RANS && !hint_tma && !hint_mid) P.stage_bytes = 12 * 1024;
        if (const char* e = getenv("B200KV_COMPACT_STAGE")) {
            const int v = atoi(e);
            if (v >= 1024 && v <= P.stage_bytes + 16 * 1024) P.stage_bytes = v & ~15;
        }
        B2_CHECK_CUDA(cudaFuncSetAttribute(compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P.stage_bytes));
        enc_scan_kernel<<<(unsigned)n_chunks, 1024, 0, stream>>>(P);
        compact_kernel<<<(unsigned)n_tiles, CT, (size_t)P.stage_bytes, stream>>>(P);
        finalize_kernel<<<(n_chunks + 127) / 128, 128, 0, stream>>>(P);
    }
    B2_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int b200kv_decode_chunks(const void* containers, int64_t containers_bytes, const int64_t* offsets,
                         const int64_t* total_bytes, const int32_t* ntokens, const int64_t* dst_tok, int32_t n_chunks,
                         int32_t max_dtype,
                         int32_t coder, const b200kv_kv_desc* dst, const float* key_bins, const float* value_bins,
                         uint32_t* status_out, void* workspace, int64_t workspace_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DecParams P;
    B2_REQUIRE(key_bins && value_bins, "bins are NULL");
    B2_REQUIRE(coder >= CODER_AC && coder <= CODER_RANS_COMPACT, "coder must be one of B200KV_CODER_*");
    P.compact = coder == CODER_RANS_COMPACT ? 1 : 0;
    if (P.compact) coder = CODER_RANS;
    if (int rc = make_plane_table(dst, key_bins, value_bins, &P.pt)) return rc;
    B2_REQUIRE(containers && offsets && total_bytes && ntokens && dst_tok && n_chunks > 0, "bad chunk arrays");
    B2_REQUIRE(max_dtype == B200KV_DT_BF16 || max_dtype == B200KV_DT_FP16, "bad max_dtype");
    B2_REQUIRE(dst->sT > 0 && dst->sT < (1ll << 23), "destination token stride out of range");
    P.sT = dst->sT; P.sH = dst->sH;
    P.slot_map = dst->slot_map;
    P.L = dst->L; P.H = dst->H; P.D = dst->D; P.C = dst->H * dst->D;
    P.out_dtype = dst->dtype; P.max_dtype = max_dtype; P.n_chunks = n_chunks;
    P.tpp = tiles_per_plane(P.C);
    int tmax = 0;
    for (int j = 0; j < n_chunks; ++j) {
        B2_REQUIRE(ntokens[j] > 0, "ntokens must be positive");
        B2_REQUIRE(!P.compact || ntokens[j] <= kGroup, "a compact container holds at most 256 tokens");
        B2_REQUIRE((offsets[j] & 15) == 0, "container offsets must be 16-byte aligned");
        // the fixed sections are addressed from (L, H, D, ntokens); the buffer must hold them in full
        const Layout lj = make_layout(P.L, P.C, ntokens[j], P.compact);
        B2_REQUIRE(total_bytes[j] >= lj.off_payload && total_bytes[j] - lj.off_payload < (1ll << 32),
                   "container shorter than its fixed sections (truncated or corrupt)");
        B2_REQUIRE(offsets[j] >= 0 && offsets[j] + total_bytes[j] + B200KV_READ_SLACK <= containers_bytes,
                   "containers buffer must extend B200KV_READ_SLACK bytes past the end of every container");
        tmax = ntokens[j] > tmax ? ntokens[j] : tmax;
    }
    const int64_t Gmax = (tmax + kGroup - 1) / kGroup;
    const int64_t tiles_max = Gmax * 2 * P.L * P.tpp;
    B2_REQUIRE(tiles_max < (1ll << 31) && n_chunks <= 65535, "too many tiles / chunks in one call");
    // table layout of the rANS decoder: the conflict-free (transposed) one pays off above ~3.6 payload bits per symbol
    // (measured: 3.28 / 3.60 ms at 0.6 bits, 4.04 / 3.72 ms at 4.1 bits, row-major / transposed); the containers say how
    // many bits they hold.  B200KV_DECODE_TABLE=rows|transposed overrides (measurement knob).
    bool transposed = false;
    {
        double bits = 0.0, syms = 0.0;
        for (int j = 0; j < n_chunks; ++j) {
            const Layout lj = make_layout(P.L, P.C, ntokens[j], P.compact);
            bits += 8.0 * (double)(total_bytes[j] - lj.off_payload);
            syms += 2.0 * P.L * (double)P.C * ntokens[j];
        }
        // a version-3 payload also carries the stream histograms: ~0.5 bits per symbol at that entropy
        transposed = bits > (P.compact ? 4.1 : 3.6) * syms && bits < 6.0 * syms;   // a slot bound instead of a size says nothing: rows
        if (const char* e = getenv("B200KV_DECODE_TABLE")) transposed = e[0] == 't';
    }
    P.tiles_max = (int32_t)tiles_max;
    size_t off_tb;
    const size_t need = dec_ws_layout(tiles_max, n_chunks, &off_tb);
    B2_REQUIRE(workspace != nullptr && workspace_bytes >= (int64_t)need, "workspace too small");
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    // chunk descriptors: small pageable -> device copy (staged by the driver before the call returns)
    {
        DecChunk* hc = static_cast<DecChunk*>(malloc(sizeof(DecChunk) * (size_t)n_chunks));
        B2_REQUIRE(hc != nullptr, "out of host memory");
        for (int j = 0; j < n_chunks; ++j) {
            hc[j].base = static_cast<const uint8_t*>(containers) + offsets[j];
            hc[j].dst_tok = dst_tok[j];
            hc[j].t = ntokens[j];
            hc[j].ngroups = (ntokens[j] + kGroup - 1) / kGroup;
            const Layout lj = make_layout(P.L, P.C, ntokens[j], P.compact);
            hc[j].payload_bytes = (uint32_t)(total_bytes[j] - lj.off_payload);
            hc[j].pad = 0;
        }
        cudaError_t e = cudaMemcpyAsync(ws, hc, sizeof(DecChunk) * (size_t)n_chunks, cudaMemcpyHostToDevice, stream);
        free(hc);
        B2_CHECK_CUDA(e);
    }
    P.chunks = reinterpret_cast<const DecChunk*>(ws);
    P.tile_base = reinterpret_cast<unsigned long long*>(ws + off_tb);
    P.status = status_out;
    if (status_out) B2_CHECK_CUDA(cudaMemsetAsync(status_out, 0, sizeof(uint32_t) * (size_t)n_chunks, stream));

    dim3 gsum((unsigned)((tiles_max + 3) / 4), (unsigned)n_chunks);
    {
        ProfScope prof(kProfTileSum, stream);
        tile_sum_kernel<<<gsum, 128, 0, stream>>>(P);
    }
    B2_CHECK_CUDA(cudaGetLastError());
    {
        ProfScope prof(kProfTileScan, stream);
        tile_scan_kernel<<<(unsigned)n_chunks, 1024, 0, stream>>>(P);
    }
    B2_CHECK_CUDA(cudaGetLastError());

    const size_t smem = (size_t)(CT * kLp + kGroup + 32) * 4;
    dim3 grid((unsigned)tiles_max, (unsigned)n_chunks);
    ProfScope prof(kProfDecode, stream);
#define B2_LAUNCH_DEC1(DT, PAGED, CODER, TR)                                                                           \
    do {                                                                                                               \
        B2_CHECK_CUDA(cudaFuncSetAttribute(decode_kernel<DT, PAGED, CODER, TR>,                                        \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                   \
        decode_kernel<DT, PAGED, CODER, TR><<<grid, CT, smem, stream>>>(P);                                            \
    } while (0)
#define B2_LAUNCH_DEC(DT, PAGED)                                                                                       \
    do {                                                                                                               \
        if (coder == CODER_RANS && transposed) B2_LAUNCH_DEC1(DT, PAGED, CODER_RANS, true);                            \
        else if (coder == CODER_RANS) B2_LAUNCH_DEC1(DT, PAGED, CODER_RANS, false);                                    \
        else B2_LAUNCH_DEC1(DT, PAGED, CODER_AC, false);                                                               \
    } while (0)
    if (P.out_dtype == B200KV_DT_BF16) { if (P.slot_map) B2_LAUNCH_DEC(0, true); else B2_LAUNCH_DEC(0, false); }
    else { if (P.slot_map) B2_LAUNCH_DEC(1, true); else B2_LAUNCH_DEC(1, false); }
#undef B2_LAUNCH_DEC
#undef B2_LAUNCH_DEC1
    B2_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int b200kv_profile_enable(int32_t on) {
    g_prof_on = on != 0;
    for (int i = 0; i < kProfCount; ++i) g_prof_have[i] = false;
    return 0;
}

int b200kv_profile_last(float* ms, int32_t n) {
    B2_REQUIRE(ms != nullptr && n >= kProfCount, "need room for 7 floats");
    for (int i = 0; i < kProfCount; ++i) {
        ms[i] = -1.0f;
        if (g_prof_on && g_prof_have[i]) {
            B2_CHECK_CUDA(cudaEventSynchronize(g_prof_ev[i][1]));
            B2_CHECK_CUDA(cudaEventElapsedTime(&ms[i], g_prof_ev[i][0], g_prof_ev[i][1]));
        }
    }
    return kProfCount;
}

}  // extern "C"
