// mover.cu -- blob pack/unpack kernels and the GPU <-> pinned-host mover primitives.
//
// Replaces (reference paths relative to the LMCache v0.1.2 tree):
//   lmcache/cache_engine.py:98-118   _tuple_kv_to_blob    (3x torch.stack + permute)
//   lmcache/cache_engine.py:131-161  _slice_kv_at         (split + .contiguous() per chunk)
//   lmcache/cache_engine.py:362-368  retrieve-side torch.cat + _blob_to_tuple_kv
//   lmcache/storage_backend/local_backend.py:82-100,141-144  pageable .to("cpu") / .to("cuda") + device sync
// with ONE gather (store) / scatter (retrieve) pass between the engine's 2L KV tensors and the per-chunk
// blobs.  When the chunk buffer is pinned host memory mapped into the device address space, the same
// kernel is the device->host (or host->device) mover: the data crosses PCIe exactly once, as 16-byte
// coalesced accesses, with no intermediate device blob.
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "common.cuh"

namespace b200kv {

struct PackParams {
    PlaneTable pt;
    int64_t sT, sH, tok_begin;
    const int64_t* slot_map;       // paged KV: token i lives in row slot_map[i]; NULL = row i
    int32_t L, H, D, n_chunks, chunk_tokens, last_chunk_tokens, hf_layout;
    uint8_t* chunks;
    int64_t chunk_stride_bytes;
};

// One grid-stride loop over (chunk, plane, token, vector) units; VEC halfs per unit.
// vllm chunk layout [L,2,t,H,D]; huggingface [L,2,H,t,D].
template <int VEC, bool PACK>
__global__ void __launch_bounds__(256) pack_kernel(PackParams P) {
    using vec_t = typename std::conditional<VEC == 8, uint4, uint16_t>::type;
    const int NL = 2 * P.L;
    const int vph = P.D / VEC;                 // vectors per head row
    const int64_t vpt = (int64_t)P.H * vph;    // vectors per token
    const int64_t per_chunk_full = (int64_t)NL * P.chunk_tokens * vpt;
    const int64_t total = (int64_t)(P.n_chunks - 1) * per_chunk_full + (int64_t)NL * P.last_chunk_tokens * vpt;
    for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < total; u += (int64_t)gridDim.x * blockDim.x) {
        int64_t j = u / per_chunk_full;
        if (j >= P.n_chunks) j = P.n_chunks - 1;
        int64_t r = u - j * per_chunk_full;
        const int t = (j == P.n_chunks - 1) ? P.last_chunk_tokens : P.chunk_tokens;
        // r indexes [l][kv][tok][h][v] of the chunk (vllm order) -- map to plane nl = kv*L + l
        const int64_t per_plane = (int64_t)t * vpt;
        const int lk = (int)(r / per_plane);       // l*2 + kv
        r -= (int64_t)lk * per_plane;
        const int tok = (int)(r / vpt);
        r -= (int64_t)tok * vpt;
        const int h = (int)(r / vph);
        const int v = (int)(r - (int64_t)h * vph);
        const int l = lk >> 1, kv = lk & 1;
        const uint16_t* plane = P.pt.p[kv * P.L + l];
        int64_t row = P.tok_begin + j * P.chunk_tokens + tok;
        if (P.slot_map) row = __ldg(P.slot_map + row);       // consecutive threads share the token: broadcast, L1 hit
        const int64_t src_off = row * P.sT + (int64_t)h * P.sH + (int64_t)v * VEC;
        int64_t dst_off;   // in halfs, inside the chunk
        if (P.hf_layout) dst_off = (((int64_t)lk * P.H + h) * t + tok) * P.D + (int64_t)v * VEC;
        else dst_off = (((int64_t)lk * t + tok) * P.H + h) * P.D + (int64_t)v * VEC;
        uint16_t* cptr = reinterpret_cast<uint16_t*>(P.chunks + j * P.chunk_stride_bytes) + dst_off;
        if (PACK) *reinterpret_cast<vec_t*>(cptr) = *reinterpret_cast<const vec_t*>(plane + src_off);
        else *reinterpret_cast<vec_t*>(const_cast<uint16_t*>(plane) + src_off) = *reinterpret_cast<const vec_t*>(cptr);
    }
}

static int launch_pack(bool pack, const b200kv_kv_desc* kv, int64_t tok_begin, int32_t n_chunks, int32_t chunk_tokens,
                       int32_t last_chunk_tokens, int32_t hf_layout, void* chunks, int64_t chunk_stride_bytes,
                       cudaStream_t stream) {
    PackParams P;
    B2_REQUIRE(kv != nullptr && kv->L > 0 && 2 * kv->L <= B200KV_MAX_PLANES, "bad kv descriptor");
    float bins[B200KV_MAX_PLANES];
    for (int i = 0; i < B200KV_MAX_PLANES; ++i) bins[i] = 32.0f;   // unused by pack/unpack; keeps the table valid
    if (int rc = make_plane_table(kv, bins, bins, &P.pt)) return rc;
    B2_REQUIRE(n_chunks > 0 && chunk_tokens > 0 && last_chunk_tokens > 0 && last_chunk_tokens <= chunk_tokens,
               "bad chunking");
    B2_REQUIRE(chunks != nullptr, "chunks is NULL");
    const int64_t chunk_bytes = 2ll * kv->L * 2 * chunk_tokens * kv->H * kv->D;
    B2_REQUIRE(chunk_stride_bytes >= chunk_bytes || n_chunks == 1, "chunk_stride_bytes too small");
    P.sT = kv->sT; P.sH = kv->sH; P.tok_begin = tok_begin;
    P.slot_map = kv->slot_map;
    P.L = kv->L; P.H = kv->H; P.D = kv->D;
    P.n_chunks = n_chunks; P.chunk_tokens = chunk_tokens; P.last_chunk_tokens = last_chunk_tokens;
    P.hf_layout = hf_layout;
    P.chunks = static_cast<uint8_t*>(chunks);
    P.chunk_stride_bytes = chunk_stride_bytes;
    bool vec = (kv->D % 8 == 0) && (kv->sT % 8 == 0) && (kv->sH % 8 == 0) &&
               ((reinterpret_cast<uintptr_t>(chunks) & 15) == 0) && (chunk_stride_bytes % 16 == 0);
    for (int nl = 0; nl < 2 * P.L && vec; ++nl) vec = (reinterpret_cast<uintptr_t>(P.pt.p[nl]) & 15) == 0;
    const int V = vec ? 8 : 1;
    const int64_t total = ((int64_t)(n_chunks - 1) * chunk_tokens + last_chunk_tokens) * 2 * kv->L * kv->H * (kv->D / V);
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = 148ll * 8 * 4;   // a few waves of 148 SMs x 8 CTAs; grid-stride covers the rest
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (pack) {
        if (vec) pack_kernel<8, true><<<(unsigned)blocks, 256, 0, stream>>>(P);
        else pack_kernel<1, true><<<(unsigned)blocks, 256, 0, stream>>>(P);
    } else {
        if (vec) pack_kernel<8, false><<<(unsigned)blocks, 256, 0, stream>>>(P);
        else pack_kernel<1, false><<<(unsigned)blocks, 256, 0, stream>>>(P);
    }
    B2_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b200kv

using namespace b200kv;

extern "C" {

int b200kv_pack_chunks(const b200kv_kv_desc* src, int64_t tok_begin, int32_t n_chunks, int32_t chunk_tokens,
                       int32_t last_chunk_tokens, int32_t hf_layout, void* chunks, int64_t chunk_stride_bytes,
                       void* stream) {
    return launch_pack(true, src, tok_begin, n_chunks, chunk_tokens, last_chunk_tokens, hf_layout, chunks,
                       chunk_stride_bytes, static_cast<cudaStream_t>(stream));
}

int b200kv_unpack_chunks(const void* chunks, int64_t chunk_stride_bytes, int32_t n_chunks, int32_t chunk_tokens,
                         int32_t last_chunk_tokens, int32_t hf_layout, const b200kv_kv_desc* dst, int64_t tok_begin,
                         void* stream) {
    return launch_pack(false, dst, tok_begin, n_chunks, chunk_tokens, last_chunk_tokens, hf_layout,
                       const_cast<void*>(chunks), chunk_stride_bytes, static_cast<cudaStream_t>(stream));
}

int b200kv_pinned_alloc(void** host_ptr, int64_t bytes) {
    B2_REQUIRE(host_ptr != nullptr && bytes > 0, "bad pinned_alloc arguments");
    B2_CHECK_CUDA(cudaHostAlloc(host_ptr, (size_t)bytes, cudaHostAllocPortable | cudaHostAllocMapped));
    return 0;
}

int b200kv_pinned_free(void* host_ptr) {
    if (host_ptr) B2_CHECK_CUDA(cudaFreeHost(host_ptr));
    return 0;
}

int b200kv_host_device_ptr(void* host_ptr, void** device_ptr) {
    B2_REQUIRE(host_ptr != nullptr && device_ptr != nullptr, "NULL pointer");
    B2_CHECK_CUDA(cudaHostGetDevicePointer(device_ptr, host_ptr, 0));
    return 0;
}

int b200kv_copy_async(void* dst, const void* src, int64_t bytes, void* stream) {
    B2_REQUIRE(dst != nullptr && src != nullptr && bytes >= 0, "bad copy arguments");
    if (bytes == 0) return 0;
    B2_CHECK_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream)));
    return 0;
}

int b200kv_copy2d_async(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t row_bytes,
                        int64_t rows, void* stream) {
    B2_REQUIRE(dst != nullptr && src != nullptr && row_bytes >= 0 && rows >= 0, "bad copy2d arguments");
    if (row_bytes == 0 || rows == 0) return 0;
    B2_CHECK_CUDA(cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src, (size_t)src_pitch, (size_t)row_bytes, (size_t)rows,
                                    cudaMemcpyDefault, static_cast<cudaStream_t>(stream)));
    return 0;
}

int b200kv_stream_create(void** stream) {
    B2_REQUIRE(stream != nullptr, "NULL pointer");
    cudaStream_t s;
    B2_CHECK_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = s;
    return 0;
}
int b200kv_stream_destroy(void* stream) {
    if (stream) B2_CHECK_CUDA(cudaStreamDestroy(static_cast<cudaStream_t>(stream)));
    return 0;
}
int b200kv_stream_sync(void* stream) {
    B2_CHECK_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    return 0;
}
int b200kv_event_create(void** event) {
    B2_REQUIRE(event != nullptr, "NULL pointer");
    cudaEvent_t e;
    B2_CHECK_CUDA(cudaEventCreate(&e));
    *event = e;
    return 0;
}
int b200kv_event_destroy(void* event) {
    if (event) B2_CHECK_CUDA(cudaEventDestroy(static_cast<cudaEvent_t>(event)));
    return 0;
}
int b200kv_event_record(void* event, void* stream) {
    B2_CHECK_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(event), static_cast<cudaStream_t>(stream)));
    return 0;
}
int b200kv_event_query(void* event) {
    cudaError_t e = cudaEventQuery(static_cast<cudaEvent_t>(event));
    if (e == cudaSuccess) return 0;
    if (e == cudaErrorNotReady) return 1;
    set_error(std::string("cudaEventQuery: ") + cudaGetErrorString(e));
    return -1;
}
int b200kv_event_sync(void* event) {
    B2_CHECK_CUDA(cudaEventSynchronize(static_cast<cudaEvent_t>(event)));
    return 0;
}
int b200kv_stream_wait_event(void* stream, void* event) {
    B2_CHECK_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), static_cast<cudaEvent_t>(event), 0));
    return 0;
}
int b200kv_event_elapsed_ms(void* start, void* stop, float* ms) {
    B2_REQUIRE(ms != nullptr, "NULL pointer");
    B2_CHECK_CUDA(cudaEventElapsedTime(ms, static_cast<cudaEvent_t>(start), static_cast<cudaEvent_t>(stop)));
    return 0;
}

int b200kv_version(void) { return B200KV_VERSION; }

int b200kv_device_count(void) {
    int n = 0;
    B2_CHECK_CUDA(cudaGetDeviceCount(&n));
    return n;
}

}  // extern "C"

namespace b200kv {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
}  // namespace b200kv

extern "C" const char* b200kv_last_error(void) { return b200kv::g_last_error.c_str(); }
