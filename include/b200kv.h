/*
 * b200kv.h -- C ABI of libb200kv.so, the B200 (sm_100a) KV-cache store/load hot path.
 *
 * This is the drop-in boundary for ONE path of LMCache v0.1.2 (paths below are relative to the
 * reference tree): CacheGen encode / decode, the chunked token-id SHA-256 prefix hash, and the
 * GPU <-> pinned-host mover.  Each entry point names the reference interface it replaces.
 *
 * Conventions
 *   - every function is extern "C", returns int: 0 = ok, <0 = error (b200kv_last_error() gives the
 *     message for the calling thread).  No torch / pybind types: raw device/host pointers, sizes,
 *     element strides and a cudaStream_t passed as void*.
 *   - the caller owns every buffer.  Work is enqueued on `stream` and is asynchronous unless the
 *     function says otherwise.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 *
 * KV layout: element (l, kv, tok, h, d) of the chunk/blob lives at
 *       base + (l*sL + kv*sKV + tok*sT + h*sH + d) * 2 bytes            (d is contiguous)
 *   or, when `planes` is non-NULL, at planes[kv*L + l] + (tok*sT + h*sH + d) * 2 bytes.
 *   vllm blob [L,2,T,H,D]: sL=2*T*H*D sKV=T*H*D sT=H*D sH=D ; huggingface [L,2,H,T,D]: sT=D sH=T*D.
 *   The planes form takes the 2L tensors of the engine's kv tuple as they are
 *   (replaces the stack/stack/stack/permute + split/.contiguous() copies of
 *   lmcache/cache_engine.py:98-118,131-161).
 */
#ifndef B200KV_H_
#define B200KV_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200KV_VERSION 4           /* ABI version; 2: b200kv_kv_desc.slot_map; 3: coder selection, decode status, total_bytes;
                                    * 4: compact container (B200KV_CODER_RANS_COMPACT), b200kv_container_layout_v,
                                    *    b200kv_sha256_chain_ready */
#define B200KV_CODER_AC 0          /* payload = torchac-lineage arithmetic coder; container version 1 */
#define B200KV_CODER_RANS 1        /* payload = rANS, 32-bit state / 16-bit renormalisation; container version 2 */
#define B200KV_CODER_RANS_COMPACT 2 /* rANS as in version 2, compact side information; container version 3 (chunks of
                                     * <= 256 tokens): the per-stream CDF row (66 B) is replaced by the symbol histogram it
                                     * is a function of, carried sparsely in front of the stream's rANS bytes (~5 B at the
                                     * headline entropy), the int32 stream length by one byte */
#define B200KV_CONTAINER_VERSION(coder) ((coder) + 1) /* "B2KV" wire container version (b200kv_header.version) */
#define B200KV_ENCODE_HINT_HIGH_ENTROPY 0x100 /* OR into `coder` of b200kv_encode_chunks: the caller expects more than ~2.7
                                               * payload bits per symbol (e.g. the previous call's sizes said so); selects the
                                               * TMA-staged encode kernel whose time does not grow with entropy.  Output bytes
                                               * are identical either way. */
#define B200KV_ENCODE_HINT_MID_ENTROPY 0x200  /* likewise: more than ~1.2 payload bits per symbol expected -- the compaction
                                               * kernel then keeps its full-size shared-memory stage (tiles of 10+ KB) */
#define B200KV_LP 33            /* CDF entries per stream (cachegen_encoder.py:287-289: int(bins.max()) + 1) */
#define B200KV_GROUP_TOKENS 256 /* CACHEGEN_GPU_MAX_TOKENS_PER_CHUNK (cachegen_basics.py:13) */
#define B200KV_MAX_PLANES 128   /* 2 * nlayers upper bound */
#define B200KV_DT_BF16 0
#define B200KV_DT_FP16 1

#define B200KV_READ_SLACK 640   /* bytes that must be readable past the end of every container handed to the decoder */
#define B200KV_MAGIC 0x564B3242u /* "B2KV" little-endian */
#define B200KV_HEADER_BYTES 64

typedef struct b200kv_kv_desc {
    const void* base;          /* used when planes == NULL */
    const void* const* planes; /* HOST array of 2L device pointers, index kv*L + l; or NULL */
    int64_t sL, sKV, sT, sH;   /* element strides (sL/sKV ignored when planes != NULL) */
    int32_t L, H, D;           /* layers, kv heads, head size; channels C = H*D */
    int32_t dtype;             /* B200KV_DT_* of the KV elements */
    const int64_t* slot_map;   /* DEVICE array or NULL.  Paged KV (vLLM's slot_mapping, lmcache-vllm's
                                * lmcache_store_kv / lmcache_retrieve_kv, docs LLM_Engine.rst:91-109): token i of the call
                                * (tok_begin + i for encode, dst_tok[j] + i for decode) lives in row slot_map[i] of every
                                * plane, i.e. `tok` in the address formula above is replaced by slot_map[tok].  The codec
                                * and pack / unpack kernels gather / scatter through it, so the paged cache is read and
                                * written in place. */
} b200kv_kv_desc;

/* Wire container of one encoded chunk ("B2KV" v1 / v2 / v3).  All sections 16-byte aligned, little-endian.
 * Logical content == CacheGenGPUEncoderOutput (cachegen_basics.py:109-142):
 *   cdf [2L,C,33] int16 | max_tensors_key/value [2,L,t] half | per <=256-token group:
 *   bytestream_lengths [2L,C] int32 + bytestream (streams in (nl,c) row-major order, no padding).
 * Flat instead of pickled CUDA tensors so it can be produced on the device in one buffer and moved
 * with one async copy.
 * Versions 1 and 2 differ ONLY in the bytes of each stream inside `bytestream`: the reference's coder lives in the
 * un-vendored torchac_cuda wheel, so the bitstream is this build's own.  v1 = 32-bit binary arithmetic
 * coder (torchac lineage, MSB-first bits); v2 = rANS over the same 16-bit CDFs: LE32 final state, then LE16
 * renormalisation words in decode order (normative description: lmcache_b200/csrc/ac_core.cuh).
 * Version 3 (one <= 256-token group per container) has no CDF section:
 *   header | nb u8[2L] (pad to 16) | maxes | half-lengths u8[2L][C] | bytestream
 * nb(plane) = 2 * (bins(plane) // 2) = the symbols a plane can emit (16 or 32 for the reference's bin tables).  Every
 * stream of the bytestream is  [mask: ceil(nb / 8) bytes LE, bit s set <=> symbol s occurs in the stream]
 *   [one count byte per set bit, ascending, except the LAST set bit, whose count is ntokens - (sum of the others)]
 *   [a zero byte if that makes the length even] [the version-2 rANS stream];  half-lengths[c] = bytes of all that / 2.
 * The CDF the reference keeps (cachegen_encoder.py:287-290) is a function of these counts n_s and t = ntokens:
 *   cdf[i] = int16(rint(fl32(sum_{k<i} fl32(n_k / t), accumulated in double) * 65504) + i)      (in-tree spec :95-126)
 * so CacheGenGPUEncoderOutput.from_bytes rebuilds the identical tensor, and the decoder evaluates it on the device. */
typedef struct b200kv_header {
    uint32_t magic, version;
    uint32_t L, H, D;
    uint32_t ntokens, ngroups;
    uint32_t max_dtype;        /* dtype of the max tensors (== input dtype) */
    uint64_t payload_bytes;    /* all groups' bytestreams, concatenated in group order */
    uint64_t total_bytes;      /* header + sections + payload */
    uint32_t status;           /* 0 = ok; nonzero = encoder detected an internal overflow */
    uint32_t reserved[3];
} b200kv_header;

/* Section offsets of a container with the given shape (pure arithmetic, host side). */
typedef struct b200kv_layout {
    int64_t off_cdf, off_maxes, off_lengths, off_payload;
    int64_t fixed_bytes;       /* == off_payload */
    int64_t max_total_bytes;   /* worst case: payload at 2 bytes/symbol + flush */
} b200kv_layout;

int b200kv_version(void);
const char* b200kv_last_error(void);
/* number of CUDA devices visible, or <0 with an error: lets callers fail loudly up front */
int b200kv_device_count(void);

int b200kv_container_layout(int32_t L, int32_t H, int32_t D, int32_t ntokens, b200kv_layout* out);   /* versions 1, 2 */
/* Same for the container `coder` produces (B200KV_CODER_RANS_COMPACT: off_cdf is the nb map, there is no CDF section). */
int b200kv_container_layout_v(int32_t L, int32_t H, int32_t D, int32_t ntokens, int32_t coder, b200kv_layout* out);

/* Bytes of device scratch b200kv_encode_chunks / b200kv_decode_chunks need for a call. */
int64_t b200kv_encode_workspace_bytes(int32_t L, int32_t H, int32_t D, int32_t chunk_tokens, int32_t n_chunks,
                                      int32_t coder);
int64_t b200kv_decode_workspace_bytes(int32_t L, int32_t H, int32_t D, int32_t chunk_tokens, int32_t n_chunks);

/*
 * CacheGen encode.  Replaces, for n_chunks consecutive chunks in ONE call:
 *   torch_quant_vectorized x2            cachegen_encoder.py:40-61,282-285
 *   torchac_cuda.calculate_cdf x2        cachegen_encoder.py:287-290   (in-tree spec :95-126,185-196)
 *   torchac_cuda.encode_fast_new + collect_bytes per <=256-token group   :225-262,301-316
 *   CacheGenGPUEncoderOutput.to_bytes    cachegen_basics.py:131-136   (container written on device)
 * i.e. the body of CacheGenSerializer.to_bytes (cachegen_encoder.py:353-389).
 *
 * Chunk j covers tokens [tok_begin + j*chunk_tokens, ...) and holds chunk_tokens tokens except the
 * last one, which holds last_chunk_tokens (1..chunk_tokens).  Its container is written at
 * out + j*out_stride (device memory, out_stride >= layout.max_total_bytes or the call fails with
 * the header status set if the payload does not fit).  sizes_out[j] (device or mapped-host memory)
 * receives total_bytes of chunk j, or 0 when the chunk's header carries a nonzero status.
 * key_bins / value_bins: HOST float arrays of length L (make_key_bins / make_value_bins, :339-350).
 * coder: B200KV_CODER_* -- which entropy coder fills the bytestreams (and hence header.version).
 */
int b200kv_encode_chunks(const b200kv_kv_desc* kv, int64_t tok_begin, int32_t n_chunks, int32_t chunk_tokens,
                         int32_t last_chunk_tokens, const float* key_bins, const float* value_bins, int32_t coder,
                         void* out, int64_t out_stride, uint64_t* sizes_out, void* workspace, int64_t workspace_bytes,
                         void* stream);

/*
 * CacheGen decode.  Replaces, for n_chunks containers in ONE call:
 *   CacheGenGPUEncoderOutput.from_bytes  cachegen_basics.py:138-142   (parsed on the host by caller)
 *   decode_chunk cumsum + torchac_cuda.decode_fast_prefsum            cachegen_decoder.py:52-66
 *   decode_function_gpu / .float()                                     :70-106
 *   do_dequantize x2 + stack/reshape/permute/.to(bf16|fp16)            :24-35,177-200
 * i.e. the body of CacheGenDeserializer.from_bytes (:143-202).
 *
 * containers: DEVICE buffer; container j starts at containers + offsets[j] (HOST int64 array, 16-byte
 * aligned offsets).  Shapes are passed by the caller (it has parsed the headers): all chunks share
 * L/H/D; ntokens[j] (HOST int32 array) tokens each.  Chunk j's tokens are written to `dst` at token
 * index dst_tok[j] (HOST int64 array) using dst's strides; dst->dtype is the output dtype
 * (bf16 for vllm, fp16 for huggingface, cachegen_decoder.py:189-200) and max_dtype the dtype of the
 * stored max tensors.  total_bytes[j] (HOST int64 array) = header.total_bytes of container j and containers_bytes = the
 * size of the `containers` buffer; the call fails unless offsets[j] + total_bytes[j] + B200KV_READ_SLACK <=
 * containers_bytes for every j (the slack bytes may hold anything, e.g. the next container).  Given that, the kernels
 * never read outside the buffer whatever the lengths section says: stream starts are clamped to the payload and a
 * stream reads a bounded number of bytes from its start (a corrupt or truncated blob from a remote tier is a cache
 * miss, not an illegal address).
 * coder = header.version - 1 of the containers (one call, one coder).
 * status_out: NULL, or DEVICE / mapped-host uint32[n_chunks], zeroed by the call and then OR-ed with
 *   1 = a rANS stream did not return to its initial state (payload or CDF bytes damaged),
 *   2 = stream offsets beyond the payload (lengths section damaged); valid once the stream has run.
 */
int b200kv_decode_chunks(const void* containers, int64_t containers_bytes, const int64_t* offsets,
                         const int64_t* total_bytes, const int32_t* ntokens, const int64_t* dst_tok, int32_t n_chunks,
                         int32_t max_dtype,
                         int32_t coder, const b200kv_kv_desc* dst, const float* key_bins, const float* value_bins,
                         uint32_t* status_out, void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Token-id prefix hash.  Replaces LMCacheEngine._chunk_tokens/_hash/_prefix_hash
 * (cache_engine.py:58-96): h_i = sha256(ascii_hex(h_{i-1}) || bytes(tokens[i*cs:(i+1)*cs])), h_{-1} = "".
 * n_seq independent sequences are hashed concurrently (one chain each).  tokens: DEVICE pointer to the
 * raw little-endian token ids (elem_size bytes each: 8 for int64, 4 for int32); sequence s covers
 * tokens [seq_offsets[s], seq_offsets[s+1]) (HOST int64 array of n_seq+1 entries).  digests: DEVICE
 * or mapped-host buffer, 32 raw bytes per chunk, sequences back to back, ceil(len/chunk_size) each.
 */
int b200kv_sha256_chain(const void* tokens, int32_t elem_size, const int64_t* seq_offsets, int32_t n_seq,
                        int32_t chunk_size, void* digests, void* stream);
/* Same, and digest k is announced as soon as it exists: ready[k] (DEVICE-visible uint32 array, one word per digest slot,
 * normally mapped host memory like `digests`; or NULL) is set to `epoch` after digest k has been made visible system-wide.
 * A chain is serial -- 38 us per 256-token chunk -- so a host thread that polls ready[] can look up and move the first
 * chunks while the later ones are still being hashed (LMCacheEngine.store / retrieve do).  The digests must be 4-byte
 * aligned. */
int b200kv_sha256_chain_ready(const void* tokens, int32_t elem_size, const int64_t* seq_offsets, int32_t n_seq,
                              int32_t chunk_size, void* digests, uint32_t* ready, uint32_t epoch, void* stream);

/*
 * Blob pack / unpack between a kv_desc (tuple-of-tensors or strided blob) and contiguous chunk blobs
 * [L,2,t,H,D] (vllm) / [L,2,H,t,D] (huggingface), t = chunk tokens.  Replaces _tuple_kv_to_blob +
 * _slice_kv_at (.contiguous() per chunk) and the retrieve-side torch.cat
 * (cache_engine.py:98-161,362-368) with one gather / scatter kernel.
 * `chunks` is DEVICE memory, or pinned host memory mapped into the device address space (then the kernel
 * itself is the GPU->host mover).  Chunk j starts at chunks + j*chunk_stride_bytes.
 * hf_layout != 0 selects the huggingface chunk layout.
 */
int b200kv_pack_chunks(const b200kv_kv_desc* src, int64_t tok_begin, int32_t n_chunks, int32_t chunk_tokens,
                       int32_t last_chunk_tokens, int32_t hf_layout, void* chunks, int64_t chunk_stride_bytes,
                       void* stream);
int b200kv_unpack_chunks(const void* chunks, int64_t chunk_stride_bytes, int32_t n_chunks, int32_t chunk_tokens,
                         int32_t last_chunk_tokens, int32_t hf_layout, const b200kv_kv_desc* dst, int64_t tok_begin,
                         void* stream);

/*
 * GPU <-> pinned-host mover.  Replaces LMCLocalBackend.put_blocking/put_nonblocking/get
 * (local_backend.py:82-100,128-144: pageable tensor.to("cpu") / .to("cuda") + torch.cuda.synchronize()).
 */
int b200kv_pinned_alloc(void** host_ptr, int64_t bytes);       /* cudaHostAlloc(portable|mapped) */
int b200kv_pinned_free(void* host_ptr);
int b200kv_host_device_ptr(void* host_ptr, void** device_ptr); /* device alias of a pinned allocation */
int b200kv_copy_async(void* dst, const void* src, int64_t bytes, void* stream);  /* cudaMemcpyDefault */
/* strided 2-D copy: `rows` rows of `row_bytes`, pitches in bytes (chunk slice of a [L,2,T,H,D] blob) */
int b200kv_copy2d_async(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t row_bytes,
                        int64_t rows, void* stream);
int b200kv_stream_create(void** stream);                       /* non-blocking side stream */
int b200kv_stream_destroy(void* stream);
int b200kv_stream_sync(void* stream);
int b200kv_event_create(void** event);
int b200kv_event_destroy(void* event);
int b200kv_event_record(void* event, void* stream);
int b200kv_event_query(void* event);                           /* 0 = complete, 1 = pending, <0 = error */
int b200kv_event_sync(void* event);
int b200kv_stream_wait_event(void* stream, void* event);
int b200kv_event_elapsed_ms(void* start, void* stop, float* ms);

/*
 * Per-kernel device timing of the most recent encode / decode call (CUDA events recorded on the call's
 * stream around each launch).  Not part of the reference surface: bench.py's roofline leg uses it so the
 * dominant kernel is timed live, outside any profiler.  Slots: 0 absmax, 1 cdf, 2 encode, 3 scan+compact+finalize,
 * 4 tile_sum, 5 tile_scan, 6 decode; -1 = not launched.  Not thread-safe; enable only while benchmarking.
 */
int b200kv_profile_enable(int32_t on);
int b200kv_profile_last(float* ms, int32_t n);

/*
 * lm:// remote tier, client and server.  Host-only: no device work, usable without a GPU.
 * Wire-compatible with lmcache/protocol.py:4-70 (158-byte client header "ii150s", 8-byte server header "ii"), so either
 * side interoperates with the reference's Python client (storage_backend/connector/lm_connector.py:15-84) and server
 * (lmcache/server/__main__.py:29-104).  Payloads are sent from / received into caller memory in one pass (Python bytes,
 * a pinned slab, ...); a connection serialises whole request / response exchanges; the server's EXIST / GET are O(1)
 * hash lookups under a reader-writer lock (the reference scans list_keys()).  Keys are <= 150 bytes.
 */
int b200kv_lm_server_start(const char* host, int32_t port, void** server); /* port 0 = ephemeral; threads run until stop */
int32_t b200kv_lm_server_port(void* server);
int64_t b200kv_lm_server_num_keys(void* server);
int b200kv_lm_server_stop(void* server);
int b200kv_lm_connect(const char* host, int32_t port, void** conn);
int b200kv_lm_close(void* conn);
int b200kv_lm_put(void* conn, const char* key, const void* data, int64_t len);      /* connection.set(); no server ack */
int b200kv_lm_exists(void* conn, const char* key);                                  /* 1 present, 0 absent, <0 error */
/* GET / LIST in two calls, because the caller allocates the destination once the length is known:
 *   n = b200kv_lm_get_begin(conn, key)   payload length >= 0, -1 = miss, < -1 = error
 *   b200kv_lm_read(conn, dst, n)         payload into caller memory (must follow a successful begin, also for n = 0) */
int64_t b200kv_lm_get_begin(void* conn, const char* key);
int64_t b200kv_lm_list_begin(void* conn);                                           /* keys joined by '\n' */
int b200kv_lm_read(void* conn, void* dst, int64_t len);

#ifdef __cplusplus
}
#endif
#endif /* B200KV_H_ */
