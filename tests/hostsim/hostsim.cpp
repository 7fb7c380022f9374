// tests/hostsim/hostsim.cpp -- TEST INFRASTRUCTURE.  Compiles the product's device arithmetic
// (lmcache_b200/csrc/ac_core.cuh, all __host__ __device__) with g++ so the exact same functions the
// CUDA kernels call can be checked against the oracle on a machine without a GPU.  Never shipped,
// never loaded by lmcache_b200/.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../lmcache_b200/csrc/ac_core.cuh"

using namespace b200kv;

namespace {
struct VecSink {
    std::vector<uint8_t>* v;
    void put_word(uint32_t w) {
        v->push_back((uint8_t)(w >> 24)); v->push_back((uint8_t)(w >> 16));
        v->push_back((uint8_t)(w >> 8)); v->push_back((uint8_t)w);
    }
};
struct BufSrc {
    const uint8_t* p; int64_t n, pos;
    uint32_t next_word() {
        uint32_t w = 0;
        for (int i = 0; i < 4; ++i) { w <<= 8; if (pos < n) w |= p[pos]; ++pos; }
        return w;
    }
};
}  // namespace

extern "C" {

// encode `g` symbols (stride sym_stride) with cdf[33] (uint16); returns byte length written to out (cap bytes)
int64_t sim_encode_stream(const uint16_t* cdf, const int8_t* sym, int64_t sym_stride, int g, uint8_t* out, int64_t cap) {
    std::vector<uint8_t> v;
    VecSink sink{&v};
    EncState st; st.init();
    for (int i = 0; i < g; ++i) {
        int s = sym[i * sym_stride];
        uint32_t c_lo = cdf[s];
        uint32_t c_hi = (s == kMaxSym) ? 0x10000u : cdf[s + 1];
        enc_symbol(st, c_lo, c_hi - c_lo, sink);
    }
    uint32_t nb = enc_finish(st, sink);
    if (nb) {
        uint32_t w = (uint32_t)(st.acc << (32u - nb));
        for (uint32_t i = 0; i < (nb + 7u) / 8u; ++i) v.push_back((uint8_t)(w >> (24 - 8 * i)));
    }
    int64_t n = (int64_t)v.size();
    if (n <= cap) memcpy(out, v.data(), (size_t)n);
    return n;
}

void sim_decode_stream(const uint16_t* cdf, const uint8_t* in, int64_t n, int g, uint8_t* out, int64_t out_stride) {
    BufSrc src{in, n, 0};
    DecState st; dec_init(st, src);
    for (int i = 0; i < g; ++i)
        out[i * out_stride] = (uint8_t)dec_symbol(st, src, [&](uint32_t k) { return (uint32_t)cdf[k]; }, i == g - 1);
}

// production (branch-light) coder paths -- the functions the kernels actually call
int64_t sim_encode_stream2(const uint16_t* cdf, const int8_t* sym, int64_t sym_stride, int g, uint8_t* out, int64_t cap) {
    const uint32_t capw = (uint32_t)(2 * g + 16) / 4 + 2;
    std::vector<uint32_t> row(capw, 0u);
    EncState2 st; st.init();
    for (int i = 0; i < g; ++i) {
        int s = sym[i * sym_stride];
        uint32_t c_lo = cdf[s];
        uint32_t c_hi = (s == kMaxSym) ? 0x10000u : cdf[s + 1];
        enc_symbol2(st, c_lo, c_hi - c_lo, row.data(), capw);
    }
    uint32_t n = enc_finish2(st, row.data(), capw);
    if ((int64_t)n <= cap)
        for (uint32_t i = 0; i < n; ++i) out[i] = (uint8_t)(row[i >> 2] >> (24u - 8u * (i & 3u)));   // MSB-first words
    return n;
}

namespace {
struct WordSrc {   // aligned big-endian word reader over a byte buffer that may start mid-word
    const uint8_t* base; int64_t pos;   // pos: byte index of the next aligned word (may be negative offset handled by caller)
    const uint8_t* buf; int64_t n;
    uint32_t next_be() {
        uint32_t w = 0;
        for (int i = 0; i < 4; ++i) { int64_t p = pos + i; w = (w << 8) | ((p >= 0 && p < n) ? buf[p] : 0xA5u); }  // garbage past the end: must not matter
        pos += 4;
        return w;
    }
};
}

// `skip` leading bytes (0..3) precede the stream inside its first aligned word (they hold foreign data)
void sim_decode_stream2(const uint16_t* cdf, const uint8_t* in, int64_t n, int g, uint8_t* out, int64_t out_stride, int skip, int nsteps) {
    WordSrc src{nullptr, -(int64_t)skip, in, n};
    DecState2 st; dec_init2(st, src, (uint32_t)skip);
    uint32_t e[kLp];
    for (uint32_t i = 0; i < (uint32_t)kLp; ++i) e[i] = dec_table_entry(i, cdf[i]);
    for (int i = 0; i < g; ++i)
        out[i * out_stride] = (uint8_t)((nsteps == 4 ? dec_symbol2<4>(st, src, e, i == g - 1) : dec_symbol2<5>(st, src, e, i == g - 1)) >> 2);   // returns 4 * symbol
}


// ---- rANS (container version 2): the product's rans_enc_symbol / rans_dec_symbol
int64_t sim_rans_encode_stream(const uint16_t* cdf, const int8_t* sym, int64_t sym_stride, int g, uint8_t* out, int64_t cap) {
    std::vector<uint16_t> stack;
    uint32_t x = kRansLow;
    for (int i = g - 1; i >= 0; --i) {
        int s = sym[i * sym_stride];
        uint32_t c_lo = cdf[s];
        uint32_t c_hi = (s == kMaxSym) ? 0x10000u : cdf[s + 1];
        rans_enc_symbol(x, c_lo, c_hi - c_lo, [&](uint32_t h) { stack.push_back((uint16_t)h); });
    }
    int64_t n = 4 + 2 * (int64_t)stack.size();
    if (n <= cap) {
        for (int i = 0; i < 4; ++i) out[i] = (uint8_t)(x >> (8 * i));
        for (size_t j = 0; j < stack.size(); ++j) {
            uint16_t h = stack[stack.size() - 1 - j];
            out[4 + 2 * j] = (uint8_t)h; out[5 + 2 * j] = (uint8_t)(h >> 8);
        }
    }
    return n;
}

namespace {
struct LeWordSrc {   // aligned little-endian word reader; the stream may start at the upper halfword of its first word
    int64_t pos; const uint8_t* buf; int64_t n;
    uint32_t next_le() {
        uint32_t w = 0;
        for (int i = 0; i < 4; ++i) { int64_t p = pos + i; w |= (uint32_t)((p >= 0 && p < n) ? buf[p] : 0xA5u) << (8 * i); }  // garbage outside: must not matter
        pos += 4;
        return w;
    }
};
}

// odd = 1: two foreign bytes precede the stream inside its first aligned word.  Returns the final state (2^16 when intact).
uint32_t sim_rans_decode_stream(const uint16_t* cdf, const uint8_t* in, int64_t n, int g, uint8_t* out, int64_t out_stride, int odd, int nsteps) {
    LeWordSrc src{-(int64_t)(2 * odd), in, n};
    RansDec st; rans_dec_init(st, src, (uint32_t)odd);
    uint32_t pk[32];
    for (uint32_t i = 0; i < 32u; ++i) pk[i] = rans_table_entry(cdf[i], i == 31u ? 0x10000u : cdf[i + 1]);
    for (int i = 0; i < g; ++i)
        out[i * out_stride] = (uint8_t)(nsteps == 4 ? rans_dec_symbol<4>(st, src, pk) : rans_dec_symbol<5>(st, src, pk));
    return st.x;
}

uint32_t sim_rans_divmod(uint32_t x, uint32_t f, uint32_t* rem) { return rans_divmod(x, f, rem); }

void sim_cdf(const uint32_t* counts, int t, uint16_t* cdf) {
    CdfAccum a; a.init(t);
    for (uint32_t i = 0; i < (uint32_t)kLp; ++i) cdf[i] = a.next(i, i < 33 ? counts[i] : 0);
}

// quantise one row of C halfs given the row max (half bits)
void sim_quant_row(const uint16_t* x, int dtype, int C, uint16_t max_bits, float maxq, uint8_t* sym) {
    float f = quant_factor(maxq, half_to_float(max_bits, dtype));
    for (int c = 0; c < C; ++c) sym[c] = (uint8_t)quant_symbol(half_to_float(x[c], dtype), f, maxq);
}

// the fused encode kernel's pass 1: "safe" factor + unchecked symbol (host build of quant_symbol_nc keeps the check, so
// this pins quant_factor_safe: an infinite factor must become NaN and nothing else may change)
void sim_quant_row_safe(const uint16_t* x, int dtype, int C, uint16_t max_bits, float maxq, uint8_t* sym, uint32_t* factor_bits) {
    float f = quant_factor_safe(maxq, half_to_float(max_bits, dtype));
    *factor_bits = f2u(f);
    for (int c = 0; c < C; ++c) sym[c] = (uint8_t)quant_symbol_nc(half_to_float(x[c], dtype), f, maxq);
}

void sim_dequant_row(const uint8_t* sym, int C, uint16_t max_bits, int max_dtype, float cq, int out_dtype, uint16_t* out) {
    float m = half_to_float(max_bits, max_dtype);
    for (int c = 0; c < C; ++c) out[c] = float_to_half(dequant_value(dequant_lut(sym[c], cq), m), out_dtype);
}

void sim_half_to_float(const uint16_t* h, int n, int dtype, float* out) { for (int i = 0; i < n; ++i) out[i] = half_to_float(h[i], dtype); }
void sim_float_to_half(const float* f, int n, int dtype, uint16_t* out) { for (int i = 0; i < n; ++i) out[i] = float_to_half(f[i], dtype); }

void sim_layout(int L, int C, int t, int compact, int64_t* out) {
    Layout lo = make_layout(L, C, t, compact);
    out[0] = lo.off_cdf; out[1] = lo.off_maxes; out[2] = lo.off_lengths; out[3] = lo.off_payload; out[4] = lo.ngroups;
}

// CdfAccum2 (value / absorb, symbols in `skip` are not absorbed -- legal when their count is 0) vs CdfAccum
void sim_cdf_skip(const uint32_t* counts, int t, uint32_t skip, uint16_t* cdf) {
    CdfAccum2 a; a.init();
    for (uint32_t i = 0; i < (uint32_t)kLp; ++i) {
        cdf[i] = (uint16_t)a.value(i);
        if (i < 32 && !((skip >> i) & 1u)) a.absorb(fdiv((float)counts[i], (float)t));
    }
}

// version-3 stream header as ac_core.cuh specifies it (hdr_write_host / hdr_len share the arithmetic with the kernels)
int sim_hdr_write(const uint32_t* cnt, int nb, uint8_t* out) { return (int)hdr_write_host(out, cnt, nb); }
int sim_hdr_len(uint32_t mask, int nb) { return (int)hdr_len(mask, nb); }
}
