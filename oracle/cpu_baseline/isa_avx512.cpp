// AVX-512 (64 x u8) instantiation — mirrors BackendAVX512U8 (src/smith_waterman/backend/avx512.rs:105-116)
// and PrefilterAVX512Backend (src/prefilter/backend/avx512.rs).  Compiled with
// -mavx512f -mavx512bw -mavx512vbmi -mbmi -mbmi2 -mlzcnt; only called after a cpuid check.
#include <immintrin.h>
#include "../../include/frz_cuda.h"
#include "kernels.inl"
#include "driver.h"

struct V512 {
    static constexpr int LANES = 64;
    using vec = __m512i;
    using msk = __mmask64;
    using bits = uint64_t;
    static inline vec zero() { return _mm512_setzero_si512(); }
    static inline msk mzero() { return 0; }
    static inline vec splat(int c) { return _mm512_set1_epi8((char)c); }
    static inline vec first_lane(uint8_t v) { return _mm512_maskz_set1_epi8(1, (char)v); }
    static inline bits all_bits() { return ~0ull; }
    static inline vec load_partial(const uint8_t* p, size_t rem) {
        return rem >= 64 ? _mm512_loadu_si512(p) : _mm512_maskz_loadu_epi8(rem ? ((~0ull) >> (64 - rem)) : 0, p);
    }
    static inline bits eq_bits(vec a, vec b) { return _mm512_cmpeq_epi8_mask(a, b); }
    static inline int tz(bits m) { return (int)_tzcnt_u64(m); }
    static inline int lz(bits m) { return (int)_lzcnt_u64(m); }
    static inline msk eq(vec a, vec b) { return _mm512_cmpeq_epi8_mask(a, b); }
    static inline msk gt(vec a, vec b) { return _mm512_cmpgt_epu8_mask(a, b); }
    static inline msk lt(vec a, vec b) { return _mm512_cmplt_epu8_mask(a, b); }
    static inline msk mand(msk a, msk b) { return a & b; }
    static inline msk mor(msk a, msk b) { return a | b; }
    static inline msk mnot(msk a) { return ~a; }
    static inline msk mshift1(msk m, msk prev) { return (m << 1) | (prev >> 63); }
    static inline vec widen(msk m) { return _mm512_movm_epi8(m); }
    static inline vec add(vec a, vec b) { return _mm512_add_epi8(a, b); }
    static inline vec subs(vec a, vec b) { return _mm512_subs_epu8(a, b); }
    static inline vec max(vec a, vec b) { return _mm512_max_epu8(a, b); }
    static inline vec band(vec a, vec b) { return _mm512_and_si512(a, b); }
    template <int S>
    static inline vec srp(vec v, vec prev) {
        // out[i] = i < S ? prev[64 - S + i] : v[i - S]   (one vpermt2b)
        alignas(64) static const struct Idx { uint8_t b[64]; Idx() { for (int i = 0; i < 64; i++) b[i] = (uint8_t)(i < S ? (64 - S + i) : (64 + i - S)); } } idx;
        return _mm512_permutex2var_epi8(prev, _mm512_load_si512(idx.b), v);
    }
    static inline unsigned hmax(vec v) {
        __m256i a = _mm256_max_epu8(_mm512_castsi512_si256(v), _mm512_extracti64x4_epi64(v, 1));
        __m128i b = _mm_max_epu8(_mm256_castsi256_si128(a), _mm256_extracti128_si256(a, 1));
        b = _mm_max_epu8(b, _mm_srli_si128(b, 8));
        b = _mm_max_epu8(b, _mm_srli_si128(b, 4));
        b = _mm_max_epu8(b, _mm_srli_si128(b, 2));
        b = _mm_max_epu8(b, _mm_srli_si128(b, 1));
        return (unsigned)_mm_extract_epi8(b, 0) & 0xff;
    }
    template <int S>
    static inline void gap_step(vec& row, vec adj, vec mm, vec amm, vec gop, vec& gex) {
        vec sr = srp<S>(row, adj), sm = srp<S>(mm, amm);
        row = max(row, subs(sr, add(gex, band(gop, sm))));
        gex = add(gex, gex);
    }
    static inline vec propagate(vec row, vec adj, vec mm, vec amm, vec gop, vec gex) {  // propagate_64_lane
        gap_step<1>(row, adj, mm, amm, gop, gex); gap_step<2>(row, adj, mm, amm, gop, gex);
        gap_step<4>(row, adj, mm, amm, gop, gex); gap_step<8>(row, adj, mm, amm, gop, gex);
        gap_step<16>(row, adj, mm, amm, gop, gex); gap_step<32>(row, adj, mm, amm, gop, gex);
        return row;
    }
};

FRZB_DEFINE_WORKER(avx512, V512)
