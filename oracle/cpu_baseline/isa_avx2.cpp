// AVX2 (32 x u8) instantiation — mirrors BackendAVXU8 (src/smith_waterman/backend/avx.rs:344-353)
// and PrefilterAVXBackend (src/prefilter/backend/avx.rs:209-244).  Compiled with -mavx2 -mbmi -mbmi2 -mlzcnt.
#include <immintrin.h>
#include "../../include/frz_cuda.h"
#include "kernels.inl"
#include "driver.h"

struct V256 {
    static constexpr int LANES = 32;
    using vec = __m256i;
    using msk = __m256i;   // 0xFF / 0x00 per byte, as in the reference's AVX2 backend
    using bits = uint32_t;
    static inline vec zero() { return _mm256_setzero_si256(); }
    static inline msk mzero() { return _mm256_setzero_si256(); }
    static inline vec splat(int c) { return _mm256_set1_epi8((char)c); }
    static inline vec first_lane(uint8_t v) { return _mm256_insert_epi8(_mm256_setzero_si256(), (char)v, 0); }
    static inline bits all_bits() { return ~0u; }
    static inline vec load_partial(const uint8_t* p, size_t rem) {
        if (rem >= 32) return _mm256_loadu_si256((const __m256i*)p);
        alignas(32) uint8_t buf[32] = {0};
        for (size_t i = 0; i < rem; i++) buf[i] = p[i];
        return _mm256_load_si256((const __m256i*)buf);
    }
    static inline bits eq_bits(vec a, vec b) { return (bits)_mm256_movemask_epi8(_mm256_cmpeq_epi8(a, b)); }
    static inline int tz(bits m) { return (int)_tzcnt_u32(m); }
    static inline int lz(bits m) { return (int)_lzcnt_u32(m); }
    static inline msk eq(vec a, vec b) { return _mm256_cmpeq_epi8(a, b); }
    static inline msk gt(vec a, vec b) {  // unsigned a > b
        const vec bias = _mm256_set1_epi8((char)0x80);
        return _mm256_cmpgt_epi8(_mm256_xor_si256(a, bias), _mm256_xor_si256(b, bias));
    }
    static inline msk lt(vec a, vec b) { return gt(b, a); }
    static inline msk mand(msk a, msk b) { return _mm256_and_si256(a, b); }
    static inline msk mor(msk a, msk b) { return _mm256_or_si256(a, b); }
    static inline msk mnot(msk a) { return _mm256_xor_si256(a, _mm256_set1_epi8(-1)); }
    template <int S>
    static inline vec srp(vec v, vec prev) {
        static_assert(S <= 16, "AVX2 u8 backend shifts by at most 16 lanes");
        vec mid = _mm256_permute2x128_si256(prev, v, 0x21);  // [prev.hi, v.lo]
        return _mm256_alignr_epi8(v, mid, 16 - S);
    }
    static inline msk mshift1(msk m, msk prev) { return srp<1>(m, prev); }
    static inline vec widen(msk m) { return m; }
    static inline vec add(vec a, vec b) { return _mm256_add_epi8(a, b); }
    static inline vec subs(vec a, vec b) { return _mm256_subs_epu8(a, b); }
    static inline vec max(vec a, vec b) { return _mm256_max_epu8(a, b); }
    static inline vec band(vec a, vec b) { return _mm256_and_si256(a, b); }
    static inline unsigned hmax(vec v) {
        __m128i b = _mm_max_epu8(_mm256_castsi256_si128(v), _mm256_extracti128_si256(v, 1));
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
    static inline vec propagate(vec row, vec adj, vec mm, vec amm, vec gop, vec gex) {  // propagate_32_lane
        gap_step<1>(row, adj, mm, amm, gop, gex); gap_step<2>(row, adj, mm, amm, gop, gex);
        gap_step<4>(row, adj, mm, amm, gop, gex); gap_step<8>(row, adj, mm, amm, gop, gex);
        gap_step<16>(row, adj, mm, amm, gop, gex);
        return row;
    }
};

FRZB_DEFINE_WORKER(avx2, V256)
