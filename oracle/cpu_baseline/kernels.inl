// kernels.inl — SIMD restatement of the reference's per-haystack pipeline, templated on an ISA
// wrapper `V` (AVX-512 64 x u8, AVX2 32 x u8).  MEASUREMENT INFRASTRUCTURE ONLY (bench.py's
// cpu_baseline / --impl reference legs); validated bit-for-bit against the scalar oracle by
// tests/test_cpu_baseline.py.  It mirrors the structure of the reference's SIMD backends:
//   prefilter  src/prefilter/algo/ascii.rs:6-54, ascii_typos.rs:15-110,363-397 (bitmask state machines)
//   SW         src/smith_waterman/algo/ascii.rs:10-158 + ascii_gap.rs (chunk-major, full matrices)
//   pipeline   src/matcher/algo.rs:78-103,229-263,331-338
// Scope: the u8 score family, ASCII needles, max_typos in {0, 1, None}; anything else makes the
// driver fall back to the scalar oracle.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

template <class V>
struct SimdMatcher {
    using vec = typename V::vec;
    using msk = typename V::msk;      // lane mask (k-register or byte vector)
    using bits = typename V::bits;    // prefilter bitmask integer
    static constexpr int L = V::LANES;
    static constexpr size_t kMaxHay = 1024;

    std::vector<uint8_t> needle;
    std::vector<vec> nc, nf;          // splatted (char, flipped) pairs
    int n = 0;
    int max_typos = 0;                // -1 = None
    size_t min_len = 0;
    bool case_sensitive = false;
    frz_scoring sc{};
    std::vector<vec> H;               // (n+1) x chunks score matrix
    std::vector<msk> M;               // (n+1) x chunks match masks
    size_t stride = 0;

    void init(const uint8_t* nd, size_t len, bool cs, int typos, size_t minl, const frz_scoring& s) {
        needle.assign(nd, nd + len);
        n = (int)len; case_sensitive = cs; max_typos = typos; min_len = minl; sc = s;
        nc.resize(len); nf.resize(len);
        for (size_t i = 0; i < len; i++) {
            uint8_t c = nd[i], f = c;
            if (!cs) { if (c >= 'a' && c <= 'z') f = c - 32; else if (c >= 'A' && c <= 'Z') f = c + 32; }
            nc[i] = V::splat(c); nf[i] = V::splat(f);
        }
        stride = kMaxHay / L + 2;
        H.assign((len + 1) * stride, V::zero());
        M.assign((len + 1) * stride, V::mzero());
    }

    // ---- prefilter -------------------------------------------------------------------------
    static inline bits first_n(size_t k) { return k >= (size_t)L ? V::all_bits() : (((bits)1 << k) - 1); }
    static inline bits ctl(bits m, bits hit) { return m & ~(hit ^ (hit - 1)); }
    inline bits occ(vec chunk, int i) const { return V::eq_bits(chunk, nc[i]) | V::eq_bits(chunk, nf[i]); }

    bool prefilter0(const uint8_t* hay, size_t len, size_t* os, size_t* oe) const {
        if (len == 0) return false;
        bool can_skip = true;
        size_t ms = 0;
        int ni = 0;
        for (size_t start = 0; start < len; start += L) {
            size_t rem = len - start;
            vec chunk = V::load_partial(hay + start, rem);
            bits cm = first_n(rem);
            for (;;) {
                bits mask = occ(chunk, ni) & cm;
                if (!mask) break;
                cm = ctl(cm, mask);
                if (can_skip) { ms = start + V::tz(mask); can_skip = false; }
                if (ni + 1 < n) { ni++; continue; }
                *os = ms;
                if (start + L >= len) { *oe = start + L - V::lz(mask); return true; }
                // find_last_char_pos on hay[start..]
                size_t sl = len - start, st = sl > (size_t)L ? sl - L : 0;
                for (;;) {
                    size_t r2 = sl - st;
                    vec c2 = V::load_partial(hay + start + st, r2);
                    bits m2 = occ(c2, n - 1) & first_n(r2);
                    if (m2) { *oe = start + st + L - V::lz(m2); return true; }
                    st = st > (size_t)L ? st - L : 0;
                }
            }
        }
        return false;
    }

    size_t end_pos_typos(const uint8_t* hay, size_t len, int k) const {
        size_t first = n - 1 - k;
        size_t start = (len - 1) / L * L;
        for (;;) {
            size_t rem = len - start;
            vec chunk = V::load_partial(hay + start, rem);
            bits m = 0;
            for (int i = (int)first; i < n; i++) m |= occ(chunk, i);
            m &= first_n(rem);
            if (m) return start + L - V::lz(m);
            if (start == 0) break;
            start -= L;
        }
        return len;
    }

    bool prefilter1(const uint8_t* hay, size_t len, size_t* os, size_t* oe) const {
        if (n <= 1) { *os = 0; *oe = len; return true; }
        if (len == 0) return false;
        int f = 0, s = 1;
        size_t ms = SIZE_MAX;
        for (size_t start = 0; start < len; start += L) {
            size_t rem = len - start;
            vec chunk = V::load_partial(hay + start, rem);
            bits cm = first_n(rem);
            bits fm = occ(chunk, f), sm = occ(chunk, s), fc = cm, scm = cm;
            for (;;) {
                bool adv = false;
                int cand = f + 1;
                if (cand > s) {
                    if (cand == n) { *os = ms; *oe = end_pos_typos(hay, len, 1); return true; }
                    s = cand; scm = fc; sm = occ(chunk, s);
                } else if (cand == s && fc > scm) scm = fc;
                bits x = fm & fc;
                if (x) { size_t p = start + V::tz(x); if (p < ms) ms = p; f++; fc = ctl(fc, x); fm = occ(chunk, f); adv = true; }
                bits y = sm & scm;
                if (y) {
                    size_t p = start + V::tz(y); if (p < ms) ms = p;
                    s++;
                    if (s >= n) { *os = ms; *oe = end_pos_typos(hay, len, 1); return true; }
                    scm = ctl(scm, y); sm = occ(chunk, s); adv = true;
                }
                if (!adv) break;
            }
        }
        return false;
    }

    // ---- Smith-Waterman (u8 lanes) ------------------------------------------------------------
    uint16_t sw(const uint8_t* hay, size_t hl, bool include_prefix) {
        const size_t chunks = (hl + L - 1) / L + 1;
        auto u8s = [](uint32_t v) { return (uint8_t)v; };
        auto sat_sub = [](uint32_t a, uint32_t b) { return a > b ? a - b : 0u; };
        auto sat_add = [](uint32_t a, uint32_t b) { uint32_t r = a + b; return r > 0xFFFFu ? 0xFFFFu : r; };
        const vec gex0 = V::splat(u8s(sc.gap_extend_penalty));
        const vec gop = V::splat(u8s(sat_sub(sc.gap_open_penalty, sc.gap_extend_penalty)));
        const vec match_score = V::splat(u8s(sat_add(sc.match_score, sc.mismatch_penalty)));
        const vec mismatch = V::splat(u8s(sc.mismatch_penalty));
        const vec case_bonus = V::splat(u8s(sc.matching_case_bonus));
        const vec cap_bonus = V::splat(u8s(sc.capitalization_bonus));
        const vec delim_bonus = V::splat(u8s(sc.delimiter_bonus));
        vec prefix_masked = include_prefix ? V::first_lane(u8s(sc.prefix_bonus)) : V::zero();
        msk prev_delim = V::mzero(), prev_lower = V::mzero();
        vec maxv = V::zero();
        for (size_t col = 1; col < chunks; col++) {
            const size_t off = (col - 1) * L;
            const vec hc = V::load_partial(hay + off, hl - off);
            const msk upper = V::mand(V::lt(hc, V::splat('Z' + 1)), V::gt(hc, V::splat('A' - 1)));
            const msk lower = V::mand(V::lt(hc, V::splat('z' + 1)), V::gt(hc, V::splat('a' - 1)));
            const msk letter = V::mor(upper, lower);
            const msk capm = V::mand(upper, V::mshift1(lower, prev_lower));
            prev_lower = lower;
            const msk digit = V::mand(V::gt(hc, V::splat('0' - 1)), V::lt(hc, V::splat('9' + 1)));
            const msk delim = V::mnot(V::mor(V::mor(letter, digit), V::gt(hc, V::splat(127))));
            const msk delimm = V::mand(V::mshift1(delim, prev_delim), V::mnot(delim));
            prev_delim = delim;
            const vec bonuses = V::add(V::add(V::add(V::band(V::widen(delimm), delim_bonus), V::band(V::widen(capm), cap_bonus)),
                                              prefix_masked), match_score);
            msk up_gap = V::mzero();
            vec prev_row = V::zero(), row = V::zero();
            for (int r = 1; r <= n; r++) {
                const msk exm = V::eq(nc[r - 1], hc);
                const msk mmk = V::mor(exm, V::eq(nf[r - 1], hc));
                const vec mm = V::widen(mmk), ex = V::widen(exm);
                vec diag = V::template srp<1>(prev_row, H[(r - 1) * stride + col - 1]);
                diag = V::add(diag, V::band(mm, bonuses));
                diag = V::subs(diag, mismatch);
                diag = V::add(diag, V::band(ex, case_bonus));
                const vec upv = V::subs(V::subs(prev_row, gex0), V::band(V::widen(up_gap), gop));
                row = V::propagate(V::max(diag, upv), H[r * stride + col - 1], mm, V::widen(M[r * stride + col - 1]), gop, gex0);
                H[r * stride + col] = row;
                M[r * stride + col] = mmk;
                prev_row = row;
                up_gap = mmk;
            }
            maxv = V::max(maxv, row);
            prefix_masked = V::zero();
        }
        return V::hmax(maxv);
    }

    // ---- per-haystack pipeline: returns true and fills m when the haystack matches ------------
    bool match_one(const uint8_t* hay, size_t len, uint32_t index, frz_match* m) {
        if (len < min_len) return false;
        size_t s = 0, e = len;
        if (max_typos == 0) { if (!prefilter0(hay, len, &s, &e)) return false; }
        else if (max_typos == 1) { if (!prefilter1(hay, len, &s, &e)) return false; }
        s = s > 0 ? s - 1 : 0;
        const bool include_exact = s == 0 && e == len;
        const uint8_t* w = hay + s;
        const size_t wl = e - s;
        uint16_t score = sw(w, wl, s == 0);
        const bool exact = include_exact && wl == (size_t)n && memcmp(w, needle.data(), wl) == 0;
        if (exact) score = (uint16_t)(score + sc.exact_match_bonus);
        m->index = index; m->score = score; m->exact = exact; m->_pad = 0;
        return true;
    }
};
