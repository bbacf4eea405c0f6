// frz_oracle.cpp — CPU restatement of the reference's match_list path.
//
// TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// leg may load it.  The product (frizbee_b200/, libfrz_cuda.so) never links or calls it.
//
// PARITY PINNING: the reference is a Rust crate and no Rust toolchain exists in this
// image, so the reference itself cannot be run here.  The oracle is pinned against
// every known-answer vector the reference's own tests hold for this path
// (tests/test_oracle_kat.py lists each with its source file:line).  The crate has no
// third-party algorithmic dependency (Cargo.toml:33-34), so the algorithm is entirely
// restated from /root/reference/src.
//
// Each function cites the reference file:line it follows (paths relative to the
// reference crate root).  Vectors are restated with *real* u8 / u16 lane arithmetic
// (wrapping add, zero-saturating sub, `value as u8` truncation of constants) and a
// runtime LANES so that any reference backend (LANES, width) can be emulated:
//   Scalar 8xu16 / 16xu8, SSE 8xu16 / 16xu8, AVX2 16xu16 / 32xu8, AVX-512 32xu16 / 64xu8
//   (src/smith_waterman/backend/scalar.rs:440-455, avx.rs, avx512.rs).
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../include/frz_cuda.h"

namespace {

constexpr int kMaxLanes = 64;
constexpr size_t kMaxHaystackLen = 1024;  // src/smith_waterman/algo/mod.rs:18

struct Pair {
    uint8_t c, flip;
};

// src/prefilter/mod.rs:49-65 (case_needle)
static std::vector<Pair> case_needle(const uint8_t* needle, size_t n, bool case_sensitive) {
    std::vector<Pair> out(n);
    for (size_t i = 0; i < n; i++) {
        uint8_t c = needle[i];
        uint8_t f;
        if (case_sensitive) f = c;
        else if (c >= 'a' && c <= 'z') f = (uint8_t)(c - 32);
        else if (c >= 'A' && c <= 'Z') f = (uint8_t)(c + 32);
        else f = c;
        out[i] = {c, f};
    }
    return out;
}

// ---------------------------------------------------------------------------------
// Prefilter.  Masks are restated as uint64_t holding LANES meaningful low bits
// (u16/u32/u64 in the reference: src/prefilter/backend/mod.rs:45-114).
// ---------------------------------------------------------------------------------
struct Pf {
    int lanes;
    const std::vector<Pair>& needle;
    const uint8_t* hay;
    size_t len;

    uint64_t all() const { return lanes == 64 ? ~0ull : ((1ull << lanes) - 1); }
    // BitMaskOps::first_n (backend/mod.rs:71-77)
    uint64_t first_n(size_t n) const { return n >= (size_t)lanes ? all() : ((1ull << n) - 1); }
    // leading_zeros of the LANES-bit mask type
    int lz(uint64_t m) const { return __builtin_clzll(m) - (64 - lanes); }
    static int tz(uint64_t m) { return __builtin_ctzll(m); }
    // clear_through_lowest (backend/mod.rs:105-107)
    static uint64_t ctl(uint64_t mask, uint64_t hit) { return mask & ~(hit ^ (hit - 1)); }
    // load_window (src/prefilter/algo/load.rs:4-26): returns the chunk mask; the chunk bytes are
    // read through occ() with the same masking effect (lanes >= remaining never survive `& mask`).
    uint64_t chunk_mask(size_t start) const { return first_n(len - start); }
    // Backend::occ (backend/scalar.rs:51-59): either-case equality bitmask of the chunk at `start`.
    // Lanes past the end of the haystack are garbage in the reference (over-read) and always
    // masked by chunk_mask by every caller; we return 0 there.
    uint64_t occ(size_t start, Pair p) const {
        uint64_t m = 0;
        for (int i = 0; i < lanes; i++) {
            size_t pos = start + i;
            if (pos >= len) break;
            uint8_t b = hay[pos];
            if (b == p.c || b == p.flip) m |= 1ull << i;
        }
        return m;
    }
};

// src/prefilter/algo/ascii.rs:57-72 (find_last_char_pos) on the sub-slice hay[off..]
static size_t find_last_char_pos(const Pf& base, Pair last, size_t off) {
    Pf s{base.lanes, base.needle, base.hay + off, base.len - off};
    size_t len = s.len;
    size_t start = len > (size_t)s.lanes ? len - s.lanes : 0;
    for (;;) {
        uint64_t mask = s.occ(start, last) & s.chunk_mask(start);
        if (mask) return start + s.lanes - s.lz(mask);
        start = start > (size_t)s.lanes ? start - s.lanes : 0;
    }
}

// src/prefilter/algo/ascii.rs:6-54 (match_haystack, 0 typos)
static bool prefilter_k0(const Pf& p, size_t* ostart, size_t* oend) {
    size_t len = p.len;
    if (len == 0) { *ostart = 0; *oend = 0; return false; }
    bool can_skip = true;
    size_t match_start = 0;
    size_t ni = 0;
    size_t n = p.needle.size();
    Pair nc = p.needle[0];
    size_t start = 0;
    while (start < len) {
        uint64_t chunk_mask = p.chunk_mask(start);
        for (;;) {
            uint64_t mask = p.occ(start, nc) & chunk_mask;
            if (!mask) break;
            chunk_mask = Pf::ctl(chunk_mask, mask);
            if (can_skip) { match_start = start + Pf::tz(mask); can_skip = false; }
            if (ni + 1 < n) {
                ni++;
                nc = p.needle[ni];
            } else if (start + p.lanes >= len) {
                *ostart = match_start;
                *oend = start + p.lanes - p.lz(mask);
                return true;
            } else {
                *ostart = match_start;
                *oend = start + find_last_char_pos(p, p.needle[n - 1], start);
                return true;
            }
        }
        start += p.lanes;
    }
    *ostart = match_start;
    *oend = len;
    return false;
}

// src/prefilter/algo/ascii_typos.rs:375-397 (find_end_pos_with_typos)
static size_t find_end_pos_with_typos(const Pf& p, size_t max_typos) {
    size_t len = p.len, n = p.needle.size();
    size_t first = n - 1 - max_typos;
    size_t start = (len - 1) / p.lanes * p.lanes;
    for (;;) {
        uint64_t mask = 0;
        for (size_t i = first; i < n; i++) mask |= p.occ(start, p.needle[i]);
        mask &= p.chunk_mask(start);
        if (mask) return start + p.lanes - p.lz(mask);
        if (start == 0) break;
        start -= p.lanes;
    }
    return len;
}

// src/prefilter/algo/ascii_typos.rs:15-110 (match_haystack_1_typo)
static bool prefilter_k1(const Pf& p, size_t* ostart, size_t* oend) {
    size_t len = p.len, n = p.needle.size();
    if (n <= 1) { *ostart = 0; *oend = len; return true; }
    if (len == 0) { *ostart = 0; *oend = 0; return false; }
    size_t f = 0, s = 1;
    size_t ms = SIZE_MAX;
    for (size_t start = 0; start < len; start += p.lanes) {
        uint64_t cm = p.chunk_mask(start);
        uint64_t fm = p.occ(start, p.needle[f]);
        uint64_t sm = p.occ(start, p.needle[s]);
        uint64_t fc = cm, sc = cm;
        for (;;) {
            bool advanced = false;
            size_t cand = f + 1;
            if (cand > s) {
                if (cand == n) { *ostart = ms; *oend = find_end_pos_with_typos(p, 1); return true; }
                s = cand;
                sc = fc;
                sm = p.occ(start, p.needle[s]);
            } else if (cand == s && fc > sc) {
                sc = fc;
            }
            uint64_t x = fm & fc;
            if (x) {
                ms = std::min(ms, start + (size_t)Pf::tz(x));
                f++;
                fc = Pf::ctl(fc, x);
                fm = p.occ(start, p.needle[f]);
                advanced = true;
            }
            uint64_t y = sm & sc;
            if (y) {
                ms = std::min(ms, start + (size_t)Pf::tz(y));
                s++;
                if (s >= n) { *ostart = ms; *oend = find_end_pos_with_typos(p, 1); return true; }
                sc = Pf::ctl(sc, y);
                sm = p.occ(start, p.needle[s]);
                advanced = true;
            }
            if (!advanced) break;
        }
    }
    *ostart = ms == SIZE_MAX ? 0 : ms;
    *oend = len;
    return false;
}

// src/prefilter/algo/ascii_typos.rs:113-251 (match_haystack_2_typos)
static bool prefilter_k2(const Pf& p, size_t* ostart, size_t* oend) {
    size_t len = p.len, n = p.needle.size();
    if (n <= 2) { *ostart = 0; *oend = len; return true; }
    if (len == 0) { *ostart = 0; *oend = 0; return false; }
    size_t i1 = 0, i2 = 1, i3 = 2;
    size_t ms = SIZE_MAX;
    auto found = [&]() { *ostart = ms; *oend = find_end_pos_with_typos(p, 2); return true; };
    for (size_t start = 0; start < len; start += p.lanes) {
        uint64_t cm = p.chunk_mask(start);
        uint64_t m1 = p.occ(start, p.needle[i1]);
        uint64_t m2 = p.occ(start, p.needle[i2]);
        uint64_t m3 = p.occ(start, p.needle[i3]);
        uint64_t c1 = cm, c2 = cm, c3 = cm;
        for (;;) {
            bool advanced = false;
            size_t cand2 = i1 + 1;
            if (cand2 > i2) {
                if (cand2 == n) return found();
                i2 = cand2; c2 = c1; m2 = p.occ(start, p.needle[i2]);
            } else if (cand2 == i2 && c1 > c2) {
                c2 = c1;
            }
            size_t cand3 = i2 + 1;
            if (cand3 > i3) {
                if (cand3 == n) return found();
                i3 = cand3; c3 = c2; m3 = p.occ(start, p.needle[i3]);
            } else if (cand3 == i3 && c2 > c3) {
                c3 = c2;
            }
            uint64_t x1 = m1 & c1;
            if (x1) {
                ms = std::min(ms, start + (size_t)Pf::tz(x1));
                i1++;
                c1 = Pf::ctl(c1, x1);
                m1 = p.occ(start, p.needle[i1]);
                advanced = true;
            }
            uint64_t x2 = m2 & c2;
            if (x2) {
                ms = std::min(ms, start + (size_t)Pf::tz(x2));
                i2++;
                if (i2 >= n) return found();
                c2 = Pf::ctl(c2, x2);
                m2 = p.occ(start, p.needle[i2]);
                advanced = true;
            }
            uint64_t x3 = m3 & c3;
            if (x3) {
                ms = std::min(ms, start + (size_t)Pf::tz(x3));
                i3++;
                if (i3 >= n) return found();
                c3 = Pf::ctl(c3, x3);
                m3 = p.occ(start, p.needle[i3]);
                advanced = true;
            }
            if (!advanced) break;
        }
    }
    *ostart = ms == SIZE_MAX ? 0 : ms;
    *oend = len;
    return false;
}

// src/prefilter/algo/ascii_typos.rs:254-360 (match_haystack_many_typos_impl)
static bool prefilter_many(const Pf& p, size_t max_typos, size_t* ostart, size_t* oend) {
    size_t len = p.len, n = p.needle.size();
    if (n <= max_typos) { *ostart = 0; *oend = len; return true; }
    if (len == 0) { *ostart = 0; *oend = 0; return false; }
    size_t path_count = max_typos + 1;
    std::vector<size_t> idx(path_count, 0);
    std::vector<uint64_t> nm(path_count, 0);
    size_t ms = SIZE_MAX;
    auto found = [&]() { *ostart = ms; *oend = find_end_pos_with_typos(p, max_typos); return true; };
    for (size_t start = 0; start < len; start += p.lanes) {
        uint64_t chunk_mask = p.chunk_mask(start);
        for (size_t k = 0; k < path_count; k++) nm[k] = p.occ(start, p.needle[idx[k]]);
        for (;;) {
            for (size_t k = 1; k < path_count; k++) {
                size_t cand = idx[k - 1] + 1;
                if (cand > idx[k]) {
                    if (cand == n) return found();
                    idx[k] = cand;
                    nm[k] = p.occ(start, p.needle[cand]);
                }
            }
            uint64_t mm = 0;
            for (size_t k = 0; k < path_count; k++) mm |= nm[k];
            uint64_t matches = mm & chunk_mask;
            if (!matches) break;
            size_t hit_pos = Pf::tz(matches);
            uint64_t hit = matches & p.first_n(hit_pos + 1);
            ms = std::min(ms, start + hit_pos);
            for (size_t k = 0; k < path_count; k++) {
                if (!(nm[k] & hit)) continue;
                idx[k]++;
                if (idx[k] == n) return found();
                nm[k] = p.occ(start, p.needle[idx[k]]);
            }
            chunk_mask = Pf::ctl(chunk_mask, hit);
        }
    }
    *ostart = ms == SIZE_MAX ? 0 : ms;
    *oend = len;
    return false;
}

// ---------------------------------------------------------------------------------
// Unicode needle path.  src/prefilter/mod.rs:21-96 (UnicodeChar, case_needle_unicode),
// src/prefilter/algo/unicode.rs, src/prefilter/algo/unicode_typos.rs.
// Case mappings come from oracle/unicode_case.inc (tools/gen_unicode_case.py): the data behind Rust's
// char::to_lowercase / to_uppercase / is_uppercase, reduced to what the reference can use.
// ---------------------------------------------------------------------------------
#include "unicode_case.inc"

struct UChar {
    uint8_t c[4];
    uint8_t f[4];
    int len;
};

static int utf8_encode(uint32_t cp, uint8_t* o) {
    if (cp < 0x80) { o[0] = (uint8_t)cp; return 1; }
    if (cp < 0x800) { o[0] = (uint8_t)(0xC0 | (cp >> 6)); o[1] = (uint8_t)(0x80 | (cp & 0x3F)); return 2; }
    if (cp < 0x10000) {
        o[0] = (uint8_t)(0xE0 | (cp >> 12)); o[1] = (uint8_t)(0x80 | ((cp >> 6) & 0x3F)); o[2] = (uint8_t)(0x80 | (cp & 0x3F));
        return 3;
    }
    o[0] = (uint8_t)(0xF0 | (cp >> 18)); o[1] = (uint8_t)(0x80 | ((cp >> 12) & 0x3F));
    o[2] = (uint8_t)(0x80 | ((cp >> 6) & 0x3F)); o[3] = (uint8_t)(0x80 | (cp & 0x3F));
    return 4;
}
// one scalar from valid UTF-8 (the needle is a Rust &str)
static uint32_t utf8_decode(const uint8_t* s, size_t n, size_t* i) {
    uint8_t b = s[*i];
    int extra = b < 0x80 ? 0 : b < 0xE0 ? 1 : b < 0xF0 ? 2 : 3;
    uint32_t cp = extra == 0 ? b : extra == 1 ? (b & 0x1Fu) : extra == 2 ? (b & 0x0Fu) : (b & 0x07u);
    (*i)++;
    for (int k = 0; k < extra && *i < n; k++, (*i)++) cp = (cp << 6) | (s[*i] & 0x3Fu);
    return cp;
}
// char::is_uppercase (used by CaseMatching::Smart, src/lib.rs:370-376)
static bool scalar_is_uppercase(uint32_t cp) {
    if (cp < 0x80) return cp >= 'A' && cp <= 'Z';
    size_t lo = 0, hi = sizeof(kUnicodeUpper) / sizeof(kUnicodeUpper[0]);
    while (lo < hi) {
        size_t mid = (lo + hi) / 2;
        if (cp < kUnicodeUpper[mid][0]) hi = mid;
        else if (cp > kUnicodeUpper[mid][1]) lo = mid + 1;
        else return true;
    }
    return false;
}
// the opposite-case scalar of src/prefilter/mod.rs:76-91, or cp itself
static uint32_t scalar_flip(uint32_t cp) {
    if (cp < 0x80) {
        if (cp >= 'a' && cp <= 'z') return cp - 32;
        if (cp >= 'A' && cp <= 'Z') return cp + 32;
        return cp;
    }
    size_t lo = 0, hi = sizeof(kUnicodeFlip) / sizeof(kUnicodeFlip[0]);
    while (lo < hi) {
        size_t mid = (lo + hi) / 2;
        if (kUnicodeFlip[mid][0] < cp) lo = mid + 1; else hi = mid;
    }
    if (lo < sizeof(kUnicodeFlip) / sizeof(kUnicodeFlip[0]) && kUnicodeFlip[lo][0] == cp) return kUnicodeFlip[lo][1];
    return cp;
}
static bool needle_has_uppercase(const uint8_t* needle, size_t n) {
    for (size_t i = 0; i < n;)
        if (scalar_is_uppercase(utf8_decode(needle, n, &i))) return true;
    return false;
}
// src/prefilter/mod.rs:71-96 (case_needle_unicode)
static std::vector<UChar> case_needle_unicode(const uint8_t* needle, size_t n, bool case_sensitive) {
    std::vector<UChar> out;
    for (size_t i = 0; i < n;) {
        uint32_t cp = utf8_decode(needle, n, &i);
        uint32_t fl = case_sensitive ? cp : scalar_flip(cp);
        UChar u{};
        u.len = utf8_encode(cp, u.c);
        int fl_len = utf8_encode(fl, u.f);
        if (fl_len != u.len) memcpy(u.f, u.c, 4);   // (the table only holds same-length pairs)
        out.push_back(u);
    }
    return out;
}

struct UPf {
    int lanes;
    const std::vector<UChar>& needle;
    const uint8_t* hay;
    size_t len;
    uint64_t all() const { return lanes == 64 ? ~0ull : ((1ull << lanes) - 1); }
    uint64_t first_n(size_t n) const { return n >= (size_t)lanes ? all() : ((1ull << n) - 1); }
    int lz(uint64_t m) const { return __builtin_clzll(m) - (64 - lanes); }
    // B::eq of the chunk loaded at `pos` (load_window / load_window_maskless, src/prefilter/algo/load.rs):
    // lanes past the end of the haystack hold over-read garbage in the reference; every caller masks them
    // (the last-byte window is the shortest), so they are returned as "no match" here.
    uint64_t eq(size_t pos, uint8_t b) const {
        uint64_t m = 0;
        for (int i = 0; i < lanes; i++) {
            size_t q = pos + i;
            if (q >= len) break;
            if (hay[q] == b) m |= 1ull << i;
        }
        return m;
    }
    // match_unicode_char_prefix (unicode.rs:8-50)
    uint64_t prefix(size_t start, int char_len, const uint8_t* chars) const {
        uint64_t m = all();
        for (int k = 0; k < char_len - 1; k++) m &= eq(start + k, chars[k]);
        return m;
    }
    // char_variant_mask (unicode.rs:54-70) on the window whose last-byte chunk sits at start + char_len - 1
    uint64_t variant(size_t start, uint64_t chunk_mask, int char_len, const uint8_t* chars) const {
        uint64_t mask = eq(start + char_len - 1, chars[char_len - 1]) & chunk_mask;
        if (mask && char_len > 1) mask &= prefix(start, char_len, chars);
        return mask;
    }
    // unicode_char_mask (unicode.rs:73-117)
    uint64_t char_mask(size_t start, const UChar& c) const {
        if (start + c.len > len) return 0;
        uint64_t chunk_mask = first_n(len - (start + c.len - 1));
        return variant(start, chunk_mask, c.len, c.c) | variant(start, chunk_mask, c.len, c.f);
    }
};

// unicode.rs:225-275 (find_last_unicode_char_pos) on the sub-slice hay[off..]
static size_t find_last_unicode_char_pos(const UPf& base, const UChar& nc, size_t off) {
    UPf s{base.lanes, base.needle, base.hay + off, base.len - off};
    size_t len = s.len;
    int cl = nc.len;
    size_t start = len > (size_t)(s.lanes + cl - 1) ? len - (s.lanes + cl - 1) : 0;
    for (;;) {
        uint64_t chunk_mask = s.first_n(len - (start + cl - 1));
        uint64_t mask = (s.eq(start + cl - 1, nc.c[cl - 1]) | s.eq(start + cl - 1, nc.f[cl - 1])) & chunk_mask;
        if (mask && cl > 1) mask &= s.prefix(start, cl, nc.c) | s.prefix(start, cl, nc.f);
        if (mask) return start + s.lanes - s.lz(mask) + cl - 1;
        if (start == 0) break;
        start = start > (size_t)s.lanes ? start - s.lanes : 0;
    }
    return len;
}

// unicode.rs:120-222 (match_haystack_unicode, 0 typos)
static bool uprefilter_k0(const UPf& p, size_t* ostart, size_t* oend) {
    size_t len = p.len;
    if (len == 0) { *ostart = 0; *oend = 0; return false; }
    bool can_skip = true;
    size_t ms = 0, ni = 0, n = p.needle.size();
    const UChar* nc = &p.needle[0];
    size_t start = 0;
    while (start + nc->len <= len) {
        int char_len = nc->len;
        uint64_t valid = p.first_n(len - (start + char_len - 1));
        uint64_t available = p.all();
        for (;;) {
            uint64_t chunk_mask = available & valid;
            uint64_t mask = p.variant(start, chunk_mask, nc->len, nc->c) | p.variant(start, chunk_mask, nc->len, nc->f);
            if (!mask) break;
            available = Pf::ctl(available, mask);
            if (can_skip) { ms = start + Pf::tz(mask); can_skip = false; }
            if (ni + 1 < n) {
                ni++;
                nc = &p.needle[ni];
                if (nc->len != char_len) {
                    if (start + nc->len > len) break;
                    char_len = nc->len;
                    valid = p.first_n(len - (start + char_len - 1));
                }
            } else if (start + nc->len - 1 + p.lanes >= len) {
                *ostart = ms;
                *oend = start + p.lanes - p.lz(mask) + nc->len - 1;
                return true;
            } else {
                *ostart = ms;
                *oend = start + find_last_unicode_char_pos(p, *nc, start);
                return true;
            }
        }
        start += p.lanes;
    }
    *ostart = ms;
    *oend = len;
    return false;
}

// unicode_typos.rs:486-508 (find_end_pos_with_unicode_typos)
static size_t find_end_pos_with_unicode_typos(const UPf& p, size_t max_typos) {
    size_t len = p.len, n = p.needle.size();
    size_t first = n - 1 - max_typos;
    size_t start = len > (size_t)p.lanes ? len - p.lanes : 0;
    for (;;) {
        size_t end_pos = 0;
        for (size_t i = first; i < n; i++) {
            uint64_t mask = p.char_mask(start, p.needle[i]);
            if (mask) end_pos = std::max(end_pos, start + p.lanes - p.lz(mask) + p.needle[i].len - 1);
        }
        if (end_pos) return end_pos;
        if (start == 0) break;
        start = start > (size_t)p.lanes ? start - p.lanes : 0;
    }
    return len;
}

// unicode_typos.rs:15-143 (match_haystack_unicode_1_typo)
static bool uprefilter_k1(const UPf& p, size_t* ostart, size_t* oend) {
    size_t len = p.len, n = p.needle.size();
    if (n <= 1) { *ostart = 0; *oend = len; return true; }
    if (len == 0) { *ostart = 0; *oend = 0; return false; }
    size_t f = 0, s = 1, ms = SIZE_MAX;
    auto found = [&]() { *ostart = ms; *oend = find_end_pos_with_unicode_typos(p, 1); return true; };
    for (size_t start = 0; start < len; start += p.lanes) {
        uint64_t fm = p.char_mask(start, p.needle[f]);
        uint64_t sm = p.char_mask(start, p.needle[s]);
        uint64_t fc = p.all(), sc = p.all();
        for (;;) {
            bool advanced = false;
            size_t cand = f + 1;
            if (cand > s) {
                if (cand == n) return found();
                s = cand; sc = fc; sm = p.char_mask(start, p.needle[s]);
            } else if (cand == s && fc > sc) {
                sc = fc;
            }
            uint64_t x = fm & fc;
            if (x) {
                ms = std::min(ms, start + (size_t)Pf::tz(x));
                f++;
                fc = Pf::ctl(fc, x);
                fm = p.char_mask(start, p.needle[f]);
                advanced = true;
            }
            uint64_t y = sm & sc;
            if (y) {
                ms = std::min(ms, start + (size_t)Pf::tz(y));
                s++;
                if (s >= n) return found();
                sc = Pf::ctl(sc, y);
                sm = p.char_mask(start, p.needle[s]);
                advanced = true;
            }
            if (!advanced) break;
        }
    }
    *ostart = ms == SIZE_MAX ? 0 : ms;
    *oend = len;
    return false;
}

// unicode_typos.rs:146-333 (match_haystack_unicode_2_typos)
static bool uprefilter_k2(const UPf& p, size_t* ostart, size_t* oend) {
    size_t len = p.len, n = p.needle.size();
    if (n <= 2) { *ostart = 0; *oend = len; return true; }
    if (len == 0) { *ostart = 0; *oend = 0; return false; }
    size_t i1 = 0, i2 = 1, i3 = 2, ms = SIZE_MAX;
    auto found = [&]() { *ostart = ms; *oend = find_end_pos_with_unicode_typos(p, 2); return true; };
    for (size_t start = 0; start < len; start += p.lanes) {
        uint64_t m1 = p.char_mask(start, p.needle[i1]);
        uint64_t m2 = p.char_mask(start, p.needle[i2]);
        uint64_t m3 = p.char_mask(start, p.needle[i3]);
        uint64_t c1 = p.all(), c2 = p.all(), c3 = p.all();
        for (;;) {
            bool advanced = false;
            size_t cand2 = i1 + 1;
            if (cand2 > i2) {
                if (cand2 == n) return found();
                i2 = cand2; c2 = c1; m2 = p.char_mask(start, p.needle[i2]);
            } else if (cand2 == i2 && c1 > c2) {
                c2 = c1;
            }
            size_t cand3 = i2 + 1;
            if (cand3 > i3) {
                if (cand3 == n) return found();
                i3 = cand3; c3 = c2; m3 = p.char_mask(start, p.needle[i3]);
            } else if (cand3 == i3 && c2 > c3) {
                c3 = c2;
            }
            uint64_t x1 = m1 & c1;
            if (x1) {
                ms = std::min(ms, start + (size_t)Pf::tz(x1));
                i1++;
                c1 = Pf::ctl(c1, x1);
                m1 = p.char_mask(start, p.needle[i1]);
                advanced = true;
            }
            uint64_t x2 = m2 & c2;
            if (x2) {
                ms = std::min(ms, start + (size_t)Pf::tz(x2));
                i2++;
                if (i2 >= n) return found();
                c2 = Pf::ctl(c2, x2);
                m2 = p.char_mask(start, p.needle[i2]);
                advanced = true;
            }
            uint64_t x3 = m3 & c3;
            if (x3) {
                ms = std::min(ms, start + (size_t)Pf::tz(x3));
                i3++;
                if (i3 >= n) return found();
                c3 = Pf::ctl(c3, x3);
                m3 = p.char_mask(start, p.needle[i3]);
                advanced = true;
            }
            if (!advanced) break;
        }
    }
    *ostart = ms == SIZE_MAX ? 0 : ms;
    *oend = len;
    return false;
}

// unicode_typos.rs:336-472 (match_haystack_unicode_many_typos_impl)
static bool uprefilter_many(const UPf& p, size_t max_typos, size_t* ostart, size_t* oend) {
    size_t len = p.len, n = p.needle.size();
    if (n <= max_typos) { *ostart = 0; *oend = len; return true; }
    if (len == 0) { *ostart = 0; *oend = 0; return false; }
    size_t path_count = max_typos + 1;
    std::vector<size_t> idx(path_count, 0);
    std::vector<uint64_t> nm(path_count, 0);
    size_t ms = SIZE_MAX;
    auto found = [&]() { *ostart = ms; *oend = find_end_pos_with_unicode_typos(p, max_typos); return true; };
    for (size_t start = 0; start < len; start += p.lanes) {
        uint64_t chunk_mask = p.all();
        for (size_t k = 0; k < path_count; k++) nm[k] = p.char_mask(start, p.needle[idx[k]]);
        for (;;) {
            for (size_t k = 1; k < path_count; k++) {
                size_t cand = idx[k - 1] + 1;
                if (cand > idx[k]) {
                    if (cand == n) return found();
                    idx[k] = cand;
                    nm[k] = p.char_mask(start, p.needle[cand]);
                }
            }
            uint64_t mm = 0;
            for (size_t k = 0; k < path_count; k++) mm |= nm[k];
            uint64_t matches = mm & chunk_mask;
            if (!matches) break;
            size_t hit_pos = Pf::tz(matches);
            uint64_t hit = matches & p.first_n(hit_pos + 1);
            ms = std::min(ms, start + hit_pos);
            for (size_t k = 0; k < path_count; k++) {
                if (!(nm[k] & hit)) continue;
                idx[k]++;
                if (idx[k] == n) return found();
                nm[k] = p.char_mask(start, p.needle[idx[k]]);
            }
            chunk_mask = Pf::ctl(chunk_mask, hit);
        }
    }
    *ostart = ms == SIZE_MAX ? 0 : ms;
    *oend = len;
    return false;
}

// src/matcher/algo.rs:171-193 (prefilter_haystack dispatch), UNICODE = true
static bool uprefilter(const std::vector<UChar>& needle, const uint8_t* hay, size_t len, int max_typos, int lanes,
                       size_t* s, size_t* e) {
    if (max_typos < 0) { *s = 0; *e = len; return true; }
    UPf p{lanes, needle, hay, len};
    switch (max_typos) {
        case 0: return uprefilter_k0(p, s, e);
        case 1: return uprefilter_k1(p, s, e);
        case 2: return uprefilter_k2(p, s, e);
        default: return uprefilter_many(p, (size_t)max_typos, s, e);
    }
}

// src/matcher/algo.rs:171-193 (prefilter_haystack dispatch); max_typos < 0 == NO_PREFILTER
static bool prefilter(const std::vector<Pair>& needle, const uint8_t* hay, size_t len, int max_typos,
                      int lanes, size_t* s, size_t* e) {
    if (max_typos < 0) { *s = 0; *e = len; return true; }
    Pf p{lanes, needle, hay, len};
    switch (max_typos) {
        case 0: return prefilter_k0(p, s, e);
        case 1: return prefilter_k1(p, s, e);
        case 2: return prefilter_k2(p, s, e);
        default: return prefilter_many(p, (size_t)max_typos, s, e);
    }
}

// ---------------------------------------------------------------------------------
// Smith-Waterman.  ScoreVec with a runtime lane count and element width
// (src/smith_waterman/backend/mod.rs:205-274, scalar.rs:163-354).
// ---------------------------------------------------------------------------------
struct SV {
    uint16_t v[kMaxLanes];
};

struct Arith {
    int lanes;
    bool u8;
    uint16_t trunc(uint16_t x) const { return u8 ? (uint16_t)(x & 0xFF) : x; }
    SV zero() const { SV r; memset(&r, 0, sizeof r); return r; }
    SV splat(uint16_t x) const { SV r = zero(); for (int i = 0; i < lanes; i++) r.v[i] = trunc(x); return r; }
    SV first_lane(uint16_t x) const { SV r = zero(); r.v[0] = trunc(x); return r; }
    SV add(const SV& a, const SV& b) const {  // wrapping
        SV r = zero();
        for (int i = 0; i < lanes; i++) r.v[i] = trunc((uint16_t)(a.v[i] + b.v[i]));
        return r;
    }
    SV subs(const SV& a, const SV& b) const {  // saturating at zero
        SV r = zero();
        for (int i = 0; i < lanes; i++) r.v[i] = a.v[i] > b.v[i] ? (uint16_t)(a.v[i] - b.v[i]) : 0;
        return r;
    }
    SV max(const SV& a, const SV& b) const {
        SV r = zero();
        for (int i = 0; i < lanes; i++) r.v[i] = std::max(a.v[i], b.v[i]);
        return r;
    }
    SV band(const SV& a, const SV& b) const {
        SV r = zero();
        for (int i = 0; i < lanes; i++) r.v[i] = a.v[i] & b.v[i];
        return r;
    }
    // shift_right_padded::<L> (scalar.rs:223-232)
    SV srp(const SV& a, const SV& prev, int L) const {
        SV r = zero();
        for (int i = 0; i < L; i++) r.v[i] = prev.v[lanes - L + i];
        for (int i = L; i < lanes; i++) r.v[i] = a.v[i - L];
        return r;
    }
    uint16_t hmax(const SV& a) const {
        uint16_t m = 0;
        for (int i = 0; i < lanes; i++) m = std::max(m, a.v[i]);
        return m;
    }
    uint16_t full() const { return u8 ? 0xFF : 0xFFFF; }
};

// src/smith_waterman/algo/ascii_gap.rs:11-105 (gap_step! / propagate_N_lane)
static SV propagate(const Arith& A, SV row, const SV& adj, const SV& mm, const SV& amm, const SV& gop, SV gex) {
    for (int s = 1; s < A.lanes; s <<= 1) {
        SV shifted_row = A.srp(row, adj, s);
        SV shifted_mm = A.srp(mm, amm, s);
        SV pen = A.add(gex, A.band(gop, shifted_mm));
        SV decayed = A.subs(shifted_row, pen);
        row = A.max(row, decayed);
        gex = A.add(gex, gex);
    }
    return row;
}

// src/smith_waterman/greedy.rs:7-91 (match_greedy); returns -1 for None
static int match_greedy(const uint8_t* needle_raw, size_t n, const uint8_t* hay, size_t hl, const frz_scoring& sc,
                        bool case_sensitive, bool include_prefix, std::vector<uint32_t>* indices = nullptr) {
    std::vector<Pair> needle = case_needle(needle_raw, n, case_sensitive);
    if (n > hl) return -1;
    auto sat_add = [](uint16_t a, uint16_t b) { uint32_t r = (uint32_t)a + b; return (uint16_t)(r > 0xFFFF ? 0xFFFF : r); };
    auto sat_sub = [](uint16_t a, uint16_t b) { return (uint16_t)(a > b ? a - b : 0); };
    auto sat_mul = [](uint16_t a, uint16_t b) { uint32_t r = (uint32_t)a * b; return (uint16_t)(r > 0xFFFF ? 0xFFFF : r); };
    uint16_t score = 0;
    size_t hi = 0;
    bool delim_enabled = false, prev_lower = false, prev_delim = false;
    for (size_t ni = 0; ni < n; ni++) {
        size_t hstart = hi;
        bool matched = false;
        while (hi <= hl - n + ni) {
            uint8_t hc = hay[hi];
            bool is_digit = hc >= '0' && hc <= '9';
            bool is_upper = hc >= 'A' && hc <= 'Z';
            bool is_lower = hc >= 'a' && hc <= 'z';
            bool is_delim = hc < 128 && !(is_lower || is_upper || is_digit);
            if (!is_delim) delim_enabled = true;
            if (needle[ni].c != hc && needle[ni].flip != hc) {
                prev_delim = delim_enabled && is_delim;
                prev_lower = is_lower;
                hi++;
                continue;
            }
            score = sat_add(score, sc.match_score);
            if (hi != hstart && ni != 0) {
                size_t gl = hi - hstart;
                gl = gl > 0 ? gl - 1 : 0;
                uint16_t gap_len = (uint16_t)std::min<size_t>(gl, 0xFFFF);
                score = sat_sub(score, sat_add(sc.gap_open_penalty, sat_mul(sc.gap_extend_penalty, gap_len)));
            }
            if (needle[ni].c == hc) score = sat_add(score, sc.matching_case_bonus);
            if (is_upper && prev_lower) score = sat_add(score, sc.capitalization_bonus);
            if (include_prefix && hi == 0) score = sat_add(score, sc.prefix_bonus);
            if (prev_delim && !is_delim) score = sat_add(score, sc.delimiter_bonus);
            prev_delim = delim_enabled && is_delim;
            prev_lower = is_lower;
            if (indices) indices->push_back((uint32_t)hi);
            hi++;
            matched = true;
            break;
        }
        if (!matched) return -1;
    }
    return score;
}

// src/smith_waterman/algo/ascii.rs:10-158 (score_haystack), chunk-major exactly as the reference,
// with the full score / match-mask matrices.
// max_cols (test-only, tests/test_oracle_kat.py::test_column_limit_property): take the final maximum over the
// last-row cells of columns < max_cols only.  Cells never depend on cells to their right, so this equals a
// computation that stops at column max_cols; the reference itself always uses every lane (SIZE_MAX).
struct SwMatrices {     // score_matrix / match_masks of the last call (src/smith_waterman/matrix.rs), chunk 0 = zero column
    std::vector<SV> H, M;
    size_t chunks = 0;
};

static uint16_t sw_score(const uint8_t* needle_raw, size_t n, const frz_scoring& sc, bool case_sensitive,
                         const uint8_t* hay, size_t hl, bool include_prefix, int lanes, bool u8, size_t max_cols = SIZE_MAX,
                         SwMatrices* mats = nullptr) {
    if (hl > kMaxHaystackLen) {
        int g = match_greedy(needle_raw, n, hay, hl, sc, case_sensitive, include_prefix);
        return g < 0 ? 0 : (uint16_t)g;
    }
    Arith A{lanes, u8};
    std::vector<Pair> needle = case_needle(needle_raw, n, case_sensitive);
    size_t chunks = (hl + lanes - 1) / lanes + 1;
    // row 0 and column 0 are zero (ascii.rs:27-31)
    std::vector<SV> H((n + 1) * chunks, A.zero());
    std::vector<SV> M((n + 1) * chunks, A.zero());
    auto sat_sub16 = [](uint16_t a, uint16_t b) { return (uint16_t)(a > b ? a - b : 0); };
    auto sat_add16 = [](uint16_t a, uint16_t b) { uint32_t r = (uint32_t)a + b; return (uint16_t)(r > 0xFFFF ? 0xFFFF : r); };
    SV gex = A.splat(sc.gap_extend_penalty);
    SV gop = A.splat(sat_sub16(sc.gap_open_penalty, sc.gap_extend_penalty));
    SV match_score = A.splat(sat_add16(sc.match_score, sc.mismatch_penalty));
    SV mismatch = A.splat(sc.mismatch_penalty);
    SV case_bonus = A.splat(sc.matching_case_bonus);
    SV cap_bonus = A.splat(sc.capitalization_bonus);
    SV delim_bonus = A.splat(sc.delimiter_bonus);
    SV prefix_masked = include_prefix ? A.first_lane(sc.prefix_bonus) : A.zero();
    bool prev_chunk_last_delim = false, prev_chunk_last_lower = false;
    SV maxv = A.zero();
    const uint16_t FULL = A.full();
    for (size_t col = 1; col < chunks; col++) {
        // load_partial: zero-filled tail (scalar.rs:78-85)
        uint8_t b[kMaxLanes];
        for (int i = 0; i < lanes; i++) {
            size_t pos = (col - 1) * lanes + i;
            b[i] = pos < hl ? hay[pos] : 0;
        }
        bool up[kMaxLanes] = {false}, lo[kMaxLanes] = {false}, dl[kMaxLanes] = {false};
        for (int i = 0; i < lanes; i++) {
            up[i] = b[i] < 'Z' + 1 && b[i] > 'A' - 1;
            lo[i] = b[i] < 'z' + 1 && b[i] > 'a' - 1;
            bool digit = b[i] > '0' - 1 && b[i] < '9' + 1;
            dl[i] = !(up[i] || lo[i] || digit || b[i] > 127);
        }
        SV cap_m = A.zero(), delim_m = A.zero();
        for (int i = 0; i < lanes; i++) {
            bool prev_lower = i == 0 ? prev_chunk_last_lower : lo[i - 1];
            bool prev_delim = i == 0 ? prev_chunk_last_delim : dl[i - 1];
            cap_m.v[i] = (up[i] && prev_lower) ? FULL : 0;
            delim_m.v[i] = (prev_delim && !dl[i]) ? FULL : 0;
        }
        prev_chunk_last_lower = lo[lanes - 1];
        prev_chunk_last_delim = dl[lanes - 1];
        SV bonuses = A.add(A.add(A.add(A.band(delim_m, delim_bonus), A.band(cap_m, cap_bonus)), prefix_masked), match_score);

        SV up_gap_mask = A.zero();
        SV prev_row = A.zero();
        SV row = A.zero();
        for (size_t r = 1; r <= n; r++) {
            SV mm = A.zero(), ex = A.zero();
            for (int i = 0; i < lanes; i++) {
                bool e = b[i] == needle[r - 1].c;
                bool f = b[i] == needle[r - 1].flip;
                mm.v[i] = (e || f) ? FULL : 0;
                ex.v[i] = e ? FULL : 0;
            }
            SV diag = A.srp(prev_row, H[(r - 1) * chunks + (col - 1)], 1);
            diag = A.add(diag, A.band(mm, bonuses));
            diag = A.subs(diag, mismatch);
            diag = A.add(diag, A.band(ex, case_bonus));
            SV upv = A.subs(A.subs(prev_row, gex), A.band(up_gap_mask, gop));
            row = propagate(A, A.max(diag, upv), H[r * chunks + (col - 1)], mm, M[r * chunks + (col - 1)], gop, gex);
            H[r * chunks + col] = row;
            M[r * chunks + col] = mm;
            prev_row = row;
            up_gap_mask = mm;
        }
        if (max_cols != SIZE_MAX)
            for (int i = 0; i < lanes; i++)
                if ((col - 1) * lanes + i >= max_cols) row.v[i] = 0;
        maxv = A.max(maxv, row);
        prefix_masked = A.zero();
    }
    if (mats) { mats->H = std::move(H); mats->M = std::move(M); mats->chunks = chunks; }
    return A.hmax(maxv);
}

// src/smith_waterman/mod.rs:91-116 (score_fits_in_u8)
static size_t max_per_char_bonus(const frz_scoring& s) {  // src/lib.rs:488-494
    uint16_t bonus = std::max(s.delimiter_bonus, s.capitalization_bonus);
    uint16_t amort = std::max<uint16_t>((uint16_t)((bonus + 1) / 2), bonus > s.gap_open_penalty ? bonus - s.gap_open_penalty : 0);
    uint32_t r = (uint32_t)amort + s.matching_case_bonus;
    return r > 0xFFFF ? 0xFFFF : r;
}
static size_t max_one_time_bonus(const frz_scoring& s) {  // src/lib.rs:497-503
    uint16_t bonus = std::max(s.delimiter_bonus, s.capitalization_bonus);
    uint16_t amort = std::max<uint16_t>((uint16_t)((bonus + 1) / 2), bonus > s.gap_open_penalty ? bonus - s.gap_open_penalty : 0);
    return bonus - amort;
}
static bool score_fits_in_u8(size_t needle_len, const frz_scoring& s) {
    size_t max_constant = (size_t)s.match_score + s.mismatch_penalty;
    max_constant = std::max<size_t>(max_constant, s.gap_open_penalty);
    max_constant = std::max<size_t>(max_constant, s.gap_extend_penalty);
    max_constant = std::max<size_t>(max_constant, s.matching_case_bonus);
    max_constant = std::max<size_t>(max_constant, s.capitalization_bonus);
    max_constant = std::max<size_t>(max_constant, s.delimiter_bonus);
    max_constant = std::max<size_t>(max_constant, s.prefix_bonus);
    if (max_constant > 255) return false;
    size_t max_gap = 64 * (size_t)s.gap_extend_penalty + s.gap_open_penalty;
    if (max_gap > 255) return false;
    size_t max_per_char = (size_t)s.match_score + max_per_char_bonus(s);
    size_t max_matrix = max_per_char * needle_len + max_one_time_bonus(s) + s.prefix_bonus;
    return max_matrix + s.mismatch_penalty <= 255;
}

// src/lib.rs:368-377 (CaseMatching::respects_case_for); ASCII needles only
static bool respects_case(int casing, const uint8_t* needle, size_t n) {
    if (casing == FRZ_CASE_IGNORE) return false;
    if (casing == FRZ_CASE_RESPECT) return true;
    return needle_has_uppercase(needle, n);   // needle.chars().any(char::is_uppercase), src/lib.rs:370-376
}

// ---------------------------------------------------------------------------------
// Unicode Smith-Waterman: src/smith_waterman/algo/unicode.rs + unicode_gap.rs.
// Rows are needle scalars, columns stay haystack BYTES; continuation bytes are transport lanes.
// ---------------------------------------------------------------------------------
// unicode_gap.rs:133-165 (unicode_gap_step)
static void unicode_gap_step(const Arith& A, int shift, SV& row, SV& pend, const SV& adj_row, const SV& adj_pend,
                             const SV& cgex, const SV& send, const SV& total, const SV& gop) {
    SV shifted_row = A.srp(row, adj_row, shift);
    SV shifted_pend = A.srp(pend, adj_pend, shift);
    SV scalar_gex = A.subs(total, cgex);
    SV crossed = A.band(shifted_pend, send);
    SV pen = A.add(scalar_gex, A.band(gop, crossed));
    SV cand = A.subs(shifted_row, pen);
    row = A.max(row, cand);
    SV cand_pend = A.subs(shifted_pend, send);
    pend = A.max(pend, cand_pend);
}
// unicode_gap.rs:168-194 (prepare_next_unicode_gap_step)
static void unicode_gap_prepare(const Arith& A, int shift, SV& cgex, SV& adj_cgex, SV& send, SV& adj_send, SV& total) {
    SV zero = A.zero();
    SV shifted_cgex = A.srp(cgex, adj_cgex, shift);
    cgex = A.add(cgex, shifted_cgex);
    adj_cgex = A.add(adj_cgex, A.srp(adj_cgex, zero, shift));
    SV shifted_send = A.srp(send, adj_send, shift);
    send = A.max(send, shifted_send);
    adj_send = A.max(adj_send, A.srp(adj_send, zero, shift));
    total = A.add(total, total);
}
// unicode_gap.rs:196-262 (unicode_propagator!: prepare shifts 1..LANES/4, final shift LANES/2)
static void propagate_unicode(const Arith& A, SV& row, SV& pend, const SV& adj_row, const SV& adj_pend, SV cgex, SV adj_cgex,
                              SV send, SV adj_send, const SV& gop, const SV& gex) {
    SV total = gex;
    int s = 1;
    for (; s < A.lanes / 2; s <<= 1) {
        unicode_gap_step(A, s, row, pend, adj_row, adj_pend, cgex, send, total, gop);
        unicode_gap_prepare(A, s, cgex, adj_cgex, send, adj_send, total);
    }
    unicode_gap_step(A, s, row, pend, adj_row, adj_pend, cgex, send, total, gop);
}

// unicode.rs:9-224 (score_haystack_unicode)
static uint16_t sw_score_unicode(const uint8_t* needle_raw, size_t nbytes, const frz_scoring& sc, bool case_sensitive,
                                 const uint8_t* hay, size_t hl, bool include_prefix, int lanes, bool u8, SwMatrices* mats = nullptr) {
    if (hl > kMaxHaystackLen) {
        int g = match_greedy(needle_raw, nbytes, hay, hl, sc, case_sensitive, include_prefix);
        return g < 0 ? 0 : (uint16_t)g;
    }
    std::vector<UChar> needle = case_needle_unicode(needle_raw, nbytes, case_sensitive);
    size_t n = needle.size();
    if (n == 0) return 0;
    Arith A{lanes, u8};
    size_t chunks = (hl + lanes - 1) / lanes + 1;
    std::vector<SV> H((n + 1) * chunks, A.zero());
    std::vector<SV> Mm((n + 1) * chunks, A.zero());
    std::vector<SV> pending(n + 1, A.zero());
    auto sat_sub16 = [](uint16_t a, uint16_t b) { return (uint16_t)(a > b ? a - b : 0); };
    auto sat_add16 = [](uint16_t a, uint16_t b) { uint32_t r = (uint32_t)a + b; return (uint16_t)(r > 0xFFFF ? 0xFFFF : r); };
    SV gex = A.splat(sc.gap_extend_penalty);
    SV gop = A.splat(sat_sub16(sc.gap_open_penalty, sc.gap_extend_penalty));
    SV match_score = A.splat(sat_add16(sc.match_score, sc.mismatch_penalty));
    SV mismatch = A.splat(sc.mismatch_penalty);
    SV case_bonus = A.splat(sc.matching_case_bonus);
    SV cap_bonus = A.splat(sc.capitalization_bonus);
    SV delim_bonus = A.splat(sc.delimiter_bonus);
    SV prefix_masked = include_prefix ? A.first_lane(sc.prefix_bonus) : A.zero();
    const uint16_t FULL = A.full();
    bool prev_chunk_last_delim = false, prev_chunk_last_lower = false;
    SV prev_cgex = A.zero(), prev_sstart = A.zero();
    SV maxv = A.zero();
    for (size_t col = 1; col < chunks; col++) {
        size_t cs = (col - 1) * lanes;
        // load_partial at chunk_start + k, zero-filled past the end (scalar.rs:78-85)
        uint8_t b[4][kMaxLanes];
        for (int k = 0; k < 4; k++)
            for (int i = 0; i < lanes; i++) {
                size_t pos = cs + k + i;
                b[k][i] = pos < hl ? hay[pos] : 0;
            }
        // unicode_scalar_masks (unicode.rs:243-259) + valid_haystack_lanes (:262-273)
        size_t valid_lanes = std::min<size_t>(hl > cs ? hl - cs : 0, (size_t)lanes);
        bool cont[kMaxLanes], sst[kMaxLanes];
        SV sstart = A.zero(), cgex = A.zero();
        for (int i = 0; i < lanes; i++) {
            bool valid = (size_t)i < valid_lanes;
            cont[i] = b[0][i] > 0x7f && b[0][i] < 0xc0 && valid;
            sst[i] = !cont[i] && valid;
            sstart.v[i] = sst[i] ? FULL : 0;
            cgex.v[i] = cont[i] ? gex.v[i] : 0;
        }
        bool up[kMaxLanes] = {false}, lo[kMaxLanes] = {false}, dl[kMaxLanes] = {false};
        for (int i = 0; i < lanes; i++) {
            uint8_t c = b[0][i];
            up[i] = c < 'Z' + 1 && c > 'A' - 1;
            lo[i] = c < 'z' + 1 && c > 'a' - 1;
            bool digit = c > '0' - 1 && c < '9' + 1;
            dl[i] = !(up[i] || lo[i] || digit || c > 127);
        }
        SV cap_m = A.zero(), delim_m = A.zero();
        for (int i = 0; i < lanes; i++) {
            bool prev_lower = i == 0 ? prev_chunk_last_lower : lo[i - 1];
            bool prev_delim = i == 0 ? prev_chunk_last_delim : dl[i - 1];
            cap_m.v[i] = (up[i] && prev_lower) ? FULL : 0;
            delim_m.v[i] = (prev_delim && !dl[i]) ? FULL : 0;
        }
        prev_chunk_last_lower = lo[lanes - 1];
        prev_chunk_last_delim = dl[lanes - 1];
        SV bonuses = A.add(A.add(A.add(A.band(delim_m, delim_bonus), A.band(cap_m, cap_bonus)), prefix_masked), match_score);
        prefix_masked = A.zero();

        SV up_gap_mask = A.zero(), prev_row = A.zero(), row = A.zero();
        for (size_t r = 1; r <= n; r++) {
            const UChar& nc = needle[r - 1];
            // unicode_char_match_mask (unicode.rs:227-245): the scalar's bytes at lane offsets 0..len-1, scalar starts only
            SV mm = A.zero(), ex = A.zero();
            for (int i = 0; i < lanes; i++) {
                bool e = sst[i], f = sst[i];
                for (int k = 0; k < nc.len; k++) {
                    e = e && b[k][i] == nc.c[k];
                    f = f && b[k][i] == nc.f[k];
                }
                mm.v[i] = (e || f) ? FULL : 0;
                ex.v[i] = e ? FULL : 0;
            }
            SV diag = A.srp(prev_row, H[(r - 1) * chunks + (col - 1)], 1);
            diag = A.add(diag, A.band(mm, bonuses));
            diag = A.subs(diag, mismatch);
            diag = A.add(diag, A.band(ex, case_bonus));
            diag = A.band(diag, sstart);
            SV upv = A.subs(A.subs(prev_row, gex), A.band(up_gap_mask, gop));
            upv = A.band(upv, sstart);
            row = A.max(diag, upv);
            SV pend = mm;
            propagate_unicode(A, row, pend, H[r * chunks + (col - 1)], pending[r], cgex, prev_cgex, sstart, prev_sstart, gop, gex);
            H[r * chunks + col] = row;
            Mm[r * chunks + col] = mm;
            pending[r] = pend;
            prev_row = row;
            up_gap_mask = mm;
        }
        maxv = A.max(maxv, row);
        prev_cgex = cgex;
        prev_sstart = sstart;
    }
    if (mats) { mats->H = std::move(H); mats->M = std::move(Mm); mats->chunks = chunks; }
    return A.hmax(maxv);
}

// ---------------------------------------------------------------------------------
// Traceback: AlignmentPathIter (src/smith_waterman/alignment_iter.rs) driven by score_haystack_indices /
// score_haystack_unicode_indices (src/smith_waterman/algo/mod.rs:49-151).  Indices come out in reverse order.
// rows = needle bytes (ASCII path) or needle scalars (unicode path: `und` and `uhay` are set).
// ---------------------------------------------------------------------------------
static std::vector<uint32_t> alignment_indices(const SwMatrices& mt, int lanes, size_t rows, size_t start_pos,
                                               const uint8_t* uhay, size_t uhay_len, const std::vector<UChar>* und,
                                               uint16_t score, int max_typos /* < 0: None */) {
    std::vector<uint32_t> indices;
    auto cell = [&](const std::vector<SV>& m, size_t row, size_t col) { return m[row * mt.chunks + col / lanes].v[col % lanes]; };
    // get_col_idx (alignment_iter.rs:73-88): first lane of the final row holding the score
    size_t col = SIZE_MAX;
    for (size_t ch = 1; ch < mt.chunks && col == SIZE_MAX; ch++)
        for (int i = 0; i < lanes; i++)
            if (mt.H[rows * mt.chunks + ch].v[i] == score) { col = ch * lanes + i; break; }
    if (col == SIZE_MAX) return indices;   // (the reference panics; unreachable for a score produced by the same matrix)
    size_t row = rows;
    uint16_t cur = score;
    int typos = 0;
    size_t prev_hidx = SIZE_MAX;
    for (;;) {
        if (row == 0) break;
        if (max_typos >= 0 && typos > max_typos) break;               // Some(None): stop collecting
        if (col < (size_t)lanes || cur == 0) break;                   // left edge / lost alignment (both end the walk)
        const size_t hidx = col - lanes;
        if (uhay && hidx < uhay_len && (uhay[hidx] & 0xC0) == 0x80) {  // continuation byte: walk left
            col -= 1;
            cur = cell(mt.H, row, col);
            continue;
        }
        if (cell(mt.M, row, col) != 0) {
            const size_t needle_idx = row - 1, hpos = hidx + start_pos;
            row -= 1; col -= 1;
            cur = cell(mt.H, row, col);
            if (und) {
                if (prev_hidx != hpos) {
                    const int len = (*und)[needle_idx].len;
                    for (int o = len - 1; o >= 0; o--) indices.push_back((uint32_t)(hpos + o));
                    prev_hidx = hpos;
                }
            } else {
                indices.push_back((uint32_t)hpos);
            }
            continue;
        }
        const uint16_t diag = cell(mt.H, row - 1, col - 1), left = cell(mt.H, row, col - 1), up = cell(mt.H, row - 1, col);
        if (diag >= left && diag >= up) { row -= 1; col -= 1; typos += 1; cur = diag; }
        else if (left >= up) { col -= 1; cur = left; }
        else { typos += 1; row -= 1; cur = up; }
    }
    return indices;
}

// score_haystack_indices / score_haystack_unicode_indices (algo/mod.rs:49-151)
static uint16_t sw_indices(const uint8_t* needle_raw, size_t nbytes, const frz_scoring& sc, bool case_sensitive, bool unicode,
                           const uint8_t* hay, size_t hl, size_t start_pos, int max_typos, int lanes, bool u8,
                           std::vector<uint32_t>* out) {
    out->clear();
    if (hl > kMaxHaystackLen) {
        std::vector<uint32_t> idx;
        int g = match_greedy(needle_raw, nbytes, hay, hl, sc, case_sensitive, start_pos == 0, &idx);
        if (g < 0) return 0;
        for (auto it = idx.rbegin(); it != idx.rend(); ++it) out->push_back((uint32_t)(*it + start_pos));
        return (uint16_t)g;
    }
    SwMatrices mt;
    if (unicode) {
        std::vector<UChar> und = case_needle_unicode(needle_raw, nbytes, case_sensitive);
        uint16_t score = sw_score_unicode(needle_raw, nbytes, sc, case_sensitive, hay, hl, start_pos == 0, lanes, u8, &mt);
        if (score == 0 || und.empty()) return score;
        *out = alignment_indices(mt, lanes, und.size(), start_pos, hay, hl, &und, score, max_typos);
        return score;
    }
    uint16_t score = sw_score(needle_raw, nbytes, sc, case_sensitive, hay, hl, start_pos == 0, lanes, u8, SIZE_MAX, &mt);
    if (score == 0) return score;
    *out = alignment_indices(mt, lanes, nbytes, start_pos, nullptr, 0, nullptr, score, max_typos);
    return score;
}

// ---------------------------------------------------------------------------------
// Literal matcher (ASCII path): src/literal/algo.rs:159-255
// ---------------------------------------------------------------------------------
static bool lit_is_delim(uint8_t b) {  // literal/algo.rs:327-330
    bool alnum = (b >= '0' && b <= '9') || (b >= 'a' && b <= 'z') || (b >= 'A' && b <= 'Z');
    return b <= 127 && !alnum;
}
static bool lit_matches_at(const std::vector<Pair>& nd, const uint8_t* hay, size_t pos) {
    for (size_t k = 0; k < nd.size(); k++) {
        uint8_t b = hay[pos + k];
        if (b != nd[k].c && b != nd[k].flip) return false;
    }
    return true;
}
static uint16_t lit_score_at(const std::vector<Pair>& nd, const frz_scoring& s, const uint8_t* hay, size_t hl, size_t pos) {
    uint16_t score = 0;
    for (size_t k = 0; k < nd.size(); k++) {
        size_t st = pos + k;
        uint16_t sc = s.match_score;
        if (hay[st] == nd[k].c) sc += s.matching_case_bonus;
        if (st == 0) sc += s.prefix_bonus;
        else {
            uint8_t byte = hay[st], prev = hay[st - 1];
            if (byte >= 'A' && byte <= 'Z' && prev >= 'a' && prev <= 'z') sc += s.capitalization_bonus;
            if (lit_is_delim(prev) && !lit_is_delim(byte)) sc += s.delimiter_bonus;
        }
        score += sc;
    }
    if (pos == 0 && nd.size() == hl) score += s.exact_match_bonus;
    return score;
}
// returns false when no match; else pos/score.  Substring: best score, earliest on ties
// (find_substring, literal/algo.rs:262-313 — the seed prefilter only prunes candidates that
// matches_at would reject, so a plain scan over every start position is equivalent).
static bool lit_find(int mode, const std::vector<Pair>& nd, const frz_scoring& s, const uint8_t* hay, size_t hl,
                     size_t* opos, uint16_t* oscore) {
    size_t n = nd.size();
    if (hl < n) return false;
    switch (mode) {
        case FRZ_MATCHING_EXACT:
            if (hl == n && lit_matches_at(nd, hay, 0)) { *opos = 0; *oscore = lit_score_at(nd, s, hay, hl, 0); return true; }
            return false;
        case FRZ_MATCHING_PREFIX:
            if (lit_matches_at(nd, hay, 0)) { *opos = 0; *oscore = lit_score_at(nd, s, hay, hl, 0); return true; }
            return false;
        case FRZ_MATCHING_SUFFIX: {
            size_t pos = hl - n;
            if (lit_matches_at(nd, hay, pos)) { *opos = pos; *oscore = lit_score_at(nd, s, hay, hl, pos); return true; }
            return false;
        }
        case FRZ_MATCHING_SUBSTRING: {
            bool have = false;
            for (size_t pos = 0; pos + n <= hl; pos++) {
                if (!lit_matches_at(nd, hay, pos)) continue;
                uint16_t sc = lit_score_at(nd, s, hay, hl, pos);
                if (!have || sc > *oscore) { have = true; *opos = pos; *oscore = sc; }
            }
            return have;
        }
    }
    return false;
}

// Literal matcher, unicode path (literal/algo.rs:159-230 with UNICODE = true): whole scalars compare against the
// original or the case-flipped encoding, and each scalar is scored once at its first byte.
static bool ulit_matches_at(const std::vector<UChar>& nd, const uint8_t* hay, size_t pos) {
    size_t k = pos;
    for (const UChar& c : nd) {
        if (memcmp(hay + k, c.c, c.len) != 0 && memcmp(hay + k, c.f, c.len) != 0) return false;
        k += c.len;
    }
    return true;
}
static uint16_t ulit_score_at(const std::vector<UChar>& nd, size_t needle_bytes, const frz_scoring& s, const uint8_t* hay, size_t hl,
                              size_t pos) {
    uint16_t score = 0;
    size_t st = pos;
    for (const UChar& c : nd) {
        uint16_t sc = s.match_score;
        if (memcmp(hay + st, c.c, c.len) == 0) sc += s.matching_case_bonus;
        if (st == 0) sc += s.prefix_bonus;
        else {
            uint8_t byte = hay[st], prev = hay[st - 1];
            if (byte >= 'A' && byte <= 'Z' && prev >= 'a' && prev <= 'z') sc += s.capitalization_bonus;
            if (lit_is_delim(prev) && !lit_is_delim(byte)) sc += s.delimiter_bonus;
        }
        score += sc;
        st += c.len;
    }
    if (pos == 0 && needle_bytes == hl) score += s.exact_match_bonus;
    return score;
}
static bool ulit_find(int mode, const std::vector<UChar>& nd, size_t n, const frz_scoring& s, const uint8_t* hay, size_t hl,
                      size_t* opos, uint16_t* oscore) {
    if (hl < n) return false;
    auto at = [&](size_t pos) {
        if (!ulit_matches_at(nd, hay, pos)) return false;
        *opos = pos; *oscore = ulit_score_at(nd, n, s, hay, hl, pos);
        return true;
    };
    switch (mode) {
        case FRZ_MATCHING_EXACT: return hl == n && at(0);
        case FRZ_MATCHING_PREFIX: return at(0);
        case FRZ_MATCHING_SUFFIX: return at(hl - n);
        case FRZ_MATCHING_SUBSTRING: {
            bool have = false;
            size_t bp = 0; uint16_t bs = 0;
            for (size_t pos = 0; pos + n <= hl; pos++) {
                if (!ulit_matches_at(nd, hay, pos)) continue;
                uint16_t sc = ulit_score_at(nd, n, s, hay, hl, pos);
                if (!have || sc > bs) { have = true; bp = pos; bs = sc; }
            }
            if (have) { *opos = bp; *oscore = bs; }
            return have;
        }
    }
    return false;
}

// One compiled pattern (src/matcher/mod.rs:193-205 compile + get_backend :448-498)
struct OPattern {
    std::vector<uint8_t> needle;
    bool negated;
    int max_typos;  // -1 = None
    int matching;
    bool case_sensitive;
    frz_scoring scoring;
    int lanes;        // SW lanes for the selected width
    int pf_lanes;     // prefilter lanes
    bool u8;
    size_t min_hay_len;
    bool unicode;     // UnicodeMatching::respects_unicode_for(needle): the UNICODE = true specialisations
};

// src/matcher/algo.rs:78-103 (match_list_into_impl) + :229-263 (smith_waterman_one) + :331-338 (trim)
// and src/literal/algo.rs:84-116 for literal modes.
static void match_list_into(const OPattern& p, const uint8_t* bytes, const uint64_t* offsets, uint64_t n,
                            uint32_t index_offset, std::vector<frz_match>& out) {
    std::vector<Pair> nd = case_needle(p.needle.data(), p.needle.size(), p.case_sensitive);
    std::vector<UChar> und = case_needle_unicode(p.needle.data(), p.needle.size(), p.case_sensitive);
    for (uint64_t i = 0; i < n; i++) {
        const uint8_t* hay = bytes + offsets[i];
        size_t hl = (size_t)(offsets[i + 1] - offsets[i]);
        uint32_t index = index_offset + (uint32_t)i;
        if (p.matching != FRZ_MATCHING_FUZZY) {
            size_t pos; uint16_t sc;
            bool hit = p.unicode ? ulit_find(p.matching, und, p.needle.size(), p.scoring, hay, hl, &pos, &sc)
                                 : lit_find(p.matching, nd, p.scoring, hay, hl, &pos, &sc);
            if (hit) out.push_back({index, sc, (uint8_t)(pos == 0 && p.needle.size() == hl), 0});
            continue;
        }
        if (hl < p.min_hay_len) continue;
        size_t s, e;
        bool pass = p.unicode ? uprefilter(und, hay, hl, p.max_typos, p.pf_lanes, &s, &e)
                              : prefilter(nd, hay, hl, p.max_typos, p.pf_lanes, &s, &e);
        if (!pass) continue;
        s = s > 0 ? s - 1 : 0;
        bool include_exact = s == 0 && e == hl;
        const uint8_t* w = hay + s;
        size_t wl = e - s;
        uint16_t score = p.unicode
            ? sw_score_unicode(p.needle.data(), p.needle.size(), p.scoring, p.case_sensitive, w, wl, s == 0, p.lanes, p.u8)
            : sw_score(p.needle.data(), p.needle.size(), p.scoring, p.case_sensitive, w, wl, s == 0, p.lanes, p.u8);
        bool exact = include_exact && wl == p.needle.size() && memcmp(w, p.needle.data(), wl) == 0;
        if (exact) score = (uint16_t)(score + p.scoring.exact_match_bonus);
        out.push_back({index, score, (uint8_t)exact, 0});
    }
}

// MatcherImpl::match_one_indices_impl (src/matcher/algo.rs:138-169) / LiteralImpl::match_one_indices_impl
// (src/literal/algo.rs:134-155) for one haystack: false = no match.
static bool match_one_indices(const OPattern& p, const uint8_t* hay, size_t hl, uint32_t index, frz_match* m,
                              std::vector<uint32_t>* indices) {
    indices->clear();
    std::vector<Pair> nd = case_needle(p.needle.data(), p.needle.size(), p.case_sensitive);
    std::vector<UChar> und = case_needle_unicode(p.needle.data(), p.needle.size(), p.case_sensitive);
    if (p.matching != FRZ_MATCHING_FUZZY) {
        size_t pos; uint16_t sc;
        bool hit = p.unicode ? ulit_find(p.matching, und, p.needle.size(), p.scoring, hay, hl, &pos, &sc)
                             : lit_find(p.matching, nd, p.scoring, hay, hl, &pos, &sc);
        if (!hit) return false;
        *m = {index, sc, (uint8_t)(pos == 0 && p.needle.size() == hl), 0};
        for (size_t i = pos + p.needle.size(); i > pos; i--) indices->push_back((uint32_t)(i - 1));
        return true;
    }
    if (hl < p.min_hay_len) return false;
    size_t s, e;
    bool pass = p.unicode ? uprefilter(und, hay, hl, p.max_typos, p.pf_lanes, &s, &e)
                          : prefilter(nd, hay, hl, p.max_typos, p.pf_lanes, &s, &e);
    if (!pass) return false;
    s = s > 0 ? s - 1 : 0;
    bool include_exact = s == 0 && e == hl;
    const uint8_t* w = hay + s;
    size_t wl = e - s;
    uint16_t score = sw_indices(p.needle.data(), p.needle.size(), p.scoring, p.case_sensitive, p.unicode, w, wl, s, p.max_typos,
                                p.lanes, p.u8, indices);
    bool exact = include_exact && wl == p.needle.size() && memcmp(w, p.needle.data(), wl) == 0;
    if (exact) score = (uint16_t)(score + p.scoring.exact_match_bonus);
    *m = {index, score, (uint8_t)exact, 0};
    return true;
}

// src/sort.rs:6-40 (radix_sort_matches) — literal 2-pass LSD radix
static void radix_sort_matches(frz_match* m, size_t n) {
    std::vector<frz_match> b(n);
    uint32_t hist[256] = {0}, off[256] = {0};
    for (size_t i = 0; i < n; i++) hist[m[i].score & 0xFF]++;
    for (int idx = 255; idx >= 1; idx--) off[idx - 1] = off[idx] + hist[idx];
    for (size_t i = 0; i < n; i++) b[off[m[i].score & 0xFF]++] = m[i];
    memset(hist, 0, sizeof hist);
    for (size_t i = 0; i < n; i++) hist[(b[i].score >> 8) & 0xFF]++;
    off[255] = 0;
    for (int idx = 255; idx >= 1; idx--) off[idx - 1] = off[idx] + hist[idx];
    for (size_t i = 0; i < n; i++) m[off[(b[i].score >> 8) & 0xFF]++] = b[i];
}

// src/matcher/multi.rs:84-152 (match_list_multi_into)
static void match_list_multi_into(const std::vector<OPattern>& pats, const uint8_t* bytes, const uint64_t* offsets,
                                  uint64_t n, uint32_t index_offset, std::vector<frz_match>& out) {
    int base = -1;
    for (size_t i = 0; i < pats.size(); i++)
        if (!pats[i].negated) { base = (int)i; break; }
    std::vector<frz_match> cand;
    if (base >= 0) match_list_into(pats[base], bytes, offsets, n, index_offset, cand);
    else
        for (uint64_t i = 0; i < n; i++) cand.push_back({index_offset + (uint32_t)i, 0, 0, 0});
    for (size_t pi = 0; pi < pats.size(); pi++) {
        if ((int)pi == base || cand.empty()) continue;
        // gather candidates
        std::vector<uint8_t> gb;
        std::vector<uint64_t> go{0};
        for (auto& c : cand) {
            uint64_t i = c.index - index_offset;
            gb.insert(gb.end(), bytes + offsets[i], bytes + offsets[i + 1]);
            go.push_back(gb.size());
        }
        std::vector<frz_match> hits;
        match_list_into(pats[pi], gb.data(), go.data(), cand.size(), 0, hits);
        if (pats[pi].negated) {
            std::vector<frz_match> keep;
            size_t h = 0;
            for (size_t pos = 0; pos < cand.size(); pos++) {
                bool matched = h < hits.size() && hits[h].index == pos;
                if (matched) h++;
                else keep.push_back(cand[pos]);
            }
            cand.swap(keep);
        } else {
            std::vector<frz_match> next;
            for (auto hit : hits) {
                frz_match c = cand[hit.index];
                hit.index = c.index;
                uint32_t s = (uint32_t)hit.score + c.score;
                hit.score = (uint16_t)(s > 0xFFFF ? 0xFFFF : s);
                hit.exact |= c.exact;
                next.push_back(hit);
            }
            cand.swap(next);
        }
    }
    out.insert(out.end(), cand.begin(), cand.end());
}

// src/matcher/mod.rs:193-205,448-498 + src/pattern.rs:250-262 (resolve)
static bool compile(const frz_pattern& src, const frz_config& cfg, OPattern* o) {
    if (src.needle_len == 0) return false;
    o->needle.assign(src.needle, src.needle + src.needle_len);
    o->negated = src.negated != 0;
    o->max_typos = src.max_typos >= 0 ? src.max_typos : cfg.max_typos;
    int casing = src.casing >= 0 ? src.casing : cfg.casing;
    o->matching = src.matching >= 0 ? src.matching : cfg.matching;
    o->scoring = src.has_scoring ? src.scoring : cfg.scoring;
    o->case_sensitive = respects_case(casing, src.needle, src.needle_len);
    int unicode = src.unicode >= 0 ? src.unicode : cfg.unicode;
    bool ascii = true;
    size_t nchars = 0;
    for (size_t i = 0; i < src.needle_len; i++) {
        if (src.needle[i] >= 0x80) ascii = false;
        if ((src.needle[i] & 0xC0) != 0x80) nchars++;
    }
    // UnicodeMatching::respects_unicode_for (src/lib.rs:394-401)
    o->unicode = (unicode == FRZ_UNICODE_ALWAYS) || (unicode == FRZ_UNICODE_SMART && !ascii);
    int lanes8 = cfg.emulate_lanes ? cfg.emulate_lanes : 64;
    o->u8 = score_fits_in_u8(src.needle_len, o->scoring);
    o->lanes = o->u8 ? lanes8 : lanes8 / 2;
    o->pf_lanes = lanes8;  // Prefilter{AVX512,AVX,SSE,Scalar} LANES = 64/32/16/16
    // min_haystack_len (src/matcher/algo.rs:62-65): needle.chars().count() - max_typos
    o->min_hay_len = o->max_typos >= 0 ? (nchars > (size_t)o->max_typos ? nchars - o->max_typos : 0) : 0;
    return true;
}

}  // namespace

extern "C" {

// ---- unit-level entry points (used by the KAT tests) ----

int frzo_prefilter(const uint8_t* needle, size_t n, int case_sensitive, const uint8_t* hay, size_t len,
                   int max_typos, int lanes, uint64_t* start, uint64_t* end) {
    std::vector<Pair> nd = case_needle(needle, n, case_sensitive != 0);
    size_t s = 0, e = 0;
    bool ok = prefilter(nd, hay, len, max_typos, lanes, &s, &e);
    *start = s;
    *end = e;
    return ok ? 1 : 0;
}

uint16_t frzo_sw_score(const uint8_t* needle, size_t n, const frz_scoring* sc, int case_sensitive,
                       const uint8_t* hay, size_t len, int include_prefix, int lanes, int score_bits) {
    return sw_score(needle, n, *sc, case_sensitive != 0, hay, len, include_prefix != 0, lanes, score_bits == 8);
}

// unicode-needle unit entry points (src/prefilter/algo/unicode*.rs, src/smith_waterman/algo/unicode.rs)
int frzo_prefilter_unicode(const uint8_t* needle, size_t n, int case_sensitive, const uint8_t* hay, size_t len,
                           int max_typos, int lanes, uint64_t* start, uint64_t* end) {
    std::vector<UChar> nd = case_needle_unicode(needle, n, case_sensitive != 0);
    size_t s = 0, e = 0;
    bool ok = uprefilter(nd, hay, len, max_typos, lanes, &s, &e);
    *start = s;
    *end = e;
    return ok ? 1 : 0;
}

uint16_t frzo_sw_score_unicode(const uint8_t* needle, size_t n, const frz_scoring* sc, int case_sensitive,
                               const uint8_t* hay, size_t len, int include_prefix, int lanes, int score_bits) {
    return sw_score_unicode(needle, n, *sc, case_sensitive != 0, hay, len, include_prefix != 0, lanes, score_bits == 8);
}

// score_haystack_indices / score_haystack_unicode_indices on a window: returns the score, writes the indices (reverse order)
uint16_t frzo_sw_indices(const uint8_t* needle, size_t n, const frz_scoring* sc, int case_sensitive, int unicode,
                         const uint8_t* hay, size_t len, uint64_t start_pos, int max_typos, int lanes, int score_bits,
                         uint32_t* out, uint32_t cap, uint32_t* n_out) {
    std::vector<uint32_t> idx;
    uint16_t score = sw_indices(needle, n, *sc, case_sensitive != 0, unicode != 0, hay, len, (size_t)start_pos, max_typos, lanes,
                                score_bits == 8, &idx);
    *n_out = (uint32_t)idx.size();
    for (size_t i = 0; i < idx.size() && i < cap; i++) out[i] = idx[i];
    return score;
}

// Matcher::match_list_indices restricted to the haystacks `which[0..n_which)`
// (src/matcher/mod.rs:234-262 → match_one_indices_impl / match_one_indices_multi).  out_cnt[j] = number of indices, or 0xFFFFFFFF when haystack
// which[j] does not match; out_idx[j * stride ...] in the reference's (reverse) order.  Returns 0, or -1 on a bad pattern.
int frzo_match_indices(const frz_pattern* patterns, size_t np, const frz_config* cfg, const uint8_t* bytes, const uint64_t* offsets,
                       const uint32_t* which, uint64_t n_which, uint32_t* out_idx, uint32_t stride, uint32_t* out_cnt,
                       frz_match* out_match) {
    std::vector<OPattern> pats;
    for (size_t i = 0; i < np; i++) {
        OPattern o;
        if (compile(patterns[i], *cfg, &o)) pats.push_back(o);
    }
    if (pats.empty()) return -1;
    std::vector<uint32_t> idx, all;
    for (uint64_t j = 0; j < n_which; j++) {
        const uint64_t i = which[j];
        const uint8_t* hay = bytes + offsets[i];
        const size_t hl = (size_t)(offsets[i + 1] - offsets[i]);
        // Single: match_one_indices_impl; Multi: match_one_indices_multi (src/matcher/multi.rs:56-79)
        frz_match comb{(uint32_t)i, 0, 0, 0};
        all.clear();
        bool alive = true;
        for (const OPattern& o : pats) {
            frz_match m{};
            const bool hit = match_one_indices(o, hay, hl, (uint32_t)i, &m, &idx);
            if (o.negated) { if (hit) { alive = false; break; } continue; }
            if (!hit) { alive = false; break; }
            uint32_t sum = (uint32_t)comb.score + m.score;
            comb.score = (uint16_t)(sum > 0xFFFF ? 0xFFFF : sum);
            comb.exact |= m.exact;
            all.insert(all.end(), idx.begin(), idx.end());
        }
        if (!alive) { out_cnt[j] = 0xFFFFFFFFu; continue; }
        if (pats.size() > 1) {   // sort_unstable_by(b.cmp(a)) + dedup
            std::sort(all.begin(), all.end(), [](uint32_t a, uint32_t b) { return a > b; });
            all.erase(std::unique(all.begin(), all.end()), all.end());
        }
        out_match[j] = comb;
        out_cnt[j] = (uint32_t)all.size();
        for (size_t k = 0; k < all.size() && k < stride; k++) out_idx[j * (uint64_t)stride + k] = all[k];
    }
    return 0;
}

// case_needle_unicode for one scalar: writes the flipped scalar's UTF-8 (same length as the input) and returns that length
int frzo_flip_scalar(const uint8_t* utf8, size_t n, uint8_t* out) {
    std::vector<UChar> nd = case_needle_unicode(utf8, n, false);
    if (nd.size() != 1) return -1;
    memcpy(out, nd[0].f, nd[0].len);
    return nd[0].len;
}

uint16_t frzo_sw_score_col_limit(const uint8_t* needle, size_t n, const frz_scoring* sc, int case_sensitive,
                                 const uint8_t* hay, size_t len, int include_prefix, int lanes, int score_bits, uint64_t max_cols) {
    return sw_score(needle, n, *sc, case_sensitive != 0, hay, len, include_prefix != 0, lanes, score_bits == 8, (size_t)max_cols);
}

// Randomised search for a window whose score changes when the final maximum ignores the columns >= len + n + slack
// (dense alphabets, the shapes that expose the reference's lane dependence).  Returns the number of counterexamples.
uint64_t frzo_col_limit_search(uint64_t seed, uint64_t trials, int lanes, int score_bits, int slack, const frz_scoring* sc,
                               uint8_t* bad_needle, uint8_t* bad_hay, uint32_t* bad_lens) {
    uint64_t x = seed * 0x9E3779B97F4A7C15ull + 1, bad = 0;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    static const char* pools[] = {"abAB_/-ab01", "ab", "aAbB_", "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789_-/."};
    for (uint64_t t = 0; t < trials; t++) {
        const char* pool = pools[rnd() % 4];
        const size_t pl = strlen(pool);
        const size_t n = 1 + rnd() % (score_bits == 8 ? 13 : 24);
        const size_t hl = 1 + rnd() % (size_t)(2 * lanes + 7);
        uint8_t needle[64], hay[256];
        for (size_t i = 0; i < n; i++) needle[i] = (uint8_t)pool[rnd() % pl];
        for (size_t i = 0; i < hl; i++) hay[i] = (uint8_t)pool[rnd() % pl];
        const bool cs = rnd() & 1, pref = rnd() & 1;
        const uint16_t full = sw_score(needle, n, *sc, cs, hay, hl, pref, lanes, score_bits == 8);
        const uint16_t lim = sw_score(needle, n, *sc, cs, hay, hl, pref, lanes, score_bits == 8, hl + n + slack);
        if (full != lim) {
            if (bad == 0 && bad_needle) {
                memcpy(bad_needle, needle, n); memcpy(bad_hay, hay, hl);
                bad_lens[0] = (uint32_t)n; bad_lens[1] = (uint32_t)hl; bad_lens[2] = full; bad_lens[3] = lim;
            }
            bad++;
        }
    }
    return bad;
}

int frzo_match_greedy(const uint8_t* needle, size_t n, const frz_scoring* sc, int case_sensitive,
                      const uint8_t* hay, size_t len, int include_prefix) {
    return match_greedy(needle, n, hay, len, *sc, case_sensitive != 0, include_prefix != 0);
}

int frzo_score_fits_in_u8(size_t needle_len, const frz_scoring* sc) { return score_fits_in_u8(needle_len, *sc) ? 1 : 0; }

void frzo_radix_sort_matches(frz_match* m, uint64_t n) { radix_sort_matches(m, (size_t)n); }

// ---- pipeline entry points ----

// Matcher::match_list_into on all patterns (index order, no sort).  Returns the number of
// matches (which may exceed cap; only the first cap are written).
uint64_t frzo_match_list_into(const frz_pattern* patterns, size_t np, const frz_config* cfg, const uint8_t* bytes,
                              const uint64_t* offsets, uint64_t n, uint32_t index_offset, frz_match* out, uint64_t cap) {
    std::vector<OPattern> pats;
    for (size_t i = 0; i < np; i++) {
        OPattern o;
        if (compile(patterns[i], *cfg, &o)) {
            pats.push_back(o);
        }
    }
    std::vector<frz_match> res;
    if (pats.empty()) {  // CompiledPatterns::Empty (src/matcher/mod.rs:380-383)
        for (uint64_t i = 0; i < n; i++) res.push_back({index_offset + (uint32_t)i, 0, 0, 0});
    } else if (pats.size() == 1 && !pats[0].negated) {
        match_list_into(pats[0], bytes, offsets, n, index_offset, res);
    } else {
        match_list_multi_into(pats, bytes, offsets, n, index_offset, res);
    }
    for (uint64_t i = 0; i < res.size() && i < cap; i++) out[i] = res[i];
    return res.size();
}

// Matcher::match_list (src/matcher/mod.rs:212-222): into + reverse? + radix sort?
uint64_t frzo_match_list(const frz_pattern* patterns, size_t np, const frz_config* cfg, const uint8_t* bytes,
                         const uint64_t* offsets, uint64_t n, frz_match* out, uint64_t cap) {
    std::vector<frz_match> tmp(n ? n : 1);
    uint64_t cnt = frzo_match_list_into(patterns, np, cfg, bytes, offsets, n, 0, tmp.data(), n);
    if (cnt == UINT64_MAX) return cnt;
    bool empty = true;
    for (size_t i = 0; i < np; i++)
        if (patterns[i].needle_len) empty = false;
    bool reversed = cfg->sort == FRZ_SORT_INDEX_DESC || cfg->sort == FRZ_SORT_SCORE_THEN_INDEX_DESC;
    bool by_score = cfg->sort == FRZ_SORT_SCORE_THEN_INDEX_ASC || cfg->sort == FRZ_SORT_SCORE_THEN_INDEX_DESC;
    if (reversed) std::reverse(tmp.begin(), tmp.begin() + cnt);
    if (!empty && by_score) radix_sort_matches(tmp.data(), cnt);
    for (uint64_t i = 0; i < cnt && i < cap; i++) out[i] = tmp[i];
    return cnt;
}

}  // extern "C"
