// indices_path.cuh — matched-character indices: the traceback behind
// Matcher::match_list_indices (src/matcher/mod.rs:234-262).
//   src/smith_waterman/algo/mod.rs:49-151   score_haystack_indices / score_haystack_unicode_indices
//   src/smith_waterman/alignment_iter.rs    AlignmentPathIter (walks score_matrix + match_masks backwards)
//   src/smith_waterman/algo/ascii.rs + ascii_gap.rs   the byte scorer, here with the FULL matrices kept
// Like unicode_path.cuh this is `__host__ __device__` code over a byte accessor: the GPU kernel (k_match_indices in
// prefilter.cu) and the CPU test build (tests/harness) compile the same source.  The reference documents this API as
// "not optimized for performance ... use after match_list" (src/matcher/mod.rs:227-229); same here: one thread per
// requested match, matrices in a global scratch.
#pragma once
#include "unicode_path.cuh"

namespace frzi {

using frzu::Ar;
using frzu::srp;

// Byte scorer with the full matrices (ascii.rs:10-158).  H, M: [(n + 1)][cols], cols = (chunks + 1) * lanes, chunk 0 and
// row 0 zero.  c / flip: case_needle pairs.  Returns horizontal_max of the running maximum of the last row.
template <class Hay>
FRZ_HD uint32_t sw_matrices_ascii(const uint8_t* c, const uint8_t* flip, int n, const FrzUScoring& sc, const Hay& hay, int W,
                                  bool include_prefix, int lanes, bool u8, uint16_t* H, uint16_t* M, int cols) {
    const Ar A{lanes, (uint16_t)(u8 ? 0xFF : 0xFFFF)};
    const uint16_t FULL = A.full;
    const int chunks = (W + lanes - 1) / lanes;
    for (int i = 0; i < cols; i++) { H[i] = 0; M[i] = 0; }
    for (int r = 1; r <= n; r++)
        for (int i = 0; i < lanes; i++) { H[r * cols + i] = 0; M[r * cols + i] = 0; }
    const uint16_t gex = (uint16_t)sc.gex, gop = (uint16_t)sc.gopx, mismatch = (uint16_t)sc.mismatch;
    bool prev_last_delim = false, prev_last_lower = false;
    uint16_t maxv[FRZ_U_MAX_LANES];
    for (int i = 0; i < lanes; i++) maxv[i] = 0;
    for (int col = 0; col < chunks; col++) {
        const int cs = col * lanes, base = (col + 1) * lanes, pbase = col * lanes;   // this / previous chunk's column offset
        uint8_t b[FRZ_U_MAX_LANES];
        uint16_t bonuses[FRZ_U_MAX_LANES];
        {
            bool pl = prev_last_lower, pd = prev_last_delim;
            for (int i = 0; i < lanes; i++) {
                const int pos = cs + i;
                b[i] = pos < W ? hay(pos) : (uint8_t)0;   // zero-filled tail (scalar.rs:78-85)
                const bool up = b[i] < 'Z' + 1 && b[i] > 'A' - 1, lo = b[i] < 'z' + 1 && b[i] > 'a' - 1;
                const bool digit = b[i] > '0' - 1 && b[i] < '9' + 1;
                const bool dl = !(up || lo || digit || b[i] > 127);
                uint16_t bn = 0;
                if (pd && !dl) bn = A.add(bn, (uint16_t)sc.delim_bonus);
                if (up && pl) bn = A.add(bn, (uint16_t)sc.cap_bonus);
                if (col == 0 && i == 0 && include_prefix) bn = A.add(bn, (uint16_t)sc.prefix_bonus);
                bonuses[i] = A.add(bn, (uint16_t)sc.match_x);
                pl = lo; pd = dl;
            }
            prev_last_lower = pl;
            prev_last_delim = pd;
        }
        uint16_t prev_row[FRZ_U_MAX_LANES], up_gap[FRZ_U_MAX_LANES], row[FRZ_U_MAX_LANES], mm[FRZ_U_MAX_LANES], t[FRZ_U_MAX_LANES];
        for (int i = 0; i < lanes; i++) { prev_row[i] = 0; up_gap[i] = 0; row[i] = 0; }
        for (int r = 1; r <= n; r++) {
            const uint16_t diag_in = H[(r - 1) * cols + pbase + lanes - 1];   // top lane of (r - 1, previous chunk)
            for (int i = 0; i < lanes; i++) {
                const bool e = b[i] == c[r - 1], f = b[i] == flip[r - 1];
                mm[i] = (e || f) ? FULL : (uint16_t)0;
                uint16_t d = i > 0 ? prev_row[i - 1] : diag_in;
                d = A.add(d, mm[i] & bonuses[i]);
                d = Ar::subs(d, mismatch);
                d = A.add(d, e ? (uint16_t)sc.case_bonus : (uint16_t)0);
                const uint16_t u = Ar::subs(Ar::subs(prev_row[i], gex), up_gap[i] & gop);
                row[i] = Ar::mx(d, u);
            }
            // propagate_N_lane (ascii_gap.rs:11-105): adjacent row / masks = (r, previous chunk), fixed over the steps
            const uint16_t* adj = H + r * cols + pbase;
            const uint16_t* amm = M + r * cols + pbase;
            uint16_t g = gex;
            for (int s = 1; s < lanes; s <<= 1) {
                for (int i = 0; i < lanes; i++) {
                    const uint16_t sh_row = srp(row, adj, lanes, s, i), sh_mm = srp(mm, amm, lanes, s, i);
                    t[i] = Ar::mx(row[i], Ar::subs(sh_row, A.add(g, gop & sh_mm)));
                }
                for (int i = 0; i < lanes; i++) row[i] = t[i];
                g = A.add(g, g);
            }
            for (int i = 0; i < lanes; i++) {
                H[r * cols + base + i] = row[i];
                M[r * cols + base + i] = mm[i];
                prev_row[i] = row[i];
                up_gap[i] = mm[i];
            }
        }
        for (int i = 0; i < lanes; i++) maxv[i] = Ar::mx(maxv[i], row[i]);
    }
    uint16_t m = 0;
    for (int i = 0; i < lanes; i++) m = Ar::mx(m, maxv[i]);
    return m;
}

// AlignmentPathIter + the collecting loops of score_haystack_indices / score_haystack_unicode_indices.
// rows = needle bytes (und == nullptr) or needle scalars.  Writes at most `cap` indices (reverse order), returns the count.
template <class Hay>
FRZ_HD int alignment_indices(const uint16_t* H, const uint16_t* M, int cols, int lanes, int rows, int start_pos, const Hay& hay,
                             int W, const FrzUNeedle* und, uint32_t score, int max_typos, uint32_t* out, int cap) {
    int count = 0;
    // get_col_idx: first lane of the final row (chunks 1..) holding the score
    int col = -1;
    for (int cidx = lanes; cidx < cols; cidx++)
        if (H[rows * cols + cidx] == score) { col = cidx; break; }
    if (col < 0) return 0;
    int row = rows, typos = 0, prev_hpos = -1;
    uint32_t cur = score;
    for (;;) {
        if (row == 0) break;
        if (max_typos >= 0 && typos > max_typos) break;
        if (col < lanes || cur == 0) break;
        const int hidx = col - lanes;
        if (und && hidx < W && (hay(hidx) & 0xC0) == 0x80) {   // continuation byte of a multi-byte scalar: walk left
            col -= 1;
            cur = H[row * cols + col];
            continue;
        }
        if (M[row * cols + col] != 0) {
            const int needle_idx = row - 1, hpos = hidx + start_pos;
            row -= 1; col -= 1;
            cur = H[row * cols + col];
            if (und) {
                if (prev_hpos != hpos) {
                    for (int o = und->len[needle_idx] - 1; o >= 0; o--) { if (count < cap) out[count] = (uint32_t)(hpos + o); count++; }
                    prev_hpos = hpos;
                }
            } else {
                if (count < cap) out[count] = (uint32_t)hpos;
                count++;
            }
            continue;
        }
        const uint32_t diag = H[(row - 1) * cols + col - 1], left = H[row * cols + col - 1], up = H[(row - 1) * cols + col];
        if (diag >= left && diag >= up) { row -= 1; col -= 1; typos += 1; cur = diag; }
        else if (left >= up) { col -= 1; cur = left; }
        else { typos += 1; row -= 1; cur = up; }
    }
    return count < cap ? count : cap;
}

// score_haystack_indices / score_haystack_unicode_indices on the window hay[0..W) that starts at byte `start_pos` of
// the haystack.  scratch: see indices_scratch_elems().  Returns the score; *n_out indices in out[] (reverse order).
FRZ_HD size_t indices_scratch_elems(int rows, int lanes) {
    const size_t cols = (size_t)((FRZ_U_MAX_WINDOW + lanes - 1) / lanes + 1) * lanes;
    return 2 * (size_t)(rows + 1) * cols + 2 * (size_t)(rows + 1) * lanes;
}
template <class Hay>
FRZ_HD uint32_t sw_indices(const FrzUNeedle& nd, bool unicode, const FrzUScoring& sc, const Hay& hay, int W, int start_pos,
                           int max_typos, int lanes, bool u8, uint16_t* scratch, uint32_t* out, int cap, int* n_out) {
    *n_out = 0;
    if (W > FRZ_U_MAX_WINDOW) {   // greedy: positions of the needle bytes, reversed
        uint32_t pos[64];
        const int g = frzu::greedy_score(nd, sc, hay, W, start_pos == 0, pos);
        if (g < 0) return 0;
        int cnt = 0;
        for (int i = nd.nbytes - 1; i >= 0; i--) { if (cnt < cap) out[cnt] = pos[i] + (uint32_t)start_pos; cnt++; }
        *n_out = cnt < cap ? cnt : cap;
        return (uint32_t)g;
    }
    const int rows = unicode ? nd.n : nd.nbytes;
    const int chunks = (W + lanes - 1) / lanes;
    const int cols = (chunks + 1) * lanes;
    uint16_t* H = scratch;
    uint16_t* M = H + (size_t)(rows + 1) * cols;
    uint16_t* rowstate = M + (size_t)(rows + 1) * cols;
    uint32_t score;
    if (unicode) {
        if (nd.n == 0) return 0;
        score = frzu::sw_score(nd, sc, hay, W, start_pos == 0, lanes, u8, rowstate, H, M, cols);
    } else {
        score = sw_matrices_ascii(nd.c, nd.bflip, nd.nbytes, sc, hay, W, start_pos == 0, lanes, u8, H, M, cols);
    }
    if (score == 0) return 0;
    *n_out = alignment_indices(H, M, cols, lanes, rows, start_pos, hay, W, unicode ? &nd : nullptr, score, max_typos, out, cap);
    return score;
}

}  // namespace frzi
