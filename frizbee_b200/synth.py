"""Deterministic synthetic haystack lists shaped like the reference's bench generator
(benches/match_list/generate.rs:48-129): per item None / Partial / Full with P(partial)=0.20,
P(full)=0.05; length = clamp(round(Normal(mu, mu/4)), 1, max_len); alphanumeric filler that avoids
the needle's letters for None/Partial; Full = the needle's bytes interleaved in order with filler.
The reference seeds rand::StdRng(12345); that stream cannot be reproduced without the rand crate,
so this uses numpy's PCG64 with the same seed."""
from __future__ import annotations

from typing import Tuple

import numpy as np

ALNUM = np.frombuffer(b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789", dtype=np.uint8)


UNICODE_SCALARS = ["é".encode(), "ن".encode(), "다".encode(), "😀".encode()]  # generator.rs:94-106


def generate(needle: str, n: int, mu: int, max_len: int, seed: int = 12345, p_partial: float = 0.20,
             p_full: float = 0.05, chunk: int = 1 << 20, alphabet: bytes = None, unicode_frac: float = 0.0,
             prefix_frac: float = 0.0, prefixes=(b"bar", b"Bar")) -> Tuple[np.ndarray, np.ndarray]:
    """Returns Arrow-style (bytes u8[total], offsets u64[n+1]).

    unicode_frac: fraction of items that get 1-3 multibyte scalars spliced in at byte positions that are
    char boundaries of the (ASCII) base string; prefix_frac: fraction of items that start with one of
    `prefixes` (BASELINE.json configs[4]: query 'foo !^bar' on mixed-unicode haystacks)."""
    rng = np.random.default_rng(seed)
    nb = np.frombuffer(needle.encode(), dtype=np.uint8)
    k = len(nb)
    alpha = ALNUM if alphabet is None else np.frombuffer(alphabet, dtype=np.uint8)
    lower = np.frombuffer(needle.lower().encode(), dtype=np.uint8)
    upper = np.frombuffer(needle.upper().encode(), dtype=np.uint8)
    clean = alpha[~np.isin(alpha, np.concatenate([lower, upper]))]
    if clean.size == 0:
        clean = alpha
    parts, lens_all = [], []
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        kind = rng.random(m)
        is_partial = kind < p_partial
        is_full = (~is_partial) & (kind < p_partial + p_full)
        length = np.clip(np.abs(np.rint(rng.normal(mu, mu / 4.0, m))), 1, max_len).astype(np.int64)
        length = np.where(is_full, np.clip(np.maximum(length, k), 1, max_len), length)
        mat = clean[rng.integers(0, clean.size, (m, max_len))]
        full_rows = np.nonzero(is_full)[0]
        if full_rows.size:
            mat[full_rows] = alpha[rng.integers(0, alpha.size, (full_rows.size, max_len))]
        # number of needle bytes to plant per row
        cnt = np.zeros(m, dtype=np.int64)
        cnt[is_full] = np.minimum(k, length[is_full])
        hi = np.minimum(length, k)
        cnt[is_partial] = (rng.random(int(is_partial.sum())) * hi[is_partial]).astype(np.int64)  # 0..min(len,k)-1
        rows = np.nonzero(cnt > 0)[0]
        if rows.size:
            r_len = length[rows]
            r_cnt = cnt[rows]
            # random distinct sorted positions < len: rank random keys, invalid columns pushed to the end
            keys = rng.random((rows.size, max_len), dtype=np.float32)
            keys[np.arange(max_len)[None, :] >= r_len[:, None]] = 2.0
            order = np.argsort(keys, axis=1)[:, :k]
            pos_valid = np.arange(k)[None, :] < r_cnt[:, None]
            order = np.where(pos_valid, order, max_len + 1)
            order.sort(axis=1)
            # which needle bytes: Full → all in order; Partial → a random sorted subset of size cnt
            nkeys = rng.random((rows.size, k), dtype=np.float32)
            sub = np.argsort(nkeys, axis=1)
            sub = np.where(pos_valid, sub, k + 1)
            sub.sort(axis=1)
            fullmask = is_full[rows]
            sub[fullmask] = np.arange(k)[None, :]
            rr, cc = np.nonzero(pos_valid)
            mat[rows[rr], order[rr, cc]] = nb[np.minimum(sub[rr, cc], k - 1)]
        if prefix_frac > 0:
            pr = np.nonzero(rng.random(m) < prefix_frac)[0]
            for k2, pre in enumerate(prefixes):
                rows2 = pr[k2::len(prefixes)]
                pb = np.frombuffer(pre, dtype=np.uint8)
                rows2 = rows2[length[rows2] >= len(pb)]
                mat[rows2[:, None], np.arange(len(pb))[None, :]] = pb[None, :]
        if unicode_frac > 0:
            # up to 3 splices per selected row, each inserting one scalar at a random byte position of the
            # current string (the base is ASCII and scalars are inserted whole, so boundaries stay valid)
            wide = np.zeros((m, max_len + 16), dtype=np.uint8)
            wide[:, :max_len] = mat
            cur = length.copy()
            sel = rng.random(m) < unicode_frac
            n_splice = np.where(sel, rng.integers(1, 4, m), 0)
            for it in range(3):
                rows3 = np.nonzero((n_splice > it) & (cur + 4 <= max_len))[0]
                if rows3.size == 0:
                    continue
                which = rng.integers(0, len(UNICODE_SCALARS), rows3.size)
                pos = (rng.random(rows3.size) * (cur[rows3] + 1)).astype(np.int64)
                # keep earlier splices intact: only splice at ASCII/lead-byte boundaries (not before a continuation byte)
                is_cont = (wide[rows3, np.minimum(pos, max_len + 15)] & 0xC0) == 0x80
                pos = np.where(is_cont, cur[rows3], pos)
                for w, sc in enumerate(UNICODE_SCALARS):
                    r4 = rows3[which == w]
                    if r4.size == 0:
                        continue
                    p4 = pos[which == w]
                    k4 = len(sc)
                    cols = np.arange(max_len + 16)[None, :]
                    src = np.where(cols < p4[:, None], cols, cols - k4)
                    moved = np.take_along_axis(wide[r4], np.clip(src, 0, max_len + 15), axis=1)
                    ins = (cols >= p4[:, None]) & (cols < p4[:, None] + k4)
                    scb = np.frombuffer(sc, dtype=np.uint8)
                    moved[ins] = scb[(cols - p4[:, None])[ins]]
                    wide[r4] = moved
                    cur[r4] += k4
            mat = wide[:, :max_len]
            length = np.minimum(cur, max_len)
            # never cut a scalar in half at max_len: back off to the previous char boundary
            for _ in range(3):
                over = (length < max_len + 1) & (length > 0)
                nxt = wide[np.arange(m), np.minimum(length, max_len + 15)]
                cut = over & ((nxt & 0xC0) == 0x80) & (cur > length)
                length = np.where(cut, length - 1, length)
        mask = np.arange(max_len)[None, :] < length[:, None]
        parts.append(mat[mask])
        lens_all.append(length)
    lens = np.concatenate(lens_all) if lens_all else np.zeros(0, dtype=np.int64)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:])
    data = np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint8)
    return np.ascontiguousarray(data), offsets


def to_list(data: np.ndarray, offsets: np.ndarray):
    b = data.tobytes()
    return [b[int(offsets[i]):int(offsets[i + 1])] for i in range(len(offsets) - 1)]
