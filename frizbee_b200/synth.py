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


def generate(needle: str, n: int, mu: int, max_len: int, seed: int = 12345, p_partial: float = 0.20,
             p_full: float = 0.05, chunk: int = 1 << 20, alphabet: bytes = None) -> Tuple[np.ndarray, np.ndarray]:
    """Returns Arrow-style (bytes u8[total], offsets u64[n+1])."""
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
