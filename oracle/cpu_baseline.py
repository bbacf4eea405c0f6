"""CPU baseline for bench.py — the reference's match_list_parallel restated on the oracle.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (bench.py's cpu_baseline and `--impl reference` legs).

Threading mirrors src/matcher/parallel.rs:18-89: 2048-item chunks claimed from a shared counter,
each worker appends index-ordered matches, sorts its run (stable, score desc), and the runs are
k-way merged (here: concatenate + stable sort, which yields the identical sequence).
The per-haystack work runs in native code (ctypes releases the GIL), so Python threads scale.

kind: "port" — the real reference is a Rust crate and cannot be built in this image (no rustc).
If oracle/libfrz_cpu_baseline.so (the SIMD restatement at the reference's lane widths) has been
built it is used; otherwise the scalar oracle is.
"""
from __future__ import annotations

import ctypes as C
import itertools
import os
import threading
import time
from typing import Optional, Tuple

import numpy as np

from frizbee_b200.types import CConfig, Config, as_pattern, pattern_array
from oracle import pyoracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_SIMD_PATH = os.path.join(_HERE, "libfrz_cpu_baseline.so")
_simd = None


def simd_lib():
    global _simd
    if _simd is None and os.path.exists(_SIMD_PATH):
        L = C.CDLL(_SIMD_PATH)
        L.frzb_match_list_parallel.restype = C.c_uint64
        L.frzb_match_list_parallel.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(CConfig), C.c_void_p, C.c_void_p,
                                               C.c_uint64, C.c_int, C.c_void_p, C.c_uint64]
        L.frzb_isa.restype = C.c_char_p
        _simd = L
    return _simd


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def describe() -> str:
    L = simd_lib()
    if L is not None:
        return "C++ restatement with " + L.frzb_isa().decode() + " intrinsics at the reference's lane widths"
    return "scalar C++ restatement (oracle/frz_oracle.cpp)"


def match_list_parallel(patterns, config: Config, data: np.ndarray, offsets: np.ndarray, threads: Optional[int] = None
                        ) -> np.ndarray:
    threads = threads or host_threads()
    n = len(offsets) - 1
    L = simd_lib()
    pats = [as_pattern(p) for p in patterns]
    if L is not None:
        arr = pattern_array(pats)
        cfg = CConfig.of(config)
        out = np.empty(max(1, n), dtype=O.MATCH_DTYPE)
        cnt = L.frzb_match_list_parallel(C.cast(arr, C.c_void_p), len(pats), C.byref(cfg), data.ctypes.data,
                                          offsets.ctypes.data, n, threads, out.ctypes.data, n)
        if cnt != 0xFFFFFFFFFFFFFFFF:
            return out[:cnt]
    # scalar oracle, threaded like match_list_parallel
    chunk = 2048
    n_chunks = (n + chunk - 1) // chunk
    counter = itertools.count()
    lock = threading.Lock()
    runs = [None] * threads
    lib = O.lib()
    arr = pattern_array(pats)
    cfg = CConfig.of(config.with_(sort=config.sort))

    def worker(t):
        local = []
        buf = np.empty(chunk, dtype=O.MATCH_DTYPE)
        while True:
            with lock:
                ci = next(counter)
            if ci >= n_chunks:
                break
            lo = ci * chunk
            m = min(chunk, n - lo)
            cnt = lib.frzo_match_list_into(C.cast(arr, C.c_void_p), len(pats), C.byref(cfg), data.ctypes.data,
                                           offsets[lo:].ctypes.data, m, C.c_uint32(lo), buf.ctypes.data, chunk)
            if cnt:
                local.append(buf[:cnt].copy())
        run = np.concatenate(local) if local else np.zeros(0, dtype=O.MATCH_DTYPE)
        run = run[np.argsort(run["index"], kind="stable")]
        if config.sort.is_reversed():
            run = run[::-1]
        if config.sort.is_by_score():
            run = run[np.argsort(-run["score"].astype(np.int64), kind="stable")]
        runs[t] = run

    ths = [threading.Thread(target=worker, args=(t,)) for t in range(threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    cat = np.concatenate(runs)
    # k-way merge of the sorted runs == one ordering of the union (ties broken by index; src/k_merge.rs:14-51)
    if config.sort.is_by_score():
        idx_key = cat["index"].astype(np.int64)
        if config.sort.is_reversed():
            idx_key = -idx_key
        order = np.lexsort((idx_key, -cat["score"].astype(np.int64)))
    else:
        order = np.argsort(cat["index"], kind="stable")
        if config.sort.is_reversed():
            order = order[::-1]
    return cat[order]


def timed(patterns, config: Config, data, offsets, threads: Optional[int] = None, repeats: int = 1
          ) -> Tuple[float, np.ndarray]:
    best, res = float("inf"), None
    for _ in range(repeats):
        t0 = time.perf_counter()
        res = match_list_parallel(patterns, config, data, offsets, threads)
        best = min(best, time.perf_counter() - t0)
    return best, res
