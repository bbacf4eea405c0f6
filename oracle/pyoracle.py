"""ctypes loader for the CPU oracle (oracle/frz_oracle.cpp).

TEST INFRASTRUCTURE ONLY — importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  The product package (frizbee_b200/) never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

from frizbee_b200.types import (CConfig, CMatch, CScoring, Config, Match, Pattern, Scoring,
                                as_pattern, pattern_array)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfrz_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "frz_oracle.cpp")
    hdr = os.path.join(_HERE, "..", "include", "frz_cuda.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(_LIB_PATH) for f in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "all"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u8p = C.c_char_p
        L.frzo_prefilter.restype = C.c_int
        L.frzo_prefilter.argtypes = [u8p, C.c_size_t, C.c_int, u8p, C.c_size_t, C.c_int, C.c_int,
                                     C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.frzo_sw_score.restype = C.c_uint16
        L.frzo_sw_score.argtypes = [u8p, C.c_size_t, C.POINTER(CScoring), C.c_int, u8p, C.c_size_t,
                                    C.c_int, C.c_int, C.c_int]
        L.frzo_prefilter_unicode.restype = C.c_int
        L.frzo_prefilter_unicode.argtypes = L.frzo_prefilter.argtypes
        L.frzo_sw_score_unicode.restype = C.c_uint16
        L.frzo_sw_score_unicode.argtypes = L.frzo_sw_score.argtypes
        L.frzo_sw_indices.restype = C.c_uint16
        L.frzo_sw_indices.argtypes = [u8p, C.c_size_t, C.POINTER(CScoring), C.c_int, C.c_int, u8p, C.c_size_t, C.c_uint64,
                                      C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        L.frzo_match_indices.restype = C.c_int
        L.frzo_match_indices.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                         C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.frzo_flip_scalar.restype = C.c_int
        L.frzo_flip_scalar.argtypes = [u8p, C.c_size_t, C.POINTER(C.c_uint8)]
        L.frzo_match_greedy.restype = C.c_int
        L.frzo_match_greedy.argtypes = [u8p, C.c_size_t, C.POINTER(CScoring), C.c_int, u8p, C.c_size_t, C.c_int]
        L.frzo_score_fits_in_u8.restype = C.c_int
        L.frzo_score_fits_in_u8.argtypes = [C.c_size_t, C.POINTER(CScoring)]
        L.frzo_radix_sort_matches.restype = None
        L.frzo_radix_sort_matches.argtypes = [C.c_void_p, C.c_uint64]
        for fn in (L.frzo_match_list_into, L.frzo_match_list):
            fn.restype = C.c_uint64
        L.frzo_match_list_into.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(CConfig), C.c_void_p, C.c_void_p,
                                           C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint64]
        L.frzo_match_list.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(CConfig), C.c_void_p, C.c_void_p,
                                      C.c_uint64, C.c_void_p, C.c_uint64]
        _lib = L
    return _lib


MATCH_DTYPE = np.dtype([("index", "<u4"), ("score", "<u2"), ("exact", "u1"), ("_pad", "u1")])


def _b(x) -> bytes:
    return x.encode("utf-8") if isinstance(x, str) else bytes(x)


def prefilter(needle, haystack, max_typos: Optional[int] = 0, lanes: int = 64,
              case_sensitive: bool = False) -> Tuple[bool, int, int]:
    """Prefilter::match_haystack* → (matched, start, end).  max_typos None = NO_PREFILTER."""
    n, h = _b(needle), _b(haystack)
    s, e = C.c_uint64(0), C.c_uint64(0)
    ok = lib().frzo_prefilter(n, len(n), int(case_sensitive), h, len(h),
                              -1 if max_typos is None else int(max_typos), lanes, C.byref(s), C.byref(e))
    return bool(ok), s.value, e.value


def sw_score(needle, haystack, scoring: Scoring = Scoring(), case_sensitive: bool = False,
             include_prefix: bool = True, lanes: int = 8, score_bits: int = 16) -> int:
    """SmithWaterman::<B>::score_haystack for backend B = (lanes, score_bits)."""
    n, h = _b(needle), _b(haystack)
    sc = CScoring.of(scoring)
    return int(lib().frzo_sw_score(n, len(n), C.byref(sc), int(case_sensitive), h, len(h),
                                   int(include_prefix), lanes, score_bits))


def prefilter_unicode(needle, haystack, max_typos: Optional[int] = 0, lanes: int = 16,
                      case_sensitive: bool = False) -> Tuple[bool, int, int]:
    """Prefilter::match_haystack_unicode* → (matched, start, end)."""
    n, h = _b(needle), _b(haystack)
    s, e = C.c_uint64(), C.c_uint64()
    ok = lib().frzo_prefilter_unicode(n, len(n), int(case_sensitive), h, len(h), -1 if max_typos is None else max_typos,
                                      lanes, C.byref(s), C.byref(e))
    return bool(ok), int(s.value), int(e.value)


def sw_score_unicode(needle, haystack, scoring: Scoring = Scoring(), case_sensitive: bool = False,
                     include_prefix: bool = True, lanes: int = 8, score_bits: int = 16) -> int:
    """SmithWaterman::<B>::score_haystack_unicode for backend B = (lanes, score_bits)."""
    n, h = _b(needle), _b(haystack)
    sc = CScoring.of(scoring)
    return int(lib().frzo_sw_score_unicode(n, len(n), C.byref(sc), int(case_sensitive), h, len(h),
                                           int(include_prefix), lanes, score_bits))


def sw_indices(needle, haystack, start_pos: int = 0, max_typos: Optional[int] = None, unicode: bool = False,
               scoring: Scoring = Scoring(), case_sensitive: bool = False, lanes: int = 8, score_bits: int = 16):
    """score_haystack_indices / score_haystack_unicode_indices → (score, indices in the reference's reverse order)."""
    n, h = _b(needle), _b(haystack)
    sc = CScoring.of(scoring)
    out = (C.c_uint32 * 4096)()
    cnt = C.c_uint32()
    score = lib().frzo_sw_indices(n, len(n), C.byref(sc), int(case_sensitive), int(unicode), h, len(h), start_pos,
                                  -1 if max_typos is None else max_typos, lanes, score_bits, out, 4096, C.byref(cnt))
    return int(score), list(out[: cnt.value])


def match_indices(pattern, config: Config, data: np.ndarray, offsets: np.ndarray, which, stride: int = 128):
    """Matcher::match_list_indices of a single-pattern matcher on the chosen haystacks.
    Returns a list of None (no match) or (score, exact, indices)."""
    plist = [as_pattern(p) for p in (pattern if isinstance(pattern, (list, tuple)) else [pattern])]
    arr = pattern_array(plist)
    cfg = CConfig.of(config)
    which = np.ascontiguousarray(which, dtype=np.uint32)
    out_idx = np.zeros((len(which), stride), dtype=np.uint32)
    out_cnt = np.zeros(len(which), dtype=np.uint32)
    out_m = np.zeros(len(which), dtype=MATCH_DTYPE)
    data = np.ascontiguousarray(data, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    rc = lib().frzo_match_indices(C.cast(arr, C.c_void_p), len(plist), C.byref(cfg), data.ctypes.data if data.size else None, offsets.ctypes.data, which.ctypes.data, len(which),
                                  out_idx.ctypes.data, stride, out_cnt.ctypes.data, out_m.ctypes.data)
    if rc != 0:
        raise ValueError("empty pattern")
    res = []
    for j in range(len(which)):
        if out_cnt[j] == 0xFFFFFFFF:
            res.append(None)
        else:
            res.append((int(out_m[j]["score"]), bool(out_m[j]["exact"]), out_idx[j, : min(int(out_cnt[j]), stride)].tolist()))
    return res


def flip_scalar(ch: str) -> str:
    """The opposite-case scalar case_needle_unicode pairs with `ch` (itself when there is none)."""
    b = ch.encode("utf-8")
    out = (C.c_uint8 * 4)()
    n = lib().frzo_flip_scalar(b, len(b), out)
    return bytes(out[:n]).decode("utf-8")


def match_greedy(needle, haystack, scoring: Scoring = Scoring(), case_sensitive: bool = False,
                 include_prefix: bool = True) -> Optional[int]:
    n, h = _b(needle), _b(haystack)
    sc = CScoring.of(scoring)
    r = lib().frzo_match_greedy(n, len(n), C.byref(sc), int(case_sensitive), h, len(h), int(include_prefix))
    return None if r < 0 else int(r)


def auto_lanes() -> int:
    """Lane width of the u8-family backend Matcher::get_backend (src/matcher/mod.rs:448-498) selects on THIS CPU:
    AVX-512 (F+BW+VBMI, BMI1/2 for the prefilter) -> 64, AVX2 -> 32, otherwise (SSE / NEON / scalar) 16."""
    flags = set()
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = set(line.split(":", 1)[1].split())
                    break
    except OSError:
        pass
    if {"avx512f", "avx512bw", "avx512vbmi", "bmi1", "bmi2"} <= flags:
        return 64
    if "avx2" in flags:
        return 32
    return 16


def score_fits_in_u8(needle_len: int, scoring: Scoring = Scoring()) -> bool:
    sc = CScoring.of(scoring)
    return bool(lib().frzo_score_fits_in_u8(needle_len, C.byref(sc)))


def pack(haystacks: Sequence) -> Tuple[np.ndarray, np.ndarray]:
    """List of str/bytes → Arrow-style (bytes u8[], offsets u64[n+1])."""
    raw = [_b(h) for h in haystacks]
    offsets = np.zeros(len(raw) + 1, dtype=np.uint64)
    if raw:
        offsets[1:] = np.cumsum([len(r) for r in raw], dtype=np.uint64)
    data = np.frombuffer(b"".join(raw), dtype=np.uint8).copy() if raw else np.zeros(0, dtype=np.uint8)
    return data, offsets


def radix_sort_matches(arr: np.ndarray) -> np.ndarray:
    arr = np.ascontiguousarray(arr, dtype=MATCH_DTYPE).copy()
    lib().frzo_radix_sort_matches(arr.ctypes.data, len(arr))
    return arr


def _run(fn, patterns, config: Config, data: np.ndarray, offsets: np.ndarray, extra):
    pats = [as_pattern(p) for p in patterns]
    arr = pattern_array(pats)
    cfg = CConfig.of(config)
    n = len(offsets) - 1
    data = np.ascontiguousarray(data, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    out = np.zeros(max(1, n), dtype=MATCH_DTYPE)
    dptr = data.ctypes.data if data.size else None
    cnt = fn(C.cast(arr, C.c_void_p), len(pats), C.byref(cfg), dptr, offsets.ctypes.data, n, *extra,
             out.ctypes.data, n)
    if cnt == 0xFFFFFFFFFFFFFFFF:
        raise NotImplementedError("oracle: unicode-needle path is not restated")
    return out[:cnt]


def match_list_into_packed(patterns, config: Config, data, offsets, index_offset: int = 0) -> np.ndarray:
    """Matcher::match_list_into: index order, unsorted (structured numpy array)."""
    return _run(lib().frzo_match_list_into, patterns, config, data, offsets, (C.c_uint32(index_offset),))


def match_list_packed(patterns, config: Config, data, offsets) -> np.ndarray:
    """Matcher::match_list: ordered per config.sort (structured numpy array)."""
    return _run(lib().frzo_match_list, patterns, config, data, offsets, ())


def match_list(patterns, haystacks: Sequence, config: Config = Config()) -> List[Match]:
    if isinstance(patterns, (str, bytes, Pattern)):
        patterns = [patterns]
    data, offsets = pack(haystacks)
    arr = match_list_packed(patterns, config, data, offsets)
    return [Match(score=int(m["score"]), index=int(m["index"]), exact=bool(m["exact"])) for m in arr]
