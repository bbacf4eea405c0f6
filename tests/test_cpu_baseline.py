"""The SIMD CPU baseline (oracle/cpu_baseline/, measurement infrastructure) must agree bit for bit with
the scalar oracle at the same emulated lane width — otherwise it is not a credible stand-in for the
reference's SIMD backends.  CPU-only."""
import random

import numpy as np
import pytest

from frizbee_b200 import synth
from frizbee_b200.types import CaseMatching, Config, SortStrategy
from oracle import cpu_baseline as cb
from oracle import pyoracle as O


def _isa_lanes():
    L = cb.simd_lib()
    if L is None:
        pytest.skip("SIMD baseline not built")
    isa = L.frzb_isa().decode()
    return [64, 32] if "512" in isa else [32] if "AVX2" in isa else pytest.skip("no AVX2")


@pytest.mark.parametrize("k", [0, 1, None])
def test_simd_baseline_matches_scalar_oracle_bench_shape(k):
    data, off = synth.generate("deadbeef", 60_000, 48, 64)
    for lanes in _isa_lanes():
        for sort in (SortStrategy.ScoreThenIndexAsc, SortStrategy.IndexAsc, SortStrategy.ScoreThenIndexDesc):
            cfg = Config(max_typos=k, emulate_lanes=lanes, sort=sort)
            want = O.match_list_packed(["deadbeef"], cfg, data, off)
            got = cb.match_list_parallel(["deadbeef"], cfg, data, off, threads=4)
            assert np.array_equal(got, want), (lanes, k, sort)


def test_simd_baseline_dense_alphabet_and_lengths():
    rng = random.Random(9)
    hs = [bytes(rng.choice(b"abAB_/-ab01") for _ in range(rng.randint(0, 200))) for _ in range(20_000)]
    data, off = O.pack(hs)
    for lanes in _isa_lanes():
        for needle in ("ab", "aB_", "b/a-", "abABab01"):
            for k in (0, 1, None):
                for casing in (CaseMatching.Smart, CaseMatching.Respect):
                    cfg = Config(max_typos=k, emulate_lanes=lanes, casing=casing)
                    want = O.match_list_packed([needle], cfg, data, off)
                    got = cb.match_list_parallel([needle], cfg, data, off, threads=3)
                    assert np.array_equal(got, want), (lanes, needle, k, casing)


def test_threaded_scalar_fallback_equals_sequential():
    # parallel == sequential across the 2048-chunk boundary (src/matcher/parallel.rs:104-130)
    hs = ["nomatch"] * 4101
    for i in (0, 2047, 2048, 2049, 4095, 4096, 4100):
        hs[i] = "foo"
    data, off = O.pack(hs)
    for threads in (1, 2, 3, 8):
        for sort in SortStrategy:
            cfg = Config(sort=sort, max_typos=2)  # max_typos=2 is outside the SIMD scope → scalar threaded path
            want = O.match_list_packed(["foo"], cfg, data, off)
            got = cb.match_list_parallel(["foo"], cfg, data, off, threads=threads)
            assert np.array_equal(got, want), (threads, sort)
