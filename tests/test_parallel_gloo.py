"""World-size-2 gloo test of the multi-GPU host logic (shard bounds, the single all-gather of padded runs, the
merge) on CPU.  The per-shard run producer is the oracle here — on the GPU box it is frz_match_shard_device.
Mirrors the reference's parallel == sequential tests (src/matcher/parallel.rs:104-173, tests/api_properties.rs:626-668)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from frizbee_b200 import parallel
from frizbee_b200.types import Config, Pattern, Matching, SortStrategy
from oracle import pyoracle as O


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _haystacks():
    hs = ["nomatch"] * 4101
    for i in (0, 2047, 2048, 2049, 4095, 4096, 4100):
        hs[i] = "foo"
    for i in (5, 1000, 3000):
        hs[i] = "f_o_o"
    hs[2050] = "xfoo"
    return hs


def _worker(rank, world, port, sort_value, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        hs = _haystacks()
        lo, hi = parallel.shard_bounds(len(hs), world)[rank]
        data, off = O.pack(hs[lo:hi])
        cfg = Config(sort=SortStrategy(sort_value))
        # this rank's locally ordered run (indices offset by the shard start)
        run = O.match_list_into_packed(["foo"], cfg, data, off, index_offset=lo)
        if cfg.sort.is_reversed():
            run = run[::-1]
        if cfg.sort.is_by_score():
            run = O.radix_sort_matches(run)
        merged = parallel.match_list_parallel_host(np.ascontiguousarray(run), cfg.sort)
        if rank == 0:
            q.put(merged.tobytes())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("sort", list(SortStrategy))
def test_parallel_equals_sequential_gloo(sort):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, int(sort), q)) for r in range(2)]
    for p in procs:
        p.start()
    got = np.frombuffer(q.get(timeout=120), dtype=O.MATCH_DTYPE)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    data, off = O.pack(_haystacks())
    want = O.match_list_packed(["foo"], Config(sort=sort), data, off)
    assert np.array_equal(got, want)


def test_shard_bounds_and_host_merge():
    assert parallel.shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert parallel.shard_bounds(0, 2) == [(0, 0), (0, 0)]
    assert parallel.shard_bounds(5, 8)[-1] == (5, 5)
    rng = np.random.default_rng(1)
    full = np.zeros(1000, dtype=O.MATCH_DTYPE)
    full["index"] = np.arange(1000)
    full["score"] = rng.integers(0, 40, 1000)
    for sort in SortStrategy:
        runs = []
        for lo, hi in parallel.shard_bounds(1000, 3):
            r = full[lo:hi]
            if sort.is_reversed():
                r = r[::-1]
            if sort.is_by_score():
                r = O.radix_sort_matches(r)
            runs.append(r)
        want = full[::-1] if sort.is_reversed() else full
        if sort.is_by_score():
            want = O.radix_sort_matches(want)
        assert np.array_equal(parallel.merge_runs_host(runs, sort), want)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_p2p_placement_arithmetic_equals_the_k_way_merge(world):
    """The position arithmetic of the P2P placement (csrc/parallel.cu: k_place, restated in parallel.placement_host): every rank
    stores its own matches straight at their merged positions inside the owning rank's slice; the concatenated slices are
    the k-way merge for all four sort strategies, skewed runs, empty runs and clamped score bins."""
    rng = np.random.default_rng(world)
    for n, hi_score in ((1000, 40), (37, 3), (world - 1, 5), (0, 5), (5000, 600)):
        full = np.zeros(n, dtype=O.MATCH_DTYPE)
        full["index"] = np.arange(n)
        full["score"] = rng.integers(0, hi_score, n)
        full["exact"] = rng.integers(0, 2, n)
        for sort in SortStrategy:
            runs = []
            for lo, hi in parallel.shard_bounds(n, world):
                r = full[lo:hi]
                if lo < hi and rng.random() < 0.3:       # a skewed shard: drop most of its matches
                    r = r[rng.random(len(r)) < 0.1]
                if sort.is_reversed():
                    r = r[::-1]
                if sort.is_by_score():
                    r = O.radix_sort_matches(r)
                runs.append(np.ascontiguousarray(r))
            want = parallel.merge_runs_host(runs, sort)
            for bins in (1024, 512):                     # 512 < 600: the top bin is shared by several scores
                if bins <= hi_score and sort.is_by_score():
                    continue                             # (the device path only uses a table that separates all scores)
                got = np.concatenate(parallel.placement_host(runs, sort, bins=bins))
                assert np.array_equal(got, want), (world, n, sort, bins)


def _placement_worker(rank, world, port, sort_value, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        hs = _haystacks()
        lo, hi = parallel.shard_bounds(len(hs), world)[rank]
        data, off = O.pack(hs[lo:hi])
        cfg = Config(sort=SortStrategy(sort_value))
        run = O.match_list_into_packed(["foo"], cfg, data, off, index_offset=lo)
        if cfg.sort.is_reversed():
            run = run[::-1]
        if cfg.sort.is_by_score():
            run = O.radix_sort_matches(run)
        piece = parallel.match_list_parallel_placement_gloo(np.ascontiguousarray(run), cfg.sort)
        q.put((rank, piece.tobytes()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("sort", list(SortStrategy))
def test_parallel_placement_protocol_gloo(sort):
    """World-size-2 run of the P2P placement protocol (tables published → every rank positions its own run → elements go to
    the rank that owns their slice): the concatenated slices equal the sequential match_list, for every sort strategy, on
    the reference's own parallel test list (src/matcher/parallel.rs:104-130)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_placement_worker, args=(r, 2, port, int(sort), q)) for r in range(2)]
    for p in procs:
        p.start()
    pieces = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got = np.concatenate([np.frombuffer(pieces[r], dtype=O.MATCH_DTYPE) for r in range(2)])
    data, off = O.pack(_haystacks())
    want = O.match_list_packed(["foo"], Config(sort=sort), data, off)
    assert np.array_equal(got, want)
