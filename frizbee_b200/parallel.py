"""Multi-GPU `match_list_parallel` (src/matcher/parallel.rs:18-89): one process per GPU, the haystack
list sharded by contiguous index range (shard g holds indices [offset_g, offset_g + n_g)), each rank
scores its shard into a locally ordered run that stays in HBM, one NCCL all-gather moves the runs,
and a device merge reproduces the reference's k-way merge (src/k_merge.rs:90-131) bit for bit.

torch / torch.distributed are plumbing here (device buffers, NCCL); the compute is the C ABI.
The shard/merge host logic is backend-agnostic and is covered on CPU with gloo (tests/test_parallel_gloo.py)
by substituting the run producer.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Tuple

import numpy as np

from . import MATCH_DTYPE, Corpus, FrizbeeError, Matcher, _check, lib
from .types import SortStrategy


def shard_bounds(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous index ranges, shard g = [g*ceil(n/world), ...)."""
    per = (n + world - 1) // world if world else 0
    return [(min(g * per, n), min((g + 1) * per, n)) for g in range(world)]


def merge_runs_host(runs: List[np.ndarray], sort: SortStrategy) -> np.ndarray:
    """Reference semantics of k_merge_matches_by_* on host arrays (used by the gloo tests and as the
    specification of the device merge): runs are index-range shards in rank order, each ordered per `sort`."""
    if not runs:
        return np.zeros(0, dtype=MATCH_DTYPE)
    cat = np.concatenate(runs[::-1] if sort.is_reversed() else runs)
    if not sort.is_by_score():
        return cat
    order = np.argsort(-cat["score"].astype(np.int64), kind="stable")
    return cat[order]


def all_gather_runs(run, count: int, group=None):
    """The collective step, backend-agnostic (NCCL on device tensors, gloo on CPU tensors): one all-gather of the
    counts and ONE all-gather of the runs padded to the longest.  `run` is a 1-D int64 tensor of 8-byte match records.
    Returns (gathered [world * stride], counts list, stride)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    cnt = torch.tensor([count], dtype=torch.int64, device=run.device)
    counts = torch.zeros(world, dtype=torch.int64, device=run.device)
    dist.all_gather_into_tensor(counts, cnt, group=group)
    counts_h = [int(c) for c in counts.cpu().tolist()]
    stride = max(max(counts_h), 1)
    send = run[:stride].contiguous() if run.numel() >= stride else torch.nn.functional.pad(run, (0, stride - run.numel()))
    gathered = torch.empty(world * stride, dtype=torch.int64, device=run.device)
    dist.all_gather_into_tensor(gathered, send, group=group)
    return gathered, counts_h, stride


def match_list_parallel_host(run: np.ndarray, sort: SortStrategy, group=None) -> np.ndarray:
    """Host/gloo form of the gather + merge (tests; the run producer is injected by the caller)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(run).view(np.int64).copy())
    gathered, counts, stride = all_gather_runs(t, len(run), group)
    g = gathered.numpy().view(MATCH_DTYPE)
    runs = [g[r * stride: r * stride + counts[r]] for r in range(len(counts))]
    return merge_runs_host(runs, sort)


class ShardRunner:
    """Persistent device buffers for repeated match_list_parallel calls on one shard (no allocation per call)."""

    def __init__(self, matcher: Matcher, shard: Corpus, index_offset: int, group=None, device: Optional[int] = None):
        import torch
        import torch.distributed as dist
        self.matcher, self.shard, self.index_offset, self.group = matcher, shard, index_offset, group
        self.dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        n_local = max(len(shard), 1)
        self.run = torch.empty(n_local, dtype=torch.int64, device=self.dev)   # 8-byte frz_match records
        self.count = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self.counts = torch.zeros(self.world, dtype=torch.int64, device=self.dev)
        self.gathered = None
        self.merged = None
        self.bound = matcher.score_bound()
        self.side = torch.cuda.Stream(self.dev) if self.world > 1 else None   # count exchange overlaps scoring

    def local(self):
        """This rank's shard → locally ordered run in HBM (asynchronous).  Returns (run tensor, count tensor)."""
        import torch
        stream = torch.cuda.current_stream(self.dev)
        _check(lib().frz_match_shard_device(self.matcher._h, self.shard._h, self.index_offset, self.run.data_ptr(),
                                            self.run.numel(), self.count.data_ptr(), stream.cuda_stream))
        return self.run, self.count

    def step(self):
        """Matcher::match_list_parallel: local run, ONE all-gather of the padded runs, device merge.
        Returns (merged tensor view, total)."""
        import torch
        import torch.distributed as dist
        run, count = self.local()
        if self.world == 1:
            return run, None
        # the count is final once the prefilter has run: exchange it on a side stream while the shard is scored
        _check(lib().frz_matcher_wait_count(self.matcher._h, self.side.cuda_stream))
        with torch.cuda.stream(self.side):
            dist.all_gather_into_tensor(self.counts, count, group=self.group)
            counts_h = np.asarray(self.counts.cpu().tolist(), dtype=np.uint64)  # the one host sync of the step
        torch.cuda.current_stream(self.dev).wait_stream(self.side)
        stride = max(int(counts_h.max()), 1)
        total = int(counts_h.sum())
        need = self.world * stride
        if self.gathered is None or self.gathered.numel() < need:
            self.gathered = torch.empty(int(need * 1.25) + 1024, dtype=torch.int64, device=self.dev)
        if self.merged is None or self.merged.numel() < total:
            self.merged = torch.empty(int(total * 1.25) + 1024, dtype=torch.int64, device=self.dev)
        g = self.gathered[:need]
        dist.all_gather_into_tensor(g, run[:stride], group=self.group)          # run is shard-sized >= stride
        stream = torch.cuda.current_stream(self.dev)
        _check(lib().frz_merge_runs_device(g.data_ptr(), stride, counts_h.ctypes.data, self.world, int(self.matcher.config.sort),
                                           self.bound, self.merged.data_ptr(), self.dev.index, stream.cuda_stream))
        return self.merged[:total], total


def match_list_parallel(matcher: Matcher, shard: Corpus, index_offset: int, group=None, device: Optional[int] = None):
    """Runs this rank's shard, all-gathers the runs over NCCL and merges them on every rank.
    Returns (merged matches as an int64 torch tensor of 8-byte records on the device, total count)."""
    r = ShardRunner(matcher, shard, index_offset, group, device)
    merged, total = r.step()
    if total is None:
        total = int(r.count.item())
        merged = merged[:total]
    return merged, total


def matches_from_tensor(t) -> np.ndarray:
    return t.cpu().numpy().view(MATCH_DTYPE)
