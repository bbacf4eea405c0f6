"""torchrun worker: shards a list over WORLD_SIZE GPUs, runs match_list_parallel (NCCL all-gather + device
merge) and checks on rank 0 that the result equals the single-GPU match_list and the oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import frizbee_b200 as F
from frizbee_b200 import parallel, synth
from frizbee_b200.types import Config, SortStrategy


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 400_003
    data, off = synth.generate("deadbeef", n, 48, 64, seed=7)
    ok = True
    for sort in SortStrategy:
        for k in (0, 1):
            cfg = Config(max_typos=k, sort=sort)
            lo, hi = parallel.shard_bounds(n, world)[rank]
            sdata = data[int(off[lo]):int(off[hi])]
            soff = (off[lo:hi + 1] - off[lo]).astype(np.uint64)
            shard = F.Corpus.from_arrow(sdata, soff, device=local)
            m = F.Matcher("deadbeef", cfg)
            merged, total = parallel.match_list_parallel(m, shard, lo, device=local)
            got = parallel.matches_from_tensor(merged)
            if rank == 0:
                full = F.Corpus.from_arrow(data, off, device=local)
                want = F.Matcher("deadbeef", cfg).match_list_array(full, device=local)
                same = len(got) == len(want) and np.array_equal(got, want)
                print(f"sort={sort.name} k={k}: {total} matches, parallel == single-GPU: {same}", flush=True)
                ok = ok and same
                full.close()
            shard.close()
            m.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
