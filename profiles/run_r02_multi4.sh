#!/bin/bash
# round 2, 4-GPU call (final tree): weak-scaling bench at N=4 with the P2P placement exchange (the N=1/2/8 partners are
# r02p_bench.json, r02k_bench_n2_p2p.json, r02l/r02n_bench_n8.json).
export FRZ_BENCH_CACHE=/tmp/frz_cache
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29714 bench.py --gpus 4 --steps 20 --warmup 5 \
    > gpurun_out/r02q_bench_n4.json 2> gpurun_out/r02q_bench_n4.err; echo "n4 rc=$?"
tail -c 300 gpurun_out/r02q_bench_n4.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02q_bench_n4.json").read().strip().splitlines()[-1])
s = d["roofline"]["stage_ms_per_step"]
print("n4 value %.2f G/s step %.4f ms dev-out %.4f ms e2e %.3f ms" % (d["value"] / 1e9, d["ms_per_step"], d["value_device_out"]["ms_per_step"], d["e2e"]["ms_per_step"]),
      {k: round(v, 4) for k, v in s.items()}, "parity", d["parity"]["mismatches"], "e2e equal", d["e2e"]["result_equals_resident_call"])
PY
