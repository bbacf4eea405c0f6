#!/bin/bash
# round 2, second 8-GPU call: weak-scaling bench at N=8 with the P2P placement exchange; BASELINE configs[3]
# (C4: 100 M haystacks, k=0, 8 shards of 12.5 M) and configs[4] (C5: 'foo !^bar' on 10 M mixed-unicode haystacks <= 128 B,
# 8 shards of 1.25 M) — every run with the full-shard parity leg (each rank checks its whole shard against the CPU restatement).
export FRZ_BENCH_CACHE=/tmp/frz_cache
mkdir -p gpurun_out
nvidia-smi -L | head -8; nproc
tr() { n=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n + RANDOM % 100)) bench.py --gpus $n "$@"; }
tr 8 --steps 20 --warmup 5 > gpurun_out/r02l_bench_n8.json 2> gpurun_out/r02l_bench_n8.err; echo "n8 rc=$?"
tail -c 400 gpurun_out/r02l_bench_n8.err
tr 8 --steps 20 --warmup 5 --haystacks-per-gpu 12500000 --max-typos 0 --e2e-steps 3 > gpurun_out/r02l_c4_n8.json 2> gpurun_out/r02l_c4_n8.err; echo "c4 rc=$?"
tail -c 400 gpurun_out/r02l_c4_n8.err
tr 8 --steps 20 --warmup 5 --haystacks-per-gpu 1250000 --query 'foo !^bar' --max-typos 0 --mu 96 --max-len 128 --unicode-frac 0.3 --prefix-frac 0.1 --e2e-steps 3 \
   > gpurun_out/r02l_c5_n8.json 2> gpurun_out/r02l_c5_n8.err; echo "c5 rc=$?"
tail -c 400 gpurun_out/r02l_c5_n8.err
python - <<'PY'
import json
for tag in ("bench_n8", "c4_n8", "c5_n8"):
    try:
        d = json.loads(open(f"gpurun_out/r02l_{tag}.json").read().strip().splitlines()[-1])
        s = d["roofline"]["stage_ms_per_step"]
        print(tag, "value %.2f G/s step %.4f ms dev-out %.4f ms e2e %.3f ms" % (d["value"] / 1e9, d["ms_per_step"], d["value_device_out"]["ms_per_step"], d["e2e"]["ms_per_step"]),
              {k: round(v, 4) for k, v in s.items()}, "matches", d["config"]["matches_per_step"], "parity", (d.get("parity") or {}).get("mismatches"))
    except Exception as e:
        print(tag, "FAILED", e)
PY
