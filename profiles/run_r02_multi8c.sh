#!/bin/bash
# round 2, third 8-GPU call: page placement of the shared host buffer (FRZ_HOST_NUMA=interleave|touch|none) — the slice copies
# of 8 GPUs into one host buffer took 0.21 ms at N=8 against 0.11 ms at N=1 (r02l).  Full record with the default policy,
# reduced runs for the other two; FRZ_PARALLEL_DEBUG prints where the pages actually are.
export FRZ_BENCH_CACHE=/tmp/frz_cache
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02n_topo.txt 2>&1; cat /sys/devices/system/node/has_memory; nproc
tr() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n + RANDOM % 100)) bench.py --gpus $n "$@"; }
FRZ_PARALLEL_DEBUG=1 tr 8 --steps 20 --warmup 5 > gpurun_out/r02n_bench_n8.json 2> gpurun_out/r02n_bench_n8.err; echo "n8 interleave rc=$?"
grep "frz numa" gpurun_out/r02n_bench_n8.err | head -20
for pol in touch none; do
  FRZ_HOST_NUMA=$pol FRZ_PARALLEL_DEBUG=1 tr 8 --steps 20 --warmup 5 --no-parity --e2e-steps -1 > gpurun_out/r02n_bench_n8_$pol.json 2> gpurun_out/r02n_bench_n8_$pol.err; echo "n8 $pol rc=$?"
  grep "frz numa" gpurun_out/r02n_bench_n8_$pol.err | head -10
done
python - <<'PY'
import json
for tag in ("bench_n8", "bench_n8_touch", "bench_n8_none"):
    try:
        d = json.loads(open(f"gpurun_out/r02n_{tag}.json").read().strip().splitlines()[-1])
        s = d["roofline"]["stage_ms_per_step"]
        print(tag, "value %.2f G/s step %.4f ms dev-out %.4f ms e2e %.3f ms" % (d["value"] / 1e9, d["ms_per_step"], d["value_device_out"]["ms_per_step"], d["e2e"]["ms_per_step"]),
              {k: round(v, 4) for k, v in s.items()}, "parity", (d.get("parity") or {}).get("mismatches"))
    except Exception as e:
        print(tag, "FAILED", e)
PY
