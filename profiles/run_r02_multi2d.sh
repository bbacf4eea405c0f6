#!/bin/bash
# round 2, 2-GPU call: the DIRECT placement form (k_place<DIRECT>: zero-copy stores into the shared mapped host buffer) —
# multi-GPU tests with all four exchange forms, then the weak-scaling bench at N=2 direct (default) vs P2P (A/B).
export FRZ_BENCH_CACHE=/tmp/frz_cache
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r02o_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02o_pytest_multi.log
tail -8 gpurun_out/r02o_pytest_multi.log
tr() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus $n "$@"; }
tr 2 --steps 20 --warmup 5 > gpurun_out/r02o_bench_n2_direct.json 2> gpurun_out/r02o_bench_n2_direct.err; echo "n2 direct rc=$?"
tail -c 300 gpurun_out/r02o_bench_n2_direct.err
FRZ_PARALLEL_EXCHANGE=p2p tr 2 --steps 20 --warmup 5 --e2e-steps -1 --no-parity > gpurun_out/r02o_bench_n2_p2p.json 2> gpurun_out/r02o_bench_n2_p2p.err; echo "n2 p2p rc=$?"
python - <<'PY'
import json
for tag in ("direct", "p2p"):
    try:
        d = json.loads(open(f"gpurun_out/r02o_bench_n2_{tag}.json").read().strip().splitlines()[-1])
        s = d["roofline"]["stage_ms_per_step"]
        print(tag, "value %.2f G/s step %.4f ms" % (d["value"] / 1e9, d["ms_per_step"]), {k: round(v, 4) for k, v in s.items()},
              "parity", (d.get("parity") or {}).get("mismatches"), "e2e ms", d["e2e"]["ms_per_step"], d["e2e"].get("result_equals_resident_call"), d["config"]["exchange"][:30])
    except Exception as e:
        print(tag, "FAILED", e)
PY
