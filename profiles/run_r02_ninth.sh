#!/bin/bash
# round 2, ninth GPU call (1 GPU): final tree — GPU tests, smoke(), the default bench line, and BASELINE config 4's list (100 M
# haystacks, k = 0) on ONE GPU as 8 logical shards: the strong-scaling partner of the 8-GPU run (profiles/r02l_c4_n8.json).
export FRZ_BENCH_CACHE=/tmp/frz_cache
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02p_pytest.log
tail -5 gpurun_out/r02p_pytest.log
python __graft_entry__.py smoke > gpurun_out/r02p_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r02p_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r02p_bench.err
timeout 600 python bench.py --gpus 1 --shards-per-gpu 8 --n 12500000 --max-typos 0 --steps 10 --warmup 3 --e2e-steps 3 --no-cpu-baseline \
    > gpurun_out/r02p_c4_n1_100M.json 2> gpurun_out/r02p_c4_n1_100M.err; echo "c4 1-gpu rc=$?"
tail -c 300 gpurun_out/r02p_c4_n1_100M.err
python - <<'PY'
import json
for tag in ("bench", "c4_n1_100M"):
    try:
        d = json.loads(open(f"gpurun_out/r02p_{tag}.json").read().strip().splitlines()[-1])
        s = d["roofline"]["stage_ms_per_step"]
        print(tag, "value %.2f G/s step %.4f ms dev-out %.4f ms e2e %.3f ms (%.3f G/s, equal=%s)" % (d["value"] / 1e9, d["ms_per_step"], d["value_device_out"]["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["value"] / 1e9, d["e2e"].get("result_equals_resident_call")),
              {k: round(v, 4) for k, v in s.items()}, "frac %.3f" % d["roofline"]["frac"], "matches", d["config"]["matches_per_step"], "parity", (d.get("parity") or {}).get("mismatches"))
    except Exception as e:
        print(tag, "FAILED", e)
PY
