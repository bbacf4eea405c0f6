#!/bin/bash
# round 2, eighth GPU call (1 GPU): the streamed end-to-end call (match pipeline overlapped with the H2D chunks) — GPU tests
# (incl. the 2 M-haystack streamed-vs-resident test and the unicode signature-scan path) and the e2e A/B.
export FRZ_BENCH_CACHE=/tmp/frz_cache
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02m_pytest.log
tail -6 gpurun_out/r02m_pytest.log
python bench.py --steps 20 --warmup 5 --e2e-steps 10 > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r02m_bench.err
FRZ_E2E_STREAM=0 python bench.py --steps 5 --warmup 3 --e2e-steps 10 --no-cpu-baseline --no-parity > gpurun_out/r02m_bench_nostream.json 2> gpurun_out/r02m_bench_nostream.err; echo "nostream rc=$?"
python - <<'PY'
import json
for tag in ("bench", "bench_nostream"):
    try:
        d = json.loads(open(f"gpurun_out/r02m_{tag}.json").read().strip().splitlines()[-1])
        print(tag, "value %.2f G/s step %.4f ms | e2e %.4f ms (%.3f G/s) streamed=%s equal=%s" % (d["value"] / 1e9, d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["value"] / 1e9, d["e2e"]["streamed"], d["e2e"]["result_equals_resident_call"]))
    except Exception as e:
        print(tag, "FAILED", e)
PY
