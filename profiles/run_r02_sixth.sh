#!/bin/bash
# round 2, sixth GPU call (1 GPU): fused sort scan, NUMA host buffer, slice-exchange build — regression check + bench
export FRZ_BENCH_CACHE=/tmp/frz_cache
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i_pytest.log
tail -4 gpurun_out/r02i_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02i_bench.json').read().strip().splitlines()[-1])
s=d['roofline']['stage_ms_per_step']
print('value %.2f G/s ms/step %.4f dev %.4f e2e %.2f ms' % (d['value']/1e9, d['ms_per_step'], d['value_device_out']['ms_per_step'], d['e2e']['ms_per_step']), {k: round(v,4) for k,v in s.items()}, 'parity', d['parity']['mismatches'], 'cpu', d['cpu_baseline']['value']/1e9)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 24 --csv --log-file gpurun_out/r02i_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps -1 > gpurun_out/r02i_ncu_list.log 2>&1
