#!/bin/bash
# full ncu capture of the <=64-byte-window SW kernel (after warm-up), single GPU
TAG=${1:-x}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k k_sw64 --launch-skip 4 --launch-count 1 -o gpurun_out/prof_sw_${TAG} -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_sw_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_sw_${TAG}.log | cut -c1-300
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(json.dumps({'value':d['value'],'stages':d['roofline']['stage_ms_per_step'],'parity':d['parity']}))"
