#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in 0 1 3 4 5 6 7; do
  echo "== FRZ_SW_VARIANT=$v"
  FRZ_SW_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(json.dumps({'value':d['value'],'ms_per_step':d['ms_per_step'],'stages':d['roofline']['stage_ms_per_step'],'parity':d['parity']['mismatches']}))"
done
