#!/bin/bash
# prefilter register-slice experiment: SLICE=4 (93 regs, 5 blocks/SM) vs SLICE=8 (125 regs, 4 blocks/SM)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for sl in 4 8; do
  for k in 1 0; do
  FRZ_PF_SLICE=$sl python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 3 --max-typos $k 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('slice=$sl k=$k', json.dumps({'value':d['value'],'stages':d['roofline']['stage_ms_per_step'],'parity':d['parity']['mismatches']}))"
  done
done
