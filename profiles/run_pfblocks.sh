#!/bin/bash
# prefilter occupancy experiment: resident blocks per SM (register cap 127 / 96 / 80)
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for b in 4 5 6; do
  for k in 1 0; do
  FRZ_PF_BLOCKS=$b python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 3 --max-typos $k 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('blocks=$b k=$k', json.dumps({'value':d['value'],'stages':d['roofline']['stage_ms_per_step'],'e2e_ms':d['e2e']['ms_per_step'],'e2e':d['e2e']['value']}))"
  done
done
