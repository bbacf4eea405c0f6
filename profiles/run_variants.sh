#!/bin/bash
# A/B of the Smith-Waterman shift-placement variants: same bench line, FRZ_SW_VARIANT in {0,1,3,5,7,8,11}
# (bits 0-2: one-lane shifts on the FMA pipe; bit 3: packed-byte bonus classification).
# Every variant is already checked against the oracle on the CPU (tests/test_kernel_logic_cpu.py); the bench line repeats
# the parity check on the GPU (parity.mismatches).
for v in 0 1 3 5 7 8 11; do
  for k in 1 0; do
    FRZ_SW_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --max-typos $k 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('variant=$v k=$k', json.dumps({'value':d['value'],'stages':d['roofline']['stage_ms_per_step'],'mismatches':d['parity']['mismatches']}))"
  done
done
