#!/bin/bash
# one full ncu capture of each dominant kernel (after warm-up), single GPU
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
ncu --set full --clock-control none --import-source on -k regex:'k_prefilter' --launch-skip 4 --launch-count 1 -o gpurun_out/prof_pf_${TAG} -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_pf_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k k_sw --launch-skip 8 --launch-count 1 -o gpurun_out/prof_sw_${TAG} -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_sw_${TAG}.log 2>&1
ls -la gpurun_out | tail -8
