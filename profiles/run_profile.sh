#!/bin/bash
# Run on the GPU box via gpurun: bench line, reference arm, ncu launch list (shares of the step), one full
# capture of the two dominant kernels.  Outputs land in gpurun_out/ (summaries copied to profiles/ by hand).
TAG=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 600 gpurun_out/bench_${TAG}.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
tail -c 400 gpurun_out/bench_ref_${TAG}.json
# launch list (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_list_${TAG}.log 2>&1
# full capture of the dominant kernels (2 launches each, after warm-up)
ncu --set full --clock-control none --import-source on -k regex:'k_prefilter|k_sw64' -s 6 -c 4 -o gpurun_out/prof_${TAG} -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/ | tail -8
