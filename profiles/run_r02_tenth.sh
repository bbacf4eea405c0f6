#!/bin/bash
# round 2, tenth GPU call (1 GPU): ncu evidence for the FINAL binary — launch list of one step and one full capture of
# k_sig_scan / k_window (5 blocks per SM) / k_sw64.
export FRZ_BENCH_CACHE=/tmp/frz_cache
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 24 --csv --log-file gpurun_out/r02r_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps -1 > gpurun_out/r02r_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_sig_scan|k_window|k_sw64' -s 9 -c 3 -o gpurun_out/r02r_prof -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps -1 > gpurun_out/r02r_ncu_full.log 2>&1
ls -la gpurun_out/r02r_prof.ncu-rep gpurun_out/r02r_launches.csv
