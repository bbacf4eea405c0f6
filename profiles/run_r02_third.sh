#!/bin/bash
# round 2, third GPU call (1 GPU): split prefilter (k_sig_scan + k_window), SW VAR 8/16/24, single-chunk A/B, ncu
export FRZ_BENCH_CACHE=/tmp/frz_cache
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
tail -4 gpurun_out/r02d_pytest.log
out=gpurun_out/r02d_variants.txt; : > $out
run() { label=$1; shift; extra=$1; shift
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 $extra 2> gpurun_out/r02d_err_$label.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
s=d['roofline']['stage_ms_per_step']
print('$label', 'pf %.4f sw %.4f sort %.4f local %.4f step %.4f dev %.4f' % (s['prefilter'], s['smith_waterman'], s['sort'], s['local_pipeline'], d['ms_per_step'], d['value_device_out']['ms_per_step']), 'matches', d['config']['matches_per_step'], 'parity', d['parity']['mismatches'])
" >> $out 2>&1 || echo "$label FAILED" >> $out
}
run base "" FRZ_SW_VARIANT=0
run var8 "" FRZ_SW_VARIANT=8
run var16 "" FRZ_SW_VARIANT=16
run var24 "" FRZ_SW_VARIANT=24
run nosingle "" FRZ_PF_SINGLE=0
run pfblocks4 "" FRZ_PF_BLOCKS=4
run k0 "--max-typos 0" A=1
run k2 "--max-typos 2" A=1
run k3 "--max-typos 3" A=1
cat $out
ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 48 --csv --log-file gpurun_out/r02d_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps -1 > gpurun_out/r02d_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_sig_scan|k_window|k_sw64' -s 9 -c 6 -o gpurun_out/r02d_prof -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps -1 > gpurun_out/r02d_ncu_full.log 2>&1
ls -la gpurun_out/r02d_prof.ncu-rep
