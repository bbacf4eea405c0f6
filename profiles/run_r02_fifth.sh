#!/bin/bash
# round 2, fifth GPU call (1 GPU): TMA vs register-prefetch k_sig_scan, cp.async record ring in k_window, VAR 8 default
export FRZ_BENCH_CACHE=/tmp/frz_cache
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest.log
tail -4 gpurun_out/r02f_pytest.log
out=gpurun_out/r02f_variants.txt; : > $out
run() { label=$1; shift; extra=$1; shift
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 $extra 2> gpurun_out/r02f_err_$label.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
s=d['roofline']['stage_ms_per_step']
print('$label', 'pf %.4f sw %.4f sort %.4f local %.4f step %.4f dev %.4f' % (s['prefilter'], s['smith_waterman'], s['sort'], s['local_pipeline'], d['ms_per_step'], d['value_device_out']['ms_per_step']), 'matches', d['config']['matches_per_step'], 'parity', d['parity']['mismatches'])
" >> $out 2>&1 || echo "$label FAILED" >> $out
}
run base "" A=1
run notma "" FRZ_PF_TMA=0
run pfblocks4 "" FRZ_PF_BLOCKS=4
run pfblocks3 "" FRZ_PF_BLOCKS=3
run k0 "--max-typos 0" A=1
cat $out
ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 30 --csv --log-file gpurun_out/r02f_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps -1 > gpurun_out/r02f_ncu_list.log 2>&1
FRZ_PF_TMA=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 30 --csv --log-file gpurun_out/r02f_launches_notma.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps -1 > gpurun_out/r02f_ncu_list2.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_sig_scan|k_window' -s 6 -c 2 -o gpurun_out/r02f_prof -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps -1 > gpurun_out/r02f_ncu_full.log 2>&1
ls -la gpurun_out/r02f_prof.ncu-rep
