#!/bin/bash
# round 2, second GPU call: SW variant sweep + prefilter A/Bs, then ncu (launch list + full capture of the two top kernels)
export FRZ_BENCH_CACHE=/tmp/frz_cache
mkdir -p gpurun_out
bash profiles/run_r02_sw_variants.sh > /dev/null 2>&1
out=gpurun_out/r02b_sw_variants.txt
run() { label=$1; shift
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 2> gpurun_out/r02b_err_$label.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
s=d['roofline']['stage_ms_per_step']
print('$label', 'pf %.4f sw %.4f sort %.4f local %.4f step %.4f dev %.4f' % (s['prefilter'], s['smith_waterman'], s['sort'], s['local_pipeline'], d['ms_per_step'], d['value_device_out']['ms_per_step']), 'parity', d['parity']['mismatches'])
" >> $out 2>&1 || echo "$label FAILED" >> $out
}
run tma FRZ_PF_TMA=1
run tma_blocks5 FRZ_PF_TMA=1 FRZ_PF_BLOCKS=5
cat $out
# ncu: every launch of one step region, then the full set on the two top kernels
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/r02b_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps -1 > gpurun_out/r02b_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_prefilter|k_sw64' -s 6 -c 4 -o gpurun_out/r02b_prof -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps -1 > gpurun_out/r02b_ncu_full.log 2>&1
ls -la gpurun_out/r02b_prof.ncu-rep
