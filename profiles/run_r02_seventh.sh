#!/bin/bash
# round 2, seventh GPU call (1 GPU): slot-major corpus layout (a haystack's units contiguous) vs the unit-interleaved layout
# (libfrz_cuda_interleaved.so = the previous commit's sources), k_window resident blocks re-swept for the new layout,
# launch list + one full ncu capture of k_window and k_sw64 (DRAM traffic).
export FRZ_BENCH_CACHE=/tmp/frz_cache
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_pytest.log
tail -4 gpurun_out/r02j_pytest.log
out=gpurun_out/r02j_variants.txt; : > $out
run() { label=$1; shift; extra=$1; shift
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 $extra 2> gpurun_out/r02j_err_$label.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
s=d['roofline']['stage_ms_per_step']
print('$label', 'pf %.4f sw %.4f sort %.4f local %.4f step %.4f dev %.4f' % (s['prefilter'], s['smith_waterman'], s['sort'], s['local_pipeline'], d['ms_per_step'], d['value_device_out']['ms_per_step']), 'matches', d['config']['matches_per_step'], 'parity', d['parity']['mismatches'])
" >> $out 2>&1 || echo "$label FAILED" >> $out
}
run slotmajor "" A=1
run interleaved "" FRZ_LIB=$PWD/frizbee_b200/libfrz_cuda_interleaved.so
run slotmajor_b5 "" FRZ_PF_BLOCKS=5
run slotmajor_b3 "" FRZ_PF_BLOCKS=3
run slotmajor_k0 "--max-typos 0" A=1
run interleaved_k0 "--max-typos 0" FRZ_LIB=$PWD/frizbee_b200/libfrz_cuda_interleaved.so
cat $out
python bench.py --steps 20 --warmup 5 > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02j_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 24 --csv --log-file gpurun_out/r02j_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps -1 > gpurun_out/r02j_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_sig_scan|k_window|k_sw64' -s 9 -c 3 -o gpurun_out/r02j_prof -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps -1 > gpurun_out/r02j_ncu_full.log 2>&1
ls -la gpurun_out/r02j_prof.ncu-rep
