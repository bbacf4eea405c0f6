#!/bin/bash
# round 2: Smith-Waterman variant sweep (SwCore VAR bits, two column classes) with the parity leg on.
# VAR bits: 1 = diagonal shift as IMAD/IMAD.HI, 2 = gap-step-1 score shift as IMAD, 4 = gap-step-1 mask shift as IMAD,
#           8 = per-column bonus classified on packed bytes
mkdir -p gpurun_out
out=gpurun_out/r02b_sw_variants.txt; : > $out
run() {  # label, env...
  label=$1; shift
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 2> gpurun_out/r02b_err_$label.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
s=d['roofline']['stage_ms_per_step']
print('$label', 'pf %.4f sw %.4f sort %.4f local %.4f step %.4f dev %.4f' % (s['prefilter'], s['smith_waterman'], s['sort'], s['local_pipeline'], d['ms_per_step'], d['value_device_out']['ms_per_step']), 'parity', d['parity']['mismatches'])
" >> $out 2>&1 || echo "$label FAILED" >> $out
}
for v in 0 1 3 5 7 8 11; do run var$v FRZ_SW_VARIANT=$v; done
run two_classes FRZ_LIB=$PWD/frizbee_b200/libfrz_cuda_two.so
run two_classes_var3 FRZ_LIB=$PWD/frizbee_b200/libfrz_cuda_two.so FRZ_SW_VARIANT=3
for b in 3 4 5; do run pf_blocks$b FRZ_PF_BLOCKS=$b; done
cat $out
# 2-typo / 3-typo prefilter: scanning forms (default) vs the occurrence-mask forms (FRZ_PF_MASKS_K2 build)
run2() { label=$1; shift; k=$1; shift
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps -1 --max-typos $k 2> gpurun_out/r02b_err_$label.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
s=d['roofline']['stage_ms_per_step']
print('$label', 'pf %.4f sw %.4f sort %.4f step %.4f' % (s['prefilter'], s['smith_waterman'], s['sort'], d['ms_per_step']), 'matches', d['config']['matches_per_step'], 'parity', d['parity']['mismatches'])
" >> $out 2>&1 || echo "$label FAILED" >> $out
}
run2 k2_scan 2
run2 k2_masks 2 FRZ_LIB=$PWD/frizbee_b200/libfrz_cuda_k2.so
run2 k3_scan 3
run2 k3_masks 3 FRZ_LIB=$PWD/frizbee_b200/libfrz_cuda_k2.so
run2 k0 0
cat $out
