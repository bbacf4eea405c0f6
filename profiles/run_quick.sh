#!/bin/bash
# quick iteration run on the GPU box: parity tests, then a short bench line with stage timings
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(json.dumps({'value':d['value'],'ms_per_step':d['ms_per_step'],'stages':d['roofline']['stage_ms_per_step'],'frac':d['roofline']['frac'],'e2e':d['e2e'],'host_out':d.get('value_host_out'),'clocks':d['clocks'],'parity':d['parity']}))"
if [ "$1" == "sanitize" ]; then
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "edge_shapes or config1 or literal or multi_pattern or long_haystacks" 2>&1 | tail -12
fi
