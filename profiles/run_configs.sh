#!/bin/bash
# numbers for the other BASELINE.json configs (parity-test cases; recorded for DESIGN.md, not the bench line)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
short() { python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(json.dumps({'workload':d['config']['workload'][:90],'value':d['value'],'ms_per_step':d['ms_per_step'],'matches':d['config']['matches_per_step'],'stages':d['roofline']['stage_ms_per_step'],'frac':d['roofline']['frac'],'e2e_ms':d['e2e']['ms_per_step'],'parity':d['parity']}))"; }
echo "== config 2 shape: needle len 6, 1M, len<=32, k=0"
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --needle deadbe --max-typos 0 --mu 24 --max-len 32 --n 1000000 2>&1 | tail -1 | short
echo "== config 4 shard shape: needle len 8, 10M (one shard), len<=64, k=0"
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --max-typos 0 2>&1 | tail -1 | short
echo "== config 3 (bench line workload)"
python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | short
echo "== config 5 shape: 'foo !^bar', 4M mixed-unicode haystacks, len<=128"
python - <<'PY'
import time, numpy as np, torch
import frizbee_b200 as F
from frizbee_b200 import synth
from frizbee_b200.types import Config, Pattern, Matching
from oracle import pyoracle as O
n=4_000_000
data,off=synth.generate("foo", n, 96, 128, unicode_frac=0.3, prefix_frac=0.1)
corpus=F.Corpus.from_arrow(data,off)
m=F.Matcher.from_query("foo !^bar", Config())
out=np.empty(n,dtype=F.MATCH_DTYPE)
for _ in range(3): r=m.match_list_array(corpus,out=out)
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(10): r=m.match_list_array(corpus,out=out)
torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/10
sub=200_000
want=O.match_list_packed([Pattern("foo"),Pattern("bar",negated=True,matching=Matching.Prefix)], Config(emulate_lanes=m.backend_info()["prefilter_lanes"]), data[:int(off[sub])], off[:sub+1])
got=r[r["index"]<sub]
ok=np.array_equal(np.sort(got,order="index"),np.sort(want,order="index"))
print({"n":n,"matches":len(r),"ms_per_call_host_out":dt*1e3,"haystacks_per_s":n/dt,"prefix_parity_ok":bool(ok),"total_bytes":int(off[-1])})
PY
