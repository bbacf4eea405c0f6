#!/bin/bash
# round 2, 8-GPU call: weak-scaling bench at N=8 and N=4, BASELINE configs[3] (C4: 100 M, k=0, 8 shards) and configs[4]
# (C5: 'foo !^bar' on 10 M mixed-unicode haystacks, 8 shards), all with the full-shard parity leg.
export FRZ_BENCH_CACHE=/tmp/frz_cache
mkdir -p gpurun_out
nvidia-smi -L | head -8; nproc
tr() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n "$@"; }
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,GRAPH tr 8 --steps 20 --warmup 5 > gpurun_out/r02g_bench_n8.json 2> gpurun_out/r02g_bench_n8.err; echo "n8 rc=$?"
grep -E "NVLS|via P2P|NET/|Channel 00|nChannels|Connected all" gpurun_out/r02g_bench_n8.err | head -12 > gpurun_out/r02g_nccl_transport.txt
tail -c 1500 gpurun_out/r02g_bench_n8.json; grep -v "NCCL INFO" gpurun_out/r02g_bench_n8.err | tail -5
tr 8 --steps 20 --warmup 5 --haystacks-per-gpu 12500000 --max-typos 0 --e2e-steps -1 > gpurun_out/r02g_c4_n8.json 2> gpurun_out/r02g_c4_n8.err; echo "c4 rc=$?"
tail -c 1200 gpurun_out/r02g_c4_n8.json; tail -3 gpurun_out/r02g_c4_n8.err
tr 8 --steps 20 --warmup 5 --haystacks-per-gpu 1250000 --query 'foo !^bar' --max-typos 0 --mu 96 --max-len 128 --unicode-frac 0.3 --prefix-frac 0.1 \
   > gpurun_out/r02g_c5_n8.json 2> gpurun_out/r02g_c5_n8.err; echo "c5 rc=$?"
tail -c 1200 gpurun_out/r02g_c5_n8.json; tail -3 gpurun_out/r02g_c5_n8.err
tr 4 --steps 20 --warmup 5 > gpurun_out/r02g_bench_n4.json 2> gpurun_out/r02g_bench_n4.err; echo "n4 rc=$?"
tail -c 600 gpurun_out/r02g_bench_n4.json; tail -3 gpurun_out/r02g_bench_n4.err
