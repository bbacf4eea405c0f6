#!/bin/bash
# round 2, 2-GPU call: multi-GPU tests (local form, torchrun worker, C client) and bench at N=2
export FRZ_BENCH_CACHE=/tmp/frz_cache
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rs > gpurun_out/r02c_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest_multi.log
tail -15 gpurun_out/r02c_pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 \
    > gpurun_out/r02c_bench_n2.json 2> gpurun_out/r02c_bench_n2.err; echo "bench n2 rc=$?"
tail -c 2500 gpurun_out/r02c_bench_n2.json; tail -5 gpurun_out/r02c_bench_n2.err
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err; echo "bench n1 rc=$?"
tail -c 1200 gpurun_out/r02c_bench_n1.json
