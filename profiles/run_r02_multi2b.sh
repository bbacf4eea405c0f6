#!/bin/bash
# round 2, second 2-GPU call: slice exchange + NUMA-placed shared buffer — multi-GPU tests (both exchange forms), bench N=2 both forms
export FRZ_BENCH_CACHE=/tmp/frz_cache
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rs > gpurun_out/r02h_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_pytest_multi.log
tail -12 gpurun_out/r02h_pytest_multi.log
tr2() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 "$@"; }
tr2 --steps 20 --warmup 5 > gpurun_out/r02h_bench_n2.json 2> gpurun_out/r02h_bench_n2.err; echo "bench n2 slices rc=$?"
tail -c 1800 gpurun_out/r02h_bench_n2.json; tail -5 gpurun_out/r02h_bench_n2.err
FRZ_PARALLEL_EXCHANGE=allgather tr2 --steps 20 --warmup 5 --e2e-steps -1 > gpurun_out/r02h_bench_n2_allgather.json 2> gpurun_out/r02h_bench_n2_allgather.err; echo "bench n2 allgather rc=$?"
tail -c 900 gpurun_out/r02h_bench_n2_allgather.json
