#!/bin/bash
# round 2, first GPU call: gpu tests, bench at N=1 (both arms)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r02a_bench.json; tail -5 gpurun_out/r02a_bench.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02a_bench_ref.json 2> gpurun_out/r02a_bench_ref.err; echo "ref rc=$?"
tail -c 1500 gpurun_out/r02a_bench_ref.json
nvidia-smi -L | head -3; nproc
