#!/usr/bin/env bash
# round-2 first GPU pass: parity suite + default bench (1 GPU)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_gpu.txt 2>&1
nproc >> gpurun_out/r2a_gpu.txt; free -g >> gpurun_out/r2a_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?" >> gpurun_out/r2a_bench.err
( time timeout 300 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
tail -5 gpurun_out/r2a_pytest.log; tail -3 gpurun_out/r2a_bench.err; cut -c1-600 gpurun_out/r2a_bench.json
