#!/usr/bin/env bash
# round-2 multi-GPU pass (gpurun --gpus N): multirank parity tests on the GPUs, then the driver's N-GPU bench command
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
N=${1:-2}
mkdir -p $O
nvidia-smi topo -m > $O/r2e_topo_$N.txt 2>&1
( time timeout 900 python -m pytest tests/test_multirank.py -m gpu -q --durations=8 ) > $O/r2e_pytest_mr_$N.log 2>&1
echo "pytest rc=$?" >> $O/r2e_pytest_mr_$N.log
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 \
    bench.py --gpus $N --steps 100 --warmup 5 ) > $O/r2e_bench_${N}gpu_f64.json 2> $O/r2e_bench_${N}gpu_f64.err
echo "bench rc=$?" >> $O/r2e_bench_${N}gpu_f64.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29572 \
    bench.py --gpus $N --steps 100 --warmup 5 --dtype f32 --no-extras ) > $O/r2e_bench_${N}gpu_f32.json 2> $O/r2e_bench_${N}gpu_f32.err
echo "bench rc=$?" >> $O/r2e_bench_${N}gpu_f32.err
tail -5 $O/r2e_pytest_mr_$N.log; tail -4 $O/r2e_bench_${N}gpu_f64.err; cut -c1-300 $O/r2e_bench_${N}gpu_f64.json
