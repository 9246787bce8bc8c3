#!/usr/bin/env bash
# round-2 pass l (gpurun --gpus 4): ranks with two remote neighbours -- multirank parity tests, then the driver's 4-GPU bench command
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
N=${1:-4}
mkdir -p $O
( time timeout 900 python -m pytest tests/test_multirank.py -m gpu -q -k "many_gpus" --durations=5 ) > $O/r2l_pytest_mr_$N.log 2>&1
echo "pytest rc=$?" >> $O/r2l_pytest_mr_$N.log
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 \
    bench.py --gpus $N --steps 100 --warmup 5 ) > $O/r2l_bench_${N}gpu_f64.json 2> $O/r2l_bench_${N}gpu_f64.err
echo "bench rc=$?" >> $O/r2l_bench_${N}gpu_f64.err
tail -5 $O/r2l_pytest_mr_$N.log; tail -4 $O/r2l_bench_${N}gpu_f64.err; cut -c1-300 $O/r2l_bench_${N}gpu_f64.json
