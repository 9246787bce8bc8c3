#!/usr/bin/env bash
# round-2 pass v (gpurun --gpus 8): the driver's 8-GPU command without the configs block (main line + parity gate)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
N=${1:-8}
mkdir -p $O
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 \
    bench.py --gpus $N --steps 100 --warmup 5 --no-extras ) > $O/r2v_bench_${N}gpu_f64.json 2> $O/r2v_bench_${N}gpu_f64.err
echo "bench rc=$?" >> $O/r2v_bench_${N}gpu_f64.err
tail -4 $O/r2v_bench_${N}gpu_f64.err; cut -c1-300 $O/r2v_bench_${N}gpu_f64.json
