#!/usr/bin/env bash
# round-2 pass y (gpurun --gpus 2): after the ring reservation fix -- noisy media across two ranks, host-model gain / noisy media
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( time timeout 400 python -m pytest tests/test_host_model.py tests/test_multirank.py -q -m gpu -k "gain or noisy" --durations=5 ) > $O/r2y_pytest_2gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2y_pytest_2gpu.log
tail -12 $O/r2y_pytest_2gpu.log
