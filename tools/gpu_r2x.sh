#!/usr/bin/env bash
# round-2 pass x (gpurun --gpus 2): multilevel-atom / noisy media across two ranks (NCCL), and the drop-in seam with them
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( time timeout 500 python -m pytest tests/test_multirank.py tests/test_dropin.py -q -m gpu -k "mlatom or noisy" --durations=5 ) > $O/r2x_pytest_2gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2x_pytest_2gpu.log
tail -25 $O/r2x_pytest_2gpu.log
