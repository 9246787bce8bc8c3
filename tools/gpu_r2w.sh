#!/usr/bin/env bash
# round-2 pass w: the new media on the B200 -- multilevel atoms and noisy Lorentzian (twins against the reference, golden replays),
# one compute-sanitizer run of each on a small cell
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( time timeout 500 python -m pytest tests/test_cuda_vs_reference.py tests/test_cuda_golden.py -q -m gpu -k "mlatom or noisy" --durations=5 ) > $O/r2w_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2w_pytest.log
tail -25 $O/r2w_pytest.log
for c in mlatom3d:10 noisy2d:16; do
  ( timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 oracle/_ref/fp64/ref_driver --mode shim --lib meep_nl_b200/libmeepcuda.so \
      --case $c --steps 4 --warmup 1 ) > $O/r2w_sanitizer_${c%%:*}.log 2>&1
  echo "sanitizer $c rc=$?" | tee -a $O/r2w_sanitizer_${c%%:*}.log
  tail -4 $O/r2w_sanitizer_${c%%:*}.log | cut -c1-300
done
