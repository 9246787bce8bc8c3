#!/usr/bin/env bash
# round-2 pass j: cylindrical coordinates on the GPU (twins against the reference), then the whole parity suite
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for c in cyl:30:0 cyl:30:1 cylper:30:2 cylpml:40:0 cylpml:40:1 cylpml:40:3 cylcond:40:1 cyllor:40:1; do
  MEEPCU_DEBUG_COMPARE=6 timeout 120 oracle/_ref/fp64/ref_driver --mode shim --lib meep_nl_b200/libmeepcuda.so --case $c --steps 100 --json $O/r2j_$c.json > $O/r2j_$c.out 2>&1
  echo "$c rc=$? $(python -c "
import json
try:
  d=json.load(open('$O/r2j_$c.json')); print({k:d.get(k) for k in ('arrays','mismatched_arrays','mismatched_elements','max_abs_diff','rel_l2','first_bad')})
except Exception as e: print('ERR', open('$O/r2j_$c.out').read()[-600:])
")"
  grep "ref " $O/r2j_$c.out | head -6
done 2>&1 | tee $O/r2j_cyl_summary.txt
( time timeout 1500 python -m pytest tests -m gpu -q --durations=6 ) > $O/r2j_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2j_pytest.log
tail -15 $O/r2j_pytest.log
