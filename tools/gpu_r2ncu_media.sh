#!/usr/bin/env bash
# round-2: ncu launch lists of the multilevel-atom and noisy-Lorentzian paths (which kernels run per step, and for how long)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for c in mlatom3d:96 noisy3d:96; do
  t=${c%%:*}
  timeout 100 oracle/_ref/fp64/ref_driver --mode gpu --lib meep_nl_b200/libmeepcuda.so --case $c --steps 20 > $O/r2n_${t}_plain.log 2>&1
  echo "plain rc=$?" >> $O/r2n_${t}_plain.log
  timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 60 -c 120 --csv --log-file $O/r2n_launches_${t}.csv \
     oracle/_ref/fp64/ref_driver --mode gpu --lib meep_nl_b200/libmeepcuda.so --case $c --steps 12 > $O/r2n_${t}_ncu.log 2>&1
  echo "ncu rc=$?" >> $O/r2n_${t}_ncu.log
  tail -2 $O/r2n_${t}_plain.log | cut -c1-300
done
