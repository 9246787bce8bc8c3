#!/usr/bin/env bash
# Run the reference-vs-CUDA twin comparison (oracle/_ref/*/ref_driver --mode shim) over a list
# of cases on the GPU box; one JSON line per case into gpurun_out/parity_<flavour>.jsonl
set -u
cd "$(dirname "$0")/.."
LIB=${LIB:-meep_nl_b200/libmeepcuda.so}
STEPS=${STEPS:-60}
mkdir -p gpurun_out
for fl in ${FLAVOURS:-fp64 fp32}; do
  out=gpurun_out/parity_$fl.jsonl
  : > $out
  for c in ${CASES:-vac3d:24 kr_metal3d kr_metal2d kr_pml2d kr_polariton1d kr_xper2d kr_xper3d kr_per2d kr_per2d_rods kr_per3d kr_per3d_rods kr_xper_ypml2d kr_xper_ypml3d wvg2d vac3d_cplx:16 phc3d:12 lorentz3d:20 kerr3d:14 kerr1d nr3d:12 nrslab3d:18 beta2d beta2d_cplx bfast2d bfast3d cond2d cond3d aniso3d magdrude2d magdrude3d gyro3d gyrodrude3d gyrosat3d dft2d dft3d}; do
    if [ $fl = fp32 ] && [[ $c == nr* ]]; then continue; fi
    OMP_NUM_THREADS=8 timeout 300 oracle/_ref/$fl/ref_driver --mode shim --lib $LIB --case $c --steps $STEPS --warmup 2 --json /tmp/p.json > /tmp/p.log 2>&1
    rc=$?
    if [ $rc -eq 0 ]; then cat /tmp/p.json >> $out; else echo "{\"case\":\"$c\",\"rc\":$rc,\"log\":\"$(tail -c 300 /tmp/p.log | tr '\n"' '  ')\"}" >> $out; fi
  done
  python - <<PY
import json
for l in open("$out"):
    d=json.loads(l)
    if 'rc' in d: print("FAIL", d); continue
    print("%-18s fp%d chunks=%-3d bad=%d/%d rel_l2=%.3g %s" % (d['case'], d['realnum_bytes']*8, d['num_chunks'], d['mismatched_elements'], d['elements'], d['rel_l2'], d['first_bad']))
PY
done
