#!/usr/bin/env bash
# Regenerate the golden ABI traces from the UNMODIFIED reference (needs /root/reference and
# `make -C oracle all`).  Each trace = problem description + initial arrays + per-step source
# samples + the arrays the reference's own CPU stepping produced (see tests/trace.py).
# The recording run itself asserts that oracle/liboracle_meep.so reproduces the reference
# bit for bit (mismatched_elements must be 0) before the file is kept.
set -euo pipefail
cd "$(dirname "$0")/../.."
OUT=tests/golden
gen() { # flavour case steps warmup nl
  local fl=$1 c=$2 steps=$3 warm=$4 nl=${5:-0}
  local tag=$(echo "$c" | tr ':.' '__')
  local f=$OUT/${tag}_${fl}.trace
  OMP_NUM_THREADS=1 oracle/_ref/$fl/ref_driver --mode shim --lib oracle/liboracle_meep.so --case "$c" \
      --steps "$steps" --warmup "$warm" --nl "$nl" --record "$f" --json /tmp/golden.json > /dev/null 2>&1
  python3 - "$f" <<'PY'
import json, sys
d = json.load(open('/tmp/golden.json'))
assert d['mismatched_elements'] == 0 and d['mismatched_arrays'] == 0, d
print("%-28s fp%d chunks=%d arrays=%d elements=%d probe=%.12g" % (sys.argv[1], 8*d['realnum_bytes'], d['num_chunks'], d['arrays'], d['elements'], d['probe_ref']))
PY
  gzip -9 -f "$f"
}
for fl in fp64 fp32; do
  gen $fl kr_polariton1d 60 0
  gen $fl kr_metal2d 40 0
  gen $fl kr_pml2d 40 2
  gen $fl kr_per2d_rods 40 0
  gen $fl kr_xper_ypml2d 40 2
  gen $fl kr_metal3d 30 0
  gen $fl kr_per3d_rods 30 0
  gen $fl kr_xper_ypml3d 30 2
  gen $fl wvg2d 60 3
  gen $fl vac3d_cplx:14 30 2
  gen $fl phc3d:10 30 0
  gen $fl lorentz3d:16 40 2
  gen $fl kerr3d:12 30 2
  gen $fl kerr1d 60 0
  gen $fl beta2d 40 2
  gen $fl bfast3d:10 40 2
  gen $fl bfast2d 40 2
  gen $fl cond3d:10 40 2
  gen $fl cond2d 40 2
  gen $fl aniso3d:12 40 2
  gen $fl dft2d:24 40 2
  gen $fl gyro3d:10 40 2
  gen $fl gyrosat3d:10 40 2
  gen $fl magdrude3d:12 40 2
  gen $fl magdrude2d:24 40 2
  gen $fl dft3d:12 40 2
  gen $fl beta2d_cplx 40 2
  gen $fl mlatom3d:10 40 2
  gen $fl mlatom2d:24 40 2
  gen $fl noisy3d:10 24 2
  gen $fl noisy2d:16 24 2
  gen $fl cylpml:14:1 40 0
  gen $fl cyllor:14:0 40 0
done
gen fp64 cylper:12:2 40 0
gen fp64 cylpmlnz:14:3 40 0
gen fp64 nr3d:10 30 0
gen fp64 nrslab3d:14 30 2
if [ -x oracle/_ref/fp64_pade/ref_driver ]; then
  gen fp64_pade kerr3d:12 30 2 1
  gen fp64_pade kerr1d 60 0 1
  gen fp64_pade kr_per3d_rods 30 0 1
  gen fp64_pade kr_per2d_rods 40 0 1
fi
ls -la $OUT | awk '{s+=$5} END {print "total bytes", s}'
