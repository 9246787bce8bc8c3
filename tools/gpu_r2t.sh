#!/usr/bin/env bash
# round-2 pass t: fp64 vector kernels at 3 CTAs/SM (80 registers, ~200 B of spills) against 2 (128 registers)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
B="--no-extras --no-dropin --no-cpu-baseline --steps 20 --warmup 5 --min-seconds 1.0"
for v in default mix3 c3dv3 both3; do
  L=$PWD/build/variants/libmeepcuda_$v.so
  [ $v = default ] && L=$PWD/meep_nl_b200/libmeepcuda.so
  MEEPCU_LIBRARY=$L timeout 300 python bench.py $B --dtype f64 > $O/r2t_${v}_vac_f64.json 2> $O/r2t_${v}_vac_f64.err
  MEEPCU_LIBRARY=$L timeout 300 python bench.py $B --workload phc3d --size 384 --dtype f64 > $O/r2t_${v}_phc384_f64.json 2> $O/r2t_${v}_phc384_f64.err
  MEEPCU_LIBRARY=$L timeout 300 python bench.py $B --workload diel3d --size 256 --dtype f64 > $O/r2t_${v}_diel256_f64.json 2> $O/r2t_${v}_diel256_f64.err
done
for f in $O/r2t_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    ks={k['name'][:16]:round(k['ms']/k['launches'],4) for k in d['roofline']['kernels']}
    print(sys.argv[1].split('/')[-1], d['config']['workload'], 'Mcell/s', round(d['value']), 'ms', round(d['ms_per_step'],4), 'min', round(d['timing']['ms_per_step_min'],4), 'clk', d['clocks']['sm_mhz'], ks)
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
