#!/usr/bin/env bash
# round-2 pass o: dipole patching instead of k_fmp (NR / Pade parity), prefetch distance 0 / 1 / 2 per workload
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -q -k "nr or kerr or pade or golden or dropin" ) > $O/r2o_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2o_pytest.log
tail -3 $O/r2o_pytest.log
B="--no-extras --no-dropin --no-cpu-baseline --steps 20 --warmup 5 --min-seconds 1.5"
timeout 300 python bench.py $B --workload nr3d --size 384 > $O/r2o_nr384.json 2> $O/r2o_nr384.err
timeout 300 python bench.py $B --workload kerr3d --size 384 > $O/r2o_kerr384.json 2> $O/r2o_kerr384.err
for pf in 0 1 2; do
  L=$PWD/build/variants/libmeepcuda_pf$pf.so
  MEEPCU_LIBRARY=$L timeout 300 python bench.py $B --dtype f64 > $O/r2o_pf${pf}_vac_f64.json 2> $O/r2o_pf${pf}_vac_f64.err
  MEEPCU_LIBRARY=$L timeout 300 python bench.py $B --dtype f32 > $O/r2o_pf${pf}_vac_f32.json 2> $O/r2o_pf${pf}_vac_f32.err
  MEEPCU_LIBRARY=$L timeout 300 python bench.py $B --workload phc3d --size 512 --dtype f32 > $O/r2o_pf${pf}_phc512_f32.json 2> $O/r2o_pf${pf}_phc512_f32.err
  MEEPCU_LIBRARY=$L timeout 300 python bench.py $B --workload phc3d --size 384 --dtype f64 > $O/r2o_pf${pf}_phc384_f64.json 2> $O/r2o_pf${pf}_phc384_f64.err
  MEEPCU_LIBRARY=$L timeout 300 python bench.py $B --workload lorentz3d --size 256 --dtype f64 > $O/r2o_pf${pf}_lor256_f64.json 2> $O/r2o_pf${pf}_lor256_f64.err
  MEEPCU_LIBRARY=$L timeout 300 python bench.py $B --workload diel3d --size 256 --dtype f64 > $O/r2o_pf${pf}_diel256_f64.json 2> $O/r2o_pf${pf}_diel256_f64.err
done
for f in $O/r2o_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    ks={k['name'][:14]:round(k['ms']/k['launches'],4) for k in d['roofline']['kernels']}
    print(sys.argv[1].split('/')[-1], d['config']['workload'], 'Mcell/s', round(d['value']), 'ms', round(d['ms_per_step'],4), 'min', round(d['timing']['ms_per_step_min'],4), 'clk', d['clocks']['sm_mhz'], ks)
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
