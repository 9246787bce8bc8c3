#!/usr/bin/env bash
# round-2 pass q: march length of the interior tiles at config-5-like sizes (fp32), k_mix on / off
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
B="--no-extras --no-dropin --no-cpu-baseline --steps 20 --warmup 5 --min-seconds 1.0"
for seg in 8 16 32 64; do
  MEEPCU_SEG=$seg timeout 300 python bench.py $B --workload phc3d --size 512 --dtype f32 > $O/r2q_seg${seg}_phc512.json 2> $O/r2q_seg${seg}_phc512.err
  MEEPCU_SEG=$seg timeout 300 python bench.py $B --dtype f32 > $O/r2q_seg${seg}_vac_f32.json 2> $O/r2q_seg${seg}_vac_f32.err
done
MEEPCU_MIX=0 timeout 300 python bench.py $B --workload phc3d --size 512 --dtype f32 > $O/r2q_mix0_phc512.json 2> $O/r2q_mix0_phc512.err
MEEPCU_MIX=1 timeout 300 python bench.py $B --workload phc3d --size 512 --dtype f32 > $O/r2q_mix1_phc512.json 2> $O/r2q_mix1_phc512.err
MEEPCU_MIX=0 timeout 300 python bench.py $B --dtype f32 > $O/r2q_mix0_vac_f32.json 2> $O/r2q_mix0_vac_f32.err
for f in $O/r2q_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    ks={k['name'][:16]:round(k['ms']/k['launches'],4) for k in d['roofline']['kernels']}
    print(sys.argv[1].split('/')[-1], d['config']['workload'], 'Mcell/s', round(d['value']), 'ms', round(d['ms_per_step'],4), 'min', round(d['timing']['ms_per_step_min'],4), 'clk', d['clocks']['sm_mhz'], ks)
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
