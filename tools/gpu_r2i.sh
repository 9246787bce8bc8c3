#!/usr/bin/env bash
# round-2 pass i: batched small-array syncs (e2e), full parity suite
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=6 ) > $O/r2i_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2i_pytest.log
B="--no-extras --no-dropin --no-cpu-baseline --warmup 5"
timeout 300 python bench.py $B --steps 100 > $O/r2i_vac_f64.json 2> $O/r2i_vac_f64.err
timeout 300 python bench.py $B --steps 20 > $O/r2i_vac_f64_k20.json 2> $O/r2i_vac_f64_k20.err
MEEPCU_NO_BATCH=1 timeout 300 python bench.py $B --steps 20 > $O/r2i_vac_f64_k20_nobatch.json 2> $O/r2i_vac_f64_k20_nobatch.err
timeout 300 python bench.py $B --steps 20 --dtype f32 > $O/r2i_vac_f32_k20.json 2> $O/r2i_vac_f32_k20.err
tail -3 $O/r2i_pytest.log
for f in $O/r2i_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], d['config']['workload'], 'K', d['steps'], 'Mcell/s', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'e2e_s', round(d['config']['cells']*d['steps']/d['e2e']['value']/1e6*1e3,1), 'ms')
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
