#!/usr/bin/env bash
# round-2 pass r: fp32 parity subset after the march-length change, config 5 at full size
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( time timeout 600 python -m pytest tests/test_cuda_vs_reference.py tests/test_cuda_golden.py tests/test_host_model.py -m gpu -q -k "fp32 or f32 or phc or full_size" ) > $O/r2r_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2r_pytest.log
tail -3 $O/r2r_pytest.log
timeout 400 python bench.py --no-extras --no-dropin --no-cpu-baseline --steps 10 --warmup 3 --workload phc3d --size 1024 --dtype f32 > $O/r2r_phc1024.json 2> $O/r2r_phc1024.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2r_phc1024.json').read().strip().splitlines()[-1])
ks={k['name'][:18]:round(k['ms']/k['launches'],4) for k in d['roofline']['kernels']}
print(d['config']['workload'], 'Mcell/s', round(d['value']), 'ms', round(d['ms_per_step'],4), 'clk', d['clocks']['sm_mhz'], ks, d['roofline']['whole_step'])
PY
