#!/usr/bin/env bash
# round-2 third GPU pass: k_eh_general mask + occupancy variants, new k_dft, ncu of k_eh_general
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests/test_cuda_golden.py tests/test_dropin.py tests/test_host_model.py tests/test_cuda_vs_reference.py -m gpu -q -k "nr or kerr or dft or lorentz or virtual or golden or dropin or flux" ) > $O/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2c_pytest.log
B="--no-extras --no-dropin --no-cpu-baseline --steps 20 --warmup 5 --min-seconds 0.5"
MEEPCU_NO_EH_MASK=1 timeout 300 python bench.py $B --workload nr3d --size 256 > $O/r2c_nr_nomask.json 2> $O/r2c_nr_nomask.err
for c in 2 3 4 5; do
  MEEPCU_EH_CTAS=$c timeout 300 python bench.py $B --workload nr3d --size 256 > $O/r2c_nr_c$c.json 2> $O/r2c_nr_c$c.err
done
MEEPCU_NO_EH_MASK=1 timeout 300 python bench.py $B --workload kerr3d --size 256 > $O/r2c_kerr_nomask.json 2> $O/r2c_kerr_nomask.err
for c in 3 5; do
  MEEPCU_EH_CTAS=$c timeout 300 python bench.py $B --workload kerr3d --size 256 > $O/r2c_kerr_c$c.json 2> $O/r2c_kerr_c$c.err
done
timeout 300 python bench.py $B --workload lorentz3d --size 256 > $O/r2c_lor256.json 2> $O/r2c_lor256.err
MEEPCU_EH_CTAS=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_eh_general --launch-skip 402 -c 1 -f -o $O/r2c_nr_full \
   python bench.py --workload nr3d --size 192 --steps 2 --warmup 3 --develop 400 --min-seconds 0.001 --no-extras --no-dropin --no-cpu-baseline > $O/r2c_ncu_nr.log 2>&1
tail -4 $O/r2c_pytest.log
for f in $O/r2c_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    ks={k['name']:round(k['ms']/k['launches'],4) for k in d['roofline']['kernels']}
    print(sys.argv[1].split('/')[-1], d['config']['workload'], 'Mcell/s', round(d['value']), 'ms', round(d['ms_per_step'],4), 'clk', d['clocks']['sm_mhz'], ks, d['details'].get('newton_raphson'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
