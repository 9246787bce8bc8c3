#!/usr/bin/env bash
# round-2 pass g: new k_dft (per-point phase + streaming phase), k_eh_general with L2 prefetch, SURVEY-count roofline
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=6 ) > $O/r2g_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2g_pytest.log
timeout 120 python tools/pcie_probe.py > $O/r2g_pcie.txt 2>&1
B="--no-extras --no-dropin --no-cpu-baseline --steps 20 --warmup 5"
timeout 300 python bench.py $B --steps 100 > $O/r2g_vac_f64.json 2> $O/r2g_vac_f64.err
timeout 300 python bench.py $B --workload lorentz3d --size 512 > $O/r2g_lor512.json 2> $O/r2g_lor512.err
timeout 300 python bench.py $B --workload lorentz3d --size 256 --dtype f32 > $O/r2g_lor256_f32.json 2> $O/r2g_lor256_f32.err
for c in 3 4 5; do
  MEEPCU_EH_CTAS=$c timeout 300 python bench.py $B --workload nr3d --size 384 > $O/r2g_nr384_c$c.json 2> $O/r2g_nr384_c$c.err
done
timeout 300 python bench.py $B --workload kerr3d --size 384 > $O/r2g_kerr384.json 2> $O/r2g_kerr384.err
NB="python bench.py --no-cpu-baseline --no-extras --no-dropin --min-seconds 0.001"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dft --launch-skip 8 -c 1 -f -o $O/r2g_dft_full \
   $NB --workload lorentz3d --size 256 --steps 2 --warmup 3 > $O/r2g_ncu_dft.log 2>&1
tail -4 $O/r2g_pytest.log; cat $O/r2g_pcie.txt
for f in $O/r2g_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    ks={k['name']:round(k['ms']/k['launches'],4) for k in d['roofline']['kernels']}
    print(sys.argv[1].split('/')[-1], d['config']['workload'], 'Mcell/s', round(d['value']), 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), 'ws', round(d['roofline']['whole_step']['frac'],3), ks)
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
