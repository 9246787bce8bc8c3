#!/usr/bin/env bash
# round-2 second GPU pass: parity suite with virtual E, A/B benches, ncu of k_eh_general / k_dft
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 ) > $O/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2b_pytest.log
B="--no-extras --no-dropin --no-cpu-baseline --steps 20 --warmup 5"
for dt in f64 f32; do
  timeout 300 python bench.py $B --dtype $dt > $O/r2b_vac_${dt}_virt.json 2> $O/r2b_vac_${dt}_virt.err
  MEEPCU_NO_VIRTUAL_E=1 timeout 300 python bench.py $B --dtype $dt > $O/r2b_vac_${dt}_stored.json 2> $O/r2b_vac_${dt}_stored.err
done
timeout 300 python bench.py $B --workload phc3d --size 512 --dtype f32 > $O/r2b_phc512_virt.json 2> $O/r2b_phc512_virt.err
MEEPCU_NO_VIRTUAL_E=1 timeout 300 python bench.py $B --workload phc3d --size 512 --dtype f32 > $O/r2b_phc512_stored.json 2> $O/r2b_phc512_stored.err
timeout 300 python bench.py $B --workload diel3d --size 256 --dtype f64 > $O/r2b_diel_virt.json 2> $O/r2b_diel_virt.err
MEEPCU_NO_VIRTUAL_E=1 timeout 300 python bench.py $B --workload diel3d --size 256 --dtype f64 > $O/r2b_diel_stored.json 2> $O/r2b_diel_stored.err
timeout 300 python bench.py $B --workload nr3d --size 256 > $O/r2b_nr256.json 2> $O/r2b_nr256.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_eh_general --launch-skip 420 -c 2 -f -o $O/r2b_nr_full \
   python bench.py --workload nr3d --size 192 --steps 2 --warmup 3 --develop 400 --min-seconds 0.001 --no-extras --no-dropin --no-cpu-baseline > $O/r2b_ncu_nr.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dft --launch-skip 8 -c 2 -f -o $O/r2b_dft_full \
   python bench.py --workload lorentz3d --size 256 --steps 2 --warmup 3 --min-seconds 0.001 --no-extras --no-dropin --no-cpu-baseline > $O/r2b_ncu_dft.log 2>&1
tail -4 $O/r2b_pytest.log
for f in $O/r2b_*_virt.json $O/r2b_*_stored.json $O/r2b_nr256.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d['config']['workload'], 'Mcell/s', round(d['value']), 'ms', round(d['ms_per_step'],4), 'min', round(d['timing']['ms_per_step_min'],4), 'virt', d['details']['virtual_e_arrays'], 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'B/cell', d['roofline']['whole_step']['bytes_per_cell'], 'frac', round(d['roofline']['whole_step']['frac'],3))
except Exception as e: print('ERR', e)
PY
done
