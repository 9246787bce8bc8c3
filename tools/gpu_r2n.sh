#!/usr/bin/env bash
# round-2 GPU pass n (state after cylindrical, reductions, batched syncs, new k_dft): full parity suite, default bench with
# the configs block, reference arm, fp32, ncu launch lists + full captures of the top kernels
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total --format=csv > $O/r2n_gpu.txt 2>&1; nproc >> $O/r2n_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -q --durations=10 ) > $O/r2n_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2n_pytest.log
( time timeout 900 python bench.py ) > $O/r2n_bench_f64.json 2> $O/r2n_bench_f64.err
echo "bench rc=$?" >> $O/r2n_bench_f64.err
( time timeout 300 python bench.py --impl reference --steps 20 --warmup 3 ) > $O/r2n_bench_ref.json 2> $O/r2n_bench_ref.err
B="--no-extras --no-dropin --no-cpu-baseline --steps 20 --warmup 5"
timeout 300 python bench.py $B --dtype f32 > $O/r2n_bench_f32.json 2> $O/r2n_bench_f32.err
timeout 300 python bench.py $B --workload lorentz3d --size 512 > $O/r2n_lor512.json 2> $O/r2n_lor512.err
NB="python bench.py --no-cpu-baseline --no-extras --no-dropin --min-seconds 0.001"
for dt in f64 f32; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2n_launches_$dt.csv $NB --dtype $dt --steps 3 --warmup 3 > $O/r2n_ncu_l_$dt.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_mix|k_curl3d_v|k_halfg|k_conn_runs' --launch-skip 12 -c 4 -f -o $O/r2n_full_$dt $NB --dtype $dt --steps 2 --warmup 3 > $O/r2n_ncu_f_$dt.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_eh_general --launch-skip 402 -c 1 -f -o $O/r2n_nr_full \
   $NB --workload nr3d --size 192 --steps 2 --warmup 3 --develop 400 > $O/r2n_ncu_nr.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dft --launch-skip 8 -c 1 -f -o $O/r2n_dft_full \
   $NB --workload lorentz3d --size 256 --steps 2 --warmup 3 > $O/r2n_ncu_dft.log 2>&1
tail -4 $O/r2n_pytest.log; tail -3 $O/r2n_bench_f64.err; cut -c1-400 $O/r2n_bench_f64.json
# prefetch-distance variants (MCU_PREFETCH_PLANES = 1 / 3 / 4 against the default 2), built into build/variants/
for pf in 1 3 4; do
  for dt in f64 f32; do
    MEEPCU_LIBRARY=$PWD/build/variants/libmeepcuda_pf$pf.so timeout 300 python bench.py $B --dtype $dt > $O/r2n_pf${pf}_vac_$dt.json 2> $O/r2n_pf${pf}_vac_$dt.err
  done
  MEEPCU_LIBRARY=$PWD/build/variants/libmeepcuda_pf$pf.so timeout 300 python bench.py $B --workload phc3d --size 512 --dtype f32 > $O/r2n_pf${pf}_phc512.json 2> $O/r2n_pf${pf}_phc512.err
done
timeout 300 python bench.py $B --workload phc3d --size 512 --dtype f32 > $O/r2n_pf2_phc512.json 2> $O/r2n_pf2_phc512.err
for f in $O/r2n_pf*.json $O/r2n_bench_f32.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], d['config']['workload'], 'Mcell/s', round(d['value']), 'ms', round(d['ms_per_step'],4), 'min', round(d['timing']['ms_per_step_min'],4), 'clk', d['clocks']['sm_mhz'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
