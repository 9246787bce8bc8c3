#!/usr/bin/env bash
# round-2 fourth GPU pass (after the container was re-created): full parity suite, default bench with
# the configs block, reference arm, fp32, ncu launch lists + full captures of the top kernels
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total --format=csv > $O/r2d_gpu.txt 2>&1; nproc >> $O/r2d_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -q --durations=10 ) > $O/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2d_pytest.log
( time timeout 900 python bench.py ) > $O/r2d_bench_f64.json 2> $O/r2d_bench_f64.err
echo "bench rc=$?" >> $O/r2d_bench_f64.err
( time timeout 300 python bench.py --impl reference --steps 20 --warmup 3 ) > $O/r2d_bench_ref.json 2> $O/r2d_bench_ref.err
B="--no-extras --no-dropin --no-cpu-baseline --steps 20 --warmup 5"
timeout 300 python bench.py $B --dtype f32 > $O/r2d_bench_f32.json 2> $O/r2d_bench_f32.err
timeout 300 python bench.py $B --workload lorentz3d --size 512 > $O/r2d_lor512.json 2> $O/r2d_lor512.err
NB="python bench.py --no-cpu-baseline --no-extras --no-dropin --min-seconds 0.001"
for dt in f64 f32; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2d_launches_$dt.csv $NB --dtype $dt --steps 3 --warmup 3 > $O/r2d_ncu_l_$dt.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_mix|k_curl3d_v|k_halfg|k_conn_runs' --launch-skip 12 -c 4 -f -o $O/r2d_full_$dt $NB --dtype $dt --steps 2 --warmup 3 > $O/r2d_ncu_f_$dt.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_eh_general --launch-skip 402 -c 1 -f -o $O/r2d_nr_full \
   $NB --workload nr3d --size 192 --steps 2 --warmup 3 --develop 400 > $O/r2d_ncu_nr.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dft --launch-skip 8 -c 1 -f -o $O/r2d_dft_full \
   $NB --workload lorentz3d --size 256 --steps 2 --warmup 3 > $O/r2d_ncu_dft.log 2>&1
tail -4 $O/r2d_pytest.log; tail -3 $O/r2d_bench_f64.err; cut -c1-400 $O/r2d_bench_f64.json
