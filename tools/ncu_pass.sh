# ncu evidence for the default bench (BASELINE config 2, 256^3): launch list + full capture of the top kernels
set -x
B="python bench.py --no-cpu-baseline"
TAG=${1:-r01c}
for dt in f64 f32; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_$dt.csv $B --dtype $dt --steps 3 --warmup 3 > gpurun_out/${TAG}_ncu_l_$dt.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:'k_mix|k_curl3d_v|k_halfg|k_conn_runs' --launch-skip 12 -c 4 -f -o gpurun_out/${TAG}_full_$dt $B --dtype $dt --steps 2 --warmup 3 > gpurun_out/${TAG}_ncu_f_$dt.log 2>&1
done
