set -x
B="python bench.py --no-cpu-baseline"
for dt in f64 f32; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s6_launches_$dt.csv $B --dtype $dt --steps 3 --warmup 3 > gpurun_out/s6_ncu_l_$dt.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:'k_curl3d_v|k_halfg' --launch-skip 8 -c 4 -f -o gpurun_out/s6_full_$dt $B --dtype $dt --steps 2 --warmup 3 > gpurun_out/s6_ncu_f_$dt.log 2>&1
done
ls -la gpurun_out
