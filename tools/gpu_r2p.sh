#!/usr/bin/env bash
# round-2 pass p: final state -- parity suite, default bench line + reference arm, fp32, ncu launch list (fp64)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q --durations=5 ) > $O/r2p_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2p_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2p_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2p_smoke.log
( time timeout 900 python bench.py ) > $O/r2p_bench_f64.json 2> $O/r2p_bench_f64.err
echo "bench rc=$?" >> $O/r2p_bench_f64.err
( time timeout 300 python bench.py --impl reference --steps 100 --warmup 5 ) > $O/r2p_bench_ref.json 2> $O/r2p_bench_ref.err
B="--no-extras --no-dropin --no-cpu-baseline --steps 20 --warmup 5"
timeout 300 python bench.py $B --dtype f32 > $O/r2p_bench_f32.json 2> $O/r2p_bench_f32.err
NB="python bench.py --no-cpu-baseline --no-extras --no-dropin --min-seconds 0.001"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2p_launches_f64.csv $NB --dtype f64 --steps 3 --warmup 3 > $O/r2p_ncu_l_f64.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_mix|k_conn_runs' --launch-skip 12 -c 4 -f -o $O/r2p_full_f32 $NB --dtype f32 --steps 2 --warmup 3 > $O/r2p_ncu_f_f32.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_curl3d_v' --launch-skip 6 -c 2 -f -o $O/r2p_full_phc $NB --workload phc3d --size 512 --dtype f32 --steps 2 --warmup 3 > $O/r2p_ncu_f_phc.log 2>&1
tail -4 $O/r2p_pytest.log; tail -3 $O/r2p_smoke.log; tail -3 $O/r2p_bench_f64.err; cut -c1-300 $O/r2p_bench_f64.json
