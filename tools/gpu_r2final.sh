#!/usr/bin/env bash
# round-2 final 1-GPU pass: the whole parity suite, smoke, a short bench line (main workload only)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=5 --durations=8 ) > $O/r2f_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2f_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2f_smoke.log
( time timeout 200 python bench.py --no-extras --no-dropin --no-cpu-baseline ) > $O/r2f_bench_f64.json 2> $O/r2f_bench_f64.err
echo "bench rc=$?" >> $O/r2f_bench_f64.err
tail -14 $O/r2f_pytest.log; tail -3 $O/r2f_smoke.log; tail -3 $O/r2f_bench_f64.err; cut -c1-400 $O/r2f_bench_f64.json
