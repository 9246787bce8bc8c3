#!/usr/bin/env bash
# round-2 pass u: the whole GPU suite on the final tree
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q --durations=5 ) > $O/r2u_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2u_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2u_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2u_smoke.log
tail -4 $O/r2u_pytest.log; tail -2 $O/r2u_smoke.log
