"""world_size-2 tests of the slab decomposition + cross-rank halo exchange:
the reference's own invariant -- a problem split over several processes must give the fields of
the unsplit problem (tests/three_d.cpp run under mpirun -np 2) -- with bit-exact comparison.
CPU: oracle backend, halo buffers moved by gloo.  -m gpu (needs 2 GPUs): libmeepcuda.so + NCCL."""
import json
import os
import subprocess
import sys
import tempfile

import pytest

from conftest import ROOT

WORKER = os.path.join(ROOT, "tests", "multirank_worker.py")


def run(kind, case, dtype, steps, world, port, extra_env=None):
    out = tempfile.mkdtemp()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER, kind, case, dtype, str(steps), out]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    env.update(extra_env or {})
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    vals = {}
    for k in range(world):
        vals.update(json.load(open(os.path.join(out, "rank%d.json" % k))))
    return vals


@pytest.mark.parametrize("case", ["pml", "per", "beta2d"])
def test_two_ranks_equal_one_rank_oracle(case):
    one = run("oracle", case + "_presplit", "f64", 40, 1, 29611)
    two = run("oracle", case, "f64", 40, 2, 29612)
    assert len(one) > 40 and max(abs(v) for v in one.values()) > 0
    assert one.keys() == two.keys()
    assert all(one[k] == two[k] for k in one), [k for k in one if one[k] != two[k]][:5]


# short tiles (MEEPCU_SEG / MEEPCU_GSEG = planes per tile) so that even this small grid has boundary
# tiles AND interior tiles: the halo exchange then really runs beside the interior tiles
SHORT_TILES = {"MEEPCU_SEG": "4", "MEEPCU_GSEG": "4"}


@pytest.mark.gpu
@pytest.mark.parametrize("case,dtype,env", [("pml", "f64", None), ("per", "f64", None), ("pml", "f32", None),
                                            ("pml", "f64", SHORT_TILES), ("per", "f64", SHORT_TILES),
                                            ("pml", "f32", SHORT_TILES),
                                            ("pml", "f64", {"MEEPCU_NO_OVERLAP": "1"})])
def test_two_gpus_equal_oracle(case, dtype, env):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ref = run("oracle", case + "_presplit", dtype, 40, 1, 29613)
    two = run("cuda", case, dtype, 40, 2, 29614, env)
    assert ref.keys() == two.keys()
    tol = 1e-12 if dtype == "f64" else 1e-5
    num = sum((ref[k] - two[k]) ** 2 for k in ref) ** 0.5
    den = sum(ref[k] ** 2 for k in ref) ** 0.5
    assert den > 0 and num / den <= tol
