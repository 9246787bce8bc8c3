"""-m gpu, run LAST: twin cases that were added after the round's GPU budget had been spent.  They are pinned
bit-for-bit between the reference and the CPU oracle (tests/test_oracle_vs_reference.py) and have never run on a
B200: their first GPU run is whoever runs this file next.  Marked xfail(strict=False) so that the outcome is reported
(XPASS = the CUDA path reproduces the reference on them too, xfailed = it does not) without deciding whether the suite
of validated tests is green; every run is bounded to two minutes."""
import json
import os
import subprocess
import tempfile

import pytest

from conftest import CUDA_LIB, ref_driver

pytestmark = pytest.mark.gpu

CASES = [("fp64", "sym3d", 60, 2, 1e-12), ("fp32", "sym3d", 60, 2, 1e-5), ("fp64", "sym2d", 80, 2, 1e-12),
         ("fp64", "symrot2d", 80, 2, 1e-12), ("fp32", "symrot2d", 80, 2, 1e-5),
         ("fp64", "srcs3d:24", 80, 2, 1e-12), ("fp32", "srcs3d:24", 80, 2, 1e-5), ("fp64", "srcs2d:48", 100, 2, 1e-12),
         ("fp64", "pmlk3d:24", 120, 2, 1e-12), ("fp32", "pmlk3d:24", 120, 2, 1e-5), ("fp64", "pmlk2d:48", 150, 2, 1e-12),
         ("fp64", "cond3d_cplx:24", 80, 2, 1e-12), ("fp32", "cond3d_cplx:24", 80, 2, 1e-5), ("fp64", "cylperk:30:1", 200, 0, 1e-12),
         ("fp64", "lor4:24", 100, 2, 1e-12), ("fp32", "lor4:24", 100, 2, 1e-5)]


@pytest.mark.xfail(strict=False, reason="first GPU run of these cases (added after the round's GPU budget was spent)")
@pytest.mark.parametrize("flavour,case,steps,warmup,tol", CASES)
def test_late_twin_cases_on_the_gpu(flavour, case, steps, warmup, tol):
    drv = ref_driver(flavour)
    if drv is None:
        pytest.skip("oracle/_ref/%s not present on this machine" % flavour)
    with tempfile.NamedTemporaryFile(suffix=".json") as tf:
        env = dict(os.environ, OMP_NUM_THREADS="8")
        subprocess.run([drv, "--mode", "shim", "--lib", CUDA_LIB, "--case", case, "--steps", str(steps),
                        "--warmup", str(warmup), "--json", tf.name], check=True, env=env,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=120)
        d = json.load(open(tf.name))
    assert d["backend"] == "cuda-sm_100a" and d["arrays"] > 0 and d["finite"] == 1
    assert d["rel_l2"] <= tol, d
