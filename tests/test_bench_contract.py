"""bench.py's reference arm (the CPU leg the driver runs beside the GPU arm) prints ONE JSON line with the
contract's keys; the GPU arm itself needs a device and is exercised by the driver."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT, ref_driver

REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


@pytest.mark.parametrize("workload,size", [("vac3d", 40), ("phc3d", 32)])
def test_reference_arm_line(workload, size):
    if ref_driver("fp64") is None:
        pytest.skip("oracle/_ref/fp64 not built (reference tree absent)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload,
                        "--size", str(size), "--steps", "4", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= set(d) and d["impl"] == "reference" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"].startswith(workload)


def test_other_ranks_of_the_reference_arm_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_both_arms_print_the_same_config():
    """the driver compares the two arms' `config`, `steps` and `warmup`: they come from one function / are
    honoured exactly"""
    import bench
    a = bench.parse(["--gpus", "4", "--steps", "7", "--warmup", "4"])
    assert bench.config_of(a, 4)["workload"] == "vac3d_1024x256x256_f64"
    assert bench.config_of(a, 4) == bench.config_of(bench.parse(["--impl", "reference", "--gpus", "4"]), 4)
    a5 = bench.parse(["--workload", "phc3d", "--size", "1024", "--dtype", "f32", "--scaling", "strong", "--slab-x", "1024"])
    assert bench.grid_dims(a5, 8) == (1024, 1024, 1536)
    a5w = bench.parse(["--workload", "phc3d", "--size", "1024", "--dtype", "f32", "--scaling", "weak", "--slab-x", "128"])
    assert bench.grid_dims(a5w, 8) == (1024, 1024, 1536) and bench.grid_dims(a5w, 2) == (256, 1024, 1536)


def test_reference_arm_honours_steps_and_warmup():
    if ref_driver("fp64") is None:
        pytest.skip("oracle/_ref/fp64 not built (reference tree absent)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "32", "--steps", "7",
                        "--warmup", "4"], capture_output=True, text=True, timeout=600)
    d = json.loads(r.stdout.strip())
    assert d["steps"] == 7 and d["warmup"] == 4 and d["config"]["workload"] == "vac3d_32x32x32_f64"


@pytest.mark.parametrize("extra", [["--workload", "vac3d", "--size", "24"],
                                   ["--workload", "phc3d", "--size", "32", "--scaling", "strong", "--slab-x", "32"]])
def test_n_rank_path_of_the_gpu_arm_dry_run(extra):
    """bench.py's own N > 1 code (slab decomposition, the N-rank == 1-rank parity gate on every array, batches,
    e2e, JSON line) with gloo + the CPU oracle standing in for NCCL + libmeepcuda.so (tests/bench_dryrun_worker.py)"""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests", "bench_dryrun_worker.py"), "--gpus", "2", "--steps", "4",
           "--warmup", "3", "--no-extras", "--no-dropin", "--no-cpu-baseline", "--min-seconds", "0.01"] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS="2"), timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["parity"]["ranks"] == 2 and d["parity"]["arrays"] > 20
    assert d["parity"]["mismatched_arrays"] == 0 and d["parity"]["mismatched_elements"] == 0
    assert d["parity"]["nonzero_arrays"] > 10 and d["details"]["probe_at_source"] not in (None, 0.0)
    assert d["timing"]["batches"] >= 3 and d["timing"]["steps_per_batch"] == 4
