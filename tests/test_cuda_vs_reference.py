"""-m gpu: the drop-in path end to end.  oracle/_ref/<flavour>/ref_driver holds the UNMODIFIED
reference core; it builds each problem twice through the reference's own C++ API, steps twin A
with the reference's CPU code and twin B through the shim + libmeepcuda.so, syncs back and
compares every array.  Skipped when oracle/_ref did not travel to this machine."""
import json
import os
import subprocess
import tempfile

import pytest

from conftest import CUDA_LIB, ref_driver

pytestmark = pytest.mark.gpu

CASES = [("fp64", "vac3d:48", 60, 3, 1e-12), ("fp32", "vac3d:48", 60, 3, 1e-5),
         ("fp64", "vac3d_cplx:24", 40, 2, 1e-12), ("fp64", "nrslab3d:20", 30, 2, 1e-12),
         ("fp64", "lorentz3d:32", 50, 2, 1e-12), ("fp32", "lorentz3d:32", 50, 2, 1e-5),
         ("fp64", "phc3d:24", 60, 0, 1e-12), ("fp64", "wvg2d", 400, 5, 1e-12),
         ("fp64", "kr_per3d_rods", 300, 0, 1e-12), ("fp64", "kr_polariton1d", 100, 0, 1e-12),
         ("fp64", "beta2d", 200, 2, 1e-12), ("fp32", "beta2d", 100, 2, 1e-5), ("fp64", "beta2d_cplx:80", 100, 2, 1e-12),
         ("fp64", "bfast3d:20", 100, 2, 1e-12), ("fp32", "bfast3d:20", 100, 2, 1e-5), ("fp64", "bfast2d", 150, 2, 1e-12),
         ("fp64", "cond3d:24", 100, 2, 1e-12), ("fp32", "cond3d:24", 100, 2, 1e-5), ("fp64", "cond2d", 150, 2, 1e-12),
         ("fp64", "magdrude3d:24", 100, 2, 1e-12), ("fp32", "magdrude3d:24", 100, 2, 1e-5), ("fp64", "magdrude2d:60", 200, 2, 1e-12),
         ("fp64", "gyro3d:20", 100, 2, 1e-12), ("fp32", "gyro3d:20", 100, 2, 1e-5), ("fp64", "gyrodrude3d:20", 100, 2, 1e-12),
         ("fp64", "gyrosat3d:20", 100, 2, 1e-12),
         ("fp64", "dft2d:60", 200, 2, 1e-12), ("fp32", "dft2d:60", 200, 2, 1e-5), ("fp64", "dft3d:24", 100, 2, 1e-12),
         ("fp64", "aniso3d:24", 100, 2, 1e-12), ("fp32", "aniso3d:24", 100, 2, 1e-5),
         # cylindrical coordinates: f_rderiv_int pre-pass, i m / r terms through the PML / conductivity chain, r = 0 rules
         ("fp64", "cyl:30:0", 200, 0, 1e-12), ("fp64", "cyl:30:1", 200, 0, 1e-12), ("fp32", "cyl:30:-1", 200, 0, 1e-5),
         ("fp64", "cylper:30:2", 200, 2, 1e-12), ("fp64", "cylpml:40:0", 200, 2, 1e-12), ("fp64", "cylpml:40:1", 200, 0, 1e-12),
         ("fp32", "cylpml:40:1", 200, 0, 1e-5), ("fp64", "cylpml:40:3", 200, 0, 1e-12), ("fp64", "cylpmlnz:40:3", 200, 0, 1e-12),
         ("fp64", "cylcond:40:1", 200, 2, 1e-12), ("fp32", "cylcond:40:2", 200, 0, 1e-5), ("fp64", "cyllor:40:1", 200, 2, 1e-12),
         ("fp32", "cyllor:40:0", 200, 0, 1e-5), ("fp64", "cyllor:36:2", 200, 0, 1e-12), ("fp64", "cylpml:600:1", 60, 0, 1e-12),
         # multilevel-atom gain media (k_ml_levels, k_ml_pol, W_prev copies) and noisy Lorentzian media (k_pol_noise fed by
         # the reference's own generator)
         ("fp64", "mlatom3d:24", 100, 2, 1e-12), ("fp32", "mlatom3d:24", 100, 2, 1e-5), ("fp64", "mlatom2d:60", 200, 2, 1e-12),
         ("fp32", "mlatom2d:60", 200, 2, 1e-5), ("fp64", "mlatom3d:40:0.6", 60, 0, 1e-12),
         ("fp64", "mlatom1d", 40000, 0, 1e-12),
         ("fp64", "noisy3d:24", 100, 2, 1e-12), ("fp32", "noisy3d:24", 100, 2, 1e-5), ("fp64", "noisy2d:60", 200, 2, 1e-12),
         ("fp32", "noisy2d:60", 200, 2, 1e-5),
         ("fp64_pade", "kerr3d:20", 40, 2, 1e-12), ("fp64_pade", "kr_per3d_rods", 100, 0, 1e-12)]


@pytest.mark.parametrize("flavour,case,steps,warmup,tol", CASES)
def test_reference_twin_comparison(flavour, case, steps, warmup, tol):
    drv = ref_driver(flavour)
    if drv is None:
        pytest.skip("oracle/_ref/%s not present on this machine" % flavour)
    with tempfile.NamedTemporaryFile(suffix=".json") as tf:
        env = dict(os.environ, OMP_NUM_THREADS="8")
        subprocess.run([drv, "--mode", "shim", "--lib", CUDA_LIB, "--case", case, "--steps", str(steps),
                        "--warmup", str(warmup), "--nl", "1" if "pade" in flavour else "0", "--json", tf.name],
                       check=True, env=env,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=900)
        d = json.load(open(tf.name))
    assert d["backend"] == "cuda-sm_100a" and d["arrays"] > 0 and d["mismatched_arrays"] <= d["arrays"]
    assert d["finite"] == 1
    assert d["rel_l2"] <= tol, d


# BASELINE.json sizes (or the largest the reference's CPU twin finishes in seconds): the long-row 128-byte-pitch
# layout, the k_mix tile ordering and the 16-plane march are exercised where they are benchmarked
BASELINE_SIZE = [("fp64", "vac3d:256", 30, 1e-12), ("fp32", "vac3d:256", 30, 1e-5),
                 ("fp64", "lorentz3d:192:100", 30, 1e-12), ("fp64", "nrslab3d:64", 12, 1e-12),  # (the reference's NR twin: ~3 s per step at 96^3)
                 ("fp64", "phc3d:160:1.0", 30, 1e-12), ("fp32", "phc3d:160:1.0", 30, 1e-5)]


@pytest.mark.parametrize("flavour,case,steps,tol", BASELINE_SIZE)
def test_baseline_size_twins(flavour, case, steps, tol):
    """every array of the reference's CPU twin against the GPU twin at benchmark sizes"""
    drv = ref_driver(flavour)
    if drv is None:
        pytest.skip("oracle/_ref/%s not present on this machine" % flavour)
    with tempfile.NamedTemporaryFile(suffix=".json") as tf:
        env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 8), OMP_PROC_BIND="spread")
        subprocess.run([drv, "--mode", "shim", "--lib", CUDA_LIB, "--case", case, "--steps", str(steps),
                        "--warmup", "2", "--json", tf.name], check=True, env=env,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=1800)
        d = json.load(open(tf.name))
    assert d["backend"] == "cuda-sm_100a" and d["arrays"] > 0 and d["elements"] > 3e6 and d["finite"] == 1
    assert d["rel_l2"] <= tol and d["mismatched_elements"] < d["elements"], d


def test_known_results_on_gpu():
    """three of the reference's own golden numbers (tests/known_results.cpp:155-169), computed by
    the reference's objects stepping through libmeepcuda.so"""
    drv = ref_driver("fp64")
    if drv is None:
        pytest.skip("oracle/_ref/fp64 not present")
    known = {"kr_metal3d": (-41.8057, 200), "kr_pml2d": (0.883776, 600), "kr_per3d_rods": (-103.844, 600)}
    for case, (val, steps) in known.items():
        with tempfile.NamedTemporaryFile(suffix=".json") as tf:
            subprocess.run([drv, "--mode", "shim", "--lib", CUDA_LIB, "--case", case, "--steps", str(steps),
                            "--json", tf.name], check=True, stdout=subprocess.DEVNULL,
                           stderr=subprocess.DEVNULL, timeout=900)
            d = json.load(open(tf.name))
        assert abs(d["probe_abi"] - val) <= 1e-5 * abs(val), (case, d["probe_abi"], val)


@pytest.mark.parametrize("flavour,case,steps,tol", [("fp64", "vac3d:48", 60, 1e-12), ("fp32", "vac3d:48", 60, 1e-5),
                                                    ("fp64", "lorentz3d:32", 50, 1e-12), ("fp64", "wvg2d", 200, 1e-12),
                                                    ("fp64", "kr_per3d_rods", 100, 1e-12), ("fp64", "cond3d:24", 60, 1e-12),
                                                    ("fp64", "magdrude3d:24", 60, 1e-12), ("fp64", "aniso3d:24", 60, 1e-12),
                                                    ("fp64", "gyrosat3d:20", 60, 1e-12), ("fp64", "cylpml:40:1", 60, 1e-12),
                                                    ("fp64", "cyllor:40:0", 60, 1e-12)])
def test_synchronize_magnetic_fields_on_device(flavour, case, steps, tol):
    """fields::synchronize_magnetic_fields / restore_magnetic_fields (energy_and_flux.cpp:145-177) every 9
    steps: twin A on the reference's CPU path, twin B through meepcu_synchronize_magnetic_fields on the GPU"""
    drv = ref_driver(flavour)
    if drv is None:
        pytest.skip("oracle/_ref/%s not present on this machine" % flavour)
    with tempfile.NamedTemporaryFile(suffix=".json") as tf:
        env = dict(os.environ, OMP_NUM_THREADS="8")
        subprocess.run([drv, "--mode", "shim", "--lib", CUDA_LIB, "--case", case, "--steps", str(steps),
                        "--warmup", "2", "--syncmag", "9", "--json", tf.name], check=True, env=env,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=900)
        d = json.load(open(tf.name))
    assert d["backend"] == "cuda-sm_100a" and d["sync_checks"] >= 5
    assert d["sync_rel_l2"] <= tol and d["rel_l2"] <= tol, d


@pytest.mark.parametrize("flavour,case,chunks,tol", [("fp64", "vac3d:48", 4, 1e-12), ("fp32", "lorentz3d:32", 3, 1e-5),
                                                     ("fp64", "nrslab3d:20", 3, 1e-12), ("fp64", "aniso3d:24", 5, 1e-12),
                                                     ("fp64", "dft3d:24", 3, 1e-12), ("fp64", "cyllor:48:1", 3, 1e-12),
                                                     ("fp32", "cylpml:48:2", 4, 1e-5)])
def test_reference_chunk_splitting_on_gpu(flavour, case, chunks, tol):
    """the reference's own cost-based chunk splitting (structure(..., num_chunks)): 40-50 chunks instead of 27"""
    drv = ref_driver(flavour)
    if drv is None:
        pytest.skip("oracle/_ref/%s not present on this machine" % flavour)
    with tempfile.NamedTemporaryFile(suffix=".json") as tf:
        env = dict(os.environ, OMP_NUM_THREADS="8")
        subprocess.run([drv, "--mode", "shim", "--lib", CUDA_LIB, "--case", case, "--chunks", str(chunks), "--steps", "60",
                        "--warmup", "2", "--json", tf.name], check=True, env=env,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=900)
        d = json.load(open(tf.name))
    assert d["backend"] == "cuda-sm_100a" and d["num_chunks"] > chunks and d["finite"] == 1
    assert d["rel_l2"] <= tol, d
