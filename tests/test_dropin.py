"""The drop-in seam itself: oracle/_ref/fp64_dropin is the reference library with the maintainer's patch
(meep_nl_b200/shim/dropin_patch.py: one data member, three methods, a one-line prologue in fields::step())
applied in the compiler's input pipe.  Twin B is stepped by the reference's OWN `fields::step()` -- unchanged
caller code -- which delegates to the C ABI library; a host-side reader runs in mid-simulation; at the end
every array must equal twin A's (stepped by the unpatched code path of the same library)."""
import json
import os
import subprocess
import tempfile

import pytest

from conftest import CUDA_LIB, ORACLE_LIB, ref_driver


def run(lib, case, steps, nl=0, syncmag=0, addsrc=0):
    drv = ref_driver("fp64_dropin")
    if drv is None:
        pytest.skip("oracle/_ref/fp64_dropin not built (reference tree absent)")
    with tempfile.NamedTemporaryFile(suffix=".json") as tf:
        env = dict(os.environ, OMP_NUM_THREADS="4")
        subprocess.run([drv, "--mode", "dropin", "--lib", lib, "--case", case, "--steps", str(steps), "--warmup", "2",
                        "--nl", str(nl), "--syncmag", str(syncmag), "--addsrc", str(addsrc), "--json", tf.name], check=True, env=env, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL, timeout=900)
        return json.load(open(tf.name))


@pytest.mark.parametrize("case", ["vac3d:32", "lorentz3d:16", "wvg2d", "nrslab3d:14", "kr_xper_ypml3d", "gyro3d:12", "dft3d:12",
                                  "mlatom3d:12", "noisy2d:20"])
def test_patched_reference_steps_through_the_abi(case):
    d = run(ORACLE_LIB, case, 40)
    assert d["attached"] == 1 and d["t_ref"] == d["t_abi"]
    assert d["arrays"] > 0 and d["mismatched_elements"] == 0, d
    assert d["probe_ref"] == d["probe_abi"]
    # mid-run readers called straight on the patched reference (no explicit sync): get_field and, through
    # loop_in_chunks, electric_energy_in_box
    assert d["probe_mid"] == d["probe_mid_ref"] and d["energy_mid_abi"] == d["energy_mid_ref"]
    # DFT consumers called on the patched reference in mid-run and at the end (dft_flux::flux), with a
    # scale_dfts() -- a host-side writer of the accumulators -- in between
    assert d["flux_mid_mismatched"] == 0 and d["flux_end_mismatched"] == 0, d
    if case in ("lorentz3d:16", "dft3d:12"):
        assert d["flux_values"] > 0
        # ... and they were SUMMED behind the ABI (meepcu_dft_dot), not fetched: dft_flux::flux, dft_energy::electric/magnetic
        assert d["device_dft_dots"] > 0
    # energy_and_flux.cpp's integrals (electric / magnetic / field_energy_in_box, flux_in_box with its synchronize /
    # restore pair) were formed behind the ABI too (meepcu_integrate_dot): same bits as the reference's own loops
    assert d["box_values"] == 4 and d["box_mismatched"] == 0 and d["device_integrals"] >= 10, d


@pytest.mark.gpu
@pytest.mark.parametrize("case,tol", [("vac3d:48", 1e-12), ("lorentz3d:32", 1e-12), ("nrslab3d:20", 1e-12), ("aniso3d:24", 1e-12),
                                      ("mlatom3d:20", 1e-12), ("noisy3d:20", 1e-12)])
def test_patched_reference_steps_on_the_gpu(case, tol):
    d = run(CUDA_LIB, case, 60)
    assert d["attached"] == 1 and d["t_ref"] == d["t_abi"]
    assert d["rel_l2"] <= tol, d
    assert abs(d["probe_mid"] - d["probe_mid_ref"]) <= tol * max(1.0, abs(d["probe_mid_ref"]))
    assert abs(d["energy_mid_abi"] - d["energy_mid_ref"]) <= 1e-9 * max(1e-30, abs(d["energy_mid_ref"]))
    # flux spectra read by the reference's own dft_flux::flux() while the accumulators live on the device
    # (stale or zero spectra would show here), before and after a host-side scale_dfts()
    assert d["flux_mid_rel"] <= tol and d["flux_end_rel"] <= tol, d
    if case.startswith("lorentz3d"):
        assert d["flux_values"] > 0 and d["device_dft_dots"] > 0
    # the box integrals are reduced on the GPU (k_integrate_dot): double tree sums against the reference's long double loop
    assert d["box_values"] == 4 and d["device_integrals"] >= 10 and d["box_rel"] <= tol, d


@pytest.mark.gpu
def test_flux_spectrum_read_from_the_device():
    d = run(CUDA_LIB, "dft3d:16", 60)
    assert d["flux_values"] > 0 and d["flux_mid_rel"] <= 1e-12 and d["flux_end_rel"] <= 1e-12, d
    assert d["device_dft_dots"] > 0 and d["box_rel"] <= 1e-12  # k_dft_dot / k_integrate_dot, incl. the centred-grid averages of flux_in_box
    assert d["rel_l2"] <= 1e-12


def run_time(lib, case, steps):
    drv = ref_driver("fp64_dropin")
    if drv is None:
        pytest.skip("oracle/_ref/fp64_dropin not built (reference tree absent)")
    with tempfile.NamedTemporaryFile(suffix=".json") as tf:
        env = dict(os.environ, OMP_NUM_THREADS="4")
        subprocess.run([drv, "--mode", "dropin_time", "--lib", lib, "--case", case, "--steps", str(steps), "--warmup", "3",
                        "--json", tf.name], check=True, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL,
                       timeout=900)
        return json.load(open(tf.name))


def test_dropin_timing_mode_runs_one_step_per_call():
    """bench.py's e2e_dropin leg: fields::step() of the patched reference called once per time step"""
    d = run_time(ORACLE_LIB, "vac3d:24", 10)
    assert d["t"] == 14 and d["steps"] == 10 and d["ms_per_step"] > 0 and d["probe"] != 0


@pytest.mark.gpu
def test_dropin_step_does_not_wait_for_the_device():
    """one fields::step() call per time step must cost about one device step: the call only queues work
    (no per-step scan of whole arrays, no per-step synchronisation)"""
    d = run_time(CUDA_LIB, "vac3d:128", 200)
    assert d["t"] == 204 and d["probe"] != 0
    assert d["kernel_launches_per_step"] <= 12, d
    assert d["enqueue_seconds"] < d["seconds"]
    assert d["ms_per_step"] <= 1.5 * d["device_ms_per_step"] + 0.05, d


@pytest.mark.parametrize("case,every", [("lorentz3d:16", 9), ("magdrude3d:12", 7), ("gyrosat3d:12", 9), ("wvg2d", 10)])
def test_patched_reference_synchronizes_magnetic_fields_through_the_abi(case, every):
    """fields::synchronize_magnetic_fields() / restore_magnetic_fields() of the patched reference delegate their
    bodies (energy_and_flux.cpp:145-177) when the fields live behind the ABI; a third twin runs the CPU bodies"""
    d = run(ORACLE_LIB, case, 40, syncmag=every)
    assert d["sync_checks"] >= 4 and d["sync_mismatched_elements"] == 0
    assert d["mismatched_elements"] == 0 and d["t_ref"] == d["t_abi"]


@pytest.mark.gpu
def test_patched_reference_synchronizes_on_the_gpu():
    d = run(CUDA_LIB, "magdrude3d:20", 45, syncmag=9)
    assert d["sync_checks"] >= 5 and d["sync_mismatched_elements"] <= d["elements"]
    assert d["rel_l2"] <= 1e-12, d


@pytest.mark.parametrize("case", ["vac3d:32", "lorentz3d:16", "wvg2d"])
def test_patched_reference_reattaches_when_the_problem_changes(case):
    """a point source and a flux monitor are added in mid-run with the reference's own calls; the patched step()
    notices that the description is outdated, syncs back, re-describes and goes on (host-side logic of the shim:
    exercised here through the CPU oracle)"""
    d = run(ORACLE_LIB, case, 45, addsrc=1)
    assert d["attached"] == 1 and d["t_ref"] == d["t_abi"]
    assert d["arrays"] > 0 and d["mismatched_elements"] == 0, d
    assert d["probe_mid"] == d["probe_mid_ref"]


@pytest.mark.parametrize("case", ["lorentz3d:16", "noisy3d:10"])
def test_environment_variable_opts_an_unchanged_program_in(case):
    """MEEP_CUDA_LIB in the environment of an UNCHANGED caller (ref_driver --mode ref only calls fields::step() and
    get_field, like python/simulation.py:2831-2834): the patched library delegates by itself, and the probe equals the
    CPU run's bit for bit (CPU oracle behind the ABI here)"""
    drv = ref_driver("fp64_dropin")
    if drv is None:
        pytest.skip("oracle/_ref/fp64_dropin not built (reference tree absent)")
    out = []
    for lib in (None, ORACLE_LIB):
        with tempfile.NamedTemporaryFile(suffix=".json") as tf:
            env = dict(os.environ, OMP_NUM_THREADS="2")
            env.pop("MEEP_CUDA_LIB", None)
            if lib:
                env["MEEP_CUDA_LIB"] = lib
            subprocess.run([drv, "--mode", "ref", "--case", case, "--steps", "40", "--json", tf.name], check=True, env=env,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600)
            out.append(json.load(open(tf.name)))
    assert out[0]["attached"] == 0 and out[1]["attached"] == 1
    assert out[0]["probe"] == out[1]["probe"] and out[0]["probe"] != 0
