"""Runs the UNMODIFIED reference (oracle/_ref/<flavour>/ref_driver, built from /root/reference by
oracle/build_ref.sh) against the oracle through the reference-side shim, in one process, and
requires every array (fields, PML auxiliaries, polarisations, DFT accumulators) to be identical.
Skipped where oracle/_ref has not been built (no reference tree)."""
import json
import os
import subprocess
import tempfile

import pytest

from conftest import ORACLE_LIB, ref_driver

CASES = [("fp64", "vac3d:24", 25, 2), ("fp64", "kr_per3d_rods", 40, 0), ("fp64", "nr3d:10", 25, 0),
         ("fp64", "lorentz3d:16", 30, 1), ("fp64", "wvg2d", 60, 3), ("fp32", "vac3d_cplx:14", 25, 2),
         ("fp32", "lorentz3d:16", 30, 1), ("fp64", "vac3d:26", 25, 0),
         # 2-D exp(i beta z) term (step_beta), real and complex fields
         ("fp64", "beta2d", 80, 2), ("fp64", "beta2d_cplx", 60, 2), ("fp32", "beta2d", 60, 2),
         # BFAST terms (step_bfast) and the conductivity + PML (f_cond) variants of step_curl
         ("fp64", "bfast3d", 60, 2), ("fp32", "bfast3d", 60, 2), ("fp64", "bfast2d", 80, 2),
         ("fp64", "cond3d", 60, 2), ("fp32", "cond3d", 60, 2), ("fp64", "cond2d", 80, 2),
         # anisotropic Lorentzian sigma: the 3x3 and 2x2 branches of update_P (susceptibility.cpp:225-251)
         ("fp64", "aniso3d", 60, 2), ("fp32", "aniso3d", 60, 2),
         # every flavour of dft_chunk: flux (decimation 1 / auto), dft_fields on the centred grid, dft_energy (D, B),
         # complex fields with a Bloch phase
         # mu != 1 (H split from B), a magnetic Lorentzian (pol[H_stuff]) and an electric Drude term
         ("fp64", "magdrude3d", 60, 2), ("fp32", "magdrude3d", 60, 2), ("fp64", "magdrude2d", 80, 2),
         # gyrotropic media: Lorentzian, Drude and saturated (Landau-Lifshitz-Gilbert) models (susceptibility.cpp:350-603)
         ("fp64", "gyro3d", 60, 2), ("fp32", "gyro3d", 60, 2), ("fp64", "gyrodrude3d", 60, 2), ("fp64", "gyrosat3d", 60, 2),
         ("fp32", "gyrosat3d", 60, 2),
         ("fp64", "dft2d", 60, 2), ("fp32", "dft2d", 60, 2), ("fp64", "dft3d", 60, 2), ("fp32", "dft3d", 60, 2),
         # cylindrical coordinates (step_db.cpp:86-122,178-457): metallic / periodic-in-z / PML cells, m = -1 .. 3 (the
         # r = 0 rules of m = 0, |m| = 1 and the zeroed rows of larger |m|, with and without zero_fields_near_cylorigin),
         # conductivity (f_cond paths of the i m / r terms), a Lorentzian rod with DFT flux planes and dft_fields
         ("fp64", "cyl:15:0", 120, 0), ("fp64", "cyl:15:1", 120, 0), ("fp32", "cyl:15:-1", 120, 0), ("fp64", "cylper:15:2", 120, 2),
         ("fp32", "cylper:16:0", 120, 0), ("fp64", "cylpml:18:0", 120, 2), ("fp64", "cylpml:18:1", 120, 0),
         ("fp32", "cylpml:18:1", 120, 0), ("fp64", "cylpml:18:3", 120, 0), ("fp64", "cylpmlnz:18:3", 120, 0),
         ("fp64", "cylcond:18:1", 120, 2), ("fp32", "cylcond:18:2", 120, 0), ("fp64", "cyllor:20:1", 150, 2),
         ("fp32", "cyllor:20:0", 150, 0), ("fp64", "cyllor:18:2", 150, 0),
         # multilevel-atom gain media (multilevel-atom.cpp: populations on the centred grid, one polarisation per
         # transition, W_prev) -- two levels / one transition with real fields, three levels / two transitions with
         # complex fields next to an ordinary Lorentzian
         ("fp64", "mlatom3d", 60, 2), ("fp32", "mlatom3d", 60, 2), ("fp64", "mlatom2d", 80, 2), ("fp32", "mlatom2d", 80, 2),
         ("fp64", "mlatom3d:18:0.6", 100, 0),
         # ... and the 1-D laser cavity of python/tests/test_multilevel_atom.py through its first relaxation spikes
         ("fp64", "mlatom1d", 120000, 0), ("fp32", "mlatom1d", 60000, 0),
         # noisy Lorentzian media (susceptibility.cpp:317-339): the reference's own generator draws the unit Gaussians in
         # its loop order for both twins; electric, and magnetic Drude + electric in one step with complex fields
         ("fp64", "noisy3d", 60, 2), ("fp32", "noisy3d", 60, 2), ("fp64", "noisy2d", 80, 2), ("fp32", "noisy2d", 80, 2),
         # cells reduced by a symmetry (tests/symmetry.cpp): two mirror planes in 3-D and 2-D, a four-fold rotation in 2-D
         # with complex fields -- the chunks connect to themselves through the symmetry's COPY / NEGATE / PHASE lists
         ("fp64", "sym3d", 60, 2), ("fp32", "sym3d", 60, 2), ("fp64", "sym2d", 80, 2), ("fp64", "symrot2d", 80, 2),
         ("fp32", "symrot2d", 80, 2),
         # every kind of current source (plane source with an amplitude profile, continuous, a NON-integrated custom source on
         # an electric component, a magnetic line source) and the remaining DFT consumers (near-to-far face, force monitor)
         ("fp64", "srcs3d", 80, 2), ("fp32", "srcs3d", 80, 2), ("fp64", "srcs2d", 100, 2), ("fp32", "srcs2d", 100, 2),
         # stretched-coordinate PML (mean_stretch != 1: kap != 1), per-face thickness / Rasymptotic, one face without PML
         ("fp64", "pmlk3d", 120, 2), ("fp32", "pmlk3d", 120, 2), ("fp64", "pmlk2d", 150, 2), ("fp32", "pmlk2d", 150, 2),
         # conductivity + PML with complex fields and a Bloch phase; a cylindrical cell periodic in z with a phase; four
         # polarisations on one chunk in a cavity whose metallic walls the medium touches
         ("fp64", "cond3d_cplx", 80, 2), ("fp32", "cond3d_cplx", 80, 2), ("fp64", "cylperk:15:1", 120, 0), ("fp32", "cylperk:16:0", 120, 0),
         ("fp64", "lor4", 100, 2), ("fp32", "lor4", 100, 2),
         # UPSTREAM_PADE against the de-blocked reference flavour (oracle/deblock_pade.py)
         ("fp64_pade", "kerr3d:14", 40, 2), ("fp64_pade", "kerr1d", 80, 0), ("fp64_pade", "kr_per3d_rods", 40, 0)]


@pytest.mark.parametrize("flavour,case,steps,warmup", CASES)
def test_shim_plus_oracle_equals_reference(flavour, case, steps, warmup):
    drv = ref_driver(flavour)
    if drv is None:
        pytest.skip("oracle/_ref/%s not built (reference tree absent)" % flavour)
    with tempfile.NamedTemporaryFile(suffix=".json") as tf:
        env = dict(os.environ, OMP_NUM_THREADS="4")
        subprocess.run([drv, "--mode", "shim", "--lib", ORACLE_LIB, "--case", case, "--steps", str(steps),
                        "--warmup", str(warmup), "--nl", "1" if "pade" in flavour else "0", "--json", tf.name],
                       check=True, env=env,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600)
        d = json.load(open(tf.name))
    assert d["arrays"] > 0 and d["mismatched_arrays"] == 0 and d["mismatched_elements"] == 0, d
    assert d["probe_ref"] == d["probe_abi"]


def test_reference_known_answers_through_the_oracle():
    """tests/known_results.cpp:155-169: the reference's own golden numbers, recomputed by
    stepping through the C ABI (oracle backend) instead of fields::step."""
    drv = ref_driver("fp64")
    if drv is None:
        pytest.skip("oracle/_ref/fp64 not built")
    # the twelve field values of the list (the thirteenth is an energy integral, host code outside the path)
    known = {"kr_polariton1d": (-0.0894851, 200), "kr_metal2d": (5.20605, 200), "kr_pml2d": (0.883776, 600),
             "kr_xper2d": (0.110425, 600), "kr_per2d_rods": (-4.78767, 600), "kr_per2d": (1.12502, 600),
             "kr_xper_ypml2d": (0.608815, 600), "kr_metal3d": (-41.8057, 200), "kr_xper3d": (-100.758, 600),
             "kr_xper_ypml3d": (-101.398, 600), "kr_per3d_rods": (-103.844, 600), "kr_per3d": (-99.1618, 600)}
    for case, (val, steps) in known.items():
        with tempfile.NamedTemporaryFile(suffix=".json") as tf:
            subprocess.run([drv, "--mode", "shim", "--lib", ORACLE_LIB, "--case", case, "--steps", str(steps),
                            "--json", tf.name], check=True, stdout=subprocess.DEVNULL,
                           stderr=subprocess.DEVNULL, timeout=600)
            d = json.load(open(tf.name))
        assert d["mismatched_elements"] == 0
        assert abs(d["probe_abi"] - val) <= 1e-5 * abs(val), (case, d["probe_abi"], val)


SYNC_CASES = [("fp64", "vac3d:32", 40), ("fp64", "wvg2d", 60), ("fp64", "lorentz3d:16", 40), ("fp32", "vac3d_cplx:14", 30),
              ("fp64", "kr_per3d_rods", 40), ("fp64", "cond3d:12", 40), ("fp64", "bfast3d:10", 40), ("fp32", "beta2d", 40),
              ("fp64", "magdrude3d", 40), ("fp32", "magdrude2d", 40), ("fp64", "gyrosat3d", 40),
              ("fp64", "cylpml:18:1", 40), ("fp64", "cyllor:18:0", 40), ("fp32", "cylcond:18:2", 40),
              ("fp64", "mlatom3d", 40), ("fp32", "noisy2d", 40), ("fp64", "pmlk3d", 40), ("fp64", "srcs2d", 40), ("fp64", "sym3d", 40)]


@pytest.mark.parametrize("flavour,case,steps", SYNC_CASES)
def test_synchronize_magnetic_fields_equals_reference(flavour, case, steps):
    """every 7 steps both twins call synchronize_magnetic_fields (reference: energy_and_flux.cpp:145-166 on
    the CPU; twin B: meepcu_synchronize_magnetic_fields), all arrays must be identical while synchronized,
    after restore_magnetic_fields, and at the end of the run"""
    drv = ref_driver(flavour)
    if drv is None:
        pytest.skip("oracle/_ref/%s not built (reference tree absent)" % flavour)
    with tempfile.NamedTemporaryFile(suffix=".json") as tf:
        env = dict(os.environ, OMP_NUM_THREADS="4")
        subprocess.run([drv, "--mode", "shim", "--lib", ORACLE_LIB, "--case", case, "--steps", str(steps),
                        "--warmup", "2", "--syncmag", "7", "--json", tf.name], check=True, env=env,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600)
        d = json.load(open(tf.name))
    assert d["sync_checks"] >= 4 and d["sync_mismatched_elements"] == 0
    assert d["arrays"] > 0 and d["mismatched_elements"] == 0


@pytest.mark.parametrize("flavour,case,chunks", [("fp64", "vac3d:32", 5), ("fp64", "lorentz3d:16", 4), ("fp32", "wvg2d", 6),
                                                 ("fp64", "nrslab3d:14", 3), ("fp64", "aniso3d:14", 3),
                                                 ("fp64", "cyllor:24:1", 3), ("fp32", "cylpml:24:2", 4),
                                                 ("fp64", "mlatom3d:16", 3), ("fp64", "noisy3d:16", 3)])
def test_reference_chunk_splitting(flavour, case, chunks):
    """structure(..., num_chunks) lets the reference split the cell by cost into many more chunks (40-50):
    the description the shim hands over changes completely, the fields must not"""
    drv = ref_driver(flavour)
    if drv is None:
        pytest.skip("oracle/_ref/%s not built (reference tree absent)" % flavour)
    with tempfile.NamedTemporaryFile(suffix=".json") as tf:
        env = dict(os.environ, OMP_NUM_THREADS="4")
        subprocess.run([drv, "--mode", "shim", "--lib", ORACLE_LIB, "--case", case, "--chunks", str(chunks), "--steps", "40",
                        "--warmup", "2", "--json", tf.name], check=True, env=env,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600)
        d = json.load(open(tf.name))
    assert d["num_chunks"] > chunks and d["arrays"] > 0 and d["mismatched_elements"] == 0, d
