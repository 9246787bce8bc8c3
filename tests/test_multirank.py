"""world_size > 1 tests of the slab decomposition + cross-rank halo exchange:
the reference's own invariant -- a problem split over several processes must give the fields of
the unsplit problem (tests/three_d.cpp:95-229 run under mpirun) -- on EVERY array every rank owns
(fields, PML auxiliaries, polarisations, DFT accumulators), incl. the not-owned halo points.
CPU: oracle backend, halo buffers moved by gloo, bit-exact.  -m gpu (needs 2 / 4 / 8 GPUs): libmeepcuda.so
+ NCCL against the oracle, 1e-12 / 1e-5."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from conftest import ROOT

WORKER = os.path.join(ROOT, "tests", "multirank_worker.py")


def start(kind, case, dtype, steps, world, port, extra_env=None, nslabs=2):
    out = tempfile.mkdtemp()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER, kind, case, dtype, str(steps), out, str(nslabs)]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    env.update(extra_env or {})
    return subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env), out, world


def finish(job):
    proc, out, world = job
    _stdout, stderr = proc.communicate(timeout=900)
    assert proc.returncode == 0, stderr[-3000:]
    vals = {}
    for k in range(world):
        with np.load(os.path.join(out, "rank%d.npz" % k)) as z:
            for name in z.files:
                assert name not in vals, "two ranks own " + name
                vals[name] = z[name]
    return vals


def run(kind, case, dtype, steps, world, port, extra_env=None, nslabs=2):
    return finish(start(kind, case, dtype, steps, world, port, extra_env, nslabs))


def run_both(one_args, many_args):
    """the single-rank and the multi-rank job of a CPU test side by side (separate rendezvous ports)"""
    a, b = start(*one_args[0], **one_args[1]), start(*many_args[0], **many_args[1])
    return finish(a), finish(b)


def bits(a):
    return a.view(np.uint64 if a.dtype.itemsize % 8 == 0 else np.uint32)


def assert_bit_equal(one, many):
    assert one.keys() == many.keys()
    assert len(one) >= 9 and sum(1 for a in one.values() if np.any(a)) >= len(one) // 2
    bad = [k for k in one if not np.array_equal(bits(one[k]), bits(many[k]))]
    assert not bad, bad[:5]


def rel_l2(ref, got):
    assert ref.keys() == got.keys()
    num = sum(float(np.sum(np.abs(ref[k].astype(np.complex128) - got[k].astype(np.complex128)) ** 2)) for k in ref)
    den = sum(float(np.sum(np.abs(ref[k].astype(np.complex128)) ** 2)) for k in ref)
    assert den > 0
    return (num / den) ** 0.5


@pytest.mark.parametrize("case", ["pml", "per", "beta2d", "lorentz", "bloch", "mlatom", "noisy"])
def test_two_ranks_equal_one_rank_oracle(case):
    one, two = run_both((("oracle", case + "_presplit", "f64", 40, 1, 29611), {}), (("oracle", case, "f64", 40, 2, 29612), {}))
    assert_bit_equal(one, two)


@pytest.mark.parametrize("case", ["pml", "bloch"])
def test_four_ranks_equal_one_rank_oracle(case):
    """middle ranks have TWO remote neighbours"""
    one, four = run_both((("oracle", case + "_presplit", "f64", 30, 1, 29615), {"nslabs": 4}),
                         (("oracle", case, "f64", 30, 4, 29616), {"nslabs": 4}))
    assert_bit_equal(one, four)


@pytest.mark.parametrize("case", ["hetero_nr", "hetero_dft"])
def test_ranks_with_different_needs_oracle(case):
    """one rank owns only linear PML chunks, the other the Newton-Raphson medium / a D-component DFT: the
    halo-exchange schedule must still be the same on both (host-side lists; the library side is the gpu test below)"""
    one, two = run_both((("oracle", case + "_presplit", "f64", 30, 1, 29617), {}), (("oracle", case, "f64", 30, 2, 29618), {}))
    assert_bit_equal(one, two)


# short tiles (MEEPCU_SEG / MEEPCU_GSEG = planes per tile) so that even this small grid has boundary
# tiles AND interior tiles: the halo exchange then really runs beside the interior tiles
SHORT_TILES = {"MEEPCU_SEG": "4", "MEEPCU_GSEG": "4"}


def gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.gpu
@pytest.mark.parametrize("case,dtype,env", [("pml", "f64", None), ("per", "f64", None), ("pml", "f32", None),
                                            ("pml", "f64", SHORT_TILES), ("per", "f64", SHORT_TILES),
                                            ("pml", "f32", SHORT_TILES),
                                            ("pml", "f64", {"MEEPCU_NO_OVERLAP": "1"}),
                                            ("lorentz", "f64", None), ("lorentz", "f64", SHORT_TILES), ("bloch", "f64", None),
                                            ("bloch", "f32", SHORT_TILES), ("beta2d", "f64", None),
                                            ("hetero_nr", "f64", None), ("hetero_dft", "f64", None),
                                            ("mlatom", "f64", None), ("mlatom", "f32", SHORT_TILES), ("noisy", "f64", None)])
def test_two_gpus_equal_oracle(case, dtype, env):
    if gpus() < 2:
        pytest.skip("needs 2 GPUs")
    ref = run("oracle", case + "_presplit", dtype, 40, 1, 29613)
    two = run("cuda", case, dtype, 40, 2, 29614, env)
    assert rel_l2(ref, two) <= (1e-12 if dtype == "f64" else 1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("case,env", [("pml", None), ("pml", SHORT_TILES), ("bloch", None), ("lorentz", SHORT_TILES),
                                      ("mlatom", None), ("noisy", None)])
def test_two_gpus_equal_one_gpu_bitwise(case, env):
    """the N-GPU result must be the 1-GPU result bit for bit (same kernels, same arithmetic per point)"""
    if gpus() < 2:
        pytest.skip("needs 2 GPUs")
    one = run("cuda", case + "_presplit", "f64", 40, 1, 29619, env)
    two = run("cuda", case, "f64", 40, 2, 29620, env)
    assert_bit_equal(one, two)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [4, 8])
@pytest.mark.parametrize("case,env", [("pml", None), ("bloch", SHORT_TILES), ("lorentz", None)])
def test_many_gpus_equal_oracle(world, case, env):
    """ranks with two remote neighbours (every middle rank of the 4- and 8-GPU runs)"""
    if gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    ref = run("oracle", case + "_presplit", "f64", 30, 1, 29621, nslabs=world)
    many = run("cuda", case, "f64", 30, world, 29622, env, nslabs=world)
    assert rel_l2(ref, many) <= 1e-12
