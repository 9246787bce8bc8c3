"""-m gpu: the CUDA path (libmeepcuda.so, through the C ABI) against the UNMODIFIED reference's
results (golden traces) and against the oracle on the same inputs.

Tolerances are the ones BASELINE.json's north_star states: max relative L2 error <= 1e-12 in
double, <= 1e-5 in the single-precision realnum build; DFT accumulators are among the compared
arrays.  (Observed: bit-identical wherever no value underflows; the only differences are
denormal-range values, which the reference's x86 FTZ/DAZ mode flushes and fp64 CUDA cannot.)"""
import ctypes as C
import os

import numpy as np
import pytest

import trace as tr
from conftest import GOLDEN
from meep_nl_b200 import _cabi as A

pytestmark = pytest.mark.gpu


def tol_for(trace):
    return 1e-12 if trace.rb == 8 else 1e-5


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-9] for p in GOLDEN])
def test_cuda_matches_reference(path, cuda_backend):
    t = tr.Trace(path)
    r = tr.Replay(t, cuda_backend, nl_mode=1 if "pade" in path else None)
    n = r.run()
    assert n == t.total_steps
    r.sync_to_host()
    rel, bad, wid = r.compare_expected()
    st = r.stats()
    r.close()
    assert st.kernel_launches > 0
    assert rel <= tol_for(t), "rel_l2 %.3g (first differing array %d, %d elements)" % (rel, wid, bad)


@pytest.mark.parametrize("name", ["kr_metal3d_fp64", "kr_per3d_rods_fp64", "nr3d_10_fp64", "kr_polariton1d_fp64",
                                  "kr_per2d_rods_fp32", "phc3d_10_fp64"])
def test_cuda_bit_exact_where_nothing_underflows(name, cuda_backend):
    """cases whose fields stay far from the denormal range must agree to the last bit"""
    path = [p for p in GOLDEN if os.path.basename(p).startswith(name)][0]
    t = tr.Trace(path)
    r = tr.Replay(t, cuda_backend)
    r.run()
    r.sync_to_host()
    rel, bad, wid = r.compare_expected()
    r.close()
    assert bad == 0, "array %d: %d elements differ (rel_l2 %.3g)" % (wid, bad, rel)


def test_cuda_equals_oracle_mid_run(cuda_backend, oracle_backend):
    """same inputs, same number of steps (not the recorded total): CUDA vs the CPU oracle"""
    path = [p for p in GOLDEN if "vac3d_cplx_14_fp64" in p][0]
    t = tr.Trace(path)
    rc, ro = tr.Replay(t, cuda_backend), tr.Replay(t, oracle_backend)
    rc.run(17)
    ro.run(17)
    rc.sync_to_host()
    num = den = 0.0
    for aid in t.expect:
        d = rc.arr[aid].astype(np.float64) - ro.arr[aid].astype(np.float64)
        num += float(np.sum(d * d))
        den += float(np.sum(ro.arr[aid].astype(np.float64) ** 2))
    rc.close()
    ro.close()
    assert den > 0 and (num / den) ** 0.5 <= 1e-12


def test_step_batching_is_invisible(cuda_backend):
    """meepcu_step(n) == n x meepcu_step(1), including DFT phases and decimation bookkeeping"""
    path = [p for p in GOLDEN if "lorentz3d_16_fp64" in p][0]
    t = tr.Trace(path)
    r = tr.Replay(t, cuda_backend)
    for ns, nsrc, raw, _noise in t.steps:
        sz = C.sizeof(A.SrcSample) * nsrc
        for s in range(ns):
            chunk = raw[s * sz:(s + 1) * sz]
            buf = C.create_string_buffer(chunk, len(chunk)) if chunk else None
            r.b.check(r.b.step(r.ctx, 1, nsrc, buf))
    r.sync_to_host()
    rel, bad, _ = r.compare_expected()
    r.close()
    assert rel <= 1e-12


def test_sync_roundtrip_is_lossless(cuda_backend):
    """sync_to_host followed by sync_from_host must not change the trajectory"""
    path = [p for p in GOLDEN if "kr_xper_ypml3d_fp64" in p][0]
    t = tr.Trace(path)
    r = tr.Replay(t, cuda_backend)
    ns, nsrc, raw, _noise = t.steps[0]
    sz = C.sizeof(A.SrcSample) * nsrc
    half = ns // 2
    a = raw[:half * sz]
    b = raw[half * sz:]
    r.b.check(r.b.step(r.ctx, half, nsrc, C.create_string_buffer(a, len(a)) if a else None))
    r.sync_to_host()
    r.b.check(r.b.sync_from_host(r.ctx))
    r.b.check(r.b.step(r.ctx, ns - half, nsrc, C.create_string_buffer(b, len(b)) if b else None))
    r.sync_to_host()
    rel, bad, _ = r.compare_expected()
    r.close()
    assert rel <= 1e-12


def test_check_finite_and_read_value(cuda_backend):
    path = [p for p in GOLDEN if "kr_metal3d_fp64" in p][0]
    t = tr.Trace(path)
    r = tr.Replay(t, cuda_backend)
    r.run()
    ok = C.c_int(0)
    r.b.check(r.b.check_finite(r.ctx, C.byref(ok)))
    assert ok.value == 1
    # pick the largest expected element and read it straight from the device mirror
    aid = max(t.expect, key=lambda k: float(np.max(np.abs(t.expect[k]))))
    i = int(np.argmax(np.abs(t.expect[aid])))
    out = C.c_double()
    r.b.check(r.b.read_value(r.ctx, r.arr[aid].ctypes.data + i * r.arr[aid].itemsize, C.byref(out)))
    r.close()
    assert out.value == float(t.expect[aid][i])
