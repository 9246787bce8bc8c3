"""The oracle (oracle/oracle_meep.cpp) must reproduce, bit for bit, what the UNMODIFIED reference
computed for every golden trace (tests/golden/*.trace.gz, recorded by oracle/ref_driver.cpp).
This is the check that pins the oracle on machines without the reference."""
import os

import numpy as np
import pytest

import trace as tr
from conftest import GOLDEN


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-9] for p in GOLDEN])
def test_oracle_reproduces_reference_bit_for_bit(path, oracle_backend):
    t = tr.Trace(path)
    nl = 1 if "pade" in path else None
    r = tr.Replay(t, oracle_backend, nl_mode=nl)
    n = r.run()
    assert n == t.total_steps and n > 0
    r.sync_to_host()
    rel, bad, wid = r.compare_expected()
    r.close()
    assert bad == 0 and rel == 0.0, "array %d differs (rel_l2 %.3g)" % (wid, rel)


def test_traces_are_not_trivial():
    """every golden trace ends with non-zero fields (the comparison is not 0 == 0)"""
    for path in GOLDEN:
        t = tr.Trace(path)
        energy = sum(float(np.sum(np.abs(v.astype(np.complex128)) ** 2)) for v in t.expect.values())
        assert energy > 0, path


def test_split_run_equals_single_run(oracle_backend):
    """meepcu_step(n) == n x meepcu_step(1): the step counter and DFT phases advance correctly"""
    path = [p for p in GOLDEN if "lorentz3d_16_fp64" in p][0]
    t = tr.Trace(path)
    r = tr.Replay(t, oracle_backend)
    done = 0
    for ns, nsrc, raw, _noise in t.steps:
        import ctypes as C
        from meep_nl_b200 import _cabi as A
        sz = C.sizeof(A.SrcSample) * nsrc
        for s in range(ns):
            chunk = raw[s * sz:(s + 1) * sz]
            buf = C.create_string_buffer(chunk, len(chunk)) if chunk else None
            r.b.check(r.b.step(r.ctx, 1, nsrc, buf))
            done += 1
    rel, bad, _ = r.compare_expected()
    r.close()
    assert done == t.total_steps and bad == 0
