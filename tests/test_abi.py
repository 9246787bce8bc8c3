"""Host-side checks that run without a GPU: the C-ABI libraries load, export every symbol that
include/meepcu.h declares, and the product fails loudly when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

from conftest import CUDA_LIB, ORACLE_LIB, ROOT
from meep_nl_b200 import _cabi as A


def header_symbols():
    src = open(os.path.join(ROOT, "include", "meepcu.h")).read()
    return sorted(set(re.findall(r"\b(meepcu_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(A.SYMBOLS)


@pytest.mark.parametrize("path", [CUDA_LIB, ORACLE_LIB])
def test_library_exports_every_symbol(path):
    lib = C.CDLL(path)
    for name in header_symbols():
        assert hasattr(lib, name), "%s lacks %s" % (path, name)
    b = A.Backend(path)
    assert b.lib.meepcu_abi_version() == A.ABI_VERSION


def test_struct_sizes_match_c():
    # sizes computed by the C compiler for include/meepcu.h (x86-64 SysV)
    assert C.sizeof(A.Config) == 112  # ABI 2: + bfast_scaled_k[3]
    assert C.sizeof(A.SrcSample) == 64
    assert C.sizeof(A.SourceDesc) == 40
    assert C.sizeof(A.KernelTime) == 72
    assert C.sizeof(A.ChunkDesc) == 24 + 8 * (4 * 40 + 100 + 20 + 20 + 100 + 100 + 18) + 24 + 32 + 8 * 40 == 4544  # ABI 2: + f_bfast


def test_bad_arguments_are_reported(oracle_backend, cuda_backend):
    for b in (oracle_backend, cuda_backend):
        cfg = A.Config()
        cfg.abi_version = 999
        ctx = A.vp()
        assert b.create(C.byref(cfg), C.byref(ctx)) == A.EINVAL
        assert b"ABI" in b.lib.meepcu_last_error()
        cfg.abi_version = A.ABI_VERSION
        cfg.realnum_bytes = 8
        cfg.dim = 7  # not a grid_volume::dim (D1, D2, D3, Dcyl are all accepted)
        assert b.create(C.byref(cfg), C.byref(ctx)) == A.EINVAL
        cfg.dim = A.D3
        cfg.realnum_bytes = 2
        assert b.create(C.byref(cfg), C.byref(ctx)) == A.EINVAL


def test_cuda_library_has_no_cpu_fallback(cuda_backend):
    """Without a CUDA device meepcu_finalize must fail with MEEPCU_ECUDA, never compute on the CPU."""
    import subprocess, sys
    code = (
        "import sys, ctypes as C; sys.path.insert(0, %r)\n"
        "from meep_nl_b200 import _cabi as A\n"
        "b = A.Backend(%r)\n"
        "cfg = A.Config(); cfg.abi_version = A.ABI_VERSION; cfg.realnum_bytes = 8; cfg.dim = A.D3; cfg.a = 10; cfg.dt = .05; cfg.courant = .5\n"
        "ctx = A.vp(); assert b.create(C.byref(cfg), C.byref(ctx)) == 0\n"
        "rc = b.finalize(ctx); print(rc, b.lib.meepcu_last_error().decode())\n" % (ROOT, CUDA_LIB))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert out.returncode == 0, out.stderr
    assert out.stdout.startswith("%d " % A.ECUDA), out.stdout
    assert "no CPU fallback" in out.stdout or "failed" in out.stdout


def test_product_package_does_not_reference_the_oracle():
    """The product path must not import / load anything under oracle/."""
    pkg = os.path.join(ROOT, "meep_nl_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if not fn.endswith((".py", ".cu", ".cuh", ".hpp", ".h")) or "shim" in dirpath:
                continue
            txt = open(os.path.join(dirpath, fn)).read()
            assert "liboracle" not in txt.replace("oracle/liboracle_meep.so, which ONLY", ""), fn
            assert "import oracle" not in txt and "from oracle" not in txt, fn


def test_header_is_plain_c_and_every_symbol_has_c_linkage(tmp_path):
    """include/meepcu.h compiled by a C99 compiler (-pedantic -Werror); the C client dlopen()s libmeepcuda.so and
    the oracle and finds all 38 entry points; struct sizes agree with the ctypes mirror"""
    import json, subprocess
    exe = str(tmp_path / "abi_c_client")
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "abi_c_client.c"), "-o", exe, "-ldl"], check=True)
    for lib in (CUDA_LIB, ORACLE_LIB):
        r = subprocess.run([exe, lib], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        d = json.loads(r.stdout)
        assert d["missing"] == 0 and d["symbols"] == len(A.SYMBOLS) == 38
        assert d["abi"] == d["header_abi"] == A.ABI_VERSION
        assert d["sizeof_config"] == C.sizeof(A.Config) and d["sizeof_chunk_desc"] == C.sizeof(A.ChunkDesc)
        assert d["sizeof_gyrotropic_desc"] == C.sizeof(A.GyrotropicDesc)
        # noise_amp took the place of two reserved ints: the size of the descriptor did not change
        assert d["sizeof_susceptibility_desc"] == C.sizeof(A.SusceptibilityDesc) == 24 + 8 + 8 * (100 + 40 + 40) + 32
        assert d["sizeof_multilevel_desc"] == C.sizeof(A.MultilevelDesc)


def test_new_media_arguments_are_checked_by_both_libraries(oracle_backend, cuda_backend):
    """meepcu_add_multilevel / meepcu_set_noise / meepcu_get_noise_count: the refusals (magnetic multilevel media, level /
    transition counts, missing arrays, a wrong number of noise values) are the same in the CUDA library and the oracle;
    none of them needs a device"""
    import numpy as np
    for b in (oracle_backend, cuda_backend):
        cfg = A.Config()
        cfg.abi_version, cfg.realnum_bytes, cfg.dim, cfg.is_real = A.ABI_VERSION, 8, A.D3, 1
        cfg.a, cfg.dt, cfg.courant = 10, 0.05, 0.5
        ctx = A.vp()
        b.check(b.create(C.byref(cfg), C.byref(ctx)))
        n = 5
        ntot = (n + 1) ** 3
        arr = {k: np.zeros(ntot) for k in ("ex", "dx", "P", "Pp", "sig")}
        cd = A.ChunkDesc()
        for k in range(3):
            cd.num[k] = n
        cd.f[A.Ex][0] = arr["ex"].ctypes.data
        cd.f[A.Dx][0] = arr["dx"].ctypes.data
        cid = C.c_int(-1)
        b.check(b.add_chunk(ctx, C.byref(cd), C.byref(cid)))
        small = np.ones(4)
        N = np.zeros(2 * ntot)
        md = A.MultilevelDesc()
        md.ft, md.L, md.T = A.H_stuff, 2, 1
        for name in ("Gamma", "GammaInv", "alpha", "omega", "gamma", "sigmat"):
            setattr(md, name, small.ctypes.data)
        md.N = N.ctypes.data
        assert b.add_multilevel(ctx, 0, C.byref(md)) == A.EUNSUPPORTED  # magnetic multilevel media
        md.ft, md.L = A.E_stuff, 9
        assert b.add_multilevel(ctx, 0, C.byref(md)) == A.EUNSUPPORTED  # more levels than the kernels hold
        md.L, md.N = 2, None
        assert b.add_multilevel(ctx, 0, C.byref(md)) == A.EINVAL        # the populations are required
        assert b.add_multilevel(ctx, 7, C.byref(md)) == A.EINVAL        # no such chunk
        # a noisy Lorentzian on Ex: one unit Gaussian per owned point and step
        sd = A.SusceptibilityDesc()
        sd.ft, sd.omega_0, sd.gamma, sd.noise_amp = A.E_stuff, 1.1, 0.05, 0.1
        sd.sigma[A.Ex][0] = arr["sig"].ctypes.data
        sd.P[A.Ex][0], sd.P_prev[A.Ex][0] = arr["P"].ctypes.data, arr["Pp"].ctypes.data
        b.check(b.add_susceptibility(ctx, 0, C.byref(sd)))
        cnt = C.c_size_t(0)
        b.check(b.get_noise_count(ctx, C.byref(cnt)))
        assert cnt.value == n ** 3  # LOOP_OVER_VOL_OWNED: num points per direction
        g = np.zeros(2 * cnt.value)
        assert b.set_noise(ctx, 2, cnt.value + 1, g.ctypes.data) == A.EINVAL
        assert b"values per step" in b.lib.meepcu_last_error()
        assert b.set_noise(ctx, 2, cnt.value, None) == A.EINVAL
        b.check(b.set_noise(ctx, 2, cnt.value, g.ctypes.data))
        b.destroy(ctx)
