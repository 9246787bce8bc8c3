"""ctypes binding of the C ABI declared in include/meepcu.h.

The structures below mirror that header field by field; `Backend` wraps one loaded shared
library.  The product library is meep_nl_b200/libmeepcuda.so (CUDA, sm_100a).  The same ABI is
exported by oracle/liboracle_meep.so, which ONLY tests / smoke / bench's CPU baseline may load
(they pass its path explicitly; nothing in this package falls back to it).
"""
from __future__ import annotations

import ctypes as C
import os

NUM_COMPONENTS = 20
NUM_FIELD_TYPES = 8
ABI_VERSION = 2
SYNC_FIELDS, SYNC_AUX, SYNC_POL, SYNC_DFT, SYNC_MATERIALS, SYNC_LOCAL = 1, 2, 4, 8, 16, 32

# component ids (reference src/meep/vec.hpp:31-52)
Ex, Ey, Er, Ep, Ez, Hx, Hy, Hr, Hp, Hz, Dx, Dy, Dr, Dp, Dz, Bx, By, Br, Bp, Bz = range(20)
D1, D2, D3, Dcyl = range(4)
X, Y, Z, R, P, NO_DIRECTION = range(6)
E_stuff, H_stuff, D_stuff, B_stuff, PE_stuff, PH_stuff, WE_stuff, WH_stuff = range(8)
CONNECT_PHASE, CONNECT_NEGATE, CONNECT_COPY = range(3)
NL_FORK_LITERAL, NL_UPSTREAM_PADE = 0, 1

OK, EINVAL, EUNSUPPORTED, ECUDA, ENOMEM, ESTATE, ENCCL, ENONFINITE = 0, -1, -2, -3, -4, -5, -6, -7

vp = C.c_void_p


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int), ("realnum_bytes", C.c_int), ("is_real", C.c_int),
                ("dim", C.c_int), ("a", C.c_double), ("dt", C.c_double), ("courant", C.c_double),
                ("beta", C.c_double), ("nl_mode", C.c_int), ("device", C.c_int), ("t", C.c_int),
                ("pin_host", C.c_int), ("m", C.c_double), ("zero_fields_near_cylorigin", C.c_int),
                ("reserved", C.c_int * 3),
                ("bfast_scaled_k", C.c_double * 3)]


class ChunkDesc(C.Structure):
    _fields_ = [("num", C.c_int * 3), ("io", C.c_int * 3),
                ("f", (vp * 2) * NUM_COMPONENTS), ("f_u", (vp * 2) * NUM_COMPONENTS),
                ("f_w", (vp * 2) * NUM_COMPONENTS), ("f_cond", (vp * 2) * NUM_COMPONENTS),
                ("chi1inv", (vp * 5) * NUM_COMPONENTS), ("chi2", vp * NUM_COMPONENTS),
                ("chi3", vp * NUM_COMPONENTS), ("conductivity", (vp * 5) * NUM_COMPONENTS),
                ("condinv", (vp * 5) * NUM_COMPONENTS),
                ("sig", vp * 6), ("kap", vp * 6), ("siginv", vp * 6), ("sigsize", C.c_int * 6),
                ("rank", C.c_int), ("reserved", C.c_int * 7), ("f_bfast", (vp * 2) * NUM_COMPONENTS)]


class SusceptibilityDesc(C.Structure):
    _fields_ = [("ft", C.c_int), ("omega_0", C.c_double), ("gamma", C.c_double),
                ("no_omega_0_denominator", C.c_int),
                ("sigma", (vp * 5) * NUM_COMPONENTS), ("P", (vp * 2) * NUM_COMPONENTS),
                ("P_prev", (vp * 2) * NUM_COMPONENTS), ("noise_amp", C.c_double), ("reserved", C.c_int * 6)]


class MultilevelDesc(C.Structure):  # meepcu_multilevel_desc
    _fields_ = [("ft", C.c_int), ("L", C.c_int), ("T", C.c_int),
                ("Gamma", vp), ("GammaInv", vp), ("alpha", vp), ("omega", vp), ("gamma", vp), ("sigmat", vp),
                ("sigma", (vp * 5) * NUM_COMPONENTS), ("P", (vp * 2) * NUM_COMPONENTS),
                ("P_prev", (vp * 2) * NUM_COMPONENTS), ("N", vp), ("W_prev", (vp * 2) * NUM_COMPONENTS),
                ("reserved", C.c_int * 8)]


class GyrotropicDesc(C.Structure):
    _fields_ = [("ft", C.c_int), ("model", C.c_int), ("omega_0", C.c_double), ("gamma", C.c_double),
                ("alpha", C.c_double), ("gyro_tensor", (C.c_double * 3) * 3), ("sigma", (vp * 5) * NUM_COMPONENTS),
                ("P", ((vp * 3) * 2) * NUM_COMPONENTS), ("P_prev", ((vp * 3) * 2) * NUM_COMPONENTS),
                ("reserved", C.c_int * 8)]


class SourceDesc(C.Structure):
    _fields_ = [("ft", C.c_int), ("c", C.c_int), ("src_time_id", C.c_int),
                ("is_integrated", C.c_int), ("n", C.c_size_t), ("index", vp), ("amp", vp)]


class DftDesc(C.Structure):
    _fields_ = [("c", C.c_int), ("is_", C.c_int * 3), ("ie", C.c_int * 3),
                ("s0", C.c_double * 3), ("s1", C.c_double * 3), ("e0", C.c_double * 3),
                ("e1", C.c_double * 3), ("dV0", C.c_double), ("dV1", C.c_double),
                ("avg1", C.c_ssize_t), ("avg2", C.c_ssize_t),
                ("include_dV_and_interp_weights", C.c_int),
                ("sqrt_dV_and_interp_weights", C.c_int), ("decimation_factor", C.c_int),
                ("Nomega", C.c_int), ("omega", vp), ("scale", C.c_double * 2), ("dft", vp),
                ("reserved", C.c_int * 8)]


class IntegrateDesc(C.Structure):  # meepcu_integrate_desc
    _fields_ = [("f", (vp * 2) * 2), ("off", (C.c_ssize_t * 2) * 2), ("ph", (C.c_double * 2) * 2),
                ("is_", C.c_int * 3), ("ie", C.c_int * 3),
                ("s0", C.c_double * 3), ("s1", C.c_double * 3), ("e0", C.c_double * 3), ("e1", C.c_double * 3),
                ("dV0", C.c_double), ("dV1", C.c_double), ("reserved", C.c_int * 4)]


class SrcSample(C.Structure):
    _fields_ = [("current_B", C.c_double * 2), ("dipole_H", C.c_double * 2),
                ("current_D", C.c_double * 2), ("dipole_E", C.c_double * 2)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("steps", C.c_int64), ("device_bytes", C.c_int64),
                ("cells", C.c_int64), ("algorithmic_bytes_per_step", C.c_int64),
                ("last_step_ms", C.c_double), ("reserved", C.c_int64 * 8)]


class KernelTime(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("launches", C.c_int64), ("ms", C.c_double),
                ("algorithmic_bytes", C.c_int64)]


# every symbol include/meepcu.h declares: (name, restype, argtypes)
_PROTOTYPES = [
    ("meepcu_last_error", C.c_char_p, []),
    ("meepcu_backend", C.c_char_p, []),
    ("meepcu_abi_version", C.c_int, []),
    ("meepcu_create", C.c_int, [C.POINTER(Config), C.POINTER(vp)]),
    ("meepcu_destroy", C.c_int, [vp]),
    ("meepcu_add_chunk", C.c_int, [vp, C.POINTER(ChunkDesc), C.POINTER(C.c_int)]),
    ("meepcu_add_susceptibility", C.c_int, [vp, C.c_int, C.POINTER(SusceptibilityDesc)]),
    ("meepcu_add_gyrotropic", C.c_int, [vp, C.c_int, C.POINTER(GyrotropicDesc)]),
    ("meepcu_add_multilevel", C.c_int, [vp, C.c_int, C.POINTER(MultilevelDesc)]),
    ("meepcu_get_noise_count", C.c_int, [vp, C.POINTER(C.c_size_t)]),
    ("meepcu_set_noise", C.c_int, [vp, C.c_int, C.c_size_t, vp]),
    ("meepcu_add_source", C.c_int, [vp, C.c_int, C.POINTER(SourceDesc)]),
    ("meepcu_add_dft", C.c_int, [vp, C.c_int, C.POINTER(DftDesc), C.POINTER(C.c_int)]),
    ("meepcu_set_metal", C.c_int, [vp, C.c_int, C.c_size_t, vp]),
    ("meepcu_add_connections", C.c_int, [vp, C.c_int, C.c_int, C.c_size_t, vp, vp, vp]),
    ("meepcu_add_remote_send", C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_size_t, vp]),
    ("meepcu_add_remote_recv", C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_size_t, vp, vp]),
    ("meepcu_comm_unique_id", C.c_int, [vp]),
    ("meepcu_comm_init", C.c_int, [vp, C.c_int, C.c_int, vp]),
    ("meepcu_finalize", C.c_int, [vp]),
    ("meepcu_step", C.c_int, [vp, C.c_int, C.c_int, vp]),
    ("meepcu_synchronize_magnetic_fields", C.c_int, [vp, C.c_int, vp]),
    ("meepcu_restore_magnetic_fields", C.c_int, [vp]),
    ("meepcu_sync_to_host", C.c_int, [vp]),
    ("meepcu_sync_from_host", C.c_int, [vp]),
    ("meepcu_sync_subset", C.c_int, [vp, C.c_int, C.c_uint]),
    ("meepcu_get_time", C.c_int, [vp, C.POINTER(C.c_int)]),
    ("meepcu_set_time", C.c_int, [vp, C.c_int]),
    ("meepcu_check_finite", C.c_int, [vp, C.POINTER(C.c_int)]),
    ("meepcu_read_value", C.c_int, [vp, vp, C.POINTER(C.c_double)]),
    ("meepcu_read_values", C.c_int, [vp, C.c_size_t, vp, C.POINTER(C.c_double)]),
    ("meepcu_dft_dot", C.c_int, [vp, C.c_size_t, vp, vp, C.c_int, C.POINTER(C.c_double)]),
    ("meepcu_integrate_dot", C.c_int, [vp, C.c_size_t, C.POINTER(IntegrateDesc), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    ("meepcu_get_timers", C.c_int, [vp, C.POINTER(C.c_double)]),
    ("meepcu_get_stats", C.c_int, [vp, C.POINTER(Stats)]),
    ("meepcu_set_profiling", C.c_int, [vp, C.c_int]),
    ("meepcu_get_kernel_times", C.c_int, [vp, C.POINTER(KernelTime), C.c_int, C.POINTER(C.c_int)]),
    ("meepcu_get_compulsory_bytes", C.c_int, [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int, C.POINTER(C.c_int)]),
]
SYMBOLS = [p[0] for p in _PROTOTYPES]


class MeepcuError(RuntimeError):
    """Raised for every non-zero return code (the reference would meep::abort, mympi.cpp:244-263)."""

    def __init__(self, code, msg):
        super().__init__("meepcu error %d: %s" % (code, msg))
        self.code = code


def default_library_path() -> str:
    """the in-tree build, or another build of the same ABI named by MEEPCU_LIBRARY (kernel experiments)"""
    return os.environ.get("MEEPCU_LIBRARY") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmeepcuda.so")


class Backend:
    """One loaded implementation of the ABI."""

    def __init__(self, path: str | None = None):
        self.path = path or default_library_path()
        if not os.path.exists(self.path):
            raise FileNotFoundError(
                "%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % self.path)
        self.lib = C.CDLL(self.path, mode=C.RTLD_LOCAL)
        for name, res, args in _PROTOTYPES:
            fn = getattr(self.lib, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        if self.lib.meepcu_abi_version() != ABI_VERSION:
            raise RuntimeError("ABI version mismatch in %s" % self.path)
        # the default backend (and any MEEPCU_LIBRARY override of it) is the CUDA library: the CPU oracle
        # is test infrastructure and is only ever loaded by an explicit path from tests/, smoke() or bench's CPU arm
        if path is None and not self.lib.meepcu_backend().decode().startswith("cuda"):
            raise RuntimeError("%s is not a CUDA build of the ABI (%s): there is no CPU fallback"
                               % (self.path, self.lib.meepcu_backend().decode()))

    @property
    def name(self) -> str:
        return self.lib.meepcu_backend().decode()

    def check(self, rc: int):
        if rc != OK:
            raise MeepcuError(rc, self.lib.meepcu_last_error().decode())

    def __getattr__(self, item):
        # b.create(...) -> lib.meepcu_create(...)
        return getattr(self.lib, "meepcu_" + item)
