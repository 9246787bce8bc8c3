"""meep_nl_b200 -- B200 (sm_100a) implementation of Meep's per-timestep field update
(PMack10/meep_nl fork) behind the C ABI of include/meepcu.h.

    csrc/            hand-written CUDA kernels + the extern "C" layer  -> libmeepcuda.so
    shim/            reference-side C++ binding (what a Meep maintainer adds to libmeep)
    _cabi.py         ctypes mirror of include/meepcu.h
    host.py          numpy mirror of the reference's grid_volume/structure/fields setup API
    build.py         nvcc driver (sm_100a only)

There is no CPU fallback: loading fails loudly if libmeepcuda.so is not built, and stepping fails
with MEEPCU_ECUDA if no CUDA device is present.
"""
from ._cabi import Backend, MeepcuError, default_library_path  # noqa: F401

__all__ = ["Backend", "MeepcuError", "default_library_path", "load"]

_backend = None


def load():
    """The product backend (libmeepcuda.so); raises if it has not been built."""
    global _backend
    if _backend is None:
        _backend = Backend(default_library_path())
    return _backend
