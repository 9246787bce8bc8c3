import glob
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

ORACLE_LIB = os.path.join(ROOT, "oracle", "liboracle_meep.so")
CUDA_LIB = os.path.join(ROOT, "meep_nl_b200", "libmeepcuda.so")
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.trace.gz")))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def ref_driver(flavour):
    p = os.path.join(ROOT, "oracle", "_ref", flavour, "ref_driver")
    return p if os.path.exists(p) else None


@pytest.fixture(scope="session")
def oracle_backend():
    from meep_nl_b200 import _cabi
    if not os.path.exists(ORACLE_LIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return _cabi.Backend(ORACLE_LIB)


@pytest.fixture(scope="session")
def cuda_backend():
    from meep_nl_b200 import _cabi
    return _cabi.Backend(CUDA_LIB)  # raises if not built: no CPU fallback
