"""Build libmeepcuda.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = [os.path.join(HERE, "csrc", "meepcu.cu")]
DEPS = SOURCES + [os.path.join(HERE, "csrc", "kernels.cuh"),
                  os.path.join(HERE, "..", "include", "meepcu.h")]
OUT = os.path.join(HERE, "libmeepcuda.so")

NVCC_FLAGS = [
    "-ccbin", "/usr/bin/g++", "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    # bit-parity with the reference's x86-64 build: no FMA contraction, IEEE div/sqrt, and
    # flush-to-zero for fp32 like the FTZ/DAZ mode the reference sets (src/mympi.cpp)
    "--fmad=false", "--prec-div=true", "--prec-sqrt=true", "--ftz=true",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fopenmp", "-shared", "-lgomp",
]


def nvcc_path() -> str:
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    raise FileNotFoundError("nvcc not found")


def up_to_date() -> bool:
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SOURCES
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if verbose:
        sys.stderr.write(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
