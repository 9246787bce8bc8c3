"""TEST INFRASTRUCTURE: runs bench.py's own N-rank code path (slab decomposition, parity gate, timed batches,
e2e leg, JSON line) on the CPU -- gloo instead of NCCL, the oracle instead of libmeepcuda.so, the oracle's
transport callback instead of NCCL send/recv.  Only the host-side logic is checked this way; the numbers mean
nothing.  Launched by tests/test_bench_contract.py under torch.distributed.run."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402
from meep_nl_b200 import _cabi  # noqa: E402
from multirank_worker import TRANSPORT, make_transport  # noqa: E402

ORACLE = os.path.join(ROOT, "oracle", "liboracle_meep.so")
_Backend = _cabi.Backend
_cabi.Backend = lambda path=None: _Backend(ORACLE)
bench.PG_BACKEND, bench.DEV = "gloo", "cpu"
_keep = []


def after_init(f, world):
    if world > 1:
        cb = make_transport()
        _keep.append(cb)
        f.b.lib.oracle_set_transport.argtypes = [C.c_void_p, TRANSPORT, C.c_void_p]
        f.b.lib.oracle_set_transport.restype = C.c_int
        f.b.check(f.b.lib.oracle_set_transport(f.ctx, cb, None))


bench.AFTER_INIT = after_init

if __name__ == "__main__":
    bench.main_ours(bench.parse())
