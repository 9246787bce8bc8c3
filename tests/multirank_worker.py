"""Worker for the world_size>1 tests: one process per rank (gloo on the CPU, NCCL on GPUs).

Every rank builds the SAME global problem with the numpy host model, owns the chunks of its
X-slab, steps, and writes the values of a probe line it owns; rank 0 gathers and compares with a
single-rank run of the same problem.  backend = "oracle" (CPU, halo buffers moved by a gloo
callback) or "cuda" (libmeepcuda.so, halo buffers moved by NCCL inside the library)."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from meep_nl_b200 import _cabi as A  # noqa: E402
from meep_nl_b200 import host as H  # noqa: E402

TRANSPORT = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p),
                        C.POINTER(C.c_longlong), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p),
                        C.POINTER(C.c_longlong))


def make_transport():
    def cb(user, ft, nsend, speers, sbufs, sbytes, nrecv, rpeers, rbufs, rbytes):
        reqs = []
        keep = []
        for k in range(nrecv):
            arr = np.ctypeslib.as_array(C.cast(rbufs[k], C.POINTER(C.c_uint8)), shape=(rbytes[k],))
            t = torch.from_numpy(arr)
            keep.append(t)
            reqs.append(dist.irecv(t, src=rpeers[k], tag=ft))
        for k in range(nsend):
            arr = np.ctypeslib.as_array(C.cast(sbufs[k], C.POINTER(C.c_uint8)), shape=(sbytes[k],))
            t = torch.from_numpy(arr.copy())
            keep.append(t)
            reqs.append(dist.isend(t, dst=speers[k], tag=ft))
        for r in reqs:
            r.wait()
    return TRANSPORT(cb)


def build(backend, rank, nranks, comm_id, case, rt):
    n = (24 * max(nranks, 2), 20, 18)
    if case.startswith("beta2d"):
        return build_beta2d(backend, rank, nranks, comm_id, case, rt)
    gv = H.vol3d(n[0] / 10.0, n[1] / 10.0, n[2] / 10.0, 10)
    split, rank_of = None, None
    if nranks > 1:
        cuts, rank_of = H.slab_decomposition(gv, nranks, 0)
        split = {0: cuts}
    elif case.endswith("_presplit"):
        cuts, _ = H.slab_decomposition(gv, 2, 0)
        split = {0: cuts}
    if case.startswith("pml"):
        s = H.structure(gv, lambda x, y, z: 1.0 + 1.5 * (y > 1.0), pml=0.4, realnum=rt, split=split)
    else:  # periodic in x (the wrap connection also crosses ranks), metallic elsewhere
        s = H.structure(gv, lambda x, y, z: 1.0 + 1.5 * (y > 1.0), pml=0.0, realnum=rt, split=split)
    cr = [rank_of(sc) for sc in s.chunks] if rank_of else None
    f = H.fields(s, backend, rank=rank, nranks=nranks, chunk_rank=cr, comm_id=comm_id)
    if case.startswith("per"):
        f.use_bloch(0, 0.0)
    f.use_real_fields()
    c = gv.center()
    f.add_point_source(A.Ez, 0.3, 0.2, 0.0, 4.0, (c[0] - 0.35, c[1], c[2] + 0.05), 1.0)
    f.add_point_source(A.Hy, 0.25, 0.2, 0.0, 4.0, (c[0] + 0.3, c[1] + 0.05, c[2] + 0.05), 0.5)
    return gv, f


def build_beta2d(backend, rank, nranks, comm_id, case, rt):
    """2-D cell with exp(i beta z) dependence (TE and TM coupled), PML, split in X slabs"""
    gv = H.vol2d(2.4 * max(nranks, 2), 2.0, 10)
    split, rank_of = None, None
    if nranks > 1:
        cuts, rank_of = H.slab_decomposition(gv, nranks, 0)
        split = {0: cuts}
    elif case.endswith("_presplit"):
        cuts, _ = H.slab_decomposition(gv, 2, 0)
        split = {0: cuts}
    s = H.structure(gv, lambda x, y, z: 1.0 + 3.0 * (np.abs(y - 1.0) < 0.3), pml=0.4, realnum=rt, split=split)
    cr = [rank_of(sc) for sc in s.chunks] if rank_of else None
    f = H.fields(s, backend, rank=rank, nranks=nranks, chunk_rank=cr, comm_id=comm_id)
    f.beta = 0.3
    f.use_real_fields()
    c = gv.center()
    f.add_point_source(A.Ez, 0.3, 0.2, 0.0, 4.0, (c[0] - 0.35, c[1] + 0.1, 0.0), 1.0)
    return gv, f


def probe(gv, f):
    c = gv.center()
    out = {}
    for k in range(1, gv.num[0]):
        for comp, off in ((A.Ez, (0.0, 0.0, 0.05)), (A.Hy, (0.05, 0.0, 0.05))):
            v = f.get_field(comp, (k * 0.1 + off[0], c[1] + off[1], c[2] + off[2]))
            if v is not None:
                out["%d_%d" % (comp, k)] = v
    return out


def main():
    kind, case, dtype, steps, outdir = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    rt = np.float64 if dtype == "f64" else np.float32
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if kind == "cuda":
        local = int(os.environ.get("LOCAL_RANK", rank))
        backend = A.Backend()
        device = local
    else:
        backend = A.Backend(os.path.join(ROOT, "oracle", "liboracle_meep.so"))
        device = 0
    idbuf = (C.c_char * 128)()
    if rank == 0:
        backend.check(backend.comm_unique_id(idbuf))
    t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8)
    dist.broadcast(t, 0)
    comm_id = bytes(t.tolist())
    gv, f = build(backend, rank, world, comm_id, case, rt)
    f.device = device
    f.init()
    cb = None
    if kind == "oracle" and world > 1:
        cb = make_transport()
        backend.lib.oracle_set_transport.argtypes = [C.c_void_p, TRANSPORT, C.c_void_p]
        backend.lib.oracle_set_transport.restype = C.c_int
        backend.check(backend.lib.oracle_set_transport(f.ctx, cb, None))
    f.step(steps)
    f.sync_to_host()
    mine = probe(gv, f)
    json.dump(mine, open(os.path.join(outdir, "rank%d.json" % rank), "w"))
    dist.barrier()
    f.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
