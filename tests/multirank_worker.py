"""Worker for the world_size>1 tests: one process per rank (gloo on the CPU, NCCL on GPUs).

Every rank builds the SAME global problem with the numpy host model, owns the chunks of its
X-slab, steps, syncs, and writes EVERY array it owns (fields, PML auxiliaries, polarisations, DFT
accumulators; names carry the global chunk index) to <outdir>/rank<r>.npz; the test compares with a
single-rank run of the same pre-split problem.  backend = "oracle" (CPU, halo buffers moved by a gloo
callback) or "cuda" (libmeepcuda.so, halo buffers moved by NCCL inside the library)."""
import ctypes as C
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


def decomposition(gv, nranks, case, nslabs):
    """(split, rank_of): nslabs equal X slabs; `_presplit` single-rank runs keep the cuts and own everything"""
    if nslabs <= 1:
        return None, None
    cuts, rank_of = H.slab_decomposition(gv, nslabs, 0)
    return {0: cuts}, (rank_of if nranks > 1 else None)


def build(backend, rank, nranks, comm_id, case, rt, nslabs):
    base = case.replace("_presplit", "")
    if base == "beta2d":
        return build_beta2d(backend, rank, nranks, comm_id, case, rt, nslabs)
    if base in ("hetero_nr", "hetero_dft"):
        return build_hetero(backend, rank, nranks, comm_id, base, rt)
    n = (24 * max(nslabs, 2), 20, 18)
    gv = H.vol3d(n[0] / 10.0, n[1] / 10.0, n[2] / 10.0, 10)
    split, rank_of = decomposition(gv, nranks, case, nslabs)
    eps = lambda x, y, z: 1.0 + 1.5 * (y > 1.0)  # noqa: E731
    c = gv.center()
    if base == "pml":
        s = H.structure(gv, eps, pml=0.4, realnum=rt, split=split)
    elif base == "lorentz":  # a Lorentzian block that straddles every cut, PML all around
        inside = lambda x, y, z: ((np.abs(y - c[1]) < 0.5) & (np.abs(z - c[2]) < 0.5) & (x > 0.8) & (x < 2 * c[0] - 0.8)) * 1.0  # noqa: E731
        s = H.structure(gv, lambda x, y, z: 1.0 + 2.0 * inside(x, y, z), pml=0.4, realnum=rt, split=split)
        s.add_susceptibility(inside, H.E_stuff, H.lorentzian_susceptibility(0.35, 0.02))
    elif base in ("mlatom", "noisy"):  # a gain / noisy block that straddles every cut, PML all around
        inside = lambda x, y, z: ((np.abs(y - c[1]) < 0.5) & (np.abs(z - c[2]) < 0.5) & (x > 0.8) & (x < 2 * c[0] - 0.8)) * 1.0  # noqa: E731
        s = H.structure(gv, lambda x, y, z: 1.0 + 1.25 * inside(x, y, z), pml=0.4, realnum=rt, split=split)
        if base == "mlatom":  # three levels, two radiative transitions (tests' mlatom2d medium)
            fa, fb = 0.3, 0.22
            al = np.zeros(6)
            al[0 * 2 + 0], al[1 * 2 + 0] = -1.0 / (2 * np.pi * fa), +1.0 / (2 * np.pi * fa)
            al[1 * 2 + 1], al[2 * 2 + 1] = -1.0 / (2 * np.pi * fb), +1.0 / (2 * np.pi * fb)
            s.add_susceptibility(inside, H.E_stuff, H.multilevel_susceptibility(
                3, 2, [0.04, -0.03, 0, -0.04, 0.05, -0.01, 0, -0.02, 0.01], [0.9, 0.3, 0.1], al, [fa, fb], [0.05, 0.08],
                [1.0, 1.2, 0.8, 0, 0, 0.7, 0.9, 1.1, 0, 0]))
            s.add_susceptibility(inside, H.E_stuff, H.lorentzian_susceptibility(0.35, 0.02))
        else:
            s.add_susceptibility(inside, H.E_stuff, H.noisy_lorentzian_susceptibility(0.05, 0.35, 0.05))
    else:  # per / bloch: periodic in x (the wrap connection also crosses ranks), metallic elsewhere
        s = H.structure(gv, eps, pml=0.0, realnum=rt, split=split)
    cr = [rank_of(sc) for sc in s.chunks] if rank_of else None
    f = H.fields(s, backend, rank=rank, nranks=nranks, chunk_rank=cr, comm_id=comm_id)
    if base == "per":
        f.use_bloch(0, 0.0)
    if base == "bloch":  # complex fields, Bloch phase exp(2 pi i k L) != 1 on the wrap between the last and the first rank
        f.use_bloch(0, 0.37)
    else:
        f.use_real_fields()
    f.add_point_source(A.Ez, 0.3, 0.2, 0.0, 4.0, (c[0] - 0.35, c[1], c[2] + 0.05), 1.0)
    f.add_point_source(A.Hy, 0.25, 0.2, 0.0, 4.0, (c[0] + 0.3, c[1] + 0.05, c[2] + 0.05), 0.5)
    if base == "lorentz":
        fr = np.linspace(0.2, 0.4, 5)
        f.add_dft_fields(A.Ez, (4, 4, 4), (n[0] - 4, n[1] - 4, 9), fr)  # a monitor plane that crosses every cut
    return gv, f


def build_hetero(backend, rank, nranks, comm_id, base, rt):
    """ranks that need DIFFERENT kernels / halo exchanges (2 ranks): the cut lies on the X-PML boundary, so rank 0
    owns nothing but PML chunks.  hetero_nr: rank 1 holds the fork's Newton-Raphson chi2 medium (needs the D halo
    every step), rank 0 only linear PML.  hetero_dft: a D-component DFT region that only rank 1 holds."""
    gv = H.vol3d(3.4, 1.6, 1.4, 10)
    npml = int(0.4 * 10 + 1 + 0.5)
    split = {0: [npml]}
    rank_of = (lambda sc: 0 if sc.gv.io[0] < 2 * npml else 1) if nranks > 1 else None
    s = H.structure(gv, lambda x, y, z: 2.25 + 0 * x, pml=0.4, realnum=rt, split=split)
    if base == "hetero_nr":
        right = lambda x, y, z: 1.0 * (x > (npml + 0.75) / 10.0)  # noqa: E731
        s.set_offdiag_flags()
        s.set_chi2(lambda x, y, z: 1e-2 * right(x, y, z), elide_trivial=True)
        s.set_chi3(lambda x, y, z: right(x, y, z), elide_trivial=True)
    cr = [rank_of(sc) for sc in s.chunks] if rank_of else None
    f = H.fields(s, backend, rank=rank, nranks=nranks, chunk_rank=cr, comm_id=comm_id)
    f.use_real_fields()
    c = gv.center()
    f.add_point_source(A.Ez, 0.3, 0.2, 0.0, 4.0, (c[0] - 0.35, c[1], c[2] + 0.05), 1.0)
    f.add_point_source(A.Ex, 0.3, 0.2, 0.0, 4.0, (c[0] + 0.2, c[1] + 0.1, c[2]), 1.0)
    if base == "hetero_dft":
        f.add_dft_fields(A.Dz, (10, 3, 3), (28, 12, 10), np.linspace(0.2, 0.4, 3))
    return gv, f


def build_beta2d(backend, rank, nranks, comm_id, case, rt, nslabs):
    """2-D cell with exp(i beta z) dependence (TE and TM coupled), PML, split in X slabs"""
    gv = H.vol2d(2.4 * max(nslabs, 2), 2.0, 10)
    split, rank_of = decomposition(gv, nranks, case, nslabs)
    s = H.structure(gv, lambda x, y, z: 1.0 + 3.0 * (np.abs(y - 1.0) < 0.3), pml=0.4, realnum=rt, split=split)
    cr = [rank_of(sc) for sc in s.chunks] if rank_of else None
    f = H.fields(s, backend, rank=rank, nranks=nranks, chunk_rank=cr, comm_id=comm_id)
    f.beta = 0.3
    f.use_real_fields()
    c = gv.center()
    f.add_point_source(A.Ez, 0.3, 0.2, 0.0, 4.0, (c[0] - 0.35, c[1] + 0.1, 0.0), 1.0)
    return gv, f


def main():
    kind, case, dtype, steps, outdir = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5]
    nslabs = int(sys.argv[6]) if len(sys.argv) > 6 else 2
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
    gv, f = build(backend, rank, world, comm_id, case, rt, nslabs)
    f.device = device
    f.init()
    cb = None
    if kind == "oracle" and world > 1:
        cb = make_transport()
        backend.lib.oracle_set_transport.argtypes = [C.c_void_p, TRANSPORT, C.c_void_p]
        backend.lib.oracle_set_transport.restype = C.c_int
        backend.check(backend.lib.oracle_set_transport(f.ctx, cb, None))
    # several calls: every call boundary (and the lazily exchanged halos at the final sync) is part of the test
    done = 0
    for nb in (steps // 3, steps // 3, steps - 2 * (steps // 3)):
        if nb:
            f.step(nb)
            done += nb
    f.sync_to_host()
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), **dict(f.owned_arrays()))
    dist.barrier()
    f.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
