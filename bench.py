#!/usr/bin/env python
"""bench.py -- Mcell-updates/s of the fused E+H time step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload vac3d|diel3d|lorentz3d|nr3d|kerr3d|phc3d] [--size 256] [--dtype f64|f32]

One "step" = one full `fields::step()` (B, H, D, E half-updates, boundaries, DFT) over the whole
grid.  Default workload = BASELINE configs[1]: 3D vacuum cube 256^3, PML 1.0 on all faces
(27 chunks), Ez point dipole, real fields; f64 (the reference's default realnum).

ours:       numpy host model (meep_nl_b200/host.py) -> C ABI -> libmeepcuda.so.
            `value`  : K steps timed on the device (CUDA events on the library's stream, recorded
                       inside meepcu_step), inputs resident in HBM.
            `e2e`    : the same K steps through the C ABI with HOST buffers: meepcu_sync_from_host
                       (all arrays H2D) + meepcu_step(K) (per-step source samples H2D) + one probe
                       read + meepcu_sync_to_host (all arrays D2H), wall clock.
            `roofline`: dominant kernel (k_half<D/E,...>), algorithmic bytes / CUDA-event time.
            `cpu_baseline`: the reference's own CPU code (oracle/_ref/<flavour>/ref_driver,
                       kind "reference") or, if that did not travel, the CPU oracle (kind "port").
reference:  times the reference's CPU implementation on the host cores, same workload/metric.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="vac3d")
    ap.add_argument("--size", dest="n", type=int, default=256)
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--pml", type=float, default=0.0, help="PML thickness of the phc3d workload (default 1.0)")
    ap.add_argument("--cpu-steps", type=int, default=20)
    ap.add_argument("--cpu-size", type=int, default=0, help="grid size of the CPU baseline sample (default: per workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
def build_problem(args, backend, device=0, rank=0, world=1, comm_id=None):
    """weak scaling: the X extent grows with the number of ranks (n*world x n x n), each rank owns
    the chunks of one X slab (contiguous halo planes), PML on the outer faces only"""
    from meep_nl_b200 import host as H
    from meep_nl_b200._cabi import Ex, Ey, Ez
    rt = np.float64 if args.dtype == "f64" else np.float32
    n = args.n
    L = n / 10.0
    w = args.workload
    gv = H.vol3d(L * world, L, 1.5 * L if w == "phc3d" else L, 10)
    split, rank_of = None, None
    if world > 1:
        cuts, rank_of = H.slab_decomposition(gv, world, 0)
        split = {0: cuts}
    _Hmod, _structure, _fields = H, H.structure, H.fields

    class _HS:  # H with the decomposition applied to every structure()/fields() call below
        def __getattr__(self, k): return getattr(_Hmod, k)
        @staticmethod
        def structure(*a, **kw): return _structure(*a, split=split, **kw)
        @staticmethod
        def fields(s, b, device=0):
            cr = [rank_of(sc) for sc in s.chunks] if rank_of else None
            return _fields(s, b, device=device, rank=rank, nranks=world, chunk_rank=cr, comm_id=comm_id)
    H = _HS()
    if w == "vac3d":
        s = H.structure(gv, None, pml=1.0, realnum=rt)
        f = H.fields(s, backend, device=device)
        f.use_real_fields()
        f.add_point_source(Ez, 0.15, 0.1, 0.0, 4.0, gv.center(), 1.0)
    elif w == "diel3d":
        s = H.structure(gv, lambda x, y, z: 1.0 + 1.25 * ((z > L / 3) & (z < 2 * L / 3)), pml=1.0, realnum=rt)
        f = H.fields(s, backend, device=device)
        f.use_real_fields()
        f.add_point_source(Ez, 0.15, 0.1, 0.0, 4.0, gv.center(), 1.0)
    elif w == "lorentz3d":  # BASELINE configs[3] shape
        def ring(x, y, z):
            r = np.sqrt((x - L / 2) ** 2 + (y - L / 2) ** 2)
            return ((r >= 0.15 * L) & (r <= 0.3 * L) & (np.abs(z - L / 2) <= 0.1 * L)) * 1.0
        s = H.structure(gv, lambda x, y, z: 1.0 + 10.56 * ring(x, y, z), pml=1.0, realnum=rt)
        s.add_susceptibility(ring, H.E_stuff, H.lorentzian_susceptibility(1.1, 1e-5))
        f = H.fields(s, backend, device=device)
        f.use_real_fields()
        f.add_point_source(Ez, 0.15, 0.1, 0.0, 4.0, (L / 2 + 0.22 * L, L / 2, L / 2 + 0.05), 1.0)
        lo, hi = int(0.15 * n), int(0.85 * n)
        fr = np.linspace(0.1, 0.2, 100)
        for c in (Ey, Ez):
            f.add_dft_fields(c, (hi, lo, lo), (hi, hi, hi), fr)
            f.add_dft_fields(c, (lo, lo, lo), (lo, hi, hi), fr)
    elif w == "nr3d":  # BASELINE configs[2] (fork Newton-Raphson chi2 slab)
        s = H.structure(gv, lambda x, y, z: 2.25 + 0 * x, pml=1.0, realnum=rt)
        s.set_offdiag_flags()
        s.set_chi2(lambda x, y, z: 1e-2 * ((z > L / 3) & (z < 2 * L / 3)))
        s.set_chi3(lambda x, y, z: 1.0 + 0 * x)
        f = H.fields(s, backend, device=device)
        f.use_real_fields()
        c = gv.center()
        # amplitudes keep chi2*|E| << eps: at chi2*|E| = eps the reference's Jacobian is singular and its
        # seed ladder never converges (newton_raphson.cpp:157-194) -- a regime parity is undefined in
        amp = 1e-6
        f.add_point_source(Ez, 0.3, 0.2, 0.0, 4.0, c, amp)
        f.add_point_source(Ex, 0.3, 0.2, 0.0, 4.0, (c[0] + 0.25, c[1] + 0.1, c[2]), amp)
        f.add_point_source(Ey, 0.3, 0.2, 0.0, 4.0, (c[0], c[1] + 0.25, c[2] + 0.1), amp)
    elif w == "kerr3d":  # BASELINE configs[2], variant 3b: upstream's Kerr (chi3) Pade update, diagonal eps
        from meep_nl_b200._cabi import NL_UPSTREAM_PADE
        s = H.structure(gv, lambda x, y, z: 2.25 + 0 * x, pml=1.0, realnum=rt)
        s.set_chi2(lambda x, y, z: 0.0 * x)
        s.set_chi3(lambda x, y, z: 1e-2 * ((z > L / 3) & (z < 2 * L / 3)))
        cr = [rank_of(sc) for sc in s.chunks] if rank_of else None
        f = _fields(s, backend, nl_mode=NL_UPSTREAM_PADE, device=device, rank=rank, nranks=world, chunk_rank=cr,
                    comm_id=comm_id)
        f.use_real_fields()
        f.add_point_source(Ex, 1.0 / 3.0, 0.2, 0.0, 4.0, gv.center(), 1.0)
    elif w == "phc3d":  # BASELINE configs[4] shape: eps-12 slab, PML in z only, periodic x,y (n x n x 1.5n)
        zlo, zhi = 0.75 * L - 0.25, 0.75 * L + 0.25
        s = H.structure(gv, lambda x, y, z: 1.0 + 11.0 * ((z > zlo) & (z < zhi)), pml=args.pml or 1.0,
                        pml_dirs=[2], realnum=rt)
        f = H.fields(s, backend, device=device)
        f.use_bloch(0, 0.0)
        f.use_bloch(1, 0.0)
        f.use_real_fields()
        f.add_point_source(Ey, 0.25, 0.2, 0.0, 4.0, gv.center(), 1.0)
    else:
        raise SystemExit("unknown workload " + w)
    return gv, f


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled through NVML every few ms while the timed regions run
    (nvidia-smi as a fallback: one sample per ~100 ms)"""

    HW_SLOWDOWN, SW_THERMAL, HW_THERMAL, SW_POWER_CAP = 0x8, 0x20, 0x40, 0x4

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.stop = threading.Event()
        self.samples = []  # (sm_mhz, reasons bitmask)
        self.index = index
        self.sm_max = None
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
        except Exception:
            self.nv = None
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        self.cmd = ["nvidia-smi", "-i", str(index), "--query-gpu=" + q, "--format=csv,noheader,nounits"]

    def _reasons(self):
        nv = self.nv
        for fn in ("nvmlDeviceGetCurrentClocksEventReasons", "nvmlDeviceGetCurrentClocksThrottleReasons"):
            if hasattr(nv, fn):
                return int(getattr(nv, fn)(self.h))
        return 0

    def run(self):
        while not self.stop.is_set():
            try:
                if self.nv is not None:
                    sm = float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                    self.samples.append((sm, self._reasons(), time.perf_counter()))
                    self.stop.wait(0.002)
                    continue
                out = subprocess.run(self.cmd, capture_output=True, text=True, timeout=5).stdout.strip()
                p = [x.strip() for x in out.split(",")]
                if len(p) >= 6:
                    act = [x.lower().startswith("active") for x in p[2:6]]
                    mask = (self.HW_SLOWDOWN * act[0]) | (self.HW_THERMAL * act[1]) | (self.SW_THERMAL * act[2]) | \
                           (self.SW_POWER_CAP * act[3])
                    self.samples.append((float(p[0]), mask, time.perf_counter()))
                    self.sm_max = float(p[1])
            except Exception:
                pass
            self.stop.wait(0.05)

    def summary(self, t0=None, t1=None):
        """clocks / throttle reasons seen inside [t0, t1] (perf_counter); the sampler itself starts earlier so
        that NVML's first-call latency on a fresh box does not eat the window"""
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        inside = [s for s in self.samples if t0 is None or (t0 <= s[2] <= t1)]
        if not inside:  # window shorter than one sampling period: the sample nearest to it
            inside = [min(self.samples, key=lambda s: abs(s[2] - 0.5 * (t0 + t1)))]
        sm = sorted(s[0] for s in inside)
        mask = 0
        for s in inside:
            mask |= s[1]
        names = [(self.HW_SLOWDOWN, "hw_slowdown"), (self.HW_THERMAL, "hw_thermal_slowdown"),
                 (self.SW_THERMAL, "sw_thermal_slowdown"), (self.SW_POWER_CAP, "sw_power_cap")]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.sm_max,
                "reasons": [n for b, n in names if mask & b], "samples": len(sm),
                "source": "nvml" if self.nv is not None else "nvidia-smi"}


def dims(args, n=None):
    n = n or args.n
    return (n, n, n * 3 // 2) if args.workload == "phc3d" else (n, n, n)


def workload_name(args):
    d = dims(args)
    return "%s_%dx%dx%d_%s" % (args.workload, d[0], d[1], d[2], args.dtype)


# reference-driver case per workload, and the largest size the CPU arm runs in a bounded time
# (kerr3d needs the de-blocked Pade flavour of the reference: its CPU arm falls back to the oracle port)
REF_CASES = {"vac3d": ("vac3d:%d", 256), "lorentz3d": ("lorentz3d:%d:100", 192), "nr3d": ("nrslab3d:%d", 96),
             "phc3d": ("phc3d:%d:1.0", 160)}


def run_reference_cpu(args, steps, warmup):
    """the reference's own CPU code on this box's host cores; returns (mcells/s, info) or None.
    A bounded sample: the same workload, at most REF_CASES[..][1] points per side."""
    fl = "fp64" if args.dtype == "f64" else "fp32"
    drv = os.path.join(ROOT, "oracle", "_ref", fl, "ref_driver")
    cores = os.cpu_count() or 1
    if os.path.exists(drv) and args.workload in REF_CASES:
        tmpl, cap = REF_CASES[args.workload]
        n = min(args.n, args.cpu_size or cap)
        case = tmpl % n
        with tempfile.NamedTemporaryFile(suffix=".json") as tf:
            env = dict(os.environ, OMP_NUM_THREADS=str(cores), OMP_PROC_BIND="spread")
            subprocess.run([drv, "--mode", "ref", "--case", case, "--steps", str(steps), "--warmup",
                            str(warmup), "--json", tf.name], check=True, env=env, stdout=subprocess.DEVNULL,
                           stderr=subprocess.DEVNULL, timeout=3000)
            d = json.load(open(tf.name))
        one = None
        if getattr(args, "cpu_one_thread", False):  # the reference's own default is ONE thread (mympi.cpp:187-192)
            with tempfile.NamedTemporaryFile(suffix=".json") as tf:
                env1 = dict(os.environ, OMP_NUM_THREADS="1")
                subprocess.run([drv, "--mode", "ref", "--case", case, "--steps", str(max(2, steps // 4)), "--warmup", "1",
                                "--json", tf.name], check=True, env=env1, stdout=subprocess.DEVNULL,
                               stderr=subprocess.DEVNULL, timeout=3000)
                one = json.load(open(tf.name))["mcells_per_s"]
        dd = dims(args, n)
        scale = "the full problem" if n == args.n else "the same problem at %dx%dx%d (the GPU arm runs %s)" % (
            dd[0], dd[1], dd[2], workload_name(args))
        return d["mcells_per_s"], {"kind": "reference", "cores": d["threads"], "cells": dd[0] * dd[1] * dd[2],
                                   "one_thread_value": one,
                                   "sample": "%d warm-up + %d timed fields::step() of %s, case %s "
                                             "(unmodified reference core, OpenMP, %d threads)"
                                             % (warmup, steps, scale, case, d["threads"])}
    # fall back to the CPU restatement (still the reference's algorithm, single translation unit)
    from meep_nl_b200 import _cabi
    lib = os.path.join(ROOT, "oracle", "liboracle_meep.so")
    if not os.path.exists(lib):
        return None
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    b = _cabi.Backend(lib)
    small = argparse.Namespace(**vars(args))
    small.n = min(args.n, args.cpu_size or 128)
    gv, f = build_problem(small, b)
    f.step(max(warmup, 1))
    t0 = time.perf_counter()
    f.step(steps)
    dt = time.perf_counter() - t0
    f.close()
    return gv.cells() * steps / dt / 1e6, {"kind": "port", "cores": cores, "cells": gv.cells(),
                                           "sample": "%d timed steps of %s at n=%d through the CPU oracle" % (
                                               steps, args.workload, small.n)}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = run_reference_cpu(args, max(2, min(args.steps, args.cpu_steps)), min(args.warmup, 3))
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "neither oracle/_ref nor the oracle library is built"}))
        return
    val, info = r
    steps = max(2, min(args.steps, args.cpu_steps))
    line = {"impl": "reference", "metric": "Mcell-updates/s (fused E+H step)", "value": val, "unit": "Mcell/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 3),
            "ms_per_step": info["cells"] / val / 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": workload_name(args)},
            "cpu_baseline": dict(info, value=val, unit="Mcell/s"),
            "e2e": {"value": val, "unit": "Mcell/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_
        dist = dist_
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from meep_nl_b200 import _cabi
    backend = _cabi.Backend()  # libmeepcuda.so or an exception: no fallback
    comm_id = None
    if world > 1:
        import torch
        idbuf = (C.c_char * 128)()
        if rank == 0:
            backend.check(backend.comm_unique_id(idbuf))
        tid = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device="cuda")
        dist.broadcast(tid, 0)
        comm_id = bytes(tid.cpu().tolist())
    sampler = ClockSampler(local)
    sampler.start()
    t_setup = time.perf_counter()
    gv, f = build_problem(args, backend, device=local, rank=rank, world=world, comm_id=comm_id)
    K, W = args.steps, max(args.warmup, 3)
    t_build = time.perf_counter()
    f.init()
    t_init = time.perf_counter()
    f.step(W)
    if rank == 0:
        sys.stderr.write("[bench] %s: host model %.1f s, attach (describe + upload) %.1f s, %d warm-up steps %.2f s\n"
                         % (workload_name(args), t_build - t_setup, t_init - t_build, W, time.perf_counter() - t_init))
    st0 = f.stats()

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    # ---- timed region: K steps, device time from CUDA events recorded inside meepcu_step ----
    smp = f.samples(K)
    barrier()
    t0 = time.perf_counter()
    t_clk0 = t0
    f.step(K, samples=smp)
    barrier()
    wall = time.perf_counter() - t0
    st1 = f.stats()
    dev_ms = st1.last_step_ms or wall * 1e3  # (0 only with a backend that has no device clock: CPU dry runs)
    # ---- the same K steps again with per-kernel CUDA events (kernels serialised on one stream so
    #      that each duration is the kernel's own): feeds `roofline` ----
    smp_p = f.samples(K)
    backend.check(backend.set_profiling(f.ctx, 1))
    f.step(K, samples=smp_p)
    prof_ms = f.stats().last_step_ms
    t_clk1 = time.perf_counter()
    sampler.stop.set()
    sampler.join()
    kt = (_cabi.KernelTime * 16)()
    nk = C.c_int(0)
    backend.check(backend.get_kernel_times(f.ctx, kt, 16, C.byref(nk)))
    backend.check(backend.set_profiling(f.ctx, 0))
    kernels = [{"name": kt[i].name.decode(), "launches": kt[i].launches, "ms": kt[i].ms,
                "bytes": kt[i].algorithmic_bytes} for i in range(nk.value) if kt[i].launches]
    if dist is not None:
        import torch
        tt = torch.tensor([dev_ms], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms = float(tt.item())
    total_cells = gv.cells()  # the whole grid, all ranks together
    cells = total_cells // world
    value = total_cells * K / (dev_ms * 1e-3) / 1e6

    # ---- end to end through the C ABI with host buffers ----
    f.sync_to_host()  # untimed: the host mirror now holds the current state (what a caller would have)
    smp2 = f.samples(K)
    barrier()
    t0 = time.perf_counter()
    f.sync_from_host()
    f.step(K, samples=smp2)
    cprobe = gv.center()
    probe = f.get_field(4, (args.n / 20.0 + rank * args.n / 10.0, cprobe[1], cprobe[2]))  # Ez, centre of my slab
    f.sync_to_host()
    barrier()
    e2e_s = time.perf_counter() - t0
    nsrc = len(f.src_times)
    hb = f.host_bytes()
    e2e = {"value": total_cells * K / e2e_s / 1e6, "unit": "Mcell/s",
           "h2d_bytes_per_step": int(f.host_bytes(mutable_only=False) / K + 64 * nsrc), "d2h_bytes_per_step": int(hb / K + 8),
           "region": "meepcu_sync_from_host + meepcu_step(%d) + probe read + meepcu_sync_to_host, wall clock" % K}

    # ---- roofline of the dominant kernel ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    with_bytes = [k for k in kernels if k["bytes"] > 0]
    dom = max(with_bytes, key=lambda k: k["ms"]) if with_bytes else None
    roof = None
    if dom:
        per_launch_ms = dom["ms"] / dom["launches"]
        ach = dom["bytes"] / (per_launch_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": dom["name"], "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom["bytes"], "ms_per_launch": per_launch_ms,
                "whole_step": {"achieved": st1.algorithmic_bytes_per_step * K / (dev_ms * 1e-3) / 1e9,
                               "frac": st1.algorithmic_bytes_per_step * K / (dev_ms * 1e-3) / 1e9 / peak,
                               "frac_of_8TBs_nominal": st1.algorithmic_bytes_per_step * K / (dev_ms * 1e-3) / 8e12,
                               "bytes_per_cell": st1.algorithmic_bytes_per_step / max(cells, 1)},
                "serialised_ms_per_step": prof_ms / K,
                "kernels": kernels}
    # DRAM traffic of the dominant kernel from the committed ncu capture of this same workload (or null)
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if roof and workload_name(args).startswith(tr["workload"]) and world == 1:
            roof["traffic"] = tr.get(args.dtype, {}).get(roof["kernel"])
            if roof["traffic"]:
                roof["traffic_source"] = "ncu --set full, profiles/r01d_ncu_full_%s.csv (bytes per launch)" % args.dtype
    except Exception:
        pass
    line = {"metric": "Mcell-updates/s (fused E+H step)", "value": value, "unit": "Mcell/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": workload_name(args), "cells": cells, "chunks": len(f.chunks),
                       "l2": "working set %.1f GB >> 126 MB L2 (no flush needed)" % (st1.device_bytes / 1e9),
                       "parallelism": "1 GPU" if world == 1 else "%d X-slabs, one rank per GPU, NCCL halo exchange (grid %dx%dx%d)" % (world, dims(args)[0] * world, dims(args)[1], dims(args)[2]),
                       "probe_Ez_center": probe, "wall_ms_per_step": wall / K * 1e3},
            "clocks": sampler.summary(t_clk0, t_clk1), "e2e": e2e, "gpu_launches": int(st1.kernel_launches - st0.kernel_launches),
            "roofline": roof}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        args.cpu_one_thread = True
        r = run_reference_cpu(args, args.cpu_steps, 3)
        if r is not None:
            line["cpu_baseline"] = dict(r[1], value=r[0], unit="Mcell/s")
    f.close()
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
