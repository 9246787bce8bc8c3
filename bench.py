#!/usr/bin/env python
"""bench.py -- Mcell-updates/s of the fused E+H time step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload vac3d|diel3d|lorentz3d|nr3d|kerr3d|phc3d] [--size 256] [--dtype f64|f32]
                    [--scaling weak|strong] [--slab-x NX] [--min-seconds 1.0] [--no-extras]

One "step" = one full `fields::step()` (B, H, D, E half-updates, boundaries, DFT) over the whole
grid.  Default workload = BASELINE configs[1]: 3D vacuum cube 256^3, PML 1.0 on all faces
(27 chunks), Ez point dipole, real fields; f64 (the reference's default realnum).

ours:       numpy host model (meep_nl_b200/host.py) -> C ABI -> libmeepcuda.so.
            `value`  : batches of exactly K steps, each timed on the device (CUDA events on the library's
                       stream, recorded inside meepcu_step; max over ranks), repeated until the timed region is
                       >= --min-seconds; the MEDIAN batch is reported (min / max beside it).
            `e2e`    : the same K steps through the C ABI with HOST buffers: meepcu_sync_from_host
                       (all mutable arrays H2D) + meepcu_step(K) (per-step source samples H2D) + one probe
                       read + meepcu_sync_to_host (all mutable arrays D2H), wall clock.
            `e2e_dropin`: the seam itself -- the reference library with the maintainer's patch
                       (meep_nl_b200/shim/dropin_patch.py), its own `fields::step()` called ONCE PER TIME
                       STEP on host objects, libmeepcuda.so underneath (oracle/_ref/fp64_dropin/ref_driver
                       --mode dropin_time); attach time reported separately.
            `parity` : (N > 1) before the timed region the same global problem, pre-split the same way, is
                       stepped on rank 0 alone as ONE context; every owned array of every rank must be
                       bit-identical (the reference's own invariant, tests/three_d.cpp:95-229).  rc != 0 if not.
            `roofline`: dominant kernel, algorithmic bytes / CUDA-event time.
            `cpu_baseline`: the reference's own CPU code (oracle/_ref/<flavour>/ref_driver,
                       kind "reference") or, if that did not travel, the CPU oracle (kind "port").
            `configs`: (unless --no-extras) the other BASELINE configs at full size -- N = 1: config 2 fp32,
                       3a (NR chi2, developed field), 3b (Pade Kerr), 4 (Lorentzian + DFT), 5 (1024x1024x1536 fp32);
                       N > 1: config 5 strong (1024x1024x1536 split over N) and weak (128 N x 1024 x 1536).
reference:  times the reference's CPU implementation on the host cores, same workload/metric/steps/warm-up.
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Mcell-updates/s (fused E+H step)"
# Hooks for tests/bench_dryrun_worker.py (the host-side logic of the N > 1 path checked without a GPU): the
# process-group backend, the device collectives' tensors live on, and a callback run after every fields.init().
PG_BACKEND, DEV, AFTER_INIT = "nccl", "cuda", None


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="vac3d")
    ap.add_argument("--size", dest="n", type=int, default=256)
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = every GPU gets a --slab-x (default --size) thick X slab, strong = the N = 1 grid is cut in N")
    ap.add_argument("--slab-x", type=int, default=0, help="X extent per GPU (weak) / of the whole grid (strong); default --size")
    ap.add_argument("--pml", type=float, default=0.0, help="PML thickness of the phc3d workload (default 1.0)")
    ap.add_argument("--amp", type=float, default=1.0, help="source amplitude of the nr3d workload")
    ap.add_argument("--develop", type=int, default=-1, help="untimed steps that let the field fill the cell before the "
                    "timed region (default: nr3d / kerr3d 2 n, others 0)")
    ap.add_argument("--min-seconds", type=float, default=1.0, help="repeat the K-step batch until the timed region is this long")
    ap.add_argument("--cpu-steps", type=int, default=20)
    ap.add_argument("--cpu-size", type=int, default=0, help="grid size of the CPU baseline sample (default: per workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="main workload only (no `configs` block)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the N-GPU == 1-GPU gate")
    ap.add_argument("--no-dropin", action="store_true", help="skip the e2e_dropin leg")
    ap.add_argument("--parity-steps", type=int, default=25)
    return ap.parse_args(argv)


# ------------------------------------------------------------------------------------------------
def grid_dims(args, nslabs=1):
    """cells per direction of the WHOLE grid for `nslabs` GPUs"""
    n = args.n
    nx = args.slab_x or n
    if args.scaling == "weak":
        nx *= nslabs
    return (nx, n, n * 3 // 2) if args.workload == "phc3d" else (nx, n, n)


def workload_name(args, nslabs=1):
    d = grid_dims(args, nslabs)
    return "%s_%dx%dx%d_%s" % (args.workload, d[0], d[1], d[2], args.dtype)


def config_of(args, nslabs=1):
    """identical in both arms (ours / reference)"""
    d = grid_dims(args, nslabs)
    note = None
    if nslabs > 1 and args.workload != "phc3d":
        note = ("PML on all six faces of the WHOLE grid: an end rank has one X-PML face, a middle rank none, the single GPU "
                "two -- per-GPU work is not constant (efficiency can exceed 1); constant per-GPU work: configs.config5_weak")
    return {"workload": workload_name(args, nslabs), "grid": list(d), "cells": d[0] * d[1] * d[2],
            "scaling": args.scaling if nslabs > 1 else "single GPU", **({"note": note} if note else {}),
            "l2": "working set >> 126 MB L2, every array streamed once per half step (no flush needed)"}


def build_problem(args, backend, device=0, rank=0, world=1, comm_id=None, nslabs=None):
    """The grid of `nslabs` GPUs (weak: X extent grows with nslabs; strong: fixed), cut into nslabs X slabs
    (contiguous halo planes).  world == nslabs: rank r steps slab r.  world == 1 < nslabs: the SAME chunks, all
    stepped by one context (the pre-split single-GPU twin of the parity gate)."""
    from meep_nl_b200 import host as H
    from meep_nl_b200._cabi import Ex, Ey, Ez
    rt = np.float64 if args.dtype == "f64" else np.float32
    nslabs = nslabs or world
    n = args.n
    w = args.workload
    d = grid_dims(args, nslabs)
    L = n / 10.0
    gv = H.vol3d(d[0] / 10.0, d[1] / 10.0, d[2] / 10.0, 10)
    split, rank_of = None, None
    if nslabs > 1:
        cuts, rank_of = H.slab_decomposition(gv, nslabs, 0)
        split = {0: cuts}
    if world == 1:
        rank_of = None
    _Hmod, _structure, _fields = H, H.structure, H.fields

    class _HS:  # H with the decomposition applied to every structure()/fields() call below
        def __getattr__(self, k): return getattr(_Hmod, k)
        @staticmethod
        def structure(*a, **kw): return _structure(*a, split=split, **kw)
        @staticmethod
        def fields(s, b, device=0, **kw):
            cr = [rank_of(sc) for sc in s.chunks] if rank_of else None
            return _fields(s, b, device=device, rank=rank, nranks=world, chunk_rank=cr, comm_id=comm_id, **kw)
    H = _HS()
    c = gv.center()
    if nslabs > 1:
        # the centre of a grid cut into an even number of X slabs lies ON a cut: a dipole there touches the faces of two
        # chunks, which then keep storing E (the virtual-E rule wants dipoles >= 2 cells inside a chunk) and set the pace
        # of the whole job.  4.5 cells off the cut; the single-GPU twin of the parity gate is built by this same function
        c = (c[0] + 0.45, c[1], c[2])
    if w == "vac3d":
        s = H.structure(gv, None, pml=1.0, realnum=rt)
        f = H.fields(s, backend, device=device)
        f.use_real_fields()
        f.add_point_source(Ez, 0.15, 0.1, 0.0, 4.0, c, 1.0)
    elif w == "diel3d":
        s = H.structure(gv, lambda x, y, z: 1.0 + 1.25 * ((z > L / 3) & (z < 2 * L / 3)), pml=1.0, realnum=rt)
        f = H.fields(s, backend, device=device)
        f.use_real_fields()
        f.add_point_source(Ez, 0.15, 0.1, 0.0, 4.0, c, 1.0)
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
        for cc in (Ey, Ez):
            f.add_dft_fields(cc, (hi, lo, lo), (hi, hi, hi), fr)
            f.add_dft_fields(cc, (lo, lo, lo), (lo, hi, hi), fr)
    elif w == "nr3d":  # BASELINE configs[2] (fork Newton-Raphson chi2 slab); the parity cases' amplitude and chi2
        s = H.structure(gv, lambda x, y, z: 2.25 + 0 * x, pml=1.0, realnum=rt)
        s.set_offdiag_flags()
        s.set_chi2(lambda x, y, z: 1e-2 * ((z > L / 3) & (z < 2 * L / 3)))
        s.set_chi3(lambda x, y, z: 1.0 + 0 * x)
        f = H.fields(s, backend, device=device)
        f.use_real_fields()
        # the parity cases' amplitude (1.0) and chi2 (1e-2), but a narrow-band source (width 3: the Gaussian peaks after
        # 800 steps and lasts 1600) so that the timed region sees a field that fills the cell, not a pulse that has left
        amp = args.amp
        f.add_point_source(Ez, 0.3, 3.0, 0.0, 4.0, c, amp)
        f.add_point_source(Ex, 0.3, 3.0, 0.0, 4.0, (c[0] + 0.25, c[1] + 0.1, c[2]), amp)
        f.add_point_source(Ey, 0.3, 3.0, 0.0, 4.0, (c[0], c[1] + 0.25, c[2] + 0.1), amp)
    elif w == "kerr3d":  # BASELINE configs[2], variant 3b: upstream's Kerr (chi3) Pade update, diagonal eps
        from meep_nl_b200._cabi import NL_UPSTREAM_PADE
        s = H.structure(gv, lambda x, y, z: 2.25 + 0 * x, pml=1.0, realnum=rt)
        s.set_chi2(lambda x, y, z: 0.0 * x)
        s.set_chi3(lambda x, y, z: 1e-2 * ((z > L / 3) & (z < 2 * L / 3)))
        f = H.fields(s, backend, device=device, nl_mode=NL_UPSTREAM_PADE)
        f.use_real_fields()
        f.add_point_source(Ex, 1.0 / 3.0, 3.0, 0.0, 4.0, c, 1.0)  # (narrow band: see nr3d)
    elif w == "phc3d":  # BASELINE configs[4] shape: eps-12 slab, PML in z only, periodic x,y (nx x n x 1.5n)
        Lz = d[2] / 10.0
        zlo, zhi = 0.5 * Lz - 0.25, 0.5 * Lz + 0.25
        s = H.structure(gv, lambda x, y, z: 1.0 + 11.0 * ((z > zlo) & (z < zhi)), pml=args.pml or 1.0,
                        pml_dirs=[2], realnum=rt)
        f = H.fields(s, backend, device=device)
        f.use_bloch(0, 0.0)
        f.use_bloch(1, 0.0)
        f.use_real_fields()
        f.add_point_source(Ey, 0.25, 0.2, 0.0, 4.0, c, 1.0)
    else:
        raise SystemExit("unknown workload " + w)
    f.probe_component = {"kerr3d": Ex, "phc3d": Ey}.get(w, Ez)
    return gv, f


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled through NVML every few ms while the timed regions run
    (nvidia-smi as a fallback: one sample per ~100 ms)"""

    HW_SLOWDOWN, SW_THERMAL, HW_THERMAL, SW_POWER_CAP = 0x8, 0x20, 0x40, 0x4

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.stop = threading.Event()
        self.samples = []  # (sm_mhz, reasons bitmask, t)
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


# reference-driver case per workload, and the largest size the CPU arm runs in a bounded time
# (kerr3d needs the de-blocked Pade flavour of the reference: its CPU arm falls back to the oracle port)
REF_CASES = {"vac3d": ("vac3d:%d", 256), "lorentz3d": ("lorentz3d:%d:100", 192), "nr3d": ("nrslab3d:%d", 96),
             "phc3d": ("phc3d:%d:1.0", 160)}


def run_reference_cpu(args, steps, warmup, nslabs=1, one_thread=False):
    """the reference's own CPU code on this box's host cores; returns (mcells/s, info) or None.
    A bounded sample: the same workload, at most REF_CASES[..][1] points per side, and never more than
    ~3 minutes of CPU work (the grid shrinks, never the step count)."""
    fl = "fp64" if args.dtype == "f64" else "fp32"
    drv = os.path.join(ROOT, "oracle", "_ref", fl, "ref_driver")
    cores = os.cpu_count() or 1
    if os.path.exists(drv) and args.workload in REF_CASES:
        tmpl, cap = REF_CASES[args.workload]
        n = min(args.n, args.cpu_size or cap)
        mult = nslabs if (args.workload == "vac3d" and args.scaling == "weak" and not args.slab_x) else 1
        # ~200 Mcell/s on 16 cores (BENCH_r01): keep (steps + warmup) * cells under ~3 minutes
        while n > 32 and (steps + warmup) * float(n) ** 3 * mult / 150e6 > 180.0:
            if mult > 1:
                mult = 1
            else:
                n = n * 3 // 4
        case = tmpl % n + (":%d" % mult if mult > 1 else "")
        with tempfile.NamedTemporaryFile(suffix=".json") as tf:
            env = dict(os.environ, OMP_NUM_THREADS=str(cores), OMP_PROC_BIND="spread")
            subprocess.run([drv, "--mode", "ref", "--case", case, "--steps", str(steps), "--warmup",
                            str(warmup), "--json", tf.name], check=True, env=env, stdout=subprocess.DEVNULL,
                           stderr=subprocess.DEVNULL, timeout=3000)
            d = json.load(open(tf.name))
        one = None
        if one_thread:  # the reference's own default is ONE thread (mympi.cpp:187-192)
            with tempfile.NamedTemporaryFile(suffix=".json") as tf:
                env1 = dict(os.environ, OMP_NUM_THREADS="1")
                subprocess.run([drv, "--mode", "ref", "--case", case, "--steps", str(max(2, min(steps, 20) // 4)), "--warmup", "1",
                                "--json", tf.name], check=True, env=env1, stdout=subprocess.DEVNULL,
                               stderr=subprocess.DEVNULL, timeout=3000)
                one = json.load(open(tf.name))["mcells_per_s"]
        full = grid_dims(args, nslabs)
        got = d["cells"]
        scale = "the full problem" if got == full[0] * full[1] * full[2] else \
            "the same problem at %d cells (the GPU arm runs %s: %d cells)" % (got, workload_name(args, nslabs), full[0] * full[1] * full[2])
        return d["mcells_per_s"], {"kind": "reference", "cores": d["threads"], "cells": got,
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
    small.slab_x = 0
    gv, f = build_problem(small, b)
    f.step(max(warmup, 1))
    t0 = time.perf_counter()
    f.step(steps)
    dt = time.perf_counter() - t0
    f.close()
    return gv.cells() * steps / dt / 1e6, {"kind": "port", "cores": cores, "cells": gv.cells(), "one_thread_value": None,
                                           "sample": "%d warm-up + %d timed steps of %s at n=%d through the CPU oracle "
                                                     "(restatement of the reference, OpenMP)" % (max(warmup, 1), steps, args.workload, small.n)}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    steps, warmup = args.steps, args.warmup  # honoured exactly: the sample shrinks in space, not in steps
    r = run_reference_cpu(args, steps, warmup, nslabs=world)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "neither oracle/_ref nor the oracle library is built"}))
        return
    val, info = r
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Mcell/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": info["cells"] / val / 1e3, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": config_of(args, world),
            "cpu_baseline": dict(info, value=val, unit="Mcell/s"),
            "e2e": {"value": val, "unit": "Mcell/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def digest_arrays(arrs):
    """name -> 128-bit digest of the array's bytes (bit-equality of two runs == equality of the digests)"""
    def one(item):
        name, a = item
        return name, hashlib.blake2b(memoryview(np.ascontiguousarray(a)).cast("B"), digest_size=16).hexdigest()
    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:  # hashlib releases the GIL
        return dict(ex.map(one, arrs))


def parity_gate(args, backend, f, rank, world, local, dist, steps):
    """N ranks == 1 rank, bit for bit, on every owned array.  `f` has been stepped `steps` times from t = 0."""
    t0 = time.perf_counter()
    f.sync_to_host()  # collective: also refreshes the lazily exchanged B / D / P halos
    mine = digest_arrays(f.owned_arrays())
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    verdict = [None]
    tmp = None
    if rank == 0:
        gv1, f1 = build_problem(args, backend, device=local, rank=0, world=1, nslabs=world)
        f1.pin_host = False  # one sync of this twin: not worth page-locking tens of GB
        f1.init()
        if AFTER_INIT: AFTER_INIT(f1, 1)
        f1.step(steps)
        f1.sync_to_host()
        ref_arrays = dict(f1.owned_arrays())
        ref = digest_arrays(list(ref_arrays.items()))
        got = {}
        for g in gathered:
            got.update(g)
        bad = sorted(k for k in ref if got.get(k) != ref[k])
        missing = sorted(set(ref) ^ set(got))
        nonzero = sum(1 for a in ref_arrays.values() if np.any(a))
        verdict[0] = {"bad": bad, "missing": missing, "arrays": len(ref), "nonzero_arrays": nonzero,
                      "elements": int(sum(a.size for a in ref_arrays.values()))}
        if bad:
            tmp = tempfile.mkdtemp(prefix="meepcu_parity_")
            verdict[0]["tmp"] = tmp
    dist.broadcast_object_list(verdict, src=0)
    v = verdict[0]
    mism = 0
    if v["bad"]:  # failure path: the ranks (one node) hand the offending arrays over through the file system
        mine_arr = dict(f.owned_arrays())
        for k in v["bad"]:
            if k in mine_arr:
                np.save(os.path.join(v["tmp"], k.replace("/", "_") + ".npy"), mine_arr[k])
        dist.barrier()
        if rank == 0:
            for k in v["bad"]:
                p = os.path.join(v["tmp"], k.replace("/", "_") + ".npy")
                if os.path.exists(p):
                    a, b = np.load(p).reshape(-1), ref_arrays[k].reshape(-1)
                    ut = np.uint64 if a.dtype.itemsize % 8 == 0 else np.uint32  # bit patterns, not values
                    mism += int(np.count_nonzero(a.view(ut) != b.view(ut))) if a.size == b.size else max(a.size, b.size)
            shutil.rmtree(v["tmp"], ignore_errors=True)
    out = None
    if rank == 0:
        f1.close()
        del f1, ref_arrays
        out = {"ranks": world, "steps": steps, "arrays": v["arrays"], "nonzero_arrays": v["nonzero_arrays"],
               "elements": v["elements"], "mismatched_arrays": len(v["bad"]) + len(v["missing"]),
               "mismatched_elements": mism if v["bad"] else 0, "first_bad": (v["bad"] + v["missing"] + [""])[0],
               "seconds": time.perf_counter() - t0,
               "how": "the same %s grid pre-split into %d X slabs, stepped as ONE context on rank 0; blake2b digest of every "
                      "owned array (fields, PML auxiliaries, polarisations, DFT) of every rank compared" % (workload_name(args, world), world)}
    return out, (len(v["bad"]) + len(v["missing"]) == 0)


def measure(args, backend, rank, world, local, dist, comm_id, sampler, with_parity):
    """build the workload, run the parity gate (N > 1), the timed batches, the profiled batch and the e2e leg"""
    from meep_nl_b200 import _cabi
    t_setup = time.perf_counter()
    gv, f = build_problem(args, backend, device=local, rank=rank, world=world, comm_id=comm_id)
    K, W = args.steps, max(args.warmup, 3)
    t_build = time.perf_counter()
    f.init()
    if AFTER_INIT: AFTER_INIT(f, world)
    t_init = time.perf_counter()

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            if DEV == "cuda": torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        tt = torch.tensor([x], device=DEV, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    parity, parity_ok = None, True
    if with_parity:
        f.step(args.parity_steps)
        parity, parity_ok = parity_gate(args, backend, f, rank, world, local, dist, args.parity_steps)
    develop = args.develop if args.develop >= 0 else (2 * args.n if args.workload in ("nr3d", "kerr3d") else 0)
    if develop:
        f.step(develop)
    f.step(W)
    f.stats()  # waits
    if rank == 0:
        sys.stderr.write("[bench] %s: host model %.1f s, attach (describe + upload) %.1f s, %d develop + %d warm-up steps %.2f s\n"
                         % (workload_name(args, world), t_build - t_setup, t_init - t_build, develop, W, time.perf_counter() - t_init))
    st0 = f.stats()
    field_rms = None
    if args.workload in ("nr3d", "kerr3d"):  # how developed the field is when the clock starts
        sl = f.get_array_slice(f.probe_component, (8, 8, gv.num[2] // 2), (gv.num[0] - 8, gv.num[1] - 8, gv.num[2] // 2))
        sl = sl[np.isfinite(sl)]
        field_rms = {"plane": "z = nz/2", "rms": float(np.sqrt(np.mean(sl ** 2))) if sl.size else None,
                     "max_abs": float(np.max(np.abs(sl))) if sl.size else None,
                     "fraction_above_1e-6_of_max": float(np.mean(np.abs(sl) > 1e-6 * np.max(np.abs(sl)))) if sl.size and np.max(np.abs(sl)) > 0 else 0.0}

    # ---- timed region: batches of exactly K steps, device time from CUDA events recorded inside meepcu_step ----
    batches, walls = [], []
    t_clk0 = time.perf_counter()
    while True:
        smp = f.samples(K)
        barrier()
        t0 = time.perf_counter()
        f.step(K, samples=smp)
        ms = f.stats().last_step_ms  # waits for the batch
        barrier()
        walls.append(time.perf_counter() - t0)
        batches.append(max_over_ranks(ms or walls[-1] * 1e3))  # (0 only with a backend without a device clock: CPU dry runs)
        if (len(batches) >= 3 and sum(batches) >= args.min_seconds * 1e3) or len(batches) >= 400:
            break
    t_clk1 = time.perf_counter()
    st1 = f.stats()
    srt = sorted(batches)
    dev_ms = srt[len(srt) // 2]
    nr_failed = int(st1.reserved[0] - st0.reserved[0])  # cells the deterministic NR seed ladder left unchanged, timed region only
    # ---- the same K steps again with per-kernel CUDA events (kernels serialised on one stream so
    #      that each duration is the kernel's own): feeds `roofline` ----
    smp_p = f.samples(K)
    backend.check(backend.set_profiling(f.ctx, 1))
    f.step(K, samples=smp_p)
    stp = f.stats()
    prof_ms = stp.last_step_ms
    kt = (_cabi.KernelTime * 16)()
    nk = C.c_int(0)
    backend.check(backend.get_kernel_times(f.ctx, kt, 16, C.byref(nk)))
    # the same byte counts with SURVEY.md 8(d)'s per-cell figures (21 / 24 values per vacuum / dielectric cell): what
    # `roofline.achieved` is quoted on; kt[i].algorithmic_bytes is what this implementation has to move (18 / 21 where a
    # chunk does not store E while stepping)
    sv_k = (C.c_int64 * 16)()
    sv_step, nsv = C.c_int64(0), C.c_int(0)
    backend.check(backend.get_compulsory_bytes(f.ctx, C.byref(sv_step), sv_k, 16, C.byref(nsv)))
    backend.check(backend.set_profiling(f.ctx, 0))
    nr_stats = None
    if args.workload == "nr3d":  # two more steps with the Newton-Raphson counters on (atomics: outside every timed leg)
        backend.check(backend.set_profiling(f.ctx, 2))
        f.step(2)
        stn = f.stats()
        backend.check(backend.set_profiling(f.ctx, 0))
        if stn.reserved[1]:
            nr_stats = {"solves_per_step": stn.reserved[1] / 2, "mean_iterations_per_solve": stn.reserved[2] / stn.reserved[1],
                        "max_iterations_of_a_solve": int(stn.reserved[6]),
                        "unconverged_cells_per_step": (stn.reserved[0] - stp.reserved[0]) / 2}
    kernels = [{"name": kt[i].name.decode(), "launches": kt[i].launches, "ms": kt[i].ms,
                "bytes": int(sv_k[i]) if i < nsv.value else kt[i].algorithmic_bytes, "bytes_moved": kt[i].algorithmic_bytes}
               for i in range(nk.value) if kt[i].launches]
    total_cells = gv.cells()  # the whole grid, all ranks together
    cells = total_cells // world
    value = total_cells * K / (dev_ms * 1e-3) / 1e6

    # ---- end to end through the C ABI with host buffers ----
    f.sync_to_host()  # untimed: the host mirror now holds the current state (what a caller would have)
    e2e_runs = []
    probe = None
    for _ in range(2 if total_cells <= 2 ** 27 else 1):
        smp2 = f.samples(K)
        barrier()
        t0 = time.perf_counter()
        f.sync_from_host()
        f.step(K, samples=smp2)
        probe = f.get_field(f.probe_component, gv.center())  # the rank that owns the source point answers
        f.sync_to_host()
        barrier()
        e2e_runs.append(time.perf_counter() - t0)
    e2e_s = min(e2e_runs)
    if dist is not None:
        probes = [None] * world
        dist.all_gather_object(probes, probe)
        probe = next((p for p in probes if p is not None), None)
    nsrc = len(f.src_times)
    hb = f.host_bytes()
    e2e = {"value": total_cells * K / e2e_s / 1e6, "unit": "Mcell/s",
           "h2d_bytes_per_step": int(hb / K + 64 * nsrc), "d2h_bytes_per_step": int(hb / K + 8),
           "region": "meepcu_sync_from_host (mutable arrays) + meepcu_step(%d) + probe read + meepcu_sync_to_host, wall clock, best of %d" % (K, len(e2e_runs))}

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
        ach_moved = dom["bytes_moved"] / (per_launch_ms * 1e-3) / 1e9
        step_bytes = int(sv_step.value) or st1.algorithmic_bytes_per_step
        step_moved = st1.algorithmic_bytes_per_step
        roof = {"bound": "hbm", "kernel": dom["name"], "achieved": ach_moved, "peak": peak, "unit": "GB/s",
                "frac": ach_moved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom["bytes_moved"], "ms_per_launch": per_launch_ms,
                "accounting": "per kernel: achieved = the bytes this launch has to move (every array element once per read and once "
                              "per write; chunks that do not store E while stepping move 18 values per vacuum cell and step, 21 "
                              "with eps) / its CUDA-event time -- comparable with `traffic` (ncu DRAM bytes).  `survey` = the same "
                              "launch credited with SURVEY.md 8(d)'s figures (21 / 24 values: E written by the D/E half, read back "
                              "by the B/H half) -- an effective bandwidth, which exceeds the moved one for the D/E half and falls "
                              "below it for the B/H half of such chunks; over the whole step (`whole_step.frac`) it is the "
                              "metric's own byte count per cell update",
                "survey": {"bytes_per_launch": dom["bytes"], "achieved": ach, "frac": ach / peak},
                "whole_step": {"achieved": step_bytes * K / (dev_ms * 1e-3) / 1e9,
                               "frac": step_bytes * K / (dev_ms * 1e-3) / 1e9 / peak,
                               "frac_of_8TBs_nominal": step_bytes * K / (dev_ms * 1e-3) / 8e12,
                               "bytes_per_cell": step_bytes / max(cells, 1),
                               "moved_bytes_per_cell": step_moved / max(cells, 1),
                               "moved_achieved": step_moved * K / (dev_ms * 1e-3) / 1e9,
                               "moved_frac": step_moved * K / (dev_ms * 1e-3) / 1e9 / peak},
                "serialised_ms_per_step": prof_ms / K,
                "kernels": kernels}
        # DRAM traffic of the dominant kernel from the committed ncu capture of this same workload (or null)
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            if workload_name(args, world).startswith(tr["workload"]) and world == 1:
                roof["traffic"] = tr.get(args.dtype, {}).get(roof["kernel"])
                if roof["traffic"]:
                    roof["traffic_source"] = tr.get("source", "ncu --set full, profiles/ (bytes per launch)")
        except Exception:
            pass
    res = {"gv": gv, "value": value, "dev_ms": dev_ms, "batches": batches, "walls": walls, "K": K, "W": W,
           "cells": cells, "total_cells": total_cells, "chunks": len(f.chunks), "probe": probe, "e2e": e2e,
           "launches": int(st1.kernel_launches - st0.kernel_launches), "roof": roof, "parity": parity,
           "parity_ok": parity_ok, "clocks": sampler.summary(t_clk0, t_clk1) if sampler else None,
           "device_bytes": int(st1.device_bytes), "develop": develop, "nr_unconverged_cells": nr_failed,
           "t_final": f.t, "nr_stats": nr_stats, "field_at_start": field_rms, "virtual_e_arrays": int(st1.reserved[5])}
    f.close()
    return res


def short_line(args, world, r):
    """what a `configs` entry keeps of a measurement"""
    roof = r["roof"] or {}
    ws = roof.get("whole_step", {})
    srt = sorted(r["batches"])
    return {"workload": workload_name(args, world), "scaling": args.scaling if world > 1 else "single GPU", "dtype": args.dtype,
            "n_gpus": world, "value": r["value"], "unit": "Mcell/s",
            "ms_per_step": r["dev_ms"] / r["K"], "ms_per_step_min": srt[0] / r["K"], "batches": len(srt), "steps": r["K"],
            "developed_steps": r["develop"], "whole_step_frac": ws.get("frac"), "whole_step_frac_of_8TBs": ws.get("frac_of_8TBs_nominal"),
            "bytes_per_cell": ws.get("bytes_per_cell"), "dominant_kernel": roof.get("kernel"), "dominant_frac": roof.get("frac"),
            "kernels": roof.get("kernels"), "probe": r["probe"], "virtual_e_arrays": r["virtual_e_arrays"],
            "e2e": r["e2e"]["value"], "parity": r["parity"]}


EXTRAS_1GPU = [  # (label, argv)
    ("config2_fp32", ["--workload", "vac3d", "--size", "256", "--dtype", "f32"]),
    ("config3a_nr_chi2", ["--workload", "nr3d", "--size", "384", "--dtype", "f64", "--steps", "20"]),
    ("config3b_pade_kerr", ["--workload", "kerr3d", "--size", "384", "--dtype", "f64", "--steps", "20"]),
    ("config4_lorentz_dft", ["--workload", "lorentz3d", "--size", "512", "--dtype", "f64", "--steps", "20"]),
    ("config5_phc_slab", ["--workload", "phc3d", "--size", "1024", "--dtype", "f32", "--steps", "10"]),
]


def dropin_leg(args):
    """fields::step() of the PATCHED reference, one call per time step, libmeepcuda.so underneath"""
    drv = os.path.join(ROOT, "oracle", "_ref", "fp64_dropin", "ref_driver")
    lib = os.path.join(ROOT, "meep_nl_b200", "libmeepcuda.so")
    if args.dtype != "f64" or args.workload not in REF_CASES or not os.path.exists(drv):
        return None
    case = REF_CASES[args.workload][0] % args.n
    steps = max(args.steps, 200)
    with tempfile.NamedTemporaryFile(suffix=".json") as tf:
        env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1))
        r = subprocess.run([drv, "--mode", "dropin_time", "--lib", lib, "--case", case, "--steps", str(steps), "--warmup", "10",
                            "--json", tf.name], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, timeout=1200)
        if r.returncode != 0:
            return {"error": r.stderr.decode(errors="replace")[-300:]}
        d = json.load(open(tf.name))
    return {"ms_per_step": d["ms_per_step"], "device_ms_per_step": d["device_ms_per_step"], "value": d["mcells_per_s"],
            "unit": "Mcell/s", "steps": d["steps"], "attach_seconds": d["attach_seconds"],
            "enqueue_seconds": d["enqueue_seconds"], "sync_to_host_seconds": d["sync_to_host_seconds"],
            "kernel_launches_per_step": d["kernel_launches_per_step"], "case": case, "probe": d["probe"],
            "region": "%d calls of the patched reference's own fields::step() (one per time step) + one get_field probe, "
                      "wall clock; attach (describe + upload) and the final full download reported separately" % d["steps"]}


def main_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_
        dist = dist_
        if DEV == "cuda":
            torch.cuda.set_device(local)
            dist.init_process_group(PG_BACKEND, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(PG_BACKEND)
    from meep_nl_b200 import _cabi
    backend = _cabi.Backend()  # libmeepcuda.so or an exception: no fallback

    def new_comm_id():
        if world == 1:
            return None
        import torch
        idbuf = (C.c_char * 128)()
        if rank == 0:
            backend.check(backend.comm_unique_id(idbuf))
        tid = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device=DEV)
        dist.broadcast(tid, 0)
        return bytes(tid.cpu().tolist())

    sampler = ClockSampler(local)
    sampler.start()
    r = measure(args, backend, rank, world, local, dist, new_comm_id(), sampler, world > 1 and not args.no_parity)
    sampler.stop.set()
    sampler.join()
    K = r["K"]
    srt = sorted(r["batches"])
    cfg = config_of(args, world)
    line = {"metric": METRIC, "value": r["value"], "unit": "Mcell/s", "n_gpus": world,
            "steps": K, "warmup": r["W"], "ms_per_step": r["dev_ms"] / K, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": cfg,
            "timing": {"batches": len(srt), "steps_per_batch": K, "timed_region_s": sum(r["batches"]) * 1e-3,
                       "ms_per_step_median": r["dev_ms"] / K, "ms_per_step_min": srt[0] / K, "ms_per_step_max": srt[-1] / K,
                       "wall_ms_per_step_median": sorted(r["walls"])[len(r["walls"]) // 2] / K * 1e3,
                       "how": "each batch = exactly K steps between barrier + synchronize, device time from CUDA events on the "
                              "library's stream, max over ranks; value = the median batch"},
            "details": {"cells_per_gpu": r["cells"], "chunks": r["chunks"], "device_bytes_per_gpu": r["device_bytes"],
                        "parallelism": "1 GPU" if world == 1 else "%d X-slabs, one rank per GPU, NCCL halo exchange beside the interior tiles" % world,
                        "probe_at_source": r["probe"], "developed_steps": r["develop"], "t_final": r["t_final"],
                        "nr_unconverged_cells_in_timed_region": r["nr_unconverged_cells"], "newton_raphson": r["nr_stats"],
                        "field_at_start_of_timed_region": r["field_at_start"], "virtual_e_arrays": r["virtual_e_arrays"]},
            "clocks": r["clocks"], "e2e": r["e2e"], "gpu_launches": r["launches"], "roofline": r["roof"]}
    if r["parity"] is not None:
        line["parity"] = r["parity"]
    ok = r["parity_ok"]
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = run_reference_cpu(args, args.cpu_steps, 3, one_thread=True)
        if cb is not None:
            line["cpu_baseline"] = dict(cb[1], value=cb[0], unit="Mcell/s")
    if rank == 0 and world == 1 and not args.no_dropin:
        try:
            d = dropin_leg(args)
            if d is not None:
                line["e2e_dropin"] = d
        except Exception as e:  # the leg needs oracle/_ref/fp64_dropin, which travels with the snapshot
            line["e2e_dropin"] = {"error": repr(e)[:300]}
    # ---- the other BASELINE configs ----
    if not args.no_extras and args.workload == "vac3d" and args.n == 256 and args.dtype == "f64" and not args.slab_x:
        configs = {}
        if world == 1:
            for label, argv in EXTRAS_1GPU:  # separate processes: a failure there cannot take the main line with it
                cmd = [sys.executable, os.path.abspath(__file__), "--no-extras", "--no-cpu-baseline", "--no-dropin",
                       "--min-seconds", "1.0"] + argv
                t0 = time.perf_counter()
                try:
                    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
                    js = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
                    if p.returncode == 0 and js:
                        d = json.loads(js[-1])
                        ws = (d.get("roofline") or {}).get("whole_step", {})
                        configs[label] = {"workload": d["config"]["workload"], "dtype": d["dtype"], "value": d["value"], "unit": "Mcell/s",
                                          "ms_per_step": d["ms_per_step"], "timing": d["timing"], "whole_step_frac": ws.get("frac"),
                                          "whole_step_frac_of_8TBs": ws.get("frac_of_8TBs_nominal"), "bytes_per_cell": ws.get("bytes_per_cell"),
                                          "dominant_kernel": (d.get("roofline") or {}).get("kernel"),
                                          "dominant_frac": (d.get("roofline") or {}).get("frac"),
                                          "kernels": (d.get("roofline") or {}).get("kernels"), "details": d["details"],
                                          "e2e": d["e2e"]["value"], "clocks": d["clocks"], "wall_s": time.perf_counter() - t0}
                    else:
                        configs[label] = {"error": (p.stderr or p.stdout)[-400:]}
                except Exception as e:
                    configs[label] = {"error": repr(e)[:300]}
        else:
            for label, scaling, slab in (("config5_strong", "strong", 1024), ("config5_weak", "weak", 128)):
                common = ["--workload", "phc3d", "--dtype", "f32", "--steps", "10", "--warmup", "3", "--scaling", scaling,
                          "--slab-x", str(slab), "--parity-steps", "12"]
                a5 = parse(common + ["--size", "1024"])
                # the gate's single-GPU twin of the full grid would need 78 GB of host memory next to the ranks' own
                # arrays: the gate runs on the same slabs with y, z extents / 4 (same X cuts, same kernels, same
                # exchange code); full-size arrays are covered by the 1-GPU twins of tests/test_cuda_vs_reference.py
                a5p = parse(common + ["--size", "256", "--min-seconds", "0.05"])
                try:
                    rp = measure(a5p, backend, rank, world, local, dist, new_comm_id(), None, not args.no_parity)
                    r5 = measure(a5, backend, rank, world, local, dist, new_comm_id(), None, False)
                    configs[label] = short_line(a5, world, r5)
                    configs[label]["parity"] = rp["parity"]
                    if rp["parity"]:
                        configs[label]["parity"]["grid"] = workload_name(a5p, world)
                    ok = ok and rp["parity_ok"]
                except Exception as e:
                    configs[label] = {"error": repr(e)[:300]}
                    raise
        line["configs"] = configs
    if rank == 0:
        print(json.dumps(line))
        sys.stdout.flush()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if not ok:
        sys.stderr.write("[bench] PARITY FAILURE: N-GPU result differs from the 1-GPU result\n")
        sys.exit(3)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
