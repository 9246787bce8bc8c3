"""Reader / replayer for the ABI traces written by oracle/trace_recorder.hpp.

A trace is what the reference-side shim told the C ABI about one `meep::fields` object built by
the UNMODIFIED reference (oracle/ref_driver.cpp --record): every array with its initial contents,
chunk / susceptibility / source / DFT descriptors, metal and connection pointer lists, the
per-step src_time samples, and -- as EXPECT records -- the arrays the reference's own CPU
stepping produced after the same number of steps.  Replaying it needs no reference code.

File format (little endian): magic "MCUTRC01", then records {u32 tag, u64 len, payload}.
Pointers are stored as (array id, element offset); id -1 == NULL.
"""
from __future__ import annotations

import ctypes as C
import gzip
import struct

import numpy as np

from meep_nl_b200 import _cabi as A

T_CONFIG, T_ARRAY, T_CHUNK, T_SUSC, T_SOURCE, T_DFT, T_METAL, T_CONN, T_FINALIZE, T_STEP, T_EXPECT, T_GYRO, T_MLATOM, T_NOISE = range(1, 15)


class Reader:
    def __init__(self, buf):
        self.b = buf
        self.o = 0

    def take(self, n):
        v = self.b[self.o:self.o + n]
        self.o += n
        return v

    def i32(self, n=1):
        v = np.frombuffer(self.take(4 * n), dtype=np.int32)
        return int(v[0]) if n == 1 else v.copy()

    def i64(self, n=1):
        v = np.frombuffer(self.take(8 * n), dtype=np.int64)
        return int(v[0]) if n == 1 else v.copy()

    def f64(self, n=1):
        v = np.frombuffer(self.take(8 * n), dtype=np.float64)
        return float(v[0]) if n == 1 else v.copy()


class Trace:
    """Parsed trace: self.arrays[id] are numpy arrays holding the INITIAL contents."""

    def __init__(self, path):
        opener = gzip.open if str(path).endswith(".gz") else open
        with opener(path, "rb") as fh:
            data = fh.read()
        assert data[:8] == b"MCUTRC01", "not a meepcu trace"
        self.records = []
        o = 8
        while o < len(data):
            tag, ln = struct.unpack_from("<IQ", data, o)
            o += 12
            self.records.append((tag, data[o:o + ln]))
            o += ln
        self.cfg = None
        self.arrays = {}
        self.expect = {}
        self.steps = []  # list of (nsteps, nsrc, bytes, noise of those steps or None)
        noise = None
        self.calls = []  # description calls in order
        for tag, pl in self.records:
            r = Reader(pl)
            if tag == T_CONFIG:
                # traces recorded with an older (shorter) meepcu_config: the appended fields are zero
                self.cfg = A.Config.from_buffer_copy(bytes(pl) + b"\0" * max(0, C.sizeof(A.Config) - len(pl)))
                self.rb = self.cfg.realnum_bytes
                self.rt = np.float64 if self.rb == 8 else np.float32
            elif tag == T_ARRAY:
                aid, eb = r.i32(), r.i32()
                n = r.i64()
                dt = {4: np.float32, 8: np.float64, 16: np.complex128}[eb] if eb != 8 else None
                if eb == 8:
                    dt = np.float64 if self.rb == 8 else np.complex64
                self.arrays[aid] = np.frombuffer(r.take(n * eb), dtype=dt).copy()
            elif tag == T_EXPECT:
                aid = r.i32()
                n = r.i64()
                self.expect[aid] = (n, pl[r.o:])
            elif tag == T_STEP:
                ns, nsrc = r.i32(), r.i32()
                self.steps.append((ns, nsrc, pl[r.o:], noise))
                noise = None
            elif tag == T_NOISE:  # unit Gaussians of the next T_STEP record (noisy Lorentzian media)
                ns = r.i32()
                npst = r.i64()
                noise = (npst, np.frombuffer(r.take(8 * ns * npst), dtype=np.float64).copy())
            else:
                self.calls.append((tag, pl))
        # expected arrays get the dtype of the array they shadow
        for aid, (n, raw) in list(self.expect.items()):
            self.expect[aid] = np.frombuffer(raw, dtype=self.arrays[aid].dtype).copy()

    @property
    def total_steps(self):
        return sum(s[0] for s in self.steps)

    def fresh_arrays(self):
        return {k: v.copy() for k, v in self.arrays.items()}


class Replay:
    """The trace's description issued against one backend; owns the host arrays it points to."""

    def __init__(self, trace: Trace, backend: A.Backend, nl_mode=None, device=0):
        self.tr = trace
        self.b = backend
        self.arr = trace.fresh_arrays()
        self.keep = []  # ctypes objects that must outlive the calls
        cfg = A.Config.from_buffer_copy(bytes(trace.cfg))
        cfg.abi_version = A.ABI_VERSION
        cfg.device = device
        if nl_mode is not None:
            cfg.nl_mode = nl_mode
        self.ctx = A.vp()
        self.b.check(self.b.create(C.byref(cfg), C.byref(self.ctx)))
        self.esz = trace.rb
        for tag, pl in trace.calls:
            getattr(self, "_t%d" % tag)(Reader(pl))

    # -- helpers
    def addr(self, aid, off=0):
        if aid < 0:
            return None
        a = self.arr[aid]
        return a.ctypes.data + off * a.itemsize

    def _ptr_table(self, r, n):
        pairs = r.i64(2 * n).reshape(n, 2) if n else np.zeros((0, 2), np.int64)
        out = (A.vp * max(n, 1))()
        for k in range(n):
            aid, off = int(pairs[k, 0]), int(pairs[k, 1])
            out[k] = self.arr[aid].ctypes.data + off * self.esz
        return out

    # -- record handlers
    def _t3(self, r):  # CHUNK
        i = r.i32()
        d = A.ChunkDesc()
        num, io, sigsize = r.i32(3), r.i32(3), r.i32(6)
        rank = r.i32()
        for k in range(3):
            d.num[k], d.io[k] = int(num[k]), int(io[k])
        for k in range(6):
            d.sigsize[k] = int(sigsize[k])
        d.rank = rank
        ids = r.i64(20 * 2 * 4 + 20 * 5 + 20 + 20 + 20 * 5 * 2 + 18)
        it = iter(ids.tolist())
        for name in ("f", "f_u", "f_w", "f_cond"):
            tab = getattr(d, name)
            for c in range(20):
                for k in range(2):
                    tab[c][k] = self.addr(next(it))
        for c in range(20):
            for k in range(5):
                d.chi1inv[c][k] = self.addr(next(it))
        for c in range(20):
            d.chi2[c] = self.addr(next(it))
        for c in range(20):
            d.chi3[c] = self.addr(next(it))
        for name in ("conductivity", "condinv"):
            tab = getattr(d, name)
            for c in range(20):
                for k in range(5):
                    tab[c][k] = self.addr(next(it))
        for name in ("sig", "kap", "siginv"):
            tab = getattr(d, name)
            for k in range(6):
                tab[k] = self.addr(next(it))
        if len(r.b) - r.o >= 8 * 40:  # ABI 2 traces: f_bfast appended
            it = iter(r.i64(40).tolist())
            for c in range(20):
                for k in range(2):
                    d.f_bfast[c][k] = self.addr(next(it))
        cid = C.c_int(-1)
        self.b.check(self.b.add_chunk(self.ctx, C.byref(d), C.byref(cid)))
        assert cid.value == i

    def _t4(self, r):  # SUSC
        i, ft, nod = r.i32(), r.i32(), r.i32()
        d = A.SusceptibilityDesc()
        d.ft, d.no_omega_0_denominator = ft, nod
        d.omega_0, d.gamma = r.f64(), r.f64()
        it = iter(r.i64(20 * 5 + 20 * 2 * 2).tolist())
        for c in range(20):
            for k in range(5):
                d.sigma[c][k] = self.addr(next(it))
        for name in ("P", "P_prev"):
            tab = getattr(d, name)
            for c in range(20):
                for k in range(2):
                    tab[c][k] = self.addr(next(it))
        if len(r.b) - r.o >= 8:  # appended: noisy Lorentzian
            d.noise_amp = r.f64()
        self.b.check(self.b.add_susceptibility(self.ctx, i, C.byref(d)))

    def _t13(self, r):  # MLATOM
        i, ft, L, T = r.i32(), r.i32(), r.i32(), r.i32()
        d = A.MultilevelDesc()
        d.ft, d.L, d.T = ft, L, T
        it = iter(r.i64(7 + 20 * 5 + 20 * 2 + 2 * 20 * 2 * T).tolist())
        for name in ("Gamma", "GammaInv", "alpha", "omega", "gamma", "sigmat", "N"):
            setattr(d, name, self.addr(next(it)))
        for c in range(20):
            for k in range(5):
                d.sigma[c][k] = self.addr(next(it))
        for c in range(20):
            for k in range(2):
                d.W_prev[c][k] = self.addr(next(it))
        for name in ("P", "P_prev"):
            tab = getattr(d, name)
            for c in range(20):
                for k in range(2):
                    ptrs = [self.addr(next(it)) for _ in range(T)]
                    if ptrs[0]:
                        arr = (A.vp * T)(*ptrs)
                        self.keep.append(arr)
                        tab[c][k] = C.cast(arr, A.vp)
        self.b.check(self.b.add_multilevel(self.ctx, i, C.byref(d)))

    def _t12(self, r):  # GYRO
        i, ft, model = r.i32(), r.i32(), r.i32()
        d = A.GyrotropicDesc()
        d.ft, d.model = ft, model
        d.omega_0, d.gamma, d.alpha = r.f64(), r.f64(), r.f64()
        gt = r.f64(9)
        for a in range(3):
            for b in range(3):
                d.gyro_tensor[a][b] = float(gt[3 * a + b])
        it = iter(r.i64(20 * 5 + 2 * 20 * 2 * 3).tolist())
        for c in range(20):
            for k in range(5):
                d.sigma[c][k] = self.addr(next(it))
        for name in ("P", "P_prev"):
            tab = getattr(d, name)
            for c in range(20):
                for k in range(2):
                    for dd in range(3):
                        tab[c][k][dd] = self.addr(next(it))
        self.b.check(self.b.add_gyrotropic(self.ctx, i, C.byref(d)))

    def _t5(self, r):  # SOURCE
        i = r.i32()
        d = A.SourceDesc()
        d.ft, d.c, d.src_time_id, d.is_integrated = r.i32(), r.i32(), r.i32(), r.i32()
        n = r.i64()
        idx = np.ascontiguousarray(r.i64(n) if n > 1 else np.array([r.i64()] if n else [], np.int64))
        amp = np.ascontiguousarray(r.f64(2 * n) if n else np.zeros(0))
        self.keep += [idx, amp]
        d.n = n
        d.index = idx.ctypes.data
        d.amp = amp.ctypes.data
        self.b.check(self.b.add_source(self.ctx, i, C.byref(d)))

    def _t6(self, r):  # DFT
        i = r.i32()
        d = A.DftDesc()
        d.c = r.i32()
        is_, ie = r.i32(3), r.i32(3)
        d.include_dV_and_interp_weights, d.sqrt_dV_and_interp_weights = r.i32(), r.i32()
        d.decimation_factor, d.Nomega = r.i32(), r.i32()
        s0, s1, e0, e1 = r.f64(3), r.f64(3), r.f64(3), r.f64(3)
        for k in range(3):
            d.is_[k], d.ie[k] = int(is_[k]), int(ie[k])
            d.s0[k], d.s1[k], d.e0[k], d.e1[k] = s0[k], s1[k], e0[k], e1[k]
        d.dV0, d.dV1 = r.f64(), r.f64()
        sc = r.f64(2)
        d.scale[0], d.scale[1] = sc[0], sc[1]
        d.avg1, d.avg2 = r.i64(), r.i64()
        om = np.ascontiguousarray(r.f64(d.Nomega) if d.Nomega > 1 else np.array([r.f64()]))
        self.keep.append(om)
        d.omega = om.ctypes.data
        aid = r.i64()
        d.dft = self.addr(aid)
        self.b.check(self.b.add_dft(self.ctx, i, C.byref(d), None))

    def _t7(self, r):  # METAL
        ft = r.i32()
        n = r.i64()
        tab = self._ptr_table(r, n)
        self.b.check(self.b.set_metal(self.ctx, ft, n, tab))

    def _t8(self, r):  # CONN
        ft, kind = r.i32(), r.i32()
        n = r.i64()
        dst = self._ptr_table(r, n)
        src = self._ptr_table(r, n)
        ph = None
        if kind == A.CONNECT_PHASE:
            raw = np.frombuffer(r.take((n // 2) * 2 * self.esz), dtype=self.tr.rt).copy()
            self.keep.append(raw)
            ph = raw.ctypes.data
        self.b.check(self.b.add_connections(self.ctx, ft, kind, n, dst, src, ph))

    def _t9(self, r):  # FINALIZE
        self.b.check(self.b.finalize(self.ctx))

    # -- running
    def run(self, max_steps=None):
        """Issue the recorded meepcu_step calls (optionally only the first max_steps steps)."""
        done = 0
        for ns, nsrc, raw, noise in self.tr.steps:
            if max_steps is not None:
                ns = min(ns, max_steps - done)
                if ns <= 0:
                    break
            if noise is not None:
                npst, g = noise
                g = np.ascontiguousarray(g[:ns * npst])
                self.b.check(self.b.set_noise(self.ctx, ns, npst, g.ctypes.data))
            buf = C.create_string_buffer(raw, len(raw)) if raw else None
            self.b.check(self.b.step(self.ctx, ns, nsrc, buf))
            done += ns
        return done

    def sync_to_host(self):
        self.b.check(self.b.sync_to_host(self.ctx))

    def stats(self):
        s = A.Stats()
        self.b.check(self.b.get_stats(self.ctx, C.byref(s)))
        return s

    def close(self):
        if self.ctx:
            self.b.destroy(self.ctx)
            self.ctx = None

    def compare_expected(self):
        """(max relative L2 error over arrays, total mismatching elements, worst array id)."""
        worst, bad, wid = 0.0, 0, -1
        num = den = 0.0
        for aid, exp in self.tr.expect.items():
            got = self.arr[aid]
            d = got.astype(np.complex128) - exp.astype(np.complex128)
            nb = int(np.count_nonzero(got != exp))
            bad += nb
            num += float(np.sum(np.abs(d) ** 2))
            den += float(np.sum(np.abs(exp.astype(np.complex128)) ** 2))
            if nb and wid < 0:
                wid = aid
        rel = (num / den) ** 0.5 if den > 0 else num ** 0.5
        return rel, bad, wid
