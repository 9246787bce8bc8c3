"""Host-side mirror of the slice of the reference's C++ API that sits above the hot path.

The reference builds its time-stepping state with `grid_volume` / `structure` / `fields`
(src/meep.hpp, src/structure.cpp, src/fields.cpp, src/boundaries.cpp); `fields::step()` then
runs on that state.  This module builds the SAME state (same array shapes, chunk division into
PML / non-PML regions, UPML sigma arrays, chunk connections, metal lists, sources) with numpy, for
machines where the reference itself is not available (the GPU box), and hands it to the C ABI
(include/meepcu.h).  Names and argument meanings follow the reference:

    gv = vol3d(25.6, 25.6, 25.6, 10)
    s  = structure(gv, eps=None, pml=1.0)            # structure.cpp:35-149
    f  = fields(s, backend)                          # fields.cpp:32-86
    f.use_real_fields(); f.add_point_source(Ez, 0.15, 0.1, 0.0, 4.0, gv.center())
    f.step(100); f.get_field(Ez, gv.center())

Only what the five BASELINE configs need is here (Cartesian 1/2/3-D, scalar eps, PML, metallic or
Bloch-periodic walls, Lorentzian susceptibility, chi2/chi3 arrays, point sources, DFT regions).
Setup is not on the hot path; stepping never touches Python between meepcu_step calls.
"""
from __future__ import annotations

import cmath
import ctypes as C
import math

import numpy as np

from . import _cabi as A
from ._cabi import (Ex, Ey, Ez, Hx, Hy, Hz, Dx, Dy, Dz, Bx, By, Bz, D1, D2, D3, X, Y, Z,
                    E_stuff, H_stuff, D_stuff, B_stuff, PE_stuff, PH_stuff)

pi = 3.141592653589793238462643383276  # src/meep.hpp:53

Metallic, Periodic = "Metallic", "Periodic"


def is_electric(c): return c < 5
def is_magnetic(c): return 5 <= c < 10
def is_D(c): return 10 <= c < 15
def is_B(c): return 15 <= c < 20
def comp_type(c): return c // 5
def comp_dir(c): return (X, Y, 3, 4, Z)[c % 5]
def dir_comp(ft, d): return 5 * ft + (0, 1, 4)[d]


class grid_volume:
    """The parts of reference grid_volume (src/meep/vec.hpp:1014-1178) the hot path needs."""

    def __init__(self, dim, a, num, io=(0, 0, 0)):
        self.dim, self.a = dim, float(a)
        self.has = [dim in (D2, D3), dim in (D2, D3), dim in (D1, D3)]
        self.num = [int(num[d]) if self.has[d] else 0 for d in range(3)]
        self.io = [int(io[d]) if self.has[d] else 0 for d in range(3)]
        nz, ny = self.num[2], self.num[1]
        self.stride = [0, 0, 0]  # set_strides, src/vec.cpp:482-494
        if self.has[2]: self.stride[2] = 1
        if self.has[1]: self.stride[1] = nz + 1
        if self.has[0]: self.stride[0] = (nz + 1) * (ny + 1)
        self.ntot = int(np.prod([self.num[d] + 1 for d in range(3) if self.has[d]]))
        self.shape = tuple(self.num[d] + 1 for d in range(3) if self.has[d])
        self.dirs = [d for d in range(3) if self.has[d]]

    def yee_shift(self, c, d):  # iyee_shift(c).in_direction(d)
        if not self.has[d]: return 0
        dc = comp_dir(c)
        if (is_electric(c) or is_D(c)) and d == dc: return 1
        if (is_magnetic(c) or is_B(c)) and d != dc: return 1
        return 0

    def has_field(self, c):
        if self.dim == D1: return c in (Ex, Hy, Dx, By)
        return comp_dir(c) < 3

    def owned_range(self, c, d):
        """inclusive index range of owned points (little_owned_corner0 .. big_corner)"""
        s = self.yee_shift(c, d)
        return 1 - s, self.num[d] - s

    def center(self):  # grid_volume::center() in physical units (io == 0 for user volumes)
        return tuple((self.io[d] / 2.0 + self.num[d] / 2.0) / self.a if self.has[d] else 0.0
                     for d in range(3))

    def cells(self):
        return int(np.prod([self.num[d] for d in self.dirs]))


def vol3d(xsize, ysize, zsize, a):  # src/vec.cpp:923-930
    return grid_volume(D3, a, (int(xsize * a + 0.5), int(ysize * a + 0.5), int(zsize * a + 0.5)))


def vol2d(xsize, ysize, a):
    return grid_volume(D2, a, (int(xsize * a + 0.5), int(ysize * a + 0.5), 0))


def vol1d(zsize, a):
    return grid_volume(D1, a, (0, 0, int(zsize * a + 0.5)))


class lorentzian_susceptibility:  # src/meep.hpp:246-278
    def __init__(self, omega_0, gamma, no_omega_0_denominator=False):
        self.omega_0, self.gamma, self.drude = omega_0, gamma, no_omega_0_denominator


class noisy_lorentzian_susceptibility(lorentzian_susceptibility):  # src/meep.hpp:282-299
    def __init__(self, noise_amp, omega_0, gamma, no_omega_0_denominator=False):
        super().__init__(omega_0, gamma, no_omega_0_denominator)
        self.noise_amp = noise_amp


class multilevel_susceptibility:  # src/meep.hpp:340-386, src/multilevel-atom.cpp:28-47
    """L levels, T radiative transitions: Gamma[L*L] relaxation rates, N0[L] initial populations, alpha[L*T]
    transition coefficients, omega[T], gamma[T], sigmat[5*T] (per-transition sigma factors by direction)."""

    def __init__(self, L, T, Gamma, N0, alpha, omega, gamma, sigmat):
        self.L, self.T = int(L), int(T)
        self.Gamma, self.N0, self.alpha = np.asarray(Gamma, float), np.asarray(N0, float), np.asarray(alpha, float)
        self.omega, self.gamma, self.sigmat = np.asarray(omega, float), np.asarray(gamma, float), np.asarray(sigmat, float)
        assert self.Gamma.size == L * L and self.N0.size == L and self.alpha.size == L * T
        assert self.omega.size == T and self.gamma.size == T and self.sigmat.size == 5 * T


class gaussian_src_time:
    """src/sources.cpp:72-108 (4-argument constructor), integrated by default"""

    def __init__(self, f, w, st, et):
        self.freq, self.width = f, w
        self.peak_time = 0.5 * (st + et)
        cutoff = (et - st) * 0.5
        while math.exp(-cutoff * cutoff / (2 * w * w)) < 1e-100:
            cutoff *= 0.9
        self.cutoff = float(np.float32(cutoff))
        self.is_integrated = True

    def dipole(self, time):
        tt = time - self.peak_time
        if float(np.float32(abs(tt))) > self.cutoff: return 0j
        amp = 1.0 / complex(0, -2 * pi * self.freq)
        return math.exp(-tt * tt / (2 * self.width * self.width)) * cmath.rect(1.0, -2 * pi * self.freq * tt) * amp

    def current(self, time, dt):  # src_time::current, src/meep.hpp:983-985
        return (self.dipole(time + dt) - self.dipole(time)) / dt


class structure_chunk:
    def __init__(self, gv, rt):
        self.gv = gv
        self.chi1inv = {}   # (c, d) -> array
        self.chi2, self.chi3 = {}, {}
        self.sig, self.kap, self.siginv = {}, {}, {}
        self.sigsize = [0] * 6
        self.pols = []      # [(susc, ft, {(c,d): sigma array})]


class structure:
    """structure::structure + choose_chunkdivision + use_pml (src/structure.cpp:35-149,509-523,
    630-691): PML regions are split off as their own chunks so that only they carry the
    auxiliary fields."""

    def __init__(self, gv, eps=None, pml=0.0, courant=0.5, realnum=np.float64, pml_dirs=None,
                 split=None):
        self.gv, self.user_volume = gv, gv
        self.a, self.Courant = gv.a, courant
        self.dt = courant / gv.a
        self.rt = np.dtype(realnum)
        self.pml_dx = float(pml)
        self.pml_dirs = [d for d in gv.dirs if (pml_dirs is None or d in pml_dirs)] if pml > 0 else []
        npml = int(self.pml_dx * gv.a + 1 + 0.5) if pml > 0 else 0  # structure.cpp:512
        # interval boundaries (in cells) per direction
        self.bounds = []
        for d in range(3):
            if not gv.has[d]:
                self.bounds.append([0, 0]); continue
            b = [0, gv.num[d]]
            if d in self.pml_dirs:
                if 2 * npml >= gv.num[d]:
                    raise ValueError("PML thicker than half the cell")
                b = [0, npml, gv.num[d] - npml, gv.num[d]]
            if split and d in split:  # extra cuts (multi-GPU slabs / tests): list of cell indices
                b = sorted(set(b) | set(int(x) for x in split[d]))
            self.bounds.append(b)
        self.chunks = []
        self.grid_index = {}
        for ix in range(max(len(self.bounds[0]) - 1, 1)):
            for iy in range(max(len(self.bounds[1]) - 1, 1)):
                for iz in range(max(len(self.bounds[2]) - 1, 1)):
                    k = (ix, iy, iz)
                    num, io = [0, 0, 0], [0, 0, 0]
                    for d in range(3):
                        if gv.has[d]:
                            num[d] = self.bounds[d][k[d] + 1] - self.bounds[d][k[d]]
                            io[d] = gv.io[d] + 2 * self.bounds[d][k[d]]
                    sc = structure_chunk(grid_volume(gv.dim, gv.a, num, io), self.rt)
                    self.grid_index[k] = len(self.chunks)
                    self.chunks.append(sc)
        for sc in self.chunks:
            for d in self.pml_dirs:
                for side in (0, 1):
                    self._use_pml(sc, d, side)
        if eps is not None:
            self.set_epsilon(eps)

    # structure_chunk::use_pml, src/structure.cpp:625-691 (quadratic profile, mean_stretch 1)
    def _use_pml(self, sc, d, side, Rasymptotic=1e-15):
        gv, a, dx = sc.gv, self.a, self.pml_dx
        uv = self.user_volume
        bloc = (uv.io[d] / 2.0 + (uv.num[d] if side else 0)) / a  # boundary_location
        prefac = (-math.log(Rasymptotic)) / (4 * dx * (1.0 / 3.0))

        def pml_x(i):
            here = i * 0.5 / a
            return 0.5 / a * (int(dx * (2 * a) + 0.5) - int(abs(bloc - here) * (2 * a) + 0.5))

        lo, hi = gv.io[d], gv.io[d] + 2 * gv.num[d] + 1
        if not any(pml_x(i) > 0 for i in range(lo, hi + 1)):
            return
        if sc.sigsize[d] <= 1:
            for dd in range(3):
                if dd not in sc.sig or dd == d:
                    n = 2 * gv.num[d] + 2 if dd == d else 1
                    sc.sigsize[dd] = n
                    sc.sig[dd] = np.zeros(n, self.rt)
                    sc.kap[dd] = np.ones(n, self.rt)
                    sc.siginv[dd] = np.ones(n, self.rt)
        for i in range(lo, hi + 1):
            x = pml_x(i)
            if x > 0:
                s = (x / dx) ** 2
                idx = i - lo
                sig = self.rt.type(0.5 * self.dt * prefac * s)
                kap = self.rt.type(1.0)
                sc.sig[d][idx] = sig
                sc.kap[d][idx] = kap
                sc.siginv[d][idx] = self.rt.type(1) / (kap + sig)

    def _yee_coords(self, sc, c):
        """physical coordinates (X,Y,Z arrays broadcastable to the chunk shape) of component c"""
        gv = sc.gv
        out = []
        for d in range(3):
            if gv.has[d]:
                h = gv.io[d] + 2 * np.arange(gv.num[d] + 1) + gv.yee_shift(c, d)
                sh = [1] * len(gv.dirs)
                sh[gv.dirs.index(d)] = -1
                out.append((h * (0.5 / gv.a)).reshape(sh))
            else:
                out.append(np.zeros([1] * len(gv.dirs)))
        return out

    def _eval(self, sc, c, fn):
        x, y, z = self._yee_coords(sc, c)
        v = np.asarray(fn(x, y, z), dtype=np.float64)
        return np.broadcast_to(v, sc.gv.shape).astype(np.float64)

    def set_epsilon(self, eps):
        """set_chi1inv without subpixel averaging (anisotropic_averaging.cpp:211-297 with
        use_anisotropic_averaging=false); trivial (all-one) arrays are elided like the reference"""
        for sc in self.chunks:
            for c in (Ex, Ey, Ez):
                if not sc.gv.has_field(c): continue
                v = self._eval(sc, c, eps)
                inv = (1.0 / v).astype(self.rt)
                if np.all(inv == 1): continue
                sc.chi1inv[(c, comp_dir(c))] = np.ascontiguousarray(inv).reshape(-1)
        # a chi1inv array present in one chunk is present in all (trivial_chi1inv is global)
        for c in (Ex, Ey, Ez):
            k = (c, comp_dir(c))
            if any(k in sc.chi1inv for sc in self.chunks):
                for sc in self.chunks:
                    if k not in sc.chi1inv and sc.gv.has_field(c):
                        sc.chi1inv[k] = np.ones(sc.gv.ntot, self.rt)

    def set_offdiag_flags(self, value=1e-30):
        """allocate the off-diagonal chi1inv rows (anisotropic_averaging.cpp:233-237) with a tiny
        value: in the fork they only act as flags selecting the NR branch"""
        for sc in self.chunks:
            for c in (Ex, Ey, Ez):
                if not sc.gv.has_field(c): continue
                for d in range(3):
                    if d != comp_dir(c) and sc.gv.has[d]:
                        sc.chi1inv[(c, d)] = np.full(sc.gv.ntot, value, self.rt)

    def set_chi2(self, fn, elide_trivial=False):  # structure_chunk::set_chi2, structure.cpp:832-866
        self._set_chi(fn, "chi2", elide_trivial)

    def set_chi3(self, fn, elide_trivial=False):
        self._set_chi(fn, "chi3", elide_trivial)

    def _set_chi(self, fn, which, elide_trivial=False):
        """elide_trivial: like the reference, a chunk where the function vanishes everywhere gets no array
        (structure.cpp:818-827,851-862) -- chunks of one structure then differ in which update they need"""
        for sc in self.chunks:
            for c in (Ex, Ey, Ez):
                if not sc.gv.has_field(c): continue
                v = np.ascontiguousarray(self._eval(sc, c, fn).astype(self.rt)).reshape(-1)
                if elide_trivial and not np.any(v):
                    getattr(sc, which).pop(c, None)
                    continue
                getattr(sc, which)[c] = v

    def add_susceptibility(self, sigma, ft, sus):  # anisotropic_averaging.cpp:300-370 (diagonal)
        assert ft == E_stuff
        for sc in self.chunks:
            arrs = {}
            for c in (Ex, Ey, Ez):
                if not sc.gv.has_field(c): continue
                arrs[(c, comp_dir(c))] = np.ascontiguousarray(self._eval(sc, c, sigma).astype(self.rt)).reshape(-1)
            sc.pols.append((sus, ft, arrs))


class _Ghost:
    """stand-in for an array that lives on another rank: same identity / aliasing behaviour as
    a numpy array in the connection builder, but no memory"""

    class _C:
        data = 0

    def __init__(self, n, itemsize):
        self.size, self.itemsize, self.nbytes = n, itemsize, 0
        self.ctypes = _Ghost._C()


def slab_decomposition(gv, nranks, d=0):
    """cuts (cells) that split direction d into nranks equal slabs, and the chunk -> rank map for a
    structure built with split={d: cuts} (the reference's split_by_cost with equal voxel counts,
    src/structure.cpp:66-94, restricted to one direction so that every halo is a contiguous plane)"""
    n = gv.num[d]
    cuts = [round(n * k / nranks) for k in range(1, nranks)]

    def rank_of(sc):
        mid = (sc.gv.io[d] - gv.io[d]) / 2.0 + sc.gv.num[d] / 2.0
        return min(int(mid * nranks / n), nranks - 1)
    return cuts, rank_of


class fields_chunk:
    def __init__(self, sc, is_real):
        self.s, self.gv = sc, sc.gv
        self.f, self.f_u, self.f_w, self.f_cond, self.f_bfast = {}, {}, {}, {}, {}
        self.f_w_prev = {}   # fields_chunk::f_w_prev (update_eh.cpp:221-226): only where a multilevel medium needs it
        self.pol_data = []   # per susceptibility: {"P": {(c,cmp): arr}, "P_prev": {...}}; multilevel: + "ml" (below)
        self.sources = []    # dicts
        self.dfts = []


class fields:
    """fields::fields (src/fields.cpp:32-86) + the lazy allocations of the first step done
    eagerly + connect_the_chunks (src/boundaries.cpp:347-623) + fields::step."""

    def __init__(self, s: structure, backend: A.Backend, nl_mode=A.NL_FORK_LITERAL, device=0,
                 rank=0, nranks=1, chunk_rank=None, comm_id=None):
        """rank / nranks / chunk_rank: the reference's MPI decomposition (structure_chunk::n_proc,
        src/structure.cpp:126-141): chunk i is time-stepped by rank chunk_rank[i]; arrays of other
        ranks' chunks are never allocated here, and connections that cross ranks become
        meepcu_add_remote_send / _recv lists (the MPI_Isend/Irecv pairs of src/mympi.cpp:107-135)."""
        self.rank, self.nranks, self.comm_id = rank, nranks, comm_id
        self.chunk_rank = list(chunk_rank) if chunk_rank is not None else [0] * len(s.chunks)
        self.s, self.gv, self.b = s, s.gv, backend
        self.a, self.dt = s.a, s.dt
        self.is_real = 0
        self.t = 0
        self.nl_mode, self.device = nl_mode, device
        self.bfast_scaled_k = [0.0, 0.0, 0.0]  # fields::bfast_scaled_k (step_db.cpp:65,129-143)
        self.synchronized_magnetic_fields = 0
        self.beta = 0.0  # fields::beta: 2-D exp(i beta z) dependence (step_db.cpp:148-175)
        self.pin_host = True  # the numpy arrays below outlive the context: sync by plain DMA (meepcu_config::pin_host)
        self.boundaries = {d: Metallic for d in range(3)}
        self.k = {d: 0.0 for d in range(3)}
        self.eikna = {d: 1 + 0j for d in range(3)}
        self.src_times = []
        self.chunks = [fields_chunk(sc, 0) for sc in s.chunks]
        self.ctx = None
        self._keep = []
        self._pending_sources = []

    # -- configuration (before the first step) ------------------------------------------------
    def use_real_fields(self):
        self.is_real = 1

    def use_bloch(self, d, kk):  # src/boundaries.cpp:88-103
        kk = complex(kk)
        if self.is_real and kk != 0.0:
            raise RuntimeError("Can't use real fields with bloch boundary conditions!")
        self.k[d] = kk
        self.boundaries[d] = Periodic
        n = self.gv.num[d]
        if kk.real * n == 0.5 * self.a:  # check b.z. edge exactly
            self.eikna[d] = -cmath.exp(-kk.imag * ((2 * pi / self.a) * n))
        else:
            self.eikna[d] = cmath.exp(1j * kk * ((2 * pi / self.a) * n))

    def add_point_source(self, c, freq, width, peaktime, cutoff, pt, amp=1.0):
        """fields::add_point_source (src/sources.cpp:193-213) -> add_volume_source with a zero-size
        volume: the amplitude is spread over the (up to 8) Yee points of c around pt with linear
        weights and scaled by a per zero-size direction (sources.cpp:474-477)."""
        width = width / freq
        cutoff = 1.0 / self.a + cutoff * width
        if peaktime <= 0.0: peaktime = self.t * self.dt + cutoff
        peaktime += (-self.dt * 0.5) if is_magnetic(c) else self.dt
        src = gaussian_src_time(freq, width, peaktime - cutoff, peaktime + cutoff)
        if is_magnetic(c): src.is_integrated = False
        self.src_times.append(src)
        self._pending_sources.append((c, len(self.src_times) - 1, pt, complex(amp)))

    def add_dft_fields(self, c, lo, hi, freqs, decimation_factor=1):
        """A DFT of component c over the box of its Yee indices lo..hi (cells, inclusive), without
        centred-grid averaging or volume weights (dft_chunk with avg1 = avg2 = 0,
        include_dV_and_interp_weights = false; src/dft.cpp:57-129)."""
        self._pending_dfts = getattr(self, "_pending_dfts", [])
        self._pending_dfts.append((c, tuple(lo), tuple(hi), np.asarray(freqs, np.float64), decimation_factor))

    # -- allocation ---------------------------------------------------------------------------
    def _alloc(self):
        rt = self.s.rt
        gvu = self.gv
        ncmp = 1 if self.is_real else 2
        comps = [c for c in range(20) if gvu.has_field(c) and comp_dir(c) < 3]
        if gvu.dim == D2:
            pass  # both TE and TM allocated (like aniso2d / 3-D); simple and always valid
        for ci, fc in enumerate(self.chunks):
            sc, gv = fc.s, fc.gv
            n = gv.ntot
            fc.mine = self.chunk_rank[ci] == self.rank
            zeros = (lambda n_, rt_: np.zeros(n_, rt_)) if fc.mine else (lambda n_, rt_: _Ghost(n_, np.dtype(rt_).itemsize))
            for cmp in range(ncmp):
                for c in comps:
                    if is_magnetic(c): continue
                    fc.f[(c, cmp)] = zeros(n, rt)
                for c in comps:  # H: alias of B unless mu / PML (fields.cpp:493-517, update_eh.cpp:204-209)
                    if not is_magnetic(c): continue
                    d = comp_dir(c)
                    bc = dir_comp(B_stuff, d)
                    fc.f[(c, cmp)] = zeros(n, rt) if sc.sigsize[d] > 1 else fc.f[(bc, cmp)]
                for c in comps:
                    d = comp_dir(c)
                    if is_D(c) or is_B(c):  # step_db.cpp:56-75
                        if sc.sigsize[(d + 2) % 3] > 1: fc.f_u[(c, cmp)] = zeros(n, rt)
                        if any(self.bfast_scaled_k): fc.f_bfast[(c, cmp)] = zeros(n, rt)  # step_db.cpp:76-79
                    else:                    # update_eh.cpp:199-216
                        if sc.sigsize[d] > 1: fc.f_w[(c, cmp)] = zeros(n, rt)
            for (sus, ft, arrs) in sc.pols:
                if isinstance(sus, multilevel_susceptibility):
                    # multilevel_data (multilevel-atom.cpp:118-181): GammaInv = inv(1 + Gamma dt/2), P / P_prev per
                    # transition, the populations N[i * L + level] = N0 everywhere; and W_prev for every E component
                    Pt = [{(c, cmp): zeros(n, rt) for (c, d) in arrs for cmp in range(ncmp)} for _ in range(sus.T)]
                    Ppt = [{(c, cmp): zeros(n, rt) for (c, d) in arrs for cmp in range(ncmp)} for _ in range(sus.T)]
                    ml = {"P": Pt, "P_prev": Ppt}
                    if fc.mine:
                        L = sus.L
                        G = (np.eye(L) + sus.Gamma.reshape(L, L) * float(np.asarray(self.dt, rt)) / 2).astype(rt)
                        ml["GammaInv"] = np.ascontiguousarray(np.linalg.inv(G.astype(np.float64)).astype(rt)).reshape(-1)
                        ml["N"] = np.ascontiguousarray(np.tile(sus.N0.astype(rt), n))
                        ml["small"] = {k: np.ascontiguousarray(getattr(sus, k).astype(rt)).reshape(-1)
                                       for k in ("Gamma", "alpha", "omega", "gamma", "sigmat")}
                        for c in comps:
                            if is_electric(c):
                                for cmp in range(ncmp):
                                    if (c, cmp) not in fc.f_w_prev: fc.f_w_prev[(c, cmp)] = zeros(n, rt)
                    fc.pol_data.append({"P": Pt[0], "P_prev": Ppt[0], "ml": ml})
                    continue
                P, Pp = {}, {}
                for (c, d) in arrs:
                    for cmp in range(ncmp):
                        P[(c, cmp)] = zeros(n, rt)
                        Pp[(c, cmp)] = zeros(n, rt)
                fc.pol_data.append({"P": P, "P_prev": Pp})

    # -- sources ------------------------------------------------------------------------------
    def _place_sources(self):
        gv = self.gv
        for (c, tid, pt, amp) in self._pending_sources:
            for d in gv.dirs:
                amp = amp * self.a  # delta function in every direction
            # linear interpolation weights on the Yee lattice of c
            per_dir = []
            for d in range(3):
                if not gv.has[d]:
                    per_dir.append([(0, 1.0)]); continue
                x = (pt[d] * 2 * self.a - gv.io[d] - gv.yee_shift(c, d)) / 2.0
                i0 = math.floor(x + 1e-9)
                w1 = x - i0
                if w1 < 1e-9: per_dir.append([(i0, 1.0)])
                else: per_dir.append([(i0, 1.0 - w1), (i0 + 1, w1)])
            ft = B_stuff if (is_magnetic(c) or is_B(c)) else D_stuff
            per_chunk = {}
            for (ix, wx) in per_dir[0]:
                for (iy, wy) in per_dir[1]:
                    for (iz, wz) in per_dir[2]:
                        g = (ix, iy, iz)  # global Yee index of component c
                        ci, idx = self._owner_index(c, g)
                        if ci is None or not self.chunks[ci].mine: continue
                        per_chunk.setdefault(ci, []).append((idx, amp * wx * wy * wz))
            for ci, pts in per_chunk.items():
                self.chunks[ci].sources.append(dict(ft=ft, c=c, tid=tid,
                                                    integrated=self.src_times[tid].is_integrated,
                                                    index=np.array([p[0] for p in pts], np.int64),
                                                    amp=np.array([[p[1].real, p[1].imag] for p in pts], np.float64)))

    def _owner_index(self, c, g):
        """chunk that owns global Yee index g of component c, and the local array index"""
        gv = self.gv
        k = []
        for d in range(3):
            if not gv.has[d]: k.append(0); continue
            h = 2 * g[d] + gv.yee_shift(c, d)  # half-pixel coordinate relative to io
            if not (0 < h <= 2 * gv.num[d]): return None, None
            b = self.s.bounds[d]
            k.append(int(np.searchsorted(2 * np.asarray(b), h, side="left")) - 1)
        ci = self.s.grid_index[tuple(k)]
        cg = self.chunks[ci].gv
        idx = 0
        for d in range(3):
            if gv.has[d]:
                idx += (g[d] - (cg.io[d] - gv.io[d]) // 2) * cg.stride[d]
        return ci, idx

    # -- connections and metal (boundaries.cpp:309-623) ------------------------------------------
    def _connect(self):
        gvu = self.gv
        rt = self.s.rt
        ncmp = 1 if self.is_real else 2
        conns = {}   # (ft, kind) -> [dst ptr arrays], [src ptr arrays], [phases]
        rsend, rrecv = {}, {}  # (ft, kind, peer) -> ([ptr arrays], [phases])
        metals = {ft: [] for ft in range(8)}
        state = {"i": 0, "j": 0}
        bnds2 = [2 * np.asarray(self.s.bounds[d]) for d in range(3)]
        kshape = [max(len(self.s.bounds[d]) - 1, 1) for d in range(3)]
        gi = np.zeros(kshape, np.int64)
        for k, v in self.s.grid_index.items(): gi[k] = v

        def add(ft, kind, dst, src, phase=None):
            mi, mj = self.chunks[state["i"]].mine, self.chunks[state["j"]].mine
            if mi and mj:
                e = conns.setdefault((ft, kind), ([], [], []))
                e[0].append(dst); e[1].append(src)
                if phase is not None: e[2].append(phase)
            elif mi:
                e = rrecv.setdefault((ft, kind, self.chunk_rank[state["j"]]), ([], []))
                e[0].append(dst)
                if phase is not None: e[1].append(phase)
            elif mj:
                e = rsend.setdefault((ft, kind, self.chunk_rank[state["i"]]), ([], []))
                e[0].append(src)

        for i, fc in enumerate(self.chunks):
            gv = fc.gv
            state["i"] = i
            for c in range(20):
                if (c, 0) not in fc.f: continue
                ft = comp_type(c)
                item = rt.itemsize
                lo = [gv.owned_range(c, d)[0] for d in range(3)]
                hi = [gv.owned_range(c, d)[1] for d in range(3)]
                # ---- metal: owned points on a Metallic high boundary (on_metal_boundary) ----
                for d in gv.dirs:
                    if self.boundaries[d] != Metallic: continue
                    if gv.io[d] + 2 * gv.num[d] != gvu.io[d] + 2 * gvu.num[d]: continue
                    if gv.yee_shift(c, d): continue
                    if not fc.mine: continue
                    idx = self._plane_indices(gv, lo, hi, d, gv.num[d])
                    for cmp in range(ncmp):
                        metals[ft].append(fc.f[(c, cmp)].ctypes.data + idx * item)
                # ---- not-owned points ----
                done = np.zeros(gv.shape, bool)
                for d in gv.dirs:
                    plane = 0 if gv.yee_shift(c, d) == 0 else gv.num[d]
                    rng_lo = [0 if gv.has[e] else 0 for e in range(3)]
                    rng_hi = [gv.num[e] for e in range(3)]
                    idx3 = self._plane_index_grids(gv, rng_lo, rng_hi, d, plane)
                    sel = ~done[tuple(idx3[e] for e in gv.dirs)]
                    done[tuple(idx3[e] for e in gv.dirs)] = True
                    pts = [g[sel] for g in idx3]
                    if pts[gv.dirs[0]].size == 0: continue
                    # global half-pixel coordinates relative to the user volume's corner
                    h = [gv.io[e] - gvu.io[e] + 2 * pts[e] + gv.yee_shift(c, e) if gv.has[e] else pts[e] * 0 for e in range(3)]
                    phase = np.ones(pts[gv.dirs[0]].size, complex)
                    ok = np.ones(pts[gv.dirs[0]].size, bool)
                    for e in gv.dirs:
                        L = 2 * gvu.num[e]
                        if self.boundaries[e] == Periodic:
                            low = h[e] <= 0
                            high = h[e] - L > 0
                            h[e] = np.where(low, h[e] + L, np.where(high, h[e] - L, h[e]))
                            phase = np.where(low, phase * np.conj(self.eikna[e]), np.where(high, phase * self.eikna[e], phase))
                        ok &= (h[e] > 0) & (h[e] <= L)
                        if self.boundaries[e] == Metallic:
                            ok &= ~(h[e] == L)  # on_metal_boundary (High side; Low is never owned)
                    if not ok.any(): continue
                    n_local = sum(pts[e][ok] * gv.stride[e] for e in gv.dirs)
                    kk = [np.searchsorted(bnds2[e], h[e][ok], side="left") - 1 if gv.has[e] else np.zeros(int(ok.sum()), np.int64) for e in range(3)]
                    owner = gi[kk[0], kk[1], kk[2]]
                    ph = phase[ok]
                    for j in np.unique(owner):
                        fj = self.chunks[int(j)]
                        state["j"] = int(j)
                        if not (fc.mine or fj.mine): continue
                        if (c, 0) not in fj.f: continue
                        selj = owner == j
                        gj = fj.gv
                        m = sum(((h[e][ok][selj] - gj.yee_shift(c, e) - (gj.io[e] - gvu.io[e])) // 2) * gj.stride[e] for e in gj.dirs)
                        n = n_local[selj]
                        # B is redundant between two chunks that both alias H to B (boundaries.cpp:357-364)
                        if is_B(c):
                            hc = dir_comp(H_stuff, comp_dir(c))
                            if fc.f[(hc, 0)] is fc.f[(c, 0)] and fj.f[(hc, 0)] is fj.f[(c, 0)]: continue
                        p = ph[selj]
                        if self.is_real or np.all(p == 1.0) or np.all(p == -1.0):
                            kind = A.CONNECT_NEGATE if np.all(p == -1.0) and not np.all(p == 1.0) else A.CONNECT_COPY
                            for cmp in range(ncmp):
                                add(ft, kind, fc.f[(c, cmp)].ctypes.data + n * item, fj.f[(c, cmp)].ctypes.data + m * item)
                        else:
                            dst = np.empty(2 * n.size, np.int64); srcp = np.empty(2 * n.size, np.int64)
                            dst[0::2] = fc.f[(c, 0)].ctypes.data + n * item
                            dst[1::2] = fc.f[(c, 1)].ctypes.data + n * item
                            srcp[0::2] = fj.f[(c, 0)].ctypes.data + m * item
                            srcp[1::2] = fj.f[(c, 1)].ctypes.data + m * item
                            pc = np.empty(2 * n.size, rt); pc[0::2] = p.real; pc[1::2] = p.imag
                            add(ft, A.CONNECT_PHASE, dst, srcp, pc)
                        # polarisations travel with E/H (boundaries.cpp:490-531), always as COPY/phase of P
                        if is_electric(c) or is_magnetic(c):
                            ftp = PE_stuff if is_electric(c) else PH_stuff
                            for ip in range(len(fc.pol_data)):
                                if ip >= len(fj.pol_data): break
                                # (a multilevel medium: one P per transition, cinternal_notowned_ptr of multilevel-atom.cpp:228-238)
                                pi_ = fc.pol_data[ip]["ml"]["P"] if "ml" in fc.pol_data[ip] else [fc.pol_data[ip]["P"]]
                                pj_ = fj.pol_data[ip]["ml"]["P"] if "ml" in fj.pol_data[ip] else [fj.pol_data[ip]["P"]]
                                for Pi, Pj in zip(pi_, pj_):
                                    for cmp in range(ncmp):
                                        if (c, cmp) in Pi and (c, cmp) in Pj and (self.is_real or np.all(p == 1.0)):
                                            add(ftp, A.CONNECT_COPY, Pi[(c, cmp)].ctypes.data + n * item,
                                                Pj[(c, cmp)].ctypes.data + m * item)
        self._rsend, self._rrecv = rsend, rrecv
        return conns, metals

    @staticmethod
    def _plane_index_grids(gv, lo, hi, d, plane):
        axes = []
        for e in range(3):
            if not gv.has[e]: axes.append(np.zeros(1, np.int64))
            elif e == d: axes.append(np.array([plane], np.int64))
            else: axes.append(np.arange(lo[e], hi[e] + 1, dtype=np.int64))
        g = np.meshgrid(*axes, indexing="ij")
        return [x.reshape(-1) for x in g]

    def _plane_indices(self, gv, lo, hi, d, plane):
        g = self._plane_index_grids(gv, lo, hi, d, plane)
        return sum(g[e] * gv.stride[e] for e in gv.dirs)

    # -- attach to the backend -----------------------------------------------------------------
    def _ptr(self, arr):
        return arr.ctypes.data if arr is not None else None

    def init(self):
        if self.ctx is not None: return
        b = self.b
        self._alloc()
        self._place_sources()
        cfg = A.Config()
        cfg.abi_version = A.ABI_VERSION
        cfg.realnum_bytes = self.s.rt.itemsize
        cfg.is_real = self.is_real
        cfg.dim = self.gv.dim
        cfg.a, cfg.dt, cfg.courant = self.a, self.dt, self.s.Courant
        cfg.beta = self.beta if self.gv.dim == D2 else 0.0
        cfg.nl_mode, cfg.device, cfg.t = self.nl_mode, self.device, self.t
        cfg.pin_host = 1 if self.pin_host else 0
        for k in range(3): cfg.bfast_scaled_k[k] = float(self.bfast_scaled_k[k])
        ctx = A.vp()
        b.check(b.create(C.byref(cfg), C.byref(ctx)))
        self.ctx = ctx
        ncmp = 1 if self.is_real else 2
        if self.nranks > 1:
            if self.comm_id is None: raise ValueError("comm_id (meepcu_comm_unique_id) required for nranks > 1")
            idbuf = (C.c_char * 128).from_buffer_copy(bytes(self.comm_id))
            b.check(b.comm_init(ctx, self.rank, self.nranks, idbuf))
        self.local_id = {}
        for gi_, fc in enumerate(self.chunks):
            if not fc.mine: continue
            i = len(self.local_id)
            self.local_id[gi_] = i
            sc, gv = fc.s, fc.gv
            d = A.ChunkDesc()
            for k in range(3):
                d.num[k], d.io[k] = gv.num[k], gv.io[k]
            for (c, cmp), arr in fc.f.items(): d.f[c][cmp] = self._ptr(arr)
            for (c, cmp), arr in fc.f_u.items(): d.f_u[c][cmp] = self._ptr(arr)
            for (c, cmp), arr in fc.f_bfast.items(): d.f_bfast[c][cmp] = self._ptr(arr)
            for (c, cmp), arr in fc.f_w.items(): d.f_w[c][cmp] = self._ptr(arr)
            for (c, dd), arr in sc.chi1inv.items(): d.chi1inv[c][dd] = self._ptr(arr)
            for c, arr in sc.chi2.items(): d.chi2[c] = self._ptr(arr)
            for c, arr in sc.chi3.items(): d.chi3[c] = self._ptr(arr)
            for dd in range(3):
                d.sigsize[dd] = sc.sigsize[dd]
                if dd in sc.sig:
                    d.sig[dd], d.kap[dd], d.siginv[dd] = self._ptr(sc.sig[dd]), self._ptr(sc.kap[dd]), self._ptr(sc.siginv[dd])
            d.rank = self.rank
            cid = C.c_int(-1)
            b.check(b.add_chunk(ctx, C.byref(d), C.byref(cid)))
            assert cid.value == i
            for ip, (sus, ft, arrs) in enumerate(sc.pols):
                if isinstance(sus, multilevel_susceptibility):
                    ml = fc.pol_data[ip]["ml"]
                    md = A.MultilevelDesc()
                    md.ft, md.L, md.T = ft, sus.L, sus.T
                    md.Gamma, md.GammaInv = self._ptr(ml["small"]["Gamma"]), self._ptr(ml["GammaInv"])
                    md.alpha, md.omega = self._ptr(ml["small"]["alpha"]), self._ptr(ml["small"]["omega"])
                    md.gamma, md.sigmat = self._ptr(ml["small"]["gamma"]), self._ptr(ml["small"]["sigmat"])
                    md.N = self._ptr(ml["N"])
                    for (c, dd), arr in arrs.items(): md.sigma[c][dd] = self._ptr(arr)
                    for name in ("P", "P_prev"):
                        tab = getattr(md, name)
                        for (c, cmp) in ml[name][0]:
                            ptrs = (A.vp * sus.T)(*[ml[name][t][(c, cmp)].ctypes.data for t in range(sus.T)])
                            self._keep.append(ptrs)
                            tab[c][cmp] = C.cast(ptrs, A.vp)
                    for (c, cmp), arr in fc.f_w_prev.items(): md.W_prev[c][cmp] = self._ptr(arr)
                    b.check(b.add_multilevel(ctx, i, C.byref(md)))
                    continue
                sd = A.SusceptibilityDesc()
                sd.ft, sd.omega_0, sd.gamma = ft, sus.omega_0, sus.gamma
                sd.no_omega_0_denominator = int(sus.drude)
                sd.noise_amp = float(getattr(sus, "noise_amp", 0.0))
                for (c, dd), arr in arrs.items(): sd.sigma[c][dd] = self._ptr(arr)
                for (c, cmp), arr in fc.pol_data[ip]["P"].items(): sd.P[c][cmp] = self._ptr(arr)
                for (c, cmp), arr in fc.pol_data[ip]["P_prev"].items(): sd.P_prev[c][cmp] = self._ptr(arr)
                b.check(b.add_susceptibility(ctx, i, C.byref(sd)))
            for s in fc.sources:
                sd = A.SourceDesc()
                sd.ft, sd.c, sd.src_time_id, sd.is_integrated = s["ft"], s["c"], s["tid"], int(s["integrated"])
                sd.n = s["index"].size
                sd.index, sd.amp = s["index"].ctypes.data, s["amp"].ctypes.data
                b.check(b.add_source(ctx, i, C.byref(sd)))
        self._add_dfts()
        conns, metals = self._connect()
        for ft, lst in metals.items():
            if not lst: continue
            p = np.ascontiguousarray(np.unique(np.concatenate(lst).astype(np.uint64)))
            b.check(b.set_metal(ctx, ft, p.size, p.ctypes.data))
        for (ft, kind), (dst, src, ph) in conns.items():
            dp = np.ascontiguousarray(np.concatenate([np.asarray(x).reshape(-1) for x in dst]).astype(np.uint64))
            sp = np.ascontiguousarray(np.concatenate([np.asarray(x).reshape(-1) for x in src]).astype(np.uint64))
            php = np.ascontiguousarray(np.concatenate(ph)) if ph else None
            b.check(b.add_connections(ctx, ft, kind, dp.size, dp.ctypes.data, sp.ctypes.data,
                                      php.ctypes.data if php is not None else None))
        for (ft, kind, peer), (ptrs, ph) in self._rsend.items():
            pp = np.ascontiguousarray(np.concatenate([np.asarray(x).reshape(-1) for x in ptrs]).astype(np.uint64))
            b.check(b.add_remote_send(ctx, ft, kind, peer, pp.size, pp.ctypes.data))
        for (ft, kind, peer), (ptrs, ph) in self._rrecv.items():
            pp = np.ascontiguousarray(np.concatenate([np.asarray(x).reshape(-1) for x in ptrs]).astype(np.uint64))
            php = np.ascontiguousarray(np.concatenate(ph)) if ph else None
            b.check(b.add_remote_recv(ctx, ft, kind, peer, pp.size, pp.ctypes.data,
                                      php.ctypes.data if php is not None else None))
        b.check(b.finalize(ctx))
        nn = C.c_size_t(0)
        b.check(b.get_noise_count(ctx, C.byref(nn)))
        self.n_noise = nn.value

    def _noise(self, nsteps):
        """unit Gaussians for noisy Lorentzian media in the order include/meepcu.h states: per step, update_pols(H_stuff)
        then update_pols(E_stuff); in each, this rank's chunks in index order, their susceptibilities, components, cmp,
        owned points.  Drawn per (time step, field type, GLOBAL chunk), so that a point sees the same value whatever
        the rank decomposition (the reference's MPI ranks each have their own generator: there the decomposition
        changes the noise; here it is the one property the N-rank == 1-rank tests need)."""
        ncmp = 1 if self.is_real else 2
        out = []
        for s in range(nsteps):
            for ft in (H_stuff, E_stuff):
                for gi_, fc in enumerate(self.chunks):
                    if not fc.mine: continue
                    cnt = 0
                    for ip, (sus, sft, arrs) in enumerate(fc.s.pols):
                        if sft != ft or not getattr(sus, "noise_amp", 0.0) or isinstance(sus, multilevel_susceptibility): continue
                        for c in range(20):
                            for cmp in range(ncmp):
                                if (c, cmp) in fc.pol_data[ip]["P"] and (c, comp_dir(c)) in arrs:
                                    cnt += int(np.prod([fc.gv.num[d] for d in fc.gv.dirs]))
                    if cnt:
                        rng = np.random.default_rng([getattr(self, "noise_seed", 1234), self.t + s, ft, gi_])
                        out.append(rng.standard_normal(cnt))
        g = np.ascontiguousarray(np.concatenate(out)) if out else np.zeros(0)
        assert g.size == nsteps * self.n_noise, (g.size, nsteps, self.n_noise)
        return g

    def _add_dfts(self):
        b, ctx = self.b, self.ctx
        rt = self.s.rt
        ct = np.complex128 if rt.itemsize == 8 else np.complex64
        for mon, (c, lo, hi, freqs, decim) in enumerate(getattr(self, "_pending_dfts", [])):
            for gi_, fc in enumerate(self.chunks):
                gv = fc.gv
                if (c, 0) not in fc.f or not fc.mine: continue
                i = self.local_id[gi_]
                is_, ie, ok = [0, 0, 0], [0, 0, 0], True
                for d in range(3):
                    if not gv.has[d]: continue
                    olo, ohi = gv.owned_range(c, d)
                    off = (gv.io[d] - self.gv.io[d]) // 2
                    a0, a1 = max(lo[d] - off, olo), min(hi[d] - off, ohi)
                    if a0 > a1: ok = False; break
                    is_[d] = gv.io[d] + 2 * a0 + gv.yee_shift(c, d)
                    ie[d] = gv.io[d] + 2 * a1 + gv.yee_shift(c, d)
                if not ok: continue
                N = int(np.prod([(ie[d] - is_[d]) // 2 + 1 for d in gv.dirs]))
                dd = A.DftDesc()
                dd.c = c
                for d in range(3):
                    dd.is_[d], dd.ie[d] = is_[d], ie[d]
                    dd.s0[d] = dd.s1[d] = dd.e0[d] = dd.e1[d] = 1.0
                dd.dV0, dd.dV1 = 1.0, 0.0
                dd.avg1 = dd.avg2 = 0
                dd.include_dV_and_interp_weights = 0
                dd.sqrt_dV_and_interp_weights = 0
                dd.decimation_factor = decim
                om = np.ascontiguousarray(2 * pi * freqs)
                dd.Nomega = om.size
                dd.omega = om.ctypes.data
                dd.scale[0], dd.scale[1] = self.dt / math.sqrt(2 * pi) * decim, 0.0  # dft.cpp:100
                arr = np.zeros(N * om.size, ct)
                dd.dft = arr.ctypes.data
                self._keep += [om]
                fc.dfts.append(dict(c=c, dft=arr, N=N, nomega=om.size, mon=mon))
                b.check(b.add_dft(ctx, i, C.byref(dd), None))

    # -- stepping -------------------------------------------------------------------------------
    def samples(self, nsteps):
        nsrc = len(self.src_times)
        buf = (A.SrcSample * max(nsteps * nsrc, 1))()
        for s in range(nsteps):
            tm = (self.t + s) * self.dt
            for j, st in enumerate(self.src_times):
                o = buf[s * nsrc + j]
                cb = st.current(tm, self.dt)
                dh = st.dipole(tm + 0.5 * self.dt)
                cd = st.current(tm + 0.5 * self.dt, self.dt)
                de = st.dipole(tm + self.dt)
                o.current_B[0], o.current_B[1] = cb.real, cb.imag
                o.dipole_H[0], o.dipole_H[1] = dh.real, dh.imag
                o.current_D[0], o.current_D[1] = cd.real, cd.imag
                o.dipole_E[0], o.dipole_E[1] = de.real, de.imag
        return buf, nsrc

    def step(self, n=1, samples=None):
        """n calls of fields::step() (src/step.cpp:35-140) without leaving the device"""
        self.init()
        saved = self.synchronized_magnetic_fields  # src/step.cpp:36-41,128-134: restore, step, synchronize again
        if saved:
            self.synchronized_magnetic_fields = 1
            self.restore_magnetic_fields()
        buf, nsrc = samples if samples is not None else self.samples(n)
        if self.n_noise:
            g = self._noise(n)
            self.b.check(self.b.set_noise(self.ctx, n, self.n_noise, g.ctypes.data))
        self.b.check(self.b.step(self.ctx, n, nsrc, buf if nsrc else None))
        self.t += n
        if saved:
            self.synchronize_magnetic_fields()
            self.synchronized_magnetic_fields = saved

    def synchronize_magnetic_fields(self):
        """fields::synchronize_magnetic_fields (src/energy_and_flux.cpp:145-166), reference-counted the
        same way: B and H are advanced half a step and averaged with their backups on the device"""
        self.init()
        self.synchronized_magnetic_fields += 1
        if self.synchronized_magnetic_fields > 1: return
        buf, nsrc = self.samples(1)  # current_B at time(), dipole_H at time() + dt/2
        self.b.check(self.b.synchronize_magnetic_fields(self.ctx, nsrc, buf if nsrc else None))

    def restore_magnetic_fields(self):  # src/energy_and_flux.cpp:168-177
        if not self.synchronized_magnetic_fields: return
        self.synchronized_magnetic_fields -= 1
        if self.synchronized_magnetic_fields: return
        self.b.check(self.b.restore_magnetic_fields(self.ctx))

    def sync_to_host(self):
        self.b.check(self.b.sync_to_host(self.ctx))

    def sync_from_host(self):
        self.b.check(self.b.sync_from_host(self.ctx))

    def get_field(self, c, pt, cmp=0):
        """value of component c at the Yee point nearest pt, read from the device mirror"""
        gv = self.gv
        g = tuple(int(round((pt[d] * 2 * self.a - gv.io[d] - gv.yee_shift(c, d)) / 2.0)) if gv.has[d] else 0 for d in range(3))
        ci, idx = self._owner_index(c, g)
        if ci is None: raise ValueError("point outside the cell")
        if not self.chunks[ci].mine: return None
        arr = self.chunks[ci].f[(c, cmp)]
        out = C.c_double()
        self.b.check(self.b.read_value(self.ctx, arr.ctypes.data + idx * arr.itemsize, C.byref(out)))
        return out.value

    def sync_subset(self, to_host=True, groups=A.SYNC_FIELDS):
        """meepcu_sync_subset: only the named array groups travel (A.SYNC_FIELDS | SYNC_AUX | SYNC_POL |
        SYNC_DFT | SYNC_MATERIALS)"""
        self.b.check(self.b.sync_subset(self.ctx, 1 if to_host else 0, groups))

    def sync_dft_to_host(self):
        """all that dft_flux::flux / dft_energy / near2far read (src/dft.cpp:533-548,630-700)"""
        self.sync_subset(True, A.SYNC_DFT)

    def dft_dot(self, mon_a, mon_b):
        """out[i] = sum over the chunks and points of monitors mon_a, mon_b (indices in add_dft_fields order; same box,
        components on the same Yee grid) of Re(a.dft * conj(b.dft)), summed on the device (meepcu_dft_dot): the loop of
        dft_flux::flux / dft_energy::electric, magnetic (src/dft.cpp:533-548,657-687) without fetching an accumulator"""
        self.init()
        pa, pb, nom = [], [], None
        for fc in self.chunks:
            if not fc.mine: continue
            da = [d for d in fc.dfts if d["mon"] == mon_a]
            db = [d for d in fc.dfts if d["mon"] == mon_b]
            if not da and not db: continue
            assert len(da) == 1 and len(db) == 1 and da[0]["N"] == db[0]["N"], "monitors do not cover the same points"
            pa.append(da[0]["dft"].ctypes.data); pb.append(db[0]["dft"].ctypes.data)
            nom = da[0]["nomega"]
        out = np.zeros(nom or 0, np.float64)
        if pa:
            a = np.asarray(pa, np.uint64); b = np.asarray(pb, np.uint64)
            self.b.check(self.b.dft_dot(self.ctx, len(pa), a.ctypes.data, b.ctypes.data, nom,
                                        out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def integrate_dot(self, c0, c1, lo, hi):
        """sum over the Yee indices lo..hi (inclusive, global cells; c0 and c1 on the same Yee grid, e.g. Ex and Dx) of
        Re(conj(f_c0) f_c1) with unit weights, and max |.|, reduced on the device (meepcu_integrate_dot): the loop of
        fields::integrate(2, cs, dot_integrand, ...) (src/integrate.cpp:47-131) with the weights of an interior box"""
        self.init()
        gv = self.gv
        pieces = []
        for fc in self.chunks:
            if not fc.mine or (c0, 0) not in fc.f or (c1, 0) not in fc.f: continue
            cg = fc.gv
            d = A.IntegrateDesc()
            ok = True
            for k in range(3):
                d.s0[k] = d.s1[k] = d.e0[k] = d.e1[k] = 1.0
                if not gv.has[k]: continue
                off = (cg.io[k] - gv.io[k]) // 2
                o0, o1 = cg.owned_range(c0, k)
                a, b = max(int(lo[k]), off + o0), min(int(hi[k]), off + o1)
                if a > b: ok = False; break
                d.is_[k] = cg.io[k] + 2 * (a - off) + cg.yee_shift(c0, k)
                d.ie[k] = cg.io[k] + 2 * (b - off) + cg.yee_shift(c0, k)
            if not ok: continue
            for i, c in enumerate((c0, c1)):
                d.f[i][0] = fc.f[(c, 0)].ctypes.data
                d.f[i][1] = fc.f[(c, 1)].ctypes.data if (c, 1) in fc.f else None
                d.ph[i][0], d.ph[i][1] = 1.0, 0.0
            d.dV0, d.dV1 = 1.0, 0.0
            pieces.append(d)
        s_, m_ = C.c_double(0), C.c_double(0)
        if pieces:
            arr = (A.IntegrateDesc * len(pieces))(*pieces)
            self.b.check(self.b.integrate_dot(self.ctx, len(pieces), arr, C.byref(s_), C.byref(m_)))
        return s_.value, m_.value

    def get_array_slice(self, c, lo, hi, cmp=0):
        """fields::get_array_slice without interpolation (src/array_slice.cpp): the values of component c
        at its Yee indices lo..hi (inclusive, global cells), gathered on the device by meepcu_read_values
        -- no sync of whole arrays.  Points that no local chunk owns come back as NaN."""
        self.init()
        gv = self.gv
        lo = [int(lo[d]) if gv.has[d] else 0 for d in range(3)]
        hi = [int(hi[d]) if gv.has[d] else 0 for d in range(3)]
        shape = [hi[d] - lo[d] + 1 for d in range(3)]
        out = np.full(shape, np.nan)
        for fc in self.chunks:
            if not fc.mine or (c, cmp) not in fc.f: continue
            cg = fc.gv
            sl, rng = [], []
            for d in range(3):
                if not gv.has[d]:
                    sl.append(slice(0, 1)); rng.append(np.zeros(1, np.int64)); continue
                off = (cg.io[d] - gv.io[d]) // 2            # global index of this chunk's local index 0
                o0, o1 = cg.owned_range(c, d)
                a, b = max(lo[d], off + o0), min(hi[d], off + o1)
                if a > b: break
                sl.append(slice(a - lo[d], b - lo[d] + 1))
                rng.append(np.arange(a - off, b - off + 1, dtype=np.int64) * cg.stride[d])
            else:
                idx = (rng[0][:, None, None] + rng[1][None, :, None] + rng[2][None, None, :]).reshape(-1)
                arr = fc.f[(c, cmp)]
                ptrs = np.ascontiguousarray((arr.ctypes.data + idx * arr.itemsize).astype(np.uint64))
                vals = np.empty(ptrs.size, np.float64)
                self.b.check(self.b.read_values(self.ctx, ptrs.size, ptrs.ctypes.data,
                                                vals.ctypes.data_as(C.POINTER(C.c_double))))
                out[sl[0], sl[1], sl[2]] = vals.reshape([s_.stop - s_.start for s_ in sl])
        return out

    def stats(self):
        s = A.Stats()
        self.b.check(self.b.get_stats(self.ctx, C.byref(s)))
        return s

    def finite(self):
        ok = C.c_int(1)
        self.b.check(self.b.check_finite(self.ctx, C.byref(ok)))
        return bool(ok.value)

    def host_bytes(self, mutable_only=True):
        n = 0
        seen = set()
        for fc in self.chunks:
            if not fc.mine: continue
            groups = [fc.f, fc.f_u, fc.f_w, fc.f_bfast]
            for g in groups:
                for arr in g.values():
                    if id(arr) in seen: continue
                    seen.add(id(arr)); n += arr.nbytes
            for pd in fc.pol_data:
                for arr in list(pd["P"].values()) + list(pd["P_prev"].values()): n += arr.nbytes
                if "ml" in pd:
                    n += pd["ml"]["N"].nbytes
                    for t in range(1, len(pd["ml"]["P"])):
                        for arr in list(pd["ml"]["P"][t].values()) + list(pd["ml"]["P_prev"][t].values()): n += arr.nbytes
            for arr in fc.f_w_prev.values(): n += arr.nbytes
            for dft in fc.dfts: n += dft["dft"].nbytes
            if not mutable_only:  # what meepcu_sync_from_host uploads on top: materials, sigma, UPML profiles
                sc = fc.s
                for arr in list(sc.chi1inv.values()) + list(sc.chi2.values()) + list(sc.chi3.values()) + \
                        list(sc.sig.values()) + list(sc.kap.values()) + list(sc.siginv.values()):
                    if id(arr) in seen: continue
                    seen.add(id(arr)); n += arr.nbytes
                for (_sus, _ft, arrs) in sc.pols:
                    for arr in arrs.values():
                        if id(arr) in seen: continue
                        seen.add(id(arr)); n += arr.nbytes
        return n

    def owned_arrays(self):
        """every mutable array of the chunks this rank steps, as (name, ndarray) in a deterministic order that
        does not depend on the rank decomposition (names carry the GLOBAL chunk index): what the N-rank ==
        1-rank comparisons hash (the reference's own invariant, tests/three_d.cpp:95-229)"""
        out = []
        for gi_, fc in enumerate(self.chunks):
            if not fc.mine: continue
            seen = set()
            for kind, g in (("f", fc.f), ("f_u", fc.f_u), ("f_w", fc.f_w), ("f_bfast", fc.f_bfast)):
                for (c, cmp) in sorted(g):
                    arr = g[(c, cmp)]
                    if id(arr) in seen: continue  # H aliasing B
                    seen.add(id(arr))
                    out.append(("chunk%d.%s[%d][%d]" % (gi_, kind, c, cmp), arr))
            for (c, cmp) in sorted(fc.f_w_prev):
                out.append(("chunk%d.f_w_prev[%d][%d]" % (gi_, c, cmp), fc.f_w_prev[(c, cmp)]))
            for ip, pd in enumerate(fc.pol_data):
                for kind in ("P", "P_prev"):
                    for (c, cmp) in sorted(pd[kind]):
                        out.append(("chunk%d.pol%d.%s[%d][%d]" % (gi_, ip, kind, c, cmp), pd[kind][(c, cmp)]))
                if "ml" in pd:
                    out.append(("chunk%d.pol%d.N" % (gi_, ip), pd["ml"]["N"]))
                    for t in range(1, len(pd["ml"]["P"])):
                        for kind in ("P", "P_prev"):
                            for (c, cmp) in sorted(pd["ml"][kind][t]):
                                out.append(("chunk%d.pol%d.%s[%d][%d][t%d]" % (gi_, ip, kind, c, cmp, t), pd["ml"][kind][t][(c, cmp)]))
            for k, d in enumerate(fc.dfts):
                out.append(("chunk%d.dft%d[%d]" % (gi_, k, d["c"]), d["dft"]))
        return out

    def close(self):
        if self.ctx is not None:
            self.b.destroy(self.ctx)
            self.ctx = None

    # the context page-locks the numpy arrays in place (pin_host) and keeps their addresses: it must go before they do
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
