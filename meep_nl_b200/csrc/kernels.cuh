// kernels.cuh -- sm_100a kernels of the per-timestep field update.
//
// Reference semantics restated here (each kernel cites the lines it replaces):
//   k_mix / k_curl3d_v / k_curl3d / k_halfg_v / k_halfg (one fused pass per half step):
//                 fields_chunk::step_db + step_curl        (src/step_db.cpp:44-146, step_generic.cpp:69-253)
//                 step_bfast, step_beta                    (src/step_generic.cpp:259-537, step_db.cpp:129-175)
//                 fields_chunk::step_source                (src/step.cpp:296-319)
//                 fields_chunk::zero_metal                 (src/boundaries.cpp:304-307)
//                 fields_chunk::update_eh + step_update_EDHB, diagonal branches
//                                                          (src/update_eh.cpp:67-283, step_generic.cpp:576-906)
//                 lorentzian_susceptibility::subtract_P / update_P, isotropic
//                                                          (src/susceptibility.cpp:188-281)
//                 so that every array element moves once per half step (DESIGN.md "bytes per cell").
//   k_eh_general: step_update_EDHB branches that read neighbours of D-P: the fork's Newton-Raphson
//                 chi2 solve (step_generic.cpp:730-816, newton_raphson.cpp:93-359) and upstream's
//                 Pade/off-diagonal formulas (UPSTREAM_PADE mode)
//   k_pol_general: update_P with an anisotropic sigma tensor (src/susceptibility.cpp:225-251)
//   k_fmp       : f_minus_p = D - P - dipoles             (src/update_eh.cpp:122-146)
//   k_conn_*, k_pack, k_unpack : fields::step_boundaries   (src/step.cpp:173-288)
//   k_dft       : dft_chunk::update_dft                    (src/dft.cpp:265-307)
//   k_average   : fields_chunk::average_with_backup        (src/energy_and_flux.cpp:135-143)
//
// Arithmetic is written with the reference's operand order and compiled with --fmad=false so
// that results are bit-identical to the reference's x86-64 build (no FMA contraction there).
//
// Thread mapping (all grid kernels): a chunk's array is [nA][nB][pitch] with the fastest axis C
// padded to `pitch` (rows start on a 16-byte vector boundary).  A CTA owns 256 consecutive
// (vectors of) positions of the flattened (B,C) plane -- fully coalesced rows -- and marches along
// A over [a0,a1).  One launch covers every chunk: CTAs find their (chunk, cmp, plane segment, A range)
// in a tile table, and copy that chunk's parameter block to shared memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mcu {

#ifndef MCU_C3DV_F64_CTAS
#define MCU_C3DV_F64_CTAS 2 // resident CTAs per SM of the fp64 interior vector kernel (128 registers)
#endif
constexpr int kThreads = 256;
constexpr int kMaxPol = 4;
// How many planes ahead of the march the operands are pulled into L2 (prefetch.global.L2).  What matters is the
// distance in bytes: measured on the B200 (r02o, ms per step at 0 / 1 / 2 planes) -- fp64: vacuum 256^3 0.717 / 0.591 /
// 0.563, eps slab 384^2 x 576 2.93 / 2.63 / 2.45, Lorentzian 256^3 1.357 / 1.157 / 1.137; fp32: vacuum 256^3 0.374 /
// 0.308 / 0.311, eps slab 512^2 x 768 3.20 / 2.82 / 2.98 (3 and 4 planes: 3.31 / 3.71).  -DMCU_PREFETCH_PLANES=n forces n.
#ifdef MCU_PREFETCH_PLANES
template <class T> __host__ __device__ constexpr int pf_planes() { return MCU_PREFETCH_PLANES; }
#else
template <class T> __host__ __device__ constexpr int pf_planes() { return sizeof(T) == 4 ? 1 : 2; }
#endif

__device__ __forceinline__ void prefetch_l2(const void *p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

struct SrcEntry { // one source point, sorted by device index (stable in reference order)
  long long idx;
  double are, aim; // src_vol::amplitude(j)
  int tid;         // src_time id
  int pad;
};

template <class T> struct PolParams { // one lorentzian_susceptibility for one component/cmp
  T *p, *pp;                          // lorentzian_data::P / P_prev   (NULL: this pol has no P here)
  const T *s;                         // sigma[c][component_direction(c)] (NULL: no update)
  T gamma1inv, gamma1, omega0dtsqr, omega0dtsqr_denom; // susceptibility.cpp:191-194
};

template <class T> struct CompParams {
  // ---- curl update of the D/B component (step_curl's arguments after the g1/g2 swap) ----
  T *f;
  const T *g1, *g2;
  long long s1, s2;
  T dtdx, dt2;
  int lo[3], hi[3]; // owned index range on the device axes A,B,C (inclusive)
  T *fu, *fcnd;
  const T *cnd, *cndinv;
  const T *sig, *kap, *siginv;
  int kaxis, koff; // k = 2*idx[kaxis] + koff   (kaxis < 0: no PML in the f update)
  const T *sigu, *kapu, *siginvu;
  int kuaxis, kuoff;
  unsigned zmask_db; // metal planes: bit d -> idx[d]==lo[d], bit 3+d -> idx[d]==hi[d]
  // ---- 2-D exp(i beta z) term (step_db.cpp:161-175, step_beta): f += betadt * gb[i] and variants ----
  const T *gb;
  T betadt;
  // ---- BFAST (step_db.cpp:129-143, step_bfast): F = k1 (g1[i+s1]+g1[i]) - k2 (g2[i+s2]+g2[i]) - F ----
  T *fb;        // fields_chunk::f_bfast (NULL: no BFAST terms)
  T bk1, bk2;   // after the g1/g2 swap rule
  // ---- sources ----
  const SrcEntry *csrc; // currents (step_source)
  const SrcEntry *isrc; // integrated sources (update_eh.cpp:136-146)
  int ncsrc, nisrc;
  int csrc_amin, csrc_amax, isrc_amin, isrc_amax; // A range that contains source points
  int src_has_cndinv;
  // ---- E/H update (NULL e: H aliases B, or the update is done by k_eh_general) ----
  T *e;
  const T *u;
  T *fw;
  const T *sigw, *kapw;
  int kwaxis, kwoff;
  unsigned zmask_eh;
  int npol;
  PolParams<T> pol[kMaxPol];
  // ---- cylindrical coordinates (step_db.cpp:177-290,393-455) ----
  const T *gm; // the OTHER cmp of the component whose phi-derivative became the i m / r term (NULL: none)
  T the_m;     // 2 m (1 - 2 cmp) (+-1 by field type) (+-1 by direction) Courant
  int r0;      // here.r() = r0 + 2 * (index along R), R being device axis B
  int zrows;   // rows r < zrows of this component (and its f_u, f_cond) are zeroed after the curl: |m| > 1 next to the axis
};

template <class T> struct HalfParams { // one chunk, one of {B/H, D/E}, one cmp
  int nA, nB, nC, pitch;
  long long sA, sB; // device strides of axes A and B (C is 1)
  int is_D;         // 1: D/E half (currents from current_D, dipoles from dipole_E)
  int cmp;
  int march;        // k_halfg: 0 = CTAs march along A over a (B,C) plane, 1 = along B over (A,C)
  int pad_;
  double dt;
  CompParams<T> comp[3];
};

struct Tile {
  int param; // index into the parameter table of the kernel
  int a0, a1;
  int p0;    // first position of the flattened plane; bit 31 set: the tile may contain source points
             // or metal planes (k_curl3d then runs the loop body that checks for them)
};
constexpr int kTileSpecial = 1 << 30;
constexpr int kTileRows = 1 << 30; // k_eh_general: p0 = (row block << 12) | column block, 8 rows x 32 columns per CTA

// meepcu_src_sample layout: current_B[2], dipole_H[2], current_D[2], dipole_E[2]
__device__ __forceinline__ double2 sample_current(const double *smp, int tid, int is_D) {
  const double *s = smp + 8 * tid + (is_D ? 4 : 0);
  return make_double2(s[0], s[1]);
}
__device__ __forceinline__ double2 sample_dipole(const double *smp, int tid, int is_D) {
  const double *s = smp + 8 * tid + (is_D ? 6 : 2);
  return make_double2(s[0], s[1]);
}

__device__ __forceinline__ int src_lower_bound(const SrcEntry *t, int n, long long idx) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (t[mid].idx < idx) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// std::complex<double> product as libstdc++/libgcc evaluate it for finite operands
__device__ __forceinline__ double2 cmul(double a, double b, double c, double d) {
  return make_double2(a * c - b * d, a * d + b * c);
}

__device__ __forceinline__ int pick3(int axis, int a, int b, int c) { return axis == 0 ? a : (axis == 1 ? b : c); }

template <class T>
__device__ __forceinline__ bool on_zero_plane(unsigned m, const int lo[3], const int hi[3], int ia,
                                              int ib, int ic) {
  if (!m) return false;
  return ((m & 1u) && ia == lo[0]) || ((m & 2u) && ib == lo[1]) || ((m & 4u) && ic == lo[2]) ||
         ((m & 8u) && ia == hi[0]) || ((m & 16u) && ib == hi[1]) || ((m & 32u) && ic == hi[2]);
}

// ------------------------------------------------------------------------------------------------
// k_halfg: the general fused half step (PML / conductivity / polarisations, any dimensionality).
// Structured for memory-level parallelism:
// every global load of the three components of a point is issued before the first dependent
// instruction (two-phase gather -> compute/store), read-only arrays go through the read-only
// path, and the CTA marches along the LONGER of the two outer axes so that thin PML slabs still
// give long, fully coalesced streams.
//   HAS_COND: conductivity terms (cnd, cndinv, fcnd) compiled in;  HAS_POL: polarisations compiled in
// ------------------------------------------------------------------------------------------------
template <class T, bool HAS_COND, bool HAS_POL, bool HAS_BB = false> struct PointVals {
  T g1p, g1, g2, g2p, f, fu, e, fw, u, gb, fb, gm;
  T cnd, cndinv, fcnd;
  T p[HAS_POL ? kMaxPol : 1], pp[HAS_POL ? kMaxPol : 1], s[HAS_POL ? kMaxPol : 1];
};

template <class T, bool HAS_COND, bool HAS_POL, bool HAS_BB>
__device__ __forceinline__ void gather_point(const CompParams<T> &c, const PolParams<T> *pol, long long i,
                                             PointVals<T, HAS_COND, HAS_POL, HAS_BB> &v) {
  v.g1p = __ldg(c.g1 + i + c.s1);
  v.g1 = __ldg(c.g1 + i);
  if (c.g2) {
    v.g2 = __ldg(c.g2 + i);
    v.g2p = __ldg(c.g2 + i + c.s2);
  }
  v.f = c.f[i];
  if (c.fu) v.fu = c.fu[i];
  if (HAS_BB && c.gb) v.gb = __ldg(c.gb + i);
  if (HAS_BB && c.fb) v.fb = c.fb[i];
  if (HAS_BB && c.gm) v.gm = __ldg(c.gm + i);
  if (HAS_COND && c.cnd) {
    v.cnd = __ldg(c.cnd + i);
    v.cndinv = __ldg(c.cndinv + i);
    if (c.fcnd) v.fcnd = c.fcnd[i];
  }
  if (c.e) {
    if (c.u) v.u = __ldg(c.u + i);
    if (c.fw) {
      v.fw = c.fw[i];
      v.e = c.e[i];
    }
    if (HAS_POL) {
#pragma unroll
      for (int ip = 0; ip < kMaxPol; ++ip)
        if (ip < c.npol && pol[ip].p) {
          v.p[ip] = pol[ip].p[i];
          if (pol[ip].s) {
            v.pp[ip] = pol[ip].pp[i];
            v.s[ip] = __ldg(pol[ip].s + i);
          }
        }
    }
  }
}

template <class T, bool HAS_COND, bool HAS_POL, bool HAS_BB>
__device__ __forceinline__ void finish_point(const CompParams<T> &c, const PolParams<T> *pol, long long i, int ia, int ib, int ic,
                                             const PointVals<T, HAS_COND, HAS_POL, HAS_BB> &v,
                                             const double *__restrict__ smp, int is_D, int cmp, double dt) {
  // ---------------- step_curl (step_generic.cpp:89-252) ----------------
  T curl;
  if (c.g2)
    curl = (v.g1p - v.g1 + v.g2 - v.g2p);
  else
    curl = (v.g1p - v.g1);
  const bool cond = HAS_COND && c.cnd;
  T fnew, fun = 0, fc = 0; // new f, f_u, f_cond (stored below, after the beta term)
  T siginv = 0, siginvu = 0;
  const T fold = v.f;
  if (c.kaxis < 0) {
    if (c.kuaxis < 0) {
      if (cond)
        fnew = ((1 - c.dt2 * v.cnd) * fold - c.dtdx * curl) * v.cndinv;
      else
        fnew = fold - c.dtdx * curl;
    }
    else {
      const int ku = 2 * pick3(c.kuaxis, ia, ib, ic) + c.kuoff;
      const T fprev = v.fu;
      if (cond)
        fun = ((1 - c.dt2 * v.cnd) * fprev - c.dtdx * curl) * v.cndinv;
      else
        fun = fprev - c.dtdx * curl;
      siginvu = __ldg(c.siginvu + ku);
      fnew = siginvu * ((__ldg(c.kapu + ku) - __ldg(c.sigu + ku)) * fold + fun - fprev);
    }
  }
  else {
    const int k = 2 * pick3(c.kaxis, ia, ib, ic) + c.koff;
    const T kap = __ldg(c.kap + k), sig = __ldg(c.sig + k);
    siginv = __ldg(c.siginv + k);
    if (c.kuaxis < 0) {
      if (cond) {
        const T fcnd_prev = v.fcnd;
        fc = ((1 - c.dt2 * v.cnd) * fcnd_prev - c.dtdx * curl) * v.cndinv;
        fnew = ((kap - sig) * fold + (fc - fcnd_prev)) * siginv;
      }
      else
        fnew = ((kap - sig) * fold - c.dtdx * curl) * siginv;
    }
    else {
      const int ku = 2 * pick3(c.kuaxis, ia, ib, ic) + c.kuoff;
      const T fprev = v.fu;
      if (cond) {
        const T fcnd_prev = v.fcnd;
        fc = ((1 - c.dt2 * v.cnd) * fcnd_prev - c.dtdx * curl) * v.cndinv;
        fun = ((kap - sig) * fprev + (fc - fcnd_prev)) * siginv;
      }
      else
        fun = ((kap - sig) * fprev - c.dtdx * curl) * siginv;
      siginvu = __ldg(c.siginvu + ku);
      fnew = siginvu * ((__ldg(c.kapu + ku) - __ldg(c.sigu + ku)) * fold + fun - fprev);
    }
  }
  // ---------------- step_bfast (step_generic.cpp:339-537) and step_beta (:259-337): both add an
  // increment to f through the same PML / conductivity chain, bfast right after the curl of the
  // component (step_db.cpp:129-143), beta after all curls (step_db.cpp:161-175) ----------------------
#pragma unroll
  for (int pass = 0; pass < (HAS_BB ? 3 : 0); ++pass) { // compiled only into the class that has these terms
    if (pass == 0 ? !c.fb : (pass == 1 ? !c.gb : !c.gm)) continue;
    T bg;
    if (pass == 0) {
      const T F_prev = v.fb;
      T Fn;
      if (c.g2) Fn = (c.bk1 * (v.g1p + v.g1) - c.bk2 * (v.g2p + v.g2)) - F_prev;
      else if (c.kaxis < 0 && c.kuaxis < 0 && !cond) Fn = c.bk1 * (v.g1p + v.g1); // :376, as written in the reference
      else Fn = c.bk1 * (v.g1p + v.g1) - F_prev;
      c.fb[i] = Fn;
      bg = Fn - F_prev;
    }
    else if (pass == 1)
      bg = c.betadt * v.gb;
    else { // cylindrical i m / r term (step_db.cpp:177-290): rinv = the_m / here.r(), increment rinv * g[i]
      const T rinv = c.the_m / (T)(c.r0 + 2 * ib);
      bg = rinv * v.gm;
    }
    if (c.kaxis >= 0) {
      if (c.kuaxis >= 0) {
        T df;
        if (cond) {
          const T dfcnd = bg * v.cndinv;
          fc = fc + dfcnd;
          fun = fun + (df = dfcnd * siginv);
        }
        else
          fun = fun + (df = bg * siginv);
        fnew = fnew + siginvu * df;
      }
      else {
        if (cond) {
          const T dfcnd = bg * v.cndinv;
          fc = fc + dfcnd;
          fnew = fnew + dfcnd * siginv;
        }
        else
          fnew = fnew + bg * siginv;
      }
    }
    else {
      if (c.kuaxis >= 0) {
        T df;
        if (cond) df = bg * v.cndinv;
        else df = bg;
        fun = fun + df;
        fnew = fnew + siginvu * df;
      }
      else {
        if (cond) fnew = fnew + bg * v.cndinv;
        else fnew = fnew + bg;
      }
    }
  }
  if (HAS_BB && ib < c.zrows) { // |m| > 1: the rows next to the axis are zeroed at the end of step_db (:393-455)
    fnew = 0;
    fun = 0;
    fc = 0;
  }
  if (c.kuaxis >= 0) c.fu[i] = fun;
  if (cond && c.kaxis >= 0) c.fcnd[i] = fc;
  // ---------------- step_source (step.cpp:296-319) ----------------
  if (c.ncsrc && ia >= c.csrc_amin && ia <= c.csrc_amax) {
    for (int j = src_lower_bound(c.csrc, c.ncsrc, i); j < c.ncsrc && c.csrc[j].idx == i; ++j) {
      const SrcEntry e = c.csrc[j];
      const double2 cur = sample_current(smp, e.tid, is_D);
      double2 A = cmul(e.are, e.aim, cur.x, cur.y);
      A.x = A.x * dt;
      A.y = A.y * dt;
      if (c.src_has_cndinv) {
        const double ci = double(c.cndinv[i]);
        A.x = A.x * ci;
        A.y = A.y * ci;
      }
      fnew = (T)((double)fnew - (cmp ? A.y : A.x));
    }
  }
  if (on_zero_plane<T>(c.zmask_db, c.lo, c.hi, ia, ib, ic)) fnew = 0;
  c.f[i] = fnew;

  // ---------------- update_eh / update_pols ----------------
  if (!c.e) return;
  T fmp = fnew;
  if (HAS_POL) {
#pragma unroll
    for (int ip = 0; ip < kMaxPol; ++ip)
      if (ip < c.npol && pol[ip].p) fmp -= v.p[ip];
  }
  if (c.nisrc && ia >= c.isrc_amin && ia <= c.isrc_amax) {
    for (int j = src_lower_bound(c.isrc, c.nisrc, i); j < c.nisrc && c.isrc[j].idx == i; ++j) {
      const SrcEntry e = c.isrc[j];
      const double2 dp = sample_dipole(smp, e.tid, is_D);
      const double2 A = cmul(e.are, e.aim, dp.x, dp.y);
      fmp = (T)((double)fmp - (cmp ? A.y : A.x));
    }
  }
  const T gu = c.u ? (fmp * v.u) : fmp;
  T enew;
  if (c.fw) {
    const int kw = 2 * pick3(c.kwaxis, ia, ib, ic) + c.kwoff;
    const T kapwkw = __ldg(c.kapw + kw), sigwkw = __ldg(c.sigw + kw);
    c.fw[i] = gu;
    enew = v.e + ((kapwkw + sigwkw) * gu - (kapwkw - sigwkw) * v.fw);
  }
  else
    enew = gu;
  if (HAS_POL) {
#pragma unroll
    for (int ip = 0; ip < kMaxPol; ++ip)
      if (ip < c.npol && pol[ip].p && pol[ip].s) {
        const PolParams<T> &q = pol[ip];
        const T pcur = v.p[ip];
        q.p[i] = q.gamma1inv * (pcur * (2 - q.omega0dtsqr_denom) - q.gamma1 * v.pp[ip] +
                                q.omega0dtsqr * (v.s[ip] * gu));
        q.pp[i] = pcur;
      }
  }
  if (on_zero_plane<T>(c.zmask_eh, c.lo, c.hi, ia, ib, ic)) enew = 0;
  c.e[i] = enew;
}

// One CTA = one COMPONENT of one tile (Tile::param = 4*HalfParams index + component): the
// component's parameters live in registers for the whole march, and the three components of a
// spatial tile sit next to each other in the tile table so that their shared operands meet in L2.
template <class T> struct HalfHeader { // leading, component-independent part of HalfParams<T>
  int nA, nB, nC, pitch;
  long long sA, sB;
  int is_D, cmp, march, pad_;
  double dt;
};

template <class T, bool HAS_COND, bool HAS_POL, bool HAS_BB = false>
__global__ void __launch_bounds__(kThreads, HAS_COND ? 2 : 3)
    k_halfg(const HalfParams<T> *__restrict__ params, const Tile *__restrict__ tiles,
            const double *__restrict__ smp) {
  __shared__ HalfHeader<T> H;
  __shared__ CompParams<T> Csm;
  const Tile t = tiles[blockIdx.x];
  const int pidx = t.param >> 2, kcomp = t.param & 3;
  {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(params + pidx);
    uint32_t *dst = reinterpret_cast<uint32_t *>(&H);
    for (int k = threadIdx.x; k < (int)(sizeof(HalfHeader<T>) / 4); k += kThreads) dst[k] = src[k];
    const uint32_t *src2 = reinterpret_cast<const uint32_t *>(&params[pidx].comp[kcomp]);
    uint32_t *dst2 = reinterpret_cast<uint32_t *>(&Csm);
    for (int k = threadIdx.x; k < (int)(sizeof(CompParams<T>) / 4); k += kThreads) dst2[k] = src2[k];
  }
  __syncthreads();
  const int p = t.p0 + threadIdx.x;
  const int io = p / H.pitch, ic = p - io * H.pitch; // io: index on the outer axis we do NOT march
  const int march = H.march;
  const int n_other = march ? H.nA : H.nB;
  if (io >= n_other || ic >= H.nC) return;
  const CompParams<T> c = Csm; // registers
  const int lo_o = march ? c.lo[0] : c.lo[1], hi_o = march ? c.hi[0] : c.hi[1];
  const int lo_m = march ? c.lo[1] : c.lo[0], hi_m = march ? c.hi[1] : c.hi[0];
  if (!(io >= lo_o && io <= hi_o && ic >= c.lo[2] && ic <= c.hi[2])) return;
  const long long s_m = march ? H.sB : H.sA, s_o = march ? H.sA : H.sB;
  const int m0 = t.a0 > lo_m ? t.a0 : lo_m;
  const int m1 = t.a1 - 1 < hi_m ? t.a1 - 1 : hi_m;
  long long i = (long long)m0 * s_m + (long long)io * s_o + ic;
  const int is_D = H.is_D, cmp = H.cmp;
  const double dt = H.dt;
#pragma unroll 2
  for (int im = m0; im <= m1; ++im, i += s_m) {
    const int ia = march ? io : im, ib = march ? im : io;
    if (pf_planes<T>() > 0 && im + pf_planes<T>() <= m1) {
      const long long ipf = i + pf_planes<T>() * s_m;
      prefetch_l2(c.g1 + ipf);
      if (c.g2) prefetch_l2(c.g2 + ipf);
      prefetch_l2(c.f + ipf);
      if (c.fu) prefetch_l2(c.fu + ipf);
      if (c.e) {
        if (c.u) prefetch_l2(c.u + ipf);
        if (c.fw) {
          prefetch_l2(c.fw + ipf);
          prefetch_l2(c.e + ipf);
        }
        if (HAS_POL) {
#pragma unroll
          for (int ip = 0; ip < kMaxPol; ++ip)
            if (ip < c.npol && Csm.pol[ip].p) {
              prefetch_l2(Csm.pol[ip].p + ipf);
              if (Csm.pol[ip].s) {
                prefetch_l2(Csm.pol[ip].pp + ipf);
                prefetch_l2(Csm.pol[ip].s + ipf);
              }
            }
        }
      }
    }
    PointVals<T, HAS_COND, HAS_POL, HAS_BB> v;
    gather_point<T, HAS_COND, HAS_POL, HAS_BB>(c, Csm.pol, i, v);
    finish_point<T, HAS_COND, HAS_POL, HAS_BB>(c, Csm.pol, i, ia, ib, ic, v, smp, is_D, cmp, dt);
  }
}

// ------------------------------------------------------------------------------------------------
// k_curl3d: the interior fast path.  Same arithmetic as k_halfg, specialised for 3-D chunks
// without PML / conductivity / polarisations (78 % of the cells of BASELINE config 2):
//   * the step plan of a Cartesian 3-D chunk (src/fields.cpp:441-469) is hard-wired:
//       f_k -= dtdx * ( g_{k+2}[i +- s_{k+1}] - g_{k+2}[i] + g_{k+1}[i] - g_{k+1}[i +- s_{k+2}] )
//     (+ for B from E, - for D from H; src/step_db.cpp:81-84), so the three components share their
//     nine operand loads;
//   * the x (= marching axis) neighbours of g_y and g_z are carried in registers from one plane to
//     the next, so each operand element is loaded once per CTA column: 7 operand loads per cell;
//   * operands are read through the read-only path (they are never written in the same half
//     step), which lets the compiler keep all loads of a plane in flight together;
//   * 32-bit indexing, <= 64 registers -> 4 CTAs / 32 warps per SM.
// ------------------------------------------------------------------------------------------------
template <class T> struct Curl3DParams {
  int nA, nB, nC, pitch;
  int sA, sB;
  int cmp, any_src;
  double dt;
  T dtdx;
  T *f[3];
  const T *g[3];
  T *e[3];        // all NULL: no E/H update in this chunk (H aliases B)
  const T *u[3];  // chi1inv diagonal (NULL: trivial)
  int lo[3][3], hi[3][3]; // [component][axis]
  unsigned zdb[3], zeh[3];
  const SrcEntry *csrc[3], *isrc[3];
  int ncsrc[3], nisrc[3];
  int ca0[3], ca1[3], ia0[3], ia1[3];
  // ---- virtual E (B/H half of a chunk whose E arrays are not kept up to date while stepping) ----
  // Where E = chi1inv * (D - dipoles) is a pure function of D (no PML, no polarisation, diagonal eps), storing E
  // in the D/E half and reading it back in the B/H half moves 6 of the 21-24 values per cell for nothing
  // (step_generic.cpp:891-894,903 write what :108 of the next half reads).  In virtual mode the D/E half of such a
  // chunk writes no E, and the B/H half recomputes each operand from D (x chi1inv) as it loads it -- the same
  // multiplication the reference did when it stored E, so every bit is the same.  Not-owned points: the E halo
  // copies of step_boundaries land in the D array's not-owned points (nobody reads those while stepping) and
  // the device copy of chi1inv is 1 there.
  int virt;            // 1: operands of the B/H half may be read as vu * vg (launch argument vmode selects)
  const T *vg[3];      // D components
  const T *vu[3];      // chi1inv diagonal of the E components (NULL: E = D)
  const SrcEntry *vsrc[3]; // integrated E sources of the chunk (update_eh.cpp:136-146): subtracted before the multiply
  int nvsrc[3], va0[3], va1[3];
};

template <class T, bool IS_D, int K, bool SPECIAL>
__device__ __forceinline__ void curl3d_finish(const Curl3DParams<T> &P, int i, int ia, int ib, int ic,
                                              T curl, T fold, T uval, T *__restrict__ f, T *__restrict__ e,
                                              bool has_u, const double *__restrict__ smp) {
  T fnew = fold - P.dtdx * curl; // step_generic.cpp:108
  if (SPECIAL && P.any_src && P.ncsrc[K] && ia >= P.ca0[K] && ia <= P.ca1[K]) { // step.cpp:296-319
    const SrcEntry *tab = P.csrc[K];
    for (int j = src_lower_bound(tab, P.ncsrc[K], i); j < P.ncsrc[K] && tab[j].idx == i; ++j) {
      const SrcEntry e = tab[j];
      const double2 cur = sample_current(smp, e.tid, IS_D);
      double2 A = cmul(e.are, e.aim, cur.x, cur.y);
      A.x = A.x * P.dt;
      A.y = A.y * P.dt;
      fnew = (T)((double)fnew - (P.cmp ? A.y : A.x));
    }
  }
  if (SPECIAL && on_zero_plane<T>(P.zdb[K], P.lo[K], P.hi[K], ia, ib, ic)) fnew = 0;
  f[i] = fnew;
  if (!e) return;
  T fmp = fnew;
  if (SPECIAL && P.any_src && P.nisrc[K] && ia >= P.ia0[K] && ia <= P.ia1[K]) { // update_eh.cpp:136-146
    const SrcEntry *tab = P.isrc[K];
    for (int j = src_lower_bound(tab, P.nisrc[K], i); j < P.nisrc[K] && tab[j].idx == i; ++j) {
      const SrcEntry se = tab[j];
      const double2 dp = sample_dipole(smp, se.tid, IS_D);
      const double2 A = cmul(se.are, se.aim, dp.x, dp.y);
      fmp = (T)((double)fmp - (P.cmp ? A.y : A.x));
    }
  }
  T enew = has_u ? (fmp * uval) : fmp; // step_generic.cpp:891-894,903
  if (SPECIAL && on_zero_plane<T>(P.zeh[K], P.lo[K], P.hi[K], ia, ib, ic)) enew = 0;
  e[i] = enew;
}

// one operand of the B/H curl in virtual-E mode: E_K[idx] recomputed as the D/E half would have stored it
// (curl3d_finish: fmp = D - dipoles, enew = fmp * u), dipoles from the PREVIOUS step's dipole_E sample
template <class T, int K, bool SPECIAL>
__device__ __forceinline__ T virt_e_load(const Curl3DParams<T> &P, const T *__restrict__ d, const T *__restrict__ u, int idx,
                                         const double *__restrict__ smp_prev) {
  T v = __ldg(d + idx);
  if (SPECIAL && P.nvsrc[K]) {
    const int pa = idx / P.sA;
    if (pa >= P.va0[K] && pa <= P.va1[K]) {
      const SrcEntry *tab = P.vsrc[K];
      for (int j = src_lower_bound(tab, P.nvsrc[K], idx); j < P.nvsrc[K] && tab[j].idx == idx; ++j) {
        const SrcEntry se = tab[j];
        const double2 dp = sample_dipole(smp_prev, se.tid, 1);
        const double2 A = cmul(se.are, se.aim, dp.x, dp.y);
        v = (T)((double)v - (P.cmp ? A.y : A.x));
      }
    }
  }
  if (u) v = v * __ldg(u + idx);
  return v;
}

template <class T, bool IS_D, bool SPECIAL>
__device__ __forceinline__ void curl3d_loop(const Curl3DParams<T> &P, const Tile t, int ib, int ic,
                                            const bool act[3], const double *__restrict__ smp,
                                            const double *__restrict__ smp_prev = 0, int vmode = 0) {
  const int sA = P.sA, nA = P.nA;
  const bool virt = !IS_D && vmode && P.virt;
  // neighbour offsets, clamped so that loads whose value no owned component consumes stay in bounds
  int oy, oz;
  if (!IS_D) {
    oy = (ib + 1 < P.nB) ? P.sB : 0;
    oz = (ic + 1 < P.nC) ? 1 : 0;
  }
  else {
    oy = (ib > 0) ? -P.sB : 0;
    oz = (ic > 0) ? -1 : 0;
  }
  const T *__restrict__ gx = virt ? P.vg[0] : P.g[0];
  const T *__restrict__ gy = virt ? P.vg[1] : P.g[1];
  const T *__restrict__ gz = virt ? P.vg[2] : P.g[2];
  const T *__restrict__ wx = virt ? P.vu[0] : 0; // multipliers of the operands (virtual E only)
  const T *__restrict__ wy = virt ? P.vu[1] : 0;
  const T *__restrict__ wz = virt ? P.vu[2] : 0;
  T *__restrict__ f0 = P.f[0];
  T *__restrict__ f1 = P.f[1];
  T *__restrict__ f2 = P.f[2];
  T *__restrict__ e0 = P.e[0];
  T *__restrict__ e1 = P.e[1];
  T *__restrict__ e2 = P.e[2];
  const T *__restrict__ u0 = P.u[0];
  const T *__restrict__ u1 = P.u[1];
  const T *__restrict__ u2 = P.u[2];
  const bool hu0 = e0 && u0, hu1 = e1 && u1, hu2 = e2 && u2;
  int i = t.a0 * sA + ib * P.sB + ic;
#define MCU_LDX(idx) (virt ? virt_e_load<T, 0, SPECIAL>(P, gx, wx, (idx), smp_prev) : __ldg(gx + (idx)))
#define MCU_LDY(idx) (virt ? virt_e_load<T, 1, SPECIAL>(P, gy, wy, (idx), smp_prev) : __ldg(gy + (idx)))
#define MCU_LDZ(idx) (virt ? virt_e_load<T, 2, SPECIAL>(P, gz, wz, (idx), smp_prev) : __ldg(gz + (idx)))
  T ry, rz; // g_y, g_z of the plane the x-difference needs, carried between iterations
  if (!IS_D) {
    ry = MCU_LDY(i);
    rz = MCU_LDZ(i);
  }
  else {
    const int ip = t.a0 > 0 ? i - sA : i;
    ry = __ldg(gy + ip);
    rz = __ldg(gz + ip);
  }
  const int lo0 = P.lo[0][0], hi0 = P.hi[0][0], lo1 = P.lo[1][0], hi1 = P.hi[1][0], lo2 = P.lo[2][0],
            hi2 = P.hi[2][0];
#pragma unroll 2
  for (int ia = t.a0; ia < t.a1; ++ia, i += sA) {
    const bool on0 = act[0] && ia >= lo0 && ia <= hi0, on1 = act[1] && ia >= lo1 && ia <= hi1,
               on2 = act[2] && ia >= lo2 && ia <= hi2;
    if (pf_planes<T>() > 0 && ia + pf_planes<T>() < nA) {
      // pull the planes this thread will need a few iterations from now into L2: turns the
      // DRAM latency of the marching loads into L2 latency without holding registers
      const int ipf = i + pf_planes<T>() * sA;
      prefetch_l2(gx + ipf);
      prefetch_l2(gy + ipf);
      prefetch_l2(gz + ipf);
      prefetch_l2(f0 + ipf);
      prefetch_l2(f1 + ipf);
      prefetch_l2(f2 + ipf);
      if (hu0) prefetch_l2(u0 + ipf);
      if (hu1) prefetch_l2(u1 + ipf);
      if (hu2) prefetch_l2(u2 + ipf);
    }
    // ---- every load of this plane first ----
    const T gx0 = MCU_LDX(i), gxy = MCU_LDX(i + oy), gxz = MCU_LDX(i + oz);
    const T gyz = MCU_LDY(i + oz);
    const T gzy = MCU_LDZ(i + oy);
    T gy0, gz0, gyx, gzx;
    if (!IS_D) {
      const int ox = (ia + 1 < nA) ? sA : 0;
      gy0 = ry;
      gz0 = rz;
      gyx = MCU_LDY(i + ox);
      gzx = MCU_LDZ(i + ox);
      ry = gyx;
      rz = gzx;
    }
    else {
      gy0 = __ldg(gy + i);
      gz0 = __ldg(gz + i);
      gyx = ry;
      gzx = rz;
      ry = gy0;
      rz = gz0;
    }
    T fo0 = 0, fo1 = 0, fo2 = 0, uv0 = 0, uv1 = 0, uv2 = 0;
    if (on0) fo0 = f0[i];
    if (on1) fo1 = f1[i];
    if (on2) fo2 = f2[i];
    if (hu0 && on0) uv0 = __ldg(u0 + i);
    if (hu1 && on1) uv1 = __ldg(u1 + i);
    if (hu2 && on2) uv2 = __ldg(u2 + i);
    // component x: g1 = g_z (d/dy), g2 = g_y (d/dz);  y: g1 = g_x (d/dz), g2 = g_z (d/dx);
    // z: g1 = g_y (d/dx), g2 = g_x (d/dy)      -- (g1[i+s1] - g1[i] + g2[i] - g2[i+s2])
    if (on0)
      curl3d_finish<T, IS_D, 0, SPECIAL>(P, i, ia, ib, ic, (gzy - gz0 + gy0 - gyz), fo0, uv0, f0, e0, hu0, smp);
    if (on1)
      curl3d_finish<T, IS_D, 1, SPECIAL>(P, i, ia, ib, ic, (gxz - gx0 + gz0 - gzx), fo1, uv1, f1, e1, hu1, smp);
    if (on2)
      curl3d_finish<T, IS_D, 2, SPECIAL>(P, i, ia, ib, ic, (gyx - gy0 + gx0 - gxy), fo2, uv2, f2, e2, hu2, smp);
  }
#undef MCU_LDX
#undef MCU_LDY
#undef MCU_LDZ
}

// hook run by a thread when it leaves a tile body (nothing today)
struct NoExit { __device__ __forceinline__ void operator()() const {} };

template <class T, bool IS_D, class OnExit = NoExit>
__device__ __forceinline__ void curl3d_body(const Curl3DParams<T> &P, const Tile t, bool special,
                                            const double *__restrict__ smp, const double *__restrict__ smp_prev, int vmode,
                                            const OnExit on_exit = OnExit());
template <class T, bool IS_D, class OnExit>
__device__ __forceinline__ void curl3d_body(const Curl3DParams<T> &P, const Tile t, bool special,
                                            const double *__restrict__ smp, const double *__restrict__ smp_prev, int vmode,
                                            const OnExit on_exit) {
  const int p = t.p0 + threadIdx.x;
  const int ib = p / P.pitch, ic = p - ib * P.pitch;
  if (ib >= P.nB || ic >= P.nC) { on_exit(); return; }
  bool act[3];
#pragma unroll
  for (int k = 0; k < 3; ++k)
    act[k] = ib >= P.lo[k][1] && ib <= P.hi[k][1] && ic >= P.lo[k][2] && ic <= P.hi[k][2];
  if (!(act[0] || act[1] || act[2])) { on_exit(); return; }
  if (special)
    curl3d_loop<T, IS_D, true>(P, t, ib, ic, act, smp, smp_prev, vmode);
  else
    curl3d_loop<T, IS_D, false>(P, t, ib, ic, act, smp, smp_prev, vmode);
  on_exit();
}

template <class U> __device__ __forceinline__ void block_copy_params(U *dst_, const U *src_) {
  const int nw = (int)(sizeof(U) / 4);
  const uint32_t *src = reinterpret_cast<const uint32_t *>(src_);
  uint32_t *dst = reinterpret_cast<uint32_t *>(dst_);
  for (int k = threadIdx.x; k < nw; k += kThreads) dst[k] = src[k];
}

template <class T, bool IS_D>
__global__ void __launch_bounds__(kThreads, sizeof(T) == 8 ? 3 : 4)
    k_curl3d(const Curl3DParams<T> *__restrict__ params, const Tile *__restrict__ tiles,
             const double *__restrict__ smp, const double *__restrict__ smp_prev, int vmode) {
  __shared__ Curl3DParams<T> P;
  Tile t = tiles[blockIdx.x];
  const bool special = (t.p0 & kTileSpecial) != 0;
  t.p0 &= ~kTileSpecial;
  block_copy_params(&P, params + t.param);
  __syncthreads();
  curl3d_body<T, IS_D>(P, t, special, smp, smp_prev, vmode);
}

// ------------------------------------------------------------------------------------------------
// k_curl3d_v: k_curl3d with V consecutive z points per thread (V = 2 doubles / 4 floats = 16-byte
// loads and stores).  Only "clean" tiles (no source point, no metal plane) are routed here; the
// arithmetic per point is the same expression sequence, so results are bit-identical.  The
// z-neighbour of a vector is its own elements shifted by one plus ONE scalar load.
// ------------------------------------------------------------------------------------------------
template <class T> struct VecOf;
template <> struct VecOf<double> { typedef double2 type; enum { V = 2 }; };
template <> struct VecOf<float> { typedef float4 type; enum { V = 4 }; };

template <class T> __device__ __forceinline__ void vload(T *out, const T *p) {
  typedef typename VecOf<T>::type VT;
  const VT v = __ldg(reinterpret_cast<const VT *>(p));
  const T *e = reinterpret_cast<const T *>(&v);
#pragma unroll
  for (int j = 0; j < VecOf<T>::V; ++j) out[j] = e[j];
}
template <class T> __device__ __forceinline__ void vload_rw(T *out, const T *p) { // array we also write
  typedef typename VecOf<T>::type VT;
  const VT v = *reinterpret_cast<const VT *>(p);
  const T *e = reinterpret_cast<const T *>(&v);
#pragma unroll
  for (int j = 0; j < VecOf<T>::V; ++j) out[j] = e[j];
}
template <class T> __device__ __forceinline__ void vstore(T *p, const T *in) {
  typedef typename VecOf<T>::type VT;
  VT v;
  T *e = reinterpret_cast<T *>(&v);
#pragma unroll
  for (int j = 0; j < VecOf<T>::V; ++j) e[j] = in[j];
  *reinterpret_cast<VT *>(p) = v;
}

// operand vector of the B/H curl: g[idx..idx+V), times u[idx..idx+V) in virtual-E mode with a chi1inv array
template <class T, bool VU> __device__ __forceinline__ void vload_e(T *out, const T *g, const T *u, int idx) {
  vload<T>(out, g + idx);
  if (VU && u) {
    T w[VecOf<T>::V];
    vload<T>(w, u + idx);
#pragma unroll
    for (int j = 0; j < VecOf<T>::V; ++j) out[j] = out[j] * w[j];
  }
}
template <class T, bool VU> __device__ __forceinline__ T sload_e(const T *g, const T *u, int idx) {
  const T v = __ldg(g + idx);
  return (VU && u) ? v * __ldg(u + idx) : v;
}

// VU: compiled with the chi1inv multiplication of virtual-E operands (only launched when some chunk needs it, so
// that the plain kernel keeps its register count)
template <class T, bool IS_D, bool VU = false, class OnExit = NoExit>
__device__ __forceinline__ void curl3d_v_body(const Curl3DParams<T> &P, const Tile t, int vmode = 0, const OnExit on_exit = OnExit()) {
  constexpr int V = VecOf<T>::V;
  const bool virt = !IS_D && vmode && P.virt;
  const int pv = P.pitch / V; // vectors per row
  const int p = t.p0 + threadIdx.x;
  const int ib = p / pv, ic = (p - ib * pv) * V;
  if (ib >= P.nB || ic >= P.nC) { on_exit(); return; }
  // per component: is this row active, and which of my V elements are owned in z
  bool rowact[3];
  unsigned zmask[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    rowact[k] = ib >= P.lo[k][1] && ib <= P.hi[k][1];
    unsigned m = 0;
#pragma unroll
    for (int j = 0; j < V; ++j)
      if (ic + j >= P.lo[k][2] && ic + j <= P.hi[k][2]) m |= 1u << j;
    zmask[k] = rowact[k] ? m : 0u;
  }
  if (!(zmask[0] | zmask[1] | zmask[2])) { on_exit(); return; }
  const unsigned full = (1u << V) - 1u;
  const int sA = P.sA, nA = P.nA;
  int oy, ozs; // row offset of the y neighbour; offset of the ONE scalar z neighbour
  if (!IS_D) {
    oy = (ib + 1 < P.nB) ? P.sB : 0;
    ozs = (ic + V < P.pitch) ? V : 0; // in bounds inside the padded row; value unused when not owned
  }
  else {
    oy = (ib > 0) ? -P.sB : 0;
    ozs = (ic > 0) ? -1 : 0;
  }
  const T *__restrict__ gx = virt ? P.vg[0] : P.g[0];
  const T *__restrict__ gy = virt ? P.vg[1] : P.g[1];
  const T *__restrict__ gz = virt ? P.vg[2] : P.g[2];
  const T *__restrict__ wx = (VU && virt) ? P.vu[0] : 0;
  const T *__restrict__ wy = (VU && virt) ? P.vu[1] : 0;
  const T *__restrict__ wz = (VU && virt) ? P.vu[2] : 0;
  T *__restrict__ f0 = P.f[0];
  T *__restrict__ f1 = P.f[1];
  T *__restrict__ f2 = P.f[2];
  T *__restrict__ e0 = P.e[0];
  T *__restrict__ e1 = P.e[1];
  T *__restrict__ e2 = P.e[2];
  const T *__restrict__ u0 = P.u[0];
  const T *__restrict__ u1 = P.u[1];
  const T *__restrict__ u2 = P.u[2];
  const bool hu0 = e0 && u0, hu1 = e1 && u1, hu2 = e2 && u2;
  const T dtdx = P.dtdx;
  int i = t.a0 * sA + ib * P.sB + ic;
  T ry[V], rz[V];
  if (!IS_D) {
    vload_e<T, VU>(ry, gy, wy, i);
    vload_e<T, VU>(rz, gz, wz, i);
  }
  else {
    const int ip = t.a0 > 0 ? i - sA : i;
    vload<T>(ry, gy + ip);
    vload<T>(rz, gz + ip);
  }
  const int lo0 = P.lo[0][0], hi0 = P.hi[0][0], lo1 = P.lo[1][0], hi1 = P.hi[1][0], lo2 = P.lo[2][0],
            hi2 = P.hi[2][0];
  for (int ia = t.a0; ia < t.a1; ++ia, i += sA) {
    const unsigned m0 = (ia >= lo0 && ia <= hi0) ? zmask[0] : 0u, m1 = (ia >= lo1 && ia <= hi1) ? zmask[1] : 0u,
                   m2 = (ia >= lo2 && ia <= hi2) ? zmask[2] : 0u;
    if (pf_planes<T>() > 0 && ia + pf_planes<T>() < nA) {
      const int ipf = i + pf_planes<T>() * sA;
      prefetch_l2(gx + ipf);
      prefetch_l2(gy + ipf);
      prefetch_l2(gz + ipf);
      prefetch_l2(f0 + ipf);
      prefetch_l2(f1 + ipf);
      prefetch_l2(f2 + ipf);
      if (hu0) prefetch_l2(u0 + ipf);
      if (hu1) prefetch_l2(u1 + ipf);
      if (hu2) prefetch_l2(u2 + ipf);
      if (VU && wx) prefetch_l2(wx + ipf);
      if (VU && wy) prefetch_l2(wy + ipf);
      if (VU && wz) prefetch_l2(wz + ipf);
    }
    T gx0[V], gxy[V], gy0[V], gz0[V], gzy[V], gyx[V], gzx[V];
    T gxs, gys; // the scalar z neighbours of g_x and g_y
    vload_e<T, VU>(gx0, gx, wx, i);
    vload_e<T, VU>(gxy, gx, wx, i + oy);
    gxs = sload_e<T, VU>(gx, wx, i + ozs);
    gys = sload_e<T, VU>(gy, wy, i + ozs);
    vload_e<T, VU>(gzy, gz, wz, i + oy);
    if (!IS_D) {
      const int ox = (ia + 1 < nA) ? sA : 0;
#pragma unroll
      for (int j = 0; j < V; ++j) { gy0[j] = ry[j]; gz0[j] = rz[j]; }
      vload_e<T, VU>(gyx, gy, wy, i + ox);
      vload_e<T, VU>(gzx, gz, wz, i + ox);
#pragma unroll
      for (int j = 0; j < V; ++j) { ry[j] = gyx[j]; rz[j] = gzx[j]; }
    }
    else {
      vload<T>(gy0, gy + i);
      vload<T>(gz0, gz + i);
#pragma unroll
      for (int j = 0; j < V; ++j) { gyx[j] = ry[j]; gzx[j] = rz[j]; ry[j] = gy0[j]; rz[j] = gz0[j]; }
    }
    T fo0[V], fo1[V], fo2[V], uv0[V], uv1[V], uv2[V];
    if (m0) vload_rw<T>(fo0, f0 + i);
    if (m1) vload_rw<T>(fo1, f1 + i);
    if (m2) vload_rw<T>(fo2, f2 + i);
    if (hu0 && m0) vload<T>(uv0, u0 + i);
    if (hu1 && m1) vload<T>(uv1, u1 + i);
    if (hu2 && m2) vload<T>(uv2, u2 + i);
    // z neighbours: element j of the shifted vector
    T gxz[V], gyz[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
      if (!IS_D) {
        gxz[j] = (j + 1 < V) ? gx0[j + 1 < V ? j + 1 : j] : gxs;
        gyz[j] = (j + 1 < V) ? gy0[j + 1 < V ? j + 1 : j] : gys;
      }
      else {
        gxz[j] = (j > 0) ? gx0[j > 0 ? j - 1 : 0] : gxs;
        gyz[j] = (j > 0) ? gy0[j > 0 ? j - 1 : 0] : gys;
      }
    }
    // (g1[i+s1] - g1[i] + g2[i] - g2[i+s2]) per component, as in k_curl3d
    T n0[V], n1[V], n2[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
      n0[j] = fo0[j] - dtdx * (gzy[j] - gz0[j] + gy0[j] - gyz[j]);
      n1[j] = fo1[j] - dtdx * (gxz[j] - gx0[j] + gz0[j] - gzx[j]);
      n2[j] = fo2[j] - dtdx * (gyx[j] - gy0[j] + gx0[j] - gxy[j]);
    }
#define MCU_V_FINISH(K, m, nn, fo, ff, ee, hu, uv)                                                     \
    if (m) {                                                                                       \
      T ev[V];                                                                                     \
      _Pragma("unroll") for (int j = 0; j < V; ++j) {                                              \
        if (!((m >> j) & 1u)) nn[j] = fo[j]; /* not owned: write back what was there */            \
        ev[j] = hu ? (nn[j] * uv[j]) : nn[j];                                                      \
      }                                                                                            \
      vstore<T>(ff + i, nn);                                                                       \
      if (ee) {                                                                                    \
        if (m == full)                                                                             \
          vstore<T>(ee + i, ev);                                                                   \
        else {                                                                                     \
          _Pragma("unroll") for (int j = 0; j < V; ++j)                                            \
            if ((m >> j) & 1u) ee[i + j] = ev[j];                                                  \
        }                                                                                          \
      }                                                                                            \
    }
    MCU_V_FINISH(0, m0, n0, fo0, f0, e0, hu0, uv0)
    MCU_V_FINISH(1, m1, n1, fo1, f1, e1, hu1, uv1)
    MCU_V_FINISH(2, m2, n2, fo2, f2, e2, hu2, uv2)
#undef MCU_V_FINISH
  }
  on_exit();
}

template <class T, bool IS_D, bool VU = false>
__global__ void __launch_bounds__(kThreads, sizeof(T) == 8 ? MCU_C3DV_F64_CTAS : 3)
    k_curl3d_v(const Curl3DParams<T> *__restrict__ params, const Tile *__restrict__ tiles, int vmode) {
  __shared__ Curl3DParams<T> P;
  const Tile t = tiles[blockIdx.x];
  block_copy_params(&P, params + t.param);
  __syncthreads();
  curl3d_v_body<T, IS_D, VU>(P, t, vmode);
}

// ------------------------------------------------------------------------------------------------
// k_halfg_v: the general fused half step (UPML f_u / f_w / in-place sigma, chi1inv, Lorentzian P)
// with V consecutive points of the fastest axis per thread, one component per CTA -- the vector
// form of k_halfg<T,false,*>.  Conductivity and tiles that contain source points stay on the scalar
// kernel.  Per point the expression sequence is the one of finish_point, so results are identical.
// ------------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ void vload_shift(T *out, const T *base, long long i, long long s, const T *own,
                                            long long edge_index) {
  // out[j] = base[i + j + s] for a stride that is +1 / -1 along the vector axis; `own` = base[i..i+V);
  // edge_index = i+V / i-1, clamped by the caller where no owned point consumes that element
  constexpr int V = VecOf<T>::V;
  const T edge = __ldg(base + edge_index);
#pragma unroll
  for (int j = 0; j < V; ++j) {
    if (s > 0) out[j] = (j + 1 < V) ? own[j + 1 < V ? j + 1 : j] : edge;
    else out[j] = (j > 0) ? own[j > 0 ? j - 1 : 0] : edge;
  }
}

// UPML coefficients of one family (sig/kap/siginv indexed by k = 2*idx[axis] + off,
// meep_internals.hpp:241-247) for the V points of a thread: a = kap[k] - sig[k], b = siginv[k]
// (WFORM: a = kap + sig, b = kap - sig).  The same operations the reference does inline, done once.
template <class T, bool WFORM>
__device__ __forceinline__ void load_coef(const T *__restrict__ sig, const T *__restrict__ kap,
                                          const T *__restrict__ siginv, int axis, int off, int ia, int ib,
                                          int ic, T *a, T *b) {
  constexpr int V = VecOf<T>::V;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    if (j == 0 || axis == 2) {
      const int k = 2 * pick3(axis, ia, ib, ic + j) + off;
      const T kp = __ldg(kap + k), sg = __ldg(sig + k);
      if (WFORM) { a[j] = kp + sg; b[j] = kp - sg; }
      else { a[j] = kp - sg; b[j] = __ldg(siginv + k); }
    }
    else { a[j] = a[0]; b[j] = b[0]; }
  }
}

template <class T> struct HalfgVals { // the operands of one thread-iteration of k_halfg_v
  T g1[VecOf<T>::V], g1p[VecOf<T>::V], g2[VecOf<T>::V], g2p[VecOf<T>::V], fo[VecOf<T>::V], fuo[VecOf<T>::V],
      eo[VecOf<T>::V], fwo[VecOf<T>::V], uv[VecOf<T>::V];
};

// (Software pipelining -- loading the operands of iteration n+1 into a second register set before
// iteration n stores -- was measured and rejected: the registers cost a resident CTA, -4...-16 %.)
template <class T, bool HAS_POL, class OnExit = NoExit>
__device__ __forceinline__ void halfg_v_body(const HalfHeader<T> &H, const CompParams<T> &Csm, const Tile t,
                                             const OnExit on_exit = OnExit()) {
  constexpr int V = VecOf<T>::V;
  const int pv = H.pitch / V;
  const int p = t.p0 + threadIdx.x;
  const int io = p / pv, ic = (p - io * pv) * V;
  const int march = H.march;
  const int n_other = march ? H.nA : H.nB;
  if (io >= n_other || ic >= H.nC) { on_exit(); return; }
  const CompParams<T> c = Csm;
  const int ax_o = march ? 0 : 1, ax_m = march ? 1 : 0;
  const int lo_o = march ? c.lo[0] : c.lo[1], hi_o = march ? c.hi[0] : c.hi[1];
  const int lo_m = march ? c.lo[1] : c.lo[0], hi_m = march ? c.hi[1] : c.hi[0];
  if (!(io >= lo_o && io <= hi_o)) { on_exit(); return; }
  unsigned own = 0;
#pragma unroll
  for (int j = 0; j < V; ++j)
    if (ic + j >= c.lo[2] && ic + j <= c.hi[2]) own |= 1u << j;
  if (!own) { on_exit(); return; }
  const unsigned full = (1u << V) - 1u;
  // metal planes: the part that does not change along the march, per element
  unsigned zfix_db = 0, zfix_eh = 0;
  {
    const bool o_db = (((c.zmask_db >> ax_o) & 1u) && io == lo_o) || (((c.zmask_db >> (3 + ax_o)) & 1u) && io == hi_o);
    const bool o_eh = (((c.zmask_eh >> ax_o) & 1u) && io == lo_o) || (((c.zmask_eh >> (3 + ax_o)) & 1u) && io == hi_o);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int z = ic + j;
      if (o_db || (((c.zmask_db >> 2) & 1u) && z == c.lo[2]) || (((c.zmask_db >> 5) & 1u) && z == c.hi[2])) zfix_db |= 1u << j;
      if (o_eh || (((c.zmask_eh >> 2) & 1u) && z == c.lo[2]) || (((c.zmask_eh >> 5) & 1u) && z == c.hi[2])) zfix_eh |= 1u << j;
    }
  }
  const bool zlo_db = (c.zmask_db >> ax_m) & 1u, zhi_db = (c.zmask_db >> (3 + ax_m)) & 1u;
  const bool zlo_eh = (c.zmask_eh >> ax_m) & 1u, zhi_eh = (c.zmask_eh >> (3 + ax_m)) & 1u;
  const long long s_m = march ? H.sB : H.sA, s_o = march ? H.sA : H.sB;
  const int m0 = t.a0 > lo_m ? t.a0 : lo_m;
  const int m1 = t.a1 - 1 < hi_m ? t.a1 - 1 : hi_m;
  long long i = (long long)m0 * s_m + (long long)io * s_o + ic;
  const bool sh1 = c.s1 == 1 || c.s1 == -1, sh2 = c.g2 && (c.s2 == 1 || c.s2 == -1);
  const T dtdx = c.dtdx;
  // UPML coefficients: constant along the march unless the PML axis IS the march axis (thin slabs
  // march along their long axis, so that is the rare case) -> loaded once, kept in registers
  T cA[V], cB[V], uA[V], uB[V], wP[V], wM[V];
  const bool has_k = c.kaxis >= 0, has_ku = c.kuaxis >= 0, has_kw = c.e && c.fw;
  const bool var_k = has_k && c.kaxis == ax_m, var_ku = has_ku && c.kuaxis == ax_m, var_kw = has_kw && c.kwaxis == ax_m;
  {
    const int ia0 = march ? io : m0, ib0 = march ? m0 : io;
    if (has_k && !var_k) load_coef<T, false>(c.sig, c.kap, c.siginv, c.kaxis, c.koff, ia0, ib0, ic, cA, cB);
    if (has_ku && !var_ku) load_coef<T, false>(c.sigu, c.kapu, c.siginvu, c.kuaxis, c.kuoff, ia0, ib0, ic, uA, uB);
    if (has_kw && !var_kw) load_coef<T, true>(c.sigw, c.kapw, c.kapw, c.kwaxis, c.kwoff, ia0, ib0, ic, wP, wM);
  }
  const int pitch = H.pitch;
  auto gather = [&](long long ii, HalfgVals<T> &v) __attribute__((always_inline)) {
    vload<T>(v.g1, c.g1 + ii);
    if (sh1) vload_shift<T>(v.g1p, c.g1, ii, c.s1, v.g1, c.s1 > 0 ? (ic + V < pitch ? ii + V : ii) : (ic > 0 ? ii - 1 : ii));
    else vload<T>(v.g1p, c.g1 + ii + c.s1);
    if (c.g2) {
      vload<T>(v.g2, c.g2 + ii);
      if (sh2) vload_shift<T>(v.g2p, c.g2, ii, c.s2, v.g2, c.s2 > 0 ? (ic + V < pitch ? ii + V : ii) : (ic > 0 ? ii - 1 : ii));
      else vload<T>(v.g2p, c.g2 + ii + c.s2);
    }
    vload_rw<T>(v.fo, c.f + ii);
    if (c.fu) vload_rw<T>(v.fuo, c.fu + ii);
    if (c.e) {
      if (c.u) vload<T>(v.uv, c.u + ii);
      if (c.fw) {
        vload_rw<T>(v.fwo, c.fw + ii);
        vload_rw<T>(v.eo, c.e + ii);
      }
    }
  };
  HalfgVals<T> cur;
  for (int im = m0; im <= m1; ++im, i += s_m) {
    const int ia = march ? io : im, ib = march ? im : io;
    if (var_k) load_coef<T, false>(c.sig, c.kap, c.siginv, c.kaxis, c.koff, ia, ib, ic, cA, cB);
    if (var_ku) load_coef<T, false>(c.sigu, c.kapu, c.siginvu, c.kuaxis, c.kuoff, ia, ib, ic, uA, uB);
    if (var_kw) load_coef<T, true>(c.sigw, c.kapw, c.kapw, c.kwaxis, c.kwoff, ia, ib, ic, wP, wM);
    if (pf_planes<T>() > 0 && im + pf_planes<T>() <= m1) {
      const long long ipf = i + pf_planes<T>() * s_m;
      prefetch_l2(c.g1 + ipf);
      if (c.g2) prefetch_l2(c.g2 + ipf);
      prefetch_l2(c.f + ipf);
      if (c.fu) prefetch_l2(c.fu + ipf);
      if (c.e) {
        if (c.u) prefetch_l2(c.u + ipf);
        if (c.fw) {
          prefetch_l2(c.fw + ipf);
          prefetch_l2(c.e + ipf);
        }
        if (HAS_POL) {
          for (int ip = 0; ip < c.npol; ++ip)
            if (Csm.pol[ip].p) {
              prefetch_l2(Csm.pol[ip].p + ipf);
              if (Csm.pol[ip].s) {
                prefetch_l2(Csm.pol[ip].pp + ipf);
                prefetch_l2(Csm.pol[ip].s + ipf);
              }
            }
        }
      }
    }
    // ---------------- gather ----------------
    gather(i, cur);
    T *const g1 = cur.g1, *const g1p = cur.g1p, *const g2 = cur.g2, *const g2p = cur.g2p, *const fo = cur.fo,
             *const fuo = cur.fuo, *const eo = cur.eo, *const fwo = cur.fwo, *const uv = cur.uv;
    // ---------------- step_curl per element (step_generic.cpp:89-252) ----------------
    const unsigned zdb = zfix_db | (((zlo_db && im == lo_m) || (zhi_db && im == hi_m)) ? full : 0u);
    const unsigned zeh = zfix_eh | (((zlo_eh && im == lo_m) || (zhi_eh && im == hi_m)) ? full : 0u);
    T fn[V], fun[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const T curl = c.g2 ? (g1p[j] - g1[j] + g2[j] - g2p[j]) : (g1p[j] - g1[j]);
      const T fold = fo[j];
      T fnew;
      if (!has_k) {
        if (!has_ku)
          fnew = fold - dtdx * curl;
        else {
          const T fprev = fuo[j];
          const T fu_ = fprev - dtdx * curl;
          fun[j] = fu_;
          fnew = uB[j] * (uA[j] * fold + fu_ - fprev);
        }
      }
      else {
        if (!has_ku)
          fnew = (cA[j] * fold - dtdx * curl) * cB[j];
        else {
          const T fprev = fuo[j];
          const T fu_ = (cA[j] * fprev - dtdx * curl) * cB[j];
          fun[j] = fu_;
          fnew = uB[j] * (uA[j] * fold + fu_ - fprev);
        }
      }
      if ((zdb >> j) & 1u) fnew = 0;
      const bool on = (own >> j) & 1u;
      fn[j] = on ? fnew : fold;
      if (c.fu && !on) fun[j] = fuo[j];
    }
    vstore<T>(c.f + i, fn);
    if (c.fu) vstore<T>(c.fu + i, fun);
    if (!c.e) continue;
    // ---------------- update_eh + update_pols per element ----------------
    T fmp[V];
#pragma unroll
    for (int j = 0; j < V; ++j) fmp[j] = fn[j];
    if (HAS_POL) {
      for (int ip = 0; ip < c.npol; ++ip)
        if (Csm.pol[ip].p) {
          T pv_[V];
          vload_rw<T>(pv_, Csm.pol[ip].p + i);
#pragma unroll
          for (int j = 0; j < V; ++j) fmp[j] -= pv_[j];
        }
    }
    T gu[V], en[V];
#pragma unroll
    for (int j = 0; j < V; ++j) gu[j] = c.u ? (fmp[j] * uv[j]) : fmp[j];
    if (c.fw) {
      T fwn[V];
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const bool on = (own >> j) & 1u;
        T enew = eo[j] + (wP[j] * gu[j] - wM[j] * fwo[j]);
        if ((zeh >> j) & 1u) enew = 0;
        en[j] = on ? enew : eo[j];
        fwn[j] = on ? gu[j] : fwo[j];
      }
      vstore<T>(c.fw + i, fwn);
      vstore<T>(c.e + i, en);
    }
    else {
#pragma unroll
      for (int j = 0; j < V; ++j) en[j] = ((zeh >> j) & 1u) ? (T)0 : gu[j];
      if (own == full)
        vstore<T>(c.e + i, en);
      else {
#pragma unroll
        for (int j = 0; j < V; ++j)
          if ((own >> j) & 1u) c.e[i + j] = en[j];
      }
    }
    if (HAS_POL) {
      for (int ip = 0; ip < c.npol; ++ip)
        if (Csm.pol[ip].p && Csm.pol[ip].s) { // susceptibility.cpp:252-259; W = f_w ?: f = gu before zero_metal
          const PolParams<T> q = Csm.pol[ip];
          T pc[V], pp[V], sv[V], pn[V], ppn[V];
          vload_rw<T>(pc, q.p + i);
          vload_rw<T>(pp, q.pp + i);
          vload<T>(sv, q.s + i);
#pragma unroll
          for (int j = 0; j < V; ++j) {
            const bool on = (own >> j) & 1u;
            const T pnew = q.gamma1inv * (pc[j] * (2 - q.omega0dtsqr_denom) - q.gamma1 * pp[j] +
                                          q.omega0dtsqr * (sv[j] * gu[j]));
            pn[j] = on ? pnew : pc[j];
            ppn[j] = on ? pc[j] : pp[j];
          }
          vstore<T>(q.p + i, pn);
          vstore<T>(q.pp + i, ppn);
        }
    }
  }
  on_exit();
}

template <class T, bool HAS_POL>
__global__ void __launch_bounds__(kThreads, sizeof(T) == 8 ? 2 : 3)
    k_halfg_v(const HalfParams<T> *__restrict__ params, const Tile *__restrict__ tiles) {
  __shared__ HalfHeader<T> H;
  __shared__ CompParams<T> Csm;
  const Tile t = tiles[blockIdx.x];
  const int pidx = t.param >> 2, kcomp = t.param & 3;
  block_copy_params(&H, reinterpret_cast<const HalfHeader<T> *>(params + pidx));
  block_copy_params(&Csm, &params[pidx].comp[kcomp]);
  __syncthreads();
  halfg_v_body<T, HAS_POL>(H, Csm, t);
}

// ------------------------------------------------------------------------------------------------
// step_boundaries: dst[n] = src[n] (COPY), = -src[n] (NEGATE), complex phase multiply (PHASE).
// Sources are owned points and destinations not-owned points, so gather and scatter can be one
// pass (the reference stages through comm_blocks only because of MPI).
// ------------------------------------------------------------------------------------------------
template <class T> struct ConnTable {
  T *const *dst;
  const T *const *src;
  const T *phase; // (re,im) per complex transfer, PHASE only
  long long n;    // number of realnum transfers (PHASE: 2 per complex)
  const T *const *mul; // per transfer: chi1inv element the source is multiplied by (virtual-E sources), or NULL table
};
template <class T> __device__ __forceinline__ T conn_value(const T *src, const T *mul) {
  const T v = *src;
  return mul ? v * *mul : v; // E = chi1inv * D of a chunk that does not store E (see Curl3DParams::virt)
}

template <class T>
__global__ void __launch_bounds__(kThreads)
    k_conn(ConnTable<T> phase, ConnTable<T> negate, ConnTable<T> copy) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long long n = t0; n < copy.n; n += stride) *copy.dst[n] = *copy.src[n];
  for (long long n = t0; n < negate.n; n += stride) *negate.dst[n] = -*negate.src[n];
  for (long long n = t0; n < phase.n / 2; n += stride) { // step.cpp:188-194
    const T pr = phase.phase[2 * n], pi = phase.phase[2 * n + 1];
    const T br = conn_value<T>(phase.src[2 * n], phase.mul ? phase.mul[2 * n] : 0),
            bi = conn_value<T>(phase.src[2 * n + 1], phase.mul ? phase.mul[2 * n + 1] : 0);
    *phase.dst[2 * n] = pr * br - pi * bi;
    *phase.dst[2 * n + 1] = pr * bi + pi * br;
  }
}

// Run-length form of the COPY / NEGATE connection lists: the reference enumerates not-owned points
// plane by plane (LOOP_OVER_VOL_NOTOWNED, src/boundaries.cpp:396), so consecutive entries advance
// by a constant stride on both sides.  One warp copies one run: 16 bytes of pointers per RUN instead
// of per element.
template <class T> struct ConnRun {
  T *dst;
  const T *src;
  int n, dst_stride, src_stride, negate;
  const T *mul; // NULL, or the chi1inv element of src[0] (same stride as src): virtual-E sources
};
constexpr int kConnPiece = 128; // runs are cut into pieces of at most this many elements
template <class T>
__global__ void __launch_bounds__(kThreads) k_conn_runs(const ConnRun<T> *__restrict__ runs, int nruns) {
  // one warp per piece, four independent loads per lane in flight before the first store
  const int lane = threadIdx.x & 31;
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= nruns) return;
  const ConnRun<T> q = runs[r];
  T v[kConnPiece / 32];
#pragma unroll
  for (int k = 0; k < kConnPiece / 32; ++k) {
    const int j = lane + 32 * k;
    if (j < q.n) {
      v[k] = q.src[(long long)j * q.src_stride];
      if (q.mul) v[k] = v[k] * q.mul[(long long)j * q.src_stride];
    }
  }
#pragma unroll
  for (int k = 0; k < kConnPiece / 32; ++k) {
    const int j = lane + 32 * k;
    if (j < q.n) q.dst[(long long)j * q.dst_stride] = q.negate ? -v[k] : v[k];
  }
}

// ------------------------------------------------------------------------------------------------
// k_mix: clean interior tiles (curl3d_v_body) and vectorised PML tiles (halfg_v_body, kTileGeneral)
// in ONE grid, interleaved in the tile table in proportion to their numbers.  Launched as two
// kernels, the latency-bound PML CTAs and the bandwidth-bound interior CTAs mostly run one after
// the other (whichever kernel the block scheduler prefers fills every SM); in one grid every SM
// holds both kinds at any time, so the interior tiles keep HBM busy while the PML tiles wait on
// their dependent loads.
// (Tried and rejected, measured: also folding the halo copies and the scalar source tiles into the
// same grid behind a "compute CTAs done" counter -- the per-thread completion signal and the extra
// bodies cost ~600 bytes of spills at 80 registers and the step went from 0.385 to 0.58 ms.)
// ------------------------------------------------------------------------------------------------
constexpr int kTileGeneral = 1 << 29; // Tile::p0 flag inside a k_mix table

#ifndef MCU_MIX_F32_CTAS
#define MCU_MIX_F32_CTAS 3
#endif
#ifndef MCU_MIX_F64_CTAS
#define MCU_MIX_F64_CTAS 2
#endif
template <class T, bool IS_D, bool VU = false>
__global__ void __launch_bounds__(kThreads, sizeof(T) == 8 ? MCU_MIX_F64_CTAS : MCU_MIX_F32_CTAS)
    k_mix(const Curl3DParams<T> *__restrict__ cparams, const HalfParams<T> *__restrict__ hparams,
          const Tile *__restrict__ tiles, int vmode) {
  struct G { HalfHeader<T> H; CompParams<T> C; };
  constexpr size_t kBytes = sizeof(Curl3DParams<T>) > sizeof(G) ? sizeof(Curl3DParams<T>) : sizeof(G);
  __shared__ __align__(16) unsigned char raw[kBytes];
  Tile t = tiles[blockIdx.x];
  if (t.p0 & kTileGeneral) {
    t.p0 &= ~kTileGeneral;
    G *g = reinterpret_cast<G *>(raw);
    const int pidx = t.param >> 2, kcomp = t.param & 3;
    block_copy_params(&g->H, reinterpret_cast<const HalfHeader<T> *>(hparams + pidx));
    block_copy_params(&g->C, &hparams[pidx].comp[kcomp]);
    __syncthreads();
    halfg_v_body<T, false>(g->H, g->C, t);
  }
  else {
    Curl3DParams<T> *P = reinterpret_cast<Curl3DParams<T> *>(raw);
    block_copy_params(P, cparams + t.param);
    __syncthreads();
    curl3d_v_body<T, IS_D, VU>(*P, t, vmode);
  }
}

// Halo exchange with another rank (replaces the MPI_Isend/Irecv pairs of src/mympi.cpp:107-135):
// k_pack gathers the outgoing values of one field type into one contiguous buffer per peer
// (order PHASE | NEGATE | COPY, src/step.cpp:257-268), NCCL moves the buffers over NVLink, and
// k_unpack applies process_incoming_chunk_data (src/step.cpp:173-224) on the receiving side.
template <class T>
__global__ void __launch_bounds__(kThreads) k_pack(const T *const *__restrict__ src, T *__restrict__ buf, long long n,
                                                   const T *const *__restrict__ mul = 0) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride)
    buf[k] = conn_value<T>(src[k], mul ? mul[k] : 0);
}

// virtual E made real again (host syncs): e = u * d over the whole array (u is 1 at not-owned points, where d holds
// the E halo), then the dipoles of the integrated sources (k_virt_src)
template <class T> __global__ void __launch_bounds__(kThreads) k_virt_materialise(T *__restrict__ e, const T *__restrict__ d, const T *__restrict__ u, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) e[i] = u ? d[i] * u[i] : d[i];
}
template <class T> __global__ void k_virt_src(T *__restrict__ e, const T *__restrict__ d, const T *__restrict__ u, const SrcEntry *__restrict__ tab, int n,
                                              int cmp, const double *__restrict__ smp_prev) {
  // one thread: the sources of a point are applied in list order (tab is sorted by index, stable)
  if (blockIdx.x || threadIdx.x) return;
  int j = 0;
  while (j < n) {
    const long long idx = tab[j].idx;
    T v = d[idx];
    for (; j < n && tab[j].idx == idx; ++j) {
      const double2 dp = sample_dipole(smp_prev, tab[j].tid, 1);
      const double2 A = cmul(tab[j].are, tab[j].aim, dp.x, dp.y);
      v = (T)((double)v - (cmp ? A.y : A.x));
    }
    e[idx] = u ? v * u[idx] : v;
  }
}
// device copy of a chi1inv array: 1 at the not-owned points of its component (see Curl3DParams::virt)
template <class T> __global__ void __launch_bounds__(kThreads) k_virt_patch_u(T *__restrict__ u, int nA, int nB, int pitch, int lo0, int hi0, int lo1, int hi1, int lo2, int hi2) {
  const long long n = (long long)nA * nB * pitch, stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int c = (int)(i % pitch);
    const long long r = i / pitch;
    const int b = (int)(r % nB), a = (int)(r / nB);
    if (a < lo0 || a > hi0 || b < lo1 || b > hi1 || c < lo2 || c > hi2) u[i] = (T)1;
  }
}
// gather with the same rule (meepcu_read_values on a virtual E array)
template <class T>
__global__ void __launch_bounds__(kThreads) k_gather(const T *const *__restrict__ src, const T *const *__restrict__ mul, T *__restrict__ buf, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) buf[k] = conn_value<T>(src[k], mul[k]);
}
template <class T>
__global__ void __launch_bounds__(kThreads)
    k_unpack(T *const *__restrict__ dst, const T *__restrict__ buf, const T *__restrict__ phase,
             const long long *__restrict__ seg, int nseg) {
  // seg: per peer segment {offset, n_phase(realnums), n_negate, n_copy, phase_offset(complex)}
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (int sgi = 0; sgi < nseg; ++sgi) {
    const long long off = seg[5 * sgi], np = seg[5 * sgi + 1], nn = seg[5 * sgi + 2], nc = seg[5 * sgi + 3],
                    po = seg[5 * sgi + 4];
    for (long long k = t0; k < np / 2; k += stride) {
      const T pr = phase[2 * (po + k)], pi = phase[2 * (po + k) + 1];
      const T br = buf[off + 2 * k], bi = buf[off + 2 * k + 1];
      *dst[off + 2 * k] = pr * br - pi * bi;
      *dst[off + 2 * k + 1] = pr * bi + pi * br;
    }
    for (long long k = t0; k < nn; k += stride) *dst[off + np + k] = -buf[off + np + k];
    for (long long k = t0; k < nc; k += stride) *dst[off + np + nn + k] = buf[off + np + nn + k];
  }
}

// zero_metal for lists that are not fused into k_half (general fallback)
template <class T> __global__ void k_zero(T *const *ptrs, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride)
    *ptrs[k] = 0;
}

// ------------------------------------------------------------------------------------------------
// f_minus_p = D - sum P - integrated dipoles over the whole array incl. not-owned points
// (update_eh.cpp:122-146), needed only where the E update reads neighbours of D-P.
// ------------------------------------------------------------------------------------------------
template <class T> struct FmpParams {
  T *fmp;
  const T *d;
  const T *p[kMaxPol];
  int npol;
  long long n; // device elements
  const SrcEntry *isrc;
  int nisrc;
  int is_D, cmp;
};
template <class T>
__global__ void __launch_bounds__(kThreads)
    k_fmp(const FmpParams<T> *__restrict__ tab, int ntab, const double *__restrict__ smp) {
  for (int it = blockIdx.y; it < ntab; it += gridDim.y) {
    const FmpParams<T> q = tab[it];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < q.n; i += stride) {
      T v = q.d[i];
      for (int ip = 0; ip < q.npol; ++ip)
        if (q.p[ip]) v -= q.p[ip][i];
      q.fmp[i] = v;
    }
  }
}
template <class T>
__global__ void k_fmp_src(const FmpParams<T> *__restrict__ tab, int ntab,
                          const double *__restrict__ smp) {
  // one thread per table; integrated sources are few and must be applied in list order
  const int it = blockIdx.x * blockDim.x + threadIdx.x;
  if (it >= ntab) return;
  const FmpParams<T> q = tab[it];
  for (int j = 0; j < q.nisrc; ++j) {
    const SrcEntry e = q.isrc[j];
    const double2 dp = sample_dipole(smp, e.tid, q.is_D);
    const double2 A = cmul(e.are, e.aim, dp.x, dp.y);
    q.fmp[e.idx] = (T)((double)q.fmp[e.idx] - (q.cmp ? A.y : A.x));
  }
}

// Where f_minus_p would be D minus dipoles only (no polarisation), it is not materialised: the E update reads D
// itself and the handful of dipole points are patched in place around it -- saved, replaced by
// (T)((double)D - dipoles) in list order like update_eh.cpp:136-146, and put back bit for bit afterwards.
template <class T> struct DipPatch {
  T *d;
  const SrcEntry *isrc; // sorted by index (stable in reference order)
  int n, is_D, cmp;
  T *saved;             // n slots; the slot of the first entry of an index holds the original value
};
template <class T>
__global__ void k_dip_patch(const DipPatch<T> *__restrict__ tab, int ntab, const double *__restrict__ smp, int restore) {
  const int it = blockIdx.x * blockDim.x + threadIdx.x;
  if (it >= ntab) return;
  const DipPatch<T> q = tab[it];
  for (int j = 0; j < q.n; ++j) {
    const SrcEntry e = q.isrc[j];
    const bool first = j == 0 || q.isrc[j - 1].idx != e.idx;
    if (restore) {
      if (first) q.d[e.idx] = q.saved[j];
      continue;
    }
    if (first) q.saved[j] = q.d[e.idx];
    const double2 dp = sample_dipole(smp, e.tid, q.is_D);
    const double2 A = cmul(e.are, e.aim, dp.x, dp.y);
    q.d[e.idx] = (T)((double)q.d[e.idx] - (q.cmp ? A.y : A.x));
  }
}

// ------------------------------------------------------------------------------------------------
// step_update_EDHB in full generality for one component (step_generic.cpp:576-906), plus the
// isotropic polarisation update and zero_metal(E), for chunks whose E update reads neighbours.
// ------------------------------------------------------------------------------------------------
template <class T> struct EhParams {
  int nA, nB, nC, pitch;
  long long sA, sB;
  int lo[3], hi[3];
  T *f;               // E (or H) component
  int fc;             // component id (0 Ex, 1 Ey, 4 Ez): selects the NR argument order
  const T *g, *g1, *g2; // dmp[dc], dmp[dc_1], dmp[dc_2] after the swap rule
  const T *u, *u1, *u2;
  long long s, s1, s2;
  const T *chi2, *chi3;
  T *fw;
  const T *sigw, *kapw;
  int kwaxis, kwoff;
  int nl_mode;
  unsigned zmask_eh;
  int npol;
  PolParams<T> pol[kMaxPol];
  // per-point class of the update, derived once from the (immutable) material arrays by k_eh_mask: 0 = the plain
  // product val = g*u (no neighbour of D-P, no chi2 / chi3 / off-diagonal flag and no old E is read), 1 = the full
  // formula.  NULL: not built (every point takes the full path).
  const unsigned char *mask;
  int *nr_fail; // device flags: [0] cells whose NR seed ladder was exhausted, [2] count solves?, [3] longest solve (iterations),
                // [4..7] as two 64-bit counters: solves, iterations
};

template <class T>
__device__ __forceinline__ T calc_nonlinear_u(const T Dsqr, const T Di, const T chi1inv,
                                              const T chi2, const T chi3) { // step_generic.cpp:546-553
  T c2 = Di * chi2 * (chi1inv * chi1inv);
  T c3 = Dsqr * chi3 * (chi1inv * chi1inv * chi1inv);
  return (1 + c2 + 2 * c3) / (1 + 2 * c2 + 3 * c3);
}

// Newton-Raphson for  A_k = B_k v_k + chi2 * (product of the other two)   (newton_raphson.cpp).
// For the zinc-blende tensor only F of eq.1, G of eq.2 and H of eq.3 are non-zero and equal to
// chi2, but the arithmetic below keeps the reference's general expressions (with the zero
// coefficients written out) so that every rounding is the same.
template <class T> struct NRP { T A, B, F, G, H; };

template <class T>
__device__ __forceinline__ void nr_equations(double x, double y, double z, const NRP<T> &p1,
                                             const NRP<T> &p2, const NRP<T> &p3, double F[3]) {
  F[0] = p1.A - (p1.B * x + p1.F * y * z + p1.G * x * z + p1.H * x * y);
  F[1] = p2.A - (p2.B * y + p2.F * y * z + p2.G * x * z + p2.H * x * y);
  F[2] = p3.A - (p3.B * z + p3.F * y * z + p3.G * x * z + p3.H * x * y);
}

template <class T> __device__ __forceinline__ bool bits_equal(T a, T b);
template <> __device__ __forceinline__ bool bits_equal<double>(double a, double b) {
  return __double_as_longlong(a) == __double_as_longlong(b);
}
template <> __device__ __forceinline__ bool bits_equal<float>(float a, float b) {
  return __float_as_int(a) == __float_as_int(b);
}

template <class T>
__device__ bool nr_iterate(T x, T y, T z, const NRP<T> &p1, const NRP<T> &p2, const NRP<T> &p3,
                           T out[3], double tol1, double tol2, double tol3, int max_iter, int &iters) {
  T px1 = x, py1 = y, pz1 = z, px2 = x, py2 = y, pz2 = z;
  bool dead = false;
  for (int iter = 0; iter < max_iter && !dead; iter++) { // newton_raphson.cpp:93-142
    ++iters;
    double F[3], J[3][3], b[3], d[3];
    nr_equations<T>(x, y, z, p1, p2, p3, F);
    J[0][0] = -p1.B - p1.G * z - p1.H * y;
    J[0][1] = -p1.F * z - p1.H * x;
    J[0][2] = -p1.F * y - p1.G * x;
    J[1][0] = -p2.G * z - p2.H * y;
    J[1][1] = -p2.B - p2.F * z - p2.H * x;
    J[1][2] = -p2.F * y - p2.G * x;
    J[2][0] = -p3.G * z - p3.H * y;
    J[2][1] = -p3.F * z - p3.H * x;
    J[2][2] = -p3.B - p3.F * y - p3.G * x;
    b[0] = F[0]; b[1] = F[1]; b[2] = F[2];
    // solveLinearSystem :170-194 (no pivoting)
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = i + 1; j < 3; j++) {
        const double factor = J[j][i] / J[i][i];
#pragma unroll
        for (int k = i; k < 3; k++) J[j][k] -= factor * J[i][k];
        b[j] -= factor * b[i];
      }
#pragma unroll
    for (int i = 2; i >= 0; i--) {
      d[i] = b[i];
#pragma unroll
      for (int j = i + 1; j < 3; j++) d[i] -= J[i][j] * d[j];
      d[i] /= J[i][i];
    }
    x -= d[0];
    y -= d[1];
    z -= d[2];
    // The iteration is a deterministic map of (x,y,z): once a state repeats (fixed point or
    // 2-cycle) or becomes NaN, no later iteration can pass the tests below, so the remaining
    // iterations of this attempt are skipped -- same outcome as running all max_iter of them.
    {
      const bool same1 = bits_equal<T>(x, px1) && bits_equal<T>(y, py1) && bits_equal<T>(z, pz1);
      const bool same2 = bits_equal<T>(x, px2) && bits_equal<T>(y, py2) && bits_equal<T>(z, pz2);
      dead = same1 || same2 || x != x || y != y || z != z;
      px2 = px1; py2 = py1; pz2 = pz1;
      px1 = x; py1 = y; pz1 = z;
    }
    if (fabs(d[0]) < tol1 && fabs(d[1]) < tol2 && fabs(d[2]) < tol3) {
      const double ax = fabs((double)x), ay = fabs((double)y), az = fabs((double)z);
      const double chk = (ax > ay ? (ax > az ? ax : az) : (ay > az ? ay : az)) * 1e-4;
      double Fc[3];
      nr_equations<T>(x, y, z, p1, p2, p3, Fc);
      if (Fc[0] <= chk && Fc[1] <= chk && Fc[2] <= chk) {
        out[0] = x; out[1] = y; out[2] = z;
        return true;
      }
    }
  }
  return false;
}

// runNR :209-359 with the deterministic part of the seed ladder
template <class T>
__device__ bool run_nr(T seed1, T seed2, T seed3, T out[3], const NRP<T> &p1, const NRP<T> &p2,
                       const NRP<T> &p3, int &iters) {
  const double TOL = 1e-8, seedMax = 1e33;
  int max_iter = 250;
  const double tol1 = fmax(fabs(TOL * out[0]) * 0.0001, TOL);
  const double tol2 = fmax(fabs(TOL * out[1]) * 0.0001, TOL);
  const double tol3 = fmax(fabs(TOL * out[2]) * 0.0001, TOL);
  double s1 = seed1, s2 = seed2, s3 = seed3;
  // After these 14 deterministic seed variants the reference draws up to 86 more seeds from
  // std::random_device (parity undefined, and from far-away seeds Newton's method lands on the
  // spurious large-field roots of the quadratic system).  We stop here and leave the field value
  // unchanged, which is also what the reference does when all of its attempts fail.
  for (int i = 0; i < 15; ++i) {
    if (nr_iterate<T>((T)s1, (T)s2, (T)s3, p1, p2, p3, out, tol1, tol2, tol3, max_iter, iters)) return true;
    switch (i) {
      case 0: s1 = seed1 * seedMax; max_iter = 600; break;
      case 1: s1 = seed1; s2 = seed2 * seedMax; break;
      case 2: s2 = seed2; s3 = seed3 * seedMax; break;
      case 3: s3 = seed3; s1 = -seed1 * seedMax; break;
      case 4: s1 = seed1; s2 = -seed2 * seedMax; break;
      case 5: s2 = seed2; s3 = -seed3 * seedMax; break;
      case 6: s1 = seed1 * seedMax; s2 = seed2 * seedMax; s3 = seed3; break;
      case 7: s1 = seed1 * seedMax; s2 = seed2; s3 = seed3 * seedMax; break;
      case 8: s1 = seed1; s2 = seed2 * seedMax; s3 = seed3 * seedMax; break;
      case 9: s1 = -seed1 * seedMax; s2 = -seed2 * seedMax; s3 = seed3; break;
      case 10: s1 = -seed1 * seedMax; s2 = seed2; s3 = -seed3 * seedMax; break;
      case 11: s1 = seed1; s2 = -seed2 * seedMax; s3 = -seed3 * seedMax; break;
      case 12: s1 = seed1 * seedMax; s2 = seed2 * seedMax; s3 = seed3 * seedMax; break;
      case 13: s1 = -seed1 * seedMax; s2 = -seed2 * seedMax; s3 = -seed3 * seedMax; break;
      default: return false;
    }
  }
  return false;
}

// susceptibility.cpp:185-186
#define MCU_POFFDIAG(u, g, sx, s)                                                                  \
  (0.25 * ((g[i] + g[i - sx]) * u[i] + (g[i + s] + g[(i + s) - sx]) * u[i + s]))

#define MCU_OFFDIAG(u, g, sx)                                                                      \
  (0.25 * ((g[i] + g[i - sx]) * u[i] + (g[i + s] + g[(i + s) - sx]) * u[i + s]))

template <class T>
__device__ __forceinline__ void eh_point(const EhParams<T> &q, long long i, int ia, int ib, int ic) {
  const T *g = q.g, *g1 = q.g1, *g2 = q.g2, *u = q.u, *u1 = q.u1, *u2 = q.u2;
  const long long s = q.s, s1 = q.s1, s2 = q.s2;
  const bool pade = q.nl_mode == 1;
  T val; // u*g with whatever nonlinear / off-diagonal terms apply
  bool done_nr = false;
  if (!pade) {
    if (!q.fw && u1 && u2 && q.chi3) { // the fork's NR branch, step_generic.cpp:732-816
      const T gs = g[i];
      const T ui = u[i];
      // :757-764: chi2 == 0, or more than one of u, u1, u2 zero -> the linear product.  The test only involves
      // material arrays, so it is evaluated once (k_eh_mask) and the 8 neighbour loads, chi2, u1, u2 and the old E
      // are skipped where it holds (two thirds of BASELINE config 3's cells)
      bool plain;
      T chi2new = 0;
      if (q.mask) plain = q.mask[i] == 0;
      else {
        chi2new = q.chi2[i];
        int zeroEps = 0;
        if (ui == 0) zeroEps += 1;
        if (u1[i] == 0) zeroEps += 1;
        if (u2[i] == 0) zeroEps += 1;
        plain = chi2new == 0 || zeroEps > 1;
      }
      if (plain) { val = (gs * ui); }
      else {
        if (q.mask) chi2new = q.chi2[i];
        const T gs_2 = (g1[i] + g1[i + s] + g1[i - s1] + g1[i + (s - s1)]) * 0.25;
        const T gs_3 = (g2[i] + g2[i + s] + g2[i - s2] + g2[i + (s - s2)]) * 0.25;
        const T us = 1 / ui;
        T out[3];
        NRP<T> p1, p2, p3;
        T sd1, sd2, sd3;
        const T z0 = 0.0;
        int own;
        if (q.fc == 0) {
          p1 = {gs, us, chi2new, z0, z0};
          p2 = {gs_2, us, z0, chi2new, z0};
          p3 = {gs_3, us, z0, z0, chi2new};
          sd1 = q.f[i]; sd2 = gs_2 * ui; sd3 = gs_3 * ui;
          own = 0;
        }
        else if (q.fc == 1) {
          p1 = {gs_3, us, chi2new, z0, z0};
          p2 = {gs, us, z0, chi2new, z0};
          p3 = {gs_2, us, z0, z0, chi2new};
          sd1 = gs_3 * ui; sd2 = q.f[i]; sd3 = gs_2 * ui;
          own = 1;
        }
        else {
          p1 = {gs_2, us, chi2new, z0, z0};
          p2 = {gs_3, us, z0, chi2new, z0};
          p3 = {gs, us, z0, z0, chi2new};
          sd1 = gs_2 * ui; sd2 = gs_3 * ui; sd3 = q.f[i];
          own = 2;
        }
        const T fprev_ = q.f[i];
        out[0] = own == 0 ? fprev_ : (T)0; // *fw of this component = previous field; the dummies are 0
        out[1] = own == 1 ? fprev_ : (T)0;
        out[2] = own == 2 ? fprev_ : (T)0;
        int iters = 0;
        const bool ok = run_nr<T>(sd1, sd2, sd3, out, p1, p2, p3, iters);
        if (!ok) {
          atomicAdd(q.nr_fail, 1);
          out[0] = out[1] = out[2] = fprev_;
        }
        if (q.nr_fail[2]) { // counting is on (meepcu_set_profiling): solves, iterations, longest solve
          unsigned long long *cnt = reinterpret_cast<unsigned long long *>(q.nr_fail + 4);
          atomicAdd(cnt, 1ull);
          atomicAdd(cnt + 1, (unsigned long long)iters);
          atomicMax(q.nr_fail + 3, iters);
        }
        val = own == 0 ? out[0] : (own == 1 ? out[1] : out[2]);
      }
      done_nr = true;
    }
    else
      val = u ? (g[i] * u[i]) : g[i];
  }
  else {
    if (u1 && u2) {
      const T gs = g[i], us = u[i];
      if (q.chi3) {
        const T g1s = g1[i] + g1[i + s] + g1[i - s1] + g1[i + (s - s1)];
        const T g2s = g2[i] + g2[i + s] + g2[i - s2] + g2[i + (s - s2)];
        val = (gs * us + MCU_OFFDIAG(u1, g1, s1) + MCU_OFFDIAG(u2, g2, s2)) *
              calc_nonlinear_u<T>(gs * gs + 0.0625 * (g1s * g1s + g2s * g2s), gs, us, q.chi2[i],
                                  q.chi3[i]);
      }
      else
        val = (gs * us + MCU_OFFDIAG(u1, g1, s1) + MCU_OFFDIAG(u2, g2, s2));
    }
    else if (u1) {
      const T gs = g[i], us = u[i];
      if (q.chi3) {
        const T g1s = g1[i] + g1[i + s] + g1[i - s1] + g1[i + (s - s1)];
        val = (gs * us + MCU_OFFDIAG(u1, g1, s1)) *
              calc_nonlinear_u<T>(gs * gs + 0.0625 * (g1s * g1s), gs, us, q.chi2[i], q.chi3[i]);
      }
      else
        val = (gs * us + MCU_OFFDIAG(u1, g1, s1));
    }
    else if (q.chi3) {
      const T gs = g[i], us = u[i];
      if (q.mask && q.mask[i] == 0) // chi2 == chi3 == 0 here: calc_nonlinear_u is (1 + 0 + 0) / (1 + 0 + 0) = 1 exactly
        val = gs * us;
      else if (g1 && g2) {
        const T g1s = g1[i] + g1[i + s] + g1[i - s1] + g1[i + (s - s1)];
        const T g2s = g2[i] + g2[i + s] + g2[i - s2] + g2[i + (s - s2)];
        val = (gs * us) * calc_nonlinear_u<T>(gs * gs + 0.0625 * (g1s * g1s + g2s * g2s), gs, us,
                                              q.chi2[i], q.chi3[i]);
      }
      else if (g1) {
        const T g1s = g1[i] + g1[i + s] + g1[i - s1] + g1[i + (s - s1)];
        val = (gs * us) *
              calc_nonlinear_u<T>(gs * gs + 0.0625 * (g1s * g1s), gs, us, q.chi2[i], q.chi3[i]);
      }
      else
        val = (gs * us) * calc_nonlinear_u<T>(gs * gs, gs, us, q.chi2[i], q.chi3[i]);
    }
    else
      val = u ? (g[i] * u[i]) : g[i];
  }
  (void)done_nr;
  T enew, w;
  if (q.fw) {
      const int kw = 2 * pick3(q.kwaxis, ia, ib, ic) + q.kwoff;
    const T fwprev = q.fw[i], kapwkw = q.kapw[kw], sigwkw = q.sigw[kw];
    q.fw[i] = val;
    enew = q.f[i] + ((kapwkw + sigwkw) * val - (kapwkw - sigwkw) * fwprev);
    w = val;
  }
  else {
    enew = val;
    w = val;
  }
#pragma unroll
  for (int ip = 0; ip < kMaxPol; ++ip)
    if (ip < q.npol && q.pol[ip].p && q.pol[ip].s) {
      const PolParams<T> &pp = q.pol[ip];
      const T pcur = pp.p[i];
      pp.p[i] = pp.gamma1inv * (pcur * (2 - pp.omega0dtsqr_denom) - pp.gamma1 * pp.pp[i] +
                                pp.omega0dtsqr * (pp.s[i] * w));
      pp.pp[i] = pcur;
    }
  if (on_zero_plane<T>(q.zmask_eh, q.lo, q.hi, ia, ib, ic)) enew = 0;
  q.f[i] = enew;
}

// EhParams::mask for one component (see there).  mode 0 (the fork's NR branch): 1 where chi2 != 0 and at most one of
// u, u1, u2 is zero (step_generic.cpp:757-764); mode 1 (UPSTREAM_PADE without off-diagonal rows): 1 where chi2 or chi3 != 0.
template <class T>
__global__ void __launch_bounds__(kThreads) k_eh_mask(unsigned char *__restrict__ mask, const T *__restrict__ u, const T *__restrict__ u1,
                                                      const T *__restrict__ u2, const T *__restrict__ chi2, const T *__restrict__ chi3,
                                                      int mode, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    unsigned char m;
    if (mode == 0) {
      int zeroEps = 0;
      if (u[i] == 0) zeroEps += 1;
      if (u1[i] == 0) zeroEps += 1;
      if (u2[i] == 0) zeroEps += 1;
      m = !(chi2[i] == 0 || zeroEps > 1);
    }
    else
      m = chi2[i] != 0 || chi3[i] != 0;
    mask[i] = m;
  }
}

#ifndef MCU_EH_CTAS
#define MCU_EH_CTAS 4 // Newton-Raphson is latency-bound.  With the L2 prefetch of the operands (r02g, 384^3, ms per launch): 5.13 / 4.34 / 4.81 at 3 / 4 / 5 CTAs per SM (before the prefetch 5 CTAs were best: 4.56)
#endif
template <class T, int CTAS = MCU_EH_CTAS>
__global__ void __launch_bounds__(kThreads, CTAS)
    k_eh_general(const EhParams<T> *__restrict__ params, const Tile *__restrict__ tiles) {
  __shared__ EhParams<T> P;
  const Tile t = tiles[blockIdx.x];
  {
    const int nw = (int)(sizeof(EhParams<T>) / 4);
    const uint32_t *src = reinterpret_cast<const uint32_t *>(params + t.param);
    uint32_t *dst = reinterpret_cast<uint32_t *>(&P);
    for (int k = threadIdx.x; k < nw; k += kThreads) dst[k] = src[k];
  }
  __syncthreads();
  int ib, ic;
  if (t.p0 & kTileRows) {
    // 8 rows x 32 columns: the warps of a CTA share their range of the fastest axis, so a CTA is
    // either all inside a nonlinear slab (every warp iterates Newton-Raphson) or all outside (done
    // at once) -- in the flat mapping 3 warps of 8 carried the whole latency-bound solve
    ib = ((t.p0 & ~kTileRows) >> 12) * 8 + (int)(threadIdx.x >> 5);
    ic = (t.p0 & 4095) * 32 + (int)(threadIdx.x & 31);
  }
  else {
    const int p = t.p0 + threadIdx.x;
    ib = p / P.pitch;
    ic = p - ib * P.pitch;
  }
  if (ib < P.lo[1] || ib > P.hi[1] || ic < P.lo[2] || ic > P.hi[2]) return;
  const int a0 = t.a0 > P.lo[0] ? t.a0 : P.lo[0];
  const int a1 = t.a1 - 1 < P.hi[0] ? t.a1 - 1 : P.hi[0];
  long long i = (long long)a0 * P.sA + (long long)ib * P.sB + ic;
  for (int ia = a0; ia <= a1; ++ia, i += P.sA) {
    if (pf_planes<T>() > 0 && ia + pf_planes<T>() <= a1) {
      // the operands of this thread's point two planes ahead, into L2 (r02d ncu: 45 % of the kernel's stall samples
      // were the first dependent instructions after the loads of a point -- DRAM latency, once per point)
      const long long ipf = i + pf_planes<T>() * P.sA;
      prefetch_l2(P.g + ipf);
      if (P.u) prefetch_l2(P.u + ipf);
      if (P.mask) prefetch_l2(P.mask + ipf);
      if (P.g1) prefetch_l2(P.g1 + ipf);
      if (P.g2) prefetch_l2(P.g2 + ipf);
      if (P.chi2) prefetch_l2(P.chi2 + ipf);
      if (P.fw || (P.u1 && P.u2 && P.chi3)) prefetch_l2(P.f + ipf);
    }
    eh_point<T>(P, i, ia, ib, ic);
  }
}

// ------------------------------------------------------------------------------------------------
// lorentzian_susceptibility::update_P with an anisotropic sigma tensor (susceptibility.cpp:188-262):
// the 3x3 / 2x2 branches average the OTHER components of W over four neighbours, so they run as
// their own pass after the E (H) update of the whole chunk and the exchange of W halos
// (step.cpp:109-118).  Isotropic components of the same susceptibility are done here too.
// ------------------------------------------------------------------------------------------------
template <class T> struct PolGenParams {
  int nA, nB, nC, pitch;
  long long sA, sB;
  int lo[3], hi[3];
  T *p, *pp;
  const T *s, *w;        // sigma[c][d], W[c]
  const T *s1, *w1, *s2, *w2; // off-diagonal sigma rows and the W components they multiply (after the swap rule)
  long long is, is1, is2;
  T gamma1inv, gamma1, omega0dtsqr, omega0dtsqr_denom;
};

template <class T>
__global__ void __launch_bounds__(kThreads)
    k_pol_general(const PolGenParams<T> *__restrict__ params, const Tile *__restrict__ tiles) {
  __shared__ PolGenParams<T> P;
  const Tile t = tiles[blockIdx.x];
  block_copy_params(&P, params + t.param);
  __syncthreads();
  const int pos = t.p0 + threadIdx.x;
  const int ib = pos / P.pitch, ic = pos - ib * P.pitch;
  if (ib < P.lo[1] || ib > P.hi[1] || ic < P.lo[2] || ic > P.hi[2]) return;
  const int a0 = t.a0 > P.lo[0] ? t.a0 : P.lo[0];
  const int a1 = t.a1 - 1 < P.hi[0] ? t.a1 - 1 : P.hi[0];
  T *p = P.p, *pp = P.pp;
  const T *s = P.s, *w = P.w, *s1 = P.s1, *w1 = P.w1, *s2 = P.s2, *w2 = P.w2;
  const long long is = P.is, is1 = P.is1, is2 = P.is2;
  const T gamma1inv = P.gamma1inv, gamma1 = P.gamma1, omega0dtsqr = P.omega0dtsqr, omega0dtsqr_denom = P.omega0dtsqr_denom;
  long long i = (long long)a0 * P.sA + (long long)ib * P.sB + ic;
  for (int ia = a0; ia <= a1; ++ia, i += P.sA) {
    if (s1 && s2) { // 3x3 anisotropic, :225-241
      if (s[i] != 0) {
        const T pcur = p[i];
        p[i] = gamma1inv * (pcur * (2 - omega0dtsqr_denom) - gamma1 * pp[i] +
                            omega0dtsqr * (s[i] * w[i] + MCU_POFFDIAG(s1, w1, is1, is) + MCU_POFFDIAG(s2, w2, is2, is)));
        pp[i] = pcur;
      }
    }
    else if (s1) { // 2x2 anisotropic, :243-251
      if (s[i] != 0) {
        const T pcur = p[i];
        p[i] = gamma1inv * (pcur * (2 - omega0dtsqr_denom) - gamma1 * pp[i] +
                            omega0dtsqr * (s[i] * w[i] + MCU_POFFDIAG(s1, w1, is1, is)));
        pp[i] = pcur;
      }
    }
    else { // isotropic, :253-259
      const T pcur = p[i];
      p[i] = gamma1inv * (pcur * (2 - omega0dtsqr_denom) - gamma1 * pp[i] + omega0dtsqr * (s[i] * w[i]));
      pp[i] = pcur;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// gyrotropic_susceptibility::update_P (susceptibility.cpp:445-584): three polarisation components
// on each driving-field position, coupled by the skew-symmetric gyration tensor; the two transverse
// ones are driven by four-point averages of the other W components, so (like k_pol_general) this is
// its own pass after the chunk's E (H) update and the W halo exchange.
// ------------------------------------------------------------------------------------------------
template <class T> struct PolGyroParams {
  int nA, nB, nC, pitch;
  long long sA, sB;
  int lo[3], hi[3];
  T *p0, *pp0, *p1, *pp1, *p2, *pp2; // P / P_prev along d0 = direction of the component, d1, d2
  const T *s, *w0, *w1, *w2;
  long long is, is1, is2;
  int saturated;                      // GYROTROPIC_SATURATED (Landau-Lifshitz-Gilbert) or Lorentzian / Drude
  T g01, g02, g10, g12, g20, g21;     // gyro_tensor[d_a][d_b]
  T inv[3][3];                        // inv[d_a][d_b], a, b in (d0, d1, d2) order
  T diag, gamma1, omega0dtsqr, pt;    // Lorentzian / Drude constants (:456-461)
  T omega2pidt, g2pidt, alpha, dt2pi; // saturated-model constants (:531, :449-450)
};

#define MCU_OFFDIAGW(g, sx, s) (0.25 * (g[i] + g[i - sx] + g[i + s] + g[i + s - sx]))

template <class T>
__global__ void __launch_bounds__(kThreads)
    k_pol_gyro(const PolGyroParams<T> *__restrict__ params, const Tile *__restrict__ tiles) {
  __shared__ PolGyroParams<T> P;
  const Tile t = tiles[blockIdx.x];
  block_copy_params(&P, params + t.param);
  __syncthreads();
  const int pos = t.p0 + threadIdx.x;
  const int ib = pos / P.pitch, ic = pos - ib * P.pitch;
  if (ib < P.lo[1] || ib > P.hi[1] || ic < P.lo[2] || ic > P.hi[2]) return;
  const int a0 = t.a0 > P.lo[0] ? t.a0 : P.lo[0];
  const int a1 = t.a1 - 1 < P.hi[0] ? t.a1 - 1 : P.hi[0];
  T *p0 = P.p0, *pp0 = P.pp0, *p1 = P.p1, *pp1 = P.pp1, *p2 = P.p2, *pp2 = P.pp2;
  const T *s = P.s, *w0 = P.w0, *w1 = P.w1, *w2 = P.w2;
  const long long is = P.is, is1 = P.is1, is2 = P.is2;
  const T g01 = P.g01, g02 = P.g02, g10 = P.g10, g12 = P.g12, g20 = P.g20, g21 = P.g21;
  const T diag = P.diag, gamma1 = P.gamma1, omega0dtsqr = P.omega0dtsqr, pt = P.pt;
  const T omega2pidt = P.omega2pidt, g2pidt = P.g2pidt, alpha = P.alpha, dt2pi = P.dt2pi;
  long long i = (long long)a0 * P.sA + (long long)ib * P.sB + ic;
  for (int ia = a0; ia <= a1; ++ia, i += P.sA) {
    T r0, r1, r2;
    if (!P.saturated) { // :484-491
      r0 = diag * p0[i] - gamma1 * pp0[i] + omega0dtsqr * s[i] * w0[i] - pt * g01 * pp1[i] - pt * g02 * pp2[i];
      r1 = diag * p1[i] - gamma1 * pp1[i] + (w1 ? omega0dtsqr * s[i] * MCU_OFFDIAGW(w1, is1, is) : 0) -
           pt * g10 * pp0[i] - pt * g12 * pp2[i];
      r2 = diag * p2[i] - gamma1 * pp2[i] + (w2 ? omega0dtsqr * s[i] * MCU_OFFDIAGW(w2, is2, is) : 0) -
           pt * g21 * pp1[i] - pt * g20 * pp0[i];
    }
    else { // :560-571
      T q0, q1, q2;
      q0 = -omega2pidt * p0[i] + 0.5 * alpha * pp0[i] + dt2pi * s[i] * w0[i];
      q1 = -omega2pidt * p1[i] + 0.5 * alpha * pp1[i] + dt2pi * s[i] * (w1 ? MCU_OFFDIAGW(w1, is1, is) : 0);
      q2 = -omega2pidt * p2[i] + 0.5 * alpha * pp2[i] + dt2pi * s[i] * (w2 ? MCU_OFFDIAGW(w2, is2, is) : 0);
      r0 = 0.5 * pp0[i] - g2pidt * p0[i] + g01 * q1 + g02 * q2;
      r1 = 0.5 * pp1[i] - g2pidt * p1[i] + g12 * q2 + g10 * q0;
      r2 = 0.5 * pp2[i] - g2pidt * p2[i] + g20 * q0 + g21 * q1;
    }
    pp0[i] = p0[i];
    pp1[i] = p1[i];
    pp2[i] = p2[i];
    p0[i] = P.inv[0][0] * r0 + P.inv[0][1] * r1 + P.inv[0][2] * r2;
    p1[i] = P.inv[1][0] * r0 + P.inv[1][1] * r1 + P.inv[1][2] * r2;
    p2[i] = P.inv[2][0] * r0 + P.inv[2][1] * r1 + P.inv[2][2] * r2;
  }
}

// ------------------------------------------------------------------------------------------------
// multilevel_susceptibility::update_P (multilevel-atom.cpp:240-374), two passes after the chunk's E update and the W
// halo exchange:
//   k_ml_levels  the populations N of the L levels at every owned point of the CENTRED grid (:260-329): relaxation
//                (I - Gamma dt/2) N, plus alpha * (E . dP/dt + gamma_perp E . P) from the eight-point sums of W, W_prev
//                and of every transition's P, P_prev around the point, then N = GammaInv * that.  N keeps the
//                reference's layout N[i_host * L + level] (the L values of a point are contiguous: one thread, one point).
//   k_ml_pol     every transition's polarisation at the owned Yee points of its component (:331-373): an oscillator
//                driven by sigma_t * sigma * W times the population inversion averaged back from four centred points.
// Every expression keeps the reference's operand order; the point-independent factors ((l1 == l2) - Gamma dt/2,
// gamma pi dt, 2 - omega0^2 dt^2 ...) are formed once on the host in realnum, as the reference forms them per point.
// ------------------------------------------------------------------------------------------------
constexpr int kMlMaxL = 8, kMlMaxT = 4;
template <class T> struct MlLevelParams {
  int nA, nB, nC, pitch;
  long long sA, sB;       // device strides of the field / polarisation arrays
  long long hA, hB, hC;   // host strides by device axis (the layout of N)
  int hi[3];              // owned centred points: 0 .. hi[axis]
  int L, Tn, ndot;
  T *N;
  const T *w[3][2], *wp[3][2];  // W = f_w ?: f and f_w_prev of the (up to three) polarised components, cmp 0 / 1
  long long o1[3], o2[3];       // grid_volume::yee2cent_offsets of those components, device elements
  const T *p[3][2][kMlMaxT], *pp[3][2][kMlMaxT];
  T A[kMlMaxL * kMlMaxL];       // (l1 == l2) - Gamma[l1 L + l2] dt/2   (:265-270)
  T Ginv[kMlMaxL * kMlMaxL];    // multilevel_data::GammaInv
  T alpha[kMlMaxL * kMlMaxT];   // alpha[l Tn + t]
  T gperpdt[kMlMaxT];           // gamma[t] pi dt                        (:297)
};

template <class T>
__global__ void __launch_bounds__(kThreads)
    k_ml_levels(const MlLevelParams<T> *__restrict__ params, const Tile *__restrict__ tiles) {
  __shared__ MlLevelParams<T> P;
  const Tile t = tiles[blockIdx.x];
  block_copy_params(&P, params + t.param);
  __syncthreads();
  const int pos = t.p0 + threadIdx.x;
  const int ib = pos / P.pitch, ic = pos - ib * P.pitch;
  if (ib > P.hi[1] || ic > P.hi[2]) return;
  const int a1 = t.a1 - 1 < P.hi[0] ? t.a1 - 1 : P.hi[0];
  const int L = P.L, Tn = P.Tn, ndot = P.ndot;
  for (int ia = t.a0; ia <= a1; ++ia) {
    const long long i = (long long)ia * P.sA + (long long)ib * P.sB + ic;
    T *N = P.N + ((long long)ia * P.hA + (long long)ib * P.hB + (long long)ic * P.hC) * L;
    T Nv[kMlMaxL], Ntmp[kMlMaxL];
    for (int l = 0; l < L; ++l) Nv[l] = N[l];
    for (int l1 = 0; l1 < L; ++l1) { // Ntmp = (I - Gamma * dt/2) * N
      T a = 0;
      for (int l2 = 0; l2 < L; ++l2) a += P.A[l1 * L + l2] * Nv[l2];
      Ntmp[l1] = a;
    }
    T E8[3][2] = {{0, 0}, {0, 0}, {0, 0}}; // E*8 at point i (:273-288)
    for (int k = 0; k < ndot; ++k) {
      const long long o1 = P.o1[k], o2 = P.o2[k];
      const T *w = P.w[k][0], *wp = P.wp[k][0];
      E8[k][0] = w[i] + w[i + o1] + w[i + o2] + w[i + o1 + o2] + wp[i] + wp[i + o1] + wp[i + o2] + wp[i + o1 + o2];
      if (P.w[k][1]) {
        w = P.w[k][1];
        wp = P.wp[k][1];
        E8[k][1] = w[i] + w[i + o1] + w[i + o2] + w[i + o1 + o2] + wp[i] + wp[i + o1] + wp[i + o2] + wp[i + o1 + o2];
      }
    }
    for (int tr = 0; tr < Tn; ++tr) { // Ntmp = Ntmp + alpha * E * dP (:291-321)
      T EdP32 = 0, EPave64 = 0;
      const T gperpdt = P.gperpdt[tr];
      for (int k = 0; k < ndot; ++k) {
        const long long o1 = P.o1[k], o2 = P.o2[k];
        for (int cmp = 0; cmp < 2; ++cmp) {
          const T *p = P.p[k][cmp][tr], *pp = P.pp[k][cmp][tr];
          if (!p) continue;
          const T sp = p[i] + p[i + o1] + p[i + o2] + p[i + o1 + o2];
          const T spp = pp[i] + pp[i + o1] + pp[i + o2] + pp[i + o1 + o2];
          const T dP = sp - spp, Pave2 = sp + spp;
          EdP32 += dP * E8[k][cmp];
          EPave64 += Pave2 * E8[k][cmp];
        }
      }
      EdP32 *= 0.03125;
      EPave64 *= 0.015625;
      for (int l = 0; l < L; ++l) Ntmp[l] += P.alpha[l * Tn + tr] * EdP32 + P.alpha[l * Tn + tr] * gperpdt * EPave64;
    }
    for (int l1 = 0; l1 < L; ++l1) { // N = GammaInv * Ntmp
      T a = 0;
      for (int l2 = 0; l2 < L; ++l2) a += P.Ginv[l1 * L + l2] * Ntmp[l2];
      N[l1] = a;
    }
  }
}

template <class T> struct MlPolParams {
  int nA, nB, nC, pitch;
  long long sA, sB;
  long long hA, hB, hC;
  int lo[3], hi[3];     // owned points of the component
  int L, lp, lm;        // the two levels the transition couples (alpha > 0 / alpha < 0, :341-347)
  T *p, *pp;
  const T *s, *w, *N;
  long long oa, ob;     // grid_volume::cent2yee_offsets in elements of N (host strides times L, negative)
  T gamma1inv, gamma1, two_minus, dtsqr, st; // :332-336,351
};

template <class T>
__global__ void __launch_bounds__(kThreads)
    k_ml_pol(const MlPolParams<T> *__restrict__ params, const Tile *__restrict__ tiles) {
  __shared__ MlPolParams<T> P;
  const Tile t = tiles[blockIdx.x];
  block_copy_params(&P, params + t.param);
  __syncthreads();
  const int pos = t.p0 + threadIdx.x;
  const int ib = pos / P.pitch, ic = pos - ib * P.pitch;
  if (ib < P.lo[1] || ib > P.hi[1] || ic < P.lo[2] || ic > P.hi[2]) return;
  const int a0 = t.a0 > P.lo[0] ? t.a0 : P.lo[0];
  const int a1 = t.a1 - 1 < P.hi[0] ? t.a1 - 1 : P.hi[0];
  T *p = P.p, *pp = P.pp;
  const T *s = P.s, *w = P.w;
  const int lp = P.lp, lm = P.lm;
  const long long oa = P.oa, ob = P.ob;
  const T gamma1inv = P.gamma1inv, gamma1 = P.gamma1, two_minus = P.two_minus, dtsqr = P.dtsqr, st = P.st;
  for (int ia = a0; ia <= a1; ++ia) {
    const long long i = (long long)ia * P.sA + (long long)ib * P.sB + ic;
    const T *Ni = P.N + ((long long)ia * P.hA + (long long)ib * P.hB + (long long)ic * P.hC) * P.L;
    const T pcur = p[i];
    // dNi is the population inversion for this transition (:366-368)
    const T dNi = 0.25 * (Ni[lp] + Ni[lp + oa] + Ni[lp + ob] + Ni[lp + oa + ob] - Ni[lm] - Ni[lm + oa] - Ni[lm + ob] -
                          Ni[lm + oa + ob]);
    p[i] = gamma1inv * (pcur * two_minus - gamma1 * pp[i] - dtsqr * (st * s[i] * w[i]) * dNi);
    pp[i] = pcur;
  }
}

// ------------------------------------------------------------------------------------------------
// noisy_lorentzian_susceptibility::update_P (susceptibility.cpp:317-339), the part after the Lorentzian update:
// p[i] += gaussian_random(0, amp sqrt(sigma[i])) at every owned point, in the reference's loop order.  The unit
// Gaussians g come from the CALLER's generator (include/meepcu.h, meepcu_set_noise): gaussian_random(0, stddev) is
// 0 + g stddev in double (random.cpp:57-69), so the value stored is the reference's, bit for bit.
// ------------------------------------------------------------------------------------------------
// whole-array device copies from a table (fields_chunk::f_w_prev = W before the E update, update_eh.cpp:221-226)
template <class T> struct CopyItem {
  T *dst;
  const T *src;
  long long n;
};
template <class T> __global__ void __launch_bounds__(kThreads) k_copy_arrays(const CopyItem<T> *__restrict__ items) {
  const CopyItem<T> it = items[blockIdx.y];
  for (long long k = (long long)blockIdx.x * kThreads + threadIdx.x; k < it.n; k += (long long)gridDim.x * kThreads)
    it.dst[k] = it.src[k];
}

template <class T> struct NoiseSeg {
  T *p;
  const T *s;
  T amp;            // w2pi noise_amp sqrt(g2pi) dt dt / (1 + g2pi dt / 2), in realnum (:325)
  long long first;  // index of the segment's first value within the half step's sequence
  long long base;   // device index of its first owned point
  long long ds[3];  // device strides X, Y, Z
  int n[3];         // owned points per direction (1: absent), X outermost = LOOP_OVER_VOL_OWNED's order
};

template <class T>
__global__ void __launch_bounds__(kThreads)
    k_pol_noise(const NoiseSeg<T> *__restrict__ segs, int nseg, long long total, const double *__restrict__ g) {
  for (long long k = (long long)blockIdx.x * kThreads + threadIdx.x; k < total; k += (long long)gridDim.x * kThreads) {
    int lo = 0, hi = nseg - 1; // last segment with first <= k
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (segs[mid].first <= k) lo = mid;
      else hi = mid - 1;
    }
    const NoiseSeg<T> &S = segs[lo];
    long long r = k - S.first;
    const int iz = (int)(r % S.n[2]);
    r /= S.n[2];
    const int iy = (int)(r % S.n[1]);
    const int ix = (int)(r / S.n[1]);
    const long long i = S.base + ix * S.ds[0] + iy * S.ds[1] + iz * S.ds[2];
    const T stddev = S.amp * sqrt(S.s[i]);
    S.p[i] = (T)((double)S.p[i] + (0.0 + g[k] * (double)stddev));
  }
}

// ------------------------------------------------------------------------------------------------
// dft_chunk::update_dft (dft.cpp:265-307).  One thread per (point, omega) element of dft[] so the
// read-modify-write of the accumulator -- the dominant traffic -- is fully coalesced.
// ------------------------------------------------------------------------------------------------
template <class T> struct DftParams {
  const T *f0, *f1;  // field component re / im (f1 NULL for real fields)
  T *dft;            // complex<T>[N*Nomega] as (re,im) pairs
  long long idx0;    // device index of the region's first point
  long long st[3];   // device strides of the three loop levels (reference loop order)
  int n[3];          // loop_n1..3
  int dir[3];        // loop_d1..3 (direction index X/Y/Z -> 0..2)
  long long avg1, avg2; // device offsets
  double s0[3], s1[3], e0[3], e1[3], dV0, dV1;
  int include_dV, sqrt_dV;
  int Nomega;
  int phase_group;   // which row of the per-step phase table
  long long N;
};
struct DftWork {
  int param;
  int npts;      // points of this CTA (<= kDftMaxPts)
  long long pt0; // first point (IVEC_LOOP_COUNTER) of this CTA
};

template <class T> struct CplxOf;
template <> struct CplxOf<double> { typedef double2 type; };
template <> struct CplxOf<float> { typedef float2 type; };
constexpr int kDftMaxPts = 1024;     // points per CTA at most (their weighted field values live in shared memory)
constexpr int kDftElemsPerCta = 4096; // accumulator elements per CTA aimed at: npts = clamp(kDftElemsPerCta / Nomega, 1, kDftMaxPts)
constexpr int kDftUnroll = 4;         // accumulator loads in flight per thread

// A CTA owns `npts` consecutive points = npts * Nomega CONSECUTIVE complex accumulators.  Phase 1: one thread per point
// evaluates the weighted (and averaged) field value once -- index decomposition, IVEC_LOOP_WEIGHT, up to four field
// loads -- and parks it in shared memory.  Phase 2: the threads stream the accumulator block with 16-byte (fp64) /
// 8-byte (fp32) accesses, kDftUnroll loads in flight, (point, omega) of an element advanced incrementally -- per
// element two multiplies, two adds, one load and one store.  (Until r02 every (point, omega) element recomputed the
// point's weight and field value: Nomega times the arithmetic, issue-bound at half the HBM rate.)  Per element the
// operations are those of dft.cpp:295-305, so the accumulators are bit-identical.
template <class T>
__global__ void __launch_bounds__(kThreads)
    k_dft(const DftParams<T> *__restrict__ params, const DftWork *__restrict__ work,
          const T *__restrict__ phases, const int *__restrict__ phase_off,
          const unsigned char *__restrict__ active) {
  typedef typename CplxOf<T>::type CT;
  __shared__ T sf[2][kDftMaxPts];
  const DftWork wk = work[blockIdx.x];
  if (!active[wk.param]) return; // decimation: dft.cpp:259
  const DftParams<T> &q = params[wk.param];
  const int Nw = q.Nomega, n2 = q.n[1], n3 = q.n[2];
  const int numcmp = q.f1 ? 2 : 1;
  for (int t = threadIdx.x; t < wk.npts; t += kThreads) {
    const long long pt = wk.pt0 + t;
    int i1, i2, i3;
    if (q.N < (1LL << 31)) {
      const unsigned p = (unsigned)pt, r = p / (unsigned)n3;
      i3 = (int)(p - r * (unsigned)n3);
      i1 = (int)(r / (unsigned)n2);
      i2 = (int)(r - (unsigned)i1 * (unsigned)n2);
    }
    else {
      i3 = (int)(pt % n3);
      const long long r = pt / n3;
      i2 = (int)(r % n2);
      i1 = (int)(r / n2);
    }
    const long long idx = q.idx0 + i1 * q.st[0] + i2 * q.st[1] + i3 * q.st[2];
    double w;
    if (q.include_dV) { // IVEC_LOOP_WEIGHT, vec.hpp:372-383
      const int li[3] = {i1, i2, i3};
      double w1[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const int i = li[d], n = q.n[d], dir = q.dir[d];
        w1[d] = (i > 1 && i < n - 2)
                    ? 1.0
                    : (i == 0 ? q.s0[dir]
                              : (i == 1 ? q.s1[dir]
                                        : i == n - 1 ? q.e0[dir] : (i == n - 2 ? q.e1[dir] : 1.0)));
      }
      w = (w1[2] * (w1[1] * ((q.dV0 + q.dV1 * i2) * w1[0])));
      if (q.sqrt_dV) w = sqrt(w);
    }
    else
      w = 1.0;
    T f[2] = {0, 0};
    for (int cmp = 0; cmp < numcmp; ++cmp) {
      const T *ff = cmp ? q.f1 : q.f0;
      if (q.avg2)
        f[cmp] = (w * 0.25) * (ff[idx] + ff[idx + q.avg1] + ff[idx + q.avg2] + ff[idx + (q.avg1 + q.avg2)]);
      else if (q.avg1)
        f[cmp] = (w * 0.5) * (ff[idx] + ff[idx + q.avg1]);
      else
        f[cmp] = w * ff[idx];
    }
    sf[0][t] = f[0];
    sf[1][t] = f[1];
  }
  __syncthreads();
  const T *__restrict__ phb = phases + 2 * phase_off[q.phase_group];
  CT *__restrict__ acc0 = reinterpret_cast<CT *>(q.dft) + wk.pt0 * Nw;
  const int n = wk.npts * Nw;
  const int dp = kThreads / Nw, dw = kThreads - dp * Nw; // (point, omega) advance of one stride of kThreads elements
  int e = threadIdx.x;
  int pt = e / Nw, iw = e - pt * Nw;
  while (e < n) {
    CT a[kDftUnroll];
    int pts[kDftUnroll], iws[kDftUnroll];
#pragma unroll
    for (int k = 0; k < kDftUnroll; ++k) {
      const int ek = e + k * kThreads;
      pts[k] = pt;
      iws[k] = iw;
      if (ek < n) a[k] = acc0[ek]; // the dominant traffic: every load of the batch before the first store
      pt += dp;
      iw += dw;
      if (iw >= Nw) { iw -= Nw; ++pt; }
    }
#pragma unroll
    for (int k = 0; k < kDftUnroll; ++k) {
      const int ek = e + k * kThreads;
      if (ek >= n) break;
      const T pr = phb[2 * iws[k]], pi = phb[2 * iws[k] + 1];
      const T f0 = sf[0][pts[k]];
      CT d = a[k];
      if (numcmp == 2) { // complex<realnum> product, dft.cpp:295-299
        const T f1 = sf[1][pts[k]];
        d.x = d.x + (pr * f0 - pi * f1);
        d.y = d.y + (pr * f1 + pi * f0);
      }
      else { // :300-305
        d.x = d.x + f0 * pr;
        d.y = d.y + f0 * pi;
      }
      acc0[ek] = d;
    }
    e += kDftUnroll * kThreads;
  }
}

// ------------------------------------------------------------------------------------------------
// Reductions for the host-side readers: the arrays stay on the device, a few
// kilobytes of per-CTA partial sums travel.  Every partial is formed in a fixed order (no atomics), and
// the host adds the partials in work-item order, so a result does not depend on scheduling.
// ------------------------------------------------------------------------------------------------
// dft_flux::flux / dft_energy::electric, magnetic (dft.cpp:533-548,657-687):
//   out[i] += Re(a[k*Nw+i] * conj(b[k*Nw+i])) = a.re*b.re + a.im*b.im   (term in realnum, sum in double)
// A CTA = (kThreads / W) point lanes x W frequency lanes (W = power of two >= min(Nw, kThreads)): consecutive
// threads read consecutive frequencies of one point -- the layout of dft[] -- and walk the points of the work item.
template <class T> struct DftDotWork {
  const T *a, *b;   // complex<T> arrays as (re,im) pairs
  long long k0, k1; // points [k0, k1) of this pair
};
template <class T>
__global__ void __launch_bounds__(kThreads)
    k_dft_dot(const DftDotWork<T> *__restrict__ work, int Nw, int W, double *__restrict__ partial) {
  typedef typename CplxOf<T>::type CT;
  __shared__ double red[kThreads];
  const DftDotWork<T> wk = work[blockIdx.x];
  const int wi = threadIdx.x % W, lk = threadIdx.x / W, lanes = kThreads / W;
  const int i = blockIdx.y * W + wi;
  double acc = 0;
  if (i < Nw) {
    const CT *a = reinterpret_cast<const CT *>(wk.a), *b = reinterpret_cast<const CT *>(wk.b);
    for (long long k = wk.k0 + lk; k < wk.k1; k += lanes) {
      const CT x = a[k * Nw + i], y = b[k * Nw + i];
      const T term = x.x * y.x + x.y * y.y;
      acc += (double)term;
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  if (lk == 0 && i < Nw) {
    double s = acc;
    for (int l = 1; l < lanes; ++l) s += red[l * W + wi];
    partial[(size_t)blockIdx.x * Nw + i] = s;
  }
}

// fields::integrate with two components and dot_integrand (integrate.cpp:47-131, energy_and_flux.cpp:61-65):
// per point  F_i = complex<realnum>(complex<double>(f_re, f_im) * ph_i),  f = 0.25 * (four-point sum in realnum),
//            v = Re(conj(F_0) F_1) = F0.re*F1.re + F0.im*F1.im in double,  sum += v * IVEC_LOOP_WEIGHT, max |v|.
template <class T> struct IntgParams {
  const T *f[2][2];
  long long off[2][2]; // device offsets of the yee2cent averages
  double ph[2][2];
  long long idx0, st[3];
  int n[3], dir[3];
  double s0[3], s1[3], e0[3], e1[3], dV0, dV1;
  long long N;
};
struct IntgWork {
  int param;
  long long p0; // first point (loop counter) of this CTA
};
constexpr int kIntgPerThread = 8;
template <class T>
__global__ void __launch_bounds__(kThreads)
    k_integrate_dot(const IntgParams<T> *__restrict__ params, const IntgWork *__restrict__ work, double *__restrict__ partial) {
  __shared__ double rs[kThreads], rm[kThreads];
  const IntgWork wk = work[blockIdx.x];
  const IntgParams<T> &q = params[wk.param];
  const int n2 = q.n[1], n3 = q.n[2];
  double sum = 0, mx = 0;
  for (int r = 0; r < kIntgPerThread; ++r) {
    const long long pt = wk.p0 + threadIdx.x + (long long)r * kThreads;
    if (pt >= q.N) break;
    const int i3 = (int)(pt % n3);
    const long long rr = pt / n3;
    const int i2 = (int)(rr % n2), i1 = (int)(rr / n2);
    const long long idx = q.idx0 + i1 * q.st[0] + i2 * q.st[1] + i3 * q.st[2];
    const int li[3] = {i1, i2, i3};
    double w1[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) { // IVEC_LOOP_WEIGHT1x, vec.hpp:372-383
      const int i = li[d], n = q.n[d], dir = q.dir[d];
      w1[d] = (i > 1 && i < n - 2) ? 1.0
                                   : (i == 0 ? q.s0[dir] : (i == 1 ? q.s1[dir] : i == n - 1 ? q.e0[dir] : (i == n - 2 ? q.e1[dir] : 1.0)));
    }
    const double w = (w1[2] * (w1[1] * ((q.dV0 + q.dV1 * i2) * w1[0])));
    T F[2][2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      double f[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const T *p = q.f[c][k];
        f[k] = p ? 0.25 * (p[idx] + p[idx + q.off[c][0]] + p[idx + q.off[c][1]] + p[idx + (q.off[c][0] + q.off[c][1])]) : 0.0;
      }
      const double2 z = cmul(f[0], f[1], q.ph[c][0], q.ph[c][1]);
      F[c][0] = (T)z.x;
      F[c][1] = (T)z.y;
    }
    const double v = (double)F[0][0] * (double)F[1][0] + (double)F[0][1] * (double)F[1][1];
    sum += v * w;
    mx = fmax(mx, fabs(v));
  }
  rs[threadIdx.x] = sum;
  rm[threadIdx.x] = mx;
  __syncthreads();
  for (int s = kThreads / 2; s > 0; s >>= 1) { // fixed-shape tree: the same partial whatever the scheduling
    if ((int)threadIdx.x < s) {
      rs[threadIdx.x] += rs[threadIdx.x + s];
      rm[threadIdx.x] = fmax(rm[threadIdx.x], rm[threadIdx.x + s]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[2 * (size_t)blockIdx.x] = rs[0];
    partial[2 * (size_t)blockIdx.x + 1] = rm[0];
  }
}

// ------------------------------------------------------------------------------------------------
// Cylindrical coordinates (step_db.cpp:86-122): the update of a z component needs (1/r) d(r F_phi)/dr; the reference
// builds the running sum over r of that expression (f_rderiv_int) so that the plain difference of step_curl yields
// it.  The sum is sequential in r by construction -- every partial sum is rounded -- so one thread owns one z
// column and walks r; consecutive threads hold consecutive z (coalesced rows).  Same expressions, same roundings.
// ------------------------------------------------------------------------------------------------
template <class T> struct CylRderiv {
  T *S;        // f_rderiv_int, device layout of the chunk
  const T *fp; // the phi component (E_p or H_p) of this cmp
  int nr, nz, pitch;
  T ir0;       // origin_r * a + 0.5 * iyee_shift(c_p).r
};
template <class T>
__global__ void __launch_bounds__(kThreads) k_cyl_rderiv(const CylRderiv<T> *__restrict__ tab, int n) {
  for (int e = blockIdx.y; e < n; e += gridDim.y) {
    const CylRderiv<T> q = tab[e];
    for (int iz = blockIdx.x * blockDim.x + threadIdx.x; iz <= q.nz; iz += gridDim.x * blockDim.x) {
      T acc = 0, prev = q.fp[iz];
      q.S[iz] = 0;
      for (int ir = 1; ir <= q.nr; ++ir) {
        const T rinv = 1.0 / ((ir + q.ir0) - 0.5);
        const long long idx = (long long)ir * q.pitch + iz;
        const T cur = q.fp[idx];
        acc = acc + rinv * (cur * (ir + q.ir0) - prev * ((ir - 1) + q.ir0));
        q.S[idx] = acc;
        prev = cur;
      }
    }
  }
}

// The r = 0 rules of step_db.cpp:292-457 for a chunk that contains the axis: the explicit update of the one
// component that lives on the axis (m = 0: D_z from H_phi; |m| = 1: D_phi from H_r, H_z / B_r from E_phi, E_z) with
// its PML / conductivity chain, and the rows that are cleared.  One thread per z.
template <class T> struct CylOrigin {
  int lo, hi, nzp1;         // z index range of the row update (inclusive); elements per row
  T *the_f, *fu;            // after the role swap of step_db.cpp:311-312 (the_f NULL: no row update)
  const T *f_p, *f_m;       // f_m: already points at the row it is read from
  int sd, m0;
  T f_m_mult, dt2;
  double courant;
  T *fcnd;
  const T *cnd, *cndinv;
  const T *sig, *kap, *siginv;
  int k_z, koff;            // k = k_z * 2 * iz + koff
  const T *sigu, *kapu, *siginvu;
  int ku_z, kuoff;
  T *zero[6];               // rows (first nzp1 elements) that are cleared
  int nzero;
};
template <class T>
__global__ void __launch_bounds__(kThreads) k_cyl_origin(const CylOrigin<T> *__restrict__ tab, int n) {
  for (int e = blockIdx.y; e < n; e += gridDim.y) {
    const CylOrigin<T> q = tab[e];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < q.nzp1; i += gridDim.x * blockDim.x) {
      if (q.the_f && i >= q.lo && i <= q.hi) {
        const T fprev = q.the_f[i];
        T dfcnd;
        if (q.m0) dfcnd = (T)(q.f_p[i] * (q.courant * 4));
        else dfcnd = (T)((q.sd * q.courant) * (q.f_p[i] - q.f_p[i - q.sd] - q.f_m_mult * q.f_m[i]));
        if (q.fcnd) {
          const T fcnd_prev = q.fcnd[i];
          const T fcn = ((1 - q.dt2 * q.cnd[i]) * fcnd_prev + dfcnd) * q.cndinv[i];
          q.fcnd[i] = fcn;
          dfcnd = fcn - fcnd_prev;
        }
        const int k = q.k_z * 2 * i + q.koff, ku = q.ku_z * 2 * i + q.kuoff;
        const T fn = ((q.kap ? q.kap[k] - q.sig[k] : (T)1) * fprev + dfcnd) * (q.siginv ? q.siginv[k] : (T)1);
        q.the_f[i] = fn;
        if (q.fu) q.fu[i] = q.siginvu[ku] * ((q.kapu ? q.kapu[ku] - q.sigu[ku] : (T)1) * q.fu[i] + fn - fprev);
      }
      for (int z = 0; z < q.nzero; ++z) q.zero[z][i] = 0;
    }
  }
}

// Row re-pitching between the dense layout of the caller's arrays (row = nC elements) and the padded
// device layout (row = pitch elements): host<->device transfers are then plain 1-D DMA at PCIe speed
// (2-D DMA with 2 KB rows runs at a fraction of it), and the re-pitching costs one HBM pass.
template <class T>
__global__ void __launch_bounds__(kThreads)
    k_rows(T *__restrict__ dst, long long dst_pitch, const T *__restrict__ src, long long src_pitch, int w,
           long long rows) {
  const long long n = rows * w, stride = (long long)gridDim.x * blockDim.x;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    const long long r = k / w;
    const int c = (int)(k - r * w);
    dst[r * dst_pitch + c] = src[r * src_pitch + c];
  }
}

// The same for MANY small arrays at once (meepcu_sync_*: the edge / corner PML chunks of a 3-D cell are hundreds of
// arrays of a few hundred KB): their dense rows travel packed in one staging block, one launch re-pitches them all.
template <class T> struct RowsEntry {
  T *dev;
  long long dpitch; // elements
  long long off;    // first element in the staging block
  int w;            // elements per dense row
  long long rows;
};
template <class T>
__global__ void __launch_bounds__(kThreads)
    k_rows_batch(const RowsEntry<T> *__restrict__ tab, int n, T *__restrict__ stage, int to_stage) {
  for (int e = blockIdx.y; e < n; e += gridDim.y) {
    const RowsEntry<T> q = tab[e];
    const long long tot = q.rows * q.w, stride = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < tot; k += stride) {
      const long long r = k / q.w;
      const int c = (int)(k - r * q.w);
      if (to_stage) stage[q.off + k] = q.dev[r * q.dpitch + c];
      else q.dev[r * q.dpitch + c] = stage[q.off + k];
    }
  }
}

// fields_chunk::average_with_backup (energy_and_flux.cpp:135-143): fc[i] = 0.5 * (fc[i] + backup[i])
template <class T> __global__ void __launch_bounds__(kThreads) k_average(T *__restrict__ f, const T *__restrict__ backup, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    f[i] = (T)(0.5 * (double)(f[i] + backup[i])); // the reference's promotions: realnum sum, double product
}

// the per-step NaN test of step.cpp:138-139: ONE launch over a table of all field arrays, testing `samples`
// evenly spaced elements of each (samples >= n: every element).  blockIdx.x walks the arrays, blockIdx.y
// splits the samples.  `flag` may live in page-locked host memory (the lazy per-call test).
struct FinEntry { const void *p; long long n; };
template <class T> __global__ void __launch_bounds__(kThreads) k_finite(const FinEntry *__restrict__ tab, int ntab, long long samples, int *flag) {
  int bad = 0;
  for (int a = blockIdx.x; a < ntab; a += gridDim.x) {
    const T *p = (const T *)tab[a].p;
    const long long n = tab[a].n;
    const long long stride = n > samples ? n / samples : 1;
    const long long cnt = (n + stride - 1) / stride;
    for (long long j = (long long)blockIdx.y * blockDim.x + threadIdx.x; j < cnt; j += (long long)gridDim.y * blockDim.x) {
      long long i = j * stride + stride / 2;
      if (i >= n) i = n - 1;
      if (!isfinite((double)p[i])) bad = 1;
    }
  }
  if (bad) *(volatile int *)flag = 1;
}

} // namespace mcu
