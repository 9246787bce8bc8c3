// meepcu.cu -- libmeepcuda.so: host side of the C ABI in include/meepcu.h.
//
// Owns the device mirrors of a `fields` object (reference src/meep.hpp:1461-1485,590-598),
// turns the reference's per-chunk state into flat parameter / tile tables once
// (meepcu_finalize), and then runs `fields::step()` (src/step.cpp:35-140) as a fixed sequence
// of batched launches per time step:
//
//     k_half<B/H>  ->  k_conn(H)  ->  k_half<D/E>  ->  k_conn(E)  ->  k_dft
//
// (+ k_conn(D/P), k_fmp, k_eh_general only for chunks whose E update reads neighbours of D-P:
// the fork's Newton-Raphson chi2 branch and the UPSTREAM_PADE mode.)  B/D/P halo copies that no
// kernel consumes are deferred to meepcu_sync_to_host.
//
// There is NO CPU fallback: without a CUDA device every entry point that needs one fails with
// MEEPCU_ECUDA.
#include "../../include/meepcu.h"
#include "kernels.cuh"

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <vector>
#include <dlfcn.h>
#include <omp.h>

using namespace mcu;

namespace {

thread_local std::string g_err;
int fail(int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(MEEPCU_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__,  \
                  __LINE__);                                                                       \
  } while (0)
#define RC(call)                                                                                   \
  do {                                                                                             \
    int rc_ = (call);                                                                              \
    if (rc_ != MEEPCU_OK) return rc_;                                                              \
  } while (0)

// ---- NCCL, bound at run time so that a single-GPU build has no dependency on it ----------------
typedef struct { char internal[128]; } nccl_uid;
typedef void *nccl_comm;
struct NcclApi {
  void *h;
  int (*GetUniqueId)(nccl_uid *);
  int (*CommInitRank)(nccl_comm *, int, nccl_uid, int);
  int (*CommDestroy)(nccl_comm);
  int (*Send)(const void *, size_t, int, int, nccl_comm, cudaStream_t);
  int (*Recv)(void *, size_t, int, int, nccl_comm, cudaStream_t);
  int (*GroupStart)();
  int (*GroupEnd)();
  int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm, cudaStream_t);
  const char *(*GetErrorString)(int);
};
NcclApi *nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    api.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.h) api.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (api.h) {
      *(void **)&api.GetUniqueId = dlsym(api.h, "ncclGetUniqueId");
      *(void **)&api.CommInitRank = dlsym(api.h, "ncclCommInitRank");
      *(void **)&api.CommDestroy = dlsym(api.h, "ncclCommDestroy");
      *(void **)&api.Send = dlsym(api.h, "ncclSend");
      *(void **)&api.Recv = dlsym(api.h, "ncclRecv");
      *(void **)&api.GroupStart = dlsym(api.h, "ncclGroupStart");
      *(void **)&api.GroupEnd = dlsym(api.h, "ncclGroupEnd");
      *(void **)&api.AllReduce = dlsym(api.h, "ncclAllReduce");
      *(void **)&api.GetErrorString = dlsym(api.h, "ncclGetErrorString");
      if (!api.GetUniqueId || !api.CommInitRank || !api.Send || !api.Recv || !api.GroupStart || !api.GroupEnd || !api.AllReduce)
        api.h = 0;
    }
  }
  return api.h ? &api : 0;
}
#define NC(call)                                                                                   \
  do {                                                                                             \
    int r_ = (call);                                                                               \
    if (r_ != 0)                                                                                   \
      return fail(MEEPCU_ENCCL, "%s failed: %s", #call,                                            \
                  nccl_api()->GetErrorString ? nccl_api()->GetErrorString(r_) : "?");              \
  } while (0)

// ---- component helpers (reference src/meep/vec.hpp:414-592, src/fields.cpp:411-425) -----------
inline bool is_electric(int c) { return c < 5; }
inline bool is_magnetic(int c) { return c >= 5 && c < 10; }
inline bool is_D(int c) { return c >= 10 && c < 15; }
inline bool is_B(int c) { return c >= 15 && c < 20; }
inline int comp_ft(int c) { return is_electric(c) ? 0 : is_magnetic(c) ? 1 : is_D(c) ? 2 : 3; }
inline int comp_dir(int c) {
  static const int d[5] = {MEEPCU_X, MEEPCU_Y, MEEPCU_R, MEEPCU_P, MEEPCU_Z};
  return d[c % 5];
}
inline int dir_comp(int ft, int d) { // direction_component(first_field_component(ft), d)
  static const int off[3] = {0, 1, 4};
  return 5 * ft + off[d];
}

// Cylindrical coordinates are handled in CANONICAL form: (R, P, Z) -> the slots of (X, Y, Z) -- which is the
// reference's own convention for everything it indexes by direction % 3 (src/meep/vec.hpp:1014-1020,
// src/fields.cpp:411-425) -- so that the library sees a Cartesian-like grid with extents in slots 0 and 2 and none
// in slot 1.  The descriptions are rewritten once, on entry; host pointers (the keys of everything) are untouched.
inline int canon_comp(int c) { // Er -> the Ex slot, Ep -> the Ey slot, Ez stays
  const int o = c % 5;
  return c - o + (o == 2 ? 0 : (o == 3 ? 1 : o));
}
inline void canon_chunk_desc(meepcu_chunk_desc &d) {
  const meepcu_chunk_desc o = d;
  static const int cyl_dir[3] = {MEEPCU_R, MEEPCU_P, MEEPCU_Z};
  for (int c = 0; c < 20; ++c) {
    for (int cmp = 0; cmp < 2; ++cmp) d.f[c][cmp] = d.f_u[c][cmp] = d.f_w[c][cmp] = d.f_cond[c][cmp] = d.f_bfast[c][cmp] = 0;
    for (int k = 0; k < 5; ++k) d.chi1inv[c][k] = d.conductivity[c][k] = d.condinv[c][k] = 0;
    d.chi2[c] = d.chi3[c] = 0;
  }
  for (int c = 0; c < 20; ++c) {
    if (c % 5 < 2) continue; // no Cartesian components in a cylindrical cell
    const int cc = canon_comp(c);
    for (int cmp = 0; cmp < 2; ++cmp) {
      d.f[cc][cmp] = o.f[c][cmp];
      d.f_u[cc][cmp] = o.f_u[c][cmp];
      d.f_w[cc][cmp] = o.f_w[c][cmp];
      d.f_cond[cc][cmp] = o.f_cond[c][cmp];
      d.f_bfast[cc][cmp] = o.f_bfast[c][cmp];
    }
    for (int k = 0; k < 3; ++k) {
      d.chi1inv[cc][k] = o.chi1inv[c][cyl_dir[k]];
      d.conductivity[cc][k] = o.conductivity[c][cyl_dir[k]];
      d.condinv[cc][k] = o.condinv[c][cyl_dir[k]];
    }
    d.chi2[cc] = o.chi2[c];
    d.chi3[cc] = o.chi3[c];
  }
  for (int k = 0; k < 6; ++k) { d.sig[k] = d.kap[k] = d.siginv[k] = 0; d.sigsize[k] = 0; }
  for (int k = 0; k < 3; ++k) {
    d.sig[k] = o.sig[cyl_dir[k]];
    d.kap[k] = o.kap[cyl_dir[k]];
    d.siginv[k] = o.siginv[cyl_dir[k]];
    d.sigsize[k] = o.sigsize[cyl_dir[k]];
  }
}
inline void canon_susc_desc(meepcu_susceptibility_desc &d) {
  const meepcu_susceptibility_desc o = d;
  static const int cyl_dir[3] = {MEEPCU_R, MEEPCU_P, MEEPCU_Z};
  for (int c = 0; c < 20; ++c) {
    for (int k = 0; k < 5; ++k) d.sigma[c][k] = 0;
    for (int cmp = 0; cmp < 2; ++cmp) d.P[c][cmp] = d.P_prev[c][cmp] = 0;
  }
  for (int c = 0; c < 20; ++c) {
    if (c % 5 < 2) continue;
    const int cc = canon_comp(c);
    for (int k = 0; k < 3; ++k) d.sigma[cc][k] = o.sigma[c][cyl_dir[k]];
    for (int cmp = 0; cmp < 2; ++cmp) { d.P[cc][cmp] = o.P[c][cmp]; d.P_prev[cc][cmp] = o.P_prev[c][cmp]; }
  }
}

enum ArrKind { A_GRID = 0, A_VEC = 1, A_DFT = 2 };

struct ArrayRec {
  const void *host;
  char *dev;      // device base
  int chunk;      // owning chunk (grid arrays)
  int kind;
  size_t n_host;  // host elements (A_DFT: complex elements)
  size_t n_dev;   // device elements
  bool mut;       // downloaded by sync_to_host
  bool is_field;  // f[c][cmp] (NaN check)
  bool pinned;    // host array page-locked in place (meepcu_config::pin_host)
  unsigned group; // MEEPCU_SYNC_* group it travels with (meepcu_sync_subset)
  int comp;       // field component of an f[c][cmp] array (-1: something else)
};

struct Geom { // device layout of one chunk
  int dim;
  int num[3], io[3];
  bool has[3];
  long long hs[3]; // host strides X,Y,Z (grid_volume::stride)
  long long ds[3]; // device strides X,Y,Z
  int axis_of_dir[3]; // direction -> device axis (0=A,1=B,2=C), -1 if absent
  int dir_of_axis[3]; // device axis -> direction, -1 if none
  int nax[3];      // points per device axis
  int pitch;
  size_t ntot_host, ntot_dev;
  int yee_shift(int c, int d) const {
    if (!has[d]) return 0;
    const int dc = comp_dir(c);
    if ((is_electric(c) || is_D(c)) && d == dc) return 1;
    if ((is_magnetic(c) || is_B(c)) && d != dc) return 1;
    return 0;
  }
  bool has_field(int c) const {
    if (dim == MEEPCU_D1) return c == 0 || c == 6 || c == 10 || c == 16;
    return comp_dir(c) < MEEPCU_R;
  }
  long long dev_index(long long hidx) const { // host element index -> device element index
    long long r = hidx, o = 0;
    for (int d = 0; d < 3; ++d)
      if (has[d]) {
        long long q = r / hs[d];
        r -= q * hs[d];
        o += q * ds[d];
      }
    return o;
  }
  long long dev_offset(long long hoff) const { // signed relative offset (e.g. yee2cent avg1)
    // decompose into per-direction steps of -1/0/+1 (offsets used by the reference are sums of
    // +-stride(d) over distinct directions)
    long long best = 0;
    for (int sx = -1; sx <= 1; ++sx)
      for (int sy = -1; sy <= 1; ++sy)
        for (int sz = -1; sz <= 1; ++sz) {
          if ((sx && !has[0]) || (sy && !has[1]) || (sz && !has[2])) continue;
          if (sx * hs[0] + sy * hs[1] + sz * hs[2] == hoff) {
            best = sx * ds[0] + sy * ds[1] + sz * ds[2];
            return best;
          }
        }
    return hoff == 0 ? 0 : (long long)1 << 62; // not representable
  }
};

struct SuscH {
  meepcu_susceptibility_desc d; // gyrotropic: sigma, and P/P_prev = the component subtract_P uses (P[c][cmp][dir(c)])
  bool gyro;
  meepcu_gyrotropic_desc g;
  // multilevel atoms: one entry per transition t, holding P[c][cmp][t] with NO sigma: subtract-only entries, in the order
  // subtract_P takes them off D (multilevel-atom.cpp:376-394); the update itself is k_ml_levels / k_ml_pol
  int ml, ml_t; // index into ChunkH::ml (-1: not a multilevel entry), transition
};
// multilevel_susceptibility + multilevel_data (src/multilevel-atom.cpp:118-128, src/meep.hpp:377-386); the small
// arrays are kept as double (exact for either realnum)
struct MlH {
  int L, Tn;
  std::vector<double> Gamma, GammaInv, alpha, omega, gamma, sigmat;
  const void *sigma[MEEPCU_NUM_COMPONENTS][5];
  std::vector<const void *> P[MEEPCU_NUM_COMPONENTS][2], P_prev[MEEPCU_NUM_COMPONENTS][2]; // [t]
  const void *N;
};
struct SrcH {
  int ft, c, tid, integrated;
  std::vector<long long> idx;
  std::vector<double> amp;
};
struct DftH {
  meepcu_dft_desc d;
  std::vector<double> omega;
  size_t N;
  int phase_group;
};
struct ChunkH {
  meepcu_chunk_desc d;
  Geom g;
  std::vector<SuscH> pol[2];
  std::vector<MlH> ml;                                // multilevel media (E_stuff)
  const void *w_prev[MEEPCU_NUM_COMPONENTS][2];       // fields_chunk::f_w_prev (update_eh.cpp:221-226)
  std::vector<SrcH> src;
  std::vector<DftH> dft;
  bool have_plus[20], have_minus[20];
  int plus_c[20], minus_c[20], plus_d[20], minus_d[20];
};
struct PtrList {
  std::vector<void *> dst;
  std::vector<const void *> src;
  std::vector<char> phases;
};

struct CtxBase {
  meepcu_config cfg;
  int t;
  bool finalized;
  virtual ~CtxBase() {}
  virtual int add_chunk(const meepcu_chunk_desc *d, int *id) = 0;
  virtual int add_susc(int chunk, const meepcu_susceptibility_desc *d) = 0;
  virtual int add_gyro(int chunk, const meepcu_gyrotropic_desc *d) = 0;
  virtual int add_multilevel(int chunk, const meepcu_multilevel_desc *d) = 0;
  virtual size_t noise_count() = 0;
  virtual int set_noise(int nsteps, size_t n_per_step, const double *g) = 0;
  virtual int add_source(int chunk, const meepcu_source_desc *d) = 0;
  virtual int add_dft(int chunk, const meepcu_dft_desc *d, int *id) = 0;
  virtual int set_metal(int ft, size_t n, void *const *ptrs) = 0;
  virtual int add_conn(int ft, int kind, size_t n, void *const *dst, const void *const *src,
                       const void *phases) = 0;
  virtual int finalize() = 0;
  virtual int step(int nsteps, int nsrc, const meepcu_src_sample *smp) = 0;
  virtual int synchronize_magnetic_fields(int nsrc, const meepcu_src_sample *smp) = 0;
  virtual int restore_magnetic_fields() = 0;
  virtual int sync_to_host() = 0;
  virtual int sync_from_host() = 0;
  virtual int sync_groups(bool to_host, unsigned groups) = 0;
  virtual int read_values(size_t n, const void *const *ptrs, double *out) = 0;
  virtual int check_finite(int *ok) = 0;
  virtual int read_value(const void *p, double *out) = 0;
  virtual int get_stats(meepcu_stats *out) = 0;
  virtual int get_timers(double out[22]) = 0;
  virtual int add_remote(int ft, int kind, int peer, size_t n, const void *const *ptrs, const void *phases, bool recv) = 0;
  virtual int comm_init(int rank, int nranks, const void *id) = 0;
  virtual int set_profiling(int on) = 0;
  virtual int get_kernel_times(meepcu_kernel_time *out, int cap, int *n) = 0;
  virtual int get_compulsory_bytes(int64_t *per_step, int64_t *per_kernel, int cap, int *n) = 0;
  virtual int dft_dot(size_t npairs, const void *const *a, const void *const *b, int Nomega, double *out) = 0;
  virtual int integrate_dot(size_t n, const meepcu_integrate_desc *pc, double *sum, double *maxabs) = 0;
};

template <class T> struct DevBuf { // simple growable device buffer
  T *p;
  size_t cap;
  DevBuf() : p(0), cap(0) {}
  int reserve(size_t n) {
    if (n <= cap) return MEEPCU_OK;
    if (p) cudaFree(p);
    p = 0;
    cap = 0;
    CU(cudaMalloc((void **)&p, std::max<size_t>(n, 1) * sizeof(T)));
    cap = n;
    return MEEPCU_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = 0;
    cap = 0;
  }
};

template <class T> struct Ctx : CtxBase {
  std::vector<ChunkH> chunks;
  std::vector<void *> metal[MEEPCU_NUM_FIELD_TYPES];
  PtrList conn[MEEPCU_NUM_FIELD_TYPES][3];
  // connections whose other end lives on another rank: [ft] -> peer -> kind -> pointers
  struct RemoteList { std::vector<const void *> ptr[3]; std::vector<char> phases; };
  std::map<int, RemoteList> rsend[MEEPCU_NUM_FIELD_TYPES], rrecv[MEEPCU_NUM_FIELD_TYPES];
  struct RemoteDev { // per field type, after finalize
    const T **d_src; const T **d_mul; T *d_sendbuf; long long n_send;
    T **d_dst; T *d_recvbuf; long long n_recv;
    T *d_phase; long long *d_seg; int nseg;
    std::vector<int> send_peer, recv_peer;
    std::vector<long long> send_off, send_n, recv_off, recv_n;
  };
  RemoteDev rdev[MEEPCU_NUM_FIELD_TYPES];
  nccl_comm comm;
  int my_rank, n_ranks;

  // ---- device state (valid after finalize) ----
  std::map<const char *, ArrayRec> arrays; // by host base address
  char *arena;
  size_t arena_bytes;
  cudaStream_t st, st2; // st2: the PML/general tiles run beside the interior tiles of the same half step
  cudaEvent_t ev0, ev1, ev_fork, ev_join;
  std::vector<char *> extra_dev; // other device allocations to free

  // tables
  HalfParams<T> *d_half;            // [nhalf]
  std::vector<HalfParams<T> > h_half;
  Tile *d_tiles[2][4];              // [half: 0=B/H 1=D/E][class: 0 pml-only, 1 +polarisations, 2 +conductivity, 3 +beta/BFAST]
  int n_tiles[2][4];
  Curl3DParams<T> *d_c3d;           // interior fast path (k_curl3d)
  std::vector<Curl3DParams<T> > h_c3d;
  Tile *d_c3d_tiles[2];
  int n_c3d_tiles[2];
  Tile *d_c3dv_tiles[2];            // clean tiles -> vectorised kernel
  int n_c3dv_tiles[2];
  Tile *d_gv_tiles[2][2];           // general vectorised: [half][0 pml, 1 +polarisations]
  int n_gv_tiles[2][2];
  // Multi-rank overlap: in every tile list the tiles that write a plane another rank receives come
  // first (nb_* of them).  They are launched first, the halo exchange then runs on its own stream
  // beside the remaining (interior) tiles of the same half step.
  int nb_tiles[2][4], nb_c3d_tiles[2], nb_c3dv_tiles[2], nb_gv_tiles[2][2];
  Tile *d_mix_tiles[2];             // k_mix: clean interior tiles and vectorised PML tiles of one half, interleaved
  int n_mix_tiles[2], nb_mix_tiles[2]; // and how many of them are cross-rank boundary tiles
  bool overlap_half[2];
  cudaStream_t st_comm, st_bnd;
  cudaEvent_t ev_bdone, ev_cdone, ev_bstart;
  EhParams<T> *d_eh;
  std::vector<EhParams<T> > h_eh;
  Tile *d_eh_tiles[2];
  int n_eh_tiles[2];
  FmpParams<T> *d_fmp[2];
  int n_fmp[2];
  DipPatch<T> *d_dip[2]; // dipole points patched in place around k_eh_general (kernels.cuh, DipPatch)
  int n_dip[2];
  bool fmp_has_src[2];
  ConnTable<T> d_conn[MEEPCU_NUM_FIELD_TYPES][3];
  ConnRun<T> *d_runs[MEEPCU_NUM_FIELD_TYPES]; // COPY + NEGATE lists in run-length form
  int n_runs[MEEPCU_NUM_FIELD_TYPES];
  // E/H halo copies no stepping kernel reads (the not-owned plane NORMAL to the component: d/dx of Ex
  // appears in no curl): copied only when the host mirror is refreshed, like the B/D/P halos
  ConnRun<T> *d_runs_def[MEEPCU_NUM_FIELD_TYPES];
  int n_runs_def[MEEPCU_NUM_FIELD_TYPES];
  std::vector<T *> metal_axis[MEEPCU_NUM_FIELD_TYPES]; // cylindrical: metal points on the axis row (see metal_masks)
  T **d_zero_late[2];                 // zero_metal(E/H) of chunks whose polarisations update in their own pass
  long long n_zero_late[2];
  T **d_zero[MEEPCU_NUM_FIELD_TYPES]; // unfused metal lists (only non-plane lists; normally empty)
  long long n_zero[MEEPCU_NUM_FIELD_TYPES];
  MlLevelParams<T> *d_mllev;       // multilevel atoms: populations (k_ml_levels) and polarisations (k_ml_pol)
  std::vector<MlLevelParams<T> > h_mllev;
  MlPolParams<T> *d_mlpol;
  std::vector<MlPolParams<T> > h_mlpol;
  Tile *d_mllev_tiles, *d_mlpol_tiles;
  int n_mllev_tiles, n_mlpol_tiles;
  CopyItem<T> *d_wprev;            // W -> W_prev copies at the start of the D/E half (update_eh.cpp:221-226)
  int n_wprev;
  long long wprev_max;
  NoiseSeg<T> *d_noise_seg[2];     // noisy Lorentzian media, per half (k_pol_noise)
  int n_noise_seg[2];
  long long noise_total[2];        // unit Gaussians per time step: [0] update_pols(H_stuff), [1] update_pols(E_stuff)
  std::vector<double> noise_host;  // meepcu_set_noise: the values of the next meepcu_step call
  int noise_steps;
  const double *cur_noise;         // device address of the current step's values
  PolGyroParams<T> *d_polgyro;     // gyrotropic polarisation updates (k_pol_gyro)
  std::vector<PolGyroParams<T> > h_polgyro;
  Tile *d_polgyro_tiles[2];
  int n_polgyro_tiles[2];
  PolGenParams<T> *d_polgen;       // anisotropic-sigma polarisation updates (k_pol_general)
  std::vector<PolGenParams<T> > h_polgen;
  Tile *d_polgen_tiles[2];
  int n_polgen_tiles[2];
  // cylindrical coordinates: per half, the f_rderiv_int pre-pass and the r = 0 rules (kernels.cuh)
  CylRderiv<T> *d_cyl_rd[2];
  int n_cyl_rd[2];
  CylOrigin<T> *d_cyl_org[2];
  int n_cyl_org[2];
  std::set<std::pair<int, int> > aniso_chunk_half; // (chunk, half) whose polarisations update in k_pol_general
  bool general_half[2];  // some chunk (of ANY rank) needs k_eh_general in this half
  bool aniso_half[2];    // some chunk (of any rank) updates polarisations in k_pol_general / k_pol_gyro
  bool eager_db[2];      // exchange B / D halos every step
  bool eager_p[2];       // exchange PH / PE halos every step
  DftParams<T> *d_dft;
  std::vector<DftParams<T> > h_dft;
  std::vector<int> dft_decim, dft_is_h;
  DftWork *d_dft_work;
  int n_dft_work;
  std::vector<int> phase_off;      // per group offset (complex elements) into a step's phase row
  std::vector<std::vector<double> > group_omega;
  std::vector<std::complex<double> > group_scale;
  std::vector<int> group_is_h;
  int phase_row;                   // complex elements per step
  int *d_phase_off;
  // Per-call tables (src_time samples, DFT phases, decimation flags) travel through a page-locked ring and a
  // device ring, so that meepcu_step never waits for the device: the reference's caller makes one call per
  // time step (src/step.cpp:35, python/simulation.py:2831-2834) and the launches of step n+1 must be queued
  // while step n runs.
  struct Ring { char *h, *d; size_t cap, used; } ring;
  static size_t ring_slot(size_t bytes) { return (std::max<size_t>(bytes, 1) + 255) & ~(size_t)255; }
  // Room for ALL the tables of one meepcu_step call at once: a wrap (or a re-allocation) between two ring_put calls of the
  // same meepcu_step would hand the second table the slot of the first (seen with the noise table of noisy media, the
  // first one big enough to wrap: the src_time samples of the same call were overwritten)
  int ring_reserve(size_t total) {
    if (ring.used + total <= ring.cap) return MEEPCU_OK;
    CU(cudaStreamSynchronize(st)); // every kernel that read earlier slots has finished (side streams join st)
    ring.used = 0;
    if (total > ring.cap) {
      if (ring.h) cudaFreeHost(ring.h);
      if (ring.d) cudaFree(ring.d);
      ring.h = ring.d = 0;
      ring.cap = 0;
      const size_t cap = std::max<size_t>(2 * total, (size_t)1 << 20);
      CU(cudaHostAlloc((void **)&ring.h, cap, cudaHostAllocDefault));
      CU(cudaMalloc((void **)&ring.d, cap));
      ring.cap = cap;
    }
    return MEEPCU_OK;
  }
  int ring_put(const void *src, size_t bytes, char **dptr) {
    const size_t need = ring_slot(bytes);
    if (ring.used + need > ring.cap) {
      CU(cudaStreamSynchronize(st)); // every kernel that read earlier slots has finished (side streams join st)
      ring.used = 0;
      if (need > ring.cap) {
        if (ring.h) cudaFreeHost(ring.h);
        if (ring.d) cudaFree(ring.d);
        ring.h = ring.d = 0;
        ring.cap = 0;
        const size_t cap = std::max<size_t>(2 * need, (size_t)1 << 20);
        CU(cudaHostAlloc((void **)&ring.h, cap, cudaHostAllocDefault));
        CU(cudaMalloc((void **)&ring.d, cap));
        ring.cap = cap;
      }
    }
    if (bytes) {
      memcpy(ring.h + ring.used, src, bytes);
      CU(cudaMemcpyAsync(ring.d + ring.used, ring.h + ring.used, bytes, cudaMemcpyHostToDevice, st));
    }
    *dptr = ring.d + ring.used;
    ring.used += need;
    return MEEPCU_OK;
  }
  int *d_flags; // [0] nr_fail, [1] nonfinite
  // the per-step NaN test (src/step.cpp:138-139), lazily: a sample of every field array is tested at the end
  // of each meepcu_step call by one small kernel that raises a flag in page-locked host memory; the host looks
  // at the flag when the next call comes in, never waiting for the device
  int *h_nonfinite, *d_hnonfinite;
  FinEntry *d_fin;
  int n_fin;
  long long fin_max_n;
  bool finite_check;
  // device time of meepcu_step calls, collected lazily (CUDA event pairs, oldest first)
  struct StepEv { cudaEvent_t e0, e1; bool pending; };
  std::vector<StepEv> step_ev;
  size_t step_ev_next;
  double total_step_ms;
  bool halo_stale; // deferred local halo copies (B/D/P, normal-plane E/H) not materialised since the last step
  // ---- virtual E (kernels.cuh, Curl3DParams::virt): chunks whose E arrays are not stored while stepping ----
  struct VirtComp { T *e; const T *d; T *u; const SrcEntry *src; int nsrc, cmp, chunk, comp; long long n; };
  std::vector<VirtComp> virt_comps;
  std::map<const char *, const char *> virt_alias; // host base of a virtual E array -> host base of its D array
  std::map<const char *, const char *> virt_mul;   // ... -> host base of its chi1inv array (absent: E = D)
  std::set<int> virt_chunks;
  bool virt_any;        // some chunk of SOME rank is virtual (the sync sequence is then the same on every rank)
  bool virt_vu;         // some virtual chunk of this rank multiplies by chi1inv
  bool e_virtual;       // stepped since the E arrays were last uploaded: those of virtual chunks are out of date
  std::vector<double> last_row; // src_time samples of the last step taken (its dipole_E is in the stored E)
  std::vector<long long> h_src_idx; // device index of every entry of d_src
  SrcEntry *d_src;

  // synchronize_magnetic_fields state (energy_and_flux.cpp:97-177)
  struct MagArr { char *dev; size_t bytes, off; bool is_f, is_pol; };
  std::vector<MagArr> mag_arrays; // B and H arrays with their f_u, f_w, f_cond, f_bfast (+ magnetic P, P_prev)
  char *mag_backup;
  bool mag_synced, in_mag_sync;

  int64_t launches, steps_done, cells, alg_bytes;
  double last_ms;
  int64_t nr_unconverged = 0; // cells left unchanged after 100 NR attempts (see check_finite)

  // per-kernel-class profiling (CUDA events around every launch, on the library's stream)
  enum { KC_HALF_B_SIMPLE = 0, KC_HALF_B_GENERAL, KC_HALF_D_SIMPLE, KC_HALF_D_GENERAL, KC_CONN,
         KC_FMP, KC_EH_GENERAL, KC_DFT, KC_CURL3D_B, KC_CURL3D_D, KC_CURL3D_SPECIAL, KC_MIX_B, KC_MIX_D, KC_POL_GENERAL, KC_COUNT };
  bool profiling;
  std::vector<cudaEvent_t> ev_pool;
  std::vector<int> ev_class;      // class of the launch between ev_pool[2k] and ev_pool[2k+1]
  size_t ev_used;
  double kc_ms[KC_COUNT];
  int64_t kc_launches[KC_COUNT];
  int64_t kc_bytes[KC_COUNT];
  // the same accounting with SURVEY.md 8(d)'s per-cell figures (E stored in the D/E half and read back in the B/H half:
  // 21 / 24 values per vacuum / dielectric cell) -- what `roofline.achieved` is quoted on; kc_bytes / alg_bytes count
  // what the kernels really have to move (virtual-E chunks: 18 / 21)
  int64_t kc_bytes_sv[KC_COUNT], alg_bytes_sv;
  double sink_ms[22];

  Ctx() : arena(0), arena_bytes(0), st(0), st2(0), ev0(0), ev1(0), ev_fork(0), ev_join(0), d_half(0), d_eh(0), d_dft(0),
          d_dft_work(0), n_dft_work(0), phase_row(0), d_phase_off(0), d_flags(0), d_src(0),
          launches(0), steps_done(0), cells(0), alg_bytes(0), last_ms(0), profiling(false), ev_used(0) {
    stage[0] = stage[1] = 0;
    dstage = 0;
    ring.h = ring.d = 0;
    ring.cap = ring.used = 0;
    h_nonfinite = d_hnonfinite = 0;
    d_fin = 0;
    n_fin = 0;
    fin_max_n = 0;
    finite_check = true;
    step_ev_next = 0;
    total_step_ms = 0;
    halo_stale = false;
    virt_any = virt_vu = e_virtual = false;
    mag_backup = 0;
    mag_synced = in_mag_sync = false;
    stage_next = 0;
    comm = 0;
    my_rank = 0;
    n_ranks = 1;
    memset(rdev, 0, sizeof(RemoteDev) * 0); // vectors inside: value-initialised below
    for (int ft = 0; ft < MEEPCU_NUM_FIELD_TYPES; ++ft) { rdev[ft].d_src = 0; rdev[ft].d_mul = 0; rdev[ft].d_sendbuf = 0; rdev[ft].n_send = 0; rdev[ft].d_dst = 0; rdev[ft].d_recvbuf = 0; rdev[ft].n_recv = 0; rdev[ft].d_phase = 0; rdev[ft].d_seg = 0; rdev[ft].nseg = 0; }
    memset(kc_ms, 0, sizeof kc_ms);
    memset(kc_launches, 0, sizeof kc_launches);
    memset(kc_bytes, 0, sizeof kc_bytes);
    memset(kc_bytes_sv, 0, sizeof kc_bytes_sv);
    alg_bytes_sv = 0;
    memset(sink_ms, 0, sizeof sink_ms);
    memset(d_tiles, 0, sizeof d_tiles);
    d_c3d = 0;
    memset(d_c3d_tiles, 0, sizeof d_c3d_tiles);
    memset(n_c3d_tiles, 0, sizeof n_c3d_tiles);
    memset(d_c3dv_tiles, 0, sizeof d_c3dv_tiles);
    memset(n_c3dv_tiles, 0, sizeof n_c3dv_tiles);
    memset(d_gv_tiles, 0, sizeof d_gv_tiles);
    memset(n_gv_tiles, 0, sizeof n_gv_tiles);
    memset(nb_tiles, 0, sizeof nb_tiles);
    memset(nb_c3d_tiles, 0, sizeof nb_c3d_tiles);
    memset(nb_c3dv_tiles, 0, sizeof nb_c3dv_tiles);
    memset(nb_gv_tiles, 0, sizeof nb_gv_tiles);
    memset(d_mix_tiles, 0, sizeof d_mix_tiles);
    memset(n_mix_tiles, 0, sizeof n_mix_tiles);
    memset(nb_mix_tiles, 0, sizeof nb_mix_tiles);
    overlap_half[0] = overlap_half[1] = false;
    st_comm = st_bnd = 0;
    ev_bdone = ev_cdone = ev_bstart = 0;
    memset(n_tiles, 0, sizeof n_tiles);
    memset(d_eh_tiles, 0, sizeof d_eh_tiles);
    memset(n_eh_tiles, 0, sizeof n_eh_tiles);
    memset(d_dip, 0, sizeof d_dip);
    memset(n_dip, 0, sizeof n_dip);
    memset(d_fmp, 0, sizeof d_fmp);
    memset(n_fmp, 0, sizeof n_fmp);
    memset(d_conn, 0, sizeof d_conn);
    memset(d_runs, 0, sizeof d_runs);
    memset(n_runs, 0, sizeof n_runs);
    memset(d_runs_def, 0, sizeof d_runs_def);
    memset(n_runs_def, 0, sizeof n_runs_def);
    memset(d_zero_late, 0, sizeof d_zero_late);
    memset(n_zero_late, 0, sizeof n_zero_late);
    d_mllev = 0;
    d_mlpol = 0;
    d_mllev_tiles = d_mlpol_tiles = 0;
    n_mllev_tiles = n_mlpol_tiles = 0;
    d_wprev = 0;
    n_wprev = 0;
    wprev_max = 0;
    memset(d_noise_seg, 0, sizeof d_noise_seg);
    memset(n_noise_seg, 0, sizeof n_noise_seg);
    memset(noise_total, 0, sizeof noise_total);
    noise_steps = 0;
    cur_noise = 0;
    d_polgyro = 0;
    memset(d_polgyro_tiles, 0, sizeof d_polgyro_tiles);
    memset(n_polgyro_tiles, 0, sizeof n_polgyro_tiles);
    d_polgen = 0;
    memset(d_polgen_tiles, 0, sizeof d_polgen_tiles);
    memset(n_polgen_tiles, 0, sizeof n_polgen_tiles);
    memset(d_cyl_rd, 0, sizeof d_cyl_rd);
    memset(n_cyl_rd, 0, sizeof n_cyl_rd);
    memset(d_cyl_org, 0, sizeof d_cyl_org);
    memset(n_cyl_org, 0, sizeof n_cyl_org);
    memset(d_zero, 0, sizeof d_zero);
    memset(n_zero, 0, sizeof n_zero);
    fmp_has_src[0] = fmp_has_src[1] = false;
    general_half[0] = general_half[1] = false;
    aniso_half[0] = aniso_half[1] = false;
    eager_db[0] = eager_db[1] = eager_p[0] = eager_p[1] = false;
  }
  ~Ctx() {
    for (int k = 0; k < 2; ++k) {
      if (stage[k]) cudaFreeHost(stage[k]);
    }
    if (dstage) cudaFree(dstage);
    if (mag_backup) cudaFree(mag_backup);
    if (comm && nccl_api() && nccl_api()->CommDestroy) nccl_api()->CommDestroy(comm);
    for (std::map<const char *, ArrayRec>::iterator it = arrays.begin(); it != arrays.end(); ++it)
      if (it->second.pinned) cudaHostUnregister((void *)it->second.host);
    if (arena) cudaFree(arena);
    for (size_t i = 0; i < extra_dev.size(); ++i) cudaFree(extra_dev[i]);
    if (ring.h) cudaFreeHost(ring.h);
    if (ring.d) cudaFree(ring.d);
    if (h_nonfinite) cudaFreeHost(h_nonfinite);
    for (size_t i = 0; i < step_ev.size(); ++i) { cudaEventDestroy(step_ev[i].e0); cudaEventDestroy(step_ev[i].e1); }
    for (size_t i = 0; i < ev_pool.size(); ++i) cudaEventDestroy(ev_pool[i]);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (ev_bdone) cudaEventDestroy(ev_bdone);
    if (ev_cdone) cudaEventDestroy(ev_cdone);
    if (ev_bstart) cudaEventDestroy(ev_bstart);
    if (st_comm) cudaStreamDestroy(st_comm);
    if (st_bnd) cudaStreamDestroy(st_bnd);
    if (st2) cudaStreamDestroy(st2);
    if (st) cudaStreamDestroy(st);
  }

  // gv.origin_r() * gv.a with the reference's roundings: origin = io * (0.5 * inva), inva = 1 / a (vec.cpp:38-41)
  double origin_r_times_a(const Geom &g) const {
    const double inva = 1.0 / cfg.a;
    const double origin_r = g.io[0] * (0.5 * inva);
    return origin_r * cfg.a;
  }
  // slot of the k-th loop level of LOOP_OVER_IVECS (yucky_dir, vec.cpp:262-272): X,Y,Z; 2-D: Z,X,Y; Dcyl: P,R,Z
  int loop_slot(int k) const {
    if (cfg.dim == MEEPCU_Dcyl) return k == 0 ? 1 : (k == 1 ? 0 : 2);
    return cfg.dim == MEEPCU_D2 ? (k + 2) % 3 : k;
  }

  // ------------------------------------------------------------------------------------------
  int add_chunk(const meepcu_chunk_desc *d, int *id) {
    if (finalized) return fail(MEEPCU_ESTATE, "context already finalized");
    ChunkH ch;
    ch.d = *d;
    memset(ch.w_prev, 0, sizeof ch.w_prev);
    if (cfg.dim == MEEPCU_Dcyl) canon_chunk_desc(ch.d);
    d = &ch.d;
    Geom &g = ch.g;
    g.dim = cfg.dim;
    g.has[0] = (cfg.dim == MEEPCU_D2 || cfg.dim == MEEPCU_D3 || cfg.dim == MEEPCU_Dcyl);
    g.has[1] = (cfg.dim == MEEPCU_D2 || cfg.dim == MEEPCU_D3);
    g.has[2] = (cfg.dim == MEEPCU_D1 || cfg.dim == MEEPCU_D3 || cfg.dim == MEEPCU_Dcyl);
    for (int k = 0; k < 3; ++k) {
      g.num[k] = g.has[k] ? d->num[k] : 0;
      g.io[k] = g.has[k] ? d->io[k] : 0;
      if (g.num[k] < 0) return fail(MEEPCU_EINVAL, "negative grid size");
    }
    // host strides: set_strides, src/vec.cpp:482-494
    g.hs[0] = g.hs[1] = g.hs[2] = 0;
    if (g.has[2]) g.hs[2] = 1;
    if (g.has[1]) g.hs[1] = g.num[2] + 1;
    if (g.has[0]) g.hs[0] = (long long)(g.num[2] + 1) * (g.num[1] + 1);
    g.ntot_host = 1;
    for (int k = 0; k < 3; ++k)
      if (g.has[k]) g.ntot_host *= (size_t)(g.num[k] + 1);
    // device axes: existing directions by decreasing host stride, right-aligned on (A,B,C)
    int dirs[3], nd = 0;
    for (int k = 0; k < 3; ++k)
      if (g.has[k]) dirs[nd++] = k; // X,Y,Z order == decreasing stride
    for (int k = 0; k < 3; ++k) {
      g.axis_of_dir[k] = -1;
      g.dir_of_axis[k] = -1;
      g.nax[k] = 1;
      g.ds[k] = 0;
    }
    for (int j = 0; j < nd; ++j) {
      const int ax = 3 - nd + j;
      g.axis_of_dir[dirs[j]] = ax;
      g.dir_of_axis[ax] = dirs[j];
      g.nax[ax] = g.num[dirs[j]] + 1;
    }
    // fastest axis padded so that rows start on a 16-byte vector boundary.  Short rows (e.g. the
    // 12-point-thick PML slabs): padding them to 32 elements would leave most lanes of every warp idle,
    // and padding fp32 rows to a whole 32-byte sector moved 16 floats of DRAM traffic for 11-12 useful ones.
    {
      // long rows: measured at 256^3 (257-point rows) -- fp32 0.3597 / 0.3546 / 0.3527 ms per step for rows
      // padded to 32 / 8 / 4 elements (fewer idle lanes beats line alignment), fp64 0.6329 / 0.6294 / 0.6325
      // for 32 / 16 / 2
      int q = g.nax[2] >= 64 ? (sizeof(T) == 4 ? 4 : 16) : (int)(16 / sizeof(T));
      if (g.nax[2] >= 64 && getenv("MEEPCU_PITCH_Q")) q = std::max((int)(16 / sizeof(T)), atoi(getenv("MEEPCU_PITCH_Q")));
      g.pitch = ((g.nax[2] + q - 1) / q) * q;
    }
    const long long sax[3] = {(long long)g.pitch * g.nax[1], (long long)g.pitch, 1};
    for (int k = 0; k < 3; ++k)
      if (g.has[k]) g.ds[k] = sax[g.axis_of_dir[k]];
    g.ntot_dev = (size_t)g.pitch * g.nax[1] * g.nax[0];

    // figure_out_step_plan, src/fields.cpp:441-469
    for (int c = 0; c < 20; ++c) ch.have_plus[c] = ch.have_minus[c] = false;
    for (int c1 = 0; c1 < 20; ++c1) {
      if (!d->f[c1][0]) continue;
      const int dc1 = comp_dir(c1);
      if (dc1 >= MEEPCU_R) return fail(MEEPCU_EUNSUPPORTED, "cylindrical field component");
      for (int c2 = 0; c2 < 20; ++c2) {
        const bool pair = ((is_electric(c1) || is_D(c1)) && is_magnetic(c2)) ||
                          ((is_magnetic(c1) || is_B(c1)) && is_electric(c2));
        if (!pair) continue;
        const int dc2 = comp_dir(c2);
        if (dc2 >= MEEPCU_R || dc1 == dc2) continue;
        const int dd = (3 + 2 * dc1 - dc2) % 3; // cross(dc1, dc2)
        if (!(g.has_field(c2) && g.has_field(c1) && g.has[dd])) continue;
        if (((3 + dc1 - dc2) % 3) == 2) { // cross_negative(dc2, dc1)
          ch.minus_c[c1] = c2; ch.have_minus[c1] = true; ch.minus_d[c1] = dd;
        }
        else {
          ch.plus_c[c1] = c2; ch.have_plus[c1] = true; ch.plus_d[c1] = dd;
        }
      }
    }
    chunks.push_back(ch);
    if (id) *id = (int)chunks.size() - 1;
    return MEEPCU_OK;
  }
  int add_susc(int chunk, const meepcu_susceptibility_desc *d) {
    if (finalized) return fail(MEEPCU_ESTATE, "context already finalized");
    if (chunk < 0 || chunk >= (int)chunks.size()) return fail(MEEPCU_EINVAL, "bad chunk id");
    if (d->ft != MEEPCU_E_STUFF && d->ft != MEEPCU_H_STUFF)
      return fail(MEEPCU_EINVAL, "susceptibility ft must be E_stuff or H_stuff");
    if (d->noise_amp != 0 && cfg.dim == MEEPCU_Dcyl)
      return fail(MEEPCU_EUNSUPPORTED, "noisy Lorentzian media in cylindrical coordinates");
    SuscH s;
    memset(&s, 0, sizeof s);
    s.ml = -1;
    s.d = *d;
    if (cfg.dim == MEEPCU_Dcyl) canon_susc_desc(s.d);
    chunks[chunk].pol[d->ft].push_back(s);
    if (chunks[chunk].pol[d->ft].size() > (size_t)kMaxPol)
      return fail(MEEPCU_EUNSUPPORTED, "more than %d susceptibilities per chunk and field type", kMaxPol);
    return MEEPCU_OK;
  }
  int add_gyro(int chunk, const meepcu_gyrotropic_desc *d) {
    if (finalized) return fail(MEEPCU_ESTATE, "context already finalized");
    if (chunk < 0 || chunk >= (int)chunks.size()) return fail(MEEPCU_EINVAL, "bad chunk id");
    if (d->ft != MEEPCU_E_STUFF && d->ft != MEEPCU_H_STUFF)
      return fail(MEEPCU_EINVAL, "susceptibility ft must be E_stuff or H_stuff");
    if (d->model < 0 || d->model > 2) return fail(MEEPCU_EINVAL, "bad gyrotropy model");
    if (cfg.dim != MEEPCU_D3) return fail(MEEPCU_EINVAL, "gyrotropic media require 3D Cartesian fields");
    SuscH s;
    memset(&s, 0, sizeof s);
    s.ml = -1;
    s.gyro = true;
    s.g = *d;
    s.d.ft = d->ft;
    for (int c = 0; c < 20; ++c) {
      for (int k = 0; k < 5; ++k) s.d.sigma[c][k] = d->sigma[c][k];
      const int d0 = comp_dir(c);
      if (d0 < MEEPCU_R)
        for (int cmp = 0; cmp < 2; ++cmp) { // subtract_P uses P[ec][cmp][component_direction(ec)] (:586-603)
          s.d.P[c][cmp] = d->P[c][cmp][d0];
          s.d.P_prev[c][cmp] = d->P_prev[c][cmp][d0];
        }
    }
    chunks[chunk].pol[d->ft].push_back(s);
    if (chunks[chunk].pol[d->ft].size() > (size_t)kMaxPol)
      return fail(MEEPCU_EUNSUPPORTED, "more than %d susceptibilities per chunk and field type", kMaxPol);
    return MEEPCU_OK;
  }
  int add_multilevel(int chunk, const meepcu_multilevel_desc *d) {
    if (finalized) return fail(MEEPCU_ESTATE, "context already finalized");
    if (chunk < 0 || chunk >= (int)chunks.size()) return fail(MEEPCU_EINVAL, "bad chunk id");
    if (d->ft != MEEPCU_E_STUFF) return fail(MEEPCU_EUNSUPPORTED, "multilevel media are supported for E_stuff only");
    if (cfg.dim == MEEPCU_Dcyl) return fail(MEEPCU_EUNSUPPORTED, "multilevel media in cylindrical coordinates");
    if (d->L < 1 || d->L > kMlMaxL || d->T < 1 || d->T > kMlMaxT)
      return fail(MEEPCU_EUNSUPPORTED, "multilevel medium: 1..%d levels, 1..%d transitions", kMlMaxL, kMlMaxT);
    if (!d->Gamma || !d->GammaInv || !d->alpha || !d->omega || !d->gamma || !d->sigmat || !d->N)
      return fail(MEEPCU_EINVAL, "multilevel medium: Gamma, GammaInv, alpha, omega, gamma, sigmat and N are required");
    ChunkH &ch = chunks[chunk];
    if (ch.pol[d->ft].size() + (size_t)d->T > (size_t)kMaxPol)
      return fail(MEEPCU_EUNSUPPORTED, "more than %d polarisations per chunk and field type (a multilevel medium counts once per transition)", kMaxPol);
    MlH M;
    const int L = d->L, Tn = d->T;
    M.L = L;
    M.Tn = Tn;
    M.Gamma.assign((const T *)d->Gamma, (const T *)d->Gamma + L * L);
    M.GammaInv.assign((const T *)d->GammaInv, (const T *)d->GammaInv + L * L);
    M.alpha.assign((const T *)d->alpha, (const T *)d->alpha + L * Tn);
    M.omega.assign((const T *)d->omega, (const T *)d->omega + Tn);
    M.gamma.assign((const T *)d->gamma, (const T *)d->gamma + Tn);
    M.sigmat.assign((const T *)d->sigmat, (const T *)d->sigmat + 5 * Tn);
    M.N = d->N;
    for (int c = 0; c < 20; ++c) {
      for (int k = 0; k < 5; ++k) M.sigma[c][k] = d->sigma[c][k];
      for (int cmp = 0; cmp < 2; ++cmp) {
        if ((d->P[c][cmp] != 0) != (d->P_prev[c][cmp] != 0)) return fail(MEEPCU_EINVAL, "P and P_prev must both be given");
        if (d->P[c][cmp])
          for (int t = 0; t < Tn; ++t) {
            if (!d->P[c][cmp][t] || !d->P_prev[c][cmp][t]) return fail(MEEPCU_EINVAL, "multilevel medium: NULL transition array");
            M.P[c][cmp].push_back(d->P[c][cmp][t]);
            M.P_prev[c][cmp].push_back(d->P_prev[c][cmp][t]);
          }
        if (d->W_prev[c][cmp]) ch.w_prev[c][cmp] = d->W_prev[c][cmp];
      }
    }
    for (int t = 0; t < Tn; ++t) {
      SuscH s;
      memset(&s, 0, sizeof s);
      s.d.ft = d->ft;
      s.ml = (int)ch.ml.size();
      s.ml_t = t;
      for (int c = 0; c < 20; ++c)
        for (int cmp = 0; cmp < 2; ++cmp)
          if (d->P[c][cmp]) {
            s.d.P[c][cmp] = d->P[c][cmp][t];
            s.d.P_prev[c][cmp] = d->P_prev[c][cmp][t];
          }
      ch.pol[d->ft].push_back(s);
    }
    ch.ml.push_back(M);
    return MEEPCU_OK;
  }
  // noisy Lorentzian media: the owned points of every noisy polarisation component, in the reference's draw order
  // (include/meepcu.h at meepcu_set_noise)
  size_t noise_points(const ChunkH &ch, int ft) const {
    size_t n = 0;
    for (size_t ip = 0; ip < ch.pol[ft].size(); ++ip) {
      const meepcu_susceptibility_desc &sd = ch.pol[ft][ip].d;
      if (ch.pol[ft][ip].gyro || ch.pol[ft][ip].ml >= 0 || sd.noise_amp == 0) continue;
      for (int c = 0; c < 20; ++c)
        for (int cmp = 0; cmp < 2; ++cmp)
          if (sd.P[c][cmp] && comp_dir(c) < MEEPCU_R && sd.sigma[c][comp_dir(c)]) {
            size_t m = 1;
            for (int dd = 0; dd < 3; ++dd)
              if (ch.g.has[dd]) m *= (size_t)ch.g.num[dd];
            n += m;
          }
    }
    return n;
  }
  size_t noise_count() {
    size_t n = 0;
    for (int ft = MEEPCU_H_STUFF; ft >= MEEPCU_E_STUFF; --ft)
      for (size_t ic = 0; ic < chunks.size(); ++ic) n += noise_points(chunks[ic], ft);
    return n;
  }
  int set_noise(int nsteps, size_t n_per_step, const double *g) {
    if (n_per_step != noise_count())
      return fail(MEEPCU_EINVAL, "meepcu_set_noise: %zu values per step given, the context needs %zu", n_per_step, noise_count());
    if (nsteps < 0 || (nsteps > 0 && n_per_step > 0 && !g)) return fail(MEEPCU_EINVAL, "meepcu_set_noise: bad arguments");
    noise_host.assign(g, g + (size_t)nsteps * n_per_step);
    noise_steps = nsteps;
    return MEEPCU_OK;
  }
  int add_source(int chunk, const meepcu_source_desc *d) {
    if (finalized) return fail(MEEPCU_ESTATE, "context already finalized");
    if (chunk < 0 || chunk >= (int)chunks.size()) return fail(MEEPCU_EINVAL, "bad chunk id");
    if (d->ft != MEEPCU_D_STUFF && d->ft != MEEPCU_B_STUFF)
      return fail(MEEPCU_EINVAL, "source ft must be D_stuff or B_stuff");
    SrcH s;
    s.ft = d->ft;
    s.c = cfg.dim == MEEPCU_Dcyl ? canon_comp(d->c) : d->c;
    s.tid = d->src_time_id;
    s.integrated = d->is_integrated;
    s.idx.assign(d->index, d->index + d->n);
    s.amp.assign(d->amp, d->amp + 2 * d->n);
    chunks[chunk].src.push_back(s);
    return MEEPCU_OK;
  }
  int add_dft(int chunk, const meepcu_dft_desc *d, int *id) {
    if (finalized) return fail(MEEPCU_ESTATE, "context already finalized");
    if (chunk < 0 || chunk >= (int)chunks.size()) return fail(MEEPCU_EINVAL, "bad chunk id");
    if (!d->dft) return fail(MEEPCU_EINVAL, "dft_chunk::dft host array required");
    if (d->decimation_factor < 1) return fail(MEEPCU_EINVAL, "decimation_factor must be >= 1");
    DftH f;
    f.d = *d;
    if (cfg.dim == MEEPCU_Dcyl && f.d.c >= 0 && f.d.c < 20) f.d.c = canon_comp(f.d.c);
    f.omega.assign(d->omega, d->omega + d->Nomega);
    f.N = 1;
    for (int k = 0; k < 3; ++k)
      if (chunks[chunk].g.has[k]) f.N *= (size_t)((d->ie[k] - d->is[k]) / 2 + 1);
    f.phase_group = -1;
    chunks[chunk].dft.push_back(f);
    if (id) *id = (int)chunks[chunk].dft.size() - 1;
    return MEEPCU_OK;
  }
  int set_metal(int ft, size_t n, void *const *ptrs) {
    if (finalized) return fail(MEEPCU_ESTATE, "context already finalized");
    if (ft < 0 || ft >= MEEPCU_NUM_FIELD_TYPES) return fail(MEEPCU_EINVAL, "bad field type");
    metal[ft].assign(ptrs, ptrs + n);
    return MEEPCU_OK;
  }
  int add_conn(int ft, int kind, size_t n, void *const *dst, const void *const *src,
               const void *phases) {
    if (finalized) return fail(MEEPCU_ESTATE, "context already finalized");
    if (ft < 0 || ft >= MEEPCU_NUM_FIELD_TYPES || kind < 0 || kind > 2)
      return fail(MEEPCU_EINVAL, "bad connection family");
    PtrList &c = conn[ft][kind];
    c.dst.insert(c.dst.end(), dst, dst + n);
    c.src.insert(c.src.end(), src, src + n);
    if (kind == MEEPCU_CONNECT_PHASE) {
      if (n % 2) return fail(MEEPCU_EINVAL, "PHASE connections come in (re,im) pairs");
      const char *p = (const char *)phases;
      c.phases.insert(c.phases.end(), p, p + (n / 2) * 2 * sizeof(T));
    }
    return MEEPCU_OK;
  }

  int add_remote(int ft, int kind, int peer, size_t n, const void *const *ptrs, const void *phases, bool recv) {
    if (finalized) return fail(MEEPCU_ESTATE, "context already finalized");
    if (ft < 0 || ft >= MEEPCU_NUM_FIELD_TYPES || kind < 0 || kind > 2 || peer < 0)
      return fail(MEEPCU_EINVAL, "bad remote connection family");
    RemoteList &L = (recv ? rrecv : rsend)[ft][peer];
    L.ptr[kind].insert(L.ptr[kind].end(), ptrs, ptrs + n);
    if (recv && kind == MEEPCU_CONNECT_PHASE) {
      if (n % 2) return fail(MEEPCU_EINVAL, "PHASE connections come in (re,im) pairs");
      const char *p = (const char *)phases;
      L.phases.insert(L.phases.end(), p, p + (n / 2) * 2 * sizeof(T));
    }
    return MEEPCU_OK;
  }
  int comm_init(int rank, int nranks, const void *id) {
    NcclApi *n = nccl_api();
    if (!n) return fail(MEEPCU_ENCCL, "libnccl.so.2 not found");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(MEEPCU_EINVAL, "bad rank");
    CU(cudaSetDevice(cfg.device));
    nccl_uid uid;
    memcpy(&uid, id, sizeof uid);
    NC(n->CommInitRank(&comm, nranks, uid, rank));
    my_rank = rank;
    n_ranks = nranks;
    return MEEPCU_OK;
  }
  int build_remote() {
    for (int ft = 0; ft < MEEPCU_NUM_FIELD_TYPES; ++ft) {
      RemoteDev &R = rdev[ft];
      if (rsend[ft].empty() && rrecv[ft].empty()) continue;
      if (!comm) return fail(MEEPCU_ESTATE, "remote connections given but meepcu_comm_init not called");
      std::vector<const T *> hsrc, hmul;
      bool any_mul = false;
      for (typename std::map<int, RemoteList>::iterator it = rsend[ft].begin(); it != rsend[ft].end(); ++it) {
        R.send_peer.push_back(it->first);
        R.send_off.push_back((long long)hsrc.size());
        for (int kind = 0; kind < 3; ++kind)
          for (size_t k = 0; k < it->second.ptr[kind].size(); ++k) {
            const T *m = 0;
            const T *d = dev_elem_v(it->second.ptr[kind][k], ft, &m);
            if (!d) return fail(MEEPCU_EINVAL, "remote send pointer outside the registered arrays");
            hsrc.push_back(d);
            hmul.push_back(m);
            if (m) any_mul = true;
          }
        R.send_n.push_back((long long)hsrc.size() - R.send_off.back());
      }
      std::vector<T *> hdst;
      std::vector<T> hphase;
      std::vector<long long> seg;
      for (typename std::map<int, RemoteList>::iterator it = rrecv[ft].begin(); it != rrecv[ft].end(); ++it) {
        R.recv_peer.push_back(it->first);
        R.recv_off.push_back((long long)hdst.size());
        seg.push_back((long long)hdst.size());
        for (int kind = 0; kind < 3; ++kind) {
          seg.push_back((long long)it->second.ptr[kind].size());
          for (size_t k = 0; k < it->second.ptr[kind].size(); ++k) {
            const T *mx = 0;
            T *d = dev_elem_v(it->second.ptr[kind][k], ft, &mx);
            if (!d) return fail(MEEPCU_EINVAL, "remote recv pointer outside the registered arrays");
            hdst.push_back(d);
          }
        }
        seg.push_back((long long)(hphase.size() / 2));
        const T *ph = (const T *)it->second.phases.data();
        hphase.insert(hphase.end(), ph, ph + it->second.phases.size() / sizeof(T));
        R.recv_n.push_back((long long)hdst.size() - R.recv_off.back());
      }
      R.n_send = (long long)hsrc.size();
      R.n_recv = (long long)hdst.size();
      R.nseg = (int)R.recv_peer.size();
      RC(upload(hsrc, &R.d_src));
      if (any_mul) RC(upload(hmul, &R.d_mul));
      RC(upload(hdst, &R.d_dst));
      RC(upload(hphase, &R.d_phase));
      RC(upload(seg, &R.d_seg));
      if (R.n_send) { CU(cudaMalloc((void **)&R.d_sendbuf, R.n_send * sizeof(T))); extra_dev.push_back((char *)R.d_sendbuf); }
      if (R.n_recv) { CU(cudaMalloc((void **)&R.d_recvbuf, R.n_recv * sizeof(T))); extra_dev.push_back((char *)R.d_recvbuf); }
    }
    return MEEPCU_OK;
  }
  int launch_remote(int ft, cudaStream_t sx) {
    RemoteDev &R = rdev[ft];
    if (!R.n_send && !R.n_recv) return MEEPCU_OK;
    NcclApi *n = nccl_api();
    const bool prof = sx == st;
    if (prof) RC(prof_begin(KC_CONN));
    if (R.n_send) {
      const int grid = (int)std::min<long long>((R.n_send + kThreads - 1) / kThreads, 148 * 8);
      k_pack<T><<<grid, kThreads, 0, sx>>>(R.d_src, R.d_sendbuf, R.n_send, R.d_mul);
      launches++;
    }
    const int dtype = sizeof(T) == 8 ? 8 /* ncclFloat64 */ : 7 /* ncclFloat32 */;
    NC(n->GroupStart());
    for (size_t k = 0; k < R.recv_peer.size(); ++k)
      NC(n->Recv(R.d_recvbuf + R.recv_off[k], (size_t)R.recv_n[k], dtype, R.recv_peer[k], comm, sx));
    for (size_t k = 0; k < R.send_peer.size(); ++k)
      NC(n->Send(R.d_sendbuf + R.send_off[k], (size_t)R.send_n[k], dtype, R.send_peer[k], comm, sx));
    NC(n->GroupEnd());
    if (R.n_recv) {
      const int grid = (int)std::min<long long>((R.n_recv + kThreads - 1) / kThreads, 148 * 8);
      k_unpack<T><<<grid, kThreads, 0, sx>>>(R.d_dst, R.d_recvbuf, R.d_phase, R.d_seg, R.nseg);
      launches++;
    }
    if (prof) RC(prof_end());
    return MEEPCU_OK;
  }

  // ------------------------------------------------------------------------------------------
  // array registry
  int reg(const void *host, int chunk, int kind, size_t n_host, size_t n_dev, bool mut,
          bool is_field = false, unsigned group = MEEPCU_SYNC_MATERIALS, int comp = -1) {
    if (!host) return MEEPCU_OK;
    const char *key = (const char *)host;
    std::map<const char *, ArrayRec>::iterator it = arrays.find(key);
    if (it != arrays.end()) {
      if (it->second.n_host != n_host || it->second.kind != kind)
        return fail(MEEPCU_EINVAL, "host array registered twice with different shapes");
      it->second.mut = it->second.mut || mut;
      it->second.is_field = it->second.is_field || is_field;
      return MEEPCU_OK; // (H aliasing B keeps the B component: same Yee position)
    }
    ArrayRec r = {host, 0, chunk, kind, n_host, n_dev, mut, is_field, false, group, comp};
    arrays[key] = r;
    return MEEPCU_OK;
  }
  const ArrayRec *find(const void *p, size_t esz, size_t *elem) const {
    if (!p) return 0;
    std::map<const char *, ArrayRec>::const_iterator it = arrays.upper_bound((const char *)p);
    if (it == arrays.begin()) return 0;
    --it;
    const size_t off = (const char *)p - it->first;
    const size_t e = it->second.kind == A_DFT ? 2 * esz : esz;
    if (off >= it->second.n_host * e) return 0;
    if (elem) *elem = off / esz;
    return &it->second;
  }
  T *dev(const void *host) const { // device base of a registered array (NULL stays NULL)
    if (!host) return 0;
    std::map<const char *, ArrayRec>::const_iterator it = arrays.find((const char *)host);
    return it == arrays.end() ? 0 : (T *)it->second.dev;
  }
  // device address of one element given a host pointer into a registered GRID array
  T *dev_elem(const void *p) const {
    size_t e = 0;
    const ArrayRec *r = find(p, sizeof(T), &e);
    if (!r) return 0;
    if (r->kind == A_GRID) return (T *)r->dev + chunks[r->chunk].g.dev_index((long long)e);
    return (T *)r->dev + e;
  }

  // device address of one element as the STEPPING kernels see it: an element of an E array that is not stored
  // while stepping is the same element of the D array (*mul: the chi1inv element a reader multiplies it by)
  T *dev_elem_v(const void *p, int ft, const T **mul) const {
    *mul = 0;
    size_t e = 0;
    const ArrayRec *r = find(p, sizeof(T), &e);
    if (!r) return 0;
    if (ft == MEEPCU_E_STUFF && r->kind == A_GRID && !virt_alias.empty()) {
      std::map<const char *, const char *>::const_iterator it = virt_alias.find((const char *)r->host);
      if (it != virt_alias.end()) {
        const long long di = chunks[r->chunk].g.dev_index((long long)e);
        std::map<const char *, const char *>::const_iterator mu = virt_mul.find((const char *)r->host);
        if (mu != virt_mul.end()) *mul = dev(mu->second) + di;
        return dev(it->second) + di;
      }
    }
    if (r->kind == A_GRID) return (T *)r->dev + chunks[r->chunk].g.dev_index((long long)e);
    return (T *)r->dev + e;
  }
  // Is the E update of this chunk a pure function of D (E = chi1inv_diag * (D - dipoles)), and does nobody but the
  // B/H curl kernels and the halo copies read its E arrays?  (3-D, no PML / conductivity / polarisation / BFAST,
  // no off-diagonal or nonlinear update, no metal plane, no DFT of E or D, dipoles away from the chunk's faces)
  bool virt_eligible(size_t ic, const std::map<const char *, unsigned> &zmask) const {
    const ChunkH &ch = chunks[ic];
    const Geom &g = ch.g;
    if (cfg.dim != MEEPCU_D3 || g.ntot_dev >= ((size_t)1 << 31)) return false;
    if (cfg.bfast_scaled_k[0] || cfg.bfast_scaled_k[1] || cfg.bfast_scaled_k[2]) return false;
    for (int d = 0; d < 3; ++d)
      if (!g.has[d] || ch.d.sigsize[d] > 1 || g.num[d] < 6) return false;
    if (!ch.pol[0].empty() || !ch.pol[1].empty()) return false;
    const int ncmp = cfg.is_real ? 1 : 2;
    for (int k = 0; k < 3; ++k) {
      const int ec = dir_comp(MEEPCU_E_STUFF, k), dc = dir_comp(MEEPCU_D_STUFF, k), hc = dir_comp(MEEPCU_H_STUFF, k),
                bc = dir_comp(MEEPCU_B_STUFF, k);
      const int cs[4] = {ec, dc, hc, bc};
      for (int q = 0; q < 4; ++q)
        for (int d = 0; d < 5; ++d)
          if (ch.d.conductivity[cs[q]][d]) return false;
      for (int cmp = 0; cmp < ncmp; ++cmp) {
        if (!ch.d.f[ec][cmp] || !ch.d.f[dc][cmp] || !ch.d.f[hc][cmp] || !ch.d.f[bc][cmp] || ch.d.f[ec][cmp] == ch.d.f[dc][cmp])
          return false;
        if (zmask.count((const char *)ch.d.f[ec][cmp]) || zmask.count((const char *)ch.d.f[dc][cmp])) return false;
      }
      const int d1 = (k + 1) % 3, d2 = (k + 2) % 3;
      const void *u1 = ch.d.chi1inv[ec][d1], *u2 = ch.d.chi1inv[ec][d2];
      if (cfg.nl_mode == MEEPCU_NL_UPSTREAM_PADE) {
        if (u1 || u2 || ch.d.chi3[ec]) return false;
      }
      else if (u1 && u2 && ch.d.chi3[ec])
        return false;
      for (int half = 0; half < 2; ++half) { // the hard-wired step plan of the interior kernels (k_curl3d)
        const int cc = dir_comp(half ? MEEPCU_D_STUFF : MEEPCU_B_STUFF, k), ftg = half ? MEEPCU_H_STUFF : MEEPCU_E_STUFF;
        if (!ch.have_plus[cc] || !ch.have_minus[cc] || ch.plus_c[cc] != dir_comp(ftg, (k + 2) % 3) || ch.plus_d[cc] != (k + 1) % 3 ||
            ch.minus_c[cc] != dir_comp(ftg, (k + 1) % 3) || ch.minus_d[cc] != (k + 2) % 3)
          return false;
      }
    }
    for (size_t id = 0; id < ch.dft.size(); ++id)
      if (is_electric(ch.dft[id].d.c) || is_D(ch.dft[id].d.c)) return false;
    for (size_t is = 0; is < ch.src.size(); ++is) {
      const SrcH &sr = ch.src[is];
      if (sr.ft != MEEPCU_D_STUFF || !sr.integrated) continue;
      for (size_t j = 0; j < sr.idx.size(); ++j) {
        long long rem = sr.idx[j];
        for (int d = 0; d < 3; ++d) {
          const long long q = rem / g.hs[d];
          rem -= q * g.hs[d];
          if (q < 2 || q > g.num[d] - 2) return false;
        }
      }
    }
    return true;
  }
  struct EhMask { unsigned char *mask; const T *u, *u1, *u2, *chi2, *chi3; int mode; long long n; };
  std::vector<EhMask> eh_masks;
  int build_eh_masks() { // after every upload of the materials
    for (size_t k = 0; k < eh_masks.size(); ++k) {
      const EhMask &m = eh_masks[k];
      k_eh_mask<T><<<148 * 8, kThreads, 0, st>>>(m.mask, m.u, m.u1, m.u2, m.chi2, m.chi3, m.mode, m.n);
      launches++;
    }
    return MEEPCU_OK;
  }
  int virt_patch_u() { // after every upload of the materials
    for (size_t k = 0; k < virt_comps.size(); ++k) {
      const VirtComp &vc = virt_comps[k];
      if (!vc.u || vc.cmp != 0) continue;
      const Geom &g = chunks[vc.chunk].g;
      int lo[3], hi[3];
      for (int ax = 0; ax < 3; ++ax) {
        const int d = g.dir_of_axis[ax];
        lo[ax] = 1 - g.yee_shift(vc.comp, d);
        hi[ax] = g.num[d] - g.yee_shift(vc.comp, d);
      }
      k_virt_patch_u<T><<<148 * 8, kThreads, 0, st>>>(vc.u, g.nax[0], g.nax[1], g.pitch, lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]);
      launches++;
    }
    return MEEPCU_OK;
  }
  // E arrays of the virtual chunks := what the D/E half would have stored (before the host reads them)
  int virt_materialise() {
    char *p_prev = 0;
    RC(ring_put(last_row.data(), last_row.size() * sizeof(double), &p_prev));
    for (size_t k = 0; k < virt_comps.size(); ++k) {
      const VirtComp &vc = virt_comps[k];
      k_virt_materialise<T><<<148 * 8, kThreads, 0, st>>>(vc.e, vc.d, vc.u, vc.n);
      launches++;
      if (vc.nsrc) {
        k_virt_src<T><<<1, 32, 0, st>>>(vc.e, vc.d, vc.u, vc.src, vc.nsrc, vc.cmp, (const double *)p_prev);
        launches++;
      }
    }
    return MEEPCU_OK;
  }

  // Host<->device copies of big PAGEABLE arrays go through two library-owned pinned staging buffers filled /
  // drained by OpenMP memcpy, so that sync runs near PCIe speed.  With meepcu_config::pin_host the caller's big
  // mutable arrays are instead page-locked IN PLACE at finalize (cudaHostRegister) and copied by plain DMA: the
  // caller then promises that those arrays outlive the context (they are unregistered in ~Ctx); arrays that
  // cannot be registered (e.g. already registered by the caller) silently take the staging path.
  enum { kStageBytes = 64 << 20 };
  char *stage[2];
  cudaEvent_t stage_ev[2];
  int stage_next;
  int stage_init() {
    if (stage[0]) return MEEPCU_OK;
    for (int k = 0; k < 2; ++k) {
      CU(cudaHostAlloc((void **)&stage[k], kStageBytes, cudaHostAllocDefault));
      CU(cudaEventCreateWithFlags(&stage_ev[k], cudaEventDisableTiming));
    }
    stage_next = 0;
    return MEEPCU_OK;
  }
  static void par_memcpy(char *dst, const char *src, size_t n) {
    const size_t blk = 1 << 20;
    const long nb = (long)((n + blk - 1) / blk);
#pragma omp parallel for schedule(static)
    for (long b = 0; b < nb; ++b) {
      const size_t o = (size_t)b * blk;
      memcpy(dst + o, src + o, std::min(blk, n - o));
    }
  }
  // Device-side staging: a dense block of rows travels by 1-D DMA and k_rows moves it to / from the
  // padded device layout (all on stream st, so one buffer is enough).
  char *dstage;
  int dstage_init() {
    if (dstage) return MEEPCU_OK;
    CU(cudaMalloc((void **)&dstage, kStageBytes));
    return MEEPCU_OK;
  }
  int rows_launch(T *dst, long long dp, const T *src, long long sp, int w, long long rows) {
    const long long n = rows * w;
    const int grid = (int)std::min<long long>((n + kThreads - 1) / kThreads, 148 * 16);
    k_rows<T><<<grid, kThreads, 0, st>>>(dst, dp, src, sp, w, rows);
    launches++;
    return MEEPCU_OK;
  }
  // rows [r0, r0+nr) of a grid array <-> a dense page-locked host block, asynchronously on st
  int rows_copy(const ArrayRec &r, const Geom &g, size_t r0, size_t nr, char *hblock, bool to_host) {
    const size_t w = (size_t)g.nax[2] * sizeof(T);
    T *d = (T *)r.dev + r0 * (size_t)g.pitch;
    if (g.pitch == g.nax[2]) {
      if (to_host) CU(cudaMemcpyAsync(hblock, d, nr * w, cudaMemcpyDeviceToHost, st));
      else CU(cudaMemcpyAsync(d, hblock, nr * w, cudaMemcpyHostToDevice, st));
      return MEEPCU_OK;
    }
    RC(dstage_init());
    if (to_host) {
      RC(rows_launch((T *)dstage, g.nax[2], d, g.pitch, g.nax[2], (long long)nr));
      CU(cudaMemcpyAsync(hblock, dstage, nr * w, cudaMemcpyDeviceToHost, st));
    }
    else {
      CU(cudaMemcpyAsync(dstage, hblock, nr * w, cudaMemcpyHostToDevice, st));
      RC(rows_launch(d, g.pitch, (const T *)dstage, g.nax[2], g.nax[2], (long long)nr));
    }
    return MEEPCU_OK;
  }

  // one dense PAGEABLE host array <-> device array, in blocks that fit a pinned staging buffer
  int staged_copy(const ArrayRec &r, bool to_host) {
    RC(stage_init());
    size_t w, rows;
    const Geom *g = 0;
    if (r.kind == A_GRID) {
      g = &chunks[r.chunk].g;
      w = (size_t)g->nax[2] * sizeof(T);
      rows = (size_t)g->nax[0] * g->nax[1];
    }
    else {
      const size_t nb = r.n_host * sizeof(T) * (r.kind == A_DFT ? 2 : 1);
      w = std::min<size_t>(nb, 1 << 20);
      rows = (nb + w - 1) / w;
      if (rows * w != nb) { // odd size: single pageable copy
        if (to_host) CU(cudaMemcpyAsync((void *)r.host, r.dev, nb, cudaMemcpyDeviceToHost, st));
        else CU(cudaMemcpyAsync(r.dev, r.host, nb, cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));
        return MEEPCU_OK;
      }
    }
    const size_t rows_per = std::max<size_t>(1, (size_t)kStageBytes / w);
    struct Pending { int buf; size_t r0, nr; } pend[2];
    int npend = 0;
    for (size_t r0 = 0; r0 < rows; r0 += rows_per) {
      const size_t nr = std::min(rows_per, rows - r0);
      const int b = stage_next;
      stage_next ^= 1;
      if (to_host) {
        // drain the buffer we are about to reuse
        for (int k = 0; k < npend; ++k)
          if (pend[k].buf == b) {
            CU(cudaEventSynchronize(stage_ev[b]));
            par_memcpy((char *)r.host + pend[k].r0 * w, stage[b], pend[k].nr * w);
            pend[k] = pend[npend - 1];
            --npend;
            break;
          }
        if (g) RC(rows_copy(r, *g, r0, nr, stage[b], true));
        else CU(cudaMemcpyAsync(stage[b], r.dev + r0 * w, nr * w, cudaMemcpyDeviceToHost, st));
        CU(cudaEventRecord(stage_ev[b], st));
        pend[npend].buf = b; pend[npend].r0 = r0; pend[npend].nr = nr; ++npend;
      }
      else {
        CU(cudaEventSynchronize(stage_ev[b])); // previous DMA out of this buffer finished
        par_memcpy(stage[b], (const char *)r.host + r0 * w, nr * w);
        if (g) RC(rows_copy(r, *g, r0, nr, stage[b], false));
        else CU(cudaMemcpyAsync(r.dev + r0 * w, stage[b], nr * w, cudaMemcpyHostToDevice, st));
        CU(cudaEventRecord(stage_ev[b], st));
      }
    }
    if (to_host)
      for (int k = 0; k < npend; ++k) {
        CU(cudaEventSynchronize(stage_ev[pend[k].buf]));
        par_memcpy((char *)r.host + pend[k].r0 * w, stage[pend[k].buf], pend[k].nr * w);
      }
    return MEEPCU_OK;
  }

  // Small arrays that are not page-locked in place (pageable copies of 2 KB rows ran at 5-7 GB/s, and a 3-D cell with
  // PML on every face has ~500 of them, ~100 MB): packed by the host cores into the pinned staging buffers, ONE DMA
  // per 64 MB block, ONE k_rows_batch launch per block.
  bool batchable(const ArrayRec &r) const {
    const size_t nb = r.n_host * sizeof(T) * (r.kind == A_DFT ? 2 : 1);
    return !r.pinned && nb <= (size_t)kStageBytes / 8 && nb > 0 && !getenv("MEEPCU_NO_BATCH");
  }
  int batch_copy(const std::vector<const ArrayRec *> &list, bool to_host) {
    if (list.empty()) return MEEPCU_OK;
    RC(stage_init());
    RC(dstage_init());
    struct Item { const ArrayRec *r; size_t off, nb; };
    struct Pending { int buf; std::vector<Item> items; };
    std::vector<Pending> pend;
    auto drain = [&](Pending &P) -> int { // D2H: the block has landed in stage[buf]: hand the rows to their arrays
      CU(cudaEventSynchronize(stage_ev[P.buf]));
      const long ni = (long)P.items.size();
#pragma omp parallel for schedule(dynamic, 4)
      for (long k = 0; k < ni; ++k) memcpy((void *)P.items[k].r->host, stage[P.buf] + P.items[k].off, P.items[k].nb);
      return MEEPCU_OK;
    };
    size_t i = 0;
    while (i < list.size()) {
      const int b = stage_next;
      stage_next ^= 1;
      for (size_t k = 0; k < pend.size(); ++k) // the buffer we are about to reuse
        if (pend[k].buf == b) { RC(drain(pend[k])); pend.erase(pend.begin() + k); break; }
      if (!to_host) CU(cudaEventSynchronize(stage_ev[b])); // the previous DMA out of this buffer has finished
      Pending P;
      P.buf = b;
      std::vector<RowsEntry<T> > tab;
      size_t used = 0;
      for (; i < list.size(); ++i) {
        const ArrayRec &r = *list[i];
        const size_t esz = sizeof(T), ne = r.n_host * (r.kind == A_DFT ? 2 : 1), nb = ne * esz;
        if (used + nb > (size_t)kStageBytes) break;
        RowsEntry<T> e;
        e.dev = (T *)r.dev;
        e.off = (long long)(used / esz);
        if (r.kind == A_GRID) {
          const Geom &g = chunks[r.chunk].g;
          e.w = g.nax[2];
          e.rows = (long long)g.nax[0] * g.nax[1];
          e.dpitch = g.pitch;
        }
        else {
          e.w = (int)std::min<size_t>(ne, (size_t)1 << 20);
          e.rows = (long long)((ne + e.w - 1) / e.w);
          e.dpitch = e.w;
          if ((size_t)e.rows * e.w != ne) { e.w = 1; e.rows = (long long)ne; e.dpitch = 1; }
        }
        tab.push_back(e);
        Item it = {&r, used, nb};
        P.items.push_back(it);
        used += (nb + 15) & ~(size_t)15;
      }
      if (tab.empty()) return fail(MEEPCU_EINVAL, "array too large for the batched path");
      char *d_tab = 0;
      RC(ring_put(tab.data(), tab.size() * sizeof(RowsEntry<T>), &d_tab));
      dim3 grid(8, (unsigned)std::min<size_t>(tab.size(), 148 * 2));
      if (to_host) {
        k_rows_batch<T><<<grid, kThreads, 0, st>>>((const RowsEntry<T> *)d_tab, (int)tab.size(), (T *)dstage, 1);
        launches++;
        CU(cudaMemcpyAsync(stage[b], dstage, used, cudaMemcpyDeviceToHost, st));
        CU(cudaEventRecord(stage_ev[b], st));
        pend.push_back(P);
      }
      else {
        const long ni = (long)P.items.size();
#pragma omp parallel for schedule(dynamic, 4)
        for (long k = 0; k < ni; ++k) memcpy(stage[b] + P.items[k].off, P.items[k].r->host, P.items[k].nb);
        CU(cudaMemcpyAsync(dstage, stage[b], used, cudaMemcpyHostToDevice, st));
        k_rows_batch<T><<<grid, kThreads, 0, st>>>((const RowsEntry<T> *)d_tab, (int)tab.size(), (T *)dstage, 0);
        launches++;
        CU(cudaEventRecord(stage_ev[b], st));
      }
    }
    for (size_t k = 0; k < pend.size(); ++k) RC(drain(pend[k]));
    return MEEPCU_OK;
  }

  // asynchronous copy on st: small arrays (pageable is fine) and big arrays page-locked in place
  int copy_array(const ArrayRec &r, bool to_host) {
    if (r.kind == A_GRID) {
      const Geom &g = chunks[r.chunk].g;
      const size_t w = (size_t)g.nax[2] * sizeof(T), rows = (size_t)g.nax[0] * g.nax[1];
      if (r.pinned) {
        const size_t rows_per = std::max<size_t>(1, (size_t)kStageBytes / w);
        for (size_t r0 = 0; r0 < rows; r0 += rows_per)
          RC(rows_copy(r, g, r0, std::min(rows_per, rows - r0), (char *)r.host + r0 * w, to_host));
        return MEEPCU_OK;
      }
      if (to_host)
        CU(cudaMemcpy2DAsync((void *)r.host, w, r.dev, (size_t)g.pitch * sizeof(T), w, rows,
                             cudaMemcpyDeviceToHost, st));
      else
        CU(cudaMemcpy2DAsync(r.dev, (size_t)g.pitch * sizeof(T), r.host, w, w, rows,
                             cudaMemcpyHostToDevice, st));
    }
    else {
      const size_t nb = r.n_host * sizeof(T) * (r.kind == A_DFT ? 2 : 1);
      if (to_host)
        CU(cudaMemcpyAsync((void *)r.host, r.dev, nb, cudaMemcpyDeviceToHost, st));
      else
        CU(cudaMemcpyAsync(r.dev, r.host, nb, cudaMemcpyHostToDevice, st));
    }
    return MEEPCU_OK;
  }

  template <class U> int upload(const std::vector<U> &v, U **out) {
    *out = 0;
    if (v.empty()) return MEEPCU_OK;
    CU(cudaMalloc((void **)out, v.size() * sizeof(U)));
    extra_dev.push_back((char *)*out);
    CU(cudaMemcpy(*out, v.data(), v.size() * sizeof(U), cudaMemcpyHostToDevice));
    return MEEPCU_OK;
  }

  // metal lists -> per-(array) plane masks; returns EUNSUPPORTED if a list is not a union of
  // boundary planes of the owned region (it always is: boundaries.cpp:309-339 lists the owned
  // points with IVEC_LOOP_AT_BOUNDARY that on_metal_boundary() accepts, i.e. whole faces)
  int metal_masks(std::map<const char *, unsigned> &mask) {
    for (int ft = 0; ft < MEEPCU_NUM_FIELD_TYPES; ++ft) {
      if (metal[ft].empty()) continue;
      // group by array
      std::map<const char *, std::vector<long long> > by;
      for (size_t k = 0; k < metal[ft].size(); ++k) {
        size_t e = 0;
        const ArrayRec *r = find(metal[ft][k], sizeof(T), &e);
        if (!r || r->kind != A_GRID || !r->is_field)
          return fail(MEEPCU_EINVAL, "metal pointer %zu of field type %d is not inside a field array", k, ft);
        by[(const char *)r->host].push_back((long long)e);
      }
      for (std::map<const char *, std::vector<long long> >::iterator it = by.begin(); it != by.end(); ++it) {
        std::sort(it->second.begin(), it->second.end());
        it->second.erase(std::unique(it->second.begin(), it->second.end()), it->second.end());
        const ArrayRec &r = arrays[it->first];
        const ChunkH &ch = chunks[r.chunk];
        const Geom &g = ch.g;
        // which component is this array?  (needed for the owned range)
        int comp = -1;
        for (int c = 0; c < 20 && comp < 0; ++c)
          for (int cmp = 0; cmp < 2; ++cmp)
            if (ch.d.f[c][cmp] == r.host && comp_ft(c) == ft) comp = c;
        if (comp < 0) return fail(MEEPCU_EINVAL, "metal list of field type %d refers to a foreign array", ft);
        int lo[3], hi[3]; // owned range per direction (host index units)
        for (int d = 0; d < 3; ++d) {
          lo[d] = g.has[d] ? 1 - g.yee_shift(comp, d) : 0;
          hi[d] = g.has[d] ? g.num[d] - g.yee_shift(comp, d) : 0;
        }
        long long cnt[6] = {0, 0, 0, 0, 0, 0};
        std::vector<int> xyz(3 * it->second.size());
        std::vector<char> outside(it->second.size(), 0);
        for (size_t k = 0; k < it->second.size(); ++k) {
          long long rem = it->second[k];
          for (int d = 0; d < 3; ++d) {
            int q = 0;
            if (g.has[d]) { q = (int)(rem / g.hs[d]); rem -= q * g.hs[d]; }
            xyz[3 * k + d] = q;
            if (g.has[d] && (q < lo[d] || q > hi[d])) outside[k] = 1;
          }
          // Cylindrical cells: find_metals loops over LOOP_OVER_VOL_OWNED, whose range includes the r = 0 row of the
          // components that live on the axis (vec.cpp:432-436) -- outside the range the stepping kernels own.  Those
          // points are only ever written by k_cyl_origin; they are cleared by the late k_zero list of the half step.
          if (outside[k]) {
            if (cfg.dim != MEEPCU_Dcyl) return fail(MEEPCU_EUNSUPPORTED, "metal point outside the owned range");
            metal_axis[ft].push_back((T *)r.dev + g.dev_index(it->second[k]));
            continue;
          }
          for (int d = 0; d < 3; ++d) {
            if (g.has[d] && xyz[3 * k + d] == lo[d]) cnt[d]++;
            if (g.has[d] && xyz[3 * k + d] == hi[d]) cnt[3 + d]++;
          }
        }
        unsigned m = 0;
        for (int d = 0; d < 3; ++d) {
          if (!g.has[d]) continue;
          long long face = 1;
          for (int e = 0; e < 3; ++e)
            if (e != d && g.has[e]) face *= (hi[e] - lo[e] + 1);
          const int ax = g.axis_of_dir[d];
          if (cnt[d] == face && lo[d] != hi[d]) m |= 1u << ax;
          if (cnt[3 + d] == face) m |= 1u << (3 + ax);
        }
        for (size_t k = 0; k < it->second.size(); ++k) {
          bool covered = outside[k] != 0;
          for (int d = 0; d < 3; ++d) {
            if (!g.has[d]) continue;
            const int ax = g.axis_of_dir[d];
            if ((m & (1u << ax)) && xyz[3 * k + d] == lo[d]) covered = true;
            if ((m & (1u << (3 + ax))) && xyz[3 * k + d] == hi[d]) covered = true;
          }
          if (!covered)
            return fail(MEEPCU_EUNSUPPORTED, "metal point list is not a union of boundary faces");
        }
        mask[it->first] = m;
      }
    }
    return MEEPCU_OK;
  }

  // ------------------------------------------------------------------------------------------
  int finalize() {
    if (finalized) return fail(MEEPCU_ESTATE, "context already finalized");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
      return fail(MEEPCU_ECUDA, "no CUDA device: libmeepcuda has no CPU fallback");
    CU(cudaSetDevice(cfg.device));
    CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    {
      int lo_p = 0, hi_p = 0;
      CU(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
      const int pr = getenv("MEEPCU_ST2_PRIO") ? atoi(getenv("MEEPCU_ST2_PRIO")) : 1; // 1: high (default), -1: low
      CU(cudaStreamCreateWithPriority(&st2, cudaStreamNonBlocking, pr > 0 ? hi_p : (pr < 0 ? lo_p : 0)));
      CU(cudaStreamCreateWithPriority(&st_comm, cudaStreamNonBlocking, hi_p));
      CU(cudaStreamCreateWithPriority(&st_bnd, cudaStreamNonBlocking, hi_p));
    }
    CU(cudaEventCreateWithFlags(&ev_bdone, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ev_cdone, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ev_bstart, cudaEventDisableTiming));
    CU(cudaEventCreate(&ev0));
    CU(cudaEventCreate(&ev1));
    CU(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    if (cfg.dim != MEEPCU_D2 && cfg.beta != 0)
      return fail(MEEPCU_EINVAL, "Nonzero beta unsupported in dimensions other than 2."); // fields.cpp:559-560

    const bool tm_on = getenv("MEEPCU_TIMING") != 0;
    double tm_last = omp_get_wtime();
    auto tm_mark = [&](const char *what) {
      if (!tm_on) return;
      cudaStreamSynchronize(st);
      const double now = omp_get_wtime();
      fprintf(stderr, "[meepcu_finalize] %-28s %8.3f s\n", what, now - tm_last);
      tm_last = now;
    };
    // ---- 1. register every array ----
    for (size_t ic = 0; ic < chunks.size(); ++ic) {
      ChunkH &ch = chunks[ic];
      const Geom &g = ch.g;
      const size_t nh = g.ntot_host, nd = g.ntot_dev;
      for (int c = 0; c < 20; ++c) {
        for (int cmp = 0; cmp < 2; ++cmp) {
          RC(reg(ch.d.f[c][cmp], (int)ic, A_GRID, nh, nd, true, true, MEEPCU_SYNC_FIELDS, c));
          RC(reg(ch.d.f_u[c][cmp], (int)ic, A_GRID, nh, nd, true, false, MEEPCU_SYNC_AUX));
          RC(reg(ch.d.f_w[c][cmp], (int)ic, A_GRID, nh, nd, true, false, MEEPCU_SYNC_AUX));
          RC(reg(ch.d.f_cond[c][cmp], (int)ic, A_GRID, nh, nd, true, false, MEEPCU_SYNC_AUX));
          RC(reg(ch.d.f_bfast[c][cmp], (int)ic, A_GRID, nh, nd, true, false, MEEPCU_SYNC_AUX));
        }
        for (int d = 0; d < 5; ++d) {
          RC(reg(ch.d.chi1inv[c][d], (int)ic, A_GRID, nh, nd, false));
          RC(reg(ch.d.conductivity[c][d], (int)ic, A_GRID, nh, nd, false));
          RC(reg(ch.d.condinv[c][d], (int)ic, A_GRID, nh, nd, false));
        }
        RC(reg(ch.d.chi2[c], (int)ic, A_GRID, nh, nd, false));
        RC(reg(ch.d.chi3[c], (int)ic, A_GRID, nh, nd, false));
      }
      for (int d = 0; d < 3; ++d)
        if (ch.d.sigsize[d] > 1) {
          if (!ch.d.sig[d] || !ch.d.kap[d] || !ch.d.siginv[d])
            return fail(MEEPCU_EINVAL, "sigsize > 1 but sig/kap/siginv missing");
          RC(reg(ch.d.sig[d], (int)ic, A_VEC, ch.d.sigsize[d], ch.d.sigsize[d], false));
          RC(reg(ch.d.kap[d], (int)ic, A_VEC, ch.d.sigsize[d], ch.d.sigsize[d], false));
          RC(reg(ch.d.siginv[d], (int)ic, A_VEC, ch.d.sigsize[d], ch.d.sigsize[d], false));
        }
      for (int k = 0; k < 2; ++k)
        for (size_t ip = 0; ip < ch.pol[k].size(); ++ip) {
          const meepcu_susceptibility_desc &sd = ch.pol[k][ip].d;
          for (int c = 0; c < 20; ++c) {
            for (int d = 0; d < 5; ++d) RC(reg(sd.sigma[c][d], (int)ic, A_GRID, nh, nd, false));
            for (int cmp = 0; cmp < 2; ++cmp) {
              if ((sd.P[c][cmp] != 0) != (sd.P_prev[c][cmp] != 0))
                return fail(MEEPCU_EINVAL, "P and P_prev must both be given");
              RC(reg(sd.P[c][cmp], (int)ic, A_GRID, nh, nd, true, false, MEEPCU_SYNC_POL));
              RC(reg(sd.P_prev[c][cmp], (int)ic, A_GRID, nh, nd, true, false, MEEPCU_SYNC_POL));
              if (ch.pol[k][ip].gyro)
                for (int dd = 0; dd < 3; ++dd) {
                  RC(reg(ch.pol[k][ip].g.P[c][cmp][dd], (int)ic, A_GRID, nh, nd, true, false, MEEPCU_SYNC_POL));
                  RC(reg(ch.pol[k][ip].g.P_prev[c][cmp][dd], (int)ic, A_GRID, nh, nd, true, false, MEEPCU_SYNC_POL));
                }
            }
          }
        }
      for (size_t im = 0; im < ch.ml.size(); ++im) { // multilevel atoms: sigma, the populations (host layout, not pitched)
        for (int c = 0; c < 20; ++c)
          for (int d = 0; d < 5; ++d) RC(reg(ch.ml[im].sigma[c][d], (int)ic, A_GRID, nh, nd, false));
        RC(reg(ch.ml[im].N, (int)ic, A_VEC, nh * ch.ml[im].L, nh * ch.ml[im].L, true, false, MEEPCU_SYNC_POL));
      }
      for (int c = 0; c < 20; ++c)
        for (int cmp = 0; cmp < 2; ++cmp) RC(reg(ch.w_prev[c][cmp], (int)ic, A_GRID, nh, nd, true, false, MEEPCU_SYNC_AUX));
      for (size_t id = 0; id < ch.dft.size(); ++id) {
        const size_t n = ch.dft[id].N * ch.dft[id].d.Nomega;
        RC(reg(ch.dft[id].d.dft, (int)ic, A_DFT, n, n, true, false, MEEPCU_SYNC_DFT));
      }
    }
    // ---- 2. one arena ----
    size_t total = 0;
    for (std::map<const char *, ArrayRec>::iterator it = arrays.begin(); it != arrays.end(); ++it) {
      const size_t nb = it->second.n_dev * sizeof(T) * (it->second.kind == A_DFT ? 2 : 1);
      it->second.dev = (char *)total; // offset for now
      total += (nb + 255) & ~(size_t)255;
    }
    arena_bytes = total + 256;
    CU(cudaMalloc((void **)&arena, arena_bytes));
    CU(cudaMemsetAsync(arena, 0, arena_bytes, st));
    for (std::map<const char *, ArrayRec>::iterator it = arrays.begin(); it != arrays.end(); ++it)
      it->second.dev = arena + (size_t)it->second.dev;
    CU(cudaMalloc((void **)&d_flags, 8 * sizeof(int)));
    extra_dev.push_back((char *)d_flags);
    CU(cudaMemsetAsync(d_flags, 0, 8 * sizeof(int), st));
    if (cfg.pin_host && !getenv("MEEPCU_NO_PIN"))
      for (std::map<const char *, ArrayRec>::iterator it = arrays.begin(); it != arrays.end(); ++it) {
        ArrayRec &r = it->second;
        if (!big(r)) continue;
        const size_t nb = r.n_host * sizeof(T) * (r.kind == A_DFT ? 2 : 1);
        // constant arrays are uploaded once: only the ones that travel at every sync are worth pinning
        if (!r.mut) continue;
        if (cudaHostRegister((void *)r.host, nb, cudaHostRegisterDefault) == cudaSuccess) r.pinned = true;
        else cudaGetLastError(); // not registrable (e.g. already registered by the caller): staging path
      }
    RC(sync_groups(false, ~0u)); // everything, materials included

    tm_mark("register + arena + upload");
    // ---- 3. metal planes ----
    std::map<const char *, unsigned> zmask;
    RC(metal_masks(zmask));

    tm_mark("metal masks");
    // ---- 4. sources -> sorted device tables ----
    std::vector<SrcEntry> all_src;
    struct SrcRange { size_t c0, c1, i0, i1; int ca0, ca1, ia0, ia1; };
    // per chunk, per half (0:B 1:D), per comp k, per cmp-independent
    std::vector<SrcRange> ranges(chunks.size() * 2 * 3);
    for (size_t ic = 0; ic < chunks.size(); ++ic) {
      ChunkH &ch = chunks[ic];
      const Geom &g = ch.g;
      for (int half = 0; half < 2; ++half) {
        const int ft2 = half ? MEEPCU_D_STUFF : MEEPCU_B_STUFF;
        const int ft = half ? MEEPCU_E_STUFF : MEEPCU_H_STUFF;
        for (int k = 0; k < 3; ++k) {
          SrcRange &R = ranges[(ic * 2 + half) * 3 + k];
          for (int pass = 0; pass < 2; ++pass) { // 0: currents, 1: integrated
            std::vector<SrcEntry> tmp;
            for (size_t is = 0; is < ch.src.size(); ++is) {
              const SrcH &s = ch.src[is];
              if (s.ft != ft2 || comp_dir(s.c) != k) continue;
              if (comp_dir(s.c) >= MEEPCU_R) return fail(MEEPCU_EUNSUPPORTED, "cylindrical source");
              bool use;
              if (pass == 0) // step.cpp:301-302
                use = !s.integrated && ch.d.f[dir_comp(ft2, k)][0] &&
                      ((ft2 == MEEPCU_D_STUFF && is_electric(s.c)) ||
                       (ft2 == MEEPCU_B_STUFF && is_magnetic(s.c)));
              else // update_eh.cpp:138
                use = s.integrated && ch.d.f[s.c][0] && ft == comp_ft(s.c);
              if (!use) continue;
              for (size_t j = 0; j < s.idx.size(); ++j) {
                if (s.idx[j] < 0 || (size_t)s.idx[j] >= g.ntot_host)
                  return fail(MEEPCU_EINVAL, "source index out of range");
                SrcEntry e;
                e.idx = g.dev_index(s.idx[j]);
                e.are = s.amp[2 * j];
                e.aim = s.amp[2 * j + 1];
                e.tid = s.tid;
                e.pad = 0;
                tmp.push_back(e);
              }
            }
            std::stable_sort(tmp.begin(), tmp.end(),
                             [](const SrcEntry &a, const SrcEntry &b) { return a.idx < b.idx; });
            int amin = 1 << 30, amax = -1;
            const long long sA = (long long)g.pitch * g.nax[1];
            for (size_t j = 0; j < tmp.size(); ++j) {
              const int ia = (int)(tmp[j].idx / sA);
              amin = std::min(amin, ia);
              amax = std::max(amax, ia);
            }
            if (pass == 0) { R.c0 = all_src.size(); R.c1 = R.c0 + tmp.size(); R.ca0 = amin; R.ca1 = amax; }
            else { R.i0 = all_src.size(); R.i1 = R.i0 + tmp.size(); R.ia0 = amin; R.ia1 = amax; }
            all_src.insert(all_src.end(), tmp.begin(), tmp.end());
          }
        }
      }
    }
    RC(upload(all_src, &d_src));
    h_src_idx.resize(all_src.size());
    for (size_t k = 0; k < all_src.size(); ++k) h_src_idx[k] = all_src[k].idx;

    // ---- 4b. chunks whose E arrays are not stored while stepping (Curl3DParams::virt) ----
    if (!getenv("MEEPCU_NO_VIRTUAL_E"))
      for (size_t ic = 0; ic < chunks.size(); ++ic)
        if (virt_eligible(ic, zmask)) virt_chunks.insert((int)ic);

    tm_mark("sources, virtual chunks");
    // ---- 5. per-(chunk, half, cmp) parameter blocks + tiles ----
    std::vector<Tile> tiles[2][4], eh_tiles[2], c3d_tiles[2], c3dv_tiles[2], gv_tiles[2][2];
    std::vector<CylRderiv<T> > cyl_rd[2];
    std::vector<CylOrigin<T> > cyl_org[2];
    std::vector<FmpParams<T> > fmps[2];
    std::vector<DipPatch<T> > dips[2];
    std::vector<Tile> polgen_tiles[2], polgyro_tiles[2], mllev_tiles, mlpol_tiles;
    std::vector<CopyItem<T> > wprev_items;
    std::vector<int> c3d_chunk, half_chunk; // owning chunk of every h_c3d / h_half entry
    // interior and PML tiles of a half in one grid (k_mix); fp64 PML tiles are vectorised only for that
    const bool mix_want = getenv("MEEPCU_MIX") ? atoi(getenv("MEEPCU_MIX")) != 0 : true;
    const int ncmp = cfg.is_real ? 1 : 2;
    cells = 0;
    alg_bytes = 0;
    alg_bytes_sv = 0;
    memset(kc_bytes_sv, 0, sizeof kc_bytes_sv);
    for (size_t ic = 0; ic < chunks.size(); ++ic) {
      ChunkH &ch = chunks[ic];
      const Geom &g = ch.g;
      long long nc = 1;
      for (int d = 0; d < 3; ++d)
        if (g.has[d]) nc *= g.num[d];
      cells += nc;
      for (int half = 0; half < 2; ++half) {
        const int ft2 = half ? MEEPCU_D_STUFF : MEEPCU_B_STUFF;
        const int ft = half ? MEEPCU_E_STUFF : MEEPCU_H_STUFF;
        // does update_eh need f_minus_p in this chunk? (update_eh.cpp:71-112)
        bool have_int_sources = false;
        for (size_t is = 0; is < ch.src.size(); ++is)
          if (ch.src[is].ft == ft2 && ch.src[is].integrated) have_int_sources = true;
        // classify the E/H update of this chunk
        bool general = false;
        for (int k = 0; k < 3; ++k) {
          const int ec = dir_comp(ft, k);
          if (!ch.d.f[ec][0]) continue;
          const int d1 = (k + 1) % 3, d2 = (k + 2) % 3;
          const bool pml = ch.d.sigsize[k] > 1;
          const void *u1 = ch.d.f[dir_comp(ft2, d1)][0] ? ch.d.chi1inv[ec][d1] : 0;
          const void *u2 = ch.d.f[dir_comp(ft2, d2)][0] ? ch.d.chi1inv[ec][d2] : 0;
          if (cfg.nl_mode == MEEPCU_NL_UPSTREAM_PADE) {
            if (u1 || u2 || ch.d.chi3[ec]) general = true;
          }
          else if (!pml && u1 && u2 && ch.d.chi3[ec])
            general = true;
          for (size_t ip = 0; ip < ch.pol[ft].size(); ++ip) {
            const meepcu_susceptibility_desc &sd = ch.pol[ft][ip].d;
            if (ch.pol[ft][ip].gyro) aniso_chunk_half.insert(std::make_pair((int)ic, half)); // -> k_pol_gyro
            if (ch.pol[ft][ip].ml >= 0 && sd.P[ec][0]) { // -> k_ml_levels reads W around every centred point before
              aniso_chunk_half.insert(std::make_pair((int)ic, half)); // zero_metal(E), and P halos every step
              eager_p[half] = true;
            }
            for (int cmp = 0; cmp < ncmp; ++cmp)
              if (sd.P[ec][cmp] && sd.sigma[ec][k] &&
                  ((sd.sigma[ec][d1] && ch.d.f[dir_comp(ft, d1)][cmp]) ||
                   (sd.sigma[ec][d2] && ch.d.f[dir_comp(ft, d2)][cmp])))
                aniso_chunk_half.insert(std::make_pair((int)ic, half)); // susceptibility.cpp:225-251 -> k_pol_general
          }
        }
        if (general) {
          general_half[half] = true;
          eager_db[half] = true;
          if (!ch.pol[ft].empty()) eager_p[half] = true;
        }
        aniso_now = aniso_chunk_half.count(std::make_pair((int)ic, half)) != 0;
        for (int cmp = 0; cmp < ncmp; ++cmp) {
          HalfParams<T> P;
          memset(&P, 0, sizeof P);
          P.nA = g.nax[0]; P.nB = g.nax[1]; P.nC = g.nax[2]; P.pitch = g.pitch;
          P.sA = (long long)g.pitch * g.nax[1];
          P.sB = g.pitch;
          P.is_D = half;
          P.cmp = cmp;
          P.dt = cfg.dt;
          bool any = false, simple = true;
          for (int k = 0; k < 3; ++k) {
            CompParams<T> &c = P.comp[k];
            c.kaxis = c.kuaxis = c.kwaxis = -1;
            const int cc = dir_comp(ft2, k), ec = dir_comp(ft, k);
            if (!ch.d.f[cc][cmp]) continue;
            const bool have_p = ch.have_plus[cc], have_m = ch.have_minus[cc];
            const T *f_p = have_p ? dev(ch.d.f[ch.plus_c[cc]][cmp]) : 0;
            const T *f_m = have_m ? dev(ch.d.f[ch.minus_c[cc]][cmp]) : 0;
            if (cfg.dim == MEEPCU_Dcyl && k == 2 && f_p) { // step_db.cpp:93-118: f_p := f_rderiv_int (k_cyl_rderiv, every step)
              if (ch.plus_d[cc] != 0 || f_m) return fail(MEEPCU_ESTATE, "internal: cylindrical step plan");
              T *S = 0;
              CU(cudaMalloc((void **)&S, g.ntot_dev * sizeof(T)));
              extra_dev.push_back((char *)S);
              CU(cudaMemsetAsync(S, 0, g.ntot_dev * sizeof(T), st));
              CylRderiv<T> rd = {S, f_p, g.num[0], g.num[2], g.pitch,
                                 (T)(origin_r_times_a(g) + 0.5 * g.yee_shift(ch.plus_c[cc], 0))};
              cyl_rd[half].push_back(rd);
              f_p = S;
            }
            if (!f_p && !f_m) continue;
            any = true;
            long long sp = have_p ? g.ds[ch.plus_d[cc]] : 0, sm = have_m ? g.ds[ch.minus_d[cc]] : 0;
            if (half) { sp = -sp; sm = -sm; } // step_db.cpp:81-84
            T dtdx = (T)cfg.courant;
            if (!f_p) { std::swap(f_p, f_m); std::swap(sp, sm); dtdx = -dtdx; } // step_generic.cpp:76-80
            c.f = dev(ch.d.f[cc][cmp]);
            c.g1 = f_p; c.g2 = f_m; c.s1 = sp; c.s2 = sm;
            c.dtdx = dtdx;
            c.dt2 = (T)((T)cfg.dt * 0.5);
            for (int ax = 0; ax < 3; ++ax) {
              const int d = g.dir_of_axis[ax];
              c.lo[ax] = d < 0 ? 0 : 1 - g.yee_shift(cc, d);
              c.hi[ax] = d < 0 ? 0 : g.num[d] - g.yee_shift(cc, d);
            }
            const int dsig0 = (k + 1) % 3, dsigu0 = (k + 2) % 3;
            const bool dsig = ch.d.sigsize[dsig0] > 1, dsigu = ch.d.sigsize[dsigu0] > 1;
            if ((dsig && !g.has[dsig0]) || (dsigu && !g.has[dsigu0]))
              return fail(MEEPCU_EINVAL, "PML in a direction the grid does not have");
            c.cnd = dev(ch.d.conductivity[cc][k]);
            c.cndinv = dev(ch.d.condinv[cc][k]);
            if (c.cnd && !c.cndinv) return fail(MEEPCU_EINVAL, "conductivity without condinv (update_condinv)");
            if (dsig) {
              c.kaxis = g.axis_of_dir[dsig0];
              c.koff = g.yee_shift(cc, dsig0);
              c.sig = dev(ch.d.sig[dsig0]); c.kap = dev(ch.d.kap[dsig0]); c.siginv = dev(ch.d.siginv[dsig0]);
              if (c.cnd) {
                c.fcnd = dev(ch.d.f_cond[cc][cmp]);
                if (!c.fcnd) return fail(MEEPCU_ESTATE, "f_cond not allocated (step_db.cpp:67-70)");
              }
            }
            if (dsigu) {
              c.kuaxis = g.axis_of_dir[dsigu0];
              c.kuoff = g.yee_shift(cc, dsigu0);
              c.sigu = dev(ch.d.sig[dsigu0]); c.kapu = dev(ch.d.kap[dsigu0]); c.siginvu = dev(ch.d.siginv[dsigu0]);
              c.fu = dev(ch.d.f_u[cc][cmp]);
              if (!c.fu) return fail(MEEPCU_ESTATE, "f_u not allocated (step_db.cpp:71-75)");
            }
            std::map<const char *, unsigned>::iterator zm = zmask.find((const char *)ch.d.f[cc][cmp]);
            c.zmask_db = zm == zmask.end() ? 0u : zm->second;
            const SrcRange &R = ranges[(ic * 2 + half) * 3 + k];
            c.csrc = d_src + R.c0; c.ncsrc = (int)(R.c1 - R.c0); c.csrc_amin = R.ca0; c.csrc_amax = R.ca1;
            c.src_has_cndinv = c.cndinv != 0;
            if (dsig || dsigu || c.cnd) simple = false;
            if (cfg.bfast_scaled_k[0] || cfg.bfast_scaled_k[1] || cfg.bfast_scaled_k[2]) { // step_db.cpp:129-143
              c.fb = dev(ch.d.f_bfast[cc][cmp]);
              if (!c.fb) return fail(MEEPCU_ESTATE, "f_bfast not allocated (step_db.cpp:76-79)");
              T k1 = have_m ? (T)cfg.bfast_scaled_k[comp_dir(ch.minus_c[cc])] : (T)0; // puts k1 in direction of g2
              T k2 = have_p ? (T)cfg.bfast_scaled_k[comp_dir(ch.plus_c[cc])] : (T)0;  // puts k2 in direction of g1
              if (half) { k1 = -k1; k2 = -k2; }
              if (!have_p || !dev(ch.d.f[ch.plus_c[cc]][cmp])) std::swap(k1, k2); // step_generic.cpp:346-350
              c.bk1 = k1;
              c.bk2 = k2;
              simple = false;
            }
            if (cfg.dim == MEEPCU_D2 && cfg.beta != 0 && k <= 1) { // step_db.cpp:161-175
              const int c_g = dir_comp(half ? MEEPCU_H_STUFF : MEEPCU_E_STUFF, k == 0 ? 1 : 0);
              const void *g_other = ch.d.f[c_g][1 - cmp];
              c.gb = dev(g_other ? g_other : ch.d.f[c_g][cmp]);
              const double pi = 3.141592653589793238462643383276; // src/meep.hpp:53
              c.betadt = (T)(2 * pi * cfg.beta * cfg.dt * (k == 0 ? +1 : -1) *
                             (g_other ? (half ? -1 : +1) * (2 * cmp - 1) : 1));
              if (c.gb) simple = false;
            }
            if (cfg.dim == MEEPCU_Dcyl) simple = false; // one code path for cylindrical cells: the most general scalar kernel
            if (cfg.dim == MEEPCU_Dcyl && cfg.m != 0 && (k == 0 || k == 2)) { // step_db.cpp:177-290
              const int c_g = dir_comp(half ? MEEPCU_H_STUFF : MEEPCU_E_STUFF, k == 0 ? 2 : 0);
              c.gm = dev(ch.d.f[c_g][1 - cmp]);
              if (!c.gm) return fail(MEEPCU_EINVAL, "m != 0 needs complex fields (step_db.cpp:183)");
              c.the_m = (T)(2 * cfg.m * (1 - 2 * cmp) * (1 - 2 * (half == 0)) * (1 - 2 * (k == 0)) * cfg.courant);
              c.r0 = g.io[0] + g.yee_shift(cc, 0);
            }
            if (cfg.dim == MEEPCU_Dcyl && cfg.m != 0 && fabs(cfg.m) != 1 && g.io[0] == 0 &&
                ((half == 1 && k >= 1) || (half == 0 && k == 0))) { // step_db.cpp:393-455: rows cleared next to the axis
              const double rmax = cfg.zero_fields_near_cylorigin ? fabs(cfg.m) - int(origin_r_times_a(g) + 0.5) : 1;
              int zr = 0;
              for (int r = 0; r <= g.num[0] && r < rmax; r++) ++zr;
              c.zrows = zr; // (row 0 itself is not in the owned range of these components: k_cyl_origin clears it)
            }

            // ---- E/H part ----
            if (!ch.d.f[ec][cmp]) continue;
            const bool pml_w = ch.d.sigsize[k] > 1;
            bool need_fmp = have_int_sources;
            for (size_t ip = 0; ip < ch.pol[ft].size(); ++ip)
              if (ch.pol[ft][ip].d.P[ec][cmp]) need_fmp = true;
            if (ch.d.f[ec][cmp] == ch.d.f[cc][cmp]) { // H aliases B
              if (ch.d.chi1inv[ec][k] || need_fmp || pml_w)
                return fail(MEEPCU_ESTATE, "H aliases B where update_eh.cpp:204-209 would split them");
              continue;
            }
            if (general) continue; // done by k_eh_general after the D halo exchange
            c.e = dev(ch.d.f[ec][cmp]);
            c.u = dev(ch.d.chi1inv[ec][k]);
            if (pml_w) {
              c.fw = dev(ch.d.f_w[ec][cmp]);
              if (!c.fw) return fail(MEEPCU_ESTATE, "f_w not allocated (update_eh.cpp:212-216)");
              c.kwaxis = g.axis_of_dir[k];
              c.kwoff = g.yee_shift(ec, k);
              c.sigw = dev(ch.d.sig[k]); c.kapw = dev(ch.d.kap[k]);
              simple = false;
            }
            zm = zmask.find((const char *)ch.d.f[ec][cmp]);
            c.zmask_eh = zm == zmask.end() ? 0u : zm->second;
            if (aniso_now) c.zmask_eh = 0; // zero_metal(E) must follow the separate polarisation pass (step.cpp:109-122)
            c.isrc = d_src + R.i0; c.nisrc = (int)(R.i1 - R.i0); c.isrc_amin = R.ia0; c.isrc_amax = R.ia1;
            c.npol = (int)ch.pol[ft].size();
            if (c.npol) simple = false;
            RC(fill_pols(ch, ft, ec, k, cmp, c.pol));
          }
          if (cfg.dim == MEEPCU_Dcyl && g.io[0] == 0) { // the r = 0 rules of step_db.cpp:292-457 -> k_cyl_origin
            const double m = cfg.m;
            const int ftg = half ? MEEPCU_H_STUFF : MEEPCU_E_STUFF;
            const int cR = dir_comp(ft2, 0), cP = dir_comp(ft2, 1), cZ = dir_comp(ft2, 2);
            CylOrigin<T> O;
            memset(&O, 0, sizeof O);
            O.nzp1 = g.num[2] + 1;
            O.hi = -1;
            auto clear_rows = [&](int c) { // ZERO_Z of f, f_cond, f_u
              const void *arrs[3] = {ch.d.f[c][cmp], ch.d.f_cond[c][cmp], ch.d.f_u[c][cmp]};
              for (int q = 0; q < 3; ++q)
                if (arrs[q] && O.nzero < 6) O.zero[O.nzero++] = dev(arrs[q]);
            };
            int upd = -1; // canonical direction of the component with an explicit row update
            if (m == 0 && half == 1 && ch.d.f[cZ][cmp]) { upd = 2; clear_rows(cP); }
            else if (m == 0 && half == 0 && ch.d.f[cR][cmp]) clear_rows(cR);
            else if (fabs(m) == 1) {
              if (ch.d.f[half ? cP : cR][cmp]) {
                upd = half ? 1 : 0;
                if (half) clear_rows(cZ);
              }
            }
            else if (m != 0) {
              if (half) { clear_rows(cP); clear_rows(cZ); }
              else clear_rows(cR);
            }
            if (upd >= 0) {
              const int cc = dir_comp(ft2, upd);
              const int dsig0 = (upd + 1) % 3, dsigu0 = (upd + 2) % 3;
              const bool hs = ch.d.sigsize[dsig0] > 1, hu = ch.d.sigsize[dsigu0] > 1;
              O.cnd = dev(ch.d.conductivity[cc][upd]);
              O.cndinv = dev(ch.d.condinv[cc][upd]);
              O.fcnd = dev(ch.d.f_cond[cc][cmp]);
              if (hs) { O.sig = dev(ch.d.sig[dsig0]); O.kap = dev(ch.d.kap[dsig0]); O.siginv = dev(ch.d.siginv[dsig0]); }
              if (hu) { O.sigu = dev(ch.d.sig[dsigu0]); O.kapu = dev(ch.d.kap[dsigu0]); O.siginvu = dev(ch.d.siginv[dsigu0]); }
              O.k_z = dsig0 == 2; O.koff = hs ? g.yee_shift(cc, dsig0) : 0;
              O.ku_z = dsigu0 == 2; O.kuoff = hu ? g.yee_shift(cc, dsigu0) : 0;
              T *fmain = dev(ch.d.f[cc][cmp]), *fuarr = dev(ch.d.f_u[cc][cmp]);
              O.fu = (O.siginvu && fuarr) ? fmain : 0; // step_db.cpp:311-312: the roles of f and f_u are swapped
              O.the_f = O.fu ? fuarr : fmain;
              O.lo = 1 - g.yee_shift(cc, 2);
              O.hi = g.num[2] - g.yee_shift(cc, 2);
              O.courant = cfg.courant;
              O.dt2 = (T)(cfg.dt * 0.5);
              if (m == 0) {
                O.m0 = 1;
                O.f_p = dev(ch.d.f[dir_comp(ftg, 1)][cmp]); // H_phi
                if (!O.f_p) O.the_f = 0;
              }
              else {
                O.sd = half ? +1 : -1;
                O.f_m_mult = half ? (T)2 : (T)((1 - 2 * cmp) * m);
                O.f_p = dev(ch.d.f[dir_comp(ftg, half ? 0 : 1)][cmp]); // H_r / E_phi
                const T *fm = half ? dev(ch.d.f[dir_comp(ftg, 2)][cmp]) : dev(ch.d.f[dir_comp(ftg, 2)][1 - cmp]);
                O.f_m = fm ? (half ? fm : fm + g.pitch) : 0; // (B: E_z of the other cmp, row r = 1)
                if (!O.f_p || !O.f_m) return fail(MEEPCU_EINVAL, "|m| = 1 needs complex fields and all components (step_db.cpp:341-386)");
              }
            }
            if (O.the_f || O.nzero) cyl_org[half].push_back(O);
          }
          if (!any) continue;
          // interior fast path: full 3-D chunk, plain curl (+ optional E = u*D), 32-bit indexable
          bool fast = simple && cfg.dim == MEEPCU_D3 && g.ntot_dev < ((size_t)1 << 31);
          const int ftg = half ? MEEPCU_H_STUFF : MEEPCU_E_STUFF;
          for (int k = 0; k < 3 && fast; ++k) {
            const int cc = dir_comp(ft2, k);
            const CompParams<T> &c = P.comp[k];
            if (!c.f || !c.g1 || !c.g2 || !ch.have_plus[cc] || !ch.have_minus[cc] ||
                ch.plus_c[cc] != dir_comp(ftg, (k + 2) % 3) || ch.plus_d[cc] != (k + 1) % 3 ||
                ch.minus_c[cc] != dir_comp(ftg, (k + 1) % 3) || ch.minus_d[cc] != (k + 2) % 3 ||
                c.dtdx != (T)cfg.courant)
              fast = false;
          }
          const int plane = g.nax[1] * g.pitch;
          // (march length, fp32, ms per step at 8 / 16 / 32 / 64 planes, r02q / r02r: eps slab 512^2 x 768 2.76 / 2.84 / 2.91 /
          //  3.02, but the full config 5 (1024^2 x 1536) 22.86 at 8 against 22.20 at 16; vacuum 256^3 0.315 / 0.310 / 0.323 / 0.372)
          int seg = 16;
          if (g.nax[0] <= 24) seg = g.nax[0];
          if (getenv("MEEPCU_SEG")) seg = std::max(1, atoi(getenv("MEEPCU_SEG")));
          const bool virt = virt_chunks.count((int)ic) != 0;
          if (virt && !fast) return fail(MEEPCU_ESTATE, "internal: virtual-E chunk off the interior fast path");
          if (virt && half == 1) // the D/E half of such a chunk stores no E
            for (int k = 0; k < 3; ++k) { P.comp[k].e = 0; P.comp[k].u = 0; P.comp[k].nisrc = 0; P.comp[k].zmask_eh = 0; }
          if (fast) {
            Curl3DParams<T> Q;
            memset(&Q, 0, sizeof Q);
            Q.nA = g.nax[0]; Q.nB = g.nax[1]; Q.nC = g.nax[2]; Q.pitch = g.pitch;
            Q.sA = g.pitch * g.nax[1]; Q.sB = g.pitch;
            Q.cmp = cmp;
            Q.dt = cfg.dt;
            Q.dtdx = (T)cfg.courant;
            for (int k = 0; k < 3; ++k) {
              const CompParams<T> &c = P.comp[k];
              Q.f[k] = c.f;
              Q.g[k] = dev(ch.d.f[dir_comp(ftg, k)][cmp]);
              Q.e[k] = c.e;
              Q.u[k] = c.u;
              for (int ax = 0; ax < 3; ++ax) { Q.lo[k][ax] = c.lo[ax]; Q.hi[k][ax] = c.hi[ax]; }
              Q.zdb[k] = c.zmask_db; Q.zeh[k] = c.zmask_eh;
              Q.csrc[k] = c.csrc; Q.ncsrc[k] = c.ncsrc; Q.ca0[k] = c.csrc_amin; Q.ca1[k] = c.csrc_amax;
              Q.isrc[k] = c.e ? c.isrc : 0; Q.nisrc[k] = c.e ? c.nisrc : 0; Q.ia0[k] = c.isrc_amin; Q.ia1[k] = c.isrc_amax;
              if (Q.ncsrc[k] || Q.nisrc[k]) Q.any_src = 1;
              if (virt && half == 0) { // the B/H half reads E_k = chi1inv_k * (D_k - dipoles) as it loads its operands
                const int ecv = dir_comp(MEEPCU_E_STUFF, k), dcv = dir_comp(MEEPCU_D_STUFF, k);
                const SrcRange &RE = ranges[(ic * 2 + 1) * 3 + k];
                Q.virt = 1;
                Q.vg[k] = dev(ch.d.f[dcv][cmp]);
                Q.vu[k] = dev(ch.d.chi1inv[ecv][k]);
                Q.vsrc[k] = d_src + RE.i0; Q.nvsrc[k] = (int)(RE.i1 - RE.i0); Q.va0[k] = RE.ia0; Q.va1[k] = RE.ia1;
                if (Q.vu[k]) virt_vu = true;
                VirtComp vc = {dev(ch.d.f[ecv][cmp]), Q.vg[k], dev(ch.d.chi1inv[ecv][k]), Q.vsrc[k], Q.nvsrc[k], cmp, (int)ic, ecv,
                               (long long)g.ntot_dev};
                virt_comps.push_back(vc);
                virt_alias[(const char *)ch.d.f[ecv][cmp]] = (const char *)ch.d.f[dcv][cmp];
                if (ch.d.chi1inv[ecv][k]) virt_mul[(const char *)ch.d.f[ecv][cmp]] = (const char *)ch.d.chi1inv[ecv][k];
              }
            }
            const int qidx = (int)h_c3d.size();
            h_c3d.push_back(Q);
            c3d_chunk.push_back((int)ic);
            bool any_zero = false;
            for (int k = 0; k < 3; ++k)
              if (Q.zdb[k] || Q.zeh[k]) any_zero = true;
            const int V = (int)(16 / sizeof(T));
            const bool vec_ok = !getenv("MEEPCU_NO_VEC") && g.pitch % V == 0;
            // device indices of every source point of this chunk half (any component)
            std::vector<long long> spts;
            std::vector<long long> vpts; // dipole points of a virtual E: special for every B/H tile that READS them
            for (int k = 0; k < 3; ++k) {
              for (int j = 0; j < Q.ncsrc[k]; ++j) spts.push_back(all_src[(Q.csrc[k] - d_src) + j].idx);
              for (int j = 0; j < Q.nisrc[k]; ++j) spts.push_back(all_src[(Q.isrc[k] - d_src) + j].idx);
              for (int j = 0; j < Q.nvsrc[k]; ++j) vpts.push_back(all_src[(Q.vsrc[k] - d_src) + j].idx);
            }
            const long long sAq = (long long)Q.sA;
            const int span = vec_ok ? kThreads * V : kThreads; // elements of the plane per clean tile
            for (int a0 = 0; a0 < g.nax[0]; a0 += seg) {
              const int a1 = std::min(a0 + seg, g.nax[0]);
              for (int e0 = 0; e0 < plane; e0 += span) {
                bool special = any_zero; // metal planes: checked per point in the special body
                for (size_t j = 0; j < spts.size() && !special; ++j) {
                  const long long a = spts[j] / sAq, pos = spts[j] % sAq;
                  if (a >= a0 && a < a1 && pos >= e0 && pos < e0 + span) special = true;
                }
                for (size_t j = 0; j < vpts.size() && !special; ++j) { // footprint of the tile's loads: +1 in x, y, z
                  const long long a = vpts[j] / sAq, pos = vpts[j] % sAq;
                  if (a >= a0 && a <= a1 && pos >= e0 && pos <= e0 + span + g.pitch) special = true;
                }
                if (!special && vec_ok) { // clean: V z-points per thread
                  Tile t = {qidx, a0, a1, e0 / V};
                  c3dv_tiles[half].push_back(t);
                }
                else // few tiles: cut them into short marches so that they finish quickly
                  for (int b0 = a0; b0 < a1; b0 += 4)
                    for (int p0 = e0; p0 < std::min(e0 + span, plane); p0 += kThreads) {
                      Tile t = {qidx, b0, std::min(b0 + 4, a1), p0 | (special ? kTileSpecial : 0)};
                      c3d_tiles[half].push_back(t);
                    }
              }
            }
          }
          else {
            // class: 0 = no conductivity, no polarisations; 1 = everything compiled in
            int heavy = 0;
            for (int k = 0; k < 3; ++k) {
              if (P.comp[k].f && P.comp[k].npol) heavy = std::max(heavy, 1);
              if (P.comp[k].f && P.comp[k].cnd) heavy = std::max(heavy, 2);
              if (P.comp[k].f && (P.comp[k].gb || P.comp[k].fb)) heavy = 3; // beta / BFAST terms: the most general kernel
              if (cfg.dim == MEEPCU_Dcyl) heavy = 3;                            // ... and the cylindrical i m / r terms
            }
            // march along the longer outer axis: thin PML slabs still give long streams
            P.march = (g.nax[1] > g.nax[0]) ? 1 : 0;
            const int n_m = P.march ? g.nax[1] : g.nax[0], n_o = P.march ? g.nax[0] : g.nax[1];
            const int pidx = (int)h_half.size();
            h_half.push_back(P);
            half_chunk.push_back((int)ic);
            const int gplane = n_o * g.pitch;
            int gseg = n_m <= 24 ? n_m : 16;
            if (getenv("MEEPCU_GSEG")) gseg = std::max(1, atoi(getenv("MEEPCU_GSEG")));
            const int V = (int)(16 / sizeof(T));
            const bool vec_ok = !getenv("MEEPCU_NO_VEC") && g.pitch % V == 0 && heavy < 2 && cfg.dim != MEEPCU_Dcyl &&
                                (sizeof(T) == 4 || getenv("MEEPCU_GVEC64") || (mix_want && heavy == 0)) &&
                                !(cfg.dim == MEEPCU_D2 && cfg.beta != 0) && // the beta and BFAST terms live in
                                !(cfg.bfast_scaled_k[0] || cfg.bfast_scaled_k[1] || cfg.bfast_scaled_k[2]); // the scalar kernel
            const int span = vec_ok ? kThreads * V : kThreads;
            const long long s_mq = P.march ? (long long)g.pitch : (long long)g.pitch * g.nax[1];
            const long long s_oq = P.march ? (long long)g.pitch * g.nax[1] : (long long)g.pitch;
            for (int a0 = 0; a0 < n_m; a0 += gseg) {
              const int a1 = std::min(a0 + gseg, n_m);
              for (int e0 = 0; e0 < gplane; e0 += span)
                for (int k = 0; k < 3; ++k) { // one CTA per component, neighbours in the table
                  const CompParams<T> &c = P.comp[k];
                  if (!c.f) continue;
                  // does a source point of this component fall into this tile?
                  bool special = false;
                  for (int pass = 0; pass < 2 && !special; ++pass) {
                    const SrcEntry *tab = pass ? c.isrc : c.csrc;
                    const int ns = pass ? (c.e ? c.nisrc : 0) : c.ncsrc;
                    for (int j = 0; j < ns && !special; ++j) {
                      const long long idx = all_src[(tab - d_src) + j].idx;
                      // decompose idx = im*s_m + io*s_o + ic on this chunk's device layout
                      const long long big = std::max(s_mq, s_oq), small = std::min(s_mq, s_oq);
                      const long long q1 = idx / big, r1 = idx % big, q2 = r1 / small, icz = r1 % small;
                      const long long im = (s_mq > s_oq) ? q1 : q2, io = (s_mq > s_oq) ? q2 : q1;
                      const long long pos = io * g.pitch + icz;
                      if (im >= a0 && im < a1 && pos >= e0 && pos < e0 + span) special = true;
                    }
                  }
                  if (vec_ok && !special) {
                    Tile t = {pidx * 4 + k, a0, a1, e0 / V};
                    gv_tiles[half][heavy].push_back(t);
                  }
                  else
                    for (int p0 = e0; p0 < std::min(e0 + span, gplane); p0 += kThreads) {
                      Tile t = {pidx * 4 + k, a0, a1, p0};
                      tiles[half][heavy].push_back(t);
                    }
                }
            }
            simple = heavy == 0; // profiling class below
          }
          // algorithmic bytes: every array element once per read and once per write
          for (int k = 0; k < 3; ++k) {
            const CompParams<T> &c = P.comp[k];
            if (!c.f) continue;
            int v = 3; // R g (one compulsory read of each curl operand is counted with its own
                       // component below: 2 operands ~ 1 per component on average), R+W f
            if (virt && half == 0 && ch.d.chi1inv[dir_comp(MEEPCU_E_STUFF, k)][k]) v += 1; // virtual E: R chi1inv with the operand
            if (c.fu) v += 2;
            if (c.fcnd) v += 2;
            if (c.cnd) v += 2;
            if (c.e) {
              v += 1;
              if (c.u) v += 1;
              if (c.fw) v += 3; // R+W fw, E becomes RMW
              for (int ip = 0; ip < c.npol; ++ip) {
                if (c.pol[ip].p) v += 1;
                if (c.pol[ip].p && c.pol[ip].s) v += 4;
              }
            }
            alg_bytes += (long long)v * nc * (long long)sizeof(T);
            const int kcl = fast ? (half ? KC_CURL3D_D : KC_CURL3D_B) : (half ? KC_HALF_D_SIMPLE : KC_HALF_B_SIMPLE) + (simple ? 0 : 1);
            kc_bytes[kcl] += (long long)v * nc * (long long)sizeof(T);
            int vs = v; // the E array is written (with its chi1inv read) in the D/E half, read in the B/H half
            if (virt) {
              const bool has_u = ch.d.chi1inv[dir_comp(MEEPCU_E_STUFF, k)][k] != 0;
              if (half == 0 && has_u) vs -= 1;
              if (half == 1) vs += 1 + (has_u ? 1 : 0);
            }
            alg_bytes_sv += (long long)vs * nc * (long long)sizeof(T);
            kc_bytes_sv[kcl] += (long long)vs * nc * (long long)sizeof(T);
          }
        }
        // ---- general E/H components of this chunk ----
        if (general) {
          for (int cmp = 0; cmp < ncmp; ++cmp) {
            const T *dmp[3] = {0, 0, 0};
            for (int k = 0; k < 3; ++k) {
              const int dc = dir_comp(ft2, k), ec = dir_comp(ft, k);
              if (!ch.d.f[dc][cmp]) continue;
              bool need_fmp = false;
              if (ch.d.f[ec][cmp]) {
                need_fmp = have_int_sources;
                for (size_t ip = 0; ip < ch.pol[ft].size(); ++ip)
                  if (ch.pol[ft][ip].d.P[ec][cmp]) need_fmp = true;
              }
              bool has_pol = false;
              for (size_t ip = 0; ip < ch.pol[ft].size(); ++ip)
                if (ch.d.f[ec][cmp] && ch.pol[ft][ip].d.P[ec][cmp]) has_pol = true;
              if (need_fmp && !has_pol && !getenv("MEEPCU_FMP_ALWAYS")) {
                // f_minus_p = D - dipoles: not materialised (a whole-array pass for a handful of points); the dipole
                // points of this component are patched in place around k_eh_general instead (DipPatch)
                const SrcRange &R = ranges[(ic * 2 + half) * 3 + k];
                if (R.i1 > R.i0) {
                  DipPatch<T> dp;
                  memset(&dp, 0, sizeof dp);
                  dp.d = dev(ch.d.f[dc][cmp]);
                  dp.isrc = d_src + R.i0;
                  dp.n = (int)(R.i1 - R.i0);
                  dp.is_D = half;
                  dp.cmp = cmp;
                  CU(cudaMalloc((void **)&dp.saved, (size_t)dp.n * sizeof(T)));
                  extra_dev.push_back((char *)dp.saved);
                  dips[half].push_back(dp);
                }
                dmp[k] = dev(ch.d.f[dc][cmp]);
              }
              else if (need_fmp) {
                T *scratch = 0;
                CU(cudaMalloc((void **)&scratch, g.ntot_dev * sizeof(T)));
                extra_dev.push_back((char *)scratch);
                CU(cudaMemset(scratch, 0, g.ntot_dev * sizeof(T)));
                FmpParams<T> F;
                memset(&F, 0, sizeof F);
                F.fmp = scratch;
                F.d = dev(ch.d.f[dc][cmp]);
                F.npol = (int)ch.pol[ft].size();
                for (size_t ip = 0; ip < ch.pol[ft].size(); ++ip) F.p[ip] = dev(ch.pol[ft][ip].d.P[ec][cmp]);
                F.n = (long long)g.ntot_dev;
                const SrcRange &R = ranges[(ic * 2 + half) * 3 + k];
                // the fused tables are sorted by index; k_fmp_src needs reference order only per
                // index, which stable sorting preserved
                F.isrc = d_src + R.i0;
                F.nisrc = (int)(R.i1 - R.i0);
                if (F.nisrc) fmp_has_src[half] = true;
                F.is_D = half;
                F.cmp = cmp;
                fmps[half].push_back(F);
                dmp[k] = scratch;
              }
              else
                dmp[k] = dev(ch.d.f[dc][cmp]);
            }
            for (int k = 0; k < 3; ++k) {
              const int dc = dir_comp(ft2, k), ec = dir_comp(ft, k);
              if (!ch.d.f[ec][cmp] || ch.d.f[ec][cmp] == ch.d.f[dc][cmp]) continue;
              EhParams<T> E;
              memset(&E, 0, sizeof E);
              E.nA = g.nax[0]; E.nB = g.nax[1]; E.nC = g.nax[2]; E.pitch = g.pitch;
              E.sA = (long long)g.pitch * g.nax[1]; E.sB = g.pitch;
              for (int ax = 0; ax < 3; ++ax) {
                const int d = g.dir_of_axis[ax];
                E.lo[ax] = d < 0 ? 0 : 1 - g.yee_shift(ec, d);
                E.hi[ax] = d < 0 ? 0 : g.num[d] - g.yee_shift(ec, d);
              }
              const int d1 = (k + 1) % 3, d2 = (k + 2) % 3;
              const int sgn = ft == MEEPCU_H_STUFF ? -1 : +1;
              E.f = dev(ch.d.f[ec][cmp]);
              E.fc = ec;
              E.g = dmp[k]; E.g1 = dmp[d1]; E.g2 = dmp[d2];
              E.u = dev(ch.d.chi1inv[ec][k]);
              E.u1 = E.g1 ? dev(ch.d.chi1inv[ec][d1]) : 0;
              E.u2 = E.g2 ? dev(ch.d.chi1inv[ec][d2]) : 0;
              E.s = g.ds[k] * sgn; E.s1 = g.ds[d1] * sgn; E.s2 = g.ds[d2] * sgn;
              if ((!E.g1 && E.g2) || (E.g1 && E.g2 && !E.u1 && E.u2)) { // step_generic.cpp:590-594
                std::swap(E.g1, E.g2); std::swap(E.u1, E.u2); std::swap(E.s1, E.s2);
              }
              if (!E.u1 && E.u2) return fail(MEEPCU_EINVAL, "bug - didn't swap off-diagonal terms!?");
              E.chi2 = dev(ch.d.chi2[ec]);
              E.chi3 = dev(ch.d.chi3[ec]);
              if ((E.u1 || E.chi3) && !E.u)
                return fail(MEEPCU_EINVAL, "chi1inv diagonal missing where the reference dereferences it");
              if (E.chi3 && !E.chi2 && (cfg.nl_mode == MEEPCU_NL_UPSTREAM_PADE || (E.u1 && E.u2)))
                return fail(MEEPCU_EINVAL, "chi3 without chi2 (the reference dereferences chi2)");
              if (cfg.nl_mode == MEEPCU_NL_FORK_LITERAL && E.u1 && E.u2 && E.chi3 && ch.d.sigsize[k] <= 1 &&
                  !(ec == 0 || ec == 1 || ec == 4))
                return fail(MEEPCU_EUNSUPPORTED, "NR branch reached for a non-E component (step_generic.cpp:808-811)");
              if (ch.d.sigsize[k] > 1) {
                E.fw = dev(ch.d.f_w[ec][cmp]);
                if (!E.fw) return fail(MEEPCU_ESTATE, "f_w not allocated (update_eh.cpp:212-216)");
                E.kwaxis = g.axis_of_dir[k];
                E.kwoff = g.yee_shift(ec, k);
                E.sigw = dev(ch.d.sig[k]); E.kapw = dev(ch.d.kap[k]);
              }
              E.nl_mode = cfg.nl_mode;
              std::map<const char *, unsigned>::iterator zm = zmask.find((const char *)ch.d.f[ec][cmp]);
              E.zmask_eh = zm == zmask.end() ? 0u : zm->second;
              if (aniso_now) E.zmask_eh = 0;
              E.npol = (int)ch.pol[ft].size();
              RC(fill_pols(ch, ft, ec, k, cmp, E.pol));
              E.nr_fail = d_flags;
              // per-point class of the update from the material arrays (EhParams::mask)
              const bool nr_branch = cfg.nl_mode == MEEPCU_NL_FORK_LITERAL && !E.fw && E.u1 && E.u2 && E.chi3;
              const bool pade_diag = cfg.nl_mode == MEEPCU_NL_UPSTREAM_PADE && !E.u1 && !E.u2 && E.chi3;
              if ((nr_branch || pade_diag) && !getenv("MEEPCU_NO_EH_MASK")) {
                unsigned char *mk = 0;
                CU(cudaMalloc((void **)&mk, g.ntot_dev));
                extra_dev.push_back((char *)mk);
                E.mask = mk;
                EhMask em = {mk, E.u, E.u1, E.u2, E.chi2, E.chi3, nr_branch ? 0 : 1, (long long)g.ntot_dev};
                eh_masks.push_back(em);
              }
              const int eidx = (int)h_eh.size();
              h_eh.push_back(E);
              const int plane = g.nax[1] * g.pitch;
              int seg = g.nax[0] <= 48 ? g.nax[0] : 32;
              const bool rows2d = g.pitch % 32 == 0 && g.pitch / 32 < 4096 && !getenv("MEEPCU_EH_FLAT");
              for (int a0 = 0; a0 < g.nax[0]; a0 += seg)
                if (rows2d) {
                  for (int rb = 0; rb < (g.nax[1] + 7) / 8; ++rb)
                    for (int zb = 0; zb < g.pitch / 32; ++zb) {
                      Tile t = {eidx, a0, std::min(a0 + seg, g.nax[0]), kTileRows | (rb << 12) | zb};
                      eh_tiles[half].push_back(t);
                    }
                }
                else
                for (int p0 = 0; p0 < plane; p0 += kThreads) {
                  Tile t = {eidx, a0, std::min(a0 + seg, g.nax[0]), p0};
                  eh_tiles[half].push_back(t);
                }
              alg_bytes += (long long)(4 + (E.fw ? 3 : 0)) * nc * (long long)sizeof(T);
              kc_bytes[KC_EH_GENERAL] += (long long)(4 + (E.fw ? 3 : 0)) * nc * (long long)sizeof(T);
              alg_bytes_sv += (long long)(4 + (E.fw ? 3 : 0)) * nc * (long long)sizeof(T);
              kc_bytes_sv[KC_EH_GENERAL] += (long long)(4 + (E.fw ? 3 : 0)) * nc * (long long)sizeof(T);
            }
          }
        }
        // ---- polarisations of a chunk with an anisotropic sigma: their own pass (k_pol_general) ----
        if (aniso_now) {
          const int sgn = ft == MEEPCU_H_STUFF ? -1 : +1;
          for (size_t ip = 0; ip < ch.pol[ft].size(); ++ip) {
            const meepcu_susceptibility_desc &sd = ch.pol[ft][ip].d;
            if (ch.pol[ft][ip].gyro) {
              RC(build_gyro(ch, (int)ic, ch.pol[ft][ip].g, ft, half, ncmp, nc, polgyro_tiles[half]));
              continue;
            }
            PolParams<T> coef[kMaxPol];
            for (int cmp = 0; cmp < ncmp; ++cmp)
              for (int k = 0; k < 3; ++k) {
                const int ec = dir_comp(ft, k);
                if (!sd.P[ec][cmp]) continue;
                // W = f_w ?: f (update_pols.cpp:43-44)
                const void *wh = ch.d.f_w[ec][cmp] ? ch.d.f_w[ec][cmp] : ch.d.f[ec][cmp];
                if (!wh || !sd.sigma[ec][k]) continue; // susceptibility.cpp:201
                const bool save = aniso_now;
                aniso_now = false;
                RC(fill_pols(ch, ft, ec, k, cmp, coef)); // the realnum coefficients of :191-194
                aniso_now = save;
                PolGenParams<T> Q;
                memset(&Q, 0, sizeof Q);
                Q.nA = g.nax[0]; Q.nB = g.nax[1]; Q.nC = g.nax[2]; Q.pitch = g.pitch;
                Q.sA = (long long)g.pitch * g.nax[1]; Q.sB = g.pitch;
                for (int ax = 0; ax < 3; ++ax) {
                  const int d = g.dir_of_axis[ax];
                  Q.lo[ax] = d < 0 ? 0 : 1 - g.yee_shift(ec, d);
                  Q.hi[ax] = d < 0 ? 0 : g.num[d] - g.yee_shift(ec, d);
                }
                Q.p = dev(sd.P[ec][cmp]); Q.pp = dev(sd.P_prev[ec][cmp]);
                Q.s = dev(sd.sigma[ec][k]); Q.w = dev(wh);
                int d1 = (k + 1) % 3, d2 = (k + 2) % 3;
                const int c1 = dir_comp(ft, d1), c2 = dir_comp(ft, d2);
                const void *w1h = ch.d.f_w[c1][cmp] ? ch.d.f_w[c1][cmp] : ch.d.f[c1][cmp];
                const void *w2h = ch.d.f_w[c2][cmp] ? ch.d.f_w[c2][cmp] : ch.d.f[c2][cmp];
                Q.w1 = dev(w1h); Q.s1 = Q.w1 ? dev(sd.sigma[ec][d1]) : 0;
                Q.w2 = dev(w2h); Q.s2 = Q.w2 ? dev(sd.sigma[ec][d2]) : 0;
                Q.is = g.ds[k] * sgn; Q.is1 = g.ds[d1] * sgn; Q.is2 = g.ds[d2] * sgn;
                if (Q.s2 && !Q.s1) { // :219-224
                  std::swap(Q.w1, Q.w2); std::swap(Q.s1, Q.s2); std::swap(Q.is1, Q.is2);
                }
                Q.gamma1inv = coef[ip].gamma1inv; Q.gamma1 = coef[ip].gamma1;
                Q.omega0dtsqr = coef[ip].omega0dtsqr; Q.omega0dtsqr_denom = coef[ip].omega0dtsqr_denom;
                const int qidx = (int)h_polgen.size();
                h_polgen.push_back(Q);
                const int plane = g.nax[1] * g.pitch;
                const int seg = g.nax[0] <= 48 ? g.nax[0] : 32;
                for (int a0 = 0; a0 < g.nax[0]; a0 += seg)
                  for (int p0 = 0; p0 < plane; p0 += kThreads) {
                    Tile t = {qidx, a0, std::min(a0 + seg, g.nax[0]), p0};
                    polgen_tiles[half].push_back(t);
                  }
                const int v = 5 + (Q.s1 ? 2 : 0) + (Q.s2 ? 2 : 0); // R,W p, R,W pp, R s (+ R s1,w1 / s2,w2; W itself counted with E)
                alg_bytes += (long long)v * nc * (long long)sizeof(T);
                kc_bytes[KC_POL_GENERAL] += (long long)v * nc * (long long)sizeof(T);
                alg_bytes_sv += (long long)v * nc * (long long)sizeof(T);
                kc_bytes_sv[KC_POL_GENERAL] += (long long)v * nc * (long long)sizeof(T);
              }
          }
        }
        aniso_now = false;
        if (half == 1) {
          for (size_t im = 0; im < ch.ml.size(); ++im) RC(build_ml(ch, ch.ml[im], ncmp, nc, mllev_tiles, mlpol_tiles));
          for (int k = 0; k < 3; ++k) // W -> W_prev before the chunk's E update
            for (int cmp = 0; cmp < ncmp; ++cmp) {
              const int ec = dir_comp(MEEPCU_E_STUFF, k);
              if (!ch.w_prev[ec][cmp] || !ch.d.f[ec][cmp]) continue;
              CopyItem<T> it = {dev(ch.w_prev[ec][cmp]), dev(ch.d.f_w[ec][cmp] ? ch.d.f_w[ec][cmp] : ch.d.f[ec][cmp]),
                                (long long)g.ntot_dev};
              wprev_items.push_back(it);
              wprev_max = std::max(wprev_max, it.n);
              alg_bytes += 2LL * nc * (long long)sizeof(T);
              alg_bytes_sv += 2LL * nc * (long long)sizeof(T);
            }
        }
      }
    }
    n_wprev = (int)wprev_items.size();
    RC(upload(wprev_items, &d_wprev));
    RC(build_noise());
    tm_mark("parameter blocks + tiles");
    // ---- 5a. which halo exchanges run every step.  The schedule must be the SAME on every rank: a rank that
    // owns only PML chunks and a rank that owns nonlinear ones would otherwise post different ncclSend/ncclRecv
    // groups (one side waiting for a D halo the other never sends).  So the flags are OR-ed over the communicator.
    for (size_t ic = 0; ic < chunks.size(); ++ic)
      for (size_t id = 0; id < chunks[ic].dft.size(); ++id) { // a DFT of D (B) reads its halos: dft.cpp:284-293
        const int c = chunks[ic].dft[id].d.c;
        if (c >= 0 && c < 20 && chunks[ic].d.f[c][0]) {
          if (is_D(c)) eager_db[1] = true;
          if (is_B(c)) eager_db[0] = true;
        }
      }
    for (int half = 0; half < 2; ++half) aniso_half[half] = !polgen_tiles[half].empty() || !polgyro_tiles[half].empty();
    if (!mllev_tiles.empty() || n_wprev) aniso_half[1] = true; // multilevel atoms: ordered path (W, P halos around the passes)
    for (int half = 0; half < 2; ++half)
      if (n_noise_seg[half]) aniso_half[half] = true; // k_pol_noise follows the fused update on the ordered path only
    virt_any = !virt_comps.empty();
    if (n_ranks > 1) {
      int hflag[9] = {general_half[0], general_half[1], eager_db[0], eager_db[1], eager_p[0], eager_p[1], aniso_half[0], aniso_half[1],
                      virt_any};
      int *dflag = 0;
      CU(cudaMalloc((void **)&dflag, sizeof hflag));
      extra_dev.push_back((char *)dflag);
      CU(cudaMemcpyAsync(dflag, hflag, sizeof hflag, cudaMemcpyHostToDevice, st));
      NC(nccl_api()->AllReduce(dflag, dflag, 9, 2 /* ncclInt32 */, 2 /* ncclMax */, comm, st));
      CU(cudaMemcpyAsync(hflag, dflag, sizeof hflag, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      general_half[0] = hflag[0]; general_half[1] = hflag[1];
      eager_db[0] = hflag[2]; eager_db[1] = hflag[3];
      eager_p[0] = hflag[4]; eager_p[1] = hflag[5];
      aniso_half[0] = hflag[6]; aniso_half[1] = hflag[7];
      virt_any = hflag[8];
    }
    RC(virt_patch_u());
    RC(build_eh_masks());
    // ---- 5b. multi-rank: boundary tiles first (see nb_* above) ----
    if (n_ranks > 1 && !getenv("MEEPCU_NO_OVERLAP"))
      for (int half = 0; half < 2; ++half) {
        if (general_half[half]) continue; // k_eh_general needs the D halo before E exists: ordered path
        if (cfg.dim == MEEPCU_Dcyl) continue; // the axis rules (k_cyl_origin, late zero list) live on the ordered path
        if (aniso_half[half]) continue;   // k_pol_general / k_pol_gyro need the W halo before P: ordered path
        if (eager_db[half] || eager_p[half]) continue; // D/B or P halos travel every step (a D/B DFT): ordered path
        const int fts[2] = {half ? MEEPCU_E_STUFF : MEEPCU_H_STUFF, half ? MEEPCU_WE_STUFF : MEEPCU_WH_STUFF};
        // per chunk: the planes (index on device axis A) that hold points another rank receives
        std::vector<std::set<int> > planes(chunks.size());
        bool any_send = false;
        for (int q = 0; q < 2; ++q)
          for (typename std::map<int, RemoteList>::iterator it = rsend[fts[q]].begin(); it != rsend[fts[q]].end(); ++it)
            for (int kind = 0; kind < 3; ++kind)
              for (size_t k = 0; k < it->second.ptr[kind].size(); ++k) {
                size_t e = 0;
                const ArrayRec *r = find(it->second.ptr[kind][k], sizeof(T), &e);
                if (!r || r->kind != A_GRID) return fail(MEEPCU_EINVAL, "remote send pointer outside the registered arrays");
                const Geom &g = chunks[r->chunk].g;
                planes[r->chunk].insert((int)(g.dev_index((long long)e) / ((long long)g.pitch * g.nax[1])));
                any_send = true;
              }
        if (!any_send && rrecv[fts[0]].empty() && rrecv[fts[1]].empty()) continue;
        struct IsBoundary {
          const std::vector<std::set<int> > *planes;
          const std::vector<int> *chunk_of;
          const std::vector<HalfParams<T> > *hp; // NULL: k_curl3d tiles (always march along A)
          int shift, V;
          bool operator()(const Tile &t) const {
            const int pi = t.param >> shift;
            const std::set<int> &S = (*planes)[(*chunk_of)[pi]];
            if (S.empty()) return false;
            if (S.size() > 8) return true; // not a slab cut along A: no split for this chunk
            int lo = t.a0, hi = t.a1 - 1;
            if (hp && (*hp)[pi].march) { // tile = positions (A,C) marching along B
              const int pitch = (*hp)[pi].pitch;
              const long long p0 = (long long)(t.p0 & ~kTileSpecial) * V;
              lo = (int)(p0 / pitch);
              hi = (int)((p0 + (long long)kThreads * V - 1) / pitch);
            }
            std::set<int>::const_iterator f = S.lower_bound(lo);
            return f != S.end() && *f <= hi;
          }
        };
        const int V = (int)(16 / sizeof(T));
        IsBoundary b3 = {&planes, &c3d_chunk, 0, 0, 1}, bg = {&planes, &half_chunk, &h_half, 2, 1};
        IsBoundary b3v = b3, bgv = bg;
        b3v.V = V; bgv.V = V;
        nb_c3d_tiles[half] = (int)(std::stable_partition(c3d_tiles[half].begin(), c3d_tiles[half].end(), b3) - c3d_tiles[half].begin());
        nb_c3dv_tiles[half] = (int)(std::stable_partition(c3dv_tiles[half].begin(), c3dv_tiles[half].end(), b3v) - c3dv_tiles[half].begin());
        for (int cl = 0; cl < 4; ++cl)
          nb_tiles[half][cl] = (int)(std::stable_partition(tiles[half][cl].begin(), tiles[half][cl].end(), bg) - tiles[half][cl].begin());
        for (int cl = 0; cl < 2; ++cl)
          nb_gv_tiles[half][cl] = (int)(std::stable_partition(gv_tiles[half][cl].begin(), gv_tiles[half][cl].end(), bgv) - gv_tiles[half][cl].begin());
        overlap_half[half] = true;
      }
    // ---- 5c. one grid for interior + PML tiles (k_mix) ----
    {
      for (int half = 0; half < 2; ++half) {
        std::vector<Tile> &A = c3dv_tiles[half], &B = gv_tiles[half][0];
        if ((A.empty() || !mix_want) && !B.empty() && sizeof(T) == 8 && !getenv("MEEPCU_GVEC64")) {
          // fp64 PML tiles on their own run faster in the scalar kernel: V scalar tiles per vector tile
          const int V = (int)(16 / sizeof(T));
          std::vector<Tile> &S0 = tiles[half][0];
          std::vector<Tile> bnd, rest;
          for (size_t k = 0; k < B.size(); ++k)
            for (int j = 0; j < V; ++j) {
              Tile t = B[k];
              t.p0 = B[k].p0 * V + j * kThreads;
              ((int)k < nb_gv_tiles[half][0] ? bnd : rest).push_back(t);
            }
          S0.insert(S0.begin() + nb_tiles[half][0], bnd.begin(), bnd.end());
          nb_tiles[half][0] += (int)bnd.size();
          S0.insert(S0.end(), rest.begin(), rest.end());
          B.clear();
          nb_gv_tiles[half][0] = 0;
        }
        if (A.empty() || B.empty() || !mix_want) continue;
        // worth it only where the PML tiles are a real share of the half step (vacuum cube with PML on
        // six faces: 28 % of the tiles; a slab with thin PML on two faces: 2 % -> two streams are better)
        if (B.size() * 20 < A.size() + B.size() && !getenv("MEEPCU_MIX")) continue;
        std::vector<Tile> M;
        M.reserve(A.size() + B.size());
        const size_t abeg[3] = {0, (size_t)nb_c3dv_tiles[half], A.size()};
        const size_t bbeg[3] = {0, (size_t)nb_gv_tiles[half][0], B.size()};
        for (int seg = 0; seg < 2; ++seg) { // cross-rank boundary tiles, then the rest
          const size_t na = abeg[seg + 1] - abeg[seg], nb = bbeg[seg + 1] - bbeg[seg], tot = na + nb;
          size_t ia = abeg[seg], ib = bbeg[seg];
          // the (slower) PML tiles are spread over the first `front` per cent of the table, so that the
          // tail of the grid is made of short interior tiles
          // (measured at 256^3, ms/step for front = 100 / 96 / 93 / 90 / 85 / 80 / 60:
          //  fp32 0.352 / 0.349 / 0.344 / 0.341 / 0.342 / 0.343 / 0.364, fp64 0.629 / 0.614 / 0.616 / 0.617 / 0.627 / 0.635 / 0.673)
          static const int front = getenv("MEEPCU_MIX_FRONT") ? std::min(100, std::max(1, atoi(getenv("MEEPCU_MIX_FRONT"))))
                                                              : (sizeof(T) == 4 ? 88 : 94);
          const size_t span = std::max<size_t>(nb, tot * front / 100);
          for (size_t k = 0; k < tot; ++k) {
            const bool pickb = nb && k < span && ((k + 1) * nb / span > k * nb / span);
            if ((pickb && ib < bbeg[seg + 1]) || ia >= abeg[seg + 1]) {
              Tile t = B[ib++];
              t.p0 |= kTileGeneral;
              M.push_back(t);
            }
            else
              M.push_back(A[ia++]);
          }
          if (seg == 0) nb_mix_tiles[half] = (int)M.size();
        }
        n_mix_tiles[half] = (int)M.size();
        RC(upload(M, &d_mix_tiles[half]));
        A.clear();
        B.clear();
        nb_c3dv_tiles[half] = nb_gv_tiles[half][0] = 0;
        const int km = half ? KC_MIX_D : KC_MIX_B, kc = half ? KC_CURL3D_D : KC_CURL3D_B,
                  kh = half ? KC_HALF_D_SIMPLE : KC_HALF_B_SIMPLE;
        kc_bytes[km] = kc_bytes[kc] + kc_bytes[kh];
        kc_bytes[kc] = kc_bytes[kh] = 0;
        kc_bytes_sv[km] = kc_bytes_sv[kc] + kc_bytes_sv[kh];
        kc_bytes_sv[kc] = kc_bytes_sv[kh] = 0;
      }
    }
    for (int half = 0; half < 2; ++half) {
      n_cyl_rd[half] = (int)cyl_rd[half].size();
      RC(upload(cyl_rd[half], &d_cyl_rd[half]));
      n_cyl_org[half] = (int)cyl_org[half].size();
      RC(upload(cyl_org[half], &d_cyl_org[half]));
    }
    RC(upload(h_half, &d_half));
    RC(upload(h_c3d, &d_c3d));
    for (int half = 0; half < 2; ++half) {
      n_c3d_tiles[half] = (int)c3d_tiles[half].size();
      RC(upload(c3d_tiles[half], &d_c3d_tiles[half]));
      n_c3dv_tiles[half] = (int)c3dv_tiles[half].size();
      RC(upload(c3dv_tiles[half], &d_c3dv_tiles[half]));
      for (int cl = 0; cl < 2; ++cl) {
        n_gv_tiles[half][cl] = (int)gv_tiles[half][cl].size();
        RC(upload(gv_tiles[half][cl], &d_gv_tiles[half][cl]));
      }
    }
    RC(upload(h_eh, &d_eh));
    for (int half = 0; half < 2; ++half) { // the metal points of E (H) in those chunks, zeroed after the polarisation pass
      const int fte = half ? MEEPCU_E_STUFF : MEEPCU_H_STUFF;
      std::vector<T *> late;
      for (size_t k = 0; k < metal[fte].size(); ++k) {
        size_t e = 0;
        const ArrayRec *r = find(metal[fte][k], sizeof(T), &e);
        if (r && r->kind == A_GRID && aniso_chunk_half.count(std::make_pair(r->chunk, half))) late.push_back(dev_elem(metal[fte][k]));
      }
      const int ftd = half ? MEEPCU_D_STUFF : MEEPCU_B_STUFF;
      late.insert(late.end(), metal_axis[ftd].begin(), metal_axis[ftd].end());
      late.insert(late.end(), metal_axis[fte].begin(), metal_axis[fte].end());
      n_zero_late[half] = (long long)late.size();
      RC(upload(late, &d_zero_late[half]));
    }
    RC(upload(h_polgen, &d_polgen));
    RC(upload(h_polgyro, &d_polgyro));
    RC(upload(h_mllev, &d_mllev));
    RC(upload(h_mlpol, &d_mlpol));
    n_mllev_tiles = (int)mllev_tiles.size();
    RC(upload(mllev_tiles, &d_mllev_tiles));
    n_mlpol_tiles = (int)mlpol_tiles.size();
    RC(upload(mlpol_tiles, &d_mlpol_tiles));
    for (int half = 0; half < 2; ++half) {
      n_polgen_tiles[half] = (int)polgen_tiles[half].size();
      RC(upload(polgen_tiles[half], &d_polgen_tiles[half]));
      n_polgyro_tiles[half] = (int)polgyro_tiles[half].size();
      RC(upload(polgyro_tiles[half], &d_polgyro_tiles[half]));
    }
    for (int half = 0; half < 2; ++half) {
      for (int cl = 0; cl < 4; ++cl) {
        n_tiles[half][cl] = (int)tiles[half][cl].size();
        RC(upload(tiles[half][cl], &d_tiles[half][cl]));
      }
      n_eh_tiles[half] = (int)eh_tiles[half].size();
      RC(upload(eh_tiles[half], &d_eh_tiles[half]));
      n_fmp[half] = (int)fmps[half].size();
      RC(upload(fmps[half], &d_fmp[half]));
      n_dip[half] = (int)dips[half].size();
      RC(upload(dips[half], &d_dip[half]));
    }

    tm_mark("schedule, boundary tiles, mix");
    // ---- 6. connections ----
    // which E/H connections can wait for the next host sync: 3-D, and nobody but the curl kernels reads
    // E/H neighbours (no k_eh_general / k_pol_general / DFT averaging / cross-rank lists built from them)
    size_t dfts_total = 0;
    for (size_t ic = 0; ic < chunks.size(); ++ic) dfts_total += chunks[ic].dft.size();
    const bool defer_normal = cfg.dim == MEEPCU_D3 && !general_half[0] && !general_half[1] && h_polgen.empty() && h_polgyro.empty() &&
                              h_mllev.empty() && dfts_total == 0 && !getenv("MEEPCU_NO_DEFER");
    std::map<const char *, int> comp_of; // host base of an f array -> component (direction is what matters)
    if (defer_normal)
      for (size_t ic = 0; ic < chunks.size(); ++ic)
        for (int c = 0; c < 10; ++c) // E and H components (H may alias B: same direction)
          for (int cmp = 0; cmp < 2; ++cmp)
            if (chunks[ic].d.f[c][cmp]) comp_of[(const char *)chunks[ic].d.f[c][cmp]] = c;
    for (int ft = 0; ft < MEEPCU_NUM_FIELD_TYPES; ++ft) {
      std::vector<ConnRun<T> > runs[2]; // [0] every step, [1] deferred
      for (int kind = MEEPCU_CONNECT_NEGATE; kind <= MEEPCU_CONNECT_COPY; ++kind) {
        const PtrList &L0 = conn[ft][kind];
        PtrList part[2];
        for (size_t k = 0; k < L0.dst.size(); ++k) {
          int which = 0;
          if (defer_normal && (ft == MEEPCU_E_STUFF || ft == MEEPCU_H_STUFF)) {
            size_t e = 0;
            const ArrayRec *r = find(L0.dst[k], sizeof(T), &e);
            std::map<const char *, int>::iterator ci = r ? comp_of.find((const char *)r->host) : comp_of.end();
            if (r && r->kind == A_GRID && ci != comp_of.end()) {
              const Geom &g = chunks[r->chunk].g;
              const int c = ci->second, d = comp_dir(c);
              const long long coord = ((long long)e / g.hs[d]) % (g.num[d] + 1);
              const long long notowned = g.yee_shift(c, d) ? g.num[d] : 0;
              if (coord == notowned) which = 1;
            }
          }
          part[which].dst.push_back(L0.dst[k]);
          part[which].src.push_back(L0.src[k]);
        }
        for (int which = 0; which < 2; ++which) {
        const PtrList &L = part[which];
        size_t k = 0;
        while (k < L.dst.size()) {
          const T *m0 = 0, *mx = 0;
          T *d0 = dev_elem_v(L.dst[k], ft, &mx);
          const T *s0 = dev_elem_v(L.src[k], ft, &m0);
          if (!d0 || !s0)
            return fail(MEEPCU_EINVAL, "connection %zu of field type %d points outside the registered arrays", k, ft);
          ConnRun<T> r = {d0, s0, 1, 0, 0, kind == MEEPCU_CONNECT_NEGATE, m0};
          if (k + 1 < L.dst.size()) {
            const T *m1 = 0;
            T *d1 = dev_elem_v(L.dst[k + 1], ft, &mx);
            const T *s1 = dev_elem_v(L.src[k + 1], ft, &m1);
            if (d1 && s1) {
              const long long ds = d1 - d0, ss = s1 - s0;
              if (ds != 0 && std::llabs(ds) < (1LL << 30) && std::llabs(ss) < (1LL << 30) && m1 == (m0 ? m0 + ss : 0)) {
                r.dst_stride = (int)ds;
                r.src_stride = (int)ss;
                T *dp = d1;
                const T *sp = s1;
                r.n = 2;
                while (k + r.n < L.dst.size() && r.n < (1 << 20)) {
                  const T *mn = 0;
                  T *dn = dev_elem_v(L.dst[k + r.n], ft, &mx);
                  const T *sn = dev_elem_v(L.src[k + r.n], ft, &mn);
                  if (!dn || !sn || dn - dp != ds || sn - sp != ss || mn != (m0 ? m0 + (sn - s0) : 0)) break;
                  dp = dn; sp = sn; r.n++;
                }
              }
            }
          }
          for (int o = 0; o < r.n; o += kConnPiece) { // k_conn_runs: one warp per piece
            ConnRun<T> pc = r;
            pc.dst = r.dst + (long long)o * r.dst_stride;
            pc.src = r.src + (long long)o * r.src_stride;
            if (r.mul) pc.mul = r.mul + (long long)o * r.src_stride;
            pc.n = std::min(kConnPiece, r.n - o);
            runs[which].push_back(pc);
          }
          k += r.n;
        }
        }
      }
      n_runs[ft] = (int)runs[0].size();
      RC(upload(runs[0], &d_runs[ft]));
      n_runs_def[ft] = (int)runs[1].size();
      RC(upload(runs[1], &d_runs_def[ft]));
    }
    for (int ft = 0; ft < MEEPCU_NUM_FIELD_TYPES; ++ft)
      for (int kind = 0; kind < 1; ++kind) { // PHASE lists keep the per-element form
        const PtrList &L = conn[ft][kind];
        ConnTable<T> &C = d_conn[ft][kind];
        memset(&C, 0, sizeof C);
        if (L.dst.empty()) continue;
        std::vector<T *> hd(L.dst.size());
        std::vector<const T *> hsrc(L.src.size()), hmul(L.src.size());
        bool any_mul = false;
        for (size_t n = 0; n < L.dst.size(); ++n) {
          const T *mx = 0;
          hd[n] = dev_elem_v(L.dst[n], ft, &mx);
          hsrc[n] = dev_elem_v(L.src[n], ft, &hmul[n]);
          if (hmul[n]) any_mul = true;
          if (!hd[n] || !hsrc[n])
            return fail(MEEPCU_EINVAL, "connection %zu of field type %d points outside the registered arrays", n, ft);
        }
        T **dd = 0;
        const T **dsrc = 0, **dmul = 0;
        RC(upload(hd, &dd));
        RC(upload(hsrc, &dsrc));
        if (any_mul) RC(upload(hmul, &dmul));
        C.dst = dd;
        C.src = dsrc;
        C.mul = dmul;
        C.n = (long long)L.dst.size();
        if (kind == MEEPCU_CONNECT_PHASE) {
          std::vector<T> ph((const T *)L.phases.data(), (const T *)(L.phases.data() + L.phases.size()));
          T *dph = 0;
          RC(upload(ph, &dph));
          C.phase = dph;
        }
      }
    // W exchanges are only non-empty for anisotropic susceptibilities (rejected above)
    RC(build_remote());

    tm_mark("connections");
    // ---- 7. DFT chunks ----
    std::vector<DftWork> work;
    phase_row = 0;
    for (size_t ic = 0; ic < chunks.size(); ++ic) {
      ChunkH &ch = chunks[ic];
      const Geom &g = ch.g;
      for (size_t id = 0; id < ch.dft.size(); ++id) {
        DftH &D = ch.dft[id];
        const meepcu_dft_desc &d = D.d;
        if (d.c < 0 || d.c >= 20 || !ch.d.f[d.c][0]) continue; // dft.cpp:266
        if (is_D(d.c)) eager_db[1] = true;
        if (is_B(d.c)) eager_db[0] = true;
        const bool is_h = is_magnetic(d.c) || is_B(d.c);
        // phase group (same omega list, scale and sampling time share one phase row)
        int grp = -1;
        for (size_t k = 0; k < group_omega.size(); ++k)
          if (group_is_h[k] == (int)is_h && group_scale[k] == std::complex<double>(d.scale[0], d.scale[1]) &&
              group_omega[k] == D.omega)
            grp = (int)k;
        if (grp < 0) {
          grp = (int)group_omega.size();
          group_omega.push_back(D.omega);
          group_scale.push_back(std::complex<double>(d.scale[0], d.scale[1]));
          group_is_h.push_back(is_h);
          phase_off.push_back(phase_row);
          phase_row += d.Nomega;
        }
        D.phase_group = grp;
        DftParams<T> q;
        memset(&q, 0, sizeof q);
        q.f0 = dev(ch.d.f[d.c][0]);
        q.f1 = dev(ch.d.f[d.c][1]);
        q.dft = dev(d.dft);
        // LOOP_OVER_IVECS in the reference's loop order (vec.hpp:151-168; yucky_dir vec.cpp:262-272)
        long long idx0 = 0;
        for (int k = 0; k < 3; ++k) {
          const int dir = loop_slot(k);
          q.dir[k] = dir;
          if (g.has[dir]) {
            q.n[k] = (d.ie[dir] - d.is[dir]) / 2 + 1;
            q.st[k] = g.ds[dir];
            idx0 += (long long)((d.is[dir] - g.io[dir]) / 2) * g.ds[dir];
          }
          else {
            q.n[k] = 1;
            q.st[k] = 0;
          }
        }
        q.idx0 = idx0;
        q.avg1 = g.dev_offset(d.avg1);
        q.avg2 = g.dev_offset(d.avg2);
        if (q.avg1 == ((long long)1 << 62) || q.avg2 == ((long long)1 << 62))
          return fail(MEEPCU_EINVAL, "dft avg offsets are not sums of +-strides");
        for (int k = 0; k < 3; ++k) { q.s0[k] = d.s0[k]; q.s1[k] = d.s1[k]; q.e0[k] = d.e0[k]; q.e1[k] = d.e1[k]; }
        q.dV0 = d.dV0; q.dV1 = d.dV1;
        q.include_dV = d.include_dV_and_interp_weights;
        q.sqrt_dV = d.sqrt_dV_and_interp_weights;
        q.Nomega = d.Nomega;
        q.phase_group = grp;
        q.N = (long long)D.N;
        const int pidx = (int)h_dft.size();
        h_dft.push_back(q);
        dft_decim.push_back(d.decimation_factor);
        dft_is_h.push_back(is_h);
        const long long ne = q.N * q.Nomega;
        const long long per = std::max<long long>(1, std::min<long long>(kDftMaxPts, kDftElemsPerCta / std::max(1, d.Nomega)));
        for (long long p0 = 0; p0 < q.N; p0 += per) {
          DftWork w = {pidx, (int)std::min<long long>(per, q.N - p0), p0};
          work.push_back(w);
        }
        alg_bytes += 4LL * ne * (long long)sizeof(T) / d.decimation_factor;
        kc_bytes[KC_DFT] += 4LL * ne * (long long)sizeof(T) / d.decimation_factor;
        alg_bytes_sv += 4LL * ne * (long long)sizeof(T) / d.decimation_factor;
        kc_bytes_sv[KC_DFT] += 4LL * ne * (long long)sizeof(T) / d.decimation_factor;
      }
    }
    RC(upload(h_dft, &d_dft));
    RC(upload(work, &d_dft_work));
    n_dft_work = (int)work.size();
    RC(upload(phase_off, &d_phase_off));

    tm_mark("dft");
    // ---- 8. the lazy NaN test (step.cpp:138-139): table of the field arrays + a flag the device can raise ----
    {
      std::vector<FinEntry> fin;
      for (std::map<const char *, ArrayRec>::iterator it = arrays.begin(); it != arrays.end(); ++it)
        if (it->second.is_field) {
          FinEntry e = {it->second.dev, (long long)it->second.n_dev};
          fin.push_back(e);
          fin_max_n = std::max(fin_max_n, e.n);
        }
      n_fin = (int)fin.size();
      RC(upload(fin, &d_fin));
      CU(cudaHostAlloc((void **)&h_nonfinite, 64, cudaHostAllocMapped));
      memset(h_nonfinite, 0, 64);
      CU(cudaHostGetDevicePointer((void **)&d_hnonfinite, h_nonfinite, 0));
      finite_check = !getenv("MEEPCU_NO_FINITE_CHECK");
    }
    CU(cudaStreamSynchronize(st));
    finalized = true;
    return MEEPCU_OK;
  }

  // parameter blocks + tiles of k_ml_levels / k_ml_pol for one multilevel medium of one chunk (constants evaluated in
  // realnum exactly like multilevel-atom.cpp:244,265-270,297,332-336)
  int build_ml(const ChunkH &ch, const MlH &M, int ncmp, long long nc, std::vector<Tile> &lev_tiles, std::vector<Tile> &pol_tiles) {
    const Geom &g = ch.g;
    const double pi = 3.141592653589793238462643383276;
    const int L = M.L, Tn = M.Tn;
    const T dt = (T)cfg.dt;
    const T dt2 = 0.5 * dt;
    std::vector<T> Gamma(M.Gamma.begin(), M.Gamma.end()), alpha(M.alpha.begin(), M.alpha.end()), omega(M.omega.begin(), M.omega.end()),
        gamma(M.gamma.begin(), M.gamma.end()), sigmat(M.sigmat.begin(), M.sigmat.end());
    long long hax[3];
    for (int ax = 0; ax < 3; ++ax) hax[ax] = g.dir_of_axis[ax] < 0 ? 0 : g.hs[g.dir_of_axis[ax]];
    T *Ndev = dev(M.N);
    if (!Ndev) return fail(MEEPCU_EINVAL, "multilevel medium: N not registered");
    const int plane = g.nax[1] * g.pitch;
    const int seg = g.nax[0] <= 48 ? g.nax[0] : 32;
    { // ---- the populations ----
      MlLevelParams<T> Q;
      memset(&Q, 0, sizeof Q);
      Q.nA = g.nax[0]; Q.nB = g.nax[1]; Q.nC = g.nax[2]; Q.pitch = g.pitch;
      Q.sA = (long long)g.pitch * g.nax[1]; Q.sB = g.pitch;
      Q.hA = hax[0]; Q.hB = hax[1]; Q.hC = hax[2];
      for (int ax = 0; ax < 3; ++ax) { // LOOP_OVER_VOL_OWNED(gv, Centered, i): num points per direction, from index 0
        const int d = g.dir_of_axis[ax];
        Q.hi[ax] = d < 0 ? 0 : g.num[d] - 1;
      }
      Q.L = L; Q.Tn = Tn;
      Q.N = Ndev;
      int ndot = 0;
      for (int c = 0; c < 20; ++c) {
        if (M.P[c][0].empty()) continue;
        if (ndot == 3) return fail(MEEPCU_EINVAL, "bug in meep: too many polarization components");
        long long o1 = 0, o2 = 0; // grid_volume::yee2cent_offsets (vec.cpp:333-344)
        for (int dd = 0; dd < 3; ++dd)
          if (g.has[dd] && !g.yee_shift(c, dd)) {
            if (o1) o2 = g.ds[dd];
            else o1 = g.ds[dd];
          }
        Q.o1[ndot] = o1; Q.o2[ndot] = o2;
        for (int cmp = 0; cmp < ncmp; ++cmp) {
          const void *wh = ch.d.f_w[c][cmp] ? ch.d.f_w[c][cmp] : ch.d.f[c][cmp]; // update_pols.cpp:43-44
          Q.w[ndot][cmp] = dev(wh);
          Q.wp[ndot][cmp] = dev(ch.w_prev[c][cmp]);
          if (Q.w[ndot][cmp] && !Q.wp[ndot][cmp]) return fail(MEEPCU_EINVAL, "multilevel medium: W_prev missing");
          if (cmp == 0 && !Q.w[ndot][0]) return fail(MEEPCU_EINVAL, "multilevel medium: W missing");
          if (!M.P[c][cmp].empty())
            for (int t = 0; t < Tn; ++t) {
              Q.p[ndot][cmp][t] = dev(M.P[c][cmp][t]);
              Q.pp[ndot][cmp][t] = dev(M.P_prev[c][cmp][t]);
              if (!Q.p[ndot][cmp][t] || !Q.pp[ndot][cmp][t]) return fail(MEEPCU_EINVAL, "multilevel medium: P not registered");
            }
        }
        ++ndot;
      }
      Q.ndot = ndot;
      for (int l1 = 0; l1 < L; ++l1)
        for (int l2 = 0; l2 < L; ++l2) {
          Q.A[l1 * L + l2] = ((l1 == l2) - Gamma[l1 * L + l2] * dt2);
          Q.Ginv[l1 * L + l2] = (T)M.GammaInv[l1 * L + l2];
        }
      for (int l = 0; l < L; ++l)
        for (int t = 0; t < Tn; ++t) Q.alpha[l * Tn + t] = alpha[l * Tn + t];
      for (int t = 0; t < Tn; ++t) Q.gperpdt[t] = gamma[t] * pi * dt;
      const int qidx = (int)h_mllev.size();
      h_mllev.push_back(Q);
      for (int a0 = 0; a0 < g.nax[0]; a0 += seg)
        for (int p0 = 0; p0 < plane; p0 += kThreads) {
          Tile t = {qidx, a0, std::min(a0 + seg, g.nax[0]), p0};
          lev_tiles.push_back(t);
        }
      const long long v = 2LL * L + (long long)ndot * ncmp * (2 + 2 * Tn); // R,W N; R W, W_prev, P, P_prev of every transition
      alg_bytes += v * nc * (long long)sizeof(T);
      kc_bytes[KC_POL_GENERAL] += v * nc * (long long)sizeof(T);
      alg_bytes_sv += v * nc * (long long)sizeof(T);
      kc_bytes_sv[KC_POL_GENERAL] += v * nc * (long long)sizeof(T);
    }
    // ---- each P is updated as a damped harmonic oscillator (:331-373) ----
    for (int t = 0; t < Tn; ++t) {
      const T omega2pi = 2 * pi * omega[t], g2pi = gamma[t] * 2 * pi, gperp = gamma[t] * pi;
      const T omega0dtsqrCorrected = omega2pi * omega2pi * dt * dt + gperp * gperp * dt * dt;
      const T gamma1inv = 1 / (1 + g2pi * dt2), gamma1 = (1 - g2pi * dt2);
      const T dtsqr = dt * dt;
      int lp = -1, lm = -1;
      for (int l = 0; l < L; ++l) {
        if (alpha[l * Tn + t] > 0) lp = l;
        if (alpha[l * Tn + t] < 0) lm = l;
      }
      if (lp < 0 || lm < 0) return fail(MEEPCU_EINVAL, "invalid alpha array for transition %d", t);
      for (int c = 0; c < 20; ++c)
        for (int cmp = 0; cmp < ncmp; ++cmp) {
          if (M.P[c][cmp].empty()) continue;
          const int d0 = comp_dir(c);
          if (d0 >= MEEPCU_R) continue;
          const void *wh = ch.d.f_w[c][cmp] ? ch.d.f_w[c][cmp] : ch.d.f[c][cmp];
          if (!wh || !M.sigma[c][d0]) continue; // :352
          const int d1 = (d0 + 1) % 3, d2 = (d0 + 2) % 3;
          const int c1 = dir_comp(MEEPCU_E_STUFF, d1), c2 = dir_comp(MEEPCU_E_STUFF, d2);
          const bool w1 = g.has_field(c1) && (ch.d.f_w[c1][cmp] || ch.d.f[c1][cmp]);
          const bool w2 = g.has_field(c2) && (ch.d.f_w[c2][cmp] || ch.d.f[c2][cmp]);
          if ((w1 && M.sigma[c][d1]) || (w2 && M.sigma[c][d2]))
            return fail(MEEPCU_EUNSUPPORTED, "nondiagonal saturable gain is not yet supported"); // the reference's words (:362)
          MlPolParams<T> Q;
          memset(&Q, 0, sizeof Q);
          Q.nA = g.nax[0]; Q.nB = g.nax[1]; Q.nC = g.nax[2]; Q.pitch = g.pitch;
          Q.sA = (long long)g.pitch * g.nax[1]; Q.sB = g.pitch;
          Q.hA = hax[0]; Q.hB = hax[1]; Q.hC = hax[2];
          for (int ax = 0; ax < 3; ++ax) {
            const int d = g.dir_of_axis[ax];
            Q.lo[ax] = d < 0 ? 0 : 1 - g.yee_shift(c, d);
            Q.hi[ax] = d < 0 ? 0 : g.num[d] - g.yee_shift(c, d);
          }
          Q.L = L; Q.lp = lp; Q.lm = lm;
          Q.p = dev(M.P[c][cmp][t]); Q.pp = dev(M.P_prev[c][cmp][t]);
          Q.s = dev(M.sigma[c][d0]); Q.w = dev(wh); Q.N = Ndev;
          if (!Q.p || !Q.pp || !Q.s || !Q.w) return fail(MEEPCU_EINVAL, "multilevel medium: array not registered");
          long long oa = 0, ob = 0; // cent2yee_offsets = -yee2cent_offsets (vec.cpp:347-351), in elements of N
          for (int dd = 0; dd < 3; ++dd)
            if (g.has[dd] && !g.yee_shift(c, dd)) {
              if (oa) ob = g.hs[dd];
              else oa = g.hs[dd];
            }
          Q.oa = -oa * L; Q.ob = -ob * L;
          Q.gamma1inv = gamma1inv; Q.gamma1 = gamma1; Q.two_minus = (2 - omega0dtsqrCorrected); Q.dtsqr = dtsqr;
          Q.st = sigmat[5 * t + d0];
          const int qidx = (int)h_mlpol.size();
          h_mlpol.push_back(Q);
          for (int a0 = 0; a0 < g.nax[0]; a0 += seg)
            for (int p0 = 0; p0 < plane; p0 += kThreads) {
              Tile tl = {qidx, a0, std::min(a0 + seg, g.nax[0]), p0};
              pol_tiles.push_back(tl);
            }
          const long long v = 5 + 2LL * L; // R,W p, R,W pp, R sigma (+ the levels around the point; W counted with E)
          alg_bytes += v * nc * (long long)sizeof(T);
          kc_bytes[KC_POL_GENERAL] += v * nc * (long long)sizeof(T);
          alg_bytes_sv += v * nc * (long long)sizeof(T);
          kc_bytes_sv[KC_POL_GENERAL] += v * nc * (long long)sizeof(T);
        }
    }
    return MEEPCU_OK;
  }

  // segments of k_pol_noise, per half, in the reference's draw order (susceptibility.cpp:327-336)
  int build_noise() {
    const double pi = 3.141592653589793238462643383276;
    const T dt = (T)cfg.dt;
    for (int half = 0; half < 2; ++half) {
      const int ft = half ? MEEPCU_E_STUFF : MEEPCU_H_STUFF;
      std::vector<NoiseSeg<T> > segs;
      long long total = 0;
      for (size_t ic = 0; ic < chunks.size(); ++ic) {
        const ChunkH &ch = chunks[ic];
        const Geom &g = ch.g;
        for (size_t ip = 0; ip < ch.pol[ft].size(); ++ip) {
          const meepcu_susceptibility_desc &sd = ch.pol[ft][ip].d;
          if (ch.pol[ft][ip].gyro || ch.pol[ft][ip].ml >= 0 || sd.noise_amp == 0) continue;
          const T noise_amp = (T)sd.noise_amp, omega_0 = (T)sd.omega_0, gamma = (T)sd.gamma;
          const T g2pi = gamma * 2 * pi;
          const T w2pi = omega_0 * 2 * pi;
          const T amp = w2pi * noise_amp * std::sqrt(g2pi) * dt * dt / (1 + g2pi * dt / 2); // :325
          for (int c = 0; c < 20; ++c)
            for (int cmp = 0; cmp < 2; ++cmp) {
              if (!sd.P[c][cmp] || comp_dir(c) >= MEEPCU_R || !sd.sigma[c][comp_dir(c)]) continue;
              NoiseSeg<T> S;
              memset(&S, 0, sizeof S);
              S.p = dev(sd.P[c][cmp]);
              S.s = dev(sd.sigma[c][comp_dir(c)]);
              if (!S.p || !S.s) return fail(MEEPCU_EINVAL, "noisy medium: array not registered");
              S.amp = amp;
              S.first = total;
              long long m = 1;
              for (int dd = 0; dd < 3; ++dd) {
                S.n[dd] = g.has[dd] ? g.num[dd] : 1;
                S.ds[dd] = g.has[dd] ? g.ds[dd] : 0;
                if (g.has[dd]) S.base += (long long)(1 - g.yee_shift(c, dd)) * g.ds[dd]; // little_owned_corner(c)
                m *= S.n[dd];
              }
              total += m;
              segs.push_back(S);
            }
        }
      }
      n_noise_seg[half] = (int)segs.size();
      noise_total[half] = total;
      RC(upload(segs, &d_noise_seg[half]));
    }
    return MEEPCU_OK;
  }

  // parameter blocks + tiles of k_pol_gyro for one gyrotropic susceptibility of one chunk
  // (constants evaluated in realnum exactly like susceptibility.cpp:449-478,531-546)
  int build_gyro(const ChunkH &ch, int ic, const meepcu_gyrotropic_desc &gd_, int ft, int half, int ncmp, long long nc,
                 std::vector<Tile> &tiles_out) {
    const Geom &g = ch.g;
    const double pi = 3.141592653589793238462643383276;
    const T dt = (T)cfg.dt;
    const T omega_0 = (T)gd_.omega_0, gamma = (T)gd_.gamma, alpha = (T)gd_.alpha;
    T gyro_tensor[3][3];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) gyro_tensor[a][b] = (T)gd_.gyro_tensor[a][b];
    const T omega2pidt = 2 * pi * omega_0 * dt;
    const T g2pidt = 2 * pi * gamma * dt;
    const int X = 0, Y = 1, Z = 2;
    const bool sat = gd_.model == MEEPCU_GYROTROPIC_SATURATED;
    T omega0dtsqr = 0, gamma1 = 0, diag = 0, pt = 0, dt2pi = 0, gd, gx, gy, gz;
    if (!sat) {
      omega0dtsqr = omega2pidt * omega2pidt;
      gamma1 = (1 - g2pidt / 2);
      diag = 2 - (gd_.model == MEEPCU_GYROTROPIC_DRUDE ? 0 : omega0dtsqr);
      pt = pi * dt;
      gd = (1 + g2pidt / 2);
      gx = pt * gyro_tensor[Y][Z];
      gy = pt * gyro_tensor[Z][X];
      gz = pt * gyro_tensor[X][Y];
    }
    else {
      dt2pi = 2 * pi * dt;
      gd = 0.5;
      gx = -0.5 * alpha * gyro_tensor[Y][Z];
      gy = -0.5 * alpha * gyro_tensor[Z][X];
      gz = -0.5 * alpha * gyro_tensor[X][Y];
    }
    const T invdet = 1.0 / gd / (gd * gd + gx * gx + gy * gy + gz * gz);
    const T inv[3][3] = {{invdet * (gd * gd + gx * gx), invdet * (gx * gy + gd * gz), invdet * (gx * gz - gd * gy)},
                         {invdet * (gy * gx - gd * gz), invdet * (gd * gd + gy * gy), invdet * (gy * gz + gd * gx)},
                         {invdet * (gz * gx + gd * gy), invdet * (gz * gy - gd * gx), invdet * (gd * gd + gz * gz)}};
    const int sgn = ft == MEEPCU_H_STUFF ? -1 : +1;
    for (int cmp = 0; cmp < ncmp; ++cmp)
      for (int k = 0; k < 3; ++k) {
        const int c = dir_comp(ft, k);
        if (!gd_.P[c][cmp][0]) continue;
        const int d0 = k, d1 = (k + 1) % 3, d2 = (k + 2) % 3;
        const int c1 = dir_comp(ft, d1), c2 = dir_comp(ft, d2);
        const void *w0h = ch.d.f_w[c][cmp] ? ch.d.f_w[c][cmp] : ch.d.f[c][cmp];
        const void *w1h = ch.d.f_w[c1][cmp] ? ch.d.f_w[c1][cmp] : ch.d.f[c1][cmp];
        const void *w2h = ch.d.f_w[c2][cmp] ? ch.d.f_w[c2][cmp] : ch.d.f[c2][cmp];
        PolGyroParams<T> Q;
        memset(&Q, 0, sizeof Q);
        Q.nA = g.nax[0]; Q.nB = g.nax[1]; Q.nC = g.nax[2]; Q.pitch = g.pitch;
        Q.sA = (long long)g.pitch * g.nax[1]; Q.sB = g.pitch;
        for (int ax = 0; ax < 3; ++ax) {
          const int d = g.dir_of_axis[ax];
          Q.lo[ax] = d < 0 ? 0 : 1 - g.yee_shift(c, d);
          Q.hi[ax] = d < 0 ? 0 : g.num[d] - g.yee_shift(c, d);
        }
        Q.p0 = dev(gd_.P[c][cmp][d0]); Q.pp0 = dev(gd_.P_prev[c][cmp][d0]);
        Q.p1 = dev(gd_.P[c][cmp][d1]); Q.pp1 = dev(gd_.P_prev[c][cmp][d1]);
        Q.p2 = dev(gd_.P[c][cmp][d2]); Q.pp2 = dev(gd_.P_prev[c][cmp][d2]);
        Q.s = dev(gd_.sigma[c][d0]);
        Q.w0 = dev(w0h); Q.w1 = dev(w1h); Q.w2 = dev(w2h);
        if (!Q.w0 || !Q.s || !Q.p0 || !Q.pp0 || !Q.p1 || !Q.pp1 || !Q.p2 || !Q.pp2)
          return fail(MEEPCU_EINVAL, "gyrotropic media require 3D Cartesian fields");
        if (gd_.sigma[c][d1] || gd_.sigma[c][d2])
          return fail(MEEPCU_EINVAL, "gyrotropic media do not support anisotropic sigma");
        Q.is = g.ds[d0] * sgn; Q.is1 = g.ds[d1] * sgn; Q.is2 = g.ds[d2] * sgn;
        Q.saturated = sat;
        Q.g01 = gyro_tensor[d0][d1]; Q.g02 = gyro_tensor[d0][d2];
        Q.g10 = gyro_tensor[d1][d0]; Q.g12 = gyro_tensor[d1][d2];
        Q.g20 = gyro_tensor[d2][d0]; Q.g21 = gyro_tensor[d2][d1];
        const int dd[3] = {d0, d1, d2};
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b) Q.inv[a][b] = inv[dd[a]][dd[b]];
        Q.diag = diag; Q.gamma1 = gamma1; Q.omega0dtsqr = omega0dtsqr; Q.pt = pt;
        Q.omega2pidt = omega2pidt; Q.g2pidt = g2pidt; Q.alpha = alpha; Q.dt2pi = dt2pi;
        const int qidx = (int)h_polgyro.size();
        h_polgyro.push_back(Q);
        const int plane = g.nax[1] * g.pitch;
        const int seg = g.nax[0] <= 48 ? g.nax[0] : 32;
        for (int a0 = 0; a0 < g.nax[0]; a0 += seg)
          for (int p0 = 0; p0 < plane; p0 += kThreads) {
            Tile t = {qidx, a0, std::min(a0 + seg, g.nax[0]), p0};
            tiles_out.push_back(t);
          }
        alg_bytes += 13LL * nc * (long long)sizeof(T); // R,W six P arrays + R sigma (W counted with E)
        kc_bytes[KC_POL_GENERAL] += 13LL * nc * (long long)sizeof(T);
        alg_bytes_sv += 13LL * nc * (long long)sizeof(T);
        kc_bytes_sv[KC_POL_GENERAL] += 13LL * nc * (long long)sizeof(T);
      }
    (void)ic;
    (void)half;
    return MEEPCU_OK;
  }

  bool aniso_now = false; // set around fill_pols calls for (chunk, half) pairs in aniso_chunk_half
  int fill_pols(const ChunkH &ch, int ft, int ec, int k, int cmp, PolParams<T> *out) {
    for (size_t ip = 0; ip < ch.pol[ft].size(); ++ip) {
      const meepcu_susceptibility_desc &sd = ch.pol[ft][ip].d;
      PolParams<T> &q = out[ip];
      memset(&q, 0, sizeof q);
      q.p = dev(sd.P[ec][cmp]);
      q.pp = dev(sd.P_prev[ec][cmp]);
      // update_P needs w && s (susceptibility.cpp:201); w = f_w ?: f exists since f[ec][cmp] does
      q.s = q.p ? dev(sd.sigma[ec][k]) : 0;
      if (aniso_now) q.s = 0; // updated by k_pol_general after the whole chunk has its new W
      // susceptibility.cpp:191-194, evaluated in realnum exactly like the reference
      const double pi = 3.141592653589793238462643383276;
      const T dt = (T)cfg.dt;
      const T omega_0 = (T)sd.omega_0, gamma = (T)sd.gamma;
      const T omega2pi = 2 * pi * omega_0, g2pi = gamma * 2 * pi;
      const T omega0dtsqr = omega2pi * omega2pi * dt * dt;
      q.gamma1inv = 1 / (1 + g2pi * dt / 2);
      q.gamma1 = (1 - g2pi * dt / 2);
      q.omega0dtsqr = omega0dtsqr;
      q.omega0dtsqr_denom = sd.no_omega_0_denominator ? 0 : omega0dtsqr;
    }
    return MEEPCU_OK;
  }

  // ------------------------------------------------------------------------------------------
  int prof_begin(int kc) {
    if (!profiling) return MEEPCU_OK;
    if (ev_used + 2 > ev_pool.size()) {
      for (int k = 0; k < 2; ++k) {
        cudaEvent_t e;
        CU(cudaEventCreate(&e));
        ev_pool.push_back(e);
      }
    }
    ev_class.push_back(kc);
    CU(cudaEventRecord(ev_pool[ev_used], st));
    return MEEPCU_OK;
  }
  int prof_end() {
    if (!profiling) return MEEPCU_OK;
    CU(cudaEventRecord(ev_pool[ev_used + 1], st));
    ev_used += 2;
    return MEEPCU_OK;
  }
  int prof_collect() { // after a stream synchronize
    for (size_t k = 0; k < ev_used / 2; ++k) {
      float ms = 0;
      CU(cudaEventElapsedTime(&ms, ev_pool[2 * k], ev_pool[2 * k + 1]));
      kc_ms[ev_class[k]] += ms;
      kc_launches[ev_class[k]]++;
    }
    ev_used = 0;
    ev_class.clear();
    return MEEPCU_OK;
  }
  int set_profiling(int on) {
    profiling = (on & 1) != 0;
    if (finalized) { // bit 1: count Newton-Raphson solves / iterations (kernels.cuh, eh_point; atomics: not while timing)
      CU(cudaSetDevice(cfg.device));
      CU(cudaStreamSynchronize(st));
      const int z[6] = {(on & 2) != 0, 0, 0, 0, 0, 0};
      if (on & 2) CU(cudaMemcpy(d_flags + 2, z, sizeof z, cudaMemcpyHostToDevice));
      else CU(cudaMemcpy(d_flags + 2, z, sizeof(int), cudaMemcpyHostToDevice));
    }
    memset(kc_ms, 0, sizeof kc_ms);
    memset(kc_launches, 0, sizeof kc_launches);
    return MEEPCU_OK;
  }
  int get_kernel_times(meepcu_kernel_time *out, int cap, int *n) {
    static const char *names[KC_COUNT] = {"k_halfg<B/H,pml>", "k_halfg<B/H,pml+pol/cond>", "k_halfg<D/E,pml>",
                                          "k_halfg<D/E,pml+pol/cond>", "k_conn", "k_fmp", "k_eh_general", "k_dft",
                                          "k_curl3d_v<B/H>", "k_curl3d_v<D/E>", "k_curl3d (source/metal tiles)",
                                          "k_mix<B/H> (interior + pml tiles)", "k_mix<D/E> (interior + pml tiles)",
                                          "k_pol_general (anisotropic sigma)"};
    int m = 0;
    for (int k = 0; k < KC_COUNT && m < cap; ++k) {
      memset(&out[m], 0, sizeof out[m]);
      snprintf(out[m].name, sizeof out[m].name, "%s", names[k]);
      out[m].launches = kc_launches[k];
      out[m].ms = kc_ms[k];
      out[m].algorithmic_bytes = kc_bytes[k];
      ++m;
    }
    *n = m;
    return MEEPCU_OK;
  }

  int launch_conn(int ft) {
    RC(launch_conn_local(ft));
    return launch_remote(ft, st);
  }
  int launch_conn_local(int ft) {
    const ConnTable<T> &a = d_conn[ft][0];
    ConnTable<T> none;
    memset(&none, 0, sizeof none);
    if (a.n) {
      const int grid = (int)std::min<long long>((a.n / 2 + kThreads - 1) / kThreads, 148 * 16);
      RC(prof_begin(KC_CONN));
      k_conn<T><<<grid, kThreads, 0, st>>>(a, none, none);
      RC(prof_end());
      launches++;
    }
    if (n_runs[ft]) {
      const int grid = (n_runs[ft] + 7) / 8;
      RC(prof_begin(KC_CONN));
      k_conn_runs<T><<<grid, kThreads, 0, st>>>(d_runs[ft], n_runs[ft]);
      RC(prof_end());
      launches++;
    }
    return MEEPCU_OK;
  }

  // The curl / PML kernels of one half step over part of every tile list.
  //   part 0: all tiles;  1: boundary tiles only (the first nb_* of each list);  2: the rest
  // sm: stream of the interior tiles; sside: stream of the PML / special tiles (== sm: no fork)
  const double *cur_prev = 0; // previous step's sample row (virtual-E dipoles), set by the callers of launch_half
  int cur_vmode = 0;          // 1: B/H half of virtual chunks reads D (x chi1inv) instead of E
  int launch_compute(int half, const double *smp, int part, cudaStream_t sm, cudaStream_t sside) {
    const int vmode = half == 0 ? cur_vmode : 0;
    const double *smp_prev = cur_prev;
    // The interior tiles (k_curl3d) and the PML / general tiles (k_halfg) of one half step touch
    // disjoint chunks, so they run side by side on two streams; everything else is ordered after
    // both.  While profiling they are serialised so that each kernel's event time is its own.
#define MCU_RANGE(total, nb, off, n)                                                               \
    const int off = part == 2 ? (nb) : 0;                                                         \
    const int n = part == 0 ? (total) : (part == 1 ? (nb) : (total) - (nb));
    MCU_RANGE(n_c3dv_tiles[half], nb_c3dv_tiles[half], o_cv, n_cv)
    MCU_RANGE(n_c3d_tiles[half], nb_c3d_tiles[half], o_c, n_c)
    MCU_RANGE(n_gv_tiles[half][0], nb_gv_tiles[half][0], o_gv0, n_gv0)
    MCU_RANGE(n_gv_tiles[half][1], nb_gv_tiles[half][1], o_gv1, n_gv1)
    MCU_RANGE(n_tiles[half][0], nb_tiles[half][0], o_t0, n_t0)
    MCU_RANGE(n_tiles[half][1], nb_tiles[half][1], o_t1, n_t1)
    MCU_RANGE(n_tiles[half][2], nb_tiles[half][2], o_t2, n_t2)
    MCU_RANGE(n_tiles[half][3], nb_tiles[half][3], o_t3, n_t3)
    MCU_RANGE(n_mix_tiles[half], nb_mix_tiles[half], o_mx, n_mx)
#undef MCU_RANGE
    const bool have_g = n_t0 || n_t1 || n_t2 || n_t3 || n_gv0 || n_gv1;
    const bool fork = (have_g || (n_c && (n_cv || n_mx))) && (n_c || n_cv || n_mx) && !profiling && sside != sm;
    cudaStream_t sg = fork ? sside : sm;
    if (fork) {
      CU(cudaEventRecord(ev_fork, sm));
      CU(cudaStreamWaitEvent(sside, ev_fork, 0));
    }
    // the bandwidth-bound interior tiles are queued first; the latency-bound PML tiles (high-priority
    // stream) take SM slots as they free up, so the two kinds of CTAs share every SM
    if (n_mx) {
      RC(prof_begin(half ? KC_MIX_D : KC_MIX_B));
      if (half)
        k_mix<T, true><<<n_mx, kThreads, 0, sm>>>(d_c3d, d_half, d_mix_tiles[half] + o_mx, 0);
      else if (vmode && virt_vu)
        k_mix<T, false, true><<<n_mx, kThreads, 0, sm>>>(d_c3d, d_half, d_mix_tiles[half] + o_mx, vmode);
      else
        k_mix<T, false><<<n_mx, kThreads, 0, sm>>>(d_c3d, d_half, d_mix_tiles[half] + o_mx, vmode);
      RC(prof_end());
      launches++;
    }
    if (n_cv) {
      RC(prof_begin(half ? KC_CURL3D_D : KC_CURL3D_B));
      if (half)
        k_curl3d_v<T, true><<<n_cv, kThreads, 0, sm>>>(d_c3d, d_c3dv_tiles[half] + o_cv, 0);
      else if (vmode && virt_vu)
        k_curl3d_v<T, false, true><<<n_cv, kThreads, 0, sm>>>(d_c3d, d_c3dv_tiles[half] + o_cv, vmode);
      else
        k_curl3d_v<T, false><<<n_cv, kThreads, 0, sm>>>(d_c3d, d_c3dv_tiles[half] + o_cv, vmode);
      RC(prof_end());
      launches++;
    }
    // one profiled region per class: the vector kernel and the scalar kernel (tiles with source points,
    // conductivity, or rows that cannot be vectorised) of a class share its algorithmic bytes
    if (n_gv0 || n_t0) {
      RC(prof_begin(half ? KC_HALF_D_SIMPLE : KC_HALF_B_SIMPLE));
      if (n_gv0) {
        k_halfg_v<T, false><<<n_gv0, kThreads, 0, sg>>>(d_half, d_gv_tiles[half][0] + o_gv0);
        launches++;
      }
      if (n_t0) {
        k_halfg<T, false, false><<<n_t0, kThreads, 0, sg>>>(d_half, d_tiles[half][0] + o_t0, smp);
        launches++;
      }
      RC(prof_end());
    }
    if (n_gv1 || n_t1 || n_t2 || n_t3) {
      RC(prof_begin(half ? KC_HALF_D_GENERAL : KC_HALF_B_GENERAL));
      if (n_gv1) {
        k_halfg_v<T, true><<<n_gv1, kThreads, 0, sg>>>(d_half, d_gv_tiles[half][1] + o_gv1);
        launches++;
      }
      if (n_t1) {
        k_halfg<T, false, true><<<n_t1, kThreads, 0, sg>>>(d_half, d_tiles[half][1] + o_t1, smp);
        launches++;
      }
      if (n_t2) {
        k_halfg<T, true, true><<<n_t2, kThreads, 0, sg>>>(d_half, d_tiles[half][2] + o_t2, smp);
        launches++;
      }
      if (n_t3) {
        k_halfg<T, true, true, true><<<n_t3, kThreads, 0, sg>>>(d_half, d_tiles[half][3] + o_t3, smp);
        launches++;
      }
      RC(prof_end());
    }
    if (n_c) { // tiles with source points / metal planes, or unvectorisable pitch
      cudaStream_t ss = (n_cv || n_mx) ? sg : sm; // beside the vector kernel when there is one
      RC(prof_begin(KC_CURL3D_SPECIAL));
      if (half)
        k_curl3d<T, true><<<n_c, kThreads, 0, ss>>>(d_c3d, d_c3d_tiles[half] + o_c, smp, smp_prev, 0);
      else
        k_curl3d<T, false><<<n_c, kThreads, 0, ss>>>(d_c3d, d_c3d_tiles[half] + o_c, smp, smp_prev, vmode);
      RC(prof_end());
      launches++;
    }
    if (fork) {
      CU(cudaEventRecord(ev_join, sside));
      CU(cudaStreamWaitEvent(sm, ev_join, 0));
    }
    return MEEPCU_OK;
  }

  int launch_cyl_origin(int half) { // the r = 0 rules of step_db.cpp:292-457, after the curls of the half step
    if (!n_cyl_org[half]) return MEEPCU_OK;
    k_cyl_origin<T><<<dim3(4, (unsigned)std::min(n_cyl_org[half], 1024)), kThreads, 0, st>>>(d_cyl_org[half], n_cyl_org[half]);
    launches++;
    return MEEPCU_OK;
  }
  int launch_half(int half, const double *smp) {
    const int ftd = half ? MEEPCU_D_STUFF : MEEPCU_B_STUFF;
    const int fte = half ? MEEPCU_E_STUFF : MEEPCU_H_STUFF;
    const int ftp = half ? MEEPCU_PE_STUFF : MEEPCU_PH_STUFF;
    const int ftw = half ? MEEPCU_WE_STUFF : MEEPCU_WH_STUFF;
    if (n_cyl_rd[half]) { // cylindrical: f_rderiv_int of this half's phi operand (step_db.cpp:93-118), before the curls
      k_cyl_rderiv<T><<<dim3(4, (unsigned)std::min(n_cyl_rd[half], 1024)), kThreads, 0, st>>>(d_cyl_rd[half], n_cyl_rd[half]);
      launches++;
    }
    if (half == 1 && n_wprev && !in_mag_sync) { // f_w_prev = W of this time step, before the E update (update_eh.cpp:221-226)
      dim3 grid((unsigned)std::min<long long>((wprev_max + kThreads * 4 - 1) / (kThreads * 4), 148 * 4), (unsigned)n_wprev);
      k_copy_arrays<T><<<grid, kThreads, 0, st>>>(d_wprev);
      launches++;
    }
    if (overlap_half[half] && !profiling) {
      // boundary tiles -> [halo exchange on st_comm  ||  interior tiles + local halo copies] -> join.
      // The exchange writes only not-owned points of E/H (W), which no kernel of this half reads.
      // The boundary tiles have their own high-priority stream: they are scheduled first but do not
      // hold back the interior tiles (both sets only depend on the previous half step).
      CU(cudaEventRecord(ev_bstart, st));
      CU(cudaStreamWaitEvent(st_bnd, ev_bstart, 0));
      RC(launch_compute(half, smp, 1, st_bnd, st_bnd));
      CU(cudaEventRecord(ev_bdone, st_bnd));
      CU(cudaStreamWaitEvent(st_comm, ev_bdone, 0));
      RC(launch_remote(ftw, st_comm));
      RC(launch_remote(fte, st_comm));
      CU(cudaEventRecord(ev_cdone, st_comm));
      RC(launch_compute(half, smp, 2, st, st2));
      CU(cudaStreamWaitEvent(st, ev_bdone, 0)); // the local halo copies read boundary tiles' results too
      RC(launch_cyl_origin(half));
      RC(launch_conn_local(ftw));
      RC(launch_conn_local(fte));
      CU(cudaStreamWaitEvent(st, ev_cdone, 0));
      return MEEPCU_OK;
    }
    RC(launch_compute(half, smp, 0, st, st2));
    RC(launch_cyl_origin(half));
    if (eager_db[half]) RC(launch_conn(ftd)); // step.cpp:72,104
    if (general_half[half]) {
      if (n_fmp[half]) {
        dim3 grid(148 * 2, (unsigned)std::min(n_fmp[half], 64));
        RC(prof_begin(KC_FMP));
        k_fmp<T><<<grid, kThreads, 0, st>>>(d_fmp[half], n_fmp[half], smp);
        RC(prof_end());
        launches++;
        if (fmp_has_src[half]) {
          k_fmp_src<T><<<(n_fmp[half] + 63) / 64, 64, 0, st>>>(d_fmp[half], n_fmp[half], smp);
          launches++;
        }
      }
      if (n_dip[half]) { // dipole points: D := D - dipoles in place (originals saved), put back after the E update
        k_dip_patch<T><<<(n_dip[half] + 63) / 64, 64, 0, st>>>(d_dip[half], n_dip[half], smp, 0);
        launches++;
      }
      if (n_eh_tiles[half]) { // (the flag is global: this rank may own no such chunk)
        RC(prof_begin(KC_EH_GENERAL));
        // resident CTAs per SM (= registers per thread): MEEPCU_EH_CTAS = 2 / 3 / 4 / 5 select builds with 128 / 80 / 64 /
        // 48 registers (measurements at MCU_EH_CTAS in kernels.cuh)
        // (the Pade update has no iteration to keep in registers: 5 CTAs, 2.55 vs 2.83 ms at 384^3; NR: 4, 4.34 vs 4.81)
        const int eh_ctas = getenv("MEEPCU_EH_CTAS") ? atoi(getenv("MEEPCU_EH_CTAS"))
                                                     : (cfg.nl_mode == MEEPCU_NL_UPSTREAM_PADE ? 5 : MCU_EH_CTAS);
        if (eh_ctas == 2) k_eh_general<T, 2><<<n_eh_tiles[half], kThreads, 0, st>>>(d_eh, d_eh_tiles[half]);
        else if (eh_ctas == 3) k_eh_general<T, 3><<<n_eh_tiles[half], kThreads, 0, st>>>(d_eh, d_eh_tiles[half]);
        else if (eh_ctas == 5) k_eh_general<T, 5><<<n_eh_tiles[half], kThreads, 0, st>>>(d_eh, d_eh_tiles[half]);
        else k_eh_general<T, 4><<<n_eh_tiles[half], kThreads, 0, st>>>(d_eh, d_eh_tiles[half]);
        RC(prof_end());
        launches++;
      }
      if (n_dip[half]) {
        k_dip_patch<T><<<(n_dip[half] + 63) / 64, 64, 0, st>>>(d_dip[half], n_dip[half], smp, 1);
        launches++;
      }
    }
    // synchronize_magnetic_fields runs neither step_boundaries(WH) nor update_pols(H) (energy_and_flux.cpp:153-160)
    if (!in_mag_sync) RC(launch_conn(ftw));  // step.cpp:81,113 (empty unless anisotropic sigma / gyrotropy)
    if (n_polgen_tiles[half] && !in_mag_sync) { // update_pols for anisotropic sigma, step.cpp:83-86,115-117
      RC(prof_begin(KC_POL_GENERAL));
      k_pol_general<T><<<n_polgen_tiles[half], kThreads, 0, st>>>(d_polgen, d_polgen_tiles[half]);
      RC(prof_end());
      launches++;
    }
    if (n_polgyro_tiles[half] && !in_mag_sync) { // gyrotropic update_P
      RC(prof_begin(KC_POL_GENERAL));
      k_pol_gyro<T><<<n_polgyro_tiles[half], kThreads, 0, st>>>(d_polgyro, d_polgyro_tiles[half]);
      RC(prof_end());
      launches++;
    }
    if (half == 1 && n_mllev_tiles && !in_mag_sync) { // multilevel atoms: populations, then every transition's P
      RC(prof_begin(KC_POL_GENERAL));
      k_ml_levels<T><<<n_mllev_tiles, kThreads, 0, st>>>(d_mllev, d_mllev_tiles);
      launches++;
      if (n_mlpol_tiles) {
        k_ml_pol<T><<<n_mlpol_tiles, kThreads, 0, st>>>(d_mlpol, d_mlpol_tiles);
        launches++;
      }
      RC(prof_end());
    }
    if (n_noise_seg[half] && !in_mag_sync) { // noisy Lorentzian media: the caller's Gaussians, after the Lorentzian update
      if (!cur_noise) return fail(MEEPCU_ESTATE, "noisy medium stepped without noise");
      const int grid = (int)std::min<long long>((noise_total[half] + kThreads - 1) / kThreads, 148 * 8);
      k_pol_noise<T><<<grid, kThreads, 0, st>>>(d_noise_seg[half], n_noise_seg[half], noise_total[half],
                                                cur_noise + (half ? noise_total[0] : 0));
      launches++;
    }
    if (n_zero_late[half]) {                 // zero_metal(E) of step_boundaries(E_stuff), step.cpp:245-246
      k_zero<T><<<(int)std::min<long long>((n_zero_late[half] + kThreads - 1) / kThreads, 148 * 8), kThreads, 0, st>>>(
          d_zero_late[half], n_zero_late[half]);
      launches++;
    }
    if (eager_p[half]) RC(launch_conn(ftp)); // step.cpp:87,118
    RC(launch_conn(fte));                    // step.cpp:91,122
    return MEEPCU_OK;
  }

  int collect_one(StepEv &e) {
    if (!e.pending) return MEEPCU_OK;
    CU(cudaEventSynchronize(e.e1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e.e0, e.e1));
    last_ms = ms;
    total_step_ms += ms;
    e.pending = false;
    return MEEPCU_OK;
  }
  int collect_timing() { // waits for every queued meepcu_step call
    for (size_t k = 0; k < step_ev.size(); ++k) RC(collect_one(step_ev[(step_ev_next + k) % step_ev.size()]));
    return MEEPCU_OK;
  }
  int nonfinite_seen() const { return h_nonfinite ? *(volatile int *)h_nonfinite : 0; }

  int step(int nsteps, int nsrc, const meepcu_src_sample *smp) {
    if (!finalized) return fail(MEEPCU_ESTATE, "meepcu_finalize not called");
    if (mag_synced) return fail(MEEPCU_ESTATE, "magnetic fields are synchronized: meepcu_restore_magnetic_fields first (step.cpp:36-41)");
    if (nsteps <= 0) return MEEPCU_OK;
    if (nonfinite_seen()) // raised by the sampled test that followed an earlier call (step.cpp:138-139)
      return fail(MEEPCU_ENONFINITE, "simulation fields are NaN or Inf");
    for (size_t ic = 0; ic < chunks.size(); ++ic)
      for (size_t is = 0; is < chunks[ic].src.size(); ++is)
        if (chunks[ic].src[is].tid >= nsrc || chunks[ic].src[is].tid < 0 || (nsrc > 0 && !smp))
          return fail(MEEPCU_EINVAL, "source refers to src_time %d but only %d samples given",
                      chunks[ic].src[is].tid, nsrc);
    CU(cudaSetDevice(cfg.device));
    // per-step tables: src_time samples, DFT phases (dft.cpp:269-270, host libm like the
    // reference), DFT decimation flags (dft.cpp:259) -- through the ring, no synchronisation
    const size_t ns = (size_t)std::max(nsrc, 1) * 8;
    char *p_smp = 0, *p_ph = 0, *p_act = 0;
    RC(ring_reserve(ring_slot((size_t)(nsteps + 1) * ns * sizeof(double)) +
                    ring_slot((size_t)(noise_total[0] + noise_total[1]) * nsteps * sizeof(double)) +
                    ring_slot((size_t)2 * phase_row * nsteps * sizeof(T)) + ring_slot(h_dft.size() * nsteps)));
    {
      // row 0: the samples of the last step taken before this call (a virtual E holds ITS dipole_E), rows 1..: this call's
      std::vector<double> tab((size_t)(nsteps + 1) * ns, 0.0);
      if (last_row.size() == ns) memcpy(tab.data(), last_row.data(), ns * sizeof(double));
      if (nsrc > 0) memcpy(tab.data() + ns, smp, sizeof(meepcu_src_sample) * (size_t)nsrc * nsteps);
      RC(ring_put(tab.data(), tab.size() * sizeof(double), &p_smp));
      last_row.assign(tab.end() - ns, tab.end());
    }
    const size_t npst = (size_t)(noise_total[0] + noise_total[1]);
    char *p_noise = 0;
    if (npst) { // noisy Lorentzian media: the unit Gaussians meepcu_set_noise left for exactly these steps
      if (noise_steps != nsteps)
        return fail(MEEPCU_ESTATE, "noisy Lorentzian media: meepcu_set_noise must provide the %d steps of this call (%d waiting)", nsteps, noise_steps);
      RC(ring_put(noise_host.data(), noise_host.size() * sizeof(double), &p_noise));
      noise_steps = 0;
    }
    const size_t ndft = h_dft.size();
    if (ndft) {
      std::vector<T> ph((size_t)2 * phase_row * nsteps);
      std::vector<unsigned char> act(ndft * nsteps);
      for (int s = 0; s < nsteps; ++s) {
        const int tt = t + s + 1; // fields::t after the increment of step.cpp:126
        const double timeE = tt * cfg.dt, timeH = tt * cfg.dt - 0.5 * cfg.dt; // dft.cpp:252
        for (size_t gI = 0; gI < group_omega.size(); ++gI) {
          const double time = group_is_h[gI] ? timeH : timeE;
          for (size_t i = 0; i < group_omega[gI].size(); ++i) {
            const std::complex<double> z = std::polar(1.0, group_omega[gI][i] * time) * group_scale[gI];
            ph[2 * ((size_t)s * phase_row + phase_off[gI] + i)] = (T)z.real();
            ph[2 * ((size_t)s * phase_row + phase_off[gI] + i) + 1] = (T)z.imag();
          }
        }
        for (size_t k = 0; k < ndft; ++k) act[(size_t)s * ndft + k] = (tt % dft_decim[k]) == 0;
      }
      RC(ring_put(ph.data(), ph.size() * sizeof(T), &p_ph));
      RC(ring_put(act.data(), act.size(), &p_act));
    }
    if (step_ev.empty()) {
      step_ev.resize(16);
      for (size_t k = 0; k < step_ev.size(); ++k) {
        CU(cudaEventCreate(&step_ev[k].e0));
        CU(cudaEventCreate(&step_ev[k].e1));
        step_ev[k].pending = false;
      }
    }
    StepEv &E = step_ev[step_ev_next];
    RC(collect_one(E)); // the oldest pair: long finished unless 16 calls are already queued
    step_ev_next = (step_ev_next + 1) % step_ev.size();
    CU(cudaEventRecord(E.e0, st));
    for (int s = 0; s < nsteps; ++s) {
      const double *sm = (const double *)p_smp + (size_t)(s + 1) * ns;
      cur_prev = (const double *)p_smp + (size_t)s * ns;
      cur_vmode = (!virt_comps.empty() && e_virtual) ? 1 : 0;
      cur_noise = npst ? (const double *)p_noise + (size_t)s * npst : 0;
      RC(launch_half(0, sm));
      RC(launch_half(1, sm));
      e_virtual = true; // the D/E half of a virtual chunk has not stored E: from now on it is read through D
      if (n_dft_work) {
        RC(prof_begin(KC_DFT));
        k_dft<T><<<n_dft_work, kThreads, 0, st>>>(d_dft, d_dft_work, (const T *)p_ph + (size_t)2 * phase_row * s,
                                                   d_phase_off, (const unsigned char *)p_act + (size_t)s * ndft);
        RC(prof_end());
        launches++;
      }
    }
    CU(cudaEventRecord(E.e1, st));
    E.pending = true;
    if (finite_check && n_fin) { // outside the timed pair: 1024 samples of every field array
      dim3 grid((unsigned)std::min(n_fin, 148 * 4), 1);
      k_finite<T><<<grid, kThreads, 0, st>>>(d_fin, n_fin, 1024, d_hnonfinite);
      launches++;
    }
    CU(cudaGetLastError());
    t += nsteps;
    steps_done += nsteps;
    halo_stale = true;
    if (profiling) { // per-kernel events are read back right away (bench.py's roofline leg)
      CU(cudaStreamSynchronize(st));
      RC(collect_timing());
      RC(prof_collect());
    }
    return MEEPCU_OK;
  }

  // fields::synchronize_magnetic_fields (src/energy_and_flux.cpp:145-166) on the device
  int collect_mag_arrays() {
    if (!mag_arrays.empty()) return MEEPCU_OK;
    std::set<const void *> seen;
    size_t off = 0;
    for (size_t ic = 0; ic < chunks.size(); ++ic) {
      const meepcu_chunk_desc &d = chunks[ic].d;
      for (int pass = 0; pass < 2; ++pass) // B components, then H components (aliases of B skipped)
        for (int c = 5 * (pass ? MEEPCU_H_STUFF : MEEPCU_B_STUFF); c < 5 * (pass ? MEEPCU_H_STUFF : MEEPCU_B_STUFF) + 5; ++c)
          for (int cmp = 0; cmp < 2; ++cmp) {
            const void *arrs[5] = {d.f[c][cmp], d.f_u[c][cmp], d.f_w[c][cmp], d.f_cond[c][cmp], d.f_bfast[c][cmp]};
            for (int k = 0; k < 5; ++k) {
              if (!arrs[k] || !seen.insert(arrs[k]).second) continue;
              std::map<const char *, ArrayRec>::iterator it = arrays.find((const char *)arrs[k]);
              if (it == arrays.end()) continue;
              MagArr m = {it->second.dev, it->second.n_dev * sizeof(T), off, k == 0, false};
              off += (m.bytes + 255) & ~(size_t)255;
              mag_arrays.push_back(m);
            }
          }
      // the fused B/H pass also advances magnetic polarisations, which the reference's
      // synchronize does not: they are put back right after the half step
      for (size_t ip = 0; ip < chunks[ic].pol[MEEPCU_H_STUFF].size(); ++ip) {
        const meepcu_susceptibility_desc &sd = chunks[ic].pol[MEEPCU_H_STUFF][ip].d;
        for (int c = 0; c < 20; ++c)
          for (int cmp = 0; cmp < 2; ++cmp) {
            const meepcu_gyrotropic_desc &gg = chunks[ic].pol[MEEPCU_H_STUFF][ip].g; // all zero unless gyrotropic
            const void *arrs[8] = {sd.P[c][cmp], sd.P_prev[c][cmp], gg.P[c][cmp][0], gg.P[c][cmp][1], gg.P[c][cmp][2],
                                   gg.P_prev[c][cmp][0], gg.P_prev[c][cmp][1], gg.P_prev[c][cmp][2]};
            for (int k = 0; k < 8; ++k) {
              if (!arrs[k] || !seen.insert(arrs[k]).second) continue;
              std::map<const char *, ArrayRec>::iterator it = arrays.find((const char *)arrs[k]);
              if (it == arrays.end()) continue;
              MagArr m = {it->second.dev, it->second.n_dev * sizeof(T), off, false, true};
              off += (m.bytes + 255) & ~(size_t)255;
              mag_arrays.push_back(m);
            }
          }
      }
    }
    if (off) CU(cudaMalloc((void **)&mag_backup, off));
    return MEEPCU_OK;
  }
  int synchronize_magnetic_fields(int nsrc, const meepcu_src_sample *smp) {
    if (!finalized) return fail(MEEPCU_ESTATE, "meepcu_finalize not called");
    if (mag_synced) return fail(MEEPCU_ESTATE, "magnetic fields are already synchronized");
    for (size_t ic = 0; ic < chunks.size(); ++ic)
      for (size_t is = 0; is < chunks[ic].src.size(); ++is)
        if (chunks[ic].src[is].tid >= nsrc || (nsrc > 0 && !smp))
          return fail(MEEPCU_EINVAL, "source refers to src_time %d but only %d samples given", chunks[ic].src[is].tid, nsrc);
    CU(cudaSetDevice(cfg.device));
    RC(collect_mag_arrays());
    char *p_smp = 0;
    {
      const size_t ns = (size_t)std::max(nsrc, 1) * 8;
      std::vector<double> tab(2 * ns, 0.0);
      if (last_row.size() == ns) memcpy(tab.data(), last_row.data(), ns * sizeof(double));
      if (nsrc > 0) memcpy(tab.data() + ns, smp, sizeof(meepcu_src_sample) * (size_t)nsrc);
      RC(ring_put(tab.data(), tab.size() * sizeof(double), &p_smp));
      cur_prev = (const double *)p_smp;
      cur_vmode = (!virt_comps.empty() && e_virtual) ? 1 : 0;
      p_smp += ns * sizeof(double);
    }
    for (size_t k = 0; k < mag_arrays.size(); ++k) // backup_component (:97-118)
      CU(cudaMemcpyAsync(mag_backup + mag_arrays[k].off, mag_arrays[k].dev, mag_arrays[k].bytes, cudaMemcpyDeviceToDevice, st));
    in_mag_sync = true;
    const int rc_half = launch_half(0, (const double *)p_smp); // step_db(B), step_source(B), step_boundaries(B), update_eh(H), step_boundaries(H)
    in_mag_sync = false;
    RC(rc_half);
    for (size_t k = 0; k < mag_arrays.size(); ++k) {
      const MagArr &m = mag_arrays[k];
      if (m.is_pol)
        CU(cudaMemcpyAsync(m.dev, mag_backup + m.off, m.bytes, cudaMemcpyDeviceToDevice, st));
      else if (m.is_f) { // average_with_backup (:135-143), every element of the array
        const long long n = (long long)(m.bytes / sizeof(T));
        k_average<T><<<(int)std::min<long long>((n + kThreads - 1) / kThreads, 148 * 16), kThreads, 0, st>>>(
            (T *)m.dev, (const T *)(mag_backup + m.off), n);
        launches++;
      }
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    mag_synced = true;
    return MEEPCU_OK;
  }
  int restore_magnetic_fields() { // src/energy_and_flux.cpp:168-177
    if (!mag_synced) return fail(MEEPCU_ESTATE, "magnetic fields are not synchronized");
    CU(cudaSetDevice(cfg.device));
    for (size_t k = 0; k < mag_arrays.size(); ++k)
      if (!mag_arrays[k].is_pol)
        CU(cudaMemcpyAsync(mag_arrays[k].dev, mag_backup + mag_arrays[k].off, mag_arrays[k].bytes, cudaMemcpyDeviceToDevice, st));
    CU(cudaStreamSynchronize(st));
    mag_synced = false;
    return MEEPCU_OK;
  }

  // every group, and NOT the MEEPCU_SYNC_LOCAL modifier: with several ranks the call is collective and refreshes the
  // halos owned by other ranks too (r02e: passing ~0u made every full sync local-only -- stale D/B halos across cuts)
  int sync_to_host() { return sync_groups(true, ~0u & ~(unsigned)MEEPCU_SYNC_LOCAL); }
  // the caller's arrays -> device: everything that changes while stepping.  Materials (chi1inv, chi2, chi3,
  // conductivity, sigma, sig/kap/siginv) were uploaded by meepcu_finalize and travel again only when asked for
  // (meepcu_sync_subset(ctx, 0, MEEPCU_SYNC_MATERIALS)): re-sending them at every sync is pure PCIe time.
  int sync_from_host() { return sync_groups(false, MEEPCU_SYNC_FIELDS | MEEPCU_SYNC_AUX | MEEPCU_SYNC_POL | MEEPCU_SYNC_DFT); }
  bool big(const ArrayRec &r) const { return r.n_host * sizeof(T) >= (size_t)(4 << 20) && !getenv("MEEPCU_NO_STAGING"); }
  // Halo copies the stepping kernels never consume (B/D/P, the normal-plane E/H points) are materialised
  // when somebody is about to read the device mirror.  `remote`: also the cross-rank ones (a collective:
  // every rank of the communicator must make the same call).  With virtual-E chunks the order matters:
  //   A  normal-plane E/H halos (for a virtual chunk they land in its D array, like every E halo)
  //   V  (make_e_real) the E arrays of virtual chunks := chi1inv * (D - dipoles), halos included
  //   B  B / D / P halos -- this overwrites the E halos parked in the D arrays of virtual chunks, so
  //   R  (after the reads) the every-step E halo copies are run again
  int halos_A(unsigned groups) {
    for (int half = 0; half < 2; ++half) {
      const int fte = half ? MEEPCU_E_STUFF : MEEPCU_H_STUFF;
      if (n_runs_def[fte] && (groups & MEEPCU_SYNC_FIELDS)) {
        k_conn_runs<T><<<(n_runs_def[fte] + 7) / 8, kThreads, 0, st>>>(d_runs_def[fte], n_runs_def[fte]);
        launches++;
      }
    }
    return MEEPCU_OK;
  }
  int halos_B(unsigned groups, bool remote) {
    for (int half = 0; half < 2; ++half) {
      const int ftd = half ? MEEPCU_D_STUFF : MEEPCU_B_STUFF, ftp = half ? MEEPCU_PE_STUFF : MEEPCU_PH_STUFF;
      if (!eager_db[half] && (groups & MEEPCU_SYNC_FIELDS)) RC(remote ? launch_conn(ftd) : launch_conn_local(ftd));
      if (!eager_p[half] && (groups & MEEPCU_SYNC_POL)) RC(remote ? launch_conn(ftp) : launch_conn_local(ftp));
    }
    return MEEPCU_OK;
  }
  bool virt_now(unsigned groups) const { return virt_any && e_virtual && (groups & MEEPCU_SYNC_FIELDS); }
  int halos_R(unsigned groups, bool remote) {
    if (!virt_now(groups)) return MEEPCU_OK;
    return remote ? launch_conn(MEEPCU_E_STUFF) : launch_conn_local(MEEPCU_E_STUFF);
  }
  // before a gather of values that may be lazily copied not-owned points, read through the virtual-E aliases:
  // B and R first, A last (B would overwrite the normal-plane E halos A parks in the D arrays of virtual chunks)
  int refresh_for_read() {
    const unsigned groups = MEEPCU_SYNC_FIELDS | MEEPCU_SYNC_POL;
    RC(halos_B(groups, false));
    RC(halos_R(groups, false));
    RC(halos_A(groups));
    if (n_ranks == 1) halo_stale = false;
    return MEEPCU_OK;
  }
  int sync_groups(bool to_host, unsigned groups) {
    if (to_host && !finalized) return fail(MEEPCU_ESTATE, "meepcu_finalize not called");
    CU(cudaSetDevice(cfg.device));
    const bool local_only = (groups & MEEPCU_SYNC_LOCAL) != 0;
    groups &= ~(unsigned)MEEPCU_SYNC_LOCAL;
    const bool remote = n_ranks > 1 && !local_only;
    if (to_host) {
      // so that the not-owned points of the host mirror hold what step_boundaries would have left there
      RC(halos_A(groups));
      if (virt_now(groups)) RC(virt_materialise());
      RC(halos_B(groups, remote));
      if (n_ranks == 1 || !local_only) halo_stale = false;
    }
    std::vector<const ArrayRec *> small;
    for (std::map<const char *, ArrayRec>::iterator it = arrays.begin(); it != arrays.end(); ++it) {
      const ArrayRec &r = it->second;
      if (!(r.group & groups) || (to_host && !r.mut)) continue;
      if (batchable(r)) { small.push_back(&r); continue; }
      RC(big(r) && !r.pinned ? staged_copy(r, to_host) : copy_array(r, to_host));
    }
    RC(batch_copy(small, to_host));
    if (to_host) {
      RC(halos_R(groups, remote));
      if (virt_now(groups) && n_runs_def[MEEPCU_E_STUFF]) halo_stale = true; // (the normal-plane E halos parked in D arrays: next A)
    }
    else {
      if (groups & MEEPCU_SYNC_FIELDS) e_virtual = false; // every E array on the device is the host's again
      if (groups & MEEPCU_SYNC_MATERIALS) { RC(virt_patch_u()); RC(build_eh_masks()); }
    }
    CU(cudaStreamSynchronize(st));
    RC(collect_timing());
    return MEEPCU_OK;
  }
  // How one requested element is read from the device.  0: as it is.  1: it is a not-owned point whose halo copy is
  // lazy (B / D / P): refresh_for_read first.  2: it needs the real E array of a virtual chunk (a dipole point) or
  // a not-owned D point of such a chunk (parked E halo): refresh_for_read(make_e_real).  Virtual E elements are
  // otherwise read as D (x chi1inv) -- *src / *mul.
  int classify_read(const void *p, const T **src, const T **mul) const {
    *src = *mul = 0;
    size_t e = 0;
    const ArrayRec *r = find(p, sizeof(T), &e);
    if (!r) return -1;
    if (r->kind != A_GRID) { *src = (const T *)r->dev + e; return 0; }
    const Geom &g = chunks[r->chunk].g;
    const long long di = g.dev_index((long long)e);
    *src = (const T *)r->dev + di;
    bool owned = true;
    if (r->comp >= 0) {
      long long rem = (long long)e;
      for (int d = 0; d < 3; ++d)
        if (g.has[d]) {
          const long long q = rem / g.hs[d];
          rem -= q * g.hs[d];
          if (q < 1 - g.yee_shift(r->comp, d) || q > g.num[d] - g.yee_shift(r->comp, d)) owned = false;
        }
    }
    const bool virt_on = !virt_comps.empty() && e_virtual;
    if (virt_on && r->comp >= 0 && virt_chunks.count(r->chunk)) {
      if (is_electric(r->comp)) {
        for (size_t k = 0; k < virt_comps.size(); ++k)
          if ((const char *)virt_comps[k].e == r->dev)
            for (int j = 0; j < virt_comps[k].nsrc; ++j)
              if (h_src_idx[(virt_comps[k].src - d_src) + j] == di) return 2; // E = chi1inv * (D - dipole): the real array
        const T *m = 0;
        *src = dev_elem_v(p, MEEPCU_E_STUFF, &m);
        *mul = m;
        // (halos of E are parked in D and chi1inv is 1 there: the product is right for them too; the normal-plane
        //  ones are copied lazily)
        return (!owned && n_runs_def[MEEPCU_E_STUFF]) ? 1 : 0;
      }
      if (is_D(r->comp) && !owned) return 2;
    }
    if (!owned && (is_D(r->comp) || is_B(r->comp)) ) return 1;
    if (r->group == MEEPCU_SYNC_POL) return 1; // polarisation halos are lazy too (which points are owned is not recorded)
    if (!owned && r->comp >= 0 && n_runs_def[is_electric(r->comp) ? MEEPCU_E_STUFF : MEEPCU_H_STUFF]) return 1; // normal-plane E/H halos
    return 0;
  }
  // n elements gathered on the device and returned as doubles in one transfer
  int read_values(size_t n, const void *const *ptrs, double *out) {
    if (!finalized) return fail(MEEPCU_ESTATE, "meepcu_finalize not called");
    if (!n) return MEEPCU_OK;
    if (!ptrs || !out) return fail(MEEPCU_EINVAL, "null argument");
    CU(cudaSetDevice(cfg.device));
    std::vector<const T *> h(2 * n);
    int worst = 0;
    for (size_t i = 0; i < n; ++i) {
      const int c = classify_read(ptrs[i], &h[i], &h[n + i]);
      if (c < 0) return fail(MEEPCU_EINVAL, "pointer %zu is not inside a registered array", i);
      worst = std::max(worst, c);
    }
    const unsigned fp = MEEPCU_SYNC_FIELDS | MEEPCU_SYNC_POL;
    if (worst == 2) { // rare: every element from the REAL arrays, between B and R (the order of meepcu_sync_to_host)
      RC(halos_A(fp));
      RC(virt_materialise());
      RC(halos_B(fp, false));
      for (size_t i = 0; i < n; ++i) { h[i] = dev_elem(ptrs[i]); h[n + i] = 0; }
    }
    else if (worst == 1 && halo_stale)
      RC(refresh_for_read());
    const T **d_ptr = 0;
    T *d_val = 0;
    CU(cudaMalloc((void **)&d_ptr, 2 * n * sizeof(T *)));
    if (cudaMalloc((void **)&d_val, n * sizeof(T)) != cudaSuccess) { cudaFree(d_ptr); return fail(MEEPCU_ENOMEM, "cudaMalloc"); }
    std::vector<T> v(n);
    cudaError_t e = cudaMemcpyAsync(d_ptr, h.data(), 2 * n * sizeof(T *), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
      k_gather<T><<<(int)std::min<size_t>((n + kThreads - 1) / kThreads, 148 * 8), kThreads, 0, st>>>(d_ptr, d_ptr + n, d_val, (long long)n);
      launches++;
      e = cudaMemcpyAsync(v.data(), d_val, n * sizeof(T), cudaMemcpyDeviceToHost, st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_ptr);
    cudaFree(d_val);
    if (e != cudaSuccess) return fail(MEEPCU_ECUDA, "%s", cudaGetErrorString(e));
    if (worst == 2) {
      RC(halos_R(fp, false));
      halo_stale = true; // the normal-plane E halos parked in D arrays are gone until the next A
    }
    for (size_t i = 0; i < n; ++i) out[i] = (double)v[i];
    return MEEPCU_OK;
  }

  // ---- reductions for the host-side readers: dft_flux::flux / dft_energy (dft.cpp:533-548,657-687) ----
  int get_compulsory_bytes(int64_t *per_step, int64_t *per_kernel, int cap, int *n) {
    if (per_step) *per_step = alg_bytes_sv;
    int m = 0;
    for (int k = 0; k < KC_COUNT && m < cap; ++k) per_kernel[m++] = kc_bytes_sv[k];
    if (n) *n = m;
    return MEEPCU_OK;
  }
  int dft_dot(size_t npairs, const void *const *a, const void *const *b, int Nomega, double *out) {
    if (!finalized) return fail(MEEPCU_ESTATE, "meepcu_finalize not called");
    if (Nomega <= 0 || !out || (npairs && (!a || !b))) return fail(MEEPCU_EINVAL, "bad argument");
    for (int i = 0; i < Nomega; ++i) out[i] = 0;
    if (!npairs) return MEEPCU_OK;
    CU(cudaSetDevice(cfg.device));
    int W = 1;
    while (W < Nomega && W < kThreads) W <<= 1;
    const int lanes = kThreads / W;
    const long long per = std::max<long long>(lanes, ((long long)1 << 15) / W); // points per work item
    std::vector<DftDotWork<T> > work;
    for (size_t p = 0; p < npairs; ++p) {
      size_t ea = 0, eb = 0;
      const ArrayRec *ra = find(a[p], sizeof(T), &ea), *rb = find(b[p], sizeof(T), &eb);
      if (!ra || !rb || ra->kind != A_DFT || rb->kind != A_DFT || ea || eb)
        return fail(MEEPCU_EINVAL, "pair %zu: not dft_chunk::dft arrays registered by meepcu_add_dft", p);
      if (ra->n_host != rb->n_host || ra->n_host % (size_t)Nomega)
        return fail(MEEPCU_EINVAL, "pair %zu: the two dft chunks differ in size, or Nomega does not divide it", p);
      const long long N = (long long)(ra->n_host / (size_t)Nomega);
      for (long long k0 = 0; k0 < N; k0 += per) {
        DftDotWork<T> w = {(const T *)ra->dev, (const T *)rb->dev, k0, std::min(N, k0 + per)};
        work.push_back(w);
      }
    }
    if (work.empty()) return MEEPCU_OK;
    DftDotWork<T> *d_work = 0;
    double *d_part = 0;
    const size_t np = work.size() * (size_t)Nomega;
    CU(cudaMalloc((void **)&d_work, work.size() * sizeof(DftDotWork<T>)));
    if (cudaMalloc((void **)&d_part, np * sizeof(double)) != cudaSuccess) { cudaFree(d_work); return fail(MEEPCU_ENOMEM, "cudaMalloc"); }
    std::vector<double> part(np);
    cudaError_t e = cudaMemcpyAsync(d_work, work.data(), work.size() * sizeof(DftDotWork<T>), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
      dim3 grid((unsigned)work.size(), (unsigned)((Nomega + W - 1) / W));
      k_dft_dot<T><<<grid, kThreads, 0, st>>>(d_work, Nomega, W, d_part);
      launches++;
      e = cudaMemcpyAsync(part.data(), d_part, np * sizeof(double), cudaMemcpyDeviceToHost, st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_work);
    cudaFree(d_part);
    if (e != cudaSuccess) return fail(MEEPCU_ECUDA, "%s", cudaGetErrorString(e));
    RC(collect_timing());
    for (size_t w = 0; w < work.size(); ++w)
      for (int i = 0; i < Nomega; ++i) out[i] += part[w * (size_t)Nomega + i];
    return MEEPCU_OK;
  }

  // ---- fields::integrate(2, cs, dot_integrand, ...) (integrate.cpp:47-131, energy_and_flux.cpp:61-85,189-215) ----
  int integrate_dot(size_t n, const meepcu_integrate_desc *pc, double *sum, double *maxabs) {
    if (!finalized) return fail(MEEPCU_ESTATE, "meepcu_finalize not called");
    if (!sum || (n && !pc)) return fail(MEEPCU_EINVAL, "null argument");
    *sum = 0;
    if (maxabs) *maxabs = 0;
    if (!n) return MEEPCU_OK;
    CU(cudaSetDevice(cfg.device));
    std::vector<IntgParams<T> > par;
    std::vector<IntgWork> work;
    for (size_t ip = 0; ip < n; ++ip) {
      const meepcu_integrate_desc &d = pc[ip];
      if (!d.f[0][0] || !d.f[1][0]) continue; // a component this chunk does not have: the integrand is 0 (integrate.cpp:110-116)
      size_t e0 = 0;
      const ArrayRec *r0 = find(d.f[0][0], sizeof(T), &e0);
      if (!r0 || r0->kind != A_GRID || e0) return fail(MEEPCU_EINVAL, "piece %zu: f[0][0] is not a registered field array", ip);
      const Geom &g = chunks[r0->chunk].g;
      IntgParams<T> q;
      memset(&q, 0, sizeof q);
      for (int c = 0; c < 2; ++c)
        for (int k = 0; k < 2; ++k) {
          if (!d.f[c][k]) continue;
          size_t e = 0;
          const ArrayRec *r = find(d.f[c][k], sizeof(T), &e);
          if (!r || r->kind != A_GRID || e || r->chunk != r0->chunk)
            return fail(MEEPCU_EINVAL, "piece %zu: f[%d][%d] is not a field array of the same chunk", ip, c, k);
          q.f[c][k] = (const T *)r->dev;
        }
      long long idx0 = 0, N = 1;
      for (int k = 0; k < 3; ++k) { // LOOP_OVER_IVECS in the reference's loop order (vec.hpp:151-168)
        const int dir = loop_slot(k);
        q.dir[k] = dir;
        if (g.has[dir]) {
          if (d.ie[dir] < d.is[dir]) { N = 0; break; }
          q.n[k] = (d.ie[dir] - d.is[dir]) / 2 + 1;
          q.st[k] = g.ds[dir];
          idx0 += (long long)((d.is[dir] - g.io[dir]) / 2) * g.ds[dir];
          N *= q.n[k];
        }
        else {
          q.n[k] = 1;
          q.st[k] = 0;
        }
      }
      if (N <= 0) continue;
      q.idx0 = idx0;
      q.N = N;
      for (int c = 0; c < 2; ++c)
        for (int k = 0; k < 2; ++k) {
          q.off[c][k] = g.dev_offset((long long)d.off[c][k]);
          if (q.off[c][k] == ((long long)1 << 62)) return fail(MEEPCU_EINVAL, "piece %zu: yee2cent offsets are not sums of +-strides", ip);
          q.ph[c][k] = d.ph[c][k];
        }
      for (int k = 0; k < 3; ++k) { q.s0[k] = d.s0[k]; q.s1[k] = d.s1[k]; q.e0[k] = d.e0[k]; q.e1[k] = d.e1[k]; }
      q.dV0 = d.dV0;
      q.dV1 = d.dV1;
      const int pidx = (int)par.size();
      par.push_back(q);
      for (long long p0 = 0; p0 < N; p0 += (long long)kThreads * kIntgPerThread) {
        IntgWork w = {pidx, p0};
        work.push_back(w);
      }
    }
    if (work.empty()) return MEEPCU_OK;
    // the device mirror as a host reader would see it: lazy halos materialised, virtual E arrays made real
    const unsigned fp = MEEPCU_SYNC_FIELDS | MEEPCU_SYNC_POL;
    const bool virt = virt_now(fp);
    if (virt) {
      RC(halos_A(fp));
      RC(virt_materialise());
      RC(halos_B(fp, false));
    }
    else if (halo_stale)
      RC(refresh_for_read());
    IntgParams<T> *d_par = 0;
    IntgWork *d_work = 0;
    double *d_part = 0;
    CU(cudaMalloc((void **)&d_par, par.size() * sizeof(IntgParams<T>)));
    if (cudaMalloc((void **)&d_work, work.size() * sizeof(IntgWork)) != cudaSuccess) { cudaFree(d_par); return fail(MEEPCU_ENOMEM, "cudaMalloc"); }
    if (cudaMalloc((void **)&d_part, 2 * work.size() * sizeof(double)) != cudaSuccess) { cudaFree(d_par); cudaFree(d_work); return fail(MEEPCU_ENOMEM, "cudaMalloc"); }
    std::vector<double> part(2 * work.size());
    cudaError_t e = cudaMemcpyAsync(d_par, par.data(), par.size() * sizeof(IntgParams<T>), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_work, work.data(), work.size() * sizeof(IntgWork), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
      k_integrate_dot<T><<<(unsigned)work.size(), kThreads, 0, st>>>(d_par, d_work, d_part);
      launches++;
      e = cudaMemcpyAsync(part.data(), d_part, part.size() * sizeof(double), cudaMemcpyDeviceToHost, st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_par);
    cudaFree(d_work);
    cudaFree(d_part);
    if (e != cudaSuccess) return fail(MEEPCU_ECUDA, "%s", cudaGetErrorString(e));
    if (virt) {
      RC(halos_R(fp, false));
      halo_stale = true; // the normal-plane E halos parked in D arrays are gone until the next A
    }
    RC(collect_timing());
    double s = 0, m = 0;
    for (size_t w = 0; w < work.size(); ++w) {
      s += part[2 * w];
      m = std::max(m, part[2 * w + 1]);
    }
    *sum = s;
    if (maxabs) *maxabs = m;
    return MEEPCU_OK;
  }
  int check_finite(int *ok) { // every element of every field array, one launch, blocking
    if (!finalized) return fail(MEEPCU_ESTATE, "meepcu_finalize not called");
    CU(cudaSetDevice(cfg.device));
    CU(cudaMemsetAsync(d_flags + 1, 0, sizeof(int), st));
    if (n_fin) {
      dim3 grid((unsigned)std::min(n_fin, 64), 148 * 4 / std::min(n_fin, 64) + 1);
      k_finite<T><<<grid, kThreads, 0, st>>>(d_fin, n_fin, fin_max_n, d_flags + 1);
      launches++;
    }
    int flags[2] = {0, 0};
    CU(cudaMemcpyAsync(flags, d_flags, sizeof flags, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    RC(collect_timing());
    *ok = flags[1] == 0 && !nonfinite_seen();
    nr_unconverged = flags[0];
    return MEEPCU_OK;
  }
  int read_value(const void *p, double *out) { return read_values(1, &p, out); }
  int get_stats(meepcu_stats *out) {
    memset(out, 0, sizeof *out);
    if (finalized) {
      CU(cudaSetDevice(cfg.device));
      RC(collect_timing()); // waits for the queued steps: last_step_ms is the device time of the last call
      int fl[8];
      CU(cudaMemcpy(fl, d_flags, sizeof fl, cudaMemcpyDeviceToHost));
      nr_unconverged = fl[0];
      unsigned long long cnt[2];
      memcpy(cnt, fl + 4, sizeof cnt);
      out->reserved[1] = (int64_t)cnt[0]; // Newton-Raphson solves counted while profiling was on
      out->reserved[2] = (int64_t)cnt[1]; // ... their iterations
      out->reserved[6] = fl[3];           // ... the longest of them
    }
    out->reserved[5] = (int64_t)virt_comps.size(); // E arrays of this rank that are not stored while stepping (virtual E)
    out->reserved[3] = nonfinite_seen();
    out->reserved[4] = (int64_t)(total_step_ms * 1e3); // device microseconds of all meepcu_step calls so far
    out->kernel_launches = launches;
    out->steps = steps_done;
    out->device_bytes = (int64_t)arena_bytes;
    out->cells = cells;
    out->algorithmic_bytes_per_step = alg_bytes;
    out->last_step_ms = last_ms;
    out->reserved[0] = nr_unconverged;
    for (std::map<const char *, ArrayRec>::const_iterator it = arrays.begin(); it != arrays.end(); ++it)
      if (it->second.pinned) // bytes of host arrays page-locked in place (meepcu_config::pin_host)
        out->reserved[7] += (int64_t)(it->second.n_host * sizeof(T) * (it->second.kind == A_DFT ? 2 : 1));
    return MEEPCU_OK;
  }
  int get_timers(double out[22]) {
    // reference time_sink numbering (src/meep.hpp:1610-1633); seconds of device time
    for (int i = 0; i < 22; ++i) out[i] = 0;
    if (finalized) RC(collect_timing());
    out[1] = total_step_ms * 1e-3; // Stepping: device time of every meepcu_step call (always on); the per-phase
                                   // entries below are filled while meepcu_set_profiling(1) serialises the kernels
    out[10] = (kc_ms[KC_HALF_B_SIMPLE] + kc_ms[KC_HALF_B_GENERAL] + kc_ms[KC_CURL3D_B] + kc_ms[KC_MIX_B]) * 1e-3; // FieldUpdateB (+H, fused)
    out[12] = (kc_ms[KC_HALF_D_SIMPLE] + kc_ms[KC_HALF_D_GENERAL] + kc_ms[KC_CURL3D_D] + kc_ms[KC_MIX_D]) * 1e-3; // FieldUpdateD (+E, fused)
    out[13] = (kc_ms[KC_FMP] + kc_ms[KC_EH_GENERAL]) * 1e-3;               // FieldUpdateE (unfused part)
    out[2] = kc_ms[KC_CONN] * 1e-3;                                        // Boundaries
    out[6] = kc_ms[KC_DFT] * 1e-3;                                         // FourierTransforming
    return MEEPCU_OK;
  }
};

} // namespace

struct meepcu_ctx {
  CtxBase *impl;
};

extern "C" {

const char *meepcu_last_error(void) { return g_err.c_str(); }
const char *meepcu_backend(void) { return "cuda-sm_100a"; }
int meepcu_abi_version(void) { return MEEPCU_ABI_VERSION; }

int meepcu_create(const meepcu_config *cfg, meepcu_ctx **out) {
  if (!cfg || !out) return fail(MEEPCU_EINVAL, "null argument");
  if (cfg->abi_version != MEEPCU_ABI_VERSION) return fail(MEEPCU_EINVAL, "ABI version mismatch");
  if (cfg->dim < 0 || cfg->dim > 3) return fail(MEEPCU_EINVAL, "bad dim");
  if (cfg->nl_mode != MEEPCU_NL_FORK_LITERAL && cfg->nl_mode != MEEPCU_NL_UPSTREAM_PADE)
    return fail(MEEPCU_EINVAL, "bad nl_mode");
  CtxBase *impl;
  if (cfg->realnum_bytes == 8)
    impl = new Ctx<double>();
  else if (cfg->realnum_bytes == 4)
    impl = new Ctx<float>();
  else
    return fail(MEEPCU_EINVAL, "realnum_bytes must be 4 or 8");
  impl->cfg = *cfg;
  impl->t = cfg->t;
  impl->finalized = false;
  *out = new meepcu_ctx;
  (*out)->impl = impl;
  return MEEPCU_OK;
}
int meepcu_destroy(meepcu_ctx *ctx) {
  if (!ctx) return MEEPCU_OK;
  delete ctx->impl;
  delete ctx;
  return MEEPCU_OK;
}
#define CHECK_CTX                                                                                  \
  if (!ctx || !ctx->impl) return fail(MEEPCU_EINVAL, "null context")
int meepcu_add_chunk(meepcu_ctx *ctx, const meepcu_chunk_desc *d, int *id) {
  CHECK_CTX;
  if (!d) return fail(MEEPCU_EINVAL, "null chunk descriptor");
  return ctx->impl->add_chunk(d, id);
}
int meepcu_add_susceptibility(meepcu_ctx *ctx, int chunk, const meepcu_susceptibility_desc *d) {
  CHECK_CTX;
  if (!d) return fail(MEEPCU_EINVAL, "null descriptor");
  return ctx->impl->add_susc(chunk, d);
}
int meepcu_add_gyrotropic(meepcu_ctx *ctx, int chunk, const meepcu_gyrotropic_desc *d) {
  CHECK_CTX;
  if (!d) return fail(MEEPCU_EINVAL, "null descriptor");
  return ctx->impl->add_gyro(chunk, d);
}
int meepcu_add_multilevel(meepcu_ctx *ctx, int chunk, const meepcu_multilevel_desc *d) {
  CHECK_CTX;
  if (!d) return fail(MEEPCU_EINVAL, "null descriptor");
  return ctx->impl->add_multilevel(chunk, d);
}
int meepcu_get_noise_count(meepcu_ctx *ctx, size_t *n) {
  CHECK_CTX;
  if (!n) return fail(MEEPCU_EINVAL, "null argument");
  *n = ctx->impl->noise_count();
  return MEEPCU_OK;
}
int meepcu_set_noise(meepcu_ctx *ctx, int nsteps, size_t n_per_step, const double *g) {
  CHECK_CTX;
  return ctx->impl->set_noise(nsteps, n_per_step, g);
}
int meepcu_add_source(meepcu_ctx *ctx, int chunk, const meepcu_source_desc *d) {
  CHECK_CTX;
  if (!d) return fail(MEEPCU_EINVAL, "null descriptor");
  return ctx->impl->add_source(chunk, d);
}
int meepcu_add_dft(meepcu_ctx *ctx, int chunk, const meepcu_dft_desc *d, int *id) {
  CHECK_CTX;
  if (!d) return fail(MEEPCU_EINVAL, "null descriptor");
  return ctx->impl->add_dft(chunk, d, id);
}
int meepcu_set_metal(meepcu_ctx *ctx, int ft, size_t n, void *const *ptrs) {
  CHECK_CTX;
  return ctx->impl->set_metal(ft, n, ptrs);
}
int meepcu_add_connections(meepcu_ctx *ctx, int ft, int kind, size_t n, void *const *dst,
                           const void *const *src, const void *phases) {
  CHECK_CTX;
  return ctx->impl->add_conn(ft, kind, n, dst, src, phases);
}
int meepcu_add_remote_send(meepcu_ctx *ctx, int ft, int kind, int peer, size_t n, const void *const *src) {
  CHECK_CTX;
  return ctx->impl->add_remote(ft, kind, peer, n, src, 0, false);
}
int meepcu_add_remote_recv(meepcu_ctx *ctx, int ft, int kind, int peer, size_t n, void *const *dst,
                           const void *phases) {
  CHECK_CTX;
  return ctx->impl->add_remote(ft, kind, peer, n, (const void *const *)dst, phases, true);
}
int meepcu_comm_unique_id(void *id128) {
  NcclApi *n = nccl_api();
  if (!n) return fail(MEEPCU_ENCCL, "libnccl.so.2 not found");
  nccl_uid uid;
  NC(n->GetUniqueId(&uid));
  memcpy(id128, &uid, sizeof uid);
  return MEEPCU_OK;
}
int meepcu_comm_init(meepcu_ctx *ctx, int rank, int nranks, const void *id128) {
  CHECK_CTX;
  if (!id128) return fail(MEEPCU_EINVAL, "null unique id");
  return ctx->impl->comm_init(rank, nranks, id128);
}
int meepcu_finalize(meepcu_ctx *ctx) {
  CHECK_CTX;
  return ctx->impl->finalize();
}
int meepcu_step(meepcu_ctx *ctx, int nsteps, int nsrc, const meepcu_src_sample *smp) {
  CHECK_CTX;
  return ctx->impl->step(nsteps, nsrc, smp);
}
int meepcu_synchronize_magnetic_fields(meepcu_ctx *ctx, int nsrc, const meepcu_src_sample *smp) {
  CHECK_CTX;
  return ctx->impl->synchronize_magnetic_fields(nsrc, smp);
}
int meepcu_restore_magnetic_fields(meepcu_ctx *ctx) {
  CHECK_CTX;
  return ctx->impl->restore_magnetic_fields();
}
int meepcu_sync_to_host(meepcu_ctx *ctx) {
  CHECK_CTX;
  return ctx->impl->sync_to_host();
}
int meepcu_sync_from_host(meepcu_ctx *ctx) {
  CHECK_CTX;
  if (!ctx->impl->finalized) return fail(MEEPCU_ESTATE, "meepcu_finalize not called");
  return ctx->impl->sync_from_host();
}
int meepcu_sync_subset(meepcu_ctx *ctx, int to_host, unsigned groups) {
  CHECK_CTX;
  if (groups & ~63u) return fail(MEEPCU_EINVAL, "unknown array group");
  return ctx->impl->sync_groups(to_host != 0, groups);
}
int meepcu_read_values(meepcu_ctx *ctx, size_t n, const void *const *host_ptrs, double *out) {
  CHECK_CTX;
  return ctx->impl->read_values(n, host_ptrs, out);
}
int meepcu_get_time(meepcu_ctx *ctx, int *t) {
  CHECK_CTX;
  *t = ctx->impl->t;
  return MEEPCU_OK;
}
int meepcu_set_time(meepcu_ctx *ctx, int t) {
  CHECK_CTX;
  ctx->impl->t = t;
  return MEEPCU_OK;
}
int meepcu_check_finite(meepcu_ctx *ctx, int *ok) {
  CHECK_CTX;
  return ctx->impl->check_finite(ok);
}
int meepcu_read_value(meepcu_ctx *ctx, const void *p, double *out) {
  CHECK_CTX;
  return ctx->impl->read_value(p, out);
}
int meepcu_get_compulsory_bytes(meepcu_ctx *ctx, int64_t *per_step, int64_t *per_kernel, int capacity, int *n) {
  if (!ctx || !ctx->impl) return fail(MEEPCU_EINVAL, "null context");
  if (capacity > 0 && !per_kernel) return fail(MEEPCU_EINVAL, "null argument");
  return ctx->impl->get_compulsory_bytes(per_step, per_kernel, capacity, n);
}
int meepcu_dft_dot(meepcu_ctx *ctx, size_t npairs, const void *const *a, const void *const *b, int Nomega, double *out) {
  if (!ctx || !ctx->impl) return fail(MEEPCU_EINVAL, "null context");
  return ctx->impl->dft_dot(npairs, a, b, Nomega, out);
}
int meepcu_integrate_dot(meepcu_ctx *ctx, size_t npieces, const meepcu_integrate_desc *pieces, double *sum, double *maxabs) {
  if (!ctx || !ctx->impl) return fail(MEEPCU_EINVAL, "null context");
  return ctx->impl->integrate_dot(npieces, pieces, sum, maxabs);
}
int meepcu_get_timers(meepcu_ctx *ctx, double out[22]) {
  CHECK_CTX;
  return ctx->impl->get_timers(out);
}
int meepcu_get_stats(meepcu_ctx *ctx, meepcu_stats *out) {
  CHECK_CTX;
  return ctx->impl->get_stats(out);
}
int meepcu_set_profiling(meepcu_ctx *ctx, int on) {
  CHECK_CTX;
  return ctx->impl->set_profiling(on);
}
int meepcu_get_kernel_times(meepcu_ctx *ctx, meepcu_kernel_time *out, int cap, int *n) {
  CHECK_CTX;
  if (!out || !n) return fail(MEEPCU_EINVAL, "null argument");
  return ctx->impl->get_kernel_times(out, cap, n);
}
}
