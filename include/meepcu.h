/* meepcu.h -- C ABI of libmeepcuda.so: the B200 (sm_100a) implementation of Meep's
 * per-timestep field update (the PMack10/meep_nl fork), behind the seam
 * `fields::step()` (reference src/step.cpp:35-140).
 *
 * Everything that crosses this boundary is a plain pointer, size or POD struct:
 * no C++ types, no torch types, no exceptions.  Every function returns
 * MEEPCU_OK (0) or a negative MEEPCU_E* code; the message is available from
 * meepcu_last_error() (the reference-side shim turns that into
 * meep::abort("%s", ...), reference src/mympi.cpp:244-263).
 *
 * Ownership: all host arrays handed in stay owned by the caller (the
 * reference's fields_chunk / structure_chunk, new[]/delete[],
 * src/fields.cpp:164-209).  The context owns device mirrors and remembers the
 * host addresses so that meepcu_sync_to_host()/meepcu_sync_from_host() can move
 * data back and forth and so that connection / metal / polarisation pointers can
 * be given exactly as the reference stores them (as realnum* into those arrays).
 *
 * Threading: one caller thread per context (the reference is single threaded
 * per process around fields::step, src/meep.hpp:1637).
 *
 * The same ABI is implemented, on the CPU and in place on the caller's arrays,
 * by oracle/liboracle_meep.so (test infrastructure, never loaded by the product).
 */
#ifndef MEEPCU_H
#define MEEPCU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MEEPCU_ABI_VERSION 2

/* ---- enums: numeric values are the reference's (src/meep/vec.hpp:27-79) ---- */
#define MEEPCU_NUM_COMPONENTS 20 /* Ex,Ey,Er,Ep,Ez,Hx..Hz,Dx..Dz,Bx..Bz           vec.hpp:31-52 */
#define MEEPCU_NUM_FIELD_TYPES 8 /* E,H,D,B,PE,PH,WE,WH                           vec.hpp:66-75 */
enum { MEEPCU_Ex = 0, MEEPCU_Ey, MEEPCU_Er, MEEPCU_Ep, MEEPCU_Ez,
       MEEPCU_Hx, MEEPCU_Hy, MEEPCU_Hr, MEEPCU_Hp, MEEPCU_Hz,
       MEEPCU_Dx, MEEPCU_Dy, MEEPCU_Dr, MEEPCU_Dp, MEEPCU_Dz,
       MEEPCU_Bx, MEEPCU_By, MEEPCU_Br, MEEPCU_Bp, MEEPCU_Bz };
enum { MEEPCU_D1 = 0, MEEPCU_D2 = 1, MEEPCU_D3 = 2, MEEPCU_Dcyl = 3 };          /* vec.hpp:65 */
enum { MEEPCU_X = 0, MEEPCU_Y, MEEPCU_Z, MEEPCU_R, MEEPCU_P, MEEPCU_NO_DIRECTION }; /* vec.hpp:77 */
enum { MEEPCU_E_STUFF = 0, MEEPCU_H_STUFF, MEEPCU_D_STUFF, MEEPCU_B_STUFF,
       MEEPCU_PE_STUFF, MEEPCU_PH_STUFF, MEEPCU_WE_STUFF, MEEPCU_WH_STUFF };
/* connect_phase, src/meep.hpp:741 */
enum { MEEPCU_CONNECT_PHASE = 0, MEEPCU_CONNECT_NEGATE = 1, MEEPCU_CONNECT_COPY = 2 };

/* how update_eh treats chi2/chi3:
 *   FORK_LITERAL  : what the fork's CPU code computes as compiled -- diagonal
 *                   E = chi1inv*D, plus the Newton-Raphson chi2 solve in non-PML
 *                   chunks whose off-diagonal chi1inv rows and chi3 are allocated
 *                   (step_generic.cpp:730-816, newton_raphson.cpp:93-359).
 *   UPSTREAM_PADE : upstream Meep's Kerr/Pockels Pade update that the fork
 *                   commented out (step_generic.cpp:546-553,617-619,772-773,...). */
enum { MEEPCU_NL_FORK_LITERAL = 0, MEEPCU_NL_UPSTREAM_PADE = 1 };

/* error codes */
enum { MEEPCU_OK = 0, MEEPCU_EINVAL = -1, MEEPCU_EUNSUPPORTED = -2, MEEPCU_ECUDA = -3,
       MEEPCU_ENOMEM = -4, MEEPCU_ESTATE = -5, MEEPCU_ENCCL = -6, MEEPCU_ENONFINITE = -7 };

typedef struct meepcu_ctx meepcu_ctx;

/* Replaces the scalar state of `fields` / `fields_chunk` (src/meep.hpp:1733-1750). */
typedef struct {
  int abi_version;   /* MEEPCU_ABI_VERSION */
  int realnum_bytes; /* sizeof(realnum): 8, or 4 for the --enable-single build (meep.hpp:42-46) */
  int is_real;       /* fields::is_real: 1 = only cmp 0 exists */
  int dim;           /* grid_volume::dim (D1/D2/D3/Dcyl) */
  double a;          /* resolution */
  double dt;         /* fields::dt = Courant/a */
  double courant;    /* fields_chunk::Courant: the dt/dx factor passed to step_curl (step_db.cpp:125) */
  double beta;       /* fields::beta, 2D exp(i beta z) term (step_db.cpp:148-175) */
  int nl_mode;       /* MEEPCU_NL_* */
  int device;        /* CUDA device ordinal for this context (ignored by the oracle) */
  int t;             /* fields::t at attach time */
  int pin_host;      /* 1: page-lock the caller's big arrays in place (cudaHostRegister) for the life of the
                        context, so that meepcu_sync_* are plain DMA at PCIe speed; the arrays must then
                        outlive the context.  0: copy through library-owned pinned staging buffers. */
  double m;          /* fields::m, the exp(i m phi) order of a cylindrical (Dcyl) simulation (step_db.cpp:178-290) */
  int zero_fields_near_cylorigin; /* fields::zero_fields_near_cylorigin (step_db.cpp:404-440) */
  int reserved[3];
  double bfast_scaled_k[3]; /* fields::bfast_scaled_k (X,Y,Z), all zero = no BFAST terms (step_db.cpp:65,129-143) */
} meepcu_config;

/* One fields_chunk + its structure_chunk (src/meep.hpp:590-598,1461-1485).  Array length is
 * ntot = prod_d (num[d]+1) over the directions the dimensionality has, z fastest, then y, then x
 * (src/vec.cpp:293-296,482-494).  A NULL pointer means "not allocated" exactly as in the
 * reference; f[Hc]==f[Bc] means H aliases B (src/fields.cpp:493-517). */
typedef struct {
  int num[3];   /* grid_volume::num_direction(X,Y,Z); 0 for absent directions.  Dcyl: slot 0 = R, slot 2 = Z (the
                   reference's own num[d % 3] layout, vec.hpp:1014-1020), here and in every is/ie/s0.. triple below */
  int io[3];    /* grid_volume::little_corner(), half-pixel units, X,Y,Z (Dcyl: R,-,Z) */
  void *f[MEEPCU_NUM_COMPONENTS][2];      /* fields_chunk::f        */
  void *f_u[MEEPCU_NUM_COMPONENTS][2];    /* fields_chunk::f_u      (PML, integrated from D/B) */
  void *f_w[MEEPCU_NUM_COMPONENTS][2];    /* fields_chunk::f_w      (PML, E/H integrated from W) */
  void *f_cond[MEEPCU_NUM_COMPONENTS][2]; /* fields_chunk::f_cond   (PML + conductivity) */
  const void *chi1inv[MEEPCU_NUM_COMPONENTS][5];      /* structure_chunk::chi1inv */
  const void *chi2[MEEPCU_NUM_COMPONENTS];            /* structure_chunk::chi2    */
  const void *chi3[MEEPCU_NUM_COMPONENTS];            /* structure_chunk::chi3    */
  const void *conductivity[MEEPCU_NUM_COMPONENTS][5]; /* structure_chunk::conductivity */
  const void *condinv[MEEPCU_NUM_COMPONENTS][5];      /* structure_chunk::condinv */
  const void *sig[6], *kap[6], *siginv[6];            /* structure_chunk::sig/kap/siginv (uPML) */
  int sigsize[6];                                     /* structure_chunk::sigsize */
  int rank;     /* owning rank (structure_chunk::n_proc()); single-GPU: 0 */
  int reserved[7];
  void *f_bfast[MEEPCU_NUM_COMPONENTS][2]; /* fields_chunk::f_bfast (BFAST auxiliary F, step_db.cpp:76-79) */
} meepcu_chunk_desc;

/* lorentzian_susceptibility (src/susceptibility.cpp:98-281, src/meep.hpp:262-330).  */
typedef struct {
  int ft;                     /* MEEPCU_E_STUFF or MEEPCU_H_STUFF (which pol[] list) */
  double omega_0, gamma;      /* as stored by the reference (frequency units, not angular) */
  int no_omega_0_denominator; /* 1 = Drude */
  const void *sigma[MEEPCU_NUM_COMPONENTS][5]; /* susceptibility::sigma[c][d], NULL when trivial */
  void *P[MEEPCU_NUM_COMPONENTS][2];           /* lorentzian_data::P (NULL: needs_P() false here) */
  void *P_prev[MEEPCU_NUM_COMPONENTS][2];      /* lorentzian_data::P_prev */
  double noise_amp;           /* noisy_lorentzian_susceptibility::noise_amp (src/meep.hpp:282-299); 0 = no noise term.
                                 Occupies the first two of the former reserved[8] ints: binary compatible. */
  int reserved[6];
} meepcu_susceptibility_desc;

/* gyrotropic_susceptibility (src/susceptibility.cpp:350-603, src/meep.hpp:300-338): three polarisation
 * components on EACH driving-field position (gyrotropy_data: P[c][cmp][X..Z], P_prev likewise). */
enum { MEEPCU_GYROTROPIC_LORENTZIAN = 0, MEEPCU_GYROTROPIC_DRUDE = 1, MEEPCU_GYROTROPIC_SATURATED = 2 };
typedef struct {
  int ft;                     /* MEEPCU_E_STUFF or MEEPCU_H_STUFF */
  int model;                  /* MEEPCU_GYROTROPIC_* (gyrotropy_model) */
  double omega_0, gamma, alpha;
  double gyro_tensor[3][3];   /* gyrotropic_susceptibility::gyro_tensor, as stored (realnum) */
  const void *sigma[MEEPCU_NUM_COMPONENTS][5];  /* susceptibility::sigma[c][d] */
  void *P[MEEPCU_NUM_COMPONENTS][2][3];         /* gyrotropy_data::P (NULL: needs_P() false here) */
  void *P_prev[MEEPCU_NUM_COMPONENTS][2][3];    /* gyrotropy_data::P_prev */
  int reserved[8];
} meepcu_gyrotropic_desc;

/* multilevel_susceptibility (src/multilevel-atom.cpp:28-407, src/meep.hpp:340-386): L atomic levels with populations
 * N on the CENTRED grid, T radiative transitions, each with its own polarisation P[c][cmp][t] on the field's Yee
 * positions.  update_P (:240-374) first advances N at every owned centred point from the eight-point sums of W, W_prev
 * and of P, P_prev around it, then every P as an oscillator driven by sigma * W * (population inversion averaged back
 * to the Yee position); subtract_P (:376-394) takes every P[.][.][t] off D.  The small arrays are copied at the call.
 * The reference adds the medium to EVERY chunk (structure_chunk::add_susceptibility, anisotropic_averaging.cpp:300-372;
 * sigma stays trivial where the medium is absent), so N relaxes and W is saved to W_prev everywhere: describe it for
 * every chunk, as the shim does.  E_stuff only (H_stuff: MEEPCU_EUNSUPPORTED). */
typedef struct {
  int ft;                /* MEEPCU_E_STUFF */
  int L, T;              /* multilevel_susceptibility::L, T */
  const void *Gamma;     /* realnum[L*L]  relaxation rates                                      (meep.hpp:380) */
  const void *GammaInv;  /* realnum[L*L]  multilevel_data::GammaInv = inv(1 + Gamma dt/2), as the reference's set-up
                            computed it with LAPACK (multilevel-atom.cpp:156-162): an input of the stepping path */
  const void *alpha;     /* realnum[L*T]  transition coefficients                               (meep.hpp:382) */
  const void *omega;     /* realnum[T]    transition frequencies                                (meep.hpp:383) */
  const void *gamma;     /* realnum[T]    optical loss rates                                    (meep.hpp:384) */
  const void *sigmat;    /* realnum[5*T]  per-transition sigma factors by direction             (meep.hpp:385) */
  const void *sigma[MEEPCU_NUM_COMPONENTS][5];  /* susceptibility::sigma[c][d] */
  void *const *P[MEEPCU_NUM_COMPONENTS][2];     /* multilevel_data::P[c][cmp]: T pointers (NULL: needs_P() false) */
  void *const *P_prev[MEEPCU_NUM_COMPONENTS][2];/* multilevel_data::P_prev[c][cmp] */
  void *N;               /* multilevel_data::N, realnum[ntot*L]: N[i*L + level], i on the centred grid */
  void *W_prev[MEEPCU_NUM_COMPONENTS][2];       /* fields_chunk::f_w_prev (update_eh.cpp:221-226): W = f_w ?: f of the
                                                   previous time step, saved before every E update of the chunk */
  int reserved[8];
} meepcu_multilevel_desc;

/* src_vol (src/meep_internals.hpp:49-82) attached to fields_chunk::sources[ft]. */
typedef struct {
  int ft;            /* MEEPCU_D_STUFF or MEEPCU_B_STUFF: which sources[] list it is in */
  int c;             /* src_vol::c (an E/H/D/B component; only its direction/type matter) */
  int src_time_id;   /* index into the per-step meepcu_src_sample table */
  int is_integrated; /* src_time::is_integrated */
  size_t n;          /* src_vol::num_points() */
  const ptrdiff_t *index; /* src_vol::index_at(j) */
  const double *amp;      /* src_vol::amplitude(j) as (re,im) pairs */
} meepcu_source_desc;

/* dft_chunk (src/dft.cpp:57-129,265-307, src/meep.hpp:1126-1221). */
typedef struct {
  int c;            /* dft_chunk::c */
  int is[3], ie[3]; /* dft_chunk::is / ie (ivec, X,Y,Z) */
  double s0[3], s1[3], e0[3], e1[3]; /* interpolation weights at the region's ends (X,Y,Z) */
  double dV0, dV1;
  ptrdiff_t avg1, avg2;  /* yee2cent offsets in host-array elements (0: none) */
  int include_dV_and_interp_weights, sqrt_dV_and_interp_weights;
  int decimation_factor;
  int Nomega;
  const double *omega;   /* dft_chunk::omega[Nomega] (angular: 2 pi f) */
  double scale[2];       /* dft_chunk::scale (re,im) */
  void *dft;             /* dft_chunk::dft, complex<realnum>[N*Nomega] (required) */
  int reserved[8];
} meepcu_dft_desc;

/* The values of one src_time at the four sampling points of one fields::step()
 * (src/step.cpp:64,74,96,106; src_time::update, src/meep.hpp:972-978), as (re,im). */
typedef struct {
  double current_B[2]; /* current() after update(time(),        dt): step_source(B_stuff) */
  double dipole_H[2];  /* dipole()  after update(time()+dt/2,   dt): integrated H sources  */
  double current_D[2]; /* current() after update(time()+dt/2,   dt): step_source(D_stuff) */
  double dipole_E[2];  /* dipole()  after update(time()+dt,     dt): integrated E sources  */
} meepcu_src_sample;

/* ---- lifetime ---------------------------------------------------------------------------- */
const char *meepcu_last_error(void);
const char *meepcu_backend(void); /* "cuda-sm_100a" or "oracle-cpu" */
int meepcu_abi_version(void);

int meepcu_create(const meepcu_config *cfg, meepcu_ctx **out);
int meepcu_destroy(meepcu_ctx *ctx);

/* ---- description of the problem (called once, or again after the host re-allocates) ------ */
int meepcu_add_chunk(meepcu_ctx *ctx, const meepcu_chunk_desc *desc, int *chunk_id);
int meepcu_add_susceptibility(meepcu_ctx *ctx, int chunk_id, const meepcu_susceptibility_desc *desc);
/* polarisations are updated / subtracted in the order they were added (Lorentzian and gyrotropic alike),
 * like the reference's polarization_state list (update_eh.cpp:130-131, update_pols.cpp:46) */
int meepcu_add_gyrotropic(meepcu_ctx *ctx, int chunk_id, const meepcu_gyrotropic_desc *desc);
int meepcu_add_multilevel(meepcu_ctx *ctx, int chunk_id, const meepcu_multilevel_desc *desc);
int meepcu_add_source(meepcu_ctx *ctx, int chunk_id, const meepcu_source_desc *desc);
int meepcu_add_dft(meepcu_ctx *ctx, int chunk_id, const meepcu_dft_desc *desc, int *dft_id);

/* fields_chunk::zeroes[ft] (src/boundaries.cpp:309-339): pointers into registered f arrays. */
int meepcu_set_metal(meepcu_ctx *ctx, int ft, size_t n, void *const *ptrs);

/* One (field type, connect_phase) family of local chunk-to-chunk connections, flattened over all
 * chunk pairs in the reference's send order (src/step.cpp:252-279, boundaries.cpp:347-623):
 * dst[i] / src[i] are fields_chunk::connections_in / connections_out entries.  For
 * MEEPCU_CONNECT_PHASE entries come in (re,im) pairs and phases holds n/2 complex<realnum>. */
int meepcu_add_connections(meepcu_ctx *ctx, int ft, int phase_kind, size_t n, void *const *dst,
                           const void *const *src, const void *phases);

/* Connections whose other end lives on another rank (the MPI_Isend/Irecv pairs of
 * src/mympi.cpp:107-135): the local end only.  Both sides must list entries in the same order. */
int meepcu_add_remote_send(meepcu_ctx *ctx, int ft, int phase_kind, int peer_rank, size_t n,
                           const void *const *src);
int meepcu_add_remote_recv(meepcu_ctx *ctx, int ft, int phase_kind, int peer_rank, size_t n,
                           void *const *dst, const void *phases);
/* NCCL bootstrap for >1 rank: rank 0 calls meepcu_comm_unique_id, the host distributes the
 * 128 bytes (MPI, torch.distributed, a file ...), every rank calls meepcu_comm_init. */
int meepcu_comm_unique_id(void *id128);
int meepcu_comm_init(meepcu_ctx *ctx, int rank, int nranks, const void *id128);

/* Freeze the description: allocate device mirrors, upload everything, build the tile lists. */
int meepcu_finalize(meepcu_ctx *ctx);

/* ---- the hot path ------------------------------------------------------------------------- */
/* Advance nsteps time steps == nsteps calls of fields::step() (src/step.cpp:35-140), including
 * step_db, step_source, step_boundaries, update_eh, update_pols and update_dfts.
 * samples[s*n_src_times + j] holds src_time j's values for step s (may be NULL if there are none).
 *
 * ASYNCHRONOUS: the call queues the work and returns; `samples` is copied before it returns.  The
 * device is waited for only by the calls that hand data back (meepcu_sync_*, meepcu_read_value(s),
 * meepcu_check_finite, meepcu_get_stats, meepcu_get_timers), so a caller that makes one call per time
 * step -- the reference's `fields::step()` -- keeps the GPU fed.
 *
 * The reference's per-step NaN test (src/step.cpp:138-139) is evaluated lazily: every call ends with one
 * small kernel that tests a sample of every field array and raises a flag; the NEXT meepcu_step call
 * (no wait involved) returns MEEPCU_ENONFINITE ("simulation fields are NaN or Inf") when the flag is up.
 * meepcu_check_finite tests every element and blocks. */
int meepcu_step(meepcu_ctx *ctx, int nsteps, int n_src_times, const meepcu_src_sample *samples);

/* noisy_lorentzian_susceptibility::update_P (src/susceptibility.cpp:317-339) adds, after the Lorentzian update,
 *   p[i] += gaussian_random(0, amp * sqrt(sigma[i]))
 * at every owned point of every polarisation component, drawing from the process-global Mersenne twister
 * (src/random.cpp:57-69: Knuth's polar method, libm log) in loop order.  Bit parity therefore needs the CALLER's
 * generator: before each meepcu_step the caller draws the unit Gaussians g = gaussian_random(0, 1) in the reference's
 * own order and hands them over; the device forms p[i] = (realnum)(p[i] + (0.0 + g * (double)(amp * sqrt(sigma[i])))),
 * the same roundings as gaussian_random's `mean + v1 * sqrt(-2 log s / s) * stddev`.
 * Order of one time step's n_per_step values: update_pols(H_stuff) then update_pols(E_stuff) (step.cpp:83-86,115-117);
 * in each, chunks in index order, their susceptibilities in the order they were added, FOR_COMPONENTS(c), cmp 0 then 1
 * where P[c][cmp] and sigma[c][direction(c)] exist, LOOP_OVER_VOL_OWNED(gv, c, i).
 * meepcu_get_noise_count: n_per_step of the finalized context (0: no noisy medium).  meepcu_set_noise: the values
 * for the steps of the NEXT meepcu_step call (nsteps * n_per_step doubles, copied before the call returns); a
 * context with noisy media refuses meepcu_step(nsteps) unless exactly that many steps of noise are waiting. */
int meepcu_get_noise_count(meepcu_ctx *ctx, size_t *n_per_step);
int meepcu_set_noise(meepcu_ctx *ctx, int nsteps, size_t n_per_step, const double *unit_gaussians);

/* fields::synchronize_magnetic_fields / restore_magnetic_fields (src/energy_and_flux.cpp:145-177) on
 * the device: back up B and H (with their f_u, f_w, f_cond, f_bfast; fields_chunk::backup_component
 * :97-118), advance them half a step (step_db(B), step_source(B), step_boundaries(B), update_eh(H),
 * step_boundaries(H) with `sample[j].current_B / dipole_H` = src_time j after update(time()) /
 * update(time()+dt/2)), and replace B, H by the average of the advanced and the backed-up arrays, so
 * that E and H are known at the same time.  restore puts the backups back.  The reference's
 * reference count (synchronized_magnetic_fields) stays with the caller: the library accepts
 * synchronize only when not synchronized, restore only when synchronized, and refuses meepcu_step
 * in between. */
int meepcu_synchronize_magnetic_fields(meepcu_ctx *ctx, int n_src_times, const meepcu_src_sample *sample);
int meepcu_restore_magnetic_fields(meepcu_ctx *ctx);

/* ---- coherence with the host object graph -------------------------------- */
/* device -> caller's arrays (fields, aux, P, dft).  The halo (not-owned) points that no stepping kernel
 * consumes -- B/D/P, and the E/H plane normal to each component -- are brought up to date first, so the
 * host mirror holds what step_boundaries (src/step.cpp:226-288) would have left there.
 * With more than one rank (meepcu_comm_init) that includes halos owned by OTHER ranks: the call is then
 * COLLECTIVE -- every rank of the communicator must make it, in the same order (like the reference's own
 * step_boundaries).  A rank that must sync alone passes MEEPCU_SYNC_LOCAL to meepcu_sync_subset: only
 * same-rank halos are refreshed; remote ones keep the values of the last collective sync. */
int meepcu_sync_to_host(meepcu_ctx *ctx);
/* caller's arrays -> device: fields, aux, P, dft.  Materials were uploaded by meepcu_finalize; after the
 * host changes them (phase_material, set_epsilon ...) call meepcu_sync_subset(ctx, 0, MEEPCU_SYNC_MATERIALS). */
int meepcu_sync_from_host(meepcu_ctx *ctx);
/* Partial coherence: only the named groups of arrays travel.  What the reference's DFT consumers
 * (dft_flux::flux, dft_energy, near2far, save_dft_hdf5; src/dft.cpp:444-462,533-548,630-700) need is
 * MEEPCU_SYNC_DFT alone; energy / flux-in-box probes need FIELDS. */
enum { MEEPCU_SYNC_FIELDS = 1,    /* fields_chunk::f */
       MEEPCU_SYNC_AUX = 2,       /* f_u, f_w, f_cond, f_bfast */
       MEEPCU_SYNC_POL = 4,       /* polarisation P, P_prev */
       MEEPCU_SYNC_DFT = 8,       /* dft_chunk::dft accumulators */
       MEEPCU_SYNC_MATERIALS = 16,/* chi1inv, chi2, chi3, conductivity, sigma, sig/kap/siginv (host -> device only) */
       MEEPCU_SYNC_LOCAL = 32     /* modifier: do not refresh halos owned by other ranks (not collective) */ };
int meepcu_sync_subset(meepcu_ctx *ctx, int to_host, unsigned groups);
int meepcu_get_time(meepcu_ctx *ctx, int *t);
int meepcu_set_time(meepcu_ctx *ctx, int t);
/* every element of every field array tested for NaN / Inf on the device (one launch), blocking; the
 * per-step test of src/step.cpp:138-139 itself is the lazy one described at meepcu_step */
int meepcu_check_finite(meepcu_ctx *ctx, int *ok);
/* one element of a registered host array, read from the device mirror (fields::get_field probes).
 * Not-owned points are valid too: same-rank halos the stepping kernels do not need are refreshed first
 * (halos owned by another rank are as of the last collective meepcu_sync_to_host). */
int meepcu_read_value(meepcu_ctx *ctx, const void *host_ptr, double *out);
/* n elements at once, gathered on the device and returned in one transfer: what fields::get_field's
 * interpolation stencil, get_array_slice and integrate (src/monitor.cpp, array_slice.cpp,
 * integrate.cpp) read, without a sync of whole arrays */
int meepcu_read_values(meepcu_ctx *ctx, size_t n, const void *const *host_ptrs, double *out);

/* ---- reductions for the host-side readers: no array travels ------------ */
/* out[i] = sum over the listed pairs p and their points k < N_p of
 *   Re(a_p.dft[k*Nomega+i] * conj(b_p.dft[k*Nomega+i]))        (each term in realnum, the sum in double)
 * for i < Nomega: dft_flux::flux (src/dft.cpp:533-548; a = the E list, b = the H list, in list order) and the
 * sums of dft_energy::electric / magnetic (src/dft.cpp:657-687; the caller multiplies by 0.5).  a[p], b[p] are
 * dft_chunk::dft host pointers as given to meepcu_add_dft; both chunks of a pair have the same N and Nomega.
 * The accumulators stay on the device. */
int meepcu_dft_dot(meepcu_ctx *ctx, size_t npairs, const void *const *a, const void *const *b, int Nomega,
                   double *out);

/* One piece of fields::integrate (src/integrate.cpp:47-131) with two field components: what
 * integrate_chunkloop receives from loop_in_chunks for one (chunk, symmetry copy, periodic shift). */
typedef struct {
  const void *f[2][2];   /* fields_chunk::f[cS[i]][cmp] of the two components (cmp 1 NULL: real fields) */
  ptrdiff_t off[2][2];   /* grid_volume::yee2cent_offsets of each component in host elements (0,0: own grid) */
  double ph[2][2];       /* shift_phase * S.phase_shift(cS[i], sn) as (re,im) */
  int is[3], ie[3];      /* loop bounds (ivec, X,Y,Z) */
  double s0[3], s1[3], e0[3], e1[3]; /* IVEC_LOOP_WEIGHT end weights (X,Y,Z) */
  double dV0, dV1;
  int reserved[4];
} meepcu_integrate_desc;
/* *sum    = sum over the pieces and their points of  w * Re(conj(F_0) F_1),
 * *maxabs = max |Re(conj(F_0) F_1)|  (may be NULL),
 * F_i = complex<realnum>(ph_i * 0.25 (f_i[idx] + f_i[idx+off0] + f_i[idx+off1] + f_i[idx+off0+off1])), w =
 * IVEC_LOOP_WEIGHT: fields::integrate(2, cs, dot_integrand, 0, where) of src/energy_and_flux.cpp:61-85,189-215
 * -- field_energy_in_box, electric / magnetic_energy_in_box, flux_in_box_wrongH -- evaluated on the device mirror
 * (E and H at the times the device holds them: call meepcu_synchronize_magnetic_fields around it like the
 * reference does).  Single rank: the caller adds the ranks' sums (sum_to_all), as the reference does. */
int meepcu_integrate_dot(meepcu_ctx *ctx, size_t npieces, const meepcu_integrate_desc *pieces, double *sum,
                         double *maxabs);

/* ---- introspection ------------------------------------------------------------------------ */
/* seconds accumulated per reference time_sink (src/meep.hpp:1610-1633), CUDA-event based: out[1]
 * (Stepping) is the device time of every meepcu_step call so far, always on; the per-phase sinks
 * (FieldUpdateB/D/E, Boundaries, FourierTransforming) accumulate while meepcu_set_profiling(1) is on */
int meepcu_get_timers(meepcu_ctx *ctx, double out[22]);
typedef struct {
  int64_t kernel_launches;    /* kernels of this library launched so far */
  int64_t steps;              /* time steps executed */
  int64_t device_bytes;       /* bytes of device memory owned by the context */
  int64_t cells;              /* sum over chunks of owned cells (nx*ny*nz) */
  int64_t algorithmic_bytes_per_step; /* compulsory HBM bytes per step (DESIGN.md accounting) */
  double last_step_ms;        /* device time of the last meepcu_step call (CUDA events; the query waits for it) */
  int64_t reserved[8];        /* [0] cells the NR seed ladder left unchanged; [1], [2], [6] Newton-Raphson solves, their
                                 iterations and the longest solve, counted while meepcu_set_profiling is on; [3] lazy NaN
                                 flag; [4] device microseconds of all meepcu_step calls; [5] E arrays not stored while
                                 stepping (virtual E); [7] bytes of host arrays page-locked in place */
} meepcu_stats;
int meepcu_get_stats(meepcu_ctx *ctx, meepcu_stats *out);

/* Per-kernel device times, measured with CUDA events on the library's own stream around every
 * launch while profiling is on (bench.py's roofline numbers come from here). */
typedef struct {
  char name[48];                  /* kernel (class) name, e.g. "k_half<D/E,simple>" */
  int64_t launches;               /* launches timed so far */
  double ms;                      /* summed device time of those launches */
  int64_t algorithmic_bytes;      /* compulsory HBM bytes of ONE launch (DESIGN.md accounting) */
} meepcu_kernel_time;
/* on: bit 0 = time every kernel (serialised on one stream; also resets the accumulated times), bit 1 = count the
 * Newton-Raphson solves and iterations of k_eh_general (meepcu_stats.reserved[1,2,6]; atomics: do not time then) */
int meepcu_set_profiling(meepcu_ctx *ctx, int on);
/* The same byte counts with SURVEY.md 8(d)'s per-cell figures, i.e. with every array of the reference's update moved once
 * per half step in which it is read and once where it is written (21 values per vacuum cell, 24 with a diagonal
 * epsilon: E is stored by the D/E half and read back by the B/H half).  meepcu_stats.algorithmic_bytes_per_step and
 * meepcu_kernel_time.algorithmic_bytes count what THIS implementation has to move, which is less where a chunk does
 * not store E while stepping (18 / 21).  per_kernel[i] belongs to entry i of meepcu_get_kernel_times. */
int meepcu_get_compulsory_bytes(meepcu_ctx *ctx, int64_t *per_step, int64_t *per_kernel, int capacity, int *n);
int meepcu_get_kernel_times(meepcu_ctx *ctx, meepcu_kernel_time *out, int capacity, int *n);

#ifdef __cplusplus
}
#endif
#endif /* MEEPCU_H */
