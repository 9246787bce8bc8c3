// meepcu_shim.hpp -- the reference-side binding of the C ABI in include/meepcu.h.
//
// This is the (header-only) piece a Meep maintainer adds to libmeep so that `fields::step()`
// (reference src/step.cpp:35-140) runs on the GPU: it walks the existing C++ object graph
// (fields -> fields_chunk -> structure_chunk, src/meep.hpp:590-598,1461-1485,1733-2339),
// describes it to libmeepcuda.so with plain pointers, and forwards step()/sync calls.
// It is compiled only where the reference headers exist (oracle/Makefile builds
// oracle/_ref/*/ref_driver with it); INTEGRATION.md shows the 12-line patch to step.cpp.
//
// The same shim drives oracle/liboracle_meep.so (same ABI, CPU, in place), which is how the
// oracle is pinned against the unmodified reference.
//
// Access to private members: a maintainer would declare `friend class meepcu::attachment;` in
// `fields` and `fields_chunk`; out of tree we include meep.hpp with private/protected opened.
#ifndef MEEPCU_SHIM_HPP
#define MEEPCU_SHIM_HPP

#include <algorithm>
#include <complex>
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>
#include <stdio.h>
#include <stddef.h>
#include <string.h>
#include <math.h>
#include <dlfcn.h>

#ifndef MEEPCU_SHIM_NO_OPEN_PRIVATE
#define private public
#define protected public
#endif
#include "meep.hpp"
#include "meep_internals.hpp"
#ifndef MEEPCU_SHIM_NO_OPEN_PRIVATE
#undef private
#undef protected
#endif

#include "meepcu.h"

namespace meepcu {

// ---- the ABI, resolved at run time so that libmeep does not link against CUDA ----------------
// direction of slot k of every X,Y,Z triple of the ABI: Cartesian X,Y,Z; cylindrical R,-,Z (grid_volume keeps
// num[d % 3], vec.hpp:1014-1020)
inline meep::direction slot_direction(meep::ndim dim, int k) {
  if (dim == meep::Dcyl) return k == 0 ? meep::R : (k == 2 ? meep::Z : meep::P); // (P: no extent, one loop iteration)
  return (meep::direction)k;
}
inline bool slot_present(meep::ndim dim, int k) {
  return meep::has_direction(dim, slot_direction(dim, k));
}

struct api {
  void *handle;
#define MEEPCU_FN(name) decltype(&::meepcu_##name) name;
  MEEPCU_FN(last_error)
  MEEPCU_FN(backend)
  MEEPCU_FN(abi_version)
  MEEPCU_FN(create)
  MEEPCU_FN(destroy)
  MEEPCU_FN(add_chunk)
  MEEPCU_FN(add_susceptibility)
  MEEPCU_FN(add_gyrotropic)
  MEEPCU_FN(add_multilevel)
  MEEPCU_FN(get_noise_count)
  MEEPCU_FN(set_noise)
  MEEPCU_FN(add_source)
  MEEPCU_FN(add_dft)
  MEEPCU_FN(set_metal)
  MEEPCU_FN(add_connections)
  MEEPCU_FN(finalize)
  MEEPCU_FN(step)
  MEEPCU_FN(synchronize_magnetic_fields)
  MEEPCU_FN(restore_magnetic_fields)
  MEEPCU_FN(sync_to_host)
  MEEPCU_FN(sync_from_host)
  MEEPCU_FN(sync_subset)
  MEEPCU_FN(read_values)
  MEEPCU_FN(get_time)
  MEEPCU_FN(set_time)
  MEEPCU_FN(check_finite)
  MEEPCU_FN(read_value)
  MEEPCU_FN(get_timers)
  MEEPCU_FN(get_stats)
  MEEPCU_FN(dft_dot)
  MEEPCU_FN(integrate_dot)
#undef MEEPCU_FN
  explicit api(const char *path) {
    handle = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!handle) meep::abort("meepcu: cannot load %s: %s", path, dlerror());
#define MEEPCU_FN(name)                                                                            \
  name = (decltype(name))dlsym(handle, "meepcu_" #name);                                           \
  if (!name) meep::abort("meepcu: %s lacks symbol meepcu_" #name, path);
    MEEPCU_FN(last_error)
    MEEPCU_FN(backend)
    MEEPCU_FN(abi_version)
    MEEPCU_FN(create)
    MEEPCU_FN(destroy)
    MEEPCU_FN(add_chunk)
    MEEPCU_FN(add_susceptibility)
    MEEPCU_FN(add_gyrotropic)
    MEEPCU_FN(add_multilevel)
    MEEPCU_FN(get_noise_count)
    MEEPCU_FN(set_noise)
    MEEPCU_FN(add_source)
    MEEPCU_FN(add_dft)
    MEEPCU_FN(set_metal)
    MEEPCU_FN(add_connections)
    MEEPCU_FN(finalize)
    MEEPCU_FN(step)
    MEEPCU_FN(synchronize_magnetic_fields)
    MEEPCU_FN(restore_magnetic_fields)
    MEEPCU_FN(sync_to_host)
    MEEPCU_FN(sync_from_host)
    MEEPCU_FN(sync_subset)
    MEEPCU_FN(read_values)
    MEEPCU_FN(get_time)
    MEEPCU_FN(set_time)
    MEEPCU_FN(check_finite)
    MEEPCU_FN(read_value)
    MEEPCU_FN(get_timers)
    MEEPCU_FN(get_stats)
    MEEPCU_FN(dft_dot)
    MEEPCU_FN(integrate_dot)
#undef MEEPCU_FN
    if (abi_version() != MEEPCU_ABI_VERSION) meep::abort("meepcu: ABI version mismatch in %s", path);
  }
};

// Optional observer of every ABI call (oracle/trace_recorder.hpp records golden traces with it).
struct observer {
  virtual ~observer() {}
  virtual void on_create(const meepcu_config &) {}
  virtual void on_chunk(int, const meepcu_chunk_desc &, size_t /*ntot*/) {}
  virtual void on_susceptibility(int, const meepcu_susceptibility_desc &, size_t /*ntot*/) {}
  virtual void on_gyrotropic(int, const meepcu_gyrotropic_desc &, size_t /*ntot*/) {}
  virtual void on_multilevel(int, const meepcu_multilevel_desc &, size_t /*ntot*/) {}
  virtual void on_noise(int /*nsteps*/, size_t /*n_per_step*/, const double *) {}
  virtual void on_source(int, const meepcu_source_desc &) {}
  virtual void on_dft(int, const meepcu_dft_desc &, size_t /*N*/) {}
  virtual void on_metal(int, size_t, void *const *) {}
  virtual void on_connections(int, int, size_t, void *const *, const void *const *, const void *) {}
  virtual void on_finalize() {}
  virtual void on_step(int, int, const meepcu_src_sample *) {}
};

// One attachment = one meepcu context mirroring one `fields` object.
class attachment {
public:
  attachment(meep::fields &f_, api &a_, int nl_mode = MEEPCU_NL_FORK_LITERAL, int device = 0,
             observer *obs_ = 0)
      : f(f_), a(a_), ctx(0), obs(obs_) {
    describe(nl_mode, device);
    live().insert(this);
  }
  ~attachment() {
    live().erase(this);
    if (ctx) a.destroy(ctx);
  }
  // every attachment alive in this process: dft_chunk objects have no way back to their `fields`, so the
  // DFT readers / writers of the reference (dft_flux::flux, save_dft_hdf5, scale_dft ...) address all of them
  static std::set<attachment *> &live() {
    static std::set<attachment *> s;
    return s;
  }
  static void for_each_live(void (attachment::*fn)()) {
    std::vector<attachment *> v(live().begin(), live().end());
    for (size_t k = 0; k < v.size(); ++k) (v[k]->*fn)();
  }

  // == fields::step() repeated n times (src/step.cpp:35-140), fields stay on the device
  void step(int n = 1) {
    const int nsrc = (int)src_times.size();
    std::vector<meepcu_src_sample> smp((size_t)n * std::max(nsrc, 1));
    for (int s = 0; s < n; ++s) {
      const double tm = (f.t + s) * f.dt; // fields::time()
      for (int j = 0; j < nsrc; ++j) {
        meep::src_time *st = src_times[j];
        meepcu_src_sample &o = smp[(size_t)s * nsrc + j];
        st->update(tm, f.dt); // step.cpp:64
        put(o.current_B, st->current());
        st->update(tm + 0.5 * f.dt, f.dt); // step.cpp:74,96
        put(o.dipole_H, st->dipole());
        put(o.current_D, st->current());
        st->update(tm + f.dt, f.dt); // step.cpp:106
        put(o.dipole_E, st->dipole());
      }
    }
    if (n_noise) { // noisy Lorentzian media: this process's generator, the reference's draw order (include/meepcu.h)
      std::vector<double> g((size_t)n * n_noise);
      for (size_t k = 0; k < g.size(); ++k) g[k] = meep::gaussian_random(0, 1);
      if (obs) obs->on_noise(n, n_noise, g.data());
      check(a.set_noise(ctx, n, n_noise, g.data()));
    }
    if (obs) obs->on_step(n, nsrc, smp.data());
    if (dft_host_dirty) { // scale_dft / operator-= / load_dft_hdf5 changed the accumulators on the host
      check(a.sync_subset(ctx, 0, MEEPCU_SYNC_DFT));
      dft_host_dirty = false;
    }
    check(a.step(ctx, n, nsrc, nsrc ? smp.data() : 0)); // queues the work and returns (include/meepcu.h)
    f.t += n;
    host_stale = true;
    dft_stale = true;
  }
  // wait for the queued steps (what a blocking fields::step() would have done)
  void wait() {
    meepcu_stats s;
    check(a.get_stats(ctx, &s));
  }
  bool stale() const { return host_stale; }
  // one element of a registered array, read from the device mirror (waits for the queued steps)
  double read(const meep::realnum *p) {
    double v = 0;
    check(a.read_value(ctx, p, &v));
    return v;
  }
  // ---- reductions on the device for the reference's readers (include/meepcu.h) ----------------------------
  // dft_flux::flux / dft_energy::electric, magnetic (dft.cpp:533-548,657-687): F[i] = sum over the chunk pairs of the
  // two lists of Re(a.dft * conj(b.dft)).  true: computed on the device (the accumulators live there and the host
  // copy is out of date); false: the host arrays are current (or belong to no attachment) -- the caller runs its loop.
  static size_t &device_reductions(int which) { // [0] dft_dot, [1] integrate_dot calls answered by the device (tests)
    static size_t n[2] = {0, 0};
    return n[which];
  }
  static bool dft_dot_on_device(meep::dft_chunk *A, meep::dft_chunk *B, size_t Nfreq, double *F) {
    if (!A || !B) return false;
    for (std::set<attachment *>::iterator it = live().begin(); it != live().end(); ++it) {
      attachment *att = *it;
      if (!att->dft_ptrs.count(A->dft)) continue;
      if (!att->dft_stale || att->dft_host_dirty) return false;
      std::vector<const void *> pa, pb;
      for (meep::dft_chunk *ca = A, *cb = B; ca && cb; ca = ca->next_in_dft, cb = cb->next_in_dft) {
        if (!att->dft_ptrs.count(ca->dft) || !att->dft_ptrs.count(cb->dft) || ca->N != cb->N ||
            ca->omega.size() != Nfreq || cb->omega.size() != Nfreq)
          return false;
        pa.push_back(ca->dft);
        pb.push_back(cb->dft);
      }
      att->check(att->a.dft_dot(att->ctx, pa.size(), pa.data(), pb.data(), (int)Nfreq, F));
      ++device_reductions(0);
      return true;
    }
    return false;
  }
  // fields::integrate(2, cs, dot_integrand, 0, where) (integrate.cpp:133-206, energy_and_flux.cpp:61-85,189-215):
  // loop_in_chunks enumerates the pieces exactly as it does for integrate_chunkloop, the sums are formed on the device
  struct integrate_collect {
    const meep::component *cs;
    std::vector<meepcu_integrate_desc> pieces;
  };
  static void integrate_collect_chunkloop(meep::fields_chunk *fc, int, meep::component cgrid, meep::ivec is, meep::ivec ie,
                                          meep::vec s0, meep::vec s1, meep::vec e0, meep::vec e1, double dV0, double dV1,
                                          meep::ivec, std::complex<double> shift_phase, const meep::symmetry &S, int sn,
                                          void *data_) {
    using namespace meep;
    integrate_collect *data = (integrate_collect *)data_;
    direction xyz[3];
    for (int k = 0; k < 3; ++k) xyz[k] = slot_direction(fc->gv.dim, k);
    meepcu_integrate_desc d;
    memset(&d, 0, sizeof d);
    for (int i = 0; i < 2; ++i) {
      const component c = S.transform(data->cs[i], -sn);
      ptrdiff_t o0 = 0, o1 = 0;
      if (cgrid == Centered) fc->gv.yee2cent_offsets(c, o0, o1);
      d.off[i][0] = o0;
      d.off[i][1] = o1;
      const std::complex<double> ph = shift_phase * S.phase_shift(c, sn);
      d.ph[i][0] = ph.real();
      d.ph[i][1] = ph.imag();
      d.f[i][0] = fc->f[c][0];
      d.f[i][1] = fc->f[c][1];
    }
    for (int k = 0; k < 3; ++k) {
      const bool has = slot_present(fc->gv.dim, k);
      d.is[k] = has ? is.in_direction(xyz[k]) : 0;
      d.ie[k] = has ? ie.in_direction(xyz[k]) : 0;
      // (every slot: the loop level of an absent direction has one iteration whose IVEC_LOOP_WEIGHT factor is
      //  s0.in_direction(that direction), vec.hpp:372-383)
      d.s0[k] = s0.in_direction(xyz[k]);
      d.s1[k] = s1.in_direction(xyz[k]);
      d.e0[k] = e0.in_direction(xyz[k]);
      d.e1[k] = e1.in_direction(xyz[k]);
    }
    d.dV0 = dV0;
    d.dV1 = dV1;
    data->pieces.push_back(d);
  }
  bool integrate_dot(const meep::component *cs, const meep::volume &where, double *out) {
    using namespace meep;
    bool same_grid = f.gv.iyee_shift(cs[1]) == f.gv.iyee_shift(cs[0]); // integrate.cpp:136-146
    integrate_collect data;
    data.cs = cs;
#ifdef MEEPCU_DROPIN_PATCHED
    f.gpu_collecting = true; // the patched loop_in_chunks only enumerates here: no fetch of the device's state
#endif
    f.loop_in_chunks(integrate_collect_chunkloop, (void *)&data, where, same_grid ? cs[0] : Centered);
#ifdef MEEPCU_DROPIN_PATCHED
    f.gpu_collecting = false;
#endif
    double sum = 0;
    check(a.integrate_dot(ctx, data.pieces.size(), data.pieces.data(), &sum, 0));
    *out = sum_to_all(sum);
    ++device_reductions(1);
    return true;
  }
  // == the body of fields::synchronize_magnetic_fields / restore_magnetic_fields
  //    (src/energy_and_flux.cpp:145-177) on the device; the caller keeps the reference's counter
  //    `synchronized_magnetic_fields` and calls these on its 0 -> 1 and 1 -> 0 transitions
  void synchronize_magnetic_fields() {
    const int nsrc = (int)src_times.size();
    std::vector<meepcu_src_sample> smp(std::max(nsrc, 1));
    memset(smp.data(), 0, smp.size() * sizeof(meepcu_src_sample));
    const double tm = f.t * f.dt;
    for (int j = 0; j < nsrc; ++j) {
      meep::src_time *st = src_times[j];
      st->update(tm, f.dt); // calc_sources(time()): B sources
      put(smp[j].current_B, st->current());
      st->update(tm + 0.5 * f.dt, f.dt); // calc_sources(time() + 0.5 dt): integrated H sources
      put(smp[j].dipole_H, st->dipole());
    }
    check(a.synchronize_magnetic_fields(ctx, nsrc, nsrc ? smp.data() : 0));
    host_stale = true;
  }
  void restore_magnetic_fields() {
    check(a.restore_magnetic_fields(ctx));
    host_stale = true;
  }
  void sync_to_host() {
    check(a.sync_to_host(ctx));
    host_stale = false;
    dft_stale = false;
  }
  // Has the host object graph changed since it was described (a source, DFT monitor or polarisation added,
  // arrays re-allocated: the reference marks the latter with chunk_connections_valid = false,
  // boundaries.cpp:171-183)?  Then the owner drops this attachment and makes a new one.
  bool description_outdated() const { return !f.chunk_connections_valid || fingerprint() != described; }
  // for readers that may be called many times between steps (fields::get_field, loop_in_chunks ...)
  void sync_to_host_if_stale() {
    if (host_stale) sync_to_host();
  }
  // only the DFT accumulators: all that dft_flux::flux / dft_energy / near2far / save_dft_hdf5 read
  void sync_dft_to_host() {
    check(a.sync_subset(ctx, 1, MEEPCU_SYNC_DFT));
    dft_stale = false;
  }
  void sync_dft_to_host_if_stale() {
    if (dft_stale) sync_dft_to_host();
  }
  void dft_will_be_written() { // a host-side writer of dft_chunk::dft: fetch now, send back before the next step
    sync_dft_to_host_if_stale();
    dft_host_dirty = true;
  }
  // fields only (get_array_slice, energy_in_box, flux_in_box ...)
  void sync_fields_to_host() { check(a.sync_subset(ctx, 1, MEEPCU_SYNC_FIELDS)); }
  void sync_from_host() {
    check(a.set_time(ctx, f.t));
    check(a.sync_from_host(ctx));
  }
  bool finite() {
    int ok = 1;
    check(a.check_finite(ctx, &ok));
    return ok != 0;
  }
  meepcu_stats stats() {
    meepcu_stats s;
    check(a.get_stats(ctx, &s));
    return s;
  }
  meepcu_ctx *context() { return ctx; }

  // Perform, eagerly and without stepping, every allocation the reference makes lazily inside its
  // first step (step_db.cpp:67-75, update_eh.cpp:204-216, update_pols.cpp:49-55), then reconnect
  // the chunks (boundaries.cpp:171-183).  After this the set of arrays is stable.
  static void allocate_lazily_created_arrays(meep::fields &f) {
    using namespace meep;
    const int is_real = f.is_real; // used by the reference's DOCMP macro
    for (int i = 0; i < f.num_chunks; i++) f.chunks[i]->s->update_condinv(); // step.cpp:61-62
    for (int i = 0; i < f.num_chunks; i++) {
      fields_chunk *fc = f.chunks[i];
      if (!fc->is_mine()) continue;
      structure_chunk *s = fc->s;
      const grid_volume &gv = fc->gv;
      const size_t ntot = gv.ntot();
      field_type dbs[2] = {B_stuff, D_stuff};
      for (int k = 0; k < 2; ++k) { // step_db.cpp:48-75
        field_type ft = dbs[k];
        DOCMP FOR_FT_COMPONENTS(ft, cc) {
          if (!fc->f[cc][cmp]) continue;
          const direction d_c = component_direction(cc);
          const direction dsig0 = cycle_direction(gv.dim, d_c, 1);
          const direction dsig = s->sigsize[dsig0] > 1 ? dsig0 : NO_DIRECTION;
          const direction dsigu0 = cycle_direction(gv.dim, d_c, 2);
          const direction dsigu = s->sigsize[dsigu0] > 1 ? dsigu0 : NO_DIRECTION;
          if (dsig != NO_DIRECTION && s->conductivity[cc][d_c] && !fc->f_cond[cc][cmp]) {
            fc->f_cond[cc][cmp] = new realnum[ntot];
            memset(fc->f_cond[cc][cmp], 0, sizeof(realnum) * ntot);
          }
          if (dsigu != NO_DIRECTION && !fc->f_u[cc][cmp]) {
            fc->f_u[cc][cmp] = new realnum[ntot];
            memcpy(fc->f_u[cc][cmp], fc->f[cc][cmp], ntot * sizeof(realnum));
          }
          const bool use_bfast = fc->bfast_scaled_k.size() >= 3 && // step_db.cpp:65,76-79
                                 (fc->bfast_scaled_k[0] || fc->bfast_scaled_k[1] || fc->bfast_scaled_k[2]);
          if (use_bfast && !fc->f_bfast[cc][cmp]) {
            fc->f_bfast[cc][cmp] = new realnum[ntot];
            memset(fc->f_bfast[cc][cmp], 0, sizeof(realnum) * ntot);
          }
        }
      }
      field_type ehs[2] = {H_stuff, E_stuff};
      for (int k = 0; k < 2; ++k) { // update_eh.cpp:67-226
        field_type ft = ehs[k];
        field_type ft2 = ft == E_stuff ? D_stuff : B_stuff;
        bool have_int_sources = false;
        for (const src_vol &sv : fc->sources[ft2])
          if (sv.t()->is_integrated) have_int_sources = true;
        bool have_f_minus_p = false;
        FOR_FT_COMPONENTS(ft, ec) DOCMP {
          if (!fc->f[ec][cmp]) continue;
          bool need = have_int_sources;
          for (polarization_state *p = fc->pol[ft]; p && !need; p = p->next)
            need = need || p->s->needs_P(ec, cmp, fc->f);
          if (need && cmp == 0) have_f_minus_p = true;
        }
        DOCMP FOR_FT_COMPONENTS(ft, ec) {
          if (!fc->f[ec][cmp]) continue;
          component dc = field_type_component(ft2, ec);
          const direction d_ec = component_direction(ec);
          const direction dsigw = s->sigsize[d_ec] > 1 ? d_ec : NO_DIRECTION;
          if (fc->f[ec][cmp] == fc->f[dc][cmp] &&
              (s->chi1inv[ec][d_ec] || have_f_minus_p || dsigw != NO_DIRECTION)) {
            fc->f[ec][cmp] = new realnum[ntot];
            memcpy(fc->f[ec][cmp], fc->f[dc][cmp], ntot * sizeof(realnum));
          }
          if (!fc->f_w[ec][cmp] && dsigw != NO_DIRECTION) {
            fc->f_w[ec][cmp] = new realnum[ntot];
            memcpy(fc->f_w[ec][cmp], fc->f[ec][cmp], ntot * sizeof(realnum));
          }
          if (fc->needs_W_prev(ec)) { // update_eh.cpp:221-226 (multilevel atoms; every chunk of such a structure)
            if (ft == H_stuff) meep::abort("meepcu: magnetic multilevel media are outside the GPU path");
            if (!fc->f_w_prev[ec][cmp]) {
              fc->f_w_prev[ec][cmp] = new realnum[ntot];
              memcpy(fc->f_w_prev[ec][cmp], fc->f_w[ec][cmp] ? fc->f_w[ec][cmp] : fc->f[ec][cmp], sizeof(realnum) * ntot);
            }
          }
        }
        for (polarization_state *p = fc->pol[ft]; p; p = p->next) { // update_pols.cpp:49-55
          if (!p->data) {
            p->data = p->s->new_internal_data(fc->f, gv);
            if (p->data) p->s->init_internal_data(fc->f, fc->dt, gv, p->data);
          }
        }
      }
    }
    f.chunk_connections_valid = false;
    f.connect_chunks();
  }

private:
  // counts of everything the description enumerates
  std::vector<size_t> fingerprint() const {
    std::vector<size_t> fp;
    size_t nst = 0;
    for (meep::src_time *st = f.sources; st; st = st->next) ++nst;
    fp.push_back(nst);
    for (int i = 0; i < f.num_chunks; i++) {
      meep::fields_chunk *fc = f.chunks[i];
      size_t nd = 0, np = 0;
      for (meep::dft_chunk *dc = fc->dft_chunks; dc; dc = dc->next_in_chunk) ++nd;
      for (int ft = 0; ft < 2; ++ft)
        for (meep::polarization_state *p = fc->pol[ft]; p; p = p->next) ++np;
      fp.push_back(nd);
      fp.push_back(np);
      for (int ft = 0; ft < meep::NUM_FIELD_TYPES; ++ft) {
        fp.push_back(fc->sources[ft].size());
        for (const meep::src_vol &sv : fc->sources[ft]) fp.push_back(sv.num_points());
      }
    }
    return fp;
  }
  std::vector<size_t> described;
  meep::fields &f;
  api &a;
  meepcu_ctx *ctx;
  observer *obs;
  bool host_stale = false; // the device has advanced since the last sync_to_host
  bool dft_stale = false;  // ... since the DFT accumulators were last fetched
  bool dft_host_dirty = false; // the host changed the accumulators since they were last sent
  std::set<const void *> dft_ptrs; // dft_chunk::dft arrays registered with the context
  std::vector<meep::src_time *> src_times;
  size_t n_noise = 0; // unit Gaussians the context wants per time step (noisy Lorentzian media)
  std::vector<std::vector<double> > keep_amp; // storage handed to the ABI during description

  static void put(double o[2], std::complex<double> z) {
    o[0] = z.real();
    o[1] = z.imag();
  }
  void check(int rc) {
    if (rc == MEEPCU_ENONFINITE) meep::abort("simulation fields are NaN or Inf"); // step.cpp:138-139, the reference's words
    if (rc != MEEPCU_OK) meep::abort("meepcu (%s): %s", a.backend(), a.last_error());
  }

  void describe(int nl_mode, int device) {
    using namespace meep;
    if (f.is_phasing()) meep::abort("meepcu: phase_material is outside the GPU path");
    for (int i = 0; i < f.num_chunks; i++)
      if (!f.chunks[i]->is_mine())
        meep::abort("meepcu: MPI-distributed chunks: use one process and several GPUs instead");
    allocate_lazily_created_arrays(f);

    meepcu_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.abi_version = MEEPCU_ABI_VERSION;
    cfg.realnum_bytes = (int)sizeof(realnum);
    cfg.is_real = f.is_real;
    cfg.dim = (int)f.gv.dim;
    cfg.a = f.a;
    cfg.dt = f.dt;
    cfg.courant = f.num_chunks ? f.chunks[0]->Courant : f.dt * f.a;
    cfg.beta = f.beta;
    cfg.m = f.m;
    cfg.zero_fields_near_cylorigin = f.num_chunks ? f.chunks[0]->zero_fields_near_cylorigin : 1; // a fields_chunk member (meep.hpp:1504)
    for (int k = 0; k < 3 && k < (int)f.bfast_scaled_k.size(); ++k) cfg.bfast_scaled_k[k] = f.bfast_scaled_k[k];
    cfg.nl_mode = nl_mode;
    cfg.device = device;
    cfg.t = f.t;
    if (obs) obs->on_create(cfg);
    check(a.create(&cfg, &ctx));

    // src_time list -> ids (fields::sources, src/meep.hpp:1736)
    std::map<const src_time *, int> src_id;
    for (src_time *s = f.sources; s; s = s->next) {
      src_id[s] = (int)src_times.size();
      src_times.push_back(s);
    }

    for (int i = 0; i < f.num_chunks; i++) {
      fields_chunk *fc = f.chunks[i];
      structure_chunk *s = fc->s;
      const grid_volume &gv = fc->gv;
      meepcu_chunk_desc d;
      memset(&d, 0, sizeof d);
      direction xyz[3];
      for (int k = 0; k < 3; ++k) {
        xyz[k] = slot_direction(gv.dim, k);
        d.num[k] = slot_present(gv.dim, k) ? gv.num_direction(xyz[k]) : 0;
        d.io[k] = slot_present(gv.dim, k) ? gv.little_corner().in_direction(xyz[k]) : 0;
      }
      FOR_COMPONENTS(c) {
        for (int cmp = 0; cmp < 2; ++cmp) {
          d.f[c][cmp] = fc->f[c][cmp];
          d.f_u[c][cmp] = fc->f_u[c][cmp];
          d.f_w[c][cmp] = fc->f_w[c][cmp];
          d.f_cond[c][cmp] = fc->f_cond[c][cmp];
          d.f_bfast[c][cmp] = fc->f_bfast[c][cmp];
        }
        for (int dd = 0; dd < 5; ++dd) {
          d.chi1inv[c][dd] = s->chi1inv[c][dd];
          d.conductivity[c][dd] = s->conductivity[c][dd];
          d.condinv[c][dd] = s->condinv[c][dd];
        }
        d.chi2[c] = s->chi2[c];
        d.chi3[c] = s->chi3[c];
      }
      for (int dd = 0; dd < 6; ++dd) {
        d.sig[dd] = s->sig[dd];
        d.kap[dd] = s->kap[dd];
        d.siginv[dd] = s->siginv[dd];
        d.sigsize[dd] = s->sigsize[dd];
      }
      d.rank = 0;
      int id = -1;
      if (obs) obs->on_chunk(i, d, gv.ntot());
      check(a.add_chunk(ctx, &d, &id));
      if (id != i) meep::abort("meepcu: chunk numbering mismatch");

      // polarizations, in list order (update_eh.cpp:130-131, update_pols.cpp:46)
      field_type ehs[2] = {E_stuff, H_stuff};
      for (int k = 0; k < 2; ++k)
        for (polarization_state *p = fc->pol[ehs[k]]; p; p = p->next) {
          if (typeid(*p->s) == typeid(gyrotropic_susceptibility)) {
            const gyrotropic_susceptibility *gs = static_cast<const gyrotropic_susceptibility *>(p->s);
            // gyrotropy_data is private to susceptibility.cpp:373-379; same layout restated here
            struct gyrotropy_data_layout {
              size_t sz_data, ntot;
              realnum *P[NUM_FIELD_COMPONENTS][2][3];
              realnum *P_prev[NUM_FIELD_COMPONENTS][2][3];
            };
            const gyrotropy_data_layout *gd = (const gyrotropy_data_layout *)p->data;
            meepcu_gyrotropic_desc gdsc;
            memset(&gdsc, 0, sizeof gdsc);
            gdsc.ft = (int)ehs[k];
            gdsc.model = (int)gs->model;
            gdsc.omega_0 = gs->omega_0;
            gdsc.gamma = gs->gamma;
            gdsc.alpha = gs->alpha;
            for (int r = 0; r < 3; ++r)
              for (int q = 0; q < 3; ++q) gdsc.gyro_tensor[r][q] = gs->gyro_tensor[r][q];
            FOR_COMPONENTS(c) {
              for (int dd = 0; dd < 5; ++dd) gdsc.sigma[c][dd] = gs->sigma[c][dd];
              if (gd)
                for (int cmp = 0; cmp < 2; ++cmp)
                  for (int dd = 0; dd < 3; ++dd) {
                    gdsc.P[c][cmp][dd] = gd->P[c][cmp][dd];
                    gdsc.P_prev[c][cmp][dd] = gd->P_prev[c][cmp][dd];
                  }
            }
            if (obs) obs->on_gyrotropic(i, gdsc, gv.ntot());
            check(a.add_gyrotropic(ctx, i, &gdsc));
            continue;
          }
          if (typeid(*p->s) == typeid(multilevel_susceptibility)) {
            const multilevel_susceptibility *ms = static_cast<const multilevel_susceptibility *>(p->s);
            // multilevel_data is private to multilevel-atom.cpp:118-128; same layout restated here
            struct multilevel_data_layout {
              size_t sz_data, ntot;
              realnum *GammaInv;
              realnum **P[NUM_FIELD_COMPONENTS][2];
              realnum **P_prev[NUM_FIELD_COMPONENTS][2];
              realnum *N, *Ntmp;
            };
            const multilevel_data_layout *md = (const multilevel_data_layout *)p->data;
            if (!md) meep::abort("meepcu: multilevel medium without internal data");
            meepcu_multilevel_desc mdsc;
            memset(&mdsc, 0, sizeof mdsc);
            mdsc.ft = (int)ehs[k];
            mdsc.L = ms->L;
            mdsc.T = ms->T;
            mdsc.Gamma = ms->Gamma;
            mdsc.GammaInv = md->GammaInv;
            mdsc.alpha = ms->alpha;
            mdsc.omega = ms->omega;
            mdsc.gamma = ms->gamma;
            mdsc.sigmat = ms->sigmat;
            mdsc.N = md->N;
            FOR_COMPONENTS(c) {
              for (int dd = 0; dd < 5; ++dd) mdsc.sigma[c][dd] = ms->sigma[c][dd];
              for (int cmp = 0; cmp < 2; ++cmp) {
                mdsc.P[c][cmp] = (void *const *)md->P[c][cmp];
                mdsc.P_prev[c][cmp] = (void *const *)md->P_prev[c][cmp];
                mdsc.W_prev[c][cmp] = fc->f_w_prev[c][cmp];
              }
            }
            if (obs) obs->on_multilevel(i, mdsc, gv.ntot());
            check(a.add_multilevel(ctx, i, &mdsc));
            continue;
          }
          const lorentzian_susceptibility *ls =
              dynamic_cast<const lorentzian_susceptibility *>(p->s);
          const bool noisy = typeid(*p->s) == typeid(noisy_lorentzian_susceptibility);
          if (!ls || (typeid(*p->s) != typeid(lorentzian_susceptibility) && !noisy))
            meep::abort("meepcu: only Lorentzian/Drude (plain or noisy), gyrotropic and multilevel susceptibilities are on the GPU path");
          meepcu_susceptibility_desc sd;
          memset(&sd, 0, sizeof sd);
          sd.ft = (int)ehs[k];
          if (noisy) sd.noise_amp = static_cast<const noisy_lorentzian_susceptibility *>(p->s)->noise_amp;
          sd.omega_0 = ls->omega_0;
          sd.gamma = ls->gamma;
          sd.no_omega_0_denominator = ls->no_omega_0_denominator;
          FOR_COMPONENTS(c) {
            for (int dd = 0; dd < 5; ++dd) sd.sigma[c][dd] = ls->sigma[c][dd];
            if (p->data)
              for (int cmp = 0; cmp < 2; ++cmp) {
                // lorentzian_data layout, susceptibility.cpp:98-139: P then P_prev, ntot apart
                realnum *P = ls->cinternal_notowned_ptr(0, c, cmp, 0, p->data);
                sd.P[c][cmp] = P;
                sd.P_prev[c][cmp] = P ? P + gv.ntot() : 0;
              }
          }
          if (obs) obs->on_susceptibility(i, sd, gv.ntot());
          check(a.add_susceptibility(ctx, i, &sd));
        }

      // sources (fields_chunk::sources[D_stuff|B_stuff])
      field_type dbs[2] = {B_stuff, D_stuff};
      for (int k = 0; k < 2; ++k)
        for (const src_vol &sv : fc->sources[dbs[k]]) {
          meepcu_source_desc sd;
          memset(&sd, 0, sizeof sd);
          sd.ft = (int)dbs[k];
          sd.c = (int)sv.c;
          sd.src_time_id = src_id.at(sv.t());
          sd.is_integrated = sv.t()->is_integrated;
          sd.n = sv.num_points();
          std::vector<ptrdiff_t> idx(sd.n);
          std::vector<double> amp(2 * sd.n);
          for (size_t j = 0; j < sd.n; ++j) {
            idx[j] = sv.index_at(j);
            amp[2 * j] = sv.amplitude(j).real();
            amp[2 * j + 1] = sv.amplitude(j).imag();
          }
          sd.index = idx.data();
          sd.amp = amp.data();
          if (obs) obs->on_source(i, sd);
          check(a.add_source(ctx, i, &sd));
        }

      // DFT chunks (fields_chunk::dft_chunks, dft.cpp:57-129)
      for (dft_chunk *dc = fc->dft_chunks; dc; dc = dc->next_in_chunk) {
        meepcu_dft_desc dd;
        memset(&dd, 0, sizeof dd);
        dd.c = (int)dc->c;
        for (int k = 0; k < 3; ++k) {
          const bool has = slot_present(gv.dim, k);
          dd.is[k] = has ? dc->is.in_direction(xyz[k]) : 0;
          dd.ie[k] = has ? dc->ie.in_direction(xyz[k]) : 0;
          dd.s0[k] = dc->s0.in_direction(xyz[k]); // (every slot: see integrate_collect_chunkloop)
          dd.s1[k] = dc->s1.in_direction(xyz[k]);
          dd.e0[k] = dc->e0.in_direction(xyz[k]);
          dd.e1[k] = dc->e1.in_direction(xyz[k]);
        }
        dd.dV0 = dc->dV0;
        dd.dV1 = dc->dV1;
        dd.avg1 = dc->avg1;
        dd.avg2 = dc->avg2;
        dd.include_dV_and_interp_weights = dc->include_dV_and_interp_weights;
        dd.sqrt_dV_and_interp_weights = dc->sqrt_dV_and_interp_weights;
        dd.decimation_factor = dc->get_decimation_factor();
        dd.Nomega = (int)dc->omega.size();
        dd.omega = dc->omega.data();
        dd.scale[0] = dc->scale.real();
        dd.scale[1] = dc->scale.imag();
        dd.dft = dc->dft;
        dft_ptrs.insert(dc->dft);
        if (obs) obs->on_dft(i, dd, dc->N);
        check(a.add_dft(ctx, i, &dd, 0));
      }
    }

    // metal points and chunk connections, per field type (boundaries.cpp:309-623)
    FOR_FIELD_TYPES(ft) {
      std::vector<void *> z;
      for (int i = 0; i < f.num_chunks; i++)
        for (size_t k = 0; k < f.chunks[i]->num_zeroes[ft]; ++k)
          z.push_back(f.chunks[i]->zeroes[ft][k]);
      if (obs) obs->on_metal((int)ft, z.size(), z.data());
      check(a.set_metal(ctx, (int)ft, z.size(), z.data()));

      for (connect_phase ip : all_connect_phases) {
        std::vector<void *> dst;
        std::vector<const void *> src;
        std::vector<std::complex<realnum> > ph;
        // same order as step_boundaries' send loop (step.cpp:252-268)
        for (const comms_operation &op : f.comms_sequence_for_field[ft].send_ops) {
          const chunk_pair pair{(int)op.my_chunk_idx, (int)op.other_chunk_idx};
          const comms_key key = {ft, ip, pair};
          if (!f.get_comm_size(key)) continue;
          const std::vector<realnum *> &out = f.chunks[pair.first]->connections_out.at(key);
          const std::vector<realnum *> &in = f.chunks[pair.second]->connections_in.at(key);
          if (out.size() != in.size()) meep::abort("meepcu: connection size mismatch");
          for (size_t n = 0; n < out.size(); ++n) {
            src.push_back(out[n]);
            dst.push_back(in[n]);
          }
          if (ip == CONNECT_PHASE) {
            const std::vector<std::complex<realnum> > &p =
                f.chunks[pair.second]->connection_phases.at(key);
            ph.insert(ph.end(), p.begin(), p.end());
          }
        }
        if (dst.empty()) continue;
        if (obs) obs->on_connections((int)ft, (int)ip, dst.size(), dst.data(), src.data(), ph.data());
        check(a.add_connections(ctx, (int)ft, (int)ip, dst.size(), dst.data(), src.data(),
                                ph.data()));
      }
    }
    if (obs) obs->on_finalize();
    check(a.finalize(ctx));
    check(a.get_noise_count(ctx, &n_noise));
    described = fingerprint();
  }
};

} // namespace meepcu
#endif
