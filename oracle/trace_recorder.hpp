// trace_recorder.hpp -- TEST INFRASTRUCTURE.  Records every C-ABI call the reference-side shim
// (meep_nl_b200/shim/meepcu_shim.hpp) makes for one `fields` object, with the initial contents of
// every array, plus the arrays the unmodified reference produced after the same number of steps.
// tests/trace.py replays such a file against libmeepcuda.so / liboracle_meep.so on machines
// that do not have the reference.  Format: see tests/trace.py (the reader is the spec).
#ifndef TRACE_RECORDER_HPP
#define TRACE_RECORDER_HPP

#include "../meep_nl_b200/shim/meepcu_shim.hpp"
#include <stdint.h>

namespace meepcu {

class trace_recorder : public observer {
public:
  enum { T_CONFIG = 1, T_ARRAY, T_CHUNK, T_SUSC, T_SOURCE, T_DFT, T_METAL, T_CONN, T_FINALIZE,
         T_STEP, T_EXPECT, T_GYRO, T_MLATOM, T_NOISE };

  explicit trace_recorder(const char *path) : rb(8) {
    fp = fopen(path, "wb");
    if (!fp) meep::abort("cannot open trace file %s", path);
    fwrite("MCUTRC01", 1, 8, fp);
  }
  ~trace_recorder() {
    if (fp) fclose(fp);
  }
  void close() {
    if (fp) fclose(fp);
    fp = 0;
  }

  // ---- observer ----
  void on_create(const meepcu_config &c) {
    rb = c.realnum_bytes;
    begin(T_CONFIG);
    put(&c, sizeof c);
    end();
  }
  void on_chunk(int i, const meepcu_chunk_desc &d, size_t ntot) {
    std::vector<int64_t> ids;
    for (int c = 0; c < 20; ++c)
      for (int k = 0; k < 2; ++k) ids.push_back(reg(d.f[c][k], ntot, rb));
    for (int c = 0; c < 20; ++c)
      for (int k = 0; k < 2; ++k) ids.push_back(reg(d.f_u[c][k], ntot, rb));
    for (int c = 0; c < 20; ++c)
      for (int k = 0; k < 2; ++k) ids.push_back(reg(d.f_w[c][k], ntot, rb));
    for (int c = 0; c < 20; ++c)
      for (int k = 0; k < 2; ++k) ids.push_back(reg(d.f_cond[c][k], ntot, rb));
    for (int c = 0; c < 20; ++c)
      for (int k = 0; k < 5; ++k) ids.push_back(reg(d.chi1inv[c][k], ntot, rb));
    for (int c = 0; c < 20; ++c) ids.push_back(reg(d.chi2[c], ntot, rb));
    for (int c = 0; c < 20; ++c) ids.push_back(reg(d.chi3[c], ntot, rb));
    for (int c = 0; c < 20; ++c)
      for (int k = 0; k < 5; ++k) ids.push_back(reg(d.conductivity[c][k], ntot, rb));
    for (int c = 0; c < 20; ++c)
      for (int k = 0; k < 5; ++k) ids.push_back(reg(d.condinv[c][k], ntot, rb));
    for (int k = 0; k < 6; ++k) ids.push_back(k < 5 ? reg(d.sig[k], d.sigsize[k], rb) : -1);
    for (int k = 0; k < 6; ++k) ids.push_back(k < 5 ? reg(d.kap[k], d.sigsize[k], rb) : -1);
    for (int k = 0; k < 6; ++k) ids.push_back(k < 5 ? reg(d.siginv[k], d.sigsize[k], rb) : -1);
    for (int c = 0; c < 20; ++c) // ABI 2: appended, older readers stop before it
      for (int k = 0; k < 2; ++k) ids.push_back(reg(d.f_bfast[c][k], ntot, rb));
    begin(T_CHUNK);
    put32(i);
    put(d.num, sizeof d.num);
    put(d.io, sizeof d.io);
    put(d.sigsize, sizeof d.sigsize);
    put32(d.rank);
    put(ids.data(), ids.size() * 8);
    end();
  }
  void on_susceptibility(int i, const meepcu_susceptibility_desc &d, size_t ntot) {
    std::vector<int64_t> ids;
    for (int c = 0; c < 20; ++c)
      for (int k = 0; k < 5; ++k) ids.push_back(reg(d.sigma[c][k], ntot, rb));
    for (int c = 0; c < 20; ++c)
      for (int k = 0; k < 2; ++k) ids.push_back(reg(d.P[c][k], ntot, rb));
    for (int c = 0; c < 20; ++c)
      for (int k = 0; k < 2; ++k) ids.push_back(reg(d.P_prev[c][k], ntot, rb));
    begin(T_SUSC);
    put32(i);
    put32(d.ft);
    put32(d.no_omega_0_denominator);
    putd(d.omega_0);
    putd(d.gamma);
    put(ids.data(), ids.size() * 8);
    if (d.noise_amp != 0) putd(d.noise_amp); // appended (noisy Lorentzian): older readers stop before it
    end();
  }
  void on_multilevel(int i, const meepcu_multilevel_desc &d, size_t ntot) {
    std::vector<int64_t> ids;
    ids.push_back(reg(d.Gamma, (size_t)d.L * d.L, rb));
    ids.push_back(reg(d.GammaInv, (size_t)d.L * d.L, rb));
    ids.push_back(reg(d.alpha, (size_t)d.L * d.T, rb));
    ids.push_back(reg(d.omega, d.T, rb));
    ids.push_back(reg(d.gamma, d.T, rb));
    ids.push_back(reg(d.sigmat, 5 * (size_t)d.T, rb));
    ids.push_back(reg(d.N, ntot * d.L, rb));
    for (int c = 0; c < 20; ++c)
      for (int k = 0; k < 5; ++k) ids.push_back(reg(d.sigma[c][k], ntot, rb));
    for (int c = 0; c < 20; ++c)
      for (int k = 0; k < 2; ++k) ids.push_back(reg(d.W_prev[c][k], ntot, rb));
    for (int c = 0; c < 20; ++c)
      for (int k = 0; k < 2; ++k)
        for (int t = 0; t < d.T; ++t) ids.push_back(d.P[c][k] ? reg(d.P[c][k][t], ntot, rb) : -1);
    for (int c = 0; c < 20; ++c)
      for (int k = 0; k < 2; ++k)
        for (int t = 0; t < d.T; ++t) ids.push_back(d.P_prev[c][k] ? reg(d.P_prev[c][k][t], ntot, rb) : -1);
    begin(T_MLATOM);
    put32(i);
    put32(d.ft);
    put32(d.L);
    put32(d.T);
    put(ids.data(), ids.size() * 8);
    end();
  }
  void on_noise(int nsteps, size_t n_per_step, const double *g) { // the values of the NEXT T_STEP record
    begin(T_NOISE);
    put32(nsteps);
    put64((int64_t)n_per_step);
    put(g, 8 * (size_t)nsteps * n_per_step);
    end();
  }
  void on_gyrotropic(int i, const meepcu_gyrotropic_desc &d, size_t ntot) {
    std::vector<int64_t> ids;
    for (int c = 0; c < 20; ++c)
      for (int k = 0; k < 5; ++k) ids.push_back(reg(d.sigma[c][k], ntot, rb));
    for (int c = 0; c < 20; ++c)
      for (int k = 0; k < 2; ++k)
        for (int dd = 0; dd < 3; ++dd) ids.push_back(reg(d.P[c][k][dd], ntot, rb));
    for (int c = 0; c < 20; ++c)
      for (int k = 0; k < 2; ++k)
        for (int dd = 0; dd < 3; ++dd) ids.push_back(reg(d.P_prev[c][k][dd], ntot, rb));
    begin(T_GYRO);
    put32(i);
    put32(d.ft);
    put32(d.model);
    putd(d.omega_0);
    putd(d.gamma);
    putd(d.alpha);
    put(d.gyro_tensor, sizeof d.gyro_tensor);
    put(ids.data(), ids.size() * 8);
    end();
  }
  void on_source(int i, const meepcu_source_desc &d) {
    begin(T_SOURCE);
    put32(i);
    put32(d.ft);
    put32(d.c);
    put32(d.src_time_id);
    put32(d.is_integrated);
    put64(d.n);
    for (size_t j = 0; j < d.n; ++j) put64((int64_t)d.index[j]);
    put(d.amp, d.n * 16);
    end();
  }
  void on_dft(int i, const meepcu_dft_desc &d, size_t N) {
    int64_t id = reg(d.dft, N * d.Nomega, 2 * rb);
    begin(T_DFT);
    put32(i);
    put32(d.c);
    put(d.is, sizeof d.is);
    put(d.ie, sizeof d.ie);
    put32(d.include_dV_and_interp_weights);
    put32(d.sqrt_dV_and_interp_weights);
    put32(d.decimation_factor);
    put32(d.Nomega);
    put(d.s0, 24);
    put(d.s1, 24);
    put(d.e0, 24);
    put(d.e1, 24);
    putd(d.dV0);
    putd(d.dV1);
    put(d.scale, 16);
    put64(d.avg1);
    put64(d.avg2);
    put(d.omega, 8 * (size_t)d.Nomega);
    put64(id);
    put64(N);
    end();
  }
  void on_metal(int ft, size_t n, void *const *p) {
    begin(T_METAL);
    put32(ft);
    put64(n);
    for (size_t k = 0; k < n; ++k) putptr(p[k]);
    end();
  }
  void on_connections(int ft, int kind, size_t n, void *const *dst, const void *const *src,
                      const void *ph) {
    begin(T_CONN);
    put32(ft);
    put32(kind);
    put64(n);
    for (size_t k = 0; k < n; ++k) putptr(dst[k]);
    for (size_t k = 0; k < n; ++k) putptr(src[k]);
    if (kind == MEEPCU_CONNECT_PHASE) put(ph, (n / 2) * 2 * rb);
    end();
  }
  void on_finalize() {
    flush_arrays();
    begin(T_FINALIZE);
    end();
  }
  void on_step(int n, int nsrc, const meepcu_src_sample *s) {
    begin(T_STEP);
    put32(n);
    put32(nsrc);
    put(s, sizeof(meepcu_src_sample) * (size_t)n * nsrc);
    end();
  }

  // after the run: `mine` is an array of the recorded fields object, `expect` the same array of
  // the reference-stepped twin
  void expect(const void *mine, const void *expected) {
    std::map<const void *, entry>::iterator it = arrays.find(mine);
    if (it == arrays.end()) meep::abort("trace: expect() on an unregistered array");
    begin(T_EXPECT);
    put32((int)it->second.id);
    put64(it->second.n);
    put(expected, it->second.n * it->second.eb);
    end();
  }

private:
  struct entry {
    int64_t id;
    size_t n, eb;
    bool written;
  };
  FILE *fp;
  int rb;
  std::map<const void *, entry> arrays; // by base address
  std::vector<char> buf;
  int cur_tag;

  int64_t reg(const void *p, size_t n, size_t eb) {
    if (!p) return -1;
    std::map<const void *, entry>::iterator it = arrays.find(p);
    if (it != arrays.end()) return it->second.id;
    entry e = {(int64_t)arrays.size(), n, eb, false};
    arrays[p] = e;
    return e.id;
  }
  void flush_arrays() {
    for (std::map<const void *, entry>::iterator it = arrays.begin(); it != arrays.end(); ++it) {
      if (it->second.written) continue;
      begin(T_ARRAY);
      put32((int)it->second.id);
      put32((int)it->second.eb);
      put64(it->second.n);
      put(it->first, it->second.n * it->second.eb);
      end();
      it->second.written = true;
    }
  }
  void putptr(const void *p) { // (array id, element offset)
    std::map<const void *, entry>::iterator it = arrays.upper_bound(p);
    if (it == arrays.begin()) meep::abort("trace: pointer outside every registered array");
    --it;
    const char *base = (const char *)it->first;
    size_t off = (const char *)p - base;
    if (off >= it->second.n * it->second.eb)
      meep::abort("trace: pointer outside every registered array");
    put64(it->second.id);
    put64(off / it->second.eb);
  }
  void begin(int tag) {
    cur_tag = tag;
    buf.clear();
  }
  void put(const void *p, size_t n) {
    const char *c = (const char *)p;
    buf.insert(buf.end(), c, c + n);
  }
  void put32(int v) {
    int32_t x = v;
    put(&x, 4);
  }
  void put64(int64_t v) { put(&v, 8); }
  void putd(double v) { put(&v, 8); }
  void end() {
    uint32_t tag = cur_tag;
    uint64_t len = buf.size();
    fwrite(&tag, 4, 1, fp);
    fwrite(&len, 8, 1, fp);
    if (len) fwrite(buf.data(), 1, len, fp);
  }
};

} // namespace meepcu
#endif
