#!/usr/bin/env python3
"""The patch a maintainer adds to the reference so that `fields::step()` itself runs on the device
(INTEGRATION.md section 2), expressed as anchored insertions -- no reference text is stored here.

    dropin_patch.py <basename> < original > patched
    (basename: meep.hpp | step.cpp | fields.cpp | energy_and_flux.cpp | loop_in_chunks.cpp | monitor.cpp |
     initialize.cpp | fields_dump.cpp | boundaries.cpp | dft.cpp | near2far.cpp)

oracle/build_ref.sh applies it in the compiler's input pipe for the `*_dropin` flavours; the result is
the reference library with ONE added data member (`fields::gpu`), three added methods and a four-line
prologue in `fields::step()`.  Everything else of the reference is untouched; with the drop-in disabled
(the default) the patched library behaves exactly like the original.
"""
import sys

MEMBERS = '''
  // ---- meepcu drop-in (meep_nl_b200/shim/dropin_patch.py) ----
  void *gpu = 0;           // meepcu::attachment*: non-null while the fields live on the device
  void gpu_sync_to_host(); // before a host-side reader of the field arrays (no-op unless the device has stepped since)
  void gpu_detach();       // sync back and drop the device mirror (writers of the arrays call this)
  void gpu_synchronize_magnetic_fields(); // bodies of synchronize / restore_magnetic_fields on the device
  void gpu_restore_magnetic_fields();
  // process-wide switch: the library the drop-in dlopen()s (NULL: off), how it treats chi2/chi3, the device
  static void gpu_enable(const char *library_path, int nl_mode = 0, int device = 0);
  // DFT accumulators (dft_chunk::dft) live on the device while stepping: readers (dft_flux::flux, dft_energy,
  // near2far, save_dft_hdf5, process_dft_component) fetch them first -- only them, not the fields; writers
  // (scale_dft, operator-=, load_dft_hdf5) fetch them and have them sent back before the next step
  static void gpu_dft_read();
  static void gpu_dft_write();
  // point probes (every fields::get_field overload ends in the (component, ivec) one) read their few values
  // straight from the device mirror instead of downloading the whole state; false: the host arrays are current
  bool gpu_get_field(component c, const ivec &origloc, bool parallel, std::complex<double> *out) const;
  // reductions on the device (no array travels): fields::integrate(2, cs, dot_integrand, 0, where) -- what
  // field_energy_in_box / electric_energy_in_box / magnetic_energy_in_box / flux_in_box_wrongH compute -- and the
  // sums of dft_flux::flux / dft_energy::electric, magnetic.  false: not on the device, the host code runs as before
  bool gpu_integrate_dot(const component *cs, const volume &where, double *out);
  static bool gpu_dft_dot(dft_chunk *a, dft_chunk *b, size_t Nfreq, double *F);
  bool gpu_collecting = false; // loop_in_chunks is being used to enumerate pieces only
#define MEEPCU_DROPIN_PATCHED 1
'''

STEP_HELPERS = r'''
// ---- meepcu drop-in: fields::step() delegates to the device when a library is enabled -----------
namespace {
meepcu::api *g_dropin_api = 0;
int g_dropin_nl_mode = 0, g_dropin_device = 0;
} // namespace
namespace meep {
void fields::gpu_enable(const char *library_path, int nl_mode, int device) {
  // loaded libraries stay loaded: attachments made earlier keep using theirs after a switch
  static std::map<std::string, meepcu::api *> loaded;
  g_dropin_api = 0;
  if (library_path) {
    meepcu::api *&a = loaded[library_path];
    if (!a) a = new meepcu::api(library_path);
    g_dropin_api = a;
  }
  g_dropin_nl_mode = nl_mode;
  g_dropin_device = device;
}
void fields::gpu_sync_to_host() {
  if (!gpu) return;
  meepcu::attachment *att = (meepcu::attachment *)gpu;
  if (!att->stale()) return;
  am_now_working_on(Stepping); // the steps queued by step() finish here: this wait IS stepping time
  att->wait();
  finished_working();
  att->sync_to_host_if_stale();
}
bool fields::gpu_get_field(component c, const ivec &origloc, bool parallel, std::complex<double> *out) const {
  meepcu::attachment *att = (meepcu::attachment *)gpu;
  if (!att || !att->stale()) return false;
  ivec iloc = origloc; // same walk as the host version (monitor.cpp:145-157), the leaf read is the device's
  std::complex<double> kphase = 1.0;
  locate_point_in_user_volume(&iloc, &kphase);
  for (int sn = 0; sn < S.multiplicity(); sn++)
    for (int i = 0; i < num_chunks; i++)
      if (chunks[i]->gv.owns(S.transform(iloc, sn))) {
        const fields_chunk *fc = chunks[i];
        const component tc = S.transform(c, sn);
        std::complex<double> raw = 0.0;
        if (fc->is_mine() && fc->f[tc][0]) {
          const ptrdiff_t idx = fc->gv.index(tc, S.transform(iloc, sn));
          raw = std::complex<double>(att->read(fc->f[tc][0] + idx), fc->f[tc][1] ? att->read(fc->f[tc][1] + idx) : 0.0);
        }
        const std::complex<double> val = S.phase_shift(c, sn) * kphase * raw;
        *out = parallel ? sum_to_all(val) : val;
        return true;
      }
  *out = 0.0;
  return true;
}
bool fields::gpu_integrate_dot(const component *cs, const volume &where, double *out) {
  meepcu::attachment *att = (meepcu::attachment *)gpu;
  if (!att || !att->stale()) return false;
  return att->integrate_dot(cs, where, out);
}
bool fields::gpu_dft_dot(dft_chunk *a, dft_chunk *b, size_t Nfreq, double *F) {
  return meepcu::attachment::dft_dot_on_device(a, b, Nfreq, F);
}
void fields::gpu_dft_read() { meepcu::attachment::for_each_live(&meepcu::attachment::sync_dft_to_host_if_stale); }
void fields::gpu_dft_write() { meepcu::attachment::for_each_live(&meepcu::attachment::dft_will_be_written); }
void fields::gpu_synchronize_magnetic_fields() { ((meepcu::attachment *)gpu)->synchronize_magnetic_fields(); }
void fields::gpu_restore_magnetic_fields() { ((meepcu::attachment *)gpu)->restore_magnetic_fields(); }
void fields::gpu_detach() {
  if (!gpu) return;
  meepcu::attachment *att = (meepcu::attachment *)gpu;
  if (synchronized_magnetic_fields) att->restore_magnetic_fields();
  att->sync_to_host();
  delete att;
  gpu = 0;
}
// true: the step was done on the device.  Mirrors the reference's own prologue / epilogue around the
// stepping phases (restore, step, re-synchronize: src/step.cpp:36-41,128-134; NaN test :138-139).
static bool dropin_step(fields *f) {
  // opt-in without touching the caller: MEEP_CUDA_LIB=/path/to/libmeepcuda.so [MEEP_CUDA_NL_MODE=pade] [MEEP_CUDA_DEVICE=n]
  // in the environment of an unchanged script (python/simulation.py:2831-2834 calls fields::step()); read once, and
  // only while nobody has called fields::gpu_enable()
  static bool env_checked = false;
  if (!g_dropin_api && !env_checked) {
    env_checked = true;
    const char *lib = getenv("MEEP_CUDA_LIB");
    if (lib && *lib) {
      const char *nl = getenv("MEEP_CUDA_NL_MODE"), *dev = getenv("MEEP_CUDA_DEVICE");
      fields::gpu_enable(lib, nl && !strcmp(nl, "pade") ? 1 : 0, dev ? atoi(dev) : 0);
    }
  }
  if (!g_dropin_api || f->is_phasing()) return false;
  // a source / DFT monitor / polarisation added or arrays re-allocated since the last step: the host arrays
  // become the truth again (sync back), and the object graph is described anew
  if (f->gpu && ((meepcu::attachment *)f->gpu)->description_outdated()) f->gpu_detach();
  if (!f->gpu) f->gpu = new meepcu::attachment(*f, *g_dropin_api, g_dropin_nl_mode, g_dropin_device);
  meepcu::attachment *att = (meepcu::attachment *)f->gpu;
  const int saved = f->synchronized_magnetic_fields;
  if (saved) att->restore_magnetic_fields();
  f->am_now_working_on(Stepping);
  // queues one time step and returns (advances fields::t).  The NaN test of step.cpp:138-139 is part of the
  // call: a sample of every field array is tested on the device after each step and the next call aborts with
  // the reference's message when it failed -- no per-step wait, no per-step scan of whole arrays
  att->step(1);
  f->finished_working();
  if (saved) att->synchronize_magnetic_fields();
  return true;
}
} // namespace meep
'''


def patch(name, text):
    if name == "meep.hpp":
        a = text.index("\nclass fields {")
        b = text.index("public:", a) + len("public:")
        return text[:b] + MEMBERS + text[b:]
    if name == "step.cpp":
        text = '#include "meepcu_shim.hpp" // first: it includes meep.hpp with the private members visible\n' + text
        # meep_internals.hpp has no include guard and the shim has just included it
        text = text.replace('#include "meep_internals.hpp"', '/* meep_internals.hpp: included by meepcu_shim.hpp */', 1)
        a = text.index("\nvoid fields::step() {")
        # helpers go before the enclosing `namespace meep {`-scoped definition: close and reopen the namespace
        head = text[:a] + "\n} // namespace meep\n" + STEP_HELPERS + "namespace meep {\n"
        body = text[a:]
        k = body.index("{") + 1
        return head + body[:k] + "\n  if (dropin_step(this)) return; // meepcu drop-in\n" + body[k:]
    if name == "energy_and_flux.cpp":
        # after the reference's own reference-count test (the first `return;` of each function) the body runs on
        # the device when the fields live there
        for fn, call in (("synchronize_magnetic_fields", "gpu_synchronize_magnetic_fields"),
                         ("restore_magnetic_fields", "gpu_restore_magnetic_fields")):
            a = text.index("\nvoid fields::%s() {" % fn)
            k = text.index("return;", a) + len("return;")
            text = text[:k] + "\n  if (gpu) { %s(); return; } // meepcu drop-in" % call + text[k:]
        # field_energy_in_box(c, where) and flux_in_box_wrongH: the integral is formed on the device
        a = text.index("return real(integrate(2, cs, dot_integrand, 0, where)) * 0.5;")
        text = text[:a] + "{ double v_; if (gpu_integrate_dot(cs, where, &v_)) return v_ * 0.5; } // meepcu drop-in\n  " + text[a:]
        a = text.index("sum += real(integrate(2, cs, dot_integrand, 0, where)) * (1 - 2 * i);")
        text = text[:a] + ("{ double v_; if (gpu_integrate_dot(cs, where, &v_)) { sum += v_ * (1 - 2 * i); continue; } } "
                           "// meepcu drop-in\n    ") + text[a:]
        return text
    if name == "loop_in_chunks.cpp":  # the iterator behind integrate / get_array_slice / output_hdf5 / energy / flux_in_box
        a = text.index("\nvoid fields::loop_in_chunks(")
        k = text.index("{", a) + 1
        return text[:k] + "\n  if (!gpu_collecting) gpu_sync_to_host(); // meepcu drop-in: readers see the device's state\n" + text[k:]
    if name == "monitor.cpp":  # point probes: every get_field overload ends in the (component, ivec) one
        a = text.index("fields::get_field(component c, const ivec &")
        k = text.index("{", a) + 1
        return text[:k] + ("\n  { std::complex<double> v_; if (gpu && gpu_get_field(c, origloc, parallel, &v_)) return v_; } "
                           "// meepcu drop-in: a point probe reads the device\n") + text[k:]
    if name == "fields.cpp":
        a = text.index("\nfields::~fields() {")
        k = text.index("{", a) + 1
        text = text[:k] + "\n  gpu_detach(); // meepcu drop-in\n" + text[k:]
    # host-side WRITERS of the field arrays / of the set of arrays: the device mirror is synced back and dropped,
    # the next step() describes the object graph anew (the rule the reference applies to chunk_connections_valid,
    # boundaries.cpp:171-183)
    for fname, sigs in WRITERS.items():
        if name == fname:
            for sig in sigs:
                text = at_function_start(text, sig, "gpu_detach(); // meepcu drop-in: host-side writer")
    for fname, sigs in DFT_READERS.items():
        if name == fname:
            for sig in sigs:
                text = at_function_start(text, sig, "fields::gpu_dft_read(); // meepcu drop-in: accumulators live on the device")
    for fname, sigs in DFT_WRITERS.items():
        if name == fname:
            for sig in sigs:
                text = at_function_start(text, sig, "fields::gpu_dft_write(); // meepcu drop-in")
    # (inserted last, so that it runs FIRST: a spectrum that can be summed on the device fetches nothing)
    if name == "dft.cpp":
        for sig, la, lb, fac in (("double *dft_flux::flux()", "E", "H", ""), ("double *dft_energy::electric()", "E", "D", "0.5 * "),
                                 ("double *dft_energy::magnetic()", "H", "B", "0.5 * ")):
            text = at_function_start(text, sig,
                "{ const size_t Nf_ = freq.size(); double *F_ = new double[Nf_]; // meepcu drop-in: summed on the device\n"
                "    if (fields::gpu_dft_dot(%s, %s, Nf_, F_)) { double *Fs_ = new double[Nf_]; for (size_t i_ = 0; i_ < Nf_; ++i_) F_[i_] = %sF_[i_];\n"
                "      sum_to_all(F_, Fs_, int(Nf_)); delete[] F_; return Fs_; }\n"
                "    delete[] F_; }" % (la, lb, fac))
    return text


def at_function_start(text, sig, line):
    a = text.index("\n" + sig)
    k = text.index("{", text.index(")", a)) + 1
    return text[:k] + "\n  " + line + "\n" + text[k:]


WRITERS = {
    "fields.cpp": ["void fields::use_real_fields()", "void fields::_require_component(", "void fields::remove_sources()",
                   "void fields::remove_susceptibilities()", "void fields::remove_fluxes()", "void fields::zero_fields()",
                   "void fields::reset()"],
    "initialize.cpp": ["void fields::initialize_field("],
    "fields_dump.cpp": ["void fields::load(const char *filename"],
    "boundaries.cpp": ["void fields::set_boundary(", "void fields::use_bloch(direction d"],
}
DFT_READERS = {
    "dft.cpp": ["void save_dft_hdf5(dft_chunk *dft_chunks, const char *name", "double *dft_flux::flux()",
                "double *dft_energy::electric()", "double *dft_energy::magnetic()",
                "complex<double> fields::process_dft_component("],
    "near2far.cpp": ["void dft_near2far::farfield_lowlevel("],
}
DFT_WRITERS = {
    "dft.cpp": ["void dft_chunk::scale_dft(", "void dft_chunk::operator-=(", "void load_dft_hdf5(dft_chunk *dft_chunks, const char *name"],
}


if __name__ == "__main__":
    sys.stdout.write(patch(sys.argv[1], sys.stdin.read()))
