#!/usr/bin/env python3
"""The patch a maintainer adds to the reference so that `fields::step()` itself runs on the device
(INTEGRATION.md section 2), expressed as anchored insertions -- no reference text is stored here.

    dropin_patch.py <basename> < original > patched
    (basename: meep.hpp | step.cpp | fields.cpp | energy_and_flux.cpp | loop_in_chunks.cpp | monitor.cpp)

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
  if (gpu) ((meepcu::attachment *)gpu)->sync_to_host_if_stale();
}
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
  if (!g_dropin_api || f->is_phasing()) return false;
  // a source / DFT monitor / polarisation added or arrays re-allocated since the last step: the host arrays
  // become the truth again (sync back), and the object graph is described anew
  if (f->gpu && ((meepcu::attachment *)f->gpu)->description_outdated()) f->gpu_detach();
  if (!f->gpu) f->gpu = new meepcu::attachment(*f, *g_dropin_api, g_dropin_nl_mode, g_dropin_device);
  meepcu::attachment *att = (meepcu::attachment *)f->gpu;
  const int saved = f->synchronized_magnetic_fields;
  if (saved) att->restore_magnetic_fields();
  f->am_now_working_on(Stepping);
  att->step(1); // advances fields::t
  f->finished_working();
  if (saved) att->synchronize_magnetic_fields();
  if (!att->finite()) meep::abort("simulation fields are NaN or Inf");
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
        return text
    if name == "loop_in_chunks.cpp":  # the iterator behind integrate / get_array_slice / output_hdf5 / energy / flux_in_box
        a = text.index("\nvoid fields::loop_in_chunks(")
        k = text.index("{", a) + 1
        return text[:k] + "\n  gpu_sync_to_host(); // meepcu drop-in: readers see the device's state\n" + text[k:]
    if name == "monitor.cpp":  # point probes: every get_field overload ends in the (component, ivec) one
        a = text.index("fields::get_field(component c, const ivec &")
        k = text.index("{", a) + 1
        return text[:k] + "\n  const_cast<fields *>(this)->gpu_sync_to_host(); // meepcu drop-in\n" + text[k:]
    if name == "fields.cpp":
        a = text.index("\nfields::~fields() {")
        k = text.index("{", a) + 1
        return text[:k] + "\n  gpu_detach(); // meepcu drop-in\n" + text[k:]
    return text


if __name__ == "__main__":
    sys.stdout.write(patch(sys.argv[1], sys.stdin.read()))
