/* A plain C99 client of include/meepcu.h: proves that the header is valid C (no C++ types cross the boundary)
 * and that the shared library given on the command line exports every declared entry point with C linkage.
 * Compiled and run by tests/test_abi.py; makes no compute call. */
#include "meepcu.h"
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#define SYM(name)                                                                                  \
  do {                                                                                             \
    if (!dlsym(h, #name)) {                                                                        \
      fprintf(stderr, "missing symbol %s\n", #name);                                               \
      ++missing;                                                                                   \
    }                                                                                              \
    ++total;                                                                                       \
  } while (0)

int main(int argc, char **argv) {
  void *h;
  int missing = 0, total = 0;
  meepcu_config cfg;
  meepcu_chunk_desc cd;
  meepcu_susceptibility_desc sd;
  meepcu_gyrotropic_desc gd;
  meepcu_multilevel_desc md;
  meepcu_source_desc src;
  meepcu_dft_desc dd;
  meepcu_src_sample smp;
  meepcu_stats st;
  meepcu_kernel_time kt;
  meepcu_integrate_desc idsc;
  int (*abi_version)(void);
  if (argc < 2) return 2;
  memset(&cfg, 0, sizeof cfg); memset(&cd, 0, sizeof cd); memset(&sd, 0, sizeof sd); memset(&gd, 0, sizeof gd);
  memset(&src, 0, sizeof src); memset(&dd, 0, sizeof dd); memset(&smp, 0, sizeof smp); memset(&st, 0, sizeof st);
  memset(&kt, 0, sizeof kt); memset(&idsc, 0, sizeof idsc); memset(&md, 0, sizeof md);
  h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!h) { fprintf(stderr, "%s\n", dlerror()); return 3; }
  SYM(meepcu_last_error); SYM(meepcu_backend); SYM(meepcu_abi_version); SYM(meepcu_create); SYM(meepcu_destroy);
  SYM(meepcu_add_chunk); SYM(meepcu_add_susceptibility); SYM(meepcu_add_gyrotropic); SYM(meepcu_add_source);
  SYM(meepcu_add_dft); SYM(meepcu_set_metal); SYM(meepcu_add_connections); SYM(meepcu_add_remote_send);
  SYM(meepcu_add_remote_recv); SYM(meepcu_comm_unique_id); SYM(meepcu_comm_init); SYM(meepcu_finalize);
  SYM(meepcu_step); SYM(meepcu_synchronize_magnetic_fields); SYM(meepcu_restore_magnetic_fields);
  SYM(meepcu_sync_to_host); SYM(meepcu_sync_from_host); SYM(meepcu_sync_subset); SYM(meepcu_get_time);
  SYM(meepcu_set_time); SYM(meepcu_check_finite); SYM(meepcu_read_value); SYM(meepcu_read_values);
  SYM(meepcu_get_timers); SYM(meepcu_get_stats); SYM(meepcu_set_profiling); SYM(meepcu_get_kernel_times);
  SYM(meepcu_dft_dot); SYM(meepcu_integrate_dot); SYM(meepcu_get_compulsory_bytes);
  SYM(meepcu_add_multilevel); SYM(meepcu_get_noise_count); SYM(meepcu_set_noise);
  *(void **)(&abi_version) = dlsym(h, "meepcu_abi_version");
  printf("{\"symbols\": %d, \"missing\": %d, \"abi\": %d, \"header_abi\": %d, \"sizeof_config\": %d, "
         "\"sizeof_chunk_desc\": %d, \"sizeof_gyrotropic_desc\": %d, \"sizeof_susceptibility_desc\": %d, "
         "\"sizeof_multilevel_desc\": %d}\n",
         total, missing, abi_version ? abi_version() : -1, MEEPCU_ABI_VERSION, (int)sizeof cfg, (int)sizeof cd,
         (int)sizeof gd, (int)sizeof sd, (int)sizeof md);
  return missing ? 1 : 0;
}
