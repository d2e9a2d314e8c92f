/* TEST INFRASTRUCTURE.  LD_PRELOAD interposer: logs the loop sets the plugin asks the library to sweep
 * (spk_cuda_sweep_color: colour and sector box in global cells) to the file SPK_TRACE_SETS, then calls the
 * real entry point.  tests/test_emulated_library.py compares them with the reference's own sets
 * (AppLattice::create_set, src/app_lattice.cpp:954-1026) on every rank. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

struct spk_cuda_handle;
typedef int (*sweep_color_fn)(struct spk_cuda_handle *, uint32_t, int, const int32_t *, int);

int spk_cuda_sweep_color(struct spk_cuda_handle *h, uint32_t sweep, int color, const int32_t *box, int nloop) {
  static sweep_color_fn real;
  if (!real) real = (sweep_color_fn)dlsym(RTLD_NEXT, "spk_cuda_sweep_color");
  const char *path = getenv("SPK_TRACE_SETS");
  if (path) {
    FILE *fp = fopen(path, "a");
    if (fp) {
      if (box) fprintf(fp, "%u %d %d %d %d %d %d %d %d\n", sweep, color, nloop, box[0], box[1], box[2], box[3], box[4], box[5]);
      else fprintf(fp, "%u %d %d -1 -1 -1 -1 -1 -1\n", sweep, color, nloop);
      fclose(fp);
    }
  }
  return real(h, sweep, color, box, nloop);
}
