/* ----------------------------------------------------------------------
   What the /cuda diagnostics (diag_energy_cuda.h, diag_cluster_cuda.h) need from an
   app_style <name>/cuda: the device lattice, holding the current spins.
------------------------------------------------------------------------- */

#ifndef SPK_CUDA_HOST_H
#define SPK_CUDA_HOST_H

#include "spparks_cuda.h"

namespace SPPARKS_NS {

class SpkCudaHost {
 public:
  virtual ~SpkCudaHost() {}

  // handle of libspparks_cuda whose device lattice equals the host copy iarray[0]
  // (uploads the host copy first if a script command may have changed it since the last sweep)
  virtual spk_cuda_handle *cuda_lattice() = 0;
};

}

#endif
