/* ----------------------------------------------------------------------
   diag_style energy/cuda : DiagEnergy (src/diag_energy.cpp:53-66) for the /cuda apps with the
   sum of site_energy() taken on the device (spk_cuda_energy) instead of a host loop over
   nlocal sites with maxneigh gathers each.  Same stats column ("Energy", %10g).
------------------------------------------------------------------------- */

#ifdef DIAG_CLASS
DiagStyle(energy/cuda,DiagEnergyCuda)

#else

#ifndef SPK_DIAG_ENERGY_CUDA_H
#define SPK_DIAG_ENERGY_CUDA_H

#include "mpi.h"
#include "stdio.h"
#include "diag.h"
#include "app.h"
#include "error.h"
#include "spk_cuda_host.h"

namespace SPPARKS_NS {

class DiagEnergyCuda : public Diag {
 public:
  DiagEnergyCuda(class SPPARKS *spk, int narg, char **arg) : Diag(spk,narg,arg)
  {
    host = NULL;
    energy = 0.0;
  }
  ~DiagEnergyCuda() {}

  void init()
  {
    host = dynamic_cast<SpkCudaHost *>(app);
    if (!host) error->all(FLERR,"diag_style energy/cuda requires an app_style <name>/cuda");
    energy = 0.0;
  }

  void compute()
  {
    double etmp = 0.0;
    if (spk_cuda_energy(host->cuda_lattice(),&etmp)) error->one(FLERR,spk_cuda_last_error());
    MPI_Allreduce(&etmp,&energy,1,MPI_DOUBLE,MPI_SUM,world);
  }

  void stats(char *strtmp) { sprintf(strtmp," %10g",energy); }
  void stats_header(char *strtmp) { sprintf(strtmp," %10s","Energy"); }

 private:
  SpkCudaHost *host;
  double energy;
};

}

#endif
#endif
