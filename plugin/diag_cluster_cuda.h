/* ----------------------------------------------------------------------
   diag_style cluster/cuda : the three stats columns of DiagCluster (Nclust, <N>, <R>;
   src/diag_cluster.cpp:585-612, 894-904) with the connected-component analysis done on the
   device (spk_cuda_cluster_stats) instead of the host's stack flood fill
   (src/diag_cluster.cpp:286-354).  Serial runs (one GPU) only; the per-cluster file and dump
   options of the host diagnostic are not offered.
------------------------------------------------------------------------- */

#ifdef DIAG_CLASS
DiagStyle(cluster/cuda,DiagClusterCuda)

#else

#ifndef SPK_DIAG_CLUSTER_CUDA_H
#define SPK_DIAG_CLUSTER_CUDA_H

#include "stdio.h"
#include "stdint.h"
#include "diag.h"
#include "app.h"
#include "error.h"
#include "spk_cuda_host.h"

namespace SPPARKS_NS {

class DiagClusterCuda : public Diag {
 public:
  DiagClusterCuda(class SPPARKS *spk, int narg, char **arg) : Diag(spk,narg,arg)
  {
    if (iarg_child < narg) error->all(FLERR,"Illegal diag_style cluster/cuda command");
    host = NULL;
    ncluster = 0;
    vav = rav = 0.0;
  }
  ~DiagClusterCuda() {}

  void init()
  {
    host = dynamic_cast<SpkCudaHost *>(app);
    if (!host) error->all(FLERR,"diag_style cluster/cuda requires an app_style <name>/cuda");
    // several ranks: the library merges the per-slab clusters across the slab boundaries and
    // hands every rank the statistics of the whole lattice (compute() is collective, as
    // DiagCluster::compute is, src/diag_cluster.cpp:230-260)
  }

  void compute()
  {
    int64_t n;
    if (spk_cuda_cluster_stats(host->cuda_lattice(),&n,&vav,&rav)) error->one(FLERR,spk_cuda_last_error());
    ncluster = n;
  }

  void stats(char *strtmp) { sprintf(strtmp," %10d %10g %10g",(int) ncluster,vav,rav); }
  void stats_header(char *strtmp) { sprintf(strtmp," %10s %10s %10s","Nclust","<N>","<R>"); }

 private:
  SpkCudaHost *host;
  int64_t ncluster;
  double vav,rav;
};

}

#endif
#endif
