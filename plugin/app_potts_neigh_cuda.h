/* ----------------------------------------------------------------------
   app_style potts/neigh/cuda : the reference's potts/neigh app with its rejection-KMC sweep
   (AppPottsNeigh::site_event_rejection) executed on a B200 by libspparks_cuda.
   See app_cuda_mixin.h.  Registered through style_app.h like any other app
   (src/Make.sh:8-45, src/input.cpp:739-743).
------------------------------------------------------------------------- */

#ifdef APP_CLASS
AppStyle(potts/neigh/cuda,AppPottsNeighCuda)

#else

#ifndef SPK_APP_POTTS_NEIGH_CUDA_H
#define SPK_APP_POTTS_NEIGH_CUDA_H

#include "app_potts_neigh.h"
#include "app_cuda_mixin.h"

namespace SPPARKS_NS {

typedef AppCudaMixin<AppPottsNeigh,SPK_POTTS_NEIGH> AppPottsNeighCuda;

}

#endif
#endif
