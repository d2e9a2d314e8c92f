/* ----------------------------------------------------------------------
   app_style potts/neighonly/cuda : the reference's potts/neighonly app with its rejection-KMC sweep
   (AppPottsNeighOnly::site_event_rejection) executed on a B200 by libspparks_cuda.
   See app_cuda_mixin.h.  Registered through style_app.h like any other app
   (src/Make.sh:8-45, src/input.cpp:739-743).
------------------------------------------------------------------------- */

#ifdef APP_CLASS
AppStyle(potts/neighonly/cuda,AppPottsNeighOnlyCuda)

#else

#ifndef SPK_APP_POTTS_NEIGHONLY_CUDA_H
#define SPK_APP_POTTS_NEIGHONLY_CUDA_H

#include "app_potts_neighonly.h"
#include "app_cuda_mixin.h"

namespace SPPARKS_NS {

typedef AppCudaMixin<AppPottsNeighOnly,SPK_POTTS_NEIGHONLY> AppPottsNeighOnlyCuda;

}

#endif
#endif
