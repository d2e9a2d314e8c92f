/* ----------------------------------------------------------------------
   app_style ising/cuda : the reference's ising app with its rejection-KMC sweep
   (AppIsing::site_event_rejection) executed on a B200 by libspparks_cuda.
   See app_cuda_mixin.h.  Registered through style_app.h like any other app
   (src/Make.sh:8-45, src/input.cpp:739-743).
------------------------------------------------------------------------- */

#ifdef APP_CLASS
AppStyle(ising/cuda,AppIsingCuda)

#else

#ifndef SPK_APP_ISING_CUDA_H
#define SPK_APP_ISING_CUDA_H

#include "app_ising.h"
#include "app_cuda_mixin.h"

namespace SPPARKS_NS {

typedef AppCudaMixin<AppIsing,SPK_ISING> AppIsingCuda;

}

#endif
#endif
