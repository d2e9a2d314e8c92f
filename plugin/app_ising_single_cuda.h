/* ----------------------------------------------------------------------
   app_style ising/single/cuda : the reference's ising/single app with its rejection-KMC sweep
   (AppIsingSingle::site_event_rejection) executed on a B200 by libspparks_cuda.
   See app_cuda_mixin.h.  Registered through style_app.h like any other app
   (src/Make.sh:8-45, src/input.cpp:739-743).
------------------------------------------------------------------------- */

#ifdef APP_CLASS
AppStyle(ising/single/cuda,AppIsingSingleCuda)

#else

#ifndef SPK_APP_ISING_SINGLE_CUDA_H
#define SPK_APP_ISING_SINGLE_CUDA_H

#include "app_ising_single.h"
#include "app_cuda_mixin.h"

namespace SPPARKS_NS {

typedef AppCudaMixin<AppIsingSingle,SPK_ISING_SINGLE> AppIsingSingleCuda;

}

#endif
#endif
