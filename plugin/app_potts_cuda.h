/* ----------------------------------------------------------------------
   app_style potts/cuda : the reference's potts app with its rejection-KMC sweep
   (AppPotts::site_event_rejection) executed on a B200 by libspparks_cuda.
   See app_cuda_mixin.h.  Registered through style_app.h like any other app
   (src/Make.sh:8-45, src/input.cpp:739-743).
------------------------------------------------------------------------- */

#ifdef APP_CLASS
AppStyle(potts/cuda,AppPottsCuda)

#else

#ifndef SPK_APP_POTTS_CUDA_H
#define SPK_APP_POTTS_CUDA_H

#include "app_potts.h"
#include "app_cuda_mixin.h"

namespace SPPARKS_NS {

typedef AppCudaMixin<AppPotts,SPK_POTTS> AppPottsCuda;

}

#endif
#endif
