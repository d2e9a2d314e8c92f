/* ----------------------------------------------------------------------
   app_style potts/pin/cuda : the reference's potts/pin app with its rejection-KMC sweep
   (AppPottsPin::site_event_rejection) executed on a B200 by libspparks_cuda.
   See app_cuda_mixin.h.  Registered through style_app.h like any other app
   (src/Make.sh:8-45, src/input.cpp:739-743).
------------------------------------------------------------------------- */

#ifdef APP_CLASS
AppStyle(potts/pin/cuda,AppPottsPinCuda)

#else

#ifndef SPK_APP_POTTS_PIN_CUDA_H
#define SPK_APP_POTTS_PIN_CUDA_H

// app_potts_pin.h has no include guard in the reference; style_app.h lists it before this
// header (alphabetical order of src/Make.sh), so AppPottsPin is already declared here.
#include "app_cuda_mixin.h"

namespace SPPARKS_NS {

typedef AppCudaMixin<AppPottsPin,SPK_POTTS_PIN> AppPottsPinCuda;

}

#endif
#endif
