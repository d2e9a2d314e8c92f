// kernels of app model SPK_POTTS
#include "model_entry.cuh"
template struct spk::ModelEntry<SPK_POTTS>;
