// kernels of app model SPK_POTTS_NEIGHONLY
#include "model_entry.cuh"
template struct spk::ModelEntry<SPK_POTTS_NEIGHONLY>;
