// kernels of app model SPK_POTTS_NEIGH
#include "model_entry.cuh"
template struct spk::ModelEntry<SPK_POTTS_NEIGH>;
