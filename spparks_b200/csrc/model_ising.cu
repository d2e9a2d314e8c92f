// kernels of app model SPK_ISING
#include "model_entry.cuh"
template struct spk::ModelEntry<SPK_ISING>;
