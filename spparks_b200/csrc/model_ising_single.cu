// kernels of app model SPK_ISING_SINGLE
#include "model_entry.cuh"
template struct spk::ModelEntry<SPK_ISING_SINGLE>;
