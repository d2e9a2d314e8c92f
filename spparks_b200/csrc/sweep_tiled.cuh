// Tuned colour-set kernels (Philox mode).  launch_tiled returns 0 when it launched a kernel,
// -1 when no tuned kernel applies (caller falls back to sweep_generic_kernel), >0 on error.
#pragma once
#include "sweep_generic.cuh"

namespace spk {

template <int LAT, int MODEL>
inline int launch_tiled(cudaStream_t, const SweepArgs &, int64_t *) { return -1; }

}  // namespace spk
