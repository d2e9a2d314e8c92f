// Internal: the handle behind the C ABI and the helpers every translation unit of the library shares.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <nccl.h>  // types only: the library is bound lazily with dlopen (spk_cuda.cu: NcclApi)

#include <cstdint>
#include <utility>
#include <vector>

#include "../../include/spparks_cuda.h"
#include "sweep_generic.cuh"

// formats the thread's last-error string (spk_cuda_last_error) and returns 1
int spk_fail(const char *fmt, ...);
#define fail spk_fail

#define CU(call)                                                                        \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) return fail("%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

using namespace spk;

struct spk_cuda_handle {
  spk_cuda_desc d;
  int dim, nb, nn, nc;
  Geo g;
  int64_t nlocal, nglobal, plane;  // plane = elements of one plane (3-D) / row (2-D) of the slab dimension
  int nslab;                       // local extent along the slab dimension
  int gh = 1;                      // ghost planes / rows on each side of the slab
  bool fused2d_ok = false;         // colour-fused 2-D tile kernel usable (tensor map built)
  bool use_fused = true;           // SPK_CUDA_NO_FUSED=1 keeps the per-colour kernels (A/B tests)
  int32_t *base = nullptr;         // allocation incl. ghost planes
  int32_t *spin = nullptr;         // first owned site
  unsigned char *mask_base = nullptr, *mask = nullptr;  // `sweep ... mask yes`: allocation incl. ghost layers, first owned site
  double *u1 = nullptr, *u2 = nullptr;
  int32_t *seeds = nullptr;
  unsigned long long *naccept = nullptr, *escratch = nullptr;
  int64_t nattempt = 0, launches = 0;
  EventParams ev;
  cudaStream_t stream = nullptr, comm_stream = nullptr;
  cudaEvent_t t0 = nullptr, t1 = nullptr, ev_halo = nullptr, ev_compute = nullptr;
  std::vector<cudaEvent_t> laps;  // spk_cuda_timer_lap events since the last timer_start
  size_t nlaps = 0;
  ncclComm_t comm = nullptr;
  bool tiled = true;  // tuned kernels allowed (SPK_CUDA_FORCE_GENERIC=1 disables, for A/B tests)
  bool use_stream = true;  // SPK_CUDA_NO_STREAM=1 keeps the strip kernel (A/B tests)
  bool smem_attr_set[17] = {};  // dynamic shared memory opt-in done: [0] tile2d, [1..8] stream variants
  uint32_t *cl_label = nullptr, *cl_size = nullptr;  // diag_style cluster: union-find parents, volumes / ids per root
  uint32_t *cl_blocksum = nullptr;
  long long *cl_dvol = nullptr;   // ordered cluster volumes (spk_cuda_cluster_ids), capacity cl_dvol_cap
  size_t cl_dvol_cap = 0;
  unsigned long long *cl_count = nullptr;             // [0] ncluster, [1] (as double) sum of radii
  // clusters across slabs: 64-bit global labels / merged volumes per local root, boundary-layer
  // label buffers (out_lo, out_hi, ghost_lo, ghost_hi), foreign-root pairs, flags
  unsigned long long *cl_glabel = nullptr, *cl_gsize = nullptr, *cl_layer = nullptr, *cl_pairs = nullptr;
  int *cl_flags = nullptr;                            // [0] a label changed this round, [1] number of foreign roots
  int32_t *edge = nullptr;     // tile-edge words of the colour-pair launches (StreamMap::edge)
  size_t edge_words = 0;
  bool use_pair = false;       // SPK_CUDA_PAIR=1: sc/26n colour pairs in one launch (half the DRAM traffic; measured slower)
  bool pk16 = false;       // every spin <= PK16_MAX_SPIN: packed 16-bit compares (SPK_CUDA_NO_PK16=1 disables, for A/B tests)
  bool stream_ok = false;  // TMA streaming kernel usable (3-D lattice, tensor map built)
  CUtensorMap tmap;        // (x, y, plane incl. the two ghost planes) view of the slab
  CUtensorMap tmap8;       // same view with the box of the 8-class-row tiles (candidate-elimination variant)
  bool tmap8_ok = false;
  // candidate-list apps on sc/26n: which first-occurrence algorithm the next launches use
  // (SPK_CUDA_NEIGH_MODE=chain|adaptive|auto).  auto follows the mean number of distinct neighbour
  // values per site that the kernels report (kstat), read back without blocking after every sweep.
  int neigh_mode = 2;            // 0 chain, 1 adaptive, 2 auto
  bool neigh_adaptive = false;   // current choice
  double neigh_switch = 2.9;     // mean distinct neighbour values at which the variants break even (SPK_CUDA_NEIGH_SWITCH)
  unsigned long long *kstat = nullptr;       // device: sum of nevent
  unsigned long long *kstat_host = nullptr;  // pinned: [0] copy of *kstat, [1] attempts it covers
  cudaEvent_t ev_kstat = nullptr;
  bool kstat_pending = false;
  bool kstat_discard = false;    // the pending read-back belongs to a lattice that has been replaced
  int64_t nattempt_base = 0;     // attempts before the last spk_cuda_reset_counters
  int64_t kstat_attempt0 = 0;    // nattempt_base + nattempt when the statistic was last reset
  double kstat_mean = -1.0;      // last mean read back (-1: none yet)
  // 16-bit shadow of the lattice (sweep_stream16.cuh): sc/26n potts / potts/pin with every spin
  // <= 65534 sweep on it inside spk_cuda_sweeps; exactly one of the two copies may be stale
  uint16_t *shadow_base = nullptr, *shadow = nullptr;  // allocation incl. the two ghost planes, first owned site
  CUtensorMap tmap16;
  bool shadow_ok = false;       // usable (eligible handle, allocation and tensor map succeeded; SPK_CUDA_NO_SHADOW16=1 disables)
  bool shadow_current = false;  // holds the current spins (incl. ghost planes when ghosts_valid)
  bool canon_current = true;    // the int32 lattice holds the current spins
  bool smem16_set = false;
  // bit shadow of a two-state 2-D lattice (sweep_bits2d.cuh): ising / ising/single on sq/4n, one GPU;
  // shares shadow_current / canon_current with the 16-bit shadow (a handle has at most one of them)
  uint32_t *bits = nullptr;
  bool bits_ok = false;
  void *peer_map16[2] = {nullptr, nullptr};            // the ring neighbours' shadow allocations (peer memory)
  uint16_t *mirror16_lo = nullptr, *mirror16_hi = nullptr;
  int32_t *base_alt = nullptr;  // second lattice buffer of the out-of-place fused 2-D sweep
  CUtensorMap tmap_alt;
  // peer-memory halo (3-D slabs through the streaming kernel): the boundary launches of a colour
  // store straight into the ring neighbours' ghost planes over NVLink, ordered by flag words
  bool p2p = false;
  void *peer_map[2] = {nullptr, nullptr};       // mapped lattice allocations of the lower / upper neighbour
  void *peer_map_alt[2] = {nullptr, nullptr};   // ... and of their second buffers (2-D ping-pong)
  int32_t *alloc0 = nullptr;                    // the allocation `base` pointed to at create (the buffers swap per fused sweep, in lockstep on all ranks)
  void *peer_flag_map[2] = {nullptr, nullptr};  // mapped flag words of the lower / upper neighbour
  uint32_t *flags = nullptr;      // [0] boundary steps finished by the lower neighbour, [1] by the upper, [2] wait timed out
  uint32_t *peer_flags[2] = {nullptr, nullptr};  // the words this rank signals: lower's [1], upper's [0]
  int32_t *mirror_lo = nullptr, *mirror_hi = nullptr;  // lower neighbour's upper ghost plane, upper neighbour's lower ghost plane
  bool mirror_active = false;     // boundary launches of spk_cuda_sweeps only
  bool ghosts_valid = false;      // ghost planes hold the neighbours' current boundary planes
  uint32_t step = 0;              // boundary steps signalled so far (same sequence on every rank)
};

namespace spk {
// sector box (global cells) -> local cell ranges of this slab; false if empty
bool local_box(const spk_cuda_handle *h, const int32_t *box, Box &b);

// Per-application entry points: each model's kernels are instantiated in their own translation
// unit (model_*.cu) so that the library builds in parallel.
template <int MODEL> struct ModelEntry {
  // one colour set (src/app_lattice.cpp:674-691): tuned kernel where it applies, else generic
  static int sweep_set(spk_cuda_handle *h, int rng, int color, const int32_t *box, int nloop);
  // colour-fused 2-D sweep (launch_tile2d): 0 launched, -1 not applicable, >0 error
  static int fused(spk_cuda_handle *h, const SweepArgs &a, int part);
  // one colour of the box on the 16-bit shadow (launch_stream16): 0 launched, -1 not applicable, >0 error
  static int shadow16(spk_cuda_handle *h, int color, const int32_t *box);
  // sc/26n colour pair: 0 launched, -1 not applicable, >0 error
  static int pair(spk_cuda_handle *h, int color, const int32_t *box);
};
}  // namespace spk
