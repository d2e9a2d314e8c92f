// libspparks_cuda: C ABI (include/spparks_cuda.h) over the sm_100a colour-sweep kernels.
// No CPU fallback: every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types only: the library is bound lazily with dlopen (see NcclApi)

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "handle.h"
#include "aux_kernels.cuh"
#include "cluster.cuh"
#include "sweep_stream.cuh"   // StreamCfg (tensor-map box shapes)
#include "sweep_stream16.cuh" // Stream16Cfg, shadow conversion kernels
#include "sweep_bits2d.cuh"   // multi-spin-coded ising kernel and its conversion kernels
#include "sweep_tile2d.cuh"   // Tile2dCfg

namespace {
thread_local std::string g_err;
}

int spk_fail(const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}

namespace spk {
extern template struct ModelEntry<SPK_ISING>;
extern template struct ModelEntry<SPK_POTTS>;
extern template struct ModelEntry<SPK_POTTS_NEIGHONLY>;
extern template struct ModelEntry<SPK_POTTS_PIN>;
extern template struct ModelEntry<SPK_ISING_SINGLE>;
extern template struct ModelEntry<SPK_POTTS_NEIGH>;
}  // namespace spk

namespace {

// NCCL is bound at the first multi-GPU call, not at link time: a host process (PyTorch, an MPI
// SPPARKS) may already carry its own libnccl.so.2 and two copies must not be mixed.  dlopen by
// SONAME returns the copy that is already loaded, if any.  SPK_CUDA_NCCL_LIB overrides the path.
struct NcclApi {
  void *dl = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
} g_nccl;

int nccl_bind() {
  if (g_nccl.dl) return 0;
  const char *override_path = getenv("SPK_CUDA_NCCL_LIB");
  void *dl = dlopen(override_path ? override_path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!dl) return fail("cannot load NCCL: %s", dlerror());
#define BIND(name)                                                          \
  *(void **)(&g_nccl.name) = dlsym(dl, "nccl" #name);                       \
  if (!g_nccl.name) return fail("NCCL symbol nccl" #name " not found")
  BIND(GetUniqueId); BIND(CommInitRank); BIND(CommDestroy); BIND(GroupStart); BIND(GroupEnd);
  BIND(Send); BIND(Recv); BIND(AllGather); BIND(AllReduce); BIND(GetErrorString);
#undef BIND
  g_nccl.dl = dl;
  return 0;
}

#define NC(call)                                                                        \
  do {                                                                                  \
    ncclResult_t r_ = (call);                                                           \
    if (r_ != ncclSuccess) return fail("%s: %s (%s:%d)", #call, g_nccl.GetErrorString(r_), __FILE__, __LINE__); \
  } while (0)

int lat_dim(int lat) { return (lat == SPK_SC_6N || lat == SPK_SC_26N) ? 3 : 2; }

}  // namespace


namespace {

void build_tables(spk_cuda_handle *h) {
  // exp((einitial-efinal)*t_inverse) of site_event_rejection (src/app_potts.cpp:130), evaluated
  // with the host libm for every possible integer dE so that device results are bit-identical
  EventParams &ev = h->ev;
  ev.nspins = ising_like(h->d.model) ? 2 : h->d.nspins;
  ev.t_zero = h->d.temperature == 0.0;
  ev.one = 1;
  ev.k0 = (uint32_t)h->d.seed;
  ev.k1 = (uint32_t)(h->d.seed >> 32);
  for (int r = 0; r < 10; r++) {  // Philox key schedule: key + r * Weyl constants
    ev.rk.k[2 * r] = ev.k0 + (uint32_t)r * 0x9E3779B9u;
    ev.rk.k[2 * r + 1] = ev.k1 + (uint32_t)r * 0xBB67AE85u;
  }
  ev.sweep = 0;
  ev.m65536 = -65536;
  ev.rep = 0x00010001u;
  for (int n = 0; n < 16; n++) ev.pow2[n] = 1u << n;
  const double t_inverse = ev.t_zero ? 0.0 : 1.0 / h->d.temperature;
  for (int dE = 0; dE < 27; dE++) {
    double p = ev.t_zero ? 0.0 : std::exp((0.0 - (double)dE) * t_inverse);
    ev.ptab[dE] = p;
    double t = std::floor(p * 4294967296.0);
    ev.thr[dE] = t > 4294967295.0 ? 4294967295u : (uint32_t)t;
  }
}

}  // namespace

// sector box (global cells) -> local cell ranges of this slab; false if empty
bool spk::local_box(const spk_cuda_handle *h, const int32_t *box, Box &b) {
  const Geo &g = h->g;
  int lo[3] = {0, 0, 0}, hi[3] = {g.gX, g.gY, g.gZ};
  if (box)
    for (int a = 0; a < 3; a++) {
      lo[a] = box[2 * a] < 0 ? 0 : box[2 * a];
      hi[a] = box[2 * a + 1] > hi[a] ? hi[a] : box[2 * a + 1];
    }
  const int sd = h->dim - 1;  // slab dimension
  lo[sd] -= g.g0; hi[sd] -= g.g0;
  if (lo[sd] < 0) lo[sd] = 0;
  if (hi[sd] > h->nslab) hi[sd] = h->nslab;
  b = Box{lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]};
  return lo[0] < hi[0] && lo[1] < hi[1] && lo[2] < hi[2];
}

namespace {

// cuTensorMapEncodeTiled is fetched through the runtime so that libcuda is not a link dependency.
int build_tensor_map(spk_cuda_handle *h) {
  h->stream_ok = false;
  h->fused2d_ok = false;
  if (h->dim == 3 && (h->nb != 1 || h->g.X % 4)) return 0;
  if (h->dim == 2 && (h->g.X * h->nb) % 8) return 0;
  typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
    cudaGetLastError();
    return 0;
  }
  // L2 promotion of the box fetches: 64 bytes measured best (potts/pin sc/26n 1024^3: 123.1 Gsites/s;
  // none / 128 B 122.1; 256 B 120.6 and +5 % DRAM reads: a box row starts 16 bytes before a 128-byte
  // line, so wide promotion drags in lines of the neighbouring tile that have left L2 again by the
  // time that tile wants them).  SPK_CUDA_L2PROMO=0..3 (none, 64, 128, 256 bytes) overrides.
  CUtensorMapL2promotion promo = h->dim == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  if (const char *e = getenv("SPK_CUDA_L2PROMO")) {
    const int v = atoi(e);
    promo = v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
            : v == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  }
  if (h->dim == 2) {  // (element, row) view of the slab, row 0 = lowest ghost row
    const int bh = h->d.lattice == SPK_SQ_4N ? Tile2dCfg<SPK_SQ_4N>::BH
                   : h->d.lattice == SPK_SQ_8N ? Tile2dCfg<SPK_SQ_8N>::BH : Tile2dCfg<SPK_TRI>::BH;
    cuuint64_t dims[2] = {(cuuint64_t)h->plane, (cuuint64_t)h->nslab + 2 * h->gh};
    cuuint64_t strides[1] = {(cuuint64_t)h->plane * 4};
    cuuint32_t box[2] = {(cuuint32_t)T2_BW, (cuuint32_t)bh};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ((encode_fn)fn)(&h->tmap, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, h->base, dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS && h->base_alt)
      r = ((encode_fn)fn)(&h->tmap_alt, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, h->base_alt, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    h->fused2d_ok = (r == CUDA_SUCCESS && h->base_alt);
    return 0;
  }
  const int rp = h->d.lattice == SPK_SC_26N ? StreamCfg<SPK_SC_26N>::RP : StreamCfg<SPK_SC_6N>::RP;
  cuuint64_t dims[3] = {(cuuint64_t)h->g.X, (cuuint64_t)h->g.Y, (cuuint64_t)h->nslab + 2};
  cuuint64_t strides[2] = {(cuuint64_t)h->g.X * 4, (cuuint64_t)h->g.X * h->g.Y * 4};
  cuuint32_t box[3] = {(cuuint32_t)ST_BX, (cuuint32_t)rp, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = ((encode_fn)fn)(&h->tmap, CU_TENSOR_MAP_DATA_TYPE_INT32, 3, h->base, dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  h->stream_ok = (r == CUDA_SUCCESS);
  // 16-bit shadow (sweep_stream16.cuh): the tensor is described in 32-bit words of two sites
  h->shadow_ok = false;
  {
    const char *ns16 = getenv("SPK_CUDA_NO_SHADOW16");
    const bool want = h->stream_ok && h->d.lattice == SPK_SC_26N && h->pk16 && h->g.X % 8 == 0 &&
                      (h->d.model == SPK_POTTS || h->d.model == SPK_POTTS_PIN) && !(ns16 && ns16[0] == '1');
    if (want) {
      const size_t n16 = (size_t)(h->nlocal + 2 * h->plane);
      if (!h->shadow_base && cudaMalloc(&h->shadow_base, n16 * sizeof(uint16_t)) != cudaSuccess) {
        cudaGetLastError();
        h->shadow_base = nullptr;
      }
      if (h->shadow_base) {
        h->shadow = h->shadow_base + h->plane;
        cuuint64_t d16[3] = {(cuuint64_t)h->g.X / 2, (cuuint64_t)h->g.Y, (cuuint64_t)h->nslab + 2};
        cuuint64_t s16[2] = {(cuuint64_t)h->g.X * 2, (cuuint64_t)h->g.X * h->g.Y * 2};
        cuuint32_t b16[3] = {(cuuint32_t)S16_BX / 2, (cuuint32_t)Stream16Cfg::RP, 1};
        r = ((encode_fn)fn)(&h->tmap16, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, h->shadow_base, d16, s16, b16, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        h->shadow_ok = (r == CUDA_SUCCESS);
      }
    }
  }
  h->tmap8_ok = false;
  if (h->stream_ok && h->d.lattice == SPK_SC_26N && neigh_like(h->d.model) && h->pk16) {
    cuuint32_t box8[3] = {(cuuint32_t)ST_BX, (cuuint32_t)StreamCfgOf<SPK_SC_26N, 1>::RP, 1};
    r = ((encode_fn)fn)(&h->tmap8, CU_TENSOR_MAP_DATA_TYPE_INT32, 3, h->base, dims, strides, box8, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    h->tmap8_ok = (r == CUDA_SUCCESS);
  }
  return 0;
}

// ---- the two copies of an eligible lattice (int32 canonical, 16-bit shadow): exactly one may be
// stale; conversions cover the ghost planes too, so `ghosts_valid` speaks for both
int ensure_canon(spk_cuda_handle *h) {
  if (h->canon_current) return 0;
  if (h->bits_ok) {
    bits2d_unpack_kernel<<<148 * 8, 256, 0, h->stream>>>(h->bits, h->spin, h->g.Y, h->g.X / 64);
    h->launches++;
    CU(cudaGetLastError());
    h->canon_current = true;
    return 0;
  }
  const long long n8 = (h->nlocal + 2 * h->plane) / 8;
  shadow_unpack_kernel<<<148 * 8, 256, 0, h->stream>>>(reinterpret_cast<const uint4 *>(h->shadow_base),
                                                       reinterpret_cast<int4 *>(h->base), n8);
  h->launches++;
  CU(cudaGetLastError());
  h->canon_current = true;
  return 0;
}
int ensure_shadow(spk_cuda_handle *h) {
  if (h->shadow_current) return 0;
  if (h->bits_ok) {
    bits2d_pack_kernel<<<148 * 8, 256, 0, h->stream>>>(h->spin, h->bits, h->g.Y, h->g.X / 64);
    h->launches++;
    CU(cudaGetLastError());
    h->shadow_current = true;
    return 0;
  }
  const long long n8 = (h->nlocal + 2 * h->plane) / 8;
  shadow_pack_kernel<<<148 * 8, 256, 0, h->stream>>>(reinterpret_cast<const int4 *>(h->base),
                                                     reinterpret_cast<uint4 *>(h->shadow_base), n8);
  h->launches++;
  CU(cudaGetLastError());
  h->shadow_current = true;
  return 0;
}
// the caller is about to read (write = false) or modify (write = true) the int32 lattice
int use_canon(spk_cuda_handle *h, bool write) {
  if (ensure_canon(h)) return 1;
  if (write) h->shadow_current = false;
  return 0;
}

// the spins were replaced (upload, randomize): forget what the statistic said about the old lattice
void neigh_mode_reset(spk_cuda_handle *h) {
  if (!h->kstat || h->neigh_mode != 2) return;
  h->neigh_adaptive = false;
  h->kstat_mean = -1.0;
  h->kstat_discard = h->kstat_pending;
  cudaMemsetAsync(h->kstat, 0, sizeof(unsigned long long), h->stream);
  h->kstat_attempt0 = h->nattempt_base + h->nattempt;
}

// Candidate-list apps on sc/26n: choose between the all-pairs chain (cost independent of the
// lattice) and candidate elimination (cost proportional to the distinct neighbour values) from
// the mean nevent the kernels accumulate in h->kstat.  Called at the first colour of a sweep;
// never blocks: the statistic is copied to pinned memory behind the launches already queued and
// picked up one or two sweeps later.  Both variants give identical spins, only the speed differs.
int neigh_mode_update(spk_cuda_handle *h) {
  if (!h->kstat || h->neigh_mode != 2) return 0;
  if (h->kstat_pending && cudaEventQuery(h->ev_kstat) == cudaSuccess) {
    h->kstat_pending = false;
    if (h->kstat_discard) h->kstat_discard = false;  // measured on a lattice that has been replaced since
    else if (h->kstat_host[1] > 0) {
      h->kstat_mean = (double)h->kstat_host[0] / (double)h->kstat_host[1];
      // hysteresis around the measured break-even (SPK_CUDA_NEIGH_SWITCH, distinct values per site)
      if (h->kstat_mean < h->neigh_switch - 0.25) h->neigh_adaptive = true;
      else if (h->kstat_mean > h->neigh_switch + 0.25) h->neigh_adaptive = false;
    }
  }
  cudaGetLastError();  // cudaErrorNotReady of the query is not an error
  const int64_t attempts = h->nattempt_base + h->nattempt;  // monotonic (reset_counters moves the base)
  if (!h->kstat_pending && attempts > h->kstat_attempt0) {
    CU(cudaMemcpyAsync(h->kstat_host, h->kstat, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemsetAsync(h->kstat, 0, sizeof(unsigned long long), h->stream));
    CU(cudaEventRecord(h->ev_kstat, h->stream));
    h->kstat_host[1] = (unsigned long long)(attempts - h->kstat_attempt0);
    h->kstat_attempt0 = attempts;
    h->kstat_pending = true;
  }
  return 0;
}

// periodic images of the slab dimension for a single-GPU run (the multi-GPU case is
// spk_cuda_halo_exchange): ghost plane -1 <- last owned plane, ghost plane n <- first owned plane
int refresh_ghosts_local(spk_cuda_handle *h) {
  const size_t bytes = (size_t)h->plane * sizeof(int32_t);
  if (h->nslab >= h->gh) {
    CU(cudaMemcpyAsync(h->base, h->spin + (int64_t)(h->nslab - h->gh) * h->plane, bytes * h->gh,
                       cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaMemcpyAsync(h->spin + h->nlocal, h->spin, bytes * h->gh, cudaMemcpyDeviceToDevice, h->stream));
    h->launches += 2;
    return 0;
  }
  for (int m = 1; m <= h->gh; m++) {  // slab thinner than the ghost depth: images wrap more than once
    const int lo = ((h->nslab - m) % h->nslab + h->nslab) % h->nslab, hi = (m - 1) % h->nslab;
    CU(cudaMemcpyAsync(h->spin - (int64_t)m * h->plane, h->spin + (int64_t)lo * h->plane, bytes,
                       cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaMemcpyAsync(h->spin + h->nlocal + (int64_t)(m - 1) * h->plane, h->spin + (int64_t)hi * h->plane, bytes,
                       cudaMemcpyDeviceToDevice, h->stream));
    h->launches += 2;
  }
  return 0;
}

int fused_sweep_m(spk_cuda_handle *h, const SweepArgs &a, int part) {
  switch (h->d.model) {
    case SPK_ISING: return ModelEntry<SPK_ISING>::fused(h, a, part);
    case SPK_POTTS: return ModelEntry<SPK_POTTS>::fused(h, a, part);
    case SPK_POTTS_NEIGHONLY: return ModelEntry<SPK_POTTS_NEIGHONLY>::fused(h, a, part);
    case SPK_POTTS_PIN: return ModelEntry<SPK_POTTS_PIN>::fused(h, a, part);
    case SPK_ISING_SINGLE: return ModelEntry<SPK_ISING_SINGLE>::fused(h, a, part);
    case SPK_POTTS_NEIGH: return ModelEntry<SPK_POTTS_NEIGH>::fused(h, a, part);
  }
  return -1;
}
// -1 not applicable, >0 error.  On one GPU the ghost rows are refreshed here.
int fused_sweep(spk_cuda_handle *h, int part) {
  if (!h->tiled || !h->use_fused || h->dim != 2 || !h->fused2d_ok) return -1;
  SweepArgs a;
  a.spin = h->spin; a.g = h->g; a.c0 = 0; a.nloop = 1; a.ev = h->ev;
  a.u1 = nullptr; a.u2 = nullptr; a.seeds = nullptr; a.naccept = h->naccept; a.kstat = nullptr; a.mask = nullptr;
  a.mirror_lo = a.mirror_hi = nullptr;
  a.b = Box{0, h->g.X, 0, h->g.Y, 0, 1};
  if (h->d.nranks == 1 && refresh_ghosts_local(h)) return 1;
  const int rc = fused_sweep_m(h, a, part);
  if (rc == 0 && part != 1) h->nattempt += h->nlocal;
  return rc;
}


bool pair_ok(const spk_cuda_handle *h) {
  return h->use_pair && h->tiled && h->use_stream && h->stream_ok && h->d.lattice == SPK_SC_26N && !ising_like(h->d.model);
}

// one colour class on the 16-bit shadow: 0 launched, -1 not applicable, >0 error
int shadow16_m(spk_cuda_handle *h, int color, const int32_t *box) {
  switch (h->d.model) {
    case SPK_POTTS: return ModelEntry<SPK_POTTS>::shadow16(h, color, box);
    case SPK_POTTS_PIN: return ModelEntry<SPK_POTTS_PIN>::shadow16(h, color, box);
    default: return -1;
  }
}
// spk_cuda_sweeps may run on the shadow: eligible handle, tuned kernels allowed, no masks, no colour pairs
bool shadow_usable(const spk_cuda_handle *h) {
  return h->shadow_ok && h->tiled && h->use_stream && !h->mask && !h->use_pair;
}

int sweep_pair_m(spk_cuda_handle *h, int color, const int32_t *box) {
  switch (h->d.model) {
    case SPK_ISING: return ModelEntry<SPK_ISING>::pair(h, color, box);
    case SPK_POTTS: return ModelEntry<SPK_POTTS>::pair(h, color, box);
    case SPK_POTTS_NEIGHONLY: return ModelEntry<SPK_POTTS_NEIGHONLY>::pair(h, color, box);
    case SPK_POTTS_PIN: return ModelEntry<SPK_POTTS_PIN>::pair(h, color, box);
    case SPK_ISING_SINGLE: return ModelEntry<SPK_ISING_SINGLE>::pair(h, color, box);
    case SPK_POTTS_NEIGH: return ModelEntry<SPK_POTTS_NEIGH>::pair(h, color, box);
  }
  return -1;
}

int sweep_pair(spk_cuda_handle *h, int color, const int32_t *box) {
  const int rc = sweep_pair_m(h, color, box);
  if (rc > 0) return fail("colour-pair launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  if (rc < 0) return fail("colour-pair launch not applicable");
  return 0;
}

int sweep_set(spk_cuda_handle *h, int rng, int color, const int32_t *box, int nloop) {
  if (!h) return fail("null handle");
  if (!h->mirror_active) h->ghosts_valid = false;
  if (color < 1 || color > h->nc) return fail("colour %d out of range 1..%d", color, h->nc);
  if (nloop < 1) return fail("nloop must be >= 1");
  CU(cudaSetDevice(h->d.device));
  if (use_canon(h, true)) return 1;
  if (color == 1 && rng == RNG_PHILOX && neigh_mode_update(h)) return 1;
  switch (h->d.model) {
    case SPK_ISING: return ModelEntry<SPK_ISING>::sweep_set(h, rng, color, box, nloop);
    case SPK_POTTS: return ModelEntry<SPK_POTTS>::sweep_set(h, rng, color, box, nloop);
    case SPK_POTTS_NEIGHONLY: return ModelEntry<SPK_POTTS_NEIGHONLY>::sweep_set(h, rng, color, box, nloop);
    case SPK_POTTS_PIN: return ModelEntry<SPK_POTTS_PIN>::sweep_set(h, rng, color, box, nloop);
    case SPK_ISING_SINGLE: return ModelEntry<SPK_ISING_SINGLE>::sweep_set(h, rng, color, box, nloop);
    case SPK_POTTS_NEIGH: return ModelEntry<SPK_POTTS_NEIGH>::sweep_set(h, rng, color, box, nloop);
  }
  return fail("unknown model %d", h->d.model);
}

}  // namespace

// ================================================================================== ABI

extern "C" {

const char *spk_cuda_last_error(void) { return g_err.c_str(); }
int spk_cuda_abi_version(void) { return SPK_CUDA_ABI_VERSION; }

int spk_cuda_device_count(int *count) {
  if (!count) return fail("null argument");
  CU(cudaGetDeviceCount(count));
  return 0;
}

int spk_cuda_ncolors(int lat) {
  switch (lat) {
    case SPK_SQ_4N: case SPK_SC_6N: return 2;
    case SPK_SQ_8N: case SPK_TRI: return 4;
    case SPK_SC_26N: return 8;
  }
  return 0;
}
int spk_cuda_maxneigh(int lat) {
  switch (lat) {
    case SPK_SQ_4N: return 4;
    case SPK_SQ_8N: return 8;
    case SPK_TRI: case SPK_SC_6N: return 6;
    case SPK_SC_26N: return 26;
  }
  return 0;
}
int spk_cuda_nbasis(int lat) { return lat == SPK_TRI ? 2 : (spk_cuda_maxneigh(lat) ? 1 : 0); }

int spk_cuda_create(const spk_cuda_desc *desc, spk_cuda_handle **out) {
  if (!desc || !out) return fail("null argument");
  *out = nullptr;
  const int nn = spk_cuda_maxneigh(desc->lattice);
  if (!nn) return fail("unsupported lattice style %d", desc->lattice);
  if (desc->model < SPK_ISING || desc->model > SPK_POTTS_NEIGH) return fail("unsupported app model %d", desc->model);
  const int dim = lat_dim(desc->lattice);
  const int nz = dim == 3 ? desc->nz : 1;
  if (desc->nx < 2 || desc->ny < 2 || nz < 1) return fail("lattice extents must be >= 2");
  // Lattice::ncolors: "Color stencil is incommensurate with lattice size" (src/lattice.cpp:198-221);
  // tri only needs nx even
  if (desc->nx % 2 || (desc->lattice != SPK_TRI && desc->ny % 2) || (dim == 3 && nz % 2))
    return fail("Color stencil is incommensurate with lattice size");
  if (!ising_like(desc->model) && desc->nspins <= 0) return fail("Illegal app_style command: nspins must be > 0");
  if (desc->temperature < 0.0) return fail("temperature must be >= 0");
  if (desc->nranks < 1 || desc->rank < 0 || desc->rank >= desc->nranks) return fail("bad rank/nranks");
  const int nslab_g = dim == 3 ? nz : desc->ny;
  // slabs must keep the colour parity of the global lattice (even thickness) and hold >= 2 layers
  if (desc->nranks > 1 && (nslab_g % desc->nranks || (nslab_g / desc->nranks) % 2 || nslab_g / desc->nranks < 2))
    return fail("slab dimension %d does not split into %d even slabs", nslab_g, desc->nranks);

  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (desc->device < 0 || desc->device >= ndev) return fail("CUDA device %d not available (%d visible)", desc->device, ndev);
  CU(cudaSetDevice(desc->device));

  spk_cuda_handle *h = new spk_cuda_handle;
  h->d = *desc;
  h->d.nz = nz;
  h->dim = dim; h->nn = nn; h->nb = spk_cuda_nbasis(desc->lattice); h->nc = spk_cuda_ncolors(desc->lattice);
  h->nslab = nslab_g / desc->nranks;
  Geo &g = h->g;
  g.gX = desc->nx; g.gY = desc->ny; g.gZ = nz;
  g.X = desc->nx; g.Y = dim == 3 ? desc->ny : h->nslab; g.Z = dim == 3 ? h->nslab : 1;
  g.g0 = desc->rank * h->nslab;
  if (desc->nranks == 1) { g.lo_wrap = h->nslab - 1; g.hi_wrap = 0; }
  else { g.lo_wrap = -1; g.hi_wrap = h->nslab; }
  h->plane = dim == 3 ? (int64_t)g.X * g.Y * h->nb : (int64_t)g.X * h->nb;
  h->nlocal = h->plane * h->nslab;
  h->nglobal = (int64_t)desc->nx * desc->ny * nz * h->nb;
  const char *fg = getenv("SPK_CUDA_FORCE_GENERIC");
  h->tiled = !(fg && fg[0] == '1');
  const char *ns = getenv("SPK_CUDA_NO_STREAM");
  h->use_stream = !(ns && ns[0] == '1');
  const char *nf = getenv("SPK_CUDA_NO_FUSED");
  h->use_fused = !(nf && nf[0] == '1');
  const char *pairenv = getenv("SPK_CUDA_PAIR");
  h->use_pair = pairenv && pairenv[0] == '1';
  const char *np = getenv("SPK_CUDA_NO_PK16");
  // spins are 1..Q (potts/pin: Q+1 marks a pinned site, src/app_potts_pin.cpp:81)
  h->pk16 = !(np && np[0] == '1') && !ising_like(desc->model) &&
            (int64_t)desc->nspins + (desc->model == SPK_POTTS_PIN ? 1 : 0) <= PK16_MAX_SPIN;
  build_tables(h);

  auto cleanup = [&](int rc) { spk_cuda_destroy(h); return rc; };
  // 2-D slabs keep T2_GH ghost rows per side: the colour-fused tile kernel needs one row per
  // fused colour step; 3-D slabs one plane (delpropensity = 1)
  h->gh = dim == 2 ? T2_GH : 1;
  if (desc->nranks > 1 && h->nslab < h->gh) {
    const int gh = h->gh;
    delete h;
    return fail("slab thinner than its %d ghost layers", gh);
  }
  const size_t bytes = (size_t)(h->nlocal + 2 * (int64_t)h->gh * h->plane) * sizeof(int32_t);
  if (cudaMalloc(&h->base, bytes) != cudaSuccess) {
    cudaGetLastError();
    return cleanup(fail("cudaMalloc of %zu bytes for the lattice failed", bytes));
  }
  h->alloc0 = h->base;
  h->spin = h->base + (int64_t)h->gh * h->plane;
  if (cudaMemset(h->base, 0, bytes) != cudaSuccess) return cleanup(fail("cudaMemset failed"));
  if (dim == 2 && h->tiled && h->use_fused && (h->plane % 8) == 0) {
    // destination buffer of the out-of-place colour-fused sweep (2-D lattices only)
    if (cudaMalloc(&h->base_alt, bytes) != cudaSuccess) {
      cudaGetLastError();
      h->base_alt = nullptr;  // not fatal: the per-colour kernels need no second buffer
    } else if (cudaMemset(h->base_alt, 0, bytes) != cudaSuccess) {
      return cleanup(fail("cudaMemset failed"));
    }
  }
  build_tensor_map(h);
  {  // bit shadow: two-state app on sq/4n, rows of whole 64-cell groups, one GPU
    const char *nb = getenv("SPK_CUDA_NO_BITS");
    if (ising_like(desc->model) && desc->lattice == SPK_SQ_4N && desc->nranks == 1 && desc->nx % 64 == 0 && h->tiled &&
        h->use_fused && !(nb && nb[0] == '1')) {
      if (cudaMalloc(&h->bits, (size_t)h->nlocal / 8) == cudaSuccess) h->bits_ok = true;
      else cudaGetLastError();
    }
  }
  if (h->tmap8_ok) {
    const char *nm = getenv("SPK_CUDA_NEIGH_MODE");
    h->neigh_mode = !nm ? 2 : !strcmp(nm, "chain") ? 0 : !strcmp(nm, "adaptive") ? 1 : 2;
    h->neigh_adaptive = h->neigh_mode == 1;
    if (const char *e = getenv("SPK_CUDA_NEIGH_SWITCH")) {
      const double v = atof(e);
      if (v > 0.0) h->neigh_switch = v;
    }
    if (cudaMalloc(&h->kstat, sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(h->kstat, 0, sizeof(unsigned long long)) != cudaSuccess ||
        cudaMallocHost(&h->kstat_host, 2 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_kstat, cudaEventDisableTiming) != cudaSuccess)
      return cleanup(fail("allocation of the candidate statistic failed"));
    h->kstat_host[0] = h->kstat_host[1] = 0;
  }
  if (cudaMalloc(&h->naccept, 2 * sizeof(unsigned long long)) != cudaSuccess) return cleanup(fail("cudaMalloc failed"));
  h->escratch = h->naccept + 1;
  cudaMemset(h->naccept, 0, 2 * sizeof(unsigned long long));
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&h->t0) != cudaSuccess || cudaEventCreate(&h->t1) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_halo, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_compute, cudaEventDisableTiming) != cudaSuccess)
    return cleanup(fail("stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError())));
  *out = h;
  return 0;
}

int spk_cuda_destroy(spk_cuda_handle *h) {
  if (!h) return 0;
  cudaSetDevice(h->d.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->comm_stream) cudaStreamSynchronize(h->comm_stream);
  for (int side = 0; side < 2; side++) {
    if (side == 1 && h->peer_map[1] == h->peer_map[0]) break;
    if (h->peer_map[side]) cudaIpcCloseMemHandle(h->peer_map[side]);
    if (h->peer_map_alt[side]) cudaIpcCloseMemHandle(h->peer_map_alt[side]);
    if (h->peer_flag_map[side]) cudaIpcCloseMemHandle(h->peer_flag_map[side]);
    if (h->peer_map16[side]) cudaIpcCloseMemHandle(h->peer_map16[side]);
  }
  // Peer memory: an exported allocation must not be freed while a neighbour still has it mapped
  // (cudaIpcGetMemHandle: "calling cudaFree on an exported memory region before calling
  // cudaIpcCloseMemHandle in the importing context will result in undefined behavior" — seen as
  // wrong spins in LATER handles of the same processes at 8 GPUs).  Every rank has closed its
  // mappings above; one all-reduce makes sure of it before anybody frees.  Destroy of a
  // multi-rank handle is therefore collective, like its create + comm_init.
  if (h->p2p && h->comm && g_nccl.AllReduce && h->flags) {
    if (g_nccl.AllReduce(h->flags + 3, h->flags + 3, 1, ncclInt32, ncclMax, h->comm, h->stream) == ncclSuccess)
      cudaStreamSynchronize(h->stream);
    cudaGetLastError();
  }
  cudaFree(h->flags);
  cudaFree(h->edge);
  cudaFree(h->kstat);
  if (h->kstat_host) cudaFreeHost(h->kstat_host);
  if (h->ev_kstat) cudaEventDestroy(h->ev_kstat);
  cudaFree(h->cl_label); cudaFree(h->cl_size); cudaFree(h->cl_blocksum); cudaFree(h->cl_count);
  cudaFree(h->cl_dvol);
  cudaFree(h->cl_glabel); cudaFree(h->cl_gsize); cudaFree(h->cl_layer); cudaFree(h->cl_pairs); cudaFree(h->cl_flags);
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  cudaFree(h->mask_base);
  cudaFree(h->shadow_base);
  cudaFree(h->bits);
  cudaFree(h->base); cudaFree(h->base_alt); cudaFree(h->u1); cudaFree(h->u2); cudaFree(h->seeds); cudaFree(h->naccept);
  for (cudaEvent_t e : h->laps) cudaEventDestroy(e);
  if (h->t0) cudaEventDestroy(h->t0);
  if (h->t1) cudaEventDestroy(h->t1);
  if (h->ev_halo) cudaEventDestroy(h->ev_halo);
  if (h->ev_compute) cudaEventDestroy(h->ev_compute);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
  delete h;
  return 0;
}

int64_t spk_cuda_nlocal(const spk_cuda_handle *h) { return h ? h->nlocal : 0; }
int64_t spk_cuda_nglobal(const spk_cuda_handle *h) { return h ? h->nglobal : 0; }

int spk_cuda_upload_spins(spk_cuda_handle *h, const int32_t *host, int64_t n) {
  if (!h || !host) return fail("null argument");
  if (n != h->nlocal) return fail("upload of %lld sites, slab owns %lld", (long long)n, (long long)h->nlocal);
  CU(cudaSetDevice(h->d.device));
  CU(cudaMemcpyAsync(h->spin, host, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  h->ghosts_valid = false;
  h->canon_current = true;   // every owned site replaced: nothing of the shadow is needed
  h->shadow_current = false;
  neigh_mode_reset(h);
  // the reference's init_app refuses spins outside 1..Q (Q+1 = pinned for potts/pin); the kernels
  // rely on it (packed 16-bit compares), so the check is repeated where the spins enter the device
  return spk_cuda_validate_spins(h);
}

int spk_cuda_validate_spins(spk_cuda_handle *h) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->d.device));
  if (use_canon(h, false)) return 1;
  const long long n = h->nlocal;
  const int hi = ising_like(h->d.model) ? 2 : h->d.nspins + (h->d.model == SPK_POTTS_PIN ? 1 : 0);
  CU(cudaMemsetAsync(h->escratch, 0, sizeof(unsigned long long), h->stream));
  validate_kernel<<<148 * 8, 256, 0, h->stream>>>(h->spin, n, 1, hi, h->escratch);
  h->launches++;
  unsigned long long bad = 0;
  CU(cudaMemcpyAsync(&bad, h->escratch, sizeof bad, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (bad) return fail("One or more sites have invalid values (%llu of %lld outside 1..%d)", bad, n, hi);
  return 0;
}

int spk_cuda_download_spins(spk_cuda_handle *h, int32_t *host, int64_t n) {
  if (!h || !host) return fail("null argument");
  if (n != h->nlocal) return fail("download of %lld sites, slab owns %lld", (long long)n, (long long)h->nlocal);
  CU(cudaSetDevice(h->d.device));
  if (use_canon(h, false)) return 1;
  CU(cudaMemcpyAsync(host, h->spin, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int spk_cuda_accept_thresholds(const spk_cuda_handle *h, uint32_t *thr, int n) {
  if (!h || !thr) return fail("null argument");
  if (n < 1 || n > 27) return fail("thresholds exist for dE = 0 .. 26");
  for (int dE = 0; dE < n; dE++) thr[dE] = h->ev.thr[dE];
  return 0;
}

int spk_cuda_set_temperature(spk_cuda_handle *h, double temperature) {
  if (!h) return fail("null handle");
  if (temperature < 0.0) return fail("temperature must be >= 0");
  h->d.temperature = temperature;
  const uint32_t sweep = h->ev.sweep;
  build_tables(h);
  h->ev.sweep = sweep;
  return 0;
}

int spk_cuda_set_seed(spk_cuda_handle *h, uint64_t seed) {
  if (!h) return fail("null handle");
  h->d.seed = seed;
  const uint32_t sweep = h->ev.sweep;
  build_tables(h);  // the Philox key and its round keys
  h->ev.sweep = sweep;
  return 0;
}

int spk_cuda_randomize(spk_cuda_handle *h, uint64_t seed, int32_t lo, int32_t hi) {
  if (!h) return fail("null handle");
  if (hi < lo) return fail("empty value range");
  {  // the kernels assume spins in 1..Q (Q+1 = pinned, potts/pin), as the reference's init_app enforces
    const int top = ising_like(h->d.model) ? 2 : h->d.nspins + (h->d.model == SPK_POTTS_PIN ? 1 : 0);
    if (lo < 1 || hi > top) return fail("One or more sites have invalid values (range %d..%d outside 1..%d)", lo, hi, top);
  }
  CU(cudaSetDevice(h->d.device));
  if (use_canon(h, true)) return 1;
  const int64_t goff = (int64_t)h->g.g0 * h->plane;
  randomize_kernel<<<148 * 8, 256, 0, h->stream>>>(h->spin, h->nlocal, goff, seed, lo, (uint32_t)(hi - lo + 1));
  h->ghosts_valid = false;
  neigh_mode_reset(h);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

int spk_cuda_set_site_range(spk_cuda_handle *h, int32_t park_seed, int32_t lo, int32_t hi) {
  if (!h) return fail("null handle");
  if (hi < lo) return fail("Illegal set command");
  if (park_seed < 1 || park_seed >= 2147483647) return fail("Park-Miller seed must lie in [1, 2^31-2]");
  const int top = ising_like(h->d.model) ? 2 : h->d.nspins + (h->d.model == SPK_POTTS_PIN ? 1 : 0);
  if (lo < 1 || hi > top) return fail("One or more sites have invalid values (range %d..%d outside 1..%d)", lo, hi, top);
  CU(cudaSetDevice(h->d.device));
  if (use_canon(h, true)) return 1;
  const long long goff = (long long)h->g.g0 * h->plane;
  set_range_kernel<<<148 * 8, 256, 0, h->stream>>>(h->spin, h->nlocal, goff, (uint32_t)park_seed, lo, hi - lo + 1);
  h->ghosts_valid = false;
  neigh_mode_reset(h);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

static int p2p_read_fence(spk_cuda_handle *h);

int spk_cuda_site_energy(spk_cuda_handle *h, unsigned char *host, int64_t n) {
  if (!h || !host) return fail("null argument");
  if (n != h->nlocal) return fail("site energies of %lld sites, slab owns %lld", (long long)n, (long long)h->nlocal);
  CU(cudaSetDevice(h->d.device));
  if (use_canon(h, false)) return 1;
  int rc = spk_cuda_halo_exchange(h);  // site_energy reads the ghosts (src/dump.cpp: comm before the pack)
  if (rc) return rc;
  unsigned char *d = nullptr;
  CU(cudaMalloc(&d, (size_t)n));
  const long long ncell = (long long)h->g.X * h->g.Y * h->g.Z;
  const int grid = (int)std::min<long long>(148 * 8, (ncell + 255) / 256);
  switch (h->d.lattice) {
    case SPK_SQ_4N: site_energy_kernel<SPK_SQ_4N><<<grid, 256, 0, h->stream>>>(h->spin, h->g, d); break;
    case SPK_SQ_8N: site_energy_kernel<SPK_SQ_8N><<<grid, 256, 0, h->stream>>>(h->spin, h->g, d); break;
    case SPK_TRI: site_energy_kernel<SPK_TRI><<<grid, 256, 0, h->stream>>>(h->spin, h->g, d); break;
    case SPK_SC_6N: site_energy_kernel<SPK_SC_6N><<<grid, 256, 0, h->stream>>>(h->spin, h->g, d); break;
    case SPK_SC_26N: site_energy_kernel<SPK_SC_26N><<<grid, 256, 0, h->stream>>>(h->spin, h->g, d); break;
  }
  h->launches++;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && p2p_read_fence(h)) e = cudaErrorUnknown;
  if (e == cudaSuccess) e = cudaMemcpyAsync(host, d, (size_t)n, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d);
  if (e != cudaSuccess) return fail("site energy: %s", cudaGetErrorString(e));
  return 0;
}

int spk_cuda_pin_random(spk_cuda_handle *h, uint64_t seed, double fraction) {
  if (!h) return fail("null handle");
  if (h->d.model != SPK_POTTS_PIN) return fail("pin is only defined for app_style potts/pin");
  if (fraction < 0.0 || fraction > 1.0) return fail("Illegal pin command");
  CU(cudaSetDevice(h->d.device));
  if (use_canon(h, true)) return 1;
  const int64_t goff = (int64_t)h->g.g0 * h->plane;
  double t = std::floor(fraction * 4294967296.0);
  const uint32_t thr = t > 4294967295.0 ? 4294967295u : (uint32_t)t;
  pin_kernel<<<148 * 8, 256, 0, h->stream>>>(h->spin, h->nlocal, goff, seed, thr, h->d.nspins + 1);
  h->ghosts_valid = false;
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

// number of pinned sites (spin > Q) of the whole lattice: the reference's recount after every round
// of pin_create (src/app_potts_pin.cpp:279-282: local count + MPI_Allreduce)
static int count_pinned(spk_cuda_handle *h, long long *npin) {
  CU(cudaMemsetAsync(h->escratch, 0, sizeof(unsigned long long), h->stream));
  validate_kernel<<<148 * 8, 256, 0, h->stream>>>(h->spin, h->nlocal, -2147483647 - 1, h->d.nspins, h->escratch);
  h->launches++;
  CU(cudaGetLastError());
  if (h->d.nranks > 1) NC(g_nccl.AllReduce(h->escratch, h->escratch, 1, ncclUint64, ncclSum, h->comm, h->stream));
  unsigned long long c = 0;
  CU(cudaMemcpyAsync(&c, h->escratch, sizeof c, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  *npin = (long long)c;
  return 0;
}

int spk_cuda_pin_create(spk_cuda_handle *h, double fraction, int multi, int nthresh, int32_t park_seed, int64_t *npinned) {
  if (!h) return fail("null handle");
  if (h->d.model != SPK_POTTS_PIN) return fail("pin is only defined for app_style potts/pin");
  if (fraction < 0.0 || fraction > 1.0 || (multi != 0 && multi != 1) || nthresh < 0) return fail("Illegal pin command");
  if (park_seed < 1 || park_seed >= 2147483647) return fail("Park-Miller seed must lie in [1, 2^31-2]");
  if (h->d.nranks > 1 && !h->comm) return fail("pin on several ranks needs spk_cuda_comm_init first");
  const bool commuting = multi == 0 && nthresh == 0;
  if (!commuting && h->d.nranks > 1)
    return fail("pin with multi = 1 or nthresh > 0 depends on the order of the attempts: one GPU only (nranks must be 1)");
  CU(cudaSetDevice(h->d.device));
  if (use_canon(h, true)) return 1;
  const long long ndesired = (long long)(fraction * (double)h->nglobal);  // static_cast<int>(pfraction*nglobal), :246
  const long long goff = (long long)h->g.g0 * h->plane;
  const int Q = h->d.nspins;
  unsigned long long draw = 0;  // draws of the command's stream consumed so far
  long long npin = 0;           // the reference's npin: 0 until the first recount, whatever was pinned before
  unsigned long long *out = nullptr;
  if (!commuting) CU(cudaMalloc(&out, 3 * sizeof(unsigned long long)));
  auto sequential = [&](long long ndraws, long long already, unsigned long long (&res)[3]) -> int {
    switch (h->d.lattice) {
#define SPK_PIN(LATV) case LATV: pin_sequential_kernel<LATV><<<1, 32, 0, h->stream>>>(h->spin, h->g, (uint32_t)park_seed, draw, ndraws, ndesired, already, multi, nthresh, Q, out); break;
      SPK_PIN(SPK_SQ_4N) SPK_PIN(SPK_SQ_8N) SPK_PIN(SPK_TRI) SPK_PIN(SPK_SC_6N) SPK_PIN(SPK_SC_26N)
#undef SPK_PIN
      default: return fail("unknown lattice %d", h->d.lattice);
    }
    h->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(res, out, sizeof res, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
  };
  int rc = 0, stalled = 0;
  if (!multi) {
    while (rc == 0 && npin < ndesired) {  // rounds of ndesired - npin attempts, recount after each (:262-283)
      const long long nattempt = ndesired - npin;
      if (commuting) {
        pin_draws_kernel<<<148 * 8, 256, 0, h->stream>>>(h->spin, h->nlocal, goff, (uint32_t)park_seed, draw, nattempt, h->nglobal, Q + 1);
        h->launches++;
        if (cudaGetLastError() != cudaSuccess) rc = fail("pin: kernel launch failed");
      } else {
        unsigned long long res[3];
        rc = sequential(nattempt, 0, res);
      }
      draw += (unsigned long long)nattempt;
      const long long before = npin;
      if (rc == 0) rc = count_pinned(h, &npin);
      // the reference would loop forever when no site can be pinned any more (nthresh too high)
      stalled = npin == before ? stalled + 1 : 0;
      if (rc == 0 && stalled >= 64) rc = fail("pin: %lld of %lld sites pinned and no attempt of the last 64 rounds succeeded", npin, ndesired);
    }
  } else {
    long long already = 0;
    rc = count_pinned(h, &already);
    if (rc == 0 && ndesired > 0) {
      unsigned long long res[3];
      const long long budget = 64 * (long long)h->nglobal + 1000000;  // draws; far beyond what a satisfiable request needs
      rc = sequential(budget, already, res);
      if (rc == 0 && res[2]) rc = fail("pin: no site left whose neighbourhood is free of pinned sites (%llu pinned, %lld wanted)",
                                       (unsigned long long)already + res[0], ndesired);
      npin = already + (long long)res[0];
    } else {
      npin = already;
    }
  }
  cudaFree(out);
  h->ghosts_valid = false;
  if (rc) return rc;
  if (npinned) *npinned = npin;
  return 0;
}

int spk_cuda_sweep_color(spk_cuda_handle *h, uint32_t sweep, int color, const int32_t *box, int nloop) {
  if (!h) return fail("null handle");
  h->ev.sweep = sweep;
  return sweep_set(h, RNG_PHILOX, color, box, nloop);
}

int spk_cuda_sweep_color_injected(spk_cuda_handle *h, int color, const int32_t *box, const double *u1,
                                  const double *u2) {
  if (!h || !u1 || !u2) return fail("null argument");
  CU(cudaSetDevice(h->d.device));
  const size_t bytes = (size_t)h->nlocal * sizeof(double);
  if (!h->u1) {
    CU(cudaMalloc(&h->u1, bytes));
    CU(cudaMalloc(&h->u2, bytes));
  }
  CU(cudaMemcpyAsync(h->u1, u1, bytes, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->u2, u2, bytes, cudaMemcpyHostToDevice, h->stream));
  int rc = sweep_set(h, RNG_INJECTED, color, box, 1);
  if (rc) return rc;
  CU(cudaStreamSynchronize(h->stream));  // u1/u2 host buffers may be reused by the caller
  return 0;
}

int spk_cuda_set_siteseeds(spk_cuda_handle *h, const int32_t *seeds, int64_t n) {
  if (!h || !seeds) return fail("null argument");
  if (n != h->nlocal) return fail("siteseeds size mismatch");
  CU(cudaSetDevice(h->d.device));
  if (!h->seeds) CU(cudaMalloc(&h->seeds, (size_t)n * sizeof(int32_t)));
  CU(cudaMemcpyAsync(h->seeds, seeds, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int spk_cuda_get_siteseeds(spk_cuda_handle *h, int32_t *seeds, int64_t n) {
  if (!h || !seeds) return fail("null argument");
  if (!h->seeds) return fail("no siteseeds set");
  if (n != h->nlocal) return fail("siteseeds size mismatch");
  CU(cudaSetDevice(h->d.device));
  CU(cudaMemcpyAsync(seeds, h->seeds, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int spk_cuda_sweep_color_strict(spk_cuda_handle *h, int color, const int32_t *box, int nloop) {
  if (!h) return fail("null handle");
  if (!h->seeds) return fail("sweep color/strict needs spk_cuda_set_siteseeds first");
  return sweep_set(h, RNG_STRICT, color, box, nloop);
}

static int halo_exchange_on(spk_cuda_handle *h, cudaStream_t st);
static int check_halo_timeout(spk_cuda_handle *h);

int spk_cuda_halo_exchange(spk_cuda_handle *h) {
  if (!h) return fail("null handle");
  if (h->d.nranks > 1 && use_canon(h, true)) return 1;  // the exchange rewrites the int32 ghost planes
  const int rc = halo_exchange_on(h, h->stream);
  if (rc == 0) h->ghosts_valid = true;
  return rc;
}

// ghost layers <- ring neighbours, enqueued on stream `st` (the compute stream for the plain API,
// the communication stream when spk_cuda_sweeps overlaps the exchange with interior updates)
static int halo_exchange_on(spk_cuda_handle *h, cudaStream_t st) {
  if (h->d.nranks == 1) return 0;  // periodic images are read in place (Geo::lo_wrap / hi_wrap)
  if (!h->comm) return fail("spk_cuda_comm_init has not been called");
  CU(cudaSetDevice(h->d.device));
  const int lower = (h->d.rank + h->d.nranks - 1) % h->d.nranks, upper = (h->d.rank + 1) % h->d.nranks;
  const size_t n = (size_t)h->plane * h->gh;  // gh layers per side
  int32_t *first = h->spin, *last = h->spin + (int64_t)(h->nslab - h->gh) * h->plane;
  int32_t *ghost_lo = h->base, *ghost_hi = h->spin + h->nlocal;
  // the exchange runs on the compute stream's timeline: it must see the previous colour's
  // writes and the next colour kernel must see the ghosts
  NC(g_nccl.GroupStart());
  NC(g_nccl.Send(first, n, ncclInt32, lower, h->comm, st));
  NC(g_nccl.Send(last, n, ncclInt32, upper, h->comm, st));
  NC(g_nccl.Recv(ghost_hi, n, ncclInt32, upper, h->comm, st));
  NC(g_nccl.Recv(ghost_lo, n, ncclInt32, lower, h->comm, st));
  NC(g_nccl.GroupEnd());
  h->launches++;
  return 0;
}

// Multi-GPU sweeps with the ghost exchange overlapped with interior updates: the
// part of a colour set (3-D: first / last plane; 2-D: first / last tile row of the fused sweep)
// that the neighbours' ghosts depend on is launched first; the exchange of its result runs on the
// communication stream while the compute stream updates the interior; the next boundary launch
// waits for the exchange.  Same launches in the same dependency order as the serial schedule,
// hence the same spins.
static int sweeps_overlapped(spk_cuda_handle *h, uint32_t first_sweep, int nsweeps) {
  const bool fused2d = h->dim == 2;
  auto exchange = [&]() -> int {  // comm stream: after the boundary launch, before the next one
    CU(cudaEventRecord(h->ev_compute, h->stream));
    CU(cudaStreamWaitEvent(h->comm_stream, h->ev_compute, 0));
    int rc = halo_exchange_on(h, h->comm_stream);
    if (rc) return rc;
    CU(cudaEventRecord(h->ev_halo, h->comm_stream));
    return 0;
  };
  int rc = exchange();  // ghosts for the first boundary launch
  if (rc) return rc;
  const int32_t lo_box[6] = {0, h->g.gX, 0, h->g.gY, h->g.g0, h->g.g0 + 1};
  const int32_t hi_box[6] = {0, h->g.gX, 0, h->g.gY, h->g.g0 + h->nslab - 1, h->g.g0 + h->nslab};
  const int32_t in_box[6] = {0, h->g.gX, 0, h->g.gY, h->g.g0 + 1, h->g.g0 + h->nslab - 1};
  for (int s = 0; s < nsweeps; s++) {
    h->ev.sweep = first_sweep + (uint32_t)s;
    if (fused2d) {
      CU(cudaStreamWaitEvent(h->stream, h->ev_halo, 0));
      rc = fused_sweep(h, 1);  // first + last tile row -> destination buffer
      if (rc) return rc > 0 ? fail("fused sweep launch failed") : fail("fused sweep not applicable");
      // the exchange must move the DESTINATION buffer's boundary rows: swap happens in part 2, so
      // enqueue the interior first (it swaps the handle's buffers), then exchange the new lattice
      CU(cudaEventRecord(h->ev_compute, h->stream));
      rc = fused_sweep(h, 2);
      if (rc) return fail("fused sweep launch failed");
      CU(cudaStreamWaitEvent(h->comm_stream, h->ev_compute, 0));
      rc = halo_exchange_on(h, h->comm_stream);
      if (rc) return rc;
      CU(cudaEventRecord(h->ev_halo, h->comm_stream));
      continue;
    }
    for (int color = 1; color <= h->nc; color++) {
      CU(cudaStreamWaitEvent(h->stream, h->ev_halo, 0));
      rc = sweep_set(h, RNG_PHILOX, color, lo_box, 1);
      if (rc) return rc;
      if (h->nslab > 1) {
        rc = sweep_set(h, RNG_PHILOX, color, hi_box, 1);
        if (rc) return rc;
      }
      rc = exchange();
      if (rc) return rc;
      if (h->nslab > 2) {
        rc = sweep_set(h, RNG_PHILOX, color, in_box, 1);
        if (rc) return rc;
      }
    }
  }
  CU(cudaStreamWaitEvent(h->stream, h->ev_halo, 0));  // later work on the compute stream sees the ghosts
  return 0;
}

// Multi-GPU sweeps over peer memory (3-D slabs): no exchange step at all.  Per colour the compute
// stream runs   wait(neighbours finished boundary step t-1)  ->  boundary launches, which store
// their updates locally AND into the neighbours' ghost planes  ->  signal(step t) into the
// neighbours' flag words  ->  interior launch (reads no ghost plane, so the neighbours' stores for
// step t land while it runs).  wait(t-1) covers both hazards: the ghosts this rank reads are
// complete (RAW), and the neighbour has finished reading the ghost this rank is about to overwrite
// (WAR; within one step the planes a colour updates and the ghost planes it reads have opposite
// parity, so same-step launches never touch the same words).
static int sweeps_p2p(spk_cuda_handle *h, uint32_t first_sweep, int nsweeps) {
  const bool on16 = shadow_usable(h) && h->mirror16_lo && h->mirror16_hi;
  if (!h->ghosts_valid) {  // spins changed outside spk_cuda_sweeps: one NCCL exchange re-seeds the ghosts
    if (use_canon(h, true)) return 1;
    int rc = halo_exchange_on(h, h->stream);
    if (rc) return rc;
  }
  if (on16) {  // boundary launches store into the neighbours' SHADOW ghost planes
    if (ensure_shadow(h)) return 1;
    h->canon_current = false;
    // the conversion above rewrites this rank's whole shadow, ghost planes included: no neighbour
    // may store its first boundary step into them before it has finished (seen at 8 GPUs as wrong
    // spins: a faster neighbour's mirrored plane overwritten by the slower rank's pack).  One
    // signal / wait step, on every call (a rank cannot know whether its neighbours had to convert).
    if (p2p_read_fence(h)) return 1;
  } else if (use_canon(h, true)) return 1;
  const int32_t lo_box[6] = {0, h->g.gX, 0, h->g.gY, h->g.g0, h->g.g0 + 1};
  const int32_t hi_box[6] = {0, h->g.gX, 0, h->g.gY, h->g.g0 + h->nslab - 1, h->g.g0 + h->nslab};
  const int32_t in_box[6] = {0, h->g.gX, 0, h->g.gY, h->g.g0 + 1, h->g.g0 + h->nslab - 1};
  int rc = 0;
  // sc/26n: a step is a colour PAIR (same z parity, so the same-step argument above holds unchanged)
  const bool pair = !on16 && pair_ok(h);
  auto run = [&](int color, const int32_t *box) -> int {
    if (on16) {
      const int rc = shadow16_m(h, color, box);
      return rc ? fail("16-bit streaming kernel launch failed: %s", cudaGetErrorString(cudaGetLastError())) : 0;
    }
    return pair ? sweep_pair(h, color, box) : sweep_set(h, RNG_PHILOX, color, box, 1);
  };
  for (int s = 0; s < nsweeps && !rc; s++) {
    h->ev.sweep = first_sweep + (uint32_t)s;
    for (int color = 1; color <= h->nc && !rc; color += pair ? 2 : 1) {
      const uint32_t t = ++h->step;
      halo_wait_kernel<<<1, 32, 0, h->stream>>>(h->flags, t - 1);
      h->mirror_active = true;
      rc = run(color, lo_box);
      if (!rc) rc = run(color, hi_box);
      h->mirror_active = false;
      halo_signal_kernel<<<1, 32, 0, h->stream>>>(h->peer_flags[0], h->peer_flags[1], t);
      h->launches += 2;
      if (!rc && h->nslab > 2) rc = run(color, in_box);
    }
  }
  // the neighbours' last stores into this rank's ghost planes must have landed before anything
  // later on the stream reads them (energy, the next call's boundary launches)
  halo_wait_kernel<<<1, 32, 0, h->stream>>>(h->flags, h->step);
  h->launches++;
  if (rc) return rc;
  CU(cudaGetLastError());
  h->ghosts_valid = true;
  return 0;
}

// 2-D slabs (colour-fused sweep, out of place): after each sweep the new boundary rows are pushed
// into the neighbours' ghost rows of the SAME buffer (the buffers swap in lockstep on all ranks)
// by one small kernel that also signals the step; wait(t-1) before the sweep kernel orders both the ghosts it reads and the
// neighbour's last reads of the ghost rows the push overwrites (two sweeps earlier).
static int sweeps_p2p_2d(spk_cuda_handle *h, uint32_t first_sweep, int nsweeps) {
  if (!h->ghosts_valid) {
    int rc = halo_exchange_on(h, h->stream);
    if (rc) return rc;
  }
  const size_t rows = (size_t)h->plane * h->gh;
  for (int s = 0; s < nsweeps; s++) {
    h->ev.sweep = first_sweep + (uint32_t)s;
    const uint32_t t = ++h->step;
    halo_wait_kernel<<<1, 32, 0, h->stream>>>(h->flags, t - 1);
    int rc = fused_sweep(h, 0);  // swaps the buffers: h->base is the new lattice
    if (rc) return rc > 0 ? fail("fused sweep launch failed") : fail("fused sweep not applicable");
    const bool cur0 = h->base == h->alloc0;
    int32_t *lower_buf = (int32_t *)(cur0 ? h->peer_map[0] : h->peer_map_alt[0]);
    int32_t *upper_buf = (int32_t *)(cur0 ? h->peer_map[1] : h->peer_map_alt[1]);
    const long long n16 = (long long)(rows / 4);  // rows are multiples of 8 elements (fused2d_ok)
    const int nblk = (int)std::min<long long>(16, (n16 + 1023) / 1024);
    halo_push_kernel<<<nblk, 1024, 0, h->stream>>>(
        reinterpret_cast<const int4 *>(h->spin), reinterpret_cast<int4 *>(lower_buf + (int64_t)(h->gh + h->nslab) * h->plane),
        reinterpret_cast<const int4 *>(h->spin + (int64_t)(h->nslab - h->gh) * h->plane), reinterpret_cast<int4 *>(upper_buf),
        n16, h->flags + 3, h->peer_flags[0], h->peer_flags[1], t);
    h->launches += 2;
  }
  halo_wait_kernel<<<1, 32, 0, h->stream>>>(h->flags, h->step);
  h->launches++;
  CU(cudaGetLastError());
  h->ghosts_valid = true;
  return 0;
}

int spk_cuda_sweeps(spk_cuda_handle *h, uint32_t first_sweep, int nsweeps) {
  if (!h) return fail("null handle");
  if (nsweeps < 0) return fail("nsweeps must be >= 0");
  CU(cudaSetDevice(h->d.device));
  if (h->mask) {
    // `sweep color mask yes` without sectors: the loop sets are the colour sets, all of whose sites
    // are border sites, so boundary_clear_mask (src/app_lattice.cpp:655) unmasks the whole set
    // before it is swept: same spins as without masks, mask bytes as the reference leaves them
    for (int s = 0; s < nsweeps; s++) {
      h->ev.sweep = first_sweep + (uint32_t)s;
      for (int color = 1; color <= h->nc; color++) {
        int rc = spk_cuda_halo_exchange(h);
        if (!rc) rc = spk_cuda_mask_clear_border(h, nullptr, color);
        if (!rc) rc = sweep_set(h, RNG_PHILOX, color, nullptr, 1);
        if (rc) return rc;
      }
    }
    return 0;
  }
  if (h->d.nranks > 1 && h->comm && h->p2p && nsweeps > 0 && h->tiled) {
    if (h->dim == 3 && h->use_stream) return sweeps_p2p(h, first_sweep, nsweeps);
    if (h->dim == 2 && h->use_fused && h->fused2d_ok && !getenv("SPK_CUDA_OVERLAP_2D"))
      return sweeps_p2p_2d(h, first_sweep, nsweeps);
  }
  if (h->bits_ok && !h->mask && nsweeps > 0) {
    // ising / ising/single on sq/4n, one GPU: multi-spin-coded colour kernels on the bit shadow
    if (ensure_shadow(h)) return 1;
    h->canon_current = false;
    Bits2dArgs ba;
    ba.bits = h->bits; ba.Y = h->g.Y; ba.W = h->g.X / 64; ba.gX = (unsigned long long)h->g.gX; ba.naccept = h->naccept;
    const long long nwords = (long long)ba.W * ba.Y;
    const int grid = (int)std::min<long long>(148 * 16, (nwords + 127) / 128);
    for (int s = 0; s < nsweeps; s++) {
      h->ev.sweep = first_sweep + (uint32_t)s;
      ba.ev = h->ev;
      for (int c0 = 0; c0 < 2; c0++) {
        ba.c0 = c0;
        if (h->d.model == SPK_ISING) sweep_bits2d_kernel<SPK_ISING><<<grid, 128, 0, h->stream>>>(ba);
        else sweep_bits2d_kernel<SPK_ISING_SINGLE><<<grid, 128, 0, h->stream>>>(ba);
      }
      h->launches += 2;
      h->nattempt += h->nlocal;
    }
    CU(cudaGetLastError());
    return 0;
  }
  if (h->d.nranks == 1 && shadow_usable(h) && nsweeps > 0) {
    // sc/26n potts / potts/pin on one GPU: every colour launch on the 16-bit shadow (half the DRAM
    // bytes); the int32 lattice is brought up to date by the next call that needs it
    if (ensure_shadow(h)) return 1;
    h->canon_current = false;
    for (int s = 0; s < nsweeps; s++) {
      h->ev.sweep = first_sweep + (uint32_t)s;
      for (int color = 1; color <= h->nc; color++) {
        const int rc = shadow16_m(h, color, nullptr);
        if (rc) return fail("16-bit streaming kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
      }
    }
    return 0;
  }
  const bool fused2d = h->tiled && h->use_fused && h->dim == 2 && h->fused2d_ok;
  // 3-D: 8 (2) exchanges per sweep, worth hiding.  2-D: one exchange of 4 rows per fused sweep; splitting
  // the single launch into boundary + interior launches costs more than the exchange it hides
  // (measured: 502 vs 538 Gsites/s at 2 GPUs), so it stays serial unless SPK_CUDA_OVERLAP_2D=1.
  const bool overlap = h->dim == 3 || (fused2d && getenv("SPK_CUDA_OVERLAP_2D"));
  if (h->d.nranks > 1 && h->comm && nsweeps > 0 && overlap && !getenv("SPK_CUDA_NO_OVERLAP"))
    return sweeps_overlapped(h, first_sweep, nsweeps);
  for (int s = 0; s < nsweeps; s++) {
    h->ev.sweep = first_sweep + (uint32_t)s;
    if (fused2d) {
      // 2-D: one launch per sweep; ghost rows (one per fused colour step) are refreshed once
      int rc = spk_cuda_halo_exchange(h);
      if (rc) return rc;
      rc = fused_sweep(h, 0);
      if (rc == 0) continue;
      if (rc > 0) return fail("fused sweep launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (h->d.nranks == 1 && pair_ok(h)) {  // sc/26n: two colours per launch (nothing to exchange on one GPU)
      h->ghosts_valid = false;
      for (int color = 1; color <= h->nc; color += 2) {
        int rc = sweep_pair(h, color, nullptr);
        if (rc) return rc;
      }
      continue;
    }
    for (int color = 1; color <= h->nc; color++) {
      int rc = spk_cuda_halo_exchange(h);
      if (rc) return rc;
      rc = sweep_set(h, RNG_PHILOX, color, nullptr, 1);
      if (rc) return rc;
    }
  }
  return 0;
}

int spk_cuda_counters(spk_cuda_handle *h, int64_t *naccept, int64_t *nattempt) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->d.device));
  unsigned long long acc = 0;
  CU(cudaMemcpyAsync(&acc, h->naccept, sizeof acc, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (naccept) *naccept = (int64_t)acc;
  if (nattempt) *nattempt = h->nattempt;
  return check_halo_timeout(h);
}

int spk_cuda_reset_counters(spk_cuda_handle *h) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->d.device));
  CU(cudaMemsetAsync(h->naccept, 0, sizeof(unsigned long long), h->stream));
  h->nattempt_base += h->nattempt;
  h->nattempt = 0;
  return 0;
}

int spk_cuda_host_register(void *ptr, size_t bytes) {
  if (!ptr || !bytes) return fail("null argument");
  cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterDefault);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail("cudaHostRegister of %zu bytes: %s", bytes, cudaGetErrorString(e));
  }
  return 0;
}

int spk_cuda_host_unregister(void *ptr) {
  if (!ptr) return 0;
  cudaHostUnregister(ptr);
  cudaGetLastError();
  return 0;
}

int spk_cuda_mask_enable(spk_cuda_handle *h, int on) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->d.device));
  if (!on) {
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(h->mask_base);
    h->mask_base = h->mask = nullptr;
    return 0;
  }
  const size_t n = (size_t)(h->nlocal + 2 * (int64_t)h->gh * h->plane);
  if (!h->mask_base) CU(cudaMalloc(&h->mask_base, n));
  CU(cudaMemsetAsync(h->mask_base, 0, n, h->stream));  // AppLattice::init: mask[i] = 0 (src/app_lattice.cpp:264-267)
  h->mask = h->mask_base + (int64_t)h->gh * h->plane;
  return 0;
}

int spk_cuda_mask_clear_border(spk_cuda_handle *h, const int32_t *box, int color) {
  if (!h) return fail("null handle");
  if (!h->mask) return fail("masking is not enabled (spk_cuda_mask_enable)");
  if (color < 0 || color > h->nc) return fail("colour %d out of range 0..%d", color, h->nc);
  CU(cudaSetDevice(h->d.device));
  Box b;
  if (!local_box(h, box, b)) return 0;
  const long long ncell = (long long)(b.ihi - b.ilo) * (b.jhi - b.jlo) * (b.khi - b.klo);
  const int grid = (int)std::min<long long>(148 * 8, (ncell + 255) / 256);
  const int multi = h->d.nranks > 1;
#define SPK_MASK_CASE(LATV)                                                                              \
  case LATV:                                                                                             \
    if (color) mask_color_clear_kernel<LATV><<<grid, 256, 0, h->stream>>>(h->mask, h->g, b, color - 1);  \
    else mask_border_kernel<LATV><<<grid, 256, 0, h->stream>>>(h->mask, h->g, b, multi);                 \
    break;
  switch (h->d.lattice) {
    SPK_MASK_CASE(SPK_SQ_4N) SPK_MASK_CASE(SPK_SQ_8N) SPK_MASK_CASE(SPK_TRI) SPK_MASK_CASE(SPK_SC_6N) SPK_MASK_CASE(SPK_SC_26N)
    default: return fail("unknown lattice %d", h->d.lattice);
  }
#undef SPK_MASK_CASE
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

int spk_cuda_download_mask(spk_cuda_handle *h, unsigned char *host, int64_t n) {
  if (!h || !host) return fail("null argument");
  if (!h->mask) return fail("masking is not enabled (spk_cuda_mask_enable)");
  if (n != h->nlocal) return fail("download of %lld mask bytes, slab owns %lld sites", (long long)n, (long long)h->nlocal);
  CU(cudaSetDevice(h->d.device));
  CU(cudaMemcpyAsync(host, h->mask, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int spk_cuda_upload_mask(spk_cuda_handle *h, const unsigned char *host, int64_t n) {
  if (!h || !host) return fail("null argument");
  if (!h->mask) return fail("masking is not enabled (spk_cuda_mask_enable)");
  if (n != h->nlocal) return fail("upload of %lld mask bytes, slab owns %lld sites", (long long)n, (long long)h->nlocal);
  CU(cudaSetDevice(h->d.device));
  CU(cudaMemcpyAsync(h->mask, host, (size_t)n, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 0;
}

int spk_cuda_digest(spk_cuda_handle *h, uint64_t *digest) {
  if (!h || !digest) return fail("null argument");
  CU(cudaSetDevice(h->d.device));
  if (use_canon(h, false)) return 1;
  CU(cudaMemsetAsync(h->escratch, 0, sizeof(unsigned long long), h->stream));
  digest_kernel<<<148 * 8, 256, 0, h->stream>>>(h->spin, h->nlocal, (long long)h->g.g0 * h->plane, h->escratch);
  h->launches++;
  unsigned long long d = 0;
  CU(cudaMemcpyAsync(&d, h->escratch, sizeof d, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  *digest = d;
  return 0;
}

int spk_cuda_sweep_path(spk_cuda_handle *h, char *name, int capacity, int *launches_per_sweep, int *bytes_per_spin) {
  if (!h || !name || capacity < 1) return fail("null argument");
  const char *k;
  int per = h->nc, width = 4;
  const bool multi = h->d.nranks > 1;
  if (h->mask || !h->tiled) k = "sweep_generic_kernel";
  else if (h->bits_ok) { k = "sweep_bits2d_kernel"; per = 2; width = 0; }
  else if (shadow_usable(h) && (!multi || (h->p2p && h->mirror16_lo))) { k = "sweep_stream16_kernel"; width = 2; }
  else if (h->dim == 2 && h->use_fused && h->fused2d_ok) { k = "sweep_tile2d_kernel"; per = 1; }
  else if (h->dim == 3 && h->use_stream && h->stream_ok) k = "sweep_stream_kernel";
  else k = "sweep_strip_kernel / sweep_generic_kernel";
  snprintf(name, (size_t)capacity, "%s", k);
  if (launches_per_sweep) *launches_per_sweep = per;
  if (bytes_per_spin) *bytes_per_spin = width;  // 0 = one bit
  return 0;
}

int spk_cuda_neigh_variant(spk_cuda_handle *h, int *variant, double *mean_distinct) {
  if (!h || !variant) return fail("null argument");
  *variant = h->tmap8_ok && h->tiled && h->use_stream ? (h->neigh_adaptive ? 1 : 0) : -1;
  if (mean_distinct) *mean_distinct = h->kstat_mean;
  return 0;
}

// Peer-memory halo: a call that READS the ghost layers outside spk_cuda_sweeps (energy, clusters,
// per-site energies) must not return while a faster neighbour could already be storing the next
// sweep's boundary values into them.  One more signal / wait step on the flag words closes the
// window: when it completes, both neighbours have finished their own reads.  Like every multi-rank
// entry point these calls are collective (every rank makes them, in the same order).
static int p2p_read_fence(spk_cuda_handle *h) {
  if (h->d.nranks == 1 || !h->p2p || !h->flags) return 0;
  const uint32_t t = ++h->step;
  halo_signal_kernel<<<1, 32, 0, h->stream>>>(h->peer_flags[0], h->peer_flags[1], t);
  halo_wait_kernel<<<1, 32, 0, h->stream>>>(h->flags, t);
  h->launches += 2;
  CU(cudaGetLastError());
  return 0;
}

int spk_cuda_energy(spk_cuda_handle *h, double *energy) {
  if (!h || !energy) return fail("null argument");
  CU(cudaSetDevice(h->d.device));
  if (use_canon(h, false)) return 1;
  int rc = spk_cuda_halo_exchange(h);  // DiagEnergy::compute calls comm->all() first (src/diag_energy.cpp:56)
  if (rc) return rc;
  CU(cudaMemsetAsync(h->escratch, 0, sizeof(unsigned long long), h->stream));
  rc = launch_energy(h->d.lattice, h->stream, h->spin, h->g, h->escratch, &h->launches);
  if (rc) return fail("energy kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  if (p2p_read_fence(h)) return 1;
  unsigned long long e = 0;
  CU(cudaMemcpyAsync(&e, h->escratch, sizeof e, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  *energy = (double)e;
  return 0;
}

// Map the ring neighbours' lattice allocations and flag words (CUDA IPC; one process per GPU) for
// the peer-memory halo.  The handles travel through one ncclAllGather; a second one makes every
// rank agree on whether the mapping worked everywhere (otherwise all keep the NCCL exchange).
static int setup_p2p(spk_cuda_handle *h) {
  struct Ipc { cudaIpcMemHandle_t base, alt, flags, shadow; int ok, ok16; int pad[2]; };
  h->p2p = false;
  const char *np = getenv("SPK_CUDA_NO_P2P");
  const bool two = h->dim == 2;  // 2-D: both buffers of the out-of-place fused sweep
  const int want = !(np && np[0] == '1') && h->nslab >= 2 &&
                   (two ? (h->fused2d_ok && h->base_alt && h->nslab >= h->gh) : h->stream_ok);
  const int n = h->d.nranks, lower = (h->d.rank + n - 1) % n, upper = (h->d.rank + 1) % n;
  int32_t *first = h->alloc0, *second = h->base == h->alloc0 ? h->base_alt : h->base;
  Ipc mine;
  memset(&mine, 0, sizeof mine);
  if (want) {
    if (!h->flags) {
      CU(cudaMalloc(&h->flags, 4 * sizeof(uint32_t)));
      CU(cudaMemset(h->flags, 0, 4 * sizeof(uint32_t)));
    }
    mine.ok = cudaIpcGetMemHandle(&mine.base, first) == cudaSuccess &&
              (!two || cudaIpcGetMemHandle(&mine.alt, second) == cudaSuccess) &&
              cudaIpcGetMemHandle(&mine.flags, h->flags) == cudaSuccess;
    // 16-bit shadow (3-D): its ghost planes are written by the neighbours too; all ranks or none
    mine.ok16 = mine.ok && h->shadow_ok && cudaIpcGetMemHandle(&mine.shadow, h->shadow_base) == cudaSuccess;
    cudaGetLastError();
  }
  Ipc *dev = nullptr;
  CU(cudaMalloc(&dev, (size_t)(n + 1) * sizeof(Ipc)));
  std::vector<Ipc> all(n);
  auto gather = [&]() -> int {
    CU(cudaMemcpyAsync(dev + n, &mine, sizeof mine, cudaMemcpyHostToDevice, h->stream));
    NC(g_nccl.AllGather(dev + n, dev, sizeof(Ipc), ncclChar, h->comm, h->stream));
    CU(cudaMemcpyAsync(all.data(), dev, (size_t)n * sizeof(Ipc), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
  };
  int rc = gather();
  bool every = rc == 0, every16 = rc == 0;
  for (int r = 0; r < n && every; r++) every = all[r].ok != 0;
  for (int r = 0; r < n && every16; r++) every16 = all[r].ok16 != 0;
  if (every) {  // open the neighbours' allocations (one mapping when both neighbours are the same rank)
    const int peers[2] = {lower, upper};
    for (int side = 0; side < 2 && mine.ok; side++) {
      if (side == 1 && upper == lower) {
        h->peer_map[1] = h->peer_map[0];
        h->peer_map_alt[1] = h->peer_map_alt[0];
        h->peer_flag_map[1] = h->peer_flag_map[0];
        h->peer_map16[1] = h->peer_map16[0];
        continue;
      }
      const Ipc &p = all[peers[side]];
      if (every16 && cudaIpcOpenMemHandle(&h->peer_map16[side], p.shadow, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        mine.ok16 = 0;
      }
      if (cudaIpcOpenMemHandle(&h->peer_map[side], p.base, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
          (two && cudaIpcOpenMemHandle(&h->peer_map_alt[side], p.alt, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ||
          cudaIpcOpenMemHandle(&h->peer_flag_map[side], p.flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        mine.ok = 0;
      }
    }
    rc = gather();  // agreement round
    every = rc == 0;
    for (int r = 0; r < n && every; r++) every = all[r].ok != 0;
    for (int r = 0; r < n && every16; r++) every16 = all[r].ok16 != 0;
  }
  cudaFree(dev);
  if (rc) return rc;
  if (every) {
    // lower neighbour's UPPER ghost layer mirrors my first layer(s); upper neighbour's LOWER ghost layer my last
    h->mirror_lo = (int32_t *)h->peer_map[0] + (int64_t)(h->gh + h->nslab) * h->plane;
    h->mirror_hi = (int32_t *)h->peer_map[1];
    h->peer_flags[0] = (uint32_t *)h->peer_flag_map[0] + 1;  // I am the lower neighbour's upper neighbour
    h->peer_flags[1] = (uint32_t *)h->peer_flag_map[1] + 0;
    h->p2p = true;
    if (every16) {  // the shadow's ghost planes: plane 0 (lower) and plane nslab+1 (upper) of its allocation
      h->mirror16_lo = (uint16_t *)h->peer_map16[0] + (int64_t)(1 + h->nslab) * h->plane;
      h->mirror16_hi = (uint16_t *)h->peer_map16[1];
    }
  }
  // without peer memory (or without every rank's shadow) a multi-rank run keeps the int32 lattice
  if (!(every && every16)) h->shadow_ok = false;
  return 0;
}

int spk_cuda_comm_id(char id[SPK_CUDA_COMM_ID_BYTES]) {
  static_assert(sizeof(ncclUniqueId) <= SPK_CUDA_COMM_ID_BYTES, "ncclUniqueId larger than the ABI buffer");
  if (!id) return fail("null argument");
  if (nccl_bind()) return 1;
  ncclUniqueId u;
  NC(g_nccl.GetUniqueId(&u));
  memset(id, 0, SPK_CUDA_COMM_ID_BYTES);
  memcpy(id, &u, sizeof u);
  return 0;
}

int spk_cuda_comm_init(spk_cuda_handle *h, const char id[SPK_CUDA_COMM_ID_BYTES]) {
  if (!h || !id) return fail("null argument");
  if (h->d.nranks == 1) return 0;
  if (h->comm) return fail("communicator already initialised");
  CU(cudaSetDevice(h->d.device));
  if (nccl_bind()) return 1;
  ncclUniqueId u;
  memcpy(&u, id, sizeof u);
  NC(g_nccl.CommInitRank(&h->comm, h->d.nranks, u, h->d.rank));
  return setup_p2p(h);
}

// the peer-memory halo's device-side wait gave up: a ring neighbour never signalled its step
static int check_halo_timeout(spk_cuda_handle *h) {
  if (!h->p2p) return 0;
  uint32_t timed_out = 0;
  CU(cudaMemcpyAsync(&timed_out, h->flags + 2, sizeof timed_out, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (timed_out) return fail("peer-memory halo: a neighbour rank did not signal its boundary step within 20 s");
  return 0;
}

// union-find over the current lattice: labels flattened to roots, volumes per root, ncluster and
// the radius sum on the device; returns ncluster and the radius sum (of the WHOLE lattice: with
// several slabs the clusters are merged across the slab boundaries and the sums reduced over ranks)
static int cluster_across_slabs(spk_cuda_handle *h, unsigned long long *ncluster, double *rsum);

static int cluster_analyse(spk_cuda_handle *h, unsigned long long *ncluster, double *rsum) {
  if (h->nlocal >= (int64_t)1 << 32) return fail("cluster analysis needs fewer than 2^32 sites per GPU");
  CU(cudaSetDevice(h->d.device));
  if (use_canon(h, false)) return 1;
  const long long n = h->nlocal;
  const bool multi = h->d.nranks > 1;
  if (multi) {
    if (!h->comm) return fail("spk_cuda_comm_init has not been called");
    const int rc = spk_cuda_halo_exchange(h);  // DiagCluster: comm->all() first (src/diag_cluster.cpp:262)
    if (rc) return rc;
  }
  if (!h->cl_label) {
    CU(cudaMalloc(&h->cl_label, (size_t)n * sizeof(uint32_t)));
    CU(cudaMalloc(&h->cl_size, (size_t)n * sizeof(uint32_t)));
    CU(cudaMalloc(&h->cl_blocksum, (size_t)((n + CL_TILE - 1) / CL_TILE) * sizeof(uint32_t)));
    CU(cudaMalloc(&h->cl_count, 2 * sizeof(unsigned long long)));
  }
  const int grid = 148 * 16;
  cluster_init_kernel<<<grid, 256, 0, h->stream>>>(h->cl_label, h->cl_size, n);
  const int X = h->g.X, Y = h->g.Y, Z = h->dim == 3 ? h->g.Z : 1;  // local extents (= global on one GPU)
  switch (h->d.lattice) {
    case SPK_SQ_4N: launch_cluster_link<SPK_SQ_4N>(h->stream, h->spin, h->cl_label, X, Y, Z, multi); break;
    case SPK_SQ_8N: launch_cluster_link<SPK_SQ_8N>(h->stream, h->spin, h->cl_label, X, Y, Z, multi); break;
    case SPK_TRI: launch_cluster_link<SPK_TRI>(h->stream, h->spin, h->cl_label, X, Y, Z, multi); break;
    case SPK_SC_6N: launch_cluster_link<SPK_SC_6N>(h->stream, h->spin, h->cl_label, X, Y, Z, multi); break;
    case SPK_SC_26N: launch_cluster_link<SPK_SC_26N>(h->stream, h->spin, h->cl_label, X, Y, Z, multi); break;
    default: return fail("unknown lattice %d", h->d.lattice);
  }
  cluster_flatten_kernel<<<grid, 256, 0, h->stream>>>(h->cl_label, h->cl_size, n);
  h->launches += 3;
  if (multi) return cluster_across_slabs(h, ncluster, rsum);
  CU(cudaMemsetAsync(h->cl_count, 0, 2 * sizeof(unsigned long long), h->stream));
  cluster_stats_kernel<<<grid, 256, 0, h->stream>>>(h->cl_label, h->cl_size, n, 1.0 / h->dim, h->cl_count,
                                                    reinterpret_cast<double *>(h->cl_count + 1));
  h->launches++;
  CU(cudaGetLastError());
  unsigned long long out[2];
  CU(cudaMemcpyAsync(out, h->cl_count, sizeof out, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  *ncluster = out[0];
  memcpy(rsum, &out[1], sizeof(double));
  return 0;
}

// Merge of the per-slab clusters (cluster.cuh: "clusters across slabs"): rounds of boundary-label
// exchange with the ring neighbours until no global label changes anywhere, then the volumes of
// the roots represented elsewhere are gathered and absorbed by their representatives, and the two
// sums are reduced over the ranks (the analogue of DiagCluster's merge on the root process,
// src/diag_cluster.cpp:372-396, without funnelling anything through one rank).
static int cluster_across_slabs(spk_cuda_handle *h, unsigned long long *ncluster, double *rsum) {
  const long long n = h->nlocal, nl = h->plane;  // nl = sites of one layer of the slab dimension
  const unsigned long long base = (unsigned long long)h->g.g0 * (unsigned long long)h->plane;
  const int nr = h->d.nranks, lower = (h->d.rank + nr - 1) % nr, upper = (h->d.rank + 1) % nr;
  const int cap = (int)std::min<long long>(2 * nl, 1 << 30);
  if (!h->cl_glabel) {
    CU(cudaMalloc(&h->cl_glabel, (size_t)n * sizeof(unsigned long long)));
    CU(cudaMalloc(&h->cl_gsize, (size_t)n * sizeof(unsigned long long)));
    CU(cudaMalloc(&h->cl_layer, (size_t)4 * nl * sizeof(unsigned long long)));
    CU(cudaMalloc(&h->cl_pairs, (size_t)2 * cap * sizeof(unsigned long long)));
    CU(cudaMalloc(&h->cl_flags, 4 * sizeof(int)));
  }
  unsigned long long *out_lo = h->cl_layer, *out_hi = out_lo + nl, *ghost_lo = out_hi + nl, *ghost_hi = ghost_lo + nl;
  const int grid = 148 * 16, lgrid = (int)std::min<long long>(148 * 8, (2 * nl + 255) / 256);
  cluster_global_init_kernel<<<grid, 256, 0, h->stream>>>(h->cl_label, h->cl_size, h->cl_glabel, h->cl_gsize, n, base);
  h->launches++;
  const int X = h->g.X, Y = h->g.Y, Z = h->dim == 3 ? h->g.Z : 1;
  for (int round = 0;; round++) {
    if (round > 4 * nr + 64) return fail("cluster merge across slabs did not converge");
    CU(cudaMemsetAsync(h->cl_flags, 0, sizeof(int), h->stream));
    cluster_pack_kernel<<<lgrid, 256, 0, h->stream>>>(h->cl_label, h->cl_glabel, n, nl, out_lo, out_hi);
    NC(g_nccl.GroupStart());
    NC(g_nccl.Send(out_lo, (size_t)nl, ncclUint64, lower, h->comm, h->stream));
    NC(g_nccl.Send(out_hi, (size_t)nl, ncclUint64, upper, h->comm, h->stream));
    NC(g_nccl.Recv(ghost_hi, (size_t)nl, ncclUint64, upper, h->comm, h->stream));
    NC(g_nccl.Recv(ghost_lo, (size_t)nl, ncclUint64, lower, h->comm, h->stream));
    NC(g_nccl.GroupEnd());
    switch (h->d.lattice) {
#define SPK_CB(LATV) case LATV: cluster_boundary_kernel<LATV><<<lgrid, 256, 0, h->stream>>>(h->spin, h->cl_label, h->cl_glabel, ghost_lo, ghost_hi, X, Y, Z, h->cl_flags); break;
      SPK_CB(SPK_SQ_4N) SPK_CB(SPK_SQ_8N) SPK_CB(SPK_TRI) SPK_CB(SPK_SC_6N) SPK_CB(SPK_SC_26N)
#undef SPK_CB
      default: return fail("unknown lattice %d", h->d.lattice);
    }
    NC(g_nccl.AllReduce(h->cl_flags, h->cl_flags, 1, ncclInt32, ncclMax, h->comm, h->stream));
    h->launches += 4;
    int changed = 0;
    CU(cudaMemcpyAsync(&changed, h->cl_flags, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (!changed) break;
  }
  // volumes of the roots represented elsewhere -> every rank -> absorbed by the representatives
  CU(cudaMemsetAsync(h->cl_flags + 1, 0, sizeof(int), h->stream));
  cluster_foreign_kernel<<<grid, 256, 0, h->stream>>>(h->cl_label, h->cl_glabel, h->cl_gsize, n, base, h->cl_pairs,
                                                      h->cl_flags + 1, cap);
  CU(cudaMemcpyAsync(h->cl_flags + 2, h->cl_flags + 1, sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
  NC(g_nccl.AllReduce(h->cl_flags + 2, h->cl_flags + 2, 1, ncclInt32, ncclMax, h->comm, h->stream));
  int counts[2] = {0, 0};  // mine, maximum over the ranks
  CU(cudaMemcpyAsync(counts, h->cl_flags + 1, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (counts[0] > cap) return fail("cluster merge: %d boundary roots exceed the buffer of %d", counts[0], cap);
  if (counts[1] > 0) {
    const size_t per = (size_t)2 * counts[1];  // padded to the largest list; unused slots carry label ~0
    unsigned long long *all = nullptr;
    CU(cudaMalloc(&all, per * nr * sizeof(unsigned long long)));
    if (counts[0] < counts[1])
      CU(cudaMemsetAsync(h->cl_pairs + 2 * (size_t)counts[0], 0xFF, (size_t)2 * (counts[1] - counts[0]) * sizeof(unsigned long long), h->stream));
    ncclResult_t r = g_nccl.AllGather(h->cl_pairs, all, per, ncclUint64, h->comm, h->stream);
    if (r == ncclSuccess) {
      const long long np = (long long)counts[1] * nr;
      cluster_absorb_kernel<<<(int)std::min<long long>(148 * 8, (np + 255) / 256), 256, 0, h->stream>>>(all, np, base, n, h->cl_gsize);
      cudaStreamSynchronize(h->stream);
    }
    cudaFree(all);
    if (r != ncclSuccess) return fail("ncclAllGather: %s", g_nccl.GetErrorString(r));
    h->launches += 2;
  }
  CU(cudaMemsetAsync(h->cl_count, 0, 2 * sizeof(unsigned long long), h->stream));
  cluster_global_stats_kernel<<<grid, 256, 0, h->stream>>>(h->cl_label, h->cl_glabel, h->cl_gsize, n, base, 1.0 / h->dim,
                                                           h->cl_count, reinterpret_cast<double *>(h->cl_count + 1));
  NC(g_nccl.AllReduce(h->cl_count, h->cl_count, 1, ncclUint64, ncclSum, h->comm, h->stream));
  NC(g_nccl.AllReduce(h->cl_count + 1, h->cl_count + 1, 1, ncclDouble, ncclSum, h->comm, h->stream));
  // (the all-reduces order every rank's ghost reads before anybody's next sweep: no extra fence)
  h->launches += 4;
  CU(cudaGetLastError());
  unsigned long long out[2];
  CU(cudaMemcpyAsync(out, h->cl_count, sizeof out, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  *ncluster = out[0];
  memcpy(rsum, &out[1], sizeof(double));
  return 0;
}

int spk_cuda_cluster_stats(spk_cuda_handle *h, int64_t *ncluster, double *mean_volume, double *mean_radius) {
  if (!h) return fail("null handle");
  unsigned long long nc = 0;
  double rsum = 0.0;
  const int rc = cluster_analyse(h, &nc, &rsum);
  if (rc) return rc;
  if (ncluster) *ncluster = (int64_t)nc;
  // volsum = every site is in exactly one cluster (src/diag_cluster.cpp:595-605)
  if (mean_volume) *mean_volume = (double)h->nglobal / (double)nc;
  if (mean_radius) *mean_radius = rsum / (double)nc;
  return 0;
}

int spk_cuda_cluster_ids(spk_cuda_handle *h, int32_t *ids, int64_t *volumes, int64_t nvolumes) {
  if (!h) return fail("null handle");
  if (h->d.nranks != 1) return fail("per-site cluster ids are a single-GPU facility (nranks must be 1); the statistics work across slabs");
  unsigned long long nc = 0;
  double rsum = 0.0;
  int rc = cluster_analyse(h, &nc, &rsum);
  if (rc) return rc;
  if (volumes && nvolumes < (int64_t)nc) return fail("volumes has room for %lld clusters, %llu found", (long long)nvolumes, nc);
  const long long n = h->nlocal;
  const int nb = (int)((n + CL_TILE - 1) / CL_TILE);
  // ordered volume list: a buffer kept with the handle and grown when more clusters turn up
  if ((size_t)nc > h->cl_dvol_cap) {
    cudaFree(h->cl_dvol);
    h->cl_dvol = nullptr;
    h->cl_dvol_cap = 0;
    const size_t cap = (size_t)nc + (size_t)nc / 4 + 1024;
    CU(cudaMalloc(&h->cl_dvol, cap * sizeof(long long)));
    h->cl_dvol_cap = cap;
  }
  long long *dvol = h->cl_dvol;
  cluster_scan_blocksums<<<nb, 256, 0, h->stream>>>(h->cl_label, n, h->cl_blocksum);
  cluster_scan_offsets<<<1, 1024, 0, h->stream>>>(h->cl_blocksum, nb);
  cluster_scan_apply<<<nb, 256, 0, h->stream>>>(h->cl_label, h->cl_size, n, h->cl_blocksum, dvol);
  h->launches += 3;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && volumes)
    e = cudaMemcpyAsync(volumes, dvol, (size_t)nc * sizeof(long long), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess && ids) {
    // the parent array is not needed any more once every site knows its root's id: reuse it
    cluster_ids_kernel<<<148 * 16, 256, 0, h->stream>>>(h->cl_label, h->cl_size, n, reinterpret_cast<int32_t *>(h->cl_label));
    h->launches++;
    e = cudaGetLastError();
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(ids, h->cl_label, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) return fail("cluster ids: %s", cudaGetErrorString(e));
  return 0;
}

int spk_cuda_halo_mode(const spk_cuda_handle *h) {
  if (!h || h->d.nranks == 1) return 0;
  if (!h->p2p || !h->tiled) return 1;
  return (h->dim == 3 ? h->use_stream : (h->use_fused && h->fused2d_ok)) ? 2 : 1;
}

int spk_cuda_sync(spk_cuda_handle *h) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->d.device));
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaStreamSynchronize(h->comm_stream));
  return check_halo_timeout(h);
}

int spk_cuda_timer_start(spk_cuda_handle *h) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->d.device));
  h->nlaps = 0;
  CU(cudaEventRecord(h->t0, h->stream));
  return 0;
}

int spk_cuda_timer_lap(spk_cuda_handle *h) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->d.device));
  if (h->nlaps == h->laps.size()) {
    cudaEvent_t e;
    CU(cudaEventCreate(&e));
    h->laps.push_back(e);
  }
  CU(cudaEventRecord(h->laps[h->nlaps++], h->stream));
  return 0;
}

int spk_cuda_timer_laps(spk_cuda_handle *h, float *milliseconds, int capacity, int *nlaps) {
  if (!h || !nlaps) return fail("null argument");
  CU(cudaSetDevice(h->d.device));
  *nlaps = (int)h->nlaps;
  cudaEvent_t prev = h->t0;
  for (size_t i = 0; i < h->nlaps && (int)i < capacity && milliseconds; i++) {
    CU(cudaEventSynchronize(h->laps[i]));
    CU(cudaEventElapsedTime(&milliseconds[i], prev, h->laps[i]));
    prev = h->laps[i];
  }
  return 0;
}

int spk_cuda_timer_stop(spk_cuda_handle *h, float *ms) {
  if (!h || !ms) return fail("null argument");
  CU(cudaSetDevice(h->d.device));
  CU(cudaEventRecord(h->t1, h->stream));
  CU(cudaEventSynchronize(h->t1));
  CU(cudaEventElapsedTime(ms, h->t0, h->t1));
  return 0;
}

int spk_cuda_launch_count(spk_cuda_handle *h, int64_t *launches) {
  if (!h || !launches) return fail("null argument");
  *launches = h->launches;
  return 0;
}

int spk_cuda_device_spins(spk_cuda_handle *h, void **dptr) {
  if (!h || !dptr) return fail("null argument");
  CU(cudaSetDevice(h->d.device));
  if (use_canon(h, true)) return 1;
  *dptr = h->spin;
  h->ghosts_valid = false;  // the caller may write through the pointer: re-seed the ghosts at the next sweep
  return 0;
}

}  // extern "C"
