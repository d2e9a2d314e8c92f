// Per-application launch logic (templates on lattice x model); included by model_*.cu only.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "handle.h"
#include "sweep_stream.cuh"
#include "sweep_stream16.cuh"
#include "sweep_tile2d.cuh"
#include "sweep_tiled.cuh"

namespace spk {

// Test switches (read at every launch so that tests can change them): small lattices otherwise take
// neither the multi-step items (several class planes streamed through the ring by one item) nor the
// loop over several items per CTA — the paths every BASELINE-size run lives on.
//   SPK_CUDA_MAX_PLANE_CHUNKS=n  at most n chunks of class planes per tile (1: one item streams all planes)
//   SPK_CUDA_MAX_CTAS=n          at most n CTAs per persistent launch
inline int env_cap(const char *name, int value) {
  const char *e = getenv(name);
  const int v = e ? atoi(e) : 0;
  return v > 0 && v < value ? v : value;
}


// sites of colour c0 inside the local box (nattempt += nselect, src/app_lattice.cpp:677)
template <int LAT> int64_t class_count(const spk_cuda_handle *h, const Box &b, int c0) {
  auto par = [](int lo, int hi, int p) -> int64_t {  // # of x in [lo,hi) with x%2 == p
    int first = lo + ((lo ^ p) & 1);
    return first < hi ? (hi - first + 1) / 2 : 0;
  };
  const int64_t ni = b.ihi - b.ilo, nj = b.jhi - b.jlo, nk = b.khi - b.klo;
  if (LAT == SPK_SQ_4N || LAT == SPK_SC_6N) {
    int64_t n = 0;
    for (int k = b.klo; k < b.khi; k++)
      for (int par_j = 0; par_j < 2; par_j++) {
        int shift = LAT == SPK_SQ_4N ? h->g.g0 : k + h->g.g0;
        n += par(b.jlo, b.jhi, par_j) * par(b.ilo, b.ihi, (c0 + par_j + shift) & 1);
      }
    return n;
  }
  if (LAT == SPK_SQ_8N) return par(b.ilo, b.ihi, c0 & 1) * par(b.jlo, b.jhi, (c0 >> 1) & 1);
  if (LAT == SPK_TRI) return par(b.ilo, b.ihi, (c0 >> 1) & 1) * nj;
  (void)ni; (void)nk;
  return par(b.ilo, b.ihi, c0 & 1) * par(b.jlo, b.jhi, (c0 >> 1) & 1) * par(b.klo, b.khi, (c0 >> 2) & 1);
}

template <int LAT, int MODEL, int RNG>
int launch_generic(spk_cuda_handle *h, const SweepArgs &a) {
  const Box &b = a.b;
  const int half_x = (b.ihi - b.ilo + 1) / 2 + 1;
  int rows = b.jhi - b.jlo, planes = b.khi - b.klo;
  if (LAT == SPK_SQ_8N || LAT == SPK_SC_26N) rows = rows / 2 + 1;
  if (LAT == SPK_SC_26N) planes = planes / 2 + 1;
  if (Lat<LAT>::DIM == 2) planes = 1;
  dim3 block(64, 4, 1);
  dim3 grid((half_x + 63) / 64, (rows + 3) / 4, planes);
  if (grid.y > 65535 || grid.z > 65535) return fail("lattice too large for the generic kernel grid");
  sweep_generic_kernel<LAT, MODEL, RNG><<<grid, block, 0, h->stream>>>(a);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}


// whole colour-fused sweep of a 2-D lattice; returns 0 launched, -1 not applicable, >0 error
// part: 0 = every tile row and swap the buffers; 1 = first and last tile row only (the rows the
// neighbours' ghost rows depend on); 2 = the tile rows in between, then swap the buffers
template <int LAT, int MODEL>
int launch_tile2d(spk_cuda_handle *h, const SweepArgs &a, int part) {
  if constexpr (Lat<LAT>::DIM != 2) {
    return -1;
  } else {
    using C = Tile2dCfg<LAT>;
    if (!h->fused2d_ok) return -1;
    const int rowlen = h->g.X * Lat<LAT>::NB;
    Tile2dMap tm;
    tm.ntx = (rowlen + C::OW - 1) / C::OW;
    tm.nty = (h->nslab + T2_OH - 1) / T2_OH;
    tm.gh = h->gh;
    tm.ty0 = 0; tm.ty_run = tm.nty; tm.ty_alt = -1;
    if (part == 1) { tm.ty_run = tm.nty > 1 ? 2 : 1; tm.ty_alt = tm.nty > 1 ? tm.nty - 1 : -1; }
    if (part == 2) { tm.ty0 = 1; tm.ty_run = tm.nty - 2; }
    tm.items = tm.ntx * tm.ty_run;
    const bool swap_after = part != 1;
    if (tm.ty_run <= 0) {
      if (swap_after) {
        std::swap(h->base, h->base_alt);
        std::swap(h->tmap, h->tmap_alt);
        h->spin = h->base + (int64_t)h->gh * h->plane;
      }
      return 0;
    }
    if (!h->smem_attr_set[0]) {  // per handle: the attribute belongs to the device the handle lives on
      if (cudaFuncSetAttribute(sweep_tile2d_kernel<LAT, MODEL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               C::SMEM_BYTES) != cudaSuccess)
        return 1;
      h->smem_attr_set[0] = true;
    }
    const int grid_max = env_cap("SPK_CUDA_MAX_CTAS", 3 * 148);
    const int grid = tm.items < grid_max ? tm.items : grid_max;
    int32_t *dst = h->base_alt + (int64_t)h->gh * h->plane;
    sweep_tile2d_kernel<LAT, MODEL><<<grid, 256, C::SMEM_BYTES, h->stream>>>(h->tmap, a, tm, dst);
    h->launches++;
    if (cudaGetLastError() != cudaSuccess) return 1;
    if (swap_after) {  // the destination becomes the lattice (stream order makes this safe for later launches)
      std::swap(h->base, h->base_alt);
      std::swap(h->tmap, h->tmap_alt);
      h->spin = h->base + (int64_t)h->gh * h->plane;
    }
    return 0;
  }
}
// returns 0 launched, -1 not applicable, >0 error
// pair: colours c0 (even) and c0 + 1 of sc/26n in ONE launch (sweep_stream.cuh: stream_chunk_pair)
template <int LAT, int MODEL>
int launch_stream(spk_cuda_handle *h, const SweepArgs &a, bool pair = false) {
  if constexpr (LAT != SPK_SC_26N && LAT != SPK_SC_6N) {
    return -1;
  } else {
    if (pair && (LAT != SPK_SC_26N || ising_like(MODEL) || (a.c0 & 1) || a.nloop != 1)) return -1;
    const Geo &g = a.g;
    const Box &b = a.b;
    if (!h->stream_ok || b.ilo != 0 || b.ihi != g.X || b.jlo != 0 || b.jhi != g.Y) return -1;
    StreamMap sm;
    sm.edge = nullptr;
    sm.nrows = 0;
    const int c0 = a.c0;
    int kfirst = b.klo, nq = b.khi - b.klo;
    if (LAT == SPK_SC_26N) {
      kfirst = b.klo + ((b.klo ^ (c0 >> 2)) & 1);
      nq = kfirst < b.khi ? (b.khi - kfirst + 1) / 2 : 0;
    }
    if (nq <= 0) return 0;
    sm.ntx = (g.X + ST_TX - 1) / ST_TX;
    sm.kfirst = kfirst;
    sm.nq = nq;
    // work items per resident CTA (plane chunks are cut so that there are about this many): more
    // items balance the tail better, fewer re-read fewer chunk-boundary planes (SPK_CUDA_ITEMS_PER_CTA, A/B tests)
    static const int items_per_cta = [] { const char *e = getenv("SPK_CUDA_ITEMS_PER_CTA"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 8; }();
    SweepArgs args = a;  // Geo::lo_wrap / hi_wrap select periodic images (1 GPU) or ghost planes (slabs)
    // P16: packed 16-bit compares (event.cuh) when every spin fits; the Ising shortcut has no compares to pack
    // AD: candidate elimination for the candidate-list apps (sweep_stream.cuh: ADAPT)
    auto launch = [&](auto ONE, auto PK, auto PAIR, auto AD) -> int {
      constexpr bool one = decltype(ONE)::value, pk = decltype(PK)::value && !ising_like(MODEL);
      constexpr bool pr = decltype(PAIR)::value && one && LAT == SPK_SC_26N && !ising_like(MODEL);
      constexpr int ad = (pk && !pr && LAT == SPK_SC_26N && neigh_like(MODEL)) ? decltype(AD)::value : 0;
      using C = StreamCfgOf<LAT, ad>;
      const int grid_max = stream_min_blocks(ad) * 148;
      sm.nty = (g.Y + C::YT - 1) / C::YT;
      int nzc = (items_per_cta * grid_max + sm.ntx * sm.nty - 1) / (sm.ntx * sm.nty);
      if (nzc > nq) nzc = nq;
      nzc = env_cap("SPK_CUDA_MAX_PLANE_CHUNKS", nzc);
      if (nzc < 1) nzc = 1;
      sm.qchunk = (nq + nzc - 1) / nzc;
      sm.nzc = (nq + sm.qchunk - 1) / sm.qchunk;
      sm.items = sm.ntx * sm.nty * sm.nzc;
      const int grid_cap = env_cap("SPK_CUDA_MAX_CTAS", grid_max);
      const int grid = sm.items < grid_cap ? sm.items : grid_cap;
      if constexpr (pr) {  // tile-edge words first (stream order: they see the lattice before the pair launch)
        sm.nrows = g.Y / 2;
        const size_t words = 2 * (size_t)sm.ntx * sm.nrows * sm.nq;
        const size_t words_max = 2 * (size_t)sm.ntx * sm.nrows * ((h->nslab + 1) / 2);  // whole-slab launch: allocate once
        if (words > h->edge_words) {
          cudaFree(h->edge);
          h->edge = nullptr;
          h->edge_words = 0;
          if (cudaMalloc(&h->edge, words_max * sizeof(int32_t)) != cudaSuccess) return 1;
          h->edge_words = words_max;
        }
        sm.edge = h->edge;
        const long long total = (long long)(words / 2);
        edge_pair_kernel<MODEL, pk><<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(args, sm, h->edge);
        h->launches++;
      }
      auto *kern = sweep_stream_kernel<LAT, MODEL, one, pk, pr, ad>;
      bool &attr = h->smem_attr_set[1 + (one ? 0 : 1) + (pk ? 2 : 0) + (pr ? 4 : 0) + 8 * ad];
      if (!attr) {  // per handle: the attribute belongs to the device the handle lives on
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) != cudaSuccess) return 1;
        attr = true;
      }
      kern<<<grid, 256, C::SMEM_BYTES, h->stream>>>(ad ? h->tmap8 : h->tmap, args, sm);
      return 0;
    };
    using T = std::true_type;
    using F = std::false_type;
    using V0 = std::integral_constant<int, 0>;
    using V1 = std::integral_constant<int, 1>;
    const bool nb8 = neigh_like(MODEL) && LAT == SPK_SC_26N && h->pk16 && !pair && h->tmap8_ok;
    const bool adapt = nb8 && h->neigh_adaptive;
    int rc;
    if (pair) rc = h->pk16 ? launch(T{}, T{}, T{}, V0{}) : launch(T{}, F{}, T{}, V0{});
    else if (adapt) rc = args.nloop == 1 ? launch(T{}, T{}, F{}, V1{}) : launch(F{}, T{}, F{}, V1{});
    else if (args.nloop == 1) rc = h->pk16 ? launch(T{}, T{}, F{}, V0{}) : launch(T{}, F{}, F{}, V0{});
    else rc = h->pk16 ? launch(F{}, T{}, F{}, V0{}) : launch(F{}, F{}, F{}, V0{});
    if (rc) return rc;
    h->launches++;
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
  }
}

template <int LAT, int MODEL>
int sweep_set_lm(spk_cuda_handle *h, int rng, int color, const int32_t *box, int nloop) {
  SweepArgs a;
  a.spin = h->spin; a.g = h->g; a.c0 = color - 1; a.nloop = nloop; a.ev = h->ev;
  a.u1 = h->u1; a.u2 = h->u2; a.seeds = h->seeds; a.naccept = h->naccept; a.kstat = h->kstat; a.mask = h->mask;
  a.mirror_lo = h->mirror_active ? h->mirror_lo : nullptr;
  a.mirror_hi = h->mirror_active ? h->mirror_hi : nullptr;
  if (!local_box(h, box, a.b)) return 0;
  h->nattempt += class_count<LAT>(h, a.b, a.c0) * nloop;
  if (rng == RNG_PHILOX) {
    if (h->tiled && !h->mask) {  // masked sweeps: generic kernel only
      int rc = h->use_stream ? launch_stream<LAT, MODEL>(h, a) : -1;
      if (rc == 0) return 0;
      if (rc > 0) return fail("stream kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
      if (h->mirror_active) return fail("peer-memory halo needs the streaming kernel");
      rc = launch_tiled<LAT, MODEL>(h->stream, a, &h->launches);
      if (rc == 0) return 0;
      if (rc > 0) return fail("tiled kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    return launch_generic<LAT, MODEL, RNG_PHILOX>(h, a);
  }
  if (rng == RNG_INJECTED) return launch_generic<LAT, MODEL, RNG_INJECTED>(h, a);
  return launch_generic<LAT, MODEL, RNG_STRICT>(h, a);
}

// colours `color` (odd, 1-based) and color + 1 of sc/26n as one launch: 0 launched, -1 not applicable
template <int MODEL> int sweep_pair_m(spk_cuda_handle *h, int color, const int32_t *box) {
  SweepArgs a;
  a.spin = h->spin; a.g = h->g; a.c0 = color - 1; a.nloop = 1; a.ev = h->ev;
  a.u1 = nullptr; a.u2 = nullptr; a.seeds = nullptr; a.naccept = h->naccept; a.kstat = nullptr; a.mask = nullptr;
  a.mirror_lo = h->mirror_active ? h->mirror_lo : nullptr;
  a.mirror_hi = h->mirror_active ? h->mirror_hi : nullptr;
  if (!local_box(h, box, a.b)) return 0;
  const int rc = launch_stream<SPK_SC_26N, MODEL>(h, a, true);
  if (rc == 0) h->nattempt += class_count<SPK_SC_26N>(h, a.b, a.c0) + class_count<SPK_SC_26N>(h, a.b, a.c0 + 1);
  return rc;
}

// one colour class of the (full x, y) box on the 16-bit shadow: 0 launched, -1 not applicable, >0 error
template <int MODEL> int launch_stream16(spk_cuda_handle *h, int color, const int32_t *box) {
  if constexpr (MODEL != SPK_POTTS && MODEL != SPK_POTTS_PIN) {
    return -1;
  } else {
    using C = Stream16Cfg;
    if (!h->shadow_ok || h->d.lattice != SPK_SC_26N) return -1;
    Box b;
    if (!local_box(h, box, b)) return 0;
    const Geo &g = h->g;
    if (b.ilo != 0 || b.ihi != g.X || b.jlo != 0 || b.jhi != g.Y) return -1;
    const int c0 = color - 1;
    StreamMap sm;
    sm.edge = nullptr;
    sm.nrows = 0;
    const int kfirst = b.klo + ((b.klo ^ (c0 >> 2)) & 1);
    const int nq = kfirst < b.khi ? (b.khi - kfirst + 1) / 2 : 0;
    h->nattempt += class_count<SPK_SC_26N>(h, b, c0);
    if (nq <= 0) return 0;
    sm.ntx = (g.X + S16_TX - 1) / S16_TX;
    sm.nty = (g.Y + C::YT - 1) / C::YT;
    sm.kfirst = kfirst;
    sm.nq = nq;
    static const int items_per_cta = [] { const char *e = getenv("SPK_CUDA_ITEMS_PER_CTA"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 8; }();
    const int grid_max = 2 * 148;
    int nzc = (items_per_cta * grid_max + sm.ntx * sm.nty - 1) / (sm.ntx * sm.nty);
    if (nzc > nq) nzc = nq;
    nzc = env_cap("SPK_CUDA_MAX_PLANE_CHUNKS", nzc);
    if (nzc < 1) nzc = 1;
    sm.qchunk = (nq + nzc - 1) / nzc;
    sm.nzc = (nq + sm.qchunk - 1) / sm.qchunk;
    sm.items = sm.ntx * sm.nty * sm.nzc;
    const int grid_cap = env_cap("SPK_CUDA_MAX_CTAS", grid_max);
    const int grid = sm.items < grid_cap ? sm.items : grid_cap;
    Sweep16Args a;
    a.spin = h->shadow; a.g = g; a.c0 = c0; a.ev = h->ev; a.naccept = h->naccept;
    a.mirror_lo = h->mirror_active ? h->mirror16_lo : nullptr;
    a.mirror_hi = h->mirror_active ? h->mirror16_hi : nullptr;
    a.minus_one = -1;
    auto *kern = sweep_stream16_kernel<MODEL>;
    if (!h->smem16_set) {
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) != cudaSuccess) return 1;
      h->smem16_set = true;
    }
    kern<<<grid, 256, C::SMEM_BYTES, h->stream>>>(h->tmap16, a, sm);
    h->launches++;
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
  }
}

template <int MODEL>
int ModelEntry<MODEL>::sweep_set(spk_cuda_handle *h, int rng, int color, const int32_t *box, int nloop) {
  switch (h->d.lattice) {
    case SPK_SQ_4N: return sweep_set_lm<SPK_SQ_4N, MODEL>(h, rng, color, box, nloop);
    case SPK_SQ_8N: return sweep_set_lm<SPK_SQ_8N, MODEL>(h, rng, color, box, nloop);
    case SPK_TRI: return sweep_set_lm<SPK_TRI, MODEL>(h, rng, color, box, nloop);
    case SPK_SC_6N: return sweep_set_lm<SPK_SC_6N, MODEL>(h, rng, color, box, nloop);
    case SPK_SC_26N: return sweep_set_lm<SPK_SC_26N, MODEL>(h, rng, color, box, nloop);
  }
  return fail("unknown lattice %d", h->d.lattice);
}

template <int MODEL> int ModelEntry<MODEL>::fused(spk_cuda_handle *h, const SweepArgs &a, int part) {
  switch (h->d.lattice) {
    case SPK_SQ_4N: return launch_tile2d<SPK_SQ_4N, MODEL>(h, a, part);
    case SPK_SQ_8N: return launch_tile2d<SPK_SQ_8N, MODEL>(h, a, part);
    case SPK_TRI: return launch_tile2d<SPK_TRI, MODEL>(h, a, part);
    default: return -1;
  }
}

template <int MODEL> int ModelEntry<MODEL>::shadow16(spk_cuda_handle *h, int color, const int32_t *box) {
  return launch_stream16<MODEL>(h, color, box);
}

template <int MODEL> int ModelEntry<MODEL>::pair(spk_cuda_handle *h, int color, const int32_t *box) {
  if constexpr (ising_like(MODEL)) return -1;
  else return sweep_pair_m<MODEL>(h, color, box);
}

}  // namespace spk
