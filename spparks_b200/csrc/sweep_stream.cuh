// Streaming colour-set kernel for the 3-D lattices (sc/26n, sc/6n), Philox mode.
//
// Persistent CTAs walk work items = (x tile of 128 sites) x (y tile of 16 class rows) x (chunk of
// class planes).  Inside an item a CTA streams along z: a ring of lattice planes lives in shared
// memory, each plane brought in by ONE TMA box load (cp.async.bulk.tensor.3d, box = 136 x RP x 1
// int32 = the tile plus a 4-element x halo chunk and a 1-row y halo on each side) that signals an
// mbarrier; the planes of the next class plane are in flight while the current one is updated.
// Every lattice plane is therefore read from L2/HBM once per tile instead of three times, loads
// cost no issue slots, and the stencil windows are read from shared memory with immediate
// offsets (one conflict-free LDS.128 + one LDS.32 per stencil row and thread).
//
// TMA zero-fills box elements outside the tensor; the periodic images that fall there (x and y
// edges of the lattice only) are patched in shared memory with plain loads after the box has
// landed.  The z direction never leaves the tensor: plane -1 / Z is the periodic image (one GPU)
// or the slab's ghost plane (multi-GPU, refreshed by spk_cuda_halo_exchange), chosen by Geo::lo_wrap / hi_wrap.
#pragma once
#include <cuda.h>
#include "sweep_tiled.cuh"

namespace spk {

constexpr int ST_TX = 128;        // tile width in sites: 32 lanes x one 16-byte chunk
constexpr int ST_BX = ST_TX + 8;  // TMA box inner extent
constexpr int ST_CR = 16;         // class rows per tile: 8 warps x 2

// CRows = class rows per tile (a multiple of the 8 warps): 16 for the register-heavy variants
// (2 CTAs per SM), 8 for the candidate-elimination variant (half the shared memory, 4 CTAs per SM)
template <int LAT, int CRows = ST_CR> struct StreamCfg {
  static constexpr int CR = CRows;
  static constexpr int JSTEP = (LAT == SPK_SC_26N) ? 2 : 1;  // class rows / planes are every
  static constexpr int KSTEP = (LAT == SPK_SC_26N) ? 2 : 1;  // JSTEP-th row / KSTEP-th plane
  static constexpr int YT = CR * JSTEP;                      // lattice rows per tile
  static constexpr int RP = YT + 2;                          // rows per plane in smem (y halo)
  static constexpr int RING = 3 + KSTEP;                     // stencil window + planes in flight
  static constexpr int PLANE_BYTES = ST_BX * RP * 4;
  static constexpr int PLANE_STRIDE = (PLANE_BYTES + 127) / 128 * 128;
  static constexpr int BAR_OFF = RING * PLANE_STRIDE;        // mbarriers, one per ring slot
  static constexpr int TAB_OFF = BAR_OFF + 64;               // nb26 offset tables (candidate elimination)
  static constexpr int SMEM_BYTES = TAB_OFF + RING * 32 * 4;
};

struct StreamMap {
  int ntx, nty, nzc, items;  // tiles in x, y; plane chunks; ntx*nty*nzc
  int kfirst, nq, qchunk;    // first class plane (local index), class planes, class planes per item
  // colour-pair launches (PAIR): per (class plane q, class row, x tile) two words made by
  // edge_pair_kernel before the launch: [0] pre-launch spin of the odd-x site left of the tile,
  // [1] post-colour-c0 spin of the tile's first even-x site
  const int32_t *edge;
  int nrows;                 // class rows of a plane (Y / 2)
};

// ---- mbarrier / TMA wrappers (PTX ISA: mbarrier, cp.async.bulk.tensor) ------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra WAIT_%=;\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}

// Periodic images of the halo cells of one smem plane that fall outside the lattice in x or y.
// Only cells the stencil can touch are patched: column x0-1, the first column past the tile (or
// past the row end), row y0-1 and the first row past the lattice.  Threads 0..RP-1 own one row of
// the two columns, threads 64..64+BX-1 one column of the two rows; the loads are issued early
// (when the TMA box is issued) and stored once the box has landed.
struct PatchPlan {
  int col_l, col_r;  // smem column to patch on the left / right, or -1
  int row_t, row_b;  // smem row to patch at the top (y = -1) / bottom (y = Y), or -1
  int gx_l, gx_r;    // lattice x of the periodic images of those columns
};

template <class C>
__device__ __forceinline__ PatchPlan make_patch_plan(int x0, int y0, int rowlen, int Y) {
  PatchPlan pp;
  pp.col_l = x0 == 0 ? 3 : -1;
  pp.gx_l = rowlen - 1;
  const int past = rowlen - x0;  // tile-relative x of the first column past the row end
  pp.col_r = past <= ST_TX ? 4 + past : -1;
  pp.gx_r = 0;
  pp.row_t = y0 == 0 ? 0 : -1;
  const int rb = Y - y0 + 1;
  pp.row_b = rb <= C::RP - 1 ? rb : -1;
  return pp;
}

template <class C>
__device__ __forceinline__ void patch_load(const PatchPlan &pp, const int32_t *gplane, int x0, int y0, int rowlen,
                                           int Y, int tid, int (&v)[2]) {
  v[0] = v[1] = 0;
  if (tid < C::RP) {  // my row of the two columns
    int y = y0 - 1 + tid;
    y = y < 0 ? Y - 1 : (y >= Y ? y - Y : y);
    if (y < Y) {
      const int32_t *row = gplane + (long long)y * rowlen;
      if (pp.col_l >= 0) v[0] = row[pp.gx_l];
      if (pp.col_r >= 0) v[1] = row[pp.gx_r];
    }
  } else if (tid >= 64 && tid < 64 + ST_BX) {  // my column of the two rows
    int x = x0 - 4 + (tid - 64);
    x = x < 0 ? x + rowlen : (x >= rowlen ? x - rowlen : x);
    if (x >= 0 && x < rowlen) {
      if (pp.row_t >= 0) v[0] = gplane[(long long)(Y - 1) * rowlen + x];
      if (pp.row_b >= 0) v[1] = gplane[x];
    }
  }
}

template <class C>
__device__ __forceinline__ void patch_store(const PatchPlan &pp, int32_t *pl, int tid, const int (&v)[2]) {
  if (tid < C::RP) {
    if (pp.col_l >= 0) pl[tid * ST_BX + pp.col_l] = v[0];
    if (pp.col_r >= 0) pl[tid * ST_BX + pp.col_r] = v[1];
  } else if (tid >= 64 && tid < 64 + ST_BX) {
    if (pp.row_t >= 0) pl[pp.row_t * ST_BX + (tid - 64)] = v[0];
    if (pp.row_b >= 0) pl[pp.row_b * ST_BX + (tid - 64)] = v[1];
  }
}

// one class row of one class plane: the thread's 16-byte chunk holds two class sites at P, P+2
// nbtab (candidate-list apps on sc/26n): byte offset of neighbour slot n relative to the site itself
// inside the shared-memory planes of this step, for the run-time fetch of the chosen candidate
template <int LAT, int MODEL, int P, bool ONEPASS, bool P16>
__device__ __forceinline__ int stream_chunk(const SweepArgs &a, const int32_t *const (&pl)[3], int r, int lane,
                                            int32_t *gchunk, int32_t *gmirror, unsigned long long gsite0, int xcell,
                                            int *kstat = nullptr, const int *nbtab = nullptr) {
  using L = Lat<LAT>;
  constexpr int NN = L::NN;
  using R = RowInfo<LAT, 0>;
  // W[row][1 + e] = element e of the chunk, e in [-1, 4]
  int W[9][6];
  static_for<9>([&](auto RR) {
    constexpr int rr = decltype(RR)::value, dy = rr % 3 - 1, dz = rr / 3 - 1;
    if constexpr (R::used(dy, dz)) {
      const int32_t *rowp = pl[dz + 1] + (r + dy) * ST_BX + 4 + 4 * lane;
      const int4 t = *reinterpret_cast<const int4 *>(rowp);
      W[rr][1] = t.x; W[rr][2] = t.y; W[rr][3] = t.z; W[rr][4] = t.w;
      if constexpr (P == 0 && R::dmin(dy, dz) < 0) W[rr][0] = rowp[-1];
      if constexpr (P == 1 && R::dmax(dy, dz) > 0) W[rr][5] = rowp[4];
    }
  });

  int cur[2] = {W[4][1 + P], W[4][3 + P]};
  const int before0 = cur[0], before1 = cur[1];
  int accepted = 0;
  // Philox block of the chunk's two class sites (event.cuh: philox_words); an ising block spans
  // two chunks, this one is its upper half when bit 2 of the (4-aligned) chunk start is set
  const bool upper = ising_like(MODEL) ? ((xcell >> 2) & 1) : false;
  const unsigned long long leader = gsite0 + P - (upper ? 4 : 0);
  const int nloop = ONEPASS ? 1 : a.nloop;  // ONEPASS: the common nloop == 1 without loop overhead
#pragma unroll
  for (int l = 0; l < nloop; l++) {
    const uint4 blk = philox4x32_10((uint32_t)leader, (uint32_t)(leader >> 32), a.ev.sweep,
                                    (uint32_t)l * 16u + (uint32_t)(a.c0 + 1), a.ev.rk);
    static_for<2>([&](auto H) {
      constexpr int h = decltype(H)::value, e = P + 2 * h;
      int v[NN];
      static_for<NN>([&](auto N_) {
        constexpr int n = decltype(N_)::value;
        constexpr Off o = L::off(0, n);
        v[n] = W[(o.dz + 1) * 3 + (o.dy + 1)][1 + e + o.dx];
      });
      Draws<RNG_PHILOX> d;
      if constexpr (ising_like(MODEL)) {
        d.w1 = upper ? (h ? blk.w : blk.z) : (h ? blk.y : blk.x);
        d.w2 = d.w1 << 1;
      } else {
        d.w1 = h ? blk.z : blk.x;
        d.w2 = h ? blk.w : blk.y;
      }
      int nw;
      if constexpr (LAT == SPK_SC_26N && neigh_like(MODEL)) {
        const uint32_t self = smem_u32(pl[1] + r * ST_BX + 4 + 4 * lane + e), tab = smem_u32(nbtab);
        auto fetch = [&](int n) {
          int off, val;
          asm("ld.shared.s32 %0, [%1];" : "=r"(off) : "r"(tab + 4u * (uint32_t)n));
          asm("ld.shared.s32 %0, [%1];" : "=r"(val) : "r"(self + (uint32_t)off));
          return val;
        };
        nw = site_event<MODEL, RNG_PHILOX, NN, P16>(a.ev, cur[h], v, d, kstat, nullptr, fetch);
      } else {
        nw = site_event<MODEL, RNG_PHILOX, NN, P16>(a.ev, cur[h], v, d, kstat);
      }
      accepted += (nw != cur[h]);
      cur[h] = nw;
    });
  }
  if (cur[0] != before0 || cur[1] != before1) {
    W[4][1 + P] = cur[0];
    W[4][3 + P] = cur[1];
    const int4 nv = make_int4(W[4][1], W[4][2], W[4][3], W[4][4]);
    *reinterpret_cast<int4 *>(gchunk) = nv;
    if (gmirror) *reinterpret_cast<int4 *>(gmirror) = nv;  // the neighbour's ghost copy of this chunk
  }
  return accepted;
}

// The same chunk for the candidate-list apps on sc/26n by candidate elimination (event.cuh:
// site_event_nb26).  The packed neighbour words of both sites are built first so that the
// register window is dead before the per-site loops start; a candidate's value is read back from
// the shared-memory planes: nbtab[b] = byte offset of the neighbour behind mask bit b relative to
// the site itself.
template <int MODEL, int P, bool ONEPASS>
__device__ __forceinline__ int stream_chunk_nb26(const SweepArgs &a, const int32_t *const (&pl)[3], int r, int lane,
                                                 int32_t *gchunk, int32_t *gmirror, unsigned long long gsite0,
                                                 const int *nbtab, int *kstat) {
  using L = Lat<SPK_SC_26N>;
  const int32_t *own = pl[1] + r * ST_BX + 4 + 4 * lane;
  uint32_t wd[2][13];
  {
    int W[9][5];  // W[row][e - P + 1]: elements P-1 .. P+3 of the chunk's stencil rows
    static_for<9>([&](auto RR) {
      constexpr int rr = decltype(RR)::value, dy = rr % 3 - 1, dz = rr / 3 - 1;
      const int32_t *rowp = pl[dz + 1] + (r + dy) * ST_BX + 4 + 4 * lane;
      const int4 t = *reinterpret_cast<const int4 *>(rowp);
      if constexpr (P == 0) { W[rr][0] = rowp[-1]; W[rr][1] = t.x; W[rr][2] = t.y; W[rr][3] = t.z; W[rr][4] = t.w; }
      else { W[rr][0] = t.x; W[rr][1] = t.y; W[rr][2] = t.z; W[rr][3] = t.w; W[rr][4] = rowp[4]; }
    });
    static_for<2>([&](auto H) {
      constexpr int h = decltype(H)::value;
      static_for<13>([&](auto J) {
        constexpr int j = decltype(J)::value;
        constexpr Off o0 = L::off(0, j), o1 = L::off(0, j + 13);
        wd[h][j] = pk_negpair(W[(o0.dz + 1) * 3 + (o0.dy + 1)][1 + 2 * h + o0.dx],
                              W[(o1.dz + 1) * 3 + (o1.dy + 1)][1 + 2 * h + o1.dx], a.ev);
      });
    });
  }
  const int4 chunk = *reinterpret_cast<const int4 *>(own);
  int cur[2] = {P ? chunk.y : chunk.x, P ? chunk.w : chunk.z};
  const int before0 = cur[0], before1 = cur[1];
  int accepted = 0;
  const unsigned long long leader = gsite0 + P;
  const uint32_t tab = smem_u32(nbtab);
  const int nloop = ONEPASS ? 1 : a.nloop;
#pragma unroll
  for (int l = 0; l < nloop; l++) {
    const uint4 blk = philox4x32_10((uint32_t)leader, (uint32_t)(leader >> 32), a.ev.sweep,
                                    (uint32_t)l * 16u + (uint32_t)(a.c0 + 1), a.ev.rk);
    const uint32_t self = smem_u32(own + P);
    const int prev0 = cur[0], prev1 = cur[1];
#if ST_NB26_JOINT
    site_event_nb26x2<MODEL>(a.ev, cur, wd, blk, [&](int h, int b) {
      int off, val;
      asm("ld.shared.s32 %0, [%1];" : "=r"(off) : "r"(tab + 4u * (uint32_t)b));
      asm("ld.shared.s32 %0, [%1];" : "=r"(val) : "r"(self + 8u * (uint32_t)h + (uint32_t)off));
      return val;
    }, kstat);
#else
    static_for<2>([&](auto H) {
      constexpr int h = decltype(H)::value;
      cur[h] = site_event_nb26<MODEL>(a.ev, cur[h], wd[h], h ? blk.z : blk.x, h ? blk.w : blk.y, [&](int b) {
        int off, val;
        asm("ld.shared.s32 %0, [%1];" : "=r"(off) : "r"(tab + 4u * (uint32_t)b));
        asm("ld.shared.s32 %0, [%1];" : "=r"(val) : "r"(self + 8u * h + (uint32_t)off));
        return val;
      }, kstat);
    });
#endif
    accepted += (cur[0] != prev0) + (cur[1] != prev1);
  }
  if (cur[0] != before0 || cur[1] != before1) {
    const int4 nv = P ? make_int4(chunk.x, cur[0], chunk.z, cur[1]) : make_int4(cur[0], chunk.y, cur[1], chunk.w);
    *reinterpret_cast<int4 *>(gchunk) = nv;
    if (gmirror) *reinterpret_cast<int4 *>(gmirror) = nv;  // the neighbour's ghost copy of this chunk
  }
  return accepted;
}

// Colour pair (c0, c0+1) of sc/26n in one pass over a class row: the two colours differ only in
// the x parity, so the odd-x sites of a class row depend on the even-x sites of the SAME row and
// on nothing else the launch changes (rows y+-1 and planes k+-1 belong to other colours).  A
// thread updates the even sites of its chunk (elements 0, 2), takes the next lane's new element 0
// by shuffle, then updates the odd sites (1, 3) and stores the chunk once.  Across the tile edge
// the two values a tile cannot know come from edge_pair_kernel (StreamMap::edge): the pre-launch
// spin left of the tile (the neighbouring tile may already have rewritten it in memory) and the
// post-c0 spin of the next tile's first site.  A launch then reads the lattice once for two
// colours: 20 instead of 36 bytes per site-update and sweep.
template <int LAT, int MODEL, bool P16>
__device__ __forceinline__ int stream_chunk_pair(const SweepArgs &a, const int32_t *const (&pl)[3], int r, int lane,
                                                 bool valid, int32_t *gchunk, int32_t *gmirror,
                                                 unsigned long long gsite0, bool use_left, int left_pre,
                                                 bool use_right, int right_new) {
  using L = Lat<LAT>;
  constexpr int NN = L::NN;
  using R = RowInfo<LAT, 0>;
  int W[9][6];  // W[row][1 + e] = element e of the chunk, e in [-1, 4]
  static_for<9>([&](auto RR) {
    constexpr int rr = decltype(RR)::value, dy = rr % 3 - 1, dz = rr / 3 - 1;
    if constexpr (R::used(dy, dz)) {
      const int32_t *rowp = pl[dz + 1] + (r + dy) * ST_BX + 4 + 4 * lane;
      const int4 t = *reinterpret_cast<const int4 *>(rowp);
      W[rr][1] = t.x; W[rr][2] = t.y; W[rr][3] = t.z; W[rr][4] = t.w;
      W[rr][0] = rowp[-1];
      W[rr][5] = rowp[4];
    }
  });
  if (use_left) W[4][0] = left_pre;
  const int before[4] = {W[4][1], W[4][2], W[4][3], W[4][4]};
  int accepted = 0;
  static_for<2>([&](auto PP) {
    constexpr int P = decltype(PP)::value;  // 0: colour c0 (even x), 1: colour c0 + 1 (odd x)
    if constexpr (P == 1) {
      const int nxt = __shfl_down_sync(0xffffffffu, W[4][1], 1);  // next lane's updated element 0
      W[4][5] = use_right ? right_new : nxt;
    }
    const unsigned long long leader = gsite0 + P;
    const uint4 blk = philox4x32_10((uint32_t)leader, (uint32_t)(leader >> 32), a.ev.sweep,
                                    (uint32_t)(a.c0 + P + 1), a.ev.rk);
    int nw[2];
    static_for<2>([&](auto H) {
      constexpr int h = decltype(H)::value, e = P + 2 * h;
      int v[NN];
      static_for<NN>([&](auto N_) {
        constexpr int n = decltype(N_)::value;
        constexpr Off o = L::off(0, n);
        v[n] = W[(o.dz + 1) * 3 + (o.dy + 1)][1 + e + o.dx];
      });
      Draws<RNG_PHILOX> d;
      d.w1 = h ? blk.z : blk.x;
      d.w2 = h ? blk.w : blk.y;
      nw[h] = site_event<MODEL, RNG_PHILOX, NN, P16>(a.ev, W[4][1 + e], v, d);
      accepted += (nw[h] != W[4][1 + e]);
    });
    // the two sites of one colour are not neighbours: write both after both were evaluated
    W[4][1 + P] = nw[0];
    W[4][3 + P] = nw[1];
  });
  if (!valid) return 0;
  if (W[4][1] != before[0] || W[4][2] != before[1] || W[4][3] != before[2] || W[4][4] != before[3]) {
    const int4 nv = make_int4(W[4][1], W[4][2], W[4][3], W[4][4]);
    *reinterpret_cast<int4 *>(gchunk) = nv;
    if (gmirror) *reinterpret_cast<int4 *>(gmirror) = nv;
  }
  return accepted;
}

// The values a colour-pair launch needs across its tile edges (see stream_chunk_pair), one thread
// per (x tile, class row, class plane): the site at the tile's first column is evaluated for
// colour c0 exactly as the tile's own lane 0 will evaluate it (same neighbours, same Philox
// words) but stored here instead of into the lattice, next to the current spin left of it.
template <int MODEL, bool P16>
__global__ void __launch_bounds__(256) edge_pair_kernel(const __grid_constant__ SweepArgs a,
                                                        const __grid_constant__ StreamMap sm, int32_t *edge) {
  using L = Lat<SPK_SC_26N>;
  constexpr int NN = L::NN;
  const Geo &g = a.g;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)sm.ntx * sm.nrows * sm.nq) return;
  const int b = (int)(idx % sm.ntx), rc = (int)((idx / sm.ntx) % sm.nrows), q = (int)(idx / ((long long)sm.ntx * sm.nrows));
  const int i = b * ST_TX, j = 2 * rc + ((a.c0 >> 1) & 1), k = sm.kfirst + 2 * q;
  const int xi[3] = {i == 0 ? g.X - 1 : i - 1, i, i == g.X - 1 ? 0 : i + 1};
  const int yj[3] = {j == 0 ? g.Y - 1 : j - 1, j, j == g.Y - 1 ? 0 : j + 1};
  const int zk[3] = {k == 0 ? g.lo_wrap : k - 1, k, k == g.Z - 1 ? g.hi_wrap : k + 1};
  const long long sy = g.X, sz = sy * g.Y;
  int v[NN];
#pragma unroll
  for (int n = 0; n < NN; n++) {
    const Off o = L::off(0, n);
    v[n] = a.spin[zk[o.dz + 1] * sz + yj[o.dy + 1] * sy + xi[o.dx + 1]];
  }
  const int cur = a.spin[zk[1] * sz + yj[1] * sy + i];
  const unsigned long long gsite = ((unsigned long long)(k + g.g0) * g.gY + j) * (unsigned long long)g.gX + i;
  const uint4 blk = philox4x32_10((uint32_t)gsite, (uint32_t)(gsite >> 32), a.ev.sweep, (uint32_t)(a.c0 + 1), a.ev.rk);
  Draws<RNG_PHILOX> d;
  d.w1 = blk.x;  // i is a multiple of 4: the site leads its Philox block (event.cuh: philox_words)
  d.w2 = blk.y;
  edge[2 * idx] = v[4];  // (dx,dy,dz) = (-1,0,0): neighbour 4 in reference order
  edge[2 * idx + 1] = site_event<MODEL, RNG_PHILOX, NN, P16>(a.ev, cur, v, d);
}

// ADAPT (sc/26n candidate-list apps, packed spins): candidate elimination, 8 class rows per tile
// and 4 CTAs per SM (the per-site loop is latency bound: more resident warps, fewer registers)
#ifndef ST_NB26_JOINT
#define ST_NB26_JOINT 0  // 1: both sites of a chunk in one loop (measured slower: the loop runs to the warp maximum over 64 sites)
#endif
#ifndef ST_ADAPT_MINB
#define ST_ADAPT_MINB 3
#endif
// NBV (sc/26n candidate-list apps with packed spins): 0 = all-pairs chain on 16-row tiles, 2 CTAs
// per SM; 1 = candidate elimination on 8-row tiles, 3 CTAs per SM.  (The chain on 8-row tiles with
// 3 / 4 CTAs per SM measured 40.3 / 33.6 against 41.8 Gsites/s: it is bound by ALU issue, not latency.)
template <int LAT, int NBV> using StreamCfgOf = StreamCfg<LAT, NBV ? 8 : ST_CR>;
constexpr int stream_min_blocks(int nbv) { return nbv == 1 ? ST_ADAPT_MINB : 2; }

template <int LAT, int MODEL, bool ONEPASS, bool P16, bool PAIR = false, int NBV = 0>
__global__ void __launch_bounds__(256, stream_min_blocks(NBV))
sweep_stream_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ SweepArgs a,
                    const __grid_constant__ StreamMap sm) {
  constexpr bool ADAPT = NBV == 1;
  using C = StreamCfgOf<LAT, NBV>;
  constexpr int TPW = C::CR / 8;  // class rows per warp
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + C::BAR_OFF);
  int *nbtabs = reinterpret_cast<int *>(smem_raw + C::TAB_OFF);
  auto plane_ptr = [&](uint32_t slot) { return reinterpret_cast<int32_t *>(smem_raw + slot * C::PLANE_STRIDE); };

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const Geo &g = a.g;
  const int rowlen = g.X, Y = g.Y;
  const long long sy = rowlen, sz = sy * Y;

  if (tid == 0) {
    for (int i = 0; i < C::RING; i++) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if constexpr (LAT == SPK_SC_26N && neigh_like(MODEL) && !PAIR) {
    // nbtabs[c][b]: the window's middle plane sits in ring slot (c + 1) % RING; entry b = mask bit b
    // (candidate elimination) or neighbour slot b (all-pairs chain) -> (dx,dy,dz) in reference
    // order (lattice.cuh) -> byte offset from the site's own element
    if (tid < C::RING * 32) {
      const int c = tid >> 5, b = tid & 31;
      int off = 0;
      if (ADAPT ? ((NB26_VALID >> b) & 1) != 0 : b < 26) {
        const Off o = Lat<SPK_SC_26N>::off(0, ADAPT ? nb26_bit_to_slot(b) : b);
        const int mid = (c + 1) % C::RING, pln = (c + 1 + o.dz + C::RING) % C::RING;
        off = (pln - mid) * C::PLANE_STRIDE + (o.dy * ST_BX + o.dx) * 4;
      }
      nbtabs[tid] = off;
    }
  }
  __syncthreads();

  // plane kk in -1..Z -> plane that holds it: the periodic image on one GPU (lo_wrap = Z-1,
  // hi_wrap = 0), the ghost plane of the slab otherwise (lo_wrap = -1, hi_wrap = Z)
  auto zmap = [&](int kk) { return kk < 0 ? g.lo_wrap : (kk >= g.Z ? g.hi_wrap : kk); };
  // Ring bookkeeping without divisions: `slot` = ring slot of the window's first plane (k-1),
  // advanced with a compare-and-subtract; `phase` bit i = parity the next wait on barrier i expects
  // (every use of a slot is one TMA box and one wait by every thread).  All CTA-uniform.
  auto adv = [](uint32_t sl, uint32_t d) { sl += d; return sl >= (uint32_t)C::RING ? sl - C::RING : sl; };
  uint32_t slot = 0, phase = 0;
  auto issue = [&](int kk, uint32_t sl, int x0, int y0) {  // plane kk (local, -1..Z) into ring slot sl
    uint64_t *bar = &full[sl];
    mbar_expect_tx(bar, C::PLANE_BYTES);
    tma_load_3d(plane_ptr(sl), &tmap, bar, x0 - 4, y0 - 1, zmap(kk) + 1);  // tensor plane 0 = lower ghost plane
  };
  auto wait = [&](uint32_t sl) {
    mbar_wait(&full[sl], (phase >> sl) & 1u);
    phase ^= 1u << sl;
  };

  int accepted = 0;
  [[maybe_unused]] int kstat = 0;  // candidate-list apps: sum of nevent over this thread's sites
  const bool mirrors = a.mirror_lo != nullptr || a.mirror_hi != nullptr;  // boundary launches of a slab run only
  const unsigned long long plane_sites = (unsigned long long)g.gY * (unsigned long long)g.gX;
  for (int item = blockIdx.x; item < sm.items; item += gridDim.x) {
    const int tx = item % sm.ntx, ty = (item / sm.ntx) % sm.nty, zc = item / (sm.ntx * sm.nty);
    const int x0 = tx * ST_TX, y0 = ty * C::YT;
    const int q0 = zc * sm.qchunk, q1 = min(q0 + sm.qchunk, sm.nq), nsteps = q1 - q0;
    const int kstart = sm.kfirst + q0 * C::KSTEP;
    const bool edge = x0 == 0 || x0 + ST_TX >= rowlen || y0 == 0 || y0 + C::YT >= Y;

    const PatchPlan pp = make_patch_plan<C>(x0, y0, rowlen, Y);

    if (tid == 0)
      for (int p = 0; p < 3; p++) issue(kstart - 1 + p, adv(slot, p), x0, y0);
    if (edge) {  // prologue: the three planes of the first stencil window
      int pv0[3][2];
      for (int p = 0; p < 3; p++)
        patch_load<C>(pp, a.spin + (long long)zmap(kstart - 1 + p) * sz, x0, y0, rowlen, Y, tid, pv0[p]);
      for (int p = 0; p < 3; p++) {
        wait(adv(slot, p));
        patch_store<C>(pp, plane_ptr(adv(slot, p)), tid, pv0[p]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
    }

    // the class rows of this warp: everything that does not depend on the plane is computed once
    // per item; the plane-dependent parts are running pointers / indices advanced per step
    const int x = x0 + 4 * lane;
    int yrow[TPW];
    bool live[TPW];
    int32_t *gptr[TPW];            // the thread's chunk in plane k of the lattice
    unsigned long long gsite[TPW];  // its global site index
#pragma unroll
    for (int t = 0; t < TPW; t++) {
      const int cr = warp + 8 * t;
      yrow[t] = LAT == SPK_SC_26N ? y0 + 2 * cr + ((a.c0 >> 1) & 1) : y0 + cr;
      live[t] = yrow[t] < Y && (PAIR || x < rowlen);  // y: warp-uniform; PAIR: lanes past the row end run along (shuffle)
      gptr[t] = a.spin + ((long long)kstart * sz + (long long)yrow[t] * sy + x);
      gsite[t] = ((unsigned long long)(kstart + g.g0) * g.gY + yrow[t]) * (unsigned long long)g.gX + x;
    }

    int pv[C::KSTEP][2];  // patch values of the planes in flight
    for (int s = 0; s < nsteps; s++) {
      const int k = kstart + s * C::KSTEP;
      const uint32_t w0 = slot, w1 = adv(slot, 1), w2 = adv(slot, 2);  // planes k-1, k, k+1

      if (s > 0) {  // planes k .. k+KSTEP-1 shifted into the window: wait, then patch
        for (int p = 0; p < C::KSTEP; p++) wait(adv(slot, 3 - C::KSTEP + p));
        if (edge) {
          for (int p = 0; p < C::KSTEP; p++) patch_store<C>(pp, plane_ptr(adv(slot, 3 - C::KSTEP + p)), tid, pv[p]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncthreads();
        }
      } else if (!edge) {
        wait(w0); wait(w1); wait(w2);
      }

      if (s + 1 < nsteps) {  // next step's planes: TMA boxes and (edge tiles) their patch loads
        if (tid == 0)
          for (int p = 0; p < C::KSTEP; p++) issue(k + 2 + p, adv(slot, 3 + p), x0, y0);
        if (edge)
          for (int p = 0; p < C::KSTEP; p++)
            patch_load<C>(pp, a.spin + (long long)zmap(k + 2 + p) * sz, x0, y0, rowlen, Y, tid, pv[p]);
      }

      const int32_t *const pl[3] = {plane_ptr(w0), plane_ptr(w1), plane_ptr(w2)};
      long long moff = 0;  // byte distance plane k -> the ring neighbour's ghost plane that mirrors it (CTA-uniform)
      bool mirror = false;
      if (mirrors) {
        const long long here = reinterpret_cast<long long>(a.spin + (long long)k * sz);
        if (k == 0 && a.mirror_lo) { mirror = true; moff = reinterpret_cast<long long>(a.mirror_lo) - here; }
        if (k == g.Z - 1 && a.mirror_hi) { mirror = true; moff = reinterpret_cast<long long>(a.mirror_hi) - here; }
      }
      [[maybe_unused]] const int *nbtab = nbtabs + w0 * 32;
#pragma unroll
      for (int t = 0; t < TPW; t++) {
        const int y = yrow[t];
        const int px = LAT == SPK_SC_26N ? (a.c0 & 1) : ((a.c0 + y + k + g.g0) & 1);
        int32_t *gchunk = gptr[t];
        int32_t *gmirror = mirror ? reinterpret_cast<int32_t *>(reinterpret_cast<long long>(gchunk) + moff) : nullptr;
        const unsigned long long gsite0 = gsite[t];
        if (live[t]) {
          const int r = y - y0 + 1;
          if constexpr (PAIR) {
            const bool valid = x < rowlen;
            const long long ebase = ((long long)(q0 + s) * sm.nrows + (y >> 1)) * sm.ntx;
            const bool use_left = lane == 0, use_right = valid && (lane == 31 || x + 4 >= rowlen);
            int left_pre = 0, right_new = 0;
            if (use_left) left_pre = sm.edge[2 * (ebase + tx)];
            if (use_right) right_new = sm.edge[2 * (ebase + (x + 4 >= rowlen ? 0 : tx + 1)) + 1];
            accepted += stream_chunk_pair<LAT, MODEL, P16>(a, pl, r, lane, valid, gchunk, gmirror, gsite0, use_left,
                                                           left_pre, use_right, right_new);
          } else if constexpr (ADAPT) {
            static_assert(!ADAPT || (LAT == SPK_SC_26N && P16 && neigh_like(MODEL)), "candidate elimination: sc/26n candidate-list apps, packed spins");
            if (px) accepted += stream_chunk_nb26<MODEL, 1, ONEPASS>(a, pl, r, lane, gchunk, gmirror, gsite0, nbtab, &kstat);
            else accepted += stream_chunk_nb26<MODEL, 0, ONEPASS>(a, pl, r, lane, gchunk, gmirror, gsite0, nbtab, &kstat);
          } else {
            if (px) accepted += stream_chunk<LAT, MODEL, 1, ONEPASS, P16>(a, pl, r, lane, gchunk, gmirror, gsite0, x, &kstat, nbtab);
            else accepted += stream_chunk<LAT, MODEL, 0, ONEPASS, P16>(a, pl, r, lane, gchunk, gmirror, gsite0, x, &kstat, nbtab);
          }
        }
        gptr[t] += C::KSTEP * sz;
        gsite[t] += (unsigned long long)C::KSTEP * plane_sites;
      }
      __syncthreads();  // planes k-1 (and k) may now be overwritten by the next step's loads
      if (s + 1 < nsteps) slot = adv(slot, C::KSTEP);
    }
    slot = adv(slot, 3);  // the last window's three slots are free again (everybody passed the barrier)
  }
  accepted = __reduce_add_sync(0xffffffffu, accepted);
  if (lane == 0 && accepted) atomicAdd(a.naccept, (unsigned long long)accepted);
  if constexpr (neigh_like(MODEL)) {
    if (a.kstat) {
      kstat = __reduce_add_sync(0xffffffffu, kstat);
      if (lane == 0 && kstat) atomicAdd(a.kstat, (unsigned long long)kstat);
    }
  }
}

}  // namespace spk
