// Colour-fused sweep kernel for the 2-D lattices (sq/4n, sq/8n, tri), Philox mode.
//
// One launch performs a WHOLE sweep (all colour classes, in colour order) with every spin read
// from and written to HBM once: a CTA brings a tile plus a halo of HY rows / HX elements into
// shared memory with one TMA box load, applies colour 1, 2, ... C to the tile in place
// (__syncthreads between colours), and stores the interior.  After colour step c the outermost c
// rings of the tile are stale (their neighbours were not all present), which is why the halo is C
// site-reaches deep: the interior that is stored is exactly what C separate per-colour kernels
// produce.  Halo sites are updated redundantly by the neighbouring tiles with the same result,
// because the random stream is keyed on (seed, global site, sweep, colour) and not on who computes.
// Traffic: 8 B per site and sweep (+ halo) instead of the 4*(C+1) B of the per-colour kernels.
// The sweep is out of place (source lattice -> destination lattice, swapped by the host after the
// launch): a tile's halo must hold PRE-sweep spins, which an in-place store of a neighbouring
// tile's interior would destroy.
//
// Inside a colour step each warp marches down its band of rows keeping the three stencil rows of
// its 8-element strips in registers, so a row is read from shared memory once per colour.
// Periodic images: y through the ghost rows of the slab (GH rows deep, refreshed once per sweep),
// x by patching the halo columns of the tiles at the row ends after the TMA load.
#pragma once
#include <cuda.h>
#include "sweep_stream.cuh"

namespace spk {

constexpr int T2_BW = 256;  // TMA box width in elements (the hardware maximum per dimension)
constexpr int T2_OH = 64;   // interior rows per tile
constexpr int T2_GH = 4;    // ghost rows kept on each side of a 2-D slab (>= HY of every lattice)

template <int LAT> struct Tile2dCfg {
  using L = Lat<LAT>;
  static constexpr int C = L::NC;                          // colour steps fused
  static constexpr int RX = L::NB;                         // x reach of one step, in elements (1 cell)
  static constexpr int HY = C;                             // halo rows
  static constexpr int HX = 8;                             // halo elements: one whole strip, so that
                                                           // strips are 8-aligned in the lattice too
  static_assert(RX * C <= HX, "x halo too shallow");
  static constexpr int BH = T2_OH + 2 * HY;                // box rows
  static constexpr int OW = T2_BW - 2 * HX;                // interior width
  static constexpr int SMEM_BYTES = T2_BW * BH * 4 + 16;
  static_assert(HY <= T2_GH, "ghost rows too shallow");
};

struct Tile2dMap {
  int ntx, nty, items;
  int gh;      // ghost rows allocated on each side of the slab (tensor row 0 = lowest ghost row)
  // tile rows of this launch: ty0, ty0+1, ... (ty_run of them), except that run index 1 is
  // ty_alt when ty_alt >= 0 (the multi-GPU path launches the first and last tile row on their own)
  int ty0, ty_run, ty_alt;
};

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}

// one 8-element strip of one row for colour step c0; P = offset of the first class site
template <int LAT, int MODEL, int P>
__device__ __forceinline__ int tile2d_strip(const SweepArgs &a, int c0, int32_t *own, const int (&Wm)[16], int (&Wc)[16],
                                            const int (&Wp)[16], unsigned long long grow, int xg, bool count) {
  using L = Lat<LAT>;
  constexpr int NN = L::NN, NB = L::NB;
  constexpr int S = 2 * NB, NS = 8 / S;
  constexpr int M = (LAT == SPK_TRI) ? (P & 1) : 0;
  using R = RowInfo<LAT, M>;
  int cur[NS], before[NS];
  static_for<NS>([&](auto S_) {
    constexpr int s = decltype(S_)::value;
    before[s] = cur[s] = Wc[4 + P + S * s];
  });
  int accepted = 0;
  constexpr int SPB = ising_like(MODEL) ? 4 : 2;  // class sites per Philox block (event.cuh)
  constexpr int NBLK = (NS + SPB - 1) / SPB;
  static_for<NBLK>([&](auto PR) {
    constexpr int pr = decltype(PR)::value;
    const bool upper = (ising_like(MODEL) && NS < SPB) ? ((xg >> 3) & 1) : false;
    const unsigned long long leader = grow + P + S * (SPB * pr) - (upper ? 8 : 0);
    const uint4 blk = philox4x32_10((uint32_t)leader, (uint32_t)(leader >> 32), a.ev.sweep, (uint32_t)(c0 + 1),
                                    a.ev.rk);
    static_for<(NS < SPB ? NS : SPB)>([&](auto H) {
      constexpr int h = decltype(H)::value, s = SPB * pr + h, p = P + S * s;
      int v[NN];
      static_for<NN>([&](auto N_) {
        constexpr int n = decltype(N_)::value;
        constexpr Off o = L::off(M, n);
        v[n] = (o.dy < 0 ? Wm : (o.dy > 0 ? Wp : Wc))[4 + p + R::delta(n)];
      });
      Draws<RNG_PHILOX> d;
      if constexpr (ising_like(MODEL)) {
        const uint32_t lo = h == 0 ? blk.x : (h == 1 ? blk.y : (h == 2 ? blk.z : blk.w));
        const uint32_t hi = h == 0 ? blk.z : blk.w;
        d.w1 = upper ? hi : lo;
        d.w2 = d.w1 << 1;
      } else {
        d.w1 = h ? blk.z : blk.x;
        d.w2 = h ? blk.w : blk.y;
      }
      // (sites of the two outermost strips whose neighbours are already stale are updated too:
      // they lie in the part of the halo that is stale from this step on and is never stored)
      const int nw = site_event<MODEL, RNG_PHILOX, NN>(a.ev, cur[s], v, d);
      accepted += (nw != cur[s]);
      cur[s] = nw;
    });
  });
  // updated spins go back to the tile; they are never read again within this colour step
  static_for<2>([&](auto C_) {
    constexpr int c = decltype(C_)::value;
    bool changed = false;
    static_for<NS>([&](auto S_) {
      constexpr int s = decltype(S_)::value, p = P + S * s;
      if constexpr (p / 4 == c) { changed |= (cur[s] != before[s]); Wc[4 + p] = cur[s]; }
    });
    if (changed)
      *reinterpret_cast<int4 *>(own + 4 * c) = make_int4(Wc[4 + 4 * c], Wc[5 + 4 * c], Wc[6 + 4 * c], Wc[7 + 4 * c]);
  });
  return count ? accepted : 0;
}

template <int LAT, int MODEL>
__global__ void __launch_bounds__(256, 3)
sweep_tile2d_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ SweepArgs a,
                    const __grid_constant__ Tile2dMap tm, int32_t *__restrict__ dst) {
  using L = Lat<LAT>;
  using C = Tile2dCfg<LAT>;
  constexpr int NB = L::NB;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  int32_t *tile = reinterpret_cast<int32_t *>(smem_raw);
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + T2_BW * C::BH * 4);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const Geo &g = a.g;
  const int rowlen = g.X * NB, Y = g.Y;  // Y = owned rows of the slab

  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  uint32_t phase = 0;
  int accepted = 0;
  for (int item = blockIdx.x; item < tm.items; item += gridDim.x) {
    const int tx = item % tm.ntx, trun = item / tm.ntx;
    const int ty = (tm.ty_alt >= 0 && trun == 1) ? tm.ty_alt : tm.ty0 + trun;
    const int xo = tx * C::OW, yo = ty * T2_OH;     // interior origin (element, owned row)
    const int xb = xo - C::HX, yb = yo - C::HY;     // box origin
    if (tid == 0) {
      mbar_expect_tx(bar, T2_BW * C::BH * 4);
      tma_load_2d(tile, &tmap, bar, xb, yb + tm.gh);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    // periodic images in x: box columns left of 0 or right of the row end
    if (xb < 0 || xb + T2_BW > rowlen) {
      for (int c = tid; c < T2_BW * C::BH; c += 256) {
        const int r = c / T2_BW, t = c - r * T2_BW;
        int x = xb + t;
        const int yrow = yb + r;
        if ((x < 0 || x >= rowlen) && yrow >= -tm.gh && yrow < Y + tm.gh) {
          x %= rowlen;
          if (x < 0) x += rowlen;
          tile[c] = a.spin[(long long)yrow * rowlen + x];
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const int e0 = 8 * lane;  // this lane's strip inside the box
    // global element of the strip start (strips are 8-aligned in the lattice and never straddle
    // the row end) and whether the strip belongs to the stored interior
    int xg = xb + e0;
    xg %= rowlen;
    if (xg < 0) xg += rowlen;
    const bool x_interior = e0 >= C::HX && e0 < T2_BW - C::HX && xb + e0 < rowlen;
    const unsigned long long rowsize = (unsigned long long)(g.gX * NB);

#pragma unroll 1
    for (int c0 = 0; c0 < C::C; c0++) {
      // rows that can be updated in step c0: all neighbours still valid (rect shrinks by one
      // site reach per colour step); this warp's band of them
      const int ylo = c0 + 1, yhi = C::BH - c0 - 1;
      const int nrow = yhi - ylo;
      const int y0 = ylo + (warp * nrow) / 8, y1 = ylo + ((warp + 1) * nrow) / 8;
      if (y0 < y1) {
        auto load_row = [&](int (&w)[16], int yl) {
          const int32_t *row = tile + yl * T2_BW + e0;
          load_chunk(w, 1, row, 0);
          load_chunk(w, 2, row, 4);
          if (e0 >= 4) load_chunk(w, 0, row, -4);
          if (e0 + 12 <= T2_BW) load_chunk(w, 3, row, 8);
        };
        int jg = (g.g0 + yb + y0) % g.gY;  // global row of tile row y0 (ghost rows are periodic images)
        if (jg < 0) jg += g.gY;
        auto step = [&](int yl, const int (&Wm)[16], int (&Wc)[16], const int (&Wp)[16]) {
          const unsigned long long grow = (unsigned long long)jg * rowsize + xg;
          const bool interior = x_interior && yl >= C::HY && yl < C::HY + T2_OH && yb + yl < Y;
          int32_t *own = tile + yl * T2_BW + e0;
          bool row_class = true;
          int P;
          if constexpr (LAT == SPK_SQ_4N) P = (c0 + jg) & 1;
          else if constexpr (LAT == SPK_SQ_8N) { P = c0 & 1; row_class = ((jg & 1) == ((c0 >> 1) & 1)); }
          else P = c0;  // tri: element offset 2*(i%2)+m inside each group of 4
          if (row_class) {
            if constexpr (LAT == SPK_TRI) {
              switch (P) {
                case 0: accepted += tile2d_strip<LAT, MODEL, 0>(a, c0, own, Wm, Wc, Wp, grow, xg, interior); break;
                case 1: accepted += tile2d_strip<LAT, MODEL, 1>(a, c0, own, Wm, Wc, Wp, grow, xg, interior); break;
                case 2: accepted += tile2d_strip<LAT, MODEL, 2>(a, c0, own, Wm, Wc, Wp, grow, xg, interior); break;
                default: accepted += tile2d_strip<LAT, MODEL, 3>(a, c0, own, Wm, Wc, Wp, grow, xg, interior); break;
              }
            } else {
              if (P) accepted += tile2d_strip<LAT, MODEL, 1>(a, c0, own, Wm, Wc, Wp, grow, xg, interior);
              else accepted += tile2d_strip<LAT, MODEL, 0>(a, c0, own, Wm, Wc, Wp, grow, xg, interior);
            }
          }
          if (++jg == g.gY) jg = 0;
        };
        // three register windows take turns as the rows above / at / below (no register copies)
        int A[16], B[16], Cw[16];
        load_row(A, y0 - 1);
        load_row(B, y0);
        for (int yl = y0; yl < y1; yl += 3) {
          load_row(Cw, yl + 1);
          step(yl, A, B, Cw);
          if (yl + 1 < y1) {
            load_row(A, yl + 2);
            step(yl + 1, B, Cw, A);
          }
          if (yl + 2 < y1) {
            load_row(B, yl + 3);
            step(yl + 2, Cw, A, B);
          }
        }
      }
      __syncthreads();
    }

    // store the interior: rows [HY, HY+OH) x elements [HX, HX+OW), 16-byte chunks
    constexpr int CPR = C::OW / 4;  // chunks per interior row
    for (int c = tid; c < CPR * T2_OH; c += 256) {
      const int r = c / CPR, q = c - r * CPR;
      const int yrow = yo + r, x = xo + 4 * q;
      if (yrow < Y && x < rowlen)
        *reinterpret_cast<int4 *>(dst + (long long)yrow * rowlen + x) =
            *reinterpret_cast<const int4 *>(tile + (C::HY + r) * T2_BW + C::HX + 4 * q);
    }
    __syncthreads();  // the tile buffer is reused by the next item
  }
  accepted = __reduce_add_sync(0xffffffffu, accepted);
  if (lane == 0 && accepted) atomicAdd(a.naccept, (unsigned long long)accepted);
}

}  // namespace spk
