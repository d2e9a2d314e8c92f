// Tuned colour-set kernel (Philox mode): "strip" kernel.
//
// Each thread owns a strip of 8 consecutive int32 elements of one colour-class row and updates
// the class sites inside it (4 for the one-site-per-cell lattices, 2 for tri).  For every stencil
// row it reads the strip with 128-bit loads plus the one 16-byte chunk that holds the x-halo, so a
// warp reads 1 KB contiguous pieces of each row; neighbour rows shared between the class rows of a
// block (RY rows x RZ planes per block) are served by L1.  Neighbour values are picked out of the
// register windows with compile-time indices, in reference neighbour order.  One Philox block is
// shared by the two class sites whose cell index differs in bit 1 (event.cuh).
//
// launch_tiled returns 0 when it launched, -1 when the geometry does not fit (caller falls back
// to sweep_generic_kernel), >0 on a CUDA error.
#pragma once
#include <utility>
#include "sweep_generic.cuh"

namespace spk {

// compile-time loop: f(std::integral_constant<int,0>{}), ..., f(std::integral_constant<int,N-1>{})
template <int... Is, class F>
__device__ __forceinline__ void static_for_impl(std::integer_sequence<int, Is...>, F &&f) {
  (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, class F> __device__ __forceinline__ void static_for(F &&f) {
  static_for_impl(std::make_integer_sequence<int, N>{}, f);
}

struct StripMap {
  int s_lo, s_hi;        // strip range in x (strip = 8 elements)
  int j0, jstep, nrows;  // class rows:   j = j0 + r*jstep, r < nrows
  int k0, kstep, nplanes;
  int ry, rz;            // class rows / planes per block (ry*rz = 8 warps)
};

// static description of one stencil row (dy,dz) for a class site of basis M
template <int LAT, int M> struct RowInfo {
  using L = Lat<LAT>;
  // element offset of neighbour n relative to the site's own element
  __host__ __device__ static constexpr int delta(int n) { return L::off(M, n).dx * L::NB + (L::off(M, n).m - M); }
  __host__ __device__ static constexpr bool used(int dy, int dz) {
    if (dy == 0 && dz == 0) return true;  // own row (self element)
    for (int n = 0; n < L::NN; n++)
      if (L::off(M, n).dy == dy && L::off(M, n).dz == dz) return true;
    return false;
  }
  __host__ __device__ static constexpr int dmin(int dy, int dz) {
    int d = 0;
    for (int n = 0; n < L::NN; n++)
      if (L::off(M, n).dy == dy && L::off(M, n).dz == dz && delta(n) < d) d = delta(n);
    return d;
  }
  __host__ __device__ static constexpr int dmax(int dy, int dz) {
    int d = 0;
    for (int n = 0; n < L::NN; n++)
      if (L::off(M, n).dy == dy && L::off(M, n).dz == dz && delta(n) > d) d = delta(n);
    return d;
  }
};

__device__ __forceinline__ void load_chunk(int (&w)[16], int c, const int32_t *row, long long e) {
  const int4 t = *reinterpret_cast<const int4 *>(row + e);
  w[4 * c + 0] = t.x; w[4 * c + 1] = t.y; w[4 * c + 2] = t.z; w[4 * c + 3] = t.w;
}

// P = element offset of the first class site inside the strip; class sites sit at P + S*s
template <int LAT, int MODEL, int P>
__device__ __forceinline__ int strip_body(const SweepArgs &a, int e0, int j, int k) {
  using L = Lat<LAT>;
  constexpr int NN = L::NN, NB = L::NB;
  constexpr int S = 2 * NB, NS = 8 / S;  // class-site stride and count inside the strip
  constexpr int M = (LAT == SPK_TRI) ? (P & 1) : 0;
  using R = RowInfo<LAT, M>;
  const Geo &g = a.g;
  const int rowlen = g.X * NB;

  int yj[3], zk[3];
  if constexpr (L::DIM == 2) {
    yj[0] = j == 0 ? g.lo_wrap : j - 1; yj[1] = j; yj[2] = j == g.Y - 1 ? g.hi_wrap : j + 1;
    zk[0] = zk[1] = zk[2] = 0;
  } else {
    yj[0] = j == 0 ? g.Y - 1 : j - 1; yj[1] = j; yj[2] = j == g.Y - 1 ? 0 : j + 1;
    zk[0] = k == 0 ? g.lo_wrap : k - 1; zk[1] = k; zk[2] = k == g.Z - 1 ? g.hi_wrap : k + 1;
  }
  const long long sy = rowlen, sz = sy * g.Y;
  // periodic x images of the two halo chunks
  const int eL = e0 == 0 ? rowlen - 4 : e0 - 4;
  const int eR = e0 + 8 == rowlen ? 0 : e0 + 8;

  // register windows: W[row][4 + p] = element e0 + p of stencil row `row`, p in [-4, 12);
  // which rows and which halo chunk are needed is decided at compile time
  int W[9][16];
  static_for<9>([&](auto RR) {
    constexpr int rr = decltype(RR)::value, dy = rr % 3 - 1, dz = rr / 3 - 1;
    constexpr bool used = (L::DIM == 3 || dz == 0) && R::used(dy, dz);
    if constexpr (used) {
      constexpr bool left = P + R::dmin(dy, dz) < 0, right = P + S * (NS - 1) + R::dmax(dy, dz) > 7;
      const int32_t *row = a.spin + zk[dz + 1] * sz + yj[dy + 1] * sy;
      load_chunk(W[rr], 1, row, e0);
      load_chunk(W[rr], 2, row, e0 + 4);
      if constexpr (left) load_chunk(W[rr], 0, row, eL);
      if constexpr (right) load_chunk(W[rr], 3, row, eR);
    }
  });

  // global index of the strip's first element keys the random stream
  long long jg = j, kg = k;
  if constexpr (L::DIM == 2) jg += g.g0; else kg += g.g0;
  const unsigned long long grow = (unsigned long long)(kg * g.gY + jg) * (unsigned long long)(g.gX * NB) + e0;

  int cur[NS], before[NS];
#pragma unroll
  for (int s = 0; s < NS; s++) before[s] = cur[s] = W[4][4 + P + S * s];

  int accepted = 0;
  for (int l = 0; l < a.nloop; l++) {
    // Philox blocks of the strip (event.cuh: philox_words): potts family one block per pair of
    // class sites, ising one block per four
    constexpr int SPB = ising_like(MODEL) ? 4 : 2;           // class sites per block
    constexpr int NBLK = (NS + SPB - 1) / SPB;                  // blocks the strip touches
    static_for<NBLK>([&](auto PR) {
      constexpr int pr = decltype(PR)::value;
      // strips are 8-element aligned; a tri strip (4 cells) is half of an ising block
      const bool upper = (ising_like(MODEL) && NS < SPB) ? ((e0 >> 3) & 1) : false;
      const unsigned long long leader = grow + P + S * (SPB * pr) - (upper ? 8 : 0);
      const uint4 blk = philox4x32_10((uint32_t)leader, (uint32_t)(leader >> 32), a.ev.sweep,
                                      (uint32_t)l * 16u + (uint32_t)(a.c0 + 1), a.ev.rk);
      static_for<(NS < SPB ? NS : SPB)>([&](auto H) {
        constexpr int h = decltype(H)::value, s = SPB * pr + h, p = P + S * s;
        int v[NN];
        static_for<NN>([&](auto N_) {
          constexpr int n = decltype(N_)::value;
          constexpr Off o = L::off(M, n);
          constexpr int row = (o.dz + 1) * 3 + (o.dy + 1), idx = 4 + p + R::delta(n);
          v[n] = W[row][idx];
        });
        Draws<RNG_PHILOX> d;
        if constexpr (ising_like(MODEL)) {
          const uint32_t lo = h == 0 ? blk.x : (h == 1 ? blk.y : (h == 2 ? blk.z : blk.w));
          const uint32_t hi = h == 0 ? blk.z : blk.w;  // (only h < 2 when the strip is the upper half)
          d.w1 = upper ? hi : lo;
          d.w2 = d.w1 << 1;
        } else {
          d.w1 = h ? blk.z : blk.x;
          d.w2 = h ? blk.w : blk.y;
        }
        const int nw = site_event<MODEL, RNG_PHILOX, NN>(a.ev, cur[s], v, d);
        accepted += (nw != cur[s]);
        cur[s] = nw;
      });
    });
  }

  // write back the 16-byte chunks that changed (class sites of one chunk travel together)
  int32_t *own = a.spin + zk[1] * sz + yj[1] * sy + e0;
  static_for<2>([&](auto C_) {
    constexpr int c = decltype(C_)::value;
    bool changed = false;
    static_for<NS>([&](auto S_) {
      constexpr int s = decltype(S_)::value, p = P + S * s;
      if constexpr (p / 4 == c) { changed |= (cur[s] != before[s]); W[4][4 + p] = cur[s]; }
    });
    if (changed)
      *reinterpret_cast<int4 *>(own + 4 * c) =
          make_int4(W[4][4 + 4 * c], W[4][5 + 4 * c], W[4][6 + 4 * c], W[4][7 + 4 * c]);
  });
  return accepted;
}

template <int LAT, int MODEL>
__global__ void __launch_bounds__(256, 2) sweep_strip_kernel(const __grid_constant__ SweepArgs a,
                                                             const __grid_constant__ StripMap sm) {
  const int lane = threadIdx.x, warp = threadIdx.y;
  const int sidx = sm.s_lo + blockIdx.x * 32 + lane;
  const int r = blockIdx.y * sm.ry + warp % sm.ry;
  const int q = blockIdx.z * sm.rz + warp / sm.ry;
  const int j = sm.j0 + r * sm.jstep, k = sm.k0 + q * sm.kstep;
  int accepted = 0;
  if (r < sm.nrows && q < sm.nplanes && sidx < sm.s_hi) {  // r,q are warp-uniform
    const int e0 = sidx * 8;
    if constexpr (LAT == SPK_TRI) {
      switch (a.c0) {  // (id-1)%4 = 2*(i%2) + m = element offset inside a group of 4
        case 0: accepted = strip_body<LAT, MODEL, 0>(a, e0, j, k); break;
        case 1: accepted = strip_body<LAT, MODEL, 1>(a, e0, j, k); break;
        case 2: accepted = strip_body<LAT, MODEL, 2>(a, e0, j, k); break;
        default: accepted = strip_body<LAT, MODEL, 3>(a, e0, j, k); break;
      }
    } else {
      int px;
      if constexpr (LAT == SPK_SQ_4N) px = (a.c0 + j + a.g.g0) & 1;
      else if constexpr (LAT == SPK_SC_6N) px = (a.c0 + j + k + a.g.g0) & 1;
      else px = a.c0 & 1;
      if (px) accepted = strip_body<LAT, MODEL, 1>(a, e0, j, k);
      else accepted = strip_body<LAT, MODEL, 0>(a, e0, j, k);
    }
  }
  accepted = __reduce_add_sync(0xffffffffu, accepted);
  if (lane == 0 && accepted) atomicAdd(a.naccept, (unsigned long long)accepted);
}

template <int LAT, int MODEL>
inline int launch_tiled(cudaStream_t st, const SweepArgs &a, int64_t *launches) {
  using L = Lat<LAT>;
  const Geo &g = a.g;
  const Box &b = a.b;
  const int rowlen = g.X * L::NB;
  if (rowlen % 8 || rowlen < 16) return -1;
  if ((b.ilo * L::NB) % 8 || (b.ihi * L::NB) % 8) return -1;  // x range of the box must be strip aligned
  StripMap sm;
  sm.s_lo = b.ilo * L::NB / 8;
  sm.s_hi = b.ihi * L::NB / 8;
  auto first = [](int lo, int p) { return lo + ((lo ^ p) & 1); };
  const int c0 = a.c0;
  sm.j0 = b.jlo; sm.jstep = 1; sm.nrows = b.jhi - b.jlo;
  sm.k0 = b.klo; sm.kstep = 1; sm.nplanes = b.khi - b.klo;
  if (LAT == SPK_SQ_8N || LAT == SPK_SC_26N) {
    sm.j0 = first(b.jlo, (c0 >> 1) & 1); sm.jstep = 2;
    sm.nrows = sm.j0 < b.jhi ? (b.jhi - sm.j0 + 1) / 2 : 0;
  }
  if (LAT == SPK_SC_26N) {
    sm.k0 = first(b.klo, (c0 >> 2) & 1); sm.kstep = 2;
    sm.nplanes = sm.k0 < b.khi ? (b.khi - sm.k0 + 1) / 2 : 0;
  }
  if (L::DIM == 2) { sm.k0 = 0; sm.kstep = 1; sm.nplanes = 1; }
  if (sm.nrows <= 0 || sm.nplanes <= 0 || sm.s_hi <= sm.s_lo) return 0;
  sm.ry = L::DIM == 2 ? 8 : 4;
  sm.rz = 8 / sm.ry;
  dim3 block(32, 8, 1);
  dim3 grid((sm.s_hi - sm.s_lo + 31) / 32, (sm.nrows + sm.ry - 1) / sm.ry, (sm.nplanes + sm.rz - 1) / sm.rz);
  if (grid.y > 65535 || grid.z > 65535) return -1;
  sweep_strip_kernel<LAT, MODEL><<<grid, block, 0, st>>>(a, sm);
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace spk
