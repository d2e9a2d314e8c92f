// diag_style cluster on the device: connected components of equal-spin sites
// over the lattice's neighbour stencil, what DiagCluster::analyze_clusters does with a serial stack
// flood fill (src/diag_cluster.cpp:286-354, src/app_lattice.cpp:1170-1201).
//
// Lock-free union-find on a label array (label[i] = parent site index, roots point to themselves;
// a union always hangs the larger root under the smaller one with atomicMin, so a cluster's root
// is its lowest site index — the site where the reference's scan opens the cluster, which makes the
// reference's numbering "k-th cluster = k-th lowest root").  Every undirected edge is visited once
// (from its lower end).  Volumes are counted per root with warp-aggregated atomics; the ordered
// volume list and the per-site cluster ids come from an exclusive scan over the root flags.
#pragma once
#include <stdint.h>
#include <type_traits>
#include "lattice.cuh"

namespace spk {

// root of x, halving the path on the way (label[x] <- grandparent: a plain racy store that can only
// replace a parent by one of its own ancestors, the intermediate pointer jumping of ECL-CC)
__device__ __forceinline__ uint32_t cl_find(uint32_t *label, uint32_t x) {
  volatile uint32_t *L = label;
  uint32_t p = L[x];
  while (p != x) {
    const uint32_t gp = L[p];
    if (gp == p) return p;
    L[x] = gp;
    x = p;
    p = gp;
  }
  return x;
}

// root of x without touching the array (flatten pass: a halving store could undo another thread's
// final label[i] = root)
__device__ __forceinline__ uint32_t cl_find_ro(const uint32_t *label, uint32_t x) {
  const volatile uint32_t *L = label;
  for (;;) {
    const uint32_t p = L[x];
    if (p == x) return x;
    x = p;
  }
}

__device__ __forceinline__ void cl_unite(uint32_t *label, uint32_t a, uint32_t b) {
  for (;;) {
    a = cl_find(label, a);
    b = cl_find(label, b);
    if (a == b) return;
    if (a < b) { const uint32_t t = a; a = b; b = t; }
    const uint32_t old = atomicMin(label + a, b);  // a is (was) a root: hang it under the smaller root
    if (old == a) return;
    a = old;  // another thread re-parented a in between: carry on from its new parent
  }
}

__global__ void __launch_bounds__(256) cluster_init_kernel(uint32_t *label, uint32_t *size, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    label[i] = (uint32_t)i;
    size[i] = 0u;
  }
}

// one thread per site: join with every equal-spin neighbour of higher index (fully periodic lattice)
// slab_open: the last dimension is cut into slabs (multi-GPU): no periodic wrap there, the
// links across slab boundaries are made by cluster_boundary_kernel
template <int LAT>
__global__ void __launch_bounds__(256) cluster_link_kernel(const int32_t *__restrict__ spin, uint32_t *label, int X, int Y,
                                                           int Z, int slab_open) {
  using L = Lat<LAT>;
  constexpr int NN = L::NN, NB = L::NB;
  const long long n = (long long)X * Y * Z * NB;
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(s % NB);
    long long c = s / NB;
    const int i = (int)(c % X);
    c /= X;
    const int j = (int)(c % Y), k = (int)(c / Y);
    const int mine = spin[s];
    auto visit = [&](auto MM) {
      constexpr int M = decltype(MM)::value;
#pragma unroll
      for (int q = 0; q < NN; q++) {
        const Off o = L::off(M, q);
        int ii = i + o.dx, jj = j + o.dy, kk = k + o.dz;
        if (slab_open && (L::DIM == 3 ? (kk < 0 || kk >= Z) : (jj < 0 || jj >= Y))) continue;
        ii = ii < 0 ? ii + X : (ii >= X ? ii - X : ii);
        jj = jj < 0 ? jj + Y : (jj >= Y ? jj - Y : jj);
        kk = kk < 0 ? kk + Z : (kk >= Z ? kk - Z : kk);
        const long long t = (((long long)kk * Y + jj) * X + ii) * NB + o.m;
        if (t > s && spin[t] == mine) cl_unite(label, (uint32_t)s, (uint32_t)t);
      }
    };
    if (NB == 2 && m == 1) visit(std::integral_constant<int, 1>{});
    else visit(std::integral_constant<int, 0>{});
  }
}

// label[i] <- root(i); volume of every root, one atomic per distinct root and warp
__global__ void __launch_bounds__(256) cluster_flatten_kernel(uint32_t *label, uint32_t *size, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long nround = (n + stride - 1) / stride * stride;  // whole warps stay converged for match_any
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
    const bool live = i < n;
    uint32_t r = 0xFFFFFFFFu;
    if (live) {
      r = cl_find_ro(label, (uint32_t)i);
      label[i] = r;  // only ever shortens a path to the same root: safe while other threads still search
    }
    const unsigned peers = __match_any_sync(0xffffffffu, r);
    if (live && (__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(size + r, (uint32_t)__popc(peers));
  }
}

// ncluster and sum of volume^(1/dim) over the roots (DiagCluster stats, src/diag_cluster.cpp:595-605);
// out[0] = ncluster, rsum = sum of radii
__global__ void __launch_bounds__(256) cluster_stats_kernel(const uint32_t *__restrict__ label, const uint32_t *__restrict__ size,
                                                            long long n, double invdim, unsigned long long *ncluster,
                                                            double *rsum) {
  unsigned long long cnt = 0;
  double r = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (label[i] == (uint32_t)i) {
      cnt++;
      r += pow((double)size[i], invdim);
    }
  for (int o = 16; o; o >>= 1) {
    cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    r += __shfl_down_sync(0xffffffffu, r, o);
  }
  __shared__ unsigned long long scnt[8];
  __shared__ double sr[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { scnt[warp] = cnt; sr[warp] = r; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++) { cnt += scnt[w]; r += sr[w]; }
    if (cnt) { atomicAdd(ncluster, cnt); atomicAdd(rsum, r); }
  }
}

// ---- exclusive scan over the root flags: reference numbering and the ordered volume list ------
constexpr int CL_ITEMS = 16;                  // sites per thread
constexpr int CL_TILE = 256 * CL_ITEMS;       // sites per block

__global__ void __launch_bounds__(256) cluster_scan_blocksums(const uint32_t *__restrict__ label, long long n, uint32_t *blocksum) {
  const long long base = (long long)blockIdx.x * CL_TILE + (long long)threadIdx.x * CL_ITEMS;
  uint32_t c = 0;
#pragma unroll
  for (int e = 0; e < CL_ITEMS; e++) {
    const long long i = base + e;
    if (i < n && label[i] == (uint32_t)i) c++;
  }
  for (int o = 16; o; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  __shared__ uint32_t s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++) c += s[w];
    blocksum[blockIdx.x] = c;
  }
}

// single block: blocksum[] -> exclusive prefix (in place)
__global__ void __launch_bounds__(1024) cluster_scan_offsets(uint32_t *blocksum, int nb) {
  __shared__ uint32_t part[1024];
  const int per = (nb + 1023) / 1024, lo = threadIdx.x * per, hi = min(lo + per, nb);
  uint32_t sum = 0;
  for (int b = lo; b < hi; b++) sum += blocksum[b];
  part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int t = 0; t < 1024; t++) { const uint32_t v = part[t]; part[t] = run; run += v; }
  }
  __syncthreads();
  uint32_t run = part[threadIdx.x];
  for (int b = lo; b < hi; b++) { const uint32_t v = blocksum[b]; blocksum[b] = run; run += v; }
}

// roots in index order: volumes[rank] = size[root]; size[root] <- rank + 1 (the cluster's id)
__global__ void __launch_bounds__(256) cluster_scan_apply(const uint32_t *__restrict__ label, uint32_t *size, long long n,
                                                          const uint32_t *__restrict__ blockoff, long long *volumes) {
  const long long base = (long long)blockIdx.x * CL_TILE + (long long)threadIdx.x * CL_ITEMS;
  uint32_t c = 0;
#pragma unroll
  for (int e = 0; e < CL_ITEMS; e++) {
    const long long i = base + e;
    if (i < n && label[i] == (uint32_t)i) c++;
  }
  // exclusive scan of c over the block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = c;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  __shared__ uint32_t wsum[8];
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  uint32_t woff = 0;
  for (int w = 0; w < warp; w++) woff += wsum[w];
  uint32_t rank = blockoff[blockIdx.x] + woff + inc - c;
#pragma unroll
  for (int e = 0; e < CL_ITEMS; e++) {
    const long long i = base + e;
    if (i < n && label[i] == (uint32_t)i) {
      volumes[rank] = (long long)size[i];
      size[i] = rank + 1;
      rank++;
    }
  }
}

// ids[i] = id of the cluster of site i (reference numbering, 1-based)
// (ids may alias label: every thread reads and writes its own element only)
__global__ void __launch_bounds__(256) cluster_ids_kernel(const uint32_t *label, const uint32_t *__restrict__ size,
                                                          long long n, int32_t *ids) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    ids[i] = (int32_t)size[label[i]];
}

template <int LAT>
inline void launch_cluster_link(cudaStream_t st, const int32_t *spin, uint32_t *label, int X, int Y, int Z, int slab_open) {
  cluster_link_kernel<LAT><<<148 * 16, 256, 0, st>>>(spin, label, X, Y, Z, slab_open);
}

// ---- clusters across slabs (multi-GPU; DiagCluster merges the per-processor clusters through
// their boundary sites, src/diag_cluster.cpp:259-285, 372-396) ------------------------------------
// Every local root carries a 64-bit GLOBAL label, initially its own global site index; the label
// of a cluster is the minimum over the roots it joins across slab boundaries.  A round: pack the
// labels of the first / last owned layer, exchange them with the ring neighbours, and lower the
// label of every root that has an equal-spin neighbour in the adjacent ghost layer with a smaller
// label.  Rounds repeat until no label changed on any rank.

// glabel[i] = global index of i (meaningful at roots), gsize[i] = local volume (at roots)
__global__ void __launch_bounds__(256) cluster_global_init_kernel(const uint32_t *__restrict__ label, const uint32_t *__restrict__ size,
                                                                  unsigned long long *glabel, unsigned long long *gsize,
                                                                  long long n, unsigned long long base) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    glabel[i] = base + (unsigned long long)i;
    gsize[i] = label[i] == (uint32_t)i ? (unsigned long long)size[i] : 0ull;
  }
}

// labels of the first / last owned layer (layer = plane in 3-D, row in 2-D; nl sites each)
__global__ void __launch_bounds__(256) cluster_pack_kernel(const uint32_t *__restrict__ label, const unsigned long long *__restrict__ glabel,
                                                           long long n, long long nl, unsigned long long *out_lo,
                                                           unsigned long long *out_hi) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < nl; p += (long long)gridDim.x * blockDim.x) {
    out_lo[p] = glabel[label[p]];
    out_hi[p] = glabel[label[n - nl + p]];
  }
}

// spin: first owned site (the adjacent ghost layers lie at [-nl, 0) and [n, n + nl))
template <int LAT>
__global__ void __launch_bounds__(256) cluster_boundary_kernel(const int32_t *__restrict__ spin, const uint32_t *__restrict__ label,
                                                               unsigned long long *glabel, const unsigned long long *__restrict__ ghost_lo,
                                                               const unsigned long long *__restrict__ ghost_hi, int X, int Y, int Z,
                                                               int *changed) {
  using L = Lat<LAT>;
  constexpr int NN = L::NN, NB = L::NB;
  const long long nl = L::DIM == 3 ? (long long)X * Y * NB : (long long)X * NB;
  const long long n = (long long)X * Y * Z * NB;
  int any = 0;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < 2 * nl; q += (long long)gridDim.x * blockDim.x) {
    const int side = q >= nl;                 // 0: first layer looks down, 1: last layer looks up
    const long long p = side ? q - nl : q;    // index inside the layer
    const long long s = side ? n - nl + p : p;
    const int m = (int)(p % NB);
    const long long c = p / NB;
    const int i = (int)(c % X), j = L::DIM == 3 ? (int)(c / X) : 0;
    const int mine = spin[s];
    unsigned long long best = ~0ull;
    auto visit = [&](auto MM) {
      constexpr int M = decltype(MM)::value;
#pragma unroll
      for (int t = 0; t < NN; t++) {
        const Off o = L::off(M, t);
        const int ds = L::DIM == 3 ? o.dz : o.dy;  // step along the slab dimension
        if (ds != (side ? 1 : -1)) continue;
        int ii = i + o.dx, jj = j + (L::DIM == 3 ? o.dy : 0);
        ii = ii < 0 ? ii + X : (ii >= X ? ii - X : ii);
        if (L::DIM == 3) jj = jj < 0 ? jj + Y : (jj >= Y ? jj - Y : jj);
        const long long g = ((long long)jj * X + ii) * NB + o.m;  // index inside the ghost layer
        const int theirs = side ? spin[n + g] : spin[g - nl];
        if (theirs == mine) {
          const unsigned long long gl = side ? ghost_hi[g] : ghost_lo[g];
          best = gl < best ? gl : best;
        }
      }
    };
    if (NB == 2 && m == 1) visit(std::integral_constant<int, 1>{});
    else visit(std::integral_constant<int, 0>{});
    if (best != ~0ull) {
      const unsigned long long old = atomicMin(glabel + label[s], best);
      any |= best < old;
    }
  }
  if (any) *changed = 1;
}

// roots whose cluster is represented by another root (glabel != own index): (glabel, volume) pairs
__global__ void __launch_bounds__(256) cluster_foreign_kernel(const uint32_t *__restrict__ label, const unsigned long long *__restrict__ glabel,
                                                              const unsigned long long *__restrict__ gsize, long long n,
                                                              unsigned long long base, unsigned long long *pairs, int *count,
                                                              int capacity) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (label[i] == (uint32_t)i && glabel[i] != base + (unsigned long long)i) {
      const int slot = atomicAdd(count, 1);
      if (slot < capacity) { pairs[2 * slot] = glabel[i]; pairs[2 * slot + 1] = gsize[i]; }
    }
}

// add the volumes of the gathered foreign roots whose representative lives in this slab
__global__ void __launch_bounds__(256) cluster_absorb_kernel(const unsigned long long *__restrict__ pairs, long long npairs,
                                                             unsigned long long base, long long n, unsigned long long *gsize) {
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < npairs; q += (long long)gridDim.x * blockDim.x) {
    const unsigned long long gl = pairs[2 * q];
    if (gl != ~0ull && gl >= base && gl < base + (unsigned long long)n) atomicAdd(gsize + (gl - base), pairs[2 * q + 1]);
  }
}

// clusters represented in this slab and their radius sum
__global__ void __launch_bounds__(256) cluster_global_stats_kernel(const uint32_t *__restrict__ label, const unsigned long long *__restrict__ glabel,
                                                                   const unsigned long long *__restrict__ gsize, long long n,
                                                                   unsigned long long base, double invdim, unsigned long long *ncluster,
                                                                   double *rsum) {
  unsigned long long cnt = 0;
  double r = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (label[i] == (uint32_t)i && glabel[i] == base + (unsigned long long)i) {
      cnt++;
      r += pow((double)gsize[i], invdim);
    }
  for (int o = 16; o; o >>= 1) {
    cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    r += __shfl_down_sync(0xffffffffu, r, o);
  }
  if ((threadIdx.x & 31) == 0 && cnt) { atomicAdd(ncluster, cnt); atomicAdd(rsum, r); }
}

}  // namespace spk
