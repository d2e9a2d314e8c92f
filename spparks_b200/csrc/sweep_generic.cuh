// Generic colour-set kernel: one thread per site of the colour class, neighbours gathered
// straight from global memory through L1.  Correct for every lattice x app x random source;
// the tuned kernels (sweep_tiled.cuh) replace it where they apply.
#pragma once
#include <type_traits>
#include "event.cuh"

namespace spk {

struct SweepArgs {
  int32_t *spin;       // first owned site (ghost plane/row lies at negative offsets)
  Geo g;
  Box b;
  int c0;              // colour - 1
  int nloop;           // passes over the set (src/app_lattice.cpp:674-676)
  EventParams ev;
  const double *u1, *u2;  // RNG_INJECTED: per owned site
  int32_t *seeds;         // RNG_STRICT: per owned site
  unsigned long long *naccept;
  unsigned char *mask;    // `sweep ... mask yes`: one byte per site, indexed like spin (null = off)
  unsigned long long *kstat;  // candidate-list apps: sum of nevent over the sites (streaming kernel; null = off)
  // peer-memory halo (streaming kernel, multi-GPU slabs): updates of local plane 0 / Z-1 are also
  // stored into the ring neighbour's ghost plane over NVLink (null = no mirror)
  int32_t *mirror_lo, *mirror_hi;
};

// colour class coordinates ------------------------------------------------------------------
// Each lattice's colour formula (src/lattice.cpp:234-283, delcolor = 1) fixes the parity of i
// (and for the 4- and 8-colourings of j, k or the basis m); (t,r,q) index the class inside box.
template <int LAT>
__device__ __forceinline__ bool class_site(const Geo &g, const Box &b, int c0, int t, int r, int q,
                                           int &i, int &j, int &k, int &m) {
  m = 0; k = 0;
  if constexpr (LAT == SPK_SQ_4N) {
    j = b.jlo + r;
    int px = (c0 + j + g.g0) & 1;               // (i + jglobal) % 2 == c0
    i = b.ilo + ((b.ilo ^ px) & 1) + 2 * t;
    return j < b.jhi && i < b.ihi;
  } else if constexpr (LAT == SPK_SQ_8N) {
    int px = c0 & 1, py = (c0 >> 1) & 1;        // 2*(j%2) + i%2 == c0 ; g0 is even
    j = b.jlo + ((b.jlo ^ py) & 1) + 2 * r;
    i = b.ilo + ((b.ilo ^ px) & 1) + 2 * t;
    return j < b.jhi && i < b.ihi;
  } else if constexpr (LAT == SPK_TRI) {
    int px = (c0 >> 1) & 1;                     // (id-1) % 4 = 2*(i%2) + m, nx even
    m = c0 & 1;
    j = b.jlo + r;
    i = b.ilo + ((b.ilo ^ px) & 1) + 2 * t;
    return j < b.jhi && i < b.ihi;
  } else if constexpr (LAT == SPK_SC_6N) {
    j = b.jlo + r;
    k = b.klo + q;
    int px = (c0 + j + k + g.g0) & 1;           // (i + j + kglobal) % 2 == c0
    i = b.ilo + ((b.ilo ^ px) & 1) + 2 * t;
    return k < b.khi && j < b.jhi && i < b.ihi;
  } else {
    int px = c0 & 1, py = (c0 >> 1) & 1, pz = (c0 >> 2) & 1;  // 4*(k%2) + 2*(j%2) + i%2 == c0
    k = b.klo + ((b.klo ^ pz) & 1) + 2 * q;
    j = b.jlo + ((b.jlo ^ py) & 1) + 2 * r;
    i = b.ilo + ((b.ilo ^ px) & 1) + 2 * t;
    return k < b.khi && j < b.jhi && i < b.ihi;
  }
}

template <int LAT, int MODEL, int RNG>
__global__ void __launch_bounds__(256) sweep_generic_kernel(const __grid_constant__ SweepArgs a) {
  using L = Lat<LAT>;
  constexpr int NN = L::NN, NB = L::NB;
  const Geo &g = a.g;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y * blockDim.y + threadIdx.y;
  const int q = blockIdx.z;
  int i, j, k, m;
  const bool live = class_site<LAT>(g, a.b, a.c0, t, r, q, i, j, k, m);
  int accepted = 0;

  if (live) {
    // periodic images in x (and y in 3-D); the slab dimension goes through lo_wrap / hi_wrap
    const int xi[3] = {i == 0 ? g.X - 1 : i - 1, i, i == g.X - 1 ? 0 : i + 1};
    int yj[3], zk[3];
    if constexpr (L::DIM == 2) {
      yj[0] = j == 0 ? g.lo_wrap : j - 1; yj[1] = j; yj[2] = j == g.Y - 1 ? g.hi_wrap : j + 1;
      zk[0] = zk[1] = zk[2] = 0;
    } else {
      yj[0] = j == 0 ? g.Y - 1 : j - 1; yj[1] = j; yj[2] = j == g.Y - 1 ? 0 : j + 1;
      zk[0] = k == 0 ? g.lo_wrap : k - 1; zk[1] = k; zk[2] = k == g.Z - 1 ? g.hi_wrap : k + 1;
    }
    const long long sy = (long long)g.X * NB, sz = sy * g.Y;
    const long long self = zk[1] * sz + yj[1] * sy + (long long)i * NB + m;

    // neighbour spins in reference order; the tri basis (uniform over the launch) picks the table
    int v[NN];
    auto gather = [&](auto mm) {
#pragma unroll
      for (int n = 0; n < NN; n++) {
        constexpr int M = decltype(mm)::value;
        const Off o = L::off(M, n);
        v[n] = a.spin[zk[o.dz + 1] * sz + yj[o.dy + 1] * sy + (long long)xi[o.dx + 1] * NB + o.m];
      }
    };
    if (LAT == SPK_TRI && m == 1) gather(std::integral_constant<int, 1>{});
    else gather(std::integral_constant<int, 0>{});

    int cur = a.spin[self];
    const int before = cur;
    // `sweep ... mask yes` (src/app_lattice.cpp:784-792, 809-819): a masked site is skipped; after an
    // event that ran to its end the site is masked if einitial < numneigh/2, and a change unmasks
    // every neighbour (src/app_potts.cpp:140-145).  Same-colour sites are not neighbours, so the
    // mask bytes a launch reads are not the ones it writes.
    bool masked = a.mask != nullptr && a.mask[self] != 0;
    auto mask_after = [&](int einit, bool changed) {
      if (a.mask == nullptr || einit < 0) return;
      if (2 * einit < NN) { a.mask[self] = 1; masked = true; }
      if (changed) {
        auto clear = [&](auto mm) {
#pragma unroll
          for (int n = 0; n < NN; n++) {
            constexpr int M = decltype(mm)::value;
            const Off o = L::off(M, n);
            a.mask[zk[o.dz + 1] * sz + yj[o.dy + 1] * sy + (long long)xi[o.dx + 1] * NB + o.m] = 0;
          }
        };
        if (LAT == SPK_TRI && m == 1) clear(std::integral_constant<int, 1>{});
        else clear(std::integral_constant<int, 0>{});
      }
    };
    // global site index keys the random stream, so any slab decomposition gives the same run
    long long jg = j, kg = k;
    if constexpr (L::DIM == 2) jg += g.g0; else kg += g.g0;
    const unsigned long long gsite = ((unsigned long long)(kg * g.gY + jg) * g.gX + i) * NB + m;
    const long long lsite = self;  // owned-site index for the per-site parity arrays

    if constexpr (RNG == RNG_STRICT) {
      Draws<RNG_STRICT> d;
      d.seed = a.seeds[lsite];
      for (int l = 0; l < a.nloop && !masked; l++) {
        int einit;
        int nw = site_event<MODEL, RNG, NN>(a.ev, cur, v, d, nullptr, &einit);
        accepted += (nw != cur);
        mask_after(einit, nw != cur);
        cur = nw;
      }
      a.seeds[lsite] = d.seed;
    } else if constexpr (RNG == RNG_INJECTED) {
      Draws<RNG_INJECTED> d;
      d.u1 = a.u1[lsite];
      d.u2 = a.u2[lsite];
      if (!masked) {
        int einit;
        int nw = site_event<MODEL, RNG, NN>(a.ev, cur, v, d, nullptr, &einit);
        accepted += (nw != cur);
        mask_after(einit, nw != cur);
        cur = nw;
      }
    } else {
      // block shared with the other class sites of the strip (event.cuh: philox_words)
      const unsigned long long leader = gsite - (unsigned long long)(philox_leader_shift<MODEL>(i) * NB);
      for (int l = 0; l < a.nloop && !masked; l++) {
        const uint4 blk = philox4x32_10((uint32_t)leader, (uint32_t)(leader >> 32), a.ev.sweep,
                                        (uint32_t)l * 16u + (uint32_t)(a.c0 + 1), a.ev.k0, a.ev.k1);
        Draws<RNG_PHILOX> d;
        philox_words<MODEL>(blk, i, d.w1, d.w2);
        int einit;
        int nw = site_event<MODEL, RNG, NN>(a.ev, cur, v, d, nullptr, &einit);
        accepted += (nw != cur);
        mask_after(einit, nw != cur);
        cur = nw;
      }
    }
    if (cur != before) a.spin[self] = cur;
  }

  // naccept (src/app_potts.cpp:133): one atomic per warp
  accepted = __reduce_add_sync(0xffffffffu, accepted);
  if (((threadIdx.y * blockDim.x + threadIdx.x) & 31) == 0 && accepted) atomicAdd(a.naccept, (unsigned long long)accepted);
}

}  // namespace spk
