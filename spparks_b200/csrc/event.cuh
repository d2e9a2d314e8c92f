// Per-site rejection event (site_event_rejection of the four apps) and the random sources.
#pragma once
#include <stdint.h>
#include "lattice.cuh"

namespace spk {

enum { RNG_PHILOX = 0, RNG_INJECTED = 1, RNG_STRICT = 2 };

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. SC'11).  counter = (site lo, site hi, sweep, iloop*16 + colour),
// key = seed.  Only words 0 and 1 of the output block are consumed (u1, u2).
// The key schedule is warp-uniform, so it folds into immediates/uniform registers.
__device__ __forceinline__ uint2 philox4x32_10_w01(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                   uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c2 = hi0 ^ c3 ^ k1;
    c1 = lo1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint2(c0, c1);
}

// RandomPark::uniform (src/random_park.cpp:79-86) as 16807*seed mod (2^31-1); the Schrage form
// of the reference and this 64-bit product give the same residue for seed in [1, 2^31-2].
__device__ __forceinline__ double park_uniform(int &seed) {
  uint64_t p = (uint64_t)16807u * (uint32_t)seed;
  uint32_t s = (uint32_t)(p & 0x7FFFFFFFu) + (uint32_t)(p >> 31);
  if (s >= 0x7FFFFFFFu) s -= 0x7FFFFFFFu;
  seed = (int)s;
  return (1.0 / 2147483647) * (double)seed;
}

// Everything an event needs besides the spins; lives in kernel parameters (constant bank).
struct EventParams {
  int nspins;               // Q
  int t_zero;               // temperature == 0.0
  uint32_t k0, k1;          // Philox key (seed)
  uint32_t sweep;           // Philox counter word 2
  uint32_t thr[27];         // Philox mode: floor(exp(-dE/T) * 2^32), dE = 1..26 (host libm)
  double ptab[27];          // double modes: exp((double)(-dE) * t_inverse)       (host libm)
};

// Source of the (at most) two uniforms of one event; drawn lazily like random->uniform().
template <int RNG> struct Draws;

template <> struct Draws<RNG_PHILOX> {
  uint2 w;
  __device__ __forceinline__ Draws(const EventParams &p, uint64_t gsite, uint32_t c3) {
    w = philox4x32_10_w01((uint32_t)gsite, (uint32_t)(gsite >> 32), p.sweep, c3, p.k0, p.k1);
  }
};
template <> struct Draws<RNG_INJECTED> {
  double u1, u2;
};
template <> struct Draws<RNG_STRICT> {
  int seed;
};

// number of neighbours unlike `value`  (site_energy, src/app_potts.cpp:97-104)
template <int NN> __device__ __forceinline__ int unlike(const int (&v)[NN], int value) {
  int e = 0;
#pragma unroll
  for (int n = 0; n < NN; n++) e += (v[n] != value);
  return e;
}

// One site_event_rejection.  `old` = spin[i], v[] = neighbour spins in reference order.
// Returns the new spin (== old when rejected or no event).
template <int MODEL, int RNG, int NN>
__device__ __forceinline__ int site_event(const EventParams &p, int old, const int (&v)[NN], Draws<RNG> &d) {
  const int Q = p.nspins;
  if (MODEL == SPK_POTTS_PIN && old > Q) return old;  // pinned: no draw (src/app_potts_pin.cpp:97)

  const int e0 = unlike<NN>(v, old);
  int trial;

  if constexpr (MODEL == SPK_POTTS_NEIGHONLY) {
    // distinct neighbour values != old in first-occurrence order (src/app_potts_neighonly.cpp:81-88)
    uint32_t first = 0;
#pragma unroll
    for (int n = 0; n < NN; n++) {
      bool dup = (v[n] == old);
#pragma unroll
      for (int q = 0; q < n; q++) dup = dup || (v[q] == v[n]);
      first |= dup ? 0u : (1u << n);
    }
    const int nevent = __popc(first);
    if (nevent == 0) return old;  // no draw (src/app_potts_neighonly.cpp:90)
    int pick;
    if constexpr (RNG == RNG_PHILOX) {
      pick = (int)__umulhi((uint32_t)nevent, d.w.x);
    } else {
      double u;
      if constexpr (RNG == RNG_INJECTED) u = d.u1;
      else u = park_uniform(d.seed);
      pick = (int)(nevent * u);
      if (pick >= nevent) pick = nevent - 1;
    }
    // value of the pick-th set bit of `first`
    trial = old;
    int seen = 0;
#pragma unroll
    for (int n = 0; n < NN; n++) {
      bool f = (first >> n) & 1u;
      if (f && seen == pick) trial = v[n];
      seen += f;
    }
  } else if constexpr (RNG == RNG_PHILOX) {
    const uint32_t w0 = d.w.x;
    if (MODEL == SPK_ISING) trial = (w0 < 0x80000000u) ? 1 : 2;
    else trial = (int)__umulhi((uint32_t)Q, w0) + 1;
  } else {
    double u;
    if constexpr (RNG == RNG_INJECTED) u = d.u1;
    else u = park_uniform(d.seed);
    if (MODEL == SPK_ISING) trial = (u < 0.5) ? 1 : 2;  // src/app_ising.cpp:111-112
    else {
      trial = (int)(Q * u) + 1;  // src/app_potts.cpp:119-120
      if (trial > Q) trial = Q;
    }
  }

  const int e1 = unlike<NN>(v, trial);
  if (e1 <= e0) return trial;
  if (p.t_zero) return old;
  const int dE = e1 - e0;
  bool reject;
  if constexpr (RNG == RNG_PHILOX) {
    reject = d.w.y > p.thr[dE];
  } else {
    double u;
    if constexpr (RNG == RNG_INJECTED) u = d.u2;
    else u = park_uniform(d.seed);
    reject = u > p.ptab[dE];  // src/app_potts.cpp:130
  }
  return reject ? old : trial;
}

}  // namespace spk
