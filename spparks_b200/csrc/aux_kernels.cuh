// Device-side helpers around the sweep: synthetic initialisation, pinning, energy reduction.
#pragma once
#include "event.cuh"

namespace spk {

// `set site range lo hi` distribution (src/set.cpp:617-627) from a counter-based stream keyed
// on the GLOBAL site index: value = lo + floor(range * w/2^32).
__global__ void randomize_kernel(int32_t *spin, long long n, long long goff, unsigned long long seed, int lo,
                                 uint32_t range) {
  for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < n; s += (long long)gridDim.x * blockDim.x) {
    unsigned long long gs = (unsigned long long)(s + goff);
    uint4 w = philox4x32_10((uint32_t)gs, (uint32_t)(gs >> 32), 0x5e7517e5u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
    spin[s] = lo + (int)__umulhi(range, w.x);
  }
}

// `set site range lo hi` of the reference, bit for bit (src/set.cpp:582-627, default `loop all`, no
// region / fraction / condition): ONE Park-Miller stream (src/random_park.cpp:79-97) is walked over
// the global ids 1..N, so the site with global index s (0-based) takes draw number s+1:
//   state = seed0 * 16807^(s+1) mod (2^31-1),  value = lo + min((int)(state*AM*range), range-1).
// A warp covers 32 consecutive sites per step; each lane jumps to its first state with a modular
// power and then advances by 32 draws at a time (multiplier 16807^32 mod (2^31-1)).
__device__ __forceinline__ uint32_t pm_mulmod(uint32_t a, uint32_t b) {  // a*b mod 2^31-1, a,b < 2^31-1
  const uint64_t p = (uint64_t)a * b;
  uint32_t r = (uint32_t)(p & 0x7FFFFFFFu) + (uint32_t)(p >> 31);
  r = (r & 0x7FFFFFFFu) + (r >> 31);
  return r == 0x7FFFFFFFu ? 0u : r;
}
__device__ __forceinline__ uint32_t pm_powmod(uint32_t base, unsigned long long e) {
  uint32_t r = 1u;
  while (e) {
    if (e & 1ull) r = pm_mulmod(r, base);
    base = pm_mulmod(base, base);
    e >>= 1;
  }
  return r;
}
__global__ void __launch_bounds__(256) set_range_kernel(int32_t *spin, long long n, long long goff, uint32_t seed0, int lo,
                                                        int range) {
  const uint32_t step32 = pm_powmod(16807u, 32ull);
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long per = ((n + 31) / 32 + nwarps - 1) / nwarps * 32;  // sites per warp, multiple of 32
  const long long first = warp * per, last = first + per < n ? first + per : n;
  if (first >= n) return;
  uint32_t state = pm_mulmod(seed0, pm_powmod(16807u, (unsigned long long)(first + lane + goff) + 1ull));
  for (long long s = first + lane; s < last; s += 32) {
    const double u = (1.0 / 2147483647) * (double)(int)state;  // RandomPark::uniform: AM*seed
    int i = (int)(u * range) + 1;                               // irandom
    if (i > range) i = range;
    spin[s] = i - 1 + lo;
    state = pm_mulmod(state, step32);
  }
}

// Bernoulli(fraction) pinning, the large-lattice stand-in for `pin f 0 0`
// (src/app_potts_pin.cpp:262-283 draws ndesired random sites; here each site is pinned
// independently with the same probability).  Pinned value = Q+1 (src/app_potts_pin.cpp:271).
__global__ void pin_kernel(int32_t *spin, long long n, long long goff, unsigned long long seed, uint32_t thr,
                           int pinned) {
  for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < n; s += (long long)gridDim.x * blockDim.x) {
    unsigned long long gs = (unsigned long long)(s + goff);
    uint4 w = philox4x32_10((uint32_t)gs, (uint32_t)(gs >> 32), 0x9199ed51u, 1u, (uint32_t)seed, (uint32_t)(seed >> 32));
    if (w.x < thr) spin[s] = pinned;
  }
}

// `pin fraction 0 0` of app_style potts/pin, bit for bit (src/app_potts_pin.cpp:262-283, multi = 0,
// nthresh = 0): every attempt draws a global site id from ONE Park-Miller stream
// (tagrandom: (tagint)(uniform()*nglobal) + 1, src/random_park.cpp:103-108) and pins it; pinning is
// idempotent, so the attempts of a round commute and draw number d of the stream (0-based over
// the whole command) is evaluated on its own with a modular power, like set_range_kernel.  Every
// rank walks the same draws and keeps the ids that fall into its slab.
__global__ void __launch_bounds__(256) pin_draws_kernel(int32_t *spin, long long n, long long goff, uint32_t seed0,
                                                        unsigned long long first_draw, long long ndraws, long long nglobal,
                                                        int pinned) {
  const uint32_t step32 = pm_powmod(16807u, 32ull);
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long per = ((ndraws + 31) / 32 + nwarps - 1) / nwarps * 32;  // draws per warp, multiple of 32
  const long long first = warp * per, last = first + per < ndraws ? first + per : ndraws;
  if (first >= ndraws) return;
  uint32_t state = pm_mulmod(seed0, pm_powmod(16807u, first_draw + (unsigned long long)(first + lane) + 1ull));
  for (long long d = first + lane; d < last; d += 32) {
    const double u = (1.0 / 2147483647) * (double)(int)state;  // RandomPark::uniform: AM*seed
    long long id = (long long)(u * (double)nglobal) + 1;      // tagrandom
    if (id > nglobal) id = nglobal;
    const long long s = id - 1 - goff;
    if (s >= 0 && s < n) spin[s] = pinned;
    state = pm_mulmod(state, step32);
  }
}

// The pin modes whose attempts do NOT commute — nthresh > 0 (a site is pinned only if at least
// nthresh neighbours differ from it, and an earlier pin of the same round changes that count,
// src/app_potts_pin.cpp:268-274) and multi = 1 (a site and all its neighbours, only where none of
// them is pinned yet, :291-335) — walked by ONE thread in the reference's order: exact, one-shot,
// slow by construction (a dependent chain of random accesses).  One GPU only.
// out[0] = pinned sites added, out[1] = draws consumed, out[2] = 1 if the draw budget ran out
// (multi: no site satisfies the conditions any more; the reference would loop forever).
template <int LAT>
__global__ void pin_sequential_kernel(int32_t *spin, Geo g, uint32_t seed0, unsigned long long first_draw, long long ndraws,
                                      long long ndesired, long long already, int multi, int nthresh, int Q,
                                      unsigned long long *out) {
  using L = Lat<LAT>;
  constexpr int NN = L::NN, NB = L::NB;
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const long long nglobal = (long long)g.X * g.Y * g.Z * NB;
  const long long sy = (long long)g.X * NB, sz = sy * g.Y;
  uint32_t state = pm_mulmod(seed0, pm_powmod(16807u, first_draw));
  long long added = 0, draws = 0, npin = 0;  // npin: the reference's variable (multi: 0 until the first insertion)
  bool exhausted = false;
  // multi = 0: exactly ndraws attempts (one round, the host recounts between rounds); multi = 1: until npin >= ndesired
  while (multi ? npin < ndesired : draws < ndraws) {
    if (multi && draws >= ndraws) { exhausted = true; break; }
    state = pm_mulmod(state, 16807u);
    draws++;
    const double u = (1.0 / 2147483647) * (double)(int)state;
    long long id = (long long)(u * (double)nglobal) + 1;
    if (id > nglobal) id = nglobal;
    const long long s = id - 1;
    const int m = (int)(s % NB);
    const long long cell = s / NB;
    const int i = (int)(cell % g.X), j = (int)((cell / g.X) % g.Y), k = (int)(cell / ((long long)g.X * g.Y));
    long long nb[NN];
#pragma unroll
    for (int n = 0; n < NN; n++) {
      const Off o = L::off(m, n);
      int ii = i + o.dx, jj = j + o.dy, kk = k + o.dz;
      ii = ii < 0 ? ii + g.X : (ii >= g.X ? ii - g.X : ii);
      jj = jj < 0 ? jj + g.Y : (jj >= g.Y ? jj - g.Y : jj);
      kk = kk < 0 ? kk + g.Z : (kk >= g.Z ? kk - g.Z : kk);
      nb[n] = kk * sz + jj * sy + (long long)ii * NB + o.m;
    }
    const int me = spin[s];
    int ndiff = 0;
    bool free_ = me <= Q;
    for (int n = 0; n < NN; n++) {
      const int v = spin[nb[n]];
      ndiff += v != me;
      free_ = free_ && v <= Q;
    }
    if (!multi) {
      if (ndiff >= nthresh && me <= Q) added++;
      if (ndiff >= nthresh) spin[s] = Q + 1;
    } else if (free_ && (nthresh == 0 || ndiff >= nthresh)) {
      spin[s] = Q + 1;
      added++;
      for (int n = 0; n < NN; n++)
        if (spin[nb[n]] <= Q) { spin[nb[n]] = Q + 1; added++; }  // (periodic images on tiny lattices repeat a neighbour)
      npin = already + added;
    }
  }
  out[0] = (unsigned long long)added;
  out[1] = (unsigned long long)draws;
  out[2] = exhausted ? 1ull : 0ull;
}

// init_app's spin check (src/app_potts.cpp:83-90, src/app_ising.cpp:74-81, src/app_potts_pin.cpp:78-84):
// number of sites whose value lies outside [lo, hi]
__global__ void __launch_bounds__(256) validate_kernel(const int32_t *__restrict__ spin, long long n, int lo, int hi,
                                                       unsigned long long *bad) {
  unsigned int c = 0;
  for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < n; s += (long long)gridDim.x * blockDim.x) {
    const int v = spin[s];
    c += (v < lo || v > hi);
  }
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(bad, (unsigned long long)c);
}

// AppLattice::boundary_clear_mask (src/app_lattice.cpp:1158-1164) for a loop set given as a box
// of local cells: the border sites found by find_border_sites (:1092-1148, one layer since
// delpropensity + delevent = 1) are the sites of the set with at least one neighbour outside it.
// color0 >= 0 restricts the set to one colour class (a colour set without sectors): every
// neighbour of such a site has another colour, so the whole class is border.
// A neighbour beyond the slab (ghost plane / row of a multi-GPU run) is outside every set.
template <int LAT>
__global__ void __launch_bounds__(256) mask_border_kernel(unsigned char *mask, Geo g, Box b, int multi) {
  using L = Lat<LAT>;
  constexpr int NN = L::NN, NB = L::NB;
  const long long ni = b.ihi - b.ilo, nj = b.jhi - b.jlo, nk = b.khi - b.klo;
  const long long ncell = ni * nj * nk;
  const long long sy = (long long)g.X * NB, sz = sy * g.Y;
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < ncell; c += (long long)gridDim.x * blockDim.x) {
    const int i = b.ilo + (int)(c % ni), j = b.jlo + (int)((c / ni) % nj), k = b.klo + (int)(c / (ni * nj));
#pragma unroll
    for (int m = 0; m < NB; m++) {
      bool border = false;
#pragma unroll
      for (int n = 0; n < NN; n++) {
        const Off o = L::off(m, n);
        int ii = i + o.dx, jj = j + o.dy, kk = k + o.dz;
        ii = ii < 0 ? ii + g.X : (ii >= g.X ? ii - g.X : ii);  // x is never the slab dimension
        bool ghost = false;
        if (L::DIM == 2) {
          if (jj < 0 || jj >= g.Y) { if (multi) ghost = true; else jj = jj < 0 ? jj + g.Y : jj - g.Y; }
        } else {
          jj = jj < 0 ? jj + g.Y : (jj >= g.Y ? jj - g.Y : jj);
          if (kk < 0 || kk >= g.Z) { if (multi) ghost = true; else kk = kk < 0 ? kk + g.Z : kk - g.Z; }
        }
        border |= ghost || ii < b.ilo || ii >= b.ihi || jj < b.jlo || jj >= b.jhi || kk < b.klo || kk >= b.khi;
      }
      if (border) mask[k * sz + j * sy + (long long)i * NB + m] = 0;
    }
  }
}

// the sites of colour c0 inside the box (a colour set as loop set: all of it is border)
template <int LAT>
__global__ void __launch_bounds__(256) mask_color_clear_kernel(unsigned char *mask, Geo g, Box b, int c0) {
  using L = Lat<LAT>;
  constexpr int NB = L::NB;
  const long long ni = b.ihi - b.ilo, nj = b.jhi - b.jlo, nk = b.khi - b.klo;
  const long long ncell = ni * nj * nk;
  const long long sy = (long long)g.X * NB, sz = sy * g.Y;
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < ncell; c += (long long)gridDim.x * blockDim.x) {
    const int i = b.ilo + (int)(c % ni), j = b.jlo + (int)((c / ni) % nj), k = b.klo + (int)(c / (ni * nj));
    const int jg = L::DIM == 2 ? j + g.g0 : j, kg = L::DIM == 3 ? k + g.g0 : 0;
#pragma unroll
    for (int m = 0; m < NB; m++) {
      int col;  // Lattice::id2color - 1 (src/lattice.cpp:234-283)
      if (LAT == SPK_SQ_4N) col = (i + jg) & 1;
      else if (LAT == SPK_SQ_8N) col = 2 * (jg & 1) + (i & 1);
      else if (LAT == SPK_TRI) col = 2 * (i & 1) + m;
      else if (LAT == SPK_SC_6N) col = (i + jg + kg) & 1;
      else col = 4 * (kg & 1) + 2 * (jg & 1) + (i & 1);
      if (col == c0) mask[k * sz + j * sy + (long long)i * NB + m] = 0;
    }
  }
}

// site_energy(i) of every owned site (src/app_potts.cpp:97-104), one byte each: what
// `dump ... energy` (DumpText::pack_energy, src/dump_text.cpp:670) asks the app for site by site
template <int LAT>
__global__ void __launch_bounds__(256) site_energy_kernel(const int32_t *__restrict__ spin, Geo g, unsigned char *out) {
  using L = Lat<LAT>;
  constexpr int NN = L::NN, NB = L::NB;
  const long long sy = (long long)g.X * NB, sz = sy * g.Y;
  const long long ncell = (long long)g.X * g.Y * g.Z;
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < ncell; c += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(c % g.X);
    const int j = (int)((c / g.X) % g.Y);
    const int k = (int)(c / ((long long)g.X * g.Y));
    const int xi[3] = {i == 0 ? g.X - 1 : i - 1, i, i == g.X - 1 ? 0 : i + 1};
    int yj[3], zk[3];
    if (L::DIM == 2) {
      yj[0] = j == 0 ? g.lo_wrap : j - 1; yj[1] = j; yj[2] = j == g.Y - 1 ? g.hi_wrap : j + 1;
      zk[0] = zk[1] = zk[2] = 0;
    } else {
      yj[0] = j == 0 ? g.Y - 1 : j - 1; yj[1] = j; yj[2] = j == g.Y - 1 ? 0 : j + 1;
      zk[0] = k == 0 ? g.lo_wrap : k - 1; zk[1] = k; zk[2] = k == g.Z - 1 ? g.hi_wrap : k + 1;
    }
#pragma unroll
    for (int m = 0; m < NB; m++) {
      const long long self = zk[1] * sz + yj[1] * sy + (long long)i * NB + m;
      const int me = spin[self];
      int e = 0;
#pragma unroll
      for (int n = 0; n < NN; n++) {
        const Off o = L::off(m, n);
        e += (spin[zk[o.dz + 1] * sz + yj[o.dy + 1] * sy + (long long)xi[o.dx + 1] * NB + o.m] != me);
      }
      out[self] = (unsigned char)e;
    }
  }
}

// Order-independent 64-bit digest of the owned sites: sum over sites of mix64(global index, spin)
// (splitmix64 finaliser).  Slabs add up to the digest of the whole lattice, so it compares
// runs of any decomposition and any kernel family at sizes no host copy is wanted for.
__global__ void __launch_bounds__(256) digest_kernel(const int32_t *__restrict__ spin, long long n, long long goff,
                                                     unsigned long long *out) {
  unsigned long long acc = 0;
  for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < n; s += (long long)gridDim.x * blockDim.x) {
    unsigned long long z = (unsigned long long)(s + goff) * 0x9E3779B97F4A7C15ull + (unsigned long long)(uint32_t)spin[s];
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    acc += z ^ (z >> 31);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// diag_style energy (src/diag_energy.cpp:53-66): sum over owned sites of site_energy(i)
template <int LAT>
__global__ void __launch_bounds__(256) energy_kernel(const int32_t *__restrict__ spin, Geo g, unsigned long long *out) {
  using L = Lat<LAT>;
  constexpr int NN = L::NN, NB = L::NB;
  const long long sy = (long long)g.X * NB, sz = sy * g.Y;
  const long long ncell = (long long)g.X * g.Y * g.Z;
  unsigned int e = 0;
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < ncell; c += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(c % g.X);
    const int j = (int)((c / g.X) % g.Y);
    const int k = (int)(c / ((long long)g.X * g.Y));
    const int xi[3] = {i == 0 ? g.X - 1 : i - 1, i, i == g.X - 1 ? 0 : i + 1};
    int yj[3], zk[3];
    if (L::DIM == 2) {
      yj[0] = j == 0 ? g.lo_wrap : j - 1; yj[1] = j; yj[2] = j == g.Y - 1 ? g.hi_wrap : j + 1;
      zk[0] = zk[1] = zk[2] = 0;
    } else {
      yj[0] = j == 0 ? g.Y - 1 : j - 1; yj[1] = j; yj[2] = j == g.Y - 1 ? 0 : j + 1;
      zk[0] = k == 0 ? g.lo_wrap : k - 1; zk[1] = k; zk[2] = k == g.Z - 1 ? g.hi_wrap : k + 1;
    }
#pragma unroll
    for (int m = 0; m < NB; m++) {
      const int me = spin[zk[1] * sz + yj[1] * sy + (long long)i * NB + m];
#pragma unroll
      for (int n = 0; n < NN; n++) {
        const Off o = L::off(m, n);
        e += (spin[zk[o.dz + 1] * sz + yj[o.dy + 1] * sy + (long long)xi[o.dx + 1] * NB + o.m] != me);
      }
    }
  }
  e = __reduce_add_sync(0xffffffffu, e);
  __shared__ unsigned int part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = e;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += part[w];
    if (t) atomicAdd(out, t);
  }
}

inline int launch_energy(int lattice, cudaStream_t st, const int32_t *spin, const Geo &g, unsigned long long *out,
                         int64_t *launches) {
  const long long ncell = (long long)g.X * g.Y * g.Z;
  long long want = (ncell + 255) / 256;
  const int grid = (int)(want < 148 * 8 ? (want < 1 ? 1 : want) : 148 * 8);
  switch (lattice) {
    case SPK_SQ_4N: energy_kernel<SPK_SQ_4N><<<grid, 256, 0, st>>>(spin, g, out); break;
    case SPK_SQ_8N: energy_kernel<SPK_SQ_8N><<<grid, 256, 0, st>>>(spin, g, out); break;
    case SPK_TRI: energy_kernel<SPK_TRI><<<grid, 256, 0, st>>>(spin, g, out); break;
    case SPK_SC_6N: energy_kernel<SPK_SC_6N><<<grid, 256, 0, st>>>(spin, g, out); break;
    case SPK_SC_26N: energy_kernel<SPK_SC_26N><<<grid, 256, 0, st>>>(spin, g, out); break;
    default: return 1;
  }
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// ---- peer-memory halo ordering (spk_cuda.cu: sweeps_p2p) --------------------------------------
// signal: this rank has finished boundary step `value` (its stores into the neighbours' ghost
// planes were made by the preceding launches of the same stream); written into the neighbours'
// flag words over NVLink
__global__ void halo_signal_kernel(uint32_t *to_lower, uint32_t *to_upper, uint32_t value) {
  if (threadIdx.x == 0) {
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t *>(to_lower) = value;
    *reinterpret_cast<volatile uint32_t *>(to_upper) = value;
    __threadfence_system();
  }
}
// 2-D slabs: copy this rank's first / last ghost-depth rows into the neighbours' ghost rows over
// NVLink (128-bit stores) and signal the step once every block's stores are out: the last block
// to finish (device-wide counter `done`, reset for the next push) writes the flags.
__global__ void __launch_bounds__(1024) halo_push_kernel(const int4 *__restrict__ first, int4 *lower_ghost,
                                                         const int4 *__restrict__ last, int4 *upper_ghost, long long n16,
                                                         uint32_t *done, uint32_t *to_lower, uint32_t *to_upper,
                                                         uint32_t value) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) {
    lower_ghost[i] = first[i];
    upper_ghost[i] = last[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(done, 1u) == gridDim.x - 1) {
      *done = 0;
      __threadfence_system();
      *reinterpret_cast<volatile uint32_t *>(to_lower) = value;
      *reinterpret_cast<volatile uint32_t *>(to_upper) = value;
      __threadfence_system();
    }
  }
}
// wait until both neighbours have signalled step >= value; gives up after ~20 s (a neighbour died)
// and raises flags[2] so that the host reports it instead of hanging the device
__global__ void halo_wait_kernel(uint32_t *flags, uint32_t value) {
  if (threadIdx.x == 0) {
    const volatile uint32_t *f = flags;
    unsigned long long t0 = 0, now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int32_t)(f[0] - value) < 0 || (int32_t)(f[1] - value) < 0) {
      __nanosleep(200);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (now - t0 > 20000000000ull) { flags[2] = 1; break; }
    }
    __threadfence_system();
  }
}

}  // namespace spk
