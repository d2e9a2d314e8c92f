// TEST INFRASTRUCTURE.  The device source of the per-site events (spparks_b200/csrc/event.cuh)
// compiled for the host through cuda_shim.h, behind a small C interface for ctypes.
#include "cuda_shim.h"
#include <cmath>
#include <cstring>
#include "event.cuh"

using namespace spk;

namespace {

// EventParams exactly as spk_cuda.cu: build_tables fills them
EventParams make_params(int model, int Q, double temperature, uint64_t seed, uint32_t sweep) {
  EventParams ev;
  memset(&ev, 0, sizeof ev);
  ev.nspins = ising_like(model) ? 2 : Q;
  ev.t_zero = temperature == 0.0;
  ev.one = 1;
  ev.k0 = (uint32_t)seed;
  ev.k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; r++) {
    ev.rk.k[2 * r] = ev.k0 + (uint32_t)r * 0x9E3779B9u;
    ev.rk.k[2 * r + 1] = ev.k1 + (uint32_t)r * 0xBB67AE85u;
  }
  ev.sweep = sweep;
  ev.m65536 = -65536;
  ev.rep = 0x00010001u;
  for (int n = 0; n < 16; n++) ev.pow2[n] = 1u << n;
  const double t_inverse = ev.t_zero ? 0.0 : 1.0 / temperature;
  for (int dE = 0; dE < 27; dE++) {
    const double p = ev.t_zero ? 0.0 : std::exp((0.0 - (double)dE) * t_inverse);
    ev.ptab[dE] = p;
    const double t = std::floor(p * 4294967296.0);
    ev.thr[dE] = t > 4294967295.0 ? 4294967295u : (uint32_t)t;
  }
  return ev;
}

template <int MODEL, int NN, bool P16> int run_chain(const EventParams &ev, int old, const int *vin, uint32_t w1, uint32_t w2) {
  int v[NN];
  for (int n = 0; n < NN; n++) v[n] = vin[n];
  Draws<RNG_PHILOX> d;
  d.w1 = w1;
  d.w2 = w2;
  return site_event<MODEL, RNG_PHILOX, NN, P16>(ev, old, v, d);
}

template <int MODEL, int NN> int run_injected(const EventParams &ev, int old, const int *vin, double u1, double u2) {
  int v[NN];
  for (int n = 0; n < NN; n++) v[n] = vin[n];
  Draws<RNG_INJECTED> d;
  d.u1 = u1;
  d.u2 = u2;
  return site_event<MODEL, RNG_INJECTED, NN, false>(ev, old, v, d);
}
template <int MODEL, int NN> int run_strict(const EventParams &ev, int old, const int *vin, int *seed) {
  int v[NN];
  for (int n = 0; n < NN; n++) v[n] = vin[n];
  Draws<RNG_STRICT> d;
  d.seed = *seed;
  const int r = site_event<MODEL, RNG_STRICT, NN, false>(ev, old, v, d);
  *seed = d.seed;
  return r;
}

template <int MODEL> int run_nb26(const EventParams &ev, int old, const int *v, uint32_t w1, uint32_t w2) {
  uint32_t wd[13];
  for (int j = 0; j < 13; j++) wd[j] = pk_negpair(v[j], v[j + 13], ev);
  int kstat = 0;
  auto fetch = [&](int b) { return v[nb26_bit_to_slot(b)]; };
  return site_event_nb26<MODEL>(ev, old, wd, w1, w2, fetch, &kstat);
}

template <int MODEL> void run_nb26x2(const EventParams &ev, int *cur, const int *v0, const int *v1, const uint4 &blk) {
  uint32_t wd[2][13];
  for (int j = 0; j < 13; j++) {
    wd[0][j] = pk_negpair(v0[j], v0[j + 13], ev);
    wd[1][j] = pk_negpair(v1[j], v1[j + 13], ev);
  }
  int kstat = 0, c[2] = {cur[0], cur[1]};
  auto fetch = [&](int h, int b) { return b == 31 ? c[h] : (h ? v1 : v0)[nb26_bit_to_slot(b)]; };
  site_event_nb26x2<MODEL>(ev, c, wd, blk, fetch, &kstat);
  cur[0] = c[0];
  cur[1] = c[1];
}

}  // namespace

extern "C" {

// one Philox-mode event through the device source: variant 0 = 32-bit compares, 1 = packed 16-bit
// chain, 2 = candidate elimination (sc/26n candidate-list apps), 3 = candidate elimination for a
// pair of sites (site 1 = the same neighbourhood rotated by `rot`); returns the new spin
int shim_site_event(int model, int nn, int variant, int Q, double temperature, int old, const int *v, uint32_t w1, uint32_t w2) {
  const EventParams ev = make_params(model, Q, temperature, 1, 0);
#define CASE(M, N)                                                                      \
  if (model == M && nn == N) {                                                          \
    if (variant == 0) return run_chain<M, N, false>(ev, old, v, w1, w2);                \
    if (variant == 1) { if constexpr (N % 2 == 0) return run_chain<M, N, true>(ev, old, v, w1, w2); } \
  }
  CASE(SPK_ISING, 4) CASE(SPK_ISING, 6) CASE(SPK_ISING, 8) CASE(SPK_ISING, 26)
  CASE(SPK_ISING_SINGLE, 4) CASE(SPK_ISING_SINGLE, 6) CASE(SPK_ISING_SINGLE, 8) CASE(SPK_ISING_SINGLE, 26)
  CASE(SPK_POTTS, 4) CASE(SPK_POTTS, 6) CASE(SPK_POTTS, 8) CASE(SPK_POTTS, 26)
  CASE(SPK_POTTS_PIN, 4) CASE(SPK_POTTS_PIN, 6) CASE(SPK_POTTS_PIN, 8) CASE(SPK_POTTS_PIN, 26)
  CASE(SPK_POTTS_NEIGHONLY, 4) CASE(SPK_POTTS_NEIGHONLY, 6) CASE(SPK_POTTS_NEIGHONLY, 8) CASE(SPK_POTTS_NEIGHONLY, 26)
  CASE(SPK_POTTS_NEIGH, 4) CASE(SPK_POTTS_NEIGH, 6) CASE(SPK_POTTS_NEIGH, 8) CASE(SPK_POTTS_NEIGH, 26)
#undef CASE
  if (variant == 2 && nn == 26) {
    if (model == SPK_POTTS_NEIGHONLY) return run_nb26<SPK_POTTS_NEIGHONLY>(ev, old, v, w1, w2);
    if (model == SPK_POTTS_NEIGH) return run_nb26<SPK_POTTS_NEIGH>(ev, old, v, w1, w2);
  }
  return -1;
}

// the double modes: injected uniforms (u1, u2) and `sweep color/strict` (per-site Park-Miller seed, updated)
#define ALL_CASES(X)                                                                                         \
  X(SPK_ISING, 4) X(SPK_ISING, 6) X(SPK_ISING, 8) X(SPK_ISING, 26)                                           \
  X(SPK_ISING_SINGLE, 4) X(SPK_ISING_SINGLE, 6) X(SPK_ISING_SINGLE, 8) X(SPK_ISING_SINGLE, 26)               \
  X(SPK_POTTS, 4) X(SPK_POTTS, 6) X(SPK_POTTS, 8) X(SPK_POTTS, 26)                                           \
  X(SPK_POTTS_PIN, 4) X(SPK_POTTS_PIN, 6) X(SPK_POTTS_PIN, 8) X(SPK_POTTS_PIN, 26)                           \
  X(SPK_POTTS_NEIGHONLY, 4) X(SPK_POTTS_NEIGHONLY, 6) X(SPK_POTTS_NEIGHONLY, 8) X(SPK_POTTS_NEIGHONLY, 26)   \
  X(SPK_POTTS_NEIGH, 4) X(SPK_POTTS_NEIGH, 6) X(SPK_POTTS_NEIGH, 8) X(SPK_POTTS_NEIGH, 26)
int shim_site_event_injected(int model, int nn, int Q, double temperature, int old, const int *v, double u1, double u2) {
  const EventParams ev = make_params(model, Q, temperature, 1, 0);
#define X(M, N) if (model == M && nn == N) return run_injected<M, N>(ev, old, v, u1, u2);
  ALL_CASES(X)
#undef X
  return -1;
}
int shim_site_event_strict(int model, int nn, int Q, double temperature, int old, const int *v, int *seed) {
  const EventParams ev = make_params(model, Q, temperature, 1, 0);
#define X(M, N) if (model == M && nn == N) return run_strict<M, N>(ev, old, v, seed);
  ALL_CASES(X)
#undef X
  return -1;
}

// the pair form of candidate elimination: two sites with their own neighbourhoods and one Philox block
int shim_site_event_nb26x2(int model, int Q, double temperature, int *cur, const int *v0, const int *v1, const uint32_t *blk) {
  const EventParams ev = make_params(model, Q, temperature, 1, 0);
  const uint4 b = make_uint4(blk[0], blk[1], blk[2], blk[3]);
  if (model == SPK_POTTS_NEIGHONLY) run_nb26x2<SPK_POTTS_NEIGHONLY>(ev, cur, v0, v1, b);
  else if (model == SPK_POTTS_NEIGH) run_nb26x2<SPK_POTTS_NEIGH>(ev, cur, v0, v1, b);
  else return -1;
  return 0;
}

// Philox4x32-10 of the device source, both forms (key schedule on the fly / precomputed)
void shim_philox(const uint32_t *ctr, const uint32_t *key, uint32_t *out, int precomputed) {
  uint4 r;
  if (precomputed) {
    PhiloxKeys rk;
    for (int i = 0; i < 10; i++) { rk.k[2 * i] = key[0] + (uint32_t)i * 0x9E3779B9u; rk.k[2 * i + 1] = key[1] + (uint32_t)i * 0xBB67AE85u; }
    r = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], rk);
  } else {
    r = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
  }
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

}  // extern "C"
