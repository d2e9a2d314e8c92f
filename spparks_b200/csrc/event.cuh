// Per-site rejection event (site_event_rejection of the four apps) and the random sources.
#pragma once
#include <stdint.h>
#include <type_traits>
#include "dedupe_chain.cuh"
#include "lattice.cuh"

namespace spk {

enum { RNG_PHILOX = 0, RNG_INJECTED = 1, RNG_STRICT = 2 };

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. SC'11).  One block serves the PAIR of same-colour sites whose cell
// index i differs only in bit 1: counter = (leader lo, leader hi, sweep, iloop*16 + colour) with
// leader = global index of the site with bit 1 of i cleared; key = seed.  The site with bit 1
// clear takes words (x,y), the other one (z,w).  The key schedule is warp-uniform.
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c2 = hi0 ^ c3 ^ k1;
    c1 = lo1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// Same function with the ten round keys precomputed on the host (EventParams::rk): the key
// schedule then costs no instructions, the keys are constant-bank operands of the XORs.
struct PhiloxKeys { uint32_t k[20]; };
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys &rk) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ rk.k[2 * r];
    c2 = hi0 ^ c3 ^ rk.k[2 * r + 1];
    c1 = lo1;
    c3 = lo0;
  }
  return make_uint4(c0, c1, c2, c3);
}

// Which Philox block and which words serve the class site at cell index i (global).
//   potts family: two sites per block (cells i, i^2): leader has bit 1 of i cleared, the site
//                 with bit 1 clear takes (x,y), the other (z,w)
//   ising:        four sites per block (cells differing in bits 1,2): the trial spin needs one
//                 bit only, so site (i>>1)&3 takes ONE word w: u1 = its top bit, u2 = (w mod 2^31)/2^31,
//                 expressed as second word = w << 1 so that `w2 > thr` is unchanged
// the two-state apps (one Philox word per site) and the candidate-list apps
__host__ __device__ constexpr bool ising_like(int model) { return model == SPK_ISING || model == SPK_ISING_SINGLE; }
__host__ __device__ constexpr bool neigh_like(int model) { return model == SPK_POTTS_NEIGHONLY || model == SPK_POTTS_NEIGH; }

template <int MODEL> __device__ __forceinline__ int philox_leader_shift(int i) {
  return ising_like(MODEL) ? (i & 6) : (i & 2);  // cells between the site and its block leader
}
template <int MODEL> __device__ __forceinline__ void philox_words(const uint4 &blk, int i, uint32_t &w1, uint32_t &w2) {
  if (ising_like(MODEL)) {
    const int s = (i >> 1) & 3;
    w1 = s == 0 ? blk.x : (s == 1 ? blk.y : (s == 2 ? blk.z : blk.w));
    w2 = w1 << 1;
  } else {
    const bool second = (i >> 1) & 1;
    w1 = second ? blk.z : blk.x;
    w2 = second ? blk.w : blk.y;
  }
}

// RandomPark::uniform (src/random_park.cpp:79-86) as 16807*seed mod (2^31-1); the Schrage form
// of the reference and this 64-bit product give the same residue for seed in [1, 2^31-2].
__device__ __forceinline__ double park_uniform(int &seed) {
  const uint64_t p = (uint64_t)16807u * (uint32_t)seed;
  uint32_t s = (uint32_t)(p & 0x7FFFFFFFu) + (uint32_t)(p >> 31);
  if (s >= 0x7FFFFFFFu) s -= 0x7FFFFFFFu;
  seed = (int)s;
  return (1.0 / 2147483647) * (double)seed;
}

// Everything an event needs besides the spins; lives in kernel parameters (constant bank).
struct EventParams {
  int nspins;               // Q
  int t_zero;               // temperature == 0.0
  int one;                  // the constant 1, opaque to the compiler (see unlike())
  uint32_t k0, k1;          // Philox key (seed)
  PhiloxKeys rk;            // its ten round keys (host-side key schedule)
  uint32_t sweep;           // Philox counter word 2
  uint32_t thr[27];         // Philox mode: floor(exp(-dE/T) * 2^32), dE = 1..26 (host libm)
  double ptab[27];          // double modes: exp((double)(-dE) * t_inverse)       (host libm)
  // packed 16-bit path (pk16 below): multipliers kept opaque to the compiler so that the packing
  // arithmetic stays IMAD (FMA pipe) instead of LEA/PRMT/SHF (ALU pipe, which the compares saturate)
  int m65536;               // -65536
  uint32_t rep;             // 0x00010001
  uint32_t pow2[16];        // 1 << n
};

// Source of the (at most) two uniforms of one event; the double modes draw lazily like
// random->uniform() does.
template <int RNG> struct Draws;
template <> struct Draws<RNG_PHILOX> { uint32_t w1, w2; };
template <> struct Draws<RNG_INJECTED> { double u1, u2; };
template <> struct Draws<RNG_STRICT> { int seed; };

// number of neighbours unlike `value`  (site_energy, src/app_potts.cpp:97-104).
// Each compare is one ISETP (ALU pipe) plus one predicated IMAD (FMA pipe): `one` is the value 1
// passed through kernel parameters so that ptxas keeps the multiply-add form and the two integer
// pipes stay balanced; nvcc's own code for `e += (v != s)` costs 2.4-3.2 ALU-heavy instructions.
template <int NN> __device__ __forceinline__ int unlike(const int (&v)[NN], int value, int one) {
  // two accumulator chains (the sites of a strip interleave, so four chains are in flight); the
  // first compare of each chain writes its result directly instead of adding to a zeroed register
  int e[2];
#ifdef SPK_HOST_SHIM  // host build of tests/host_shim: the same count without PTX
  e[0] = e[1] = 0;
  for (int n = 0; n < NN; n++) e[n & 1] += (v[n] != value) ? one : 0;
  return NN > 1 ? e[0] + e[1] : e[0];
#else
#pragma unroll
  for (int n = 0; n < NN; n++) {
    if (n < 2)
      asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %1, %2;\n\tselp.s32 %0, 1, 0, p;\n\t}"
          : "=r"(e[n]) : "r"(v[n]), "r"(value));
    else
      asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %1, %2;\n\t@p mad.lo.s32 %0, %3, %3, %0;\n\t}"
          : "+r"(e[n & 1]) : "r"(v[n]), "r"(value), "r"(one));
  }
  return NN > 1 ? e[0] + e[1] : e[0];
#endif
}

// ---------------------------------------------------------------------------------------------
// Packed 16-bit compares (all spins of the lattice <= PK16_MAX_SPIN): VIADDMNMX.U16x2, the DPX
// "add then min" instruction of sm_90+, does TWO equality tests and their accumulation in one
// ALU-pipe instruction:  min(a + b, acc) per halfword with a = (x, x), b = (-y0, -y1) mod 2^16
// leaves a zero halfword exactly where x == y.  The 32-bit ISETP form costs one instruction per
// test on the same pipe (measured with tools/ubench: both 1.5 cycles per warp instruction per
// SMSP, and no compare instruction of sm_100 runs on the FMA pipe).
//
// Encoding: a neighbour pair (lo, hi) is stored as the 32-bit NEGATIVE of (hi << 16 | lo): its low
// halfword is -lo, its high halfword -hi - 1 (lo >= 1 always borrows); the "+1" is put into the
// replicated operand instead, rep(x) = (x + 1) << 16 | x, hence the x <= 65534 limit.
constexpr int PK16_MAX_SPIN = 65534;

__device__ __forceinline__ uint32_t pk_rep(int x, const EventParams &p) { return (uint32_t)x * p.rep + 0x10000u; }
__device__ __forceinline__ uint32_t pk_negpair(int lo, int hi, const EventParams &p) {
  return (uint32_t)(hi * p.m65536 - lo);
}
// number of neighbours unlike `value` from the packed pairs: min(x - y, 1) per halfword, summed
// with packed IMADs
template <int NN> __device__ __forceinline__ int unlike16(const uint32_t (&negp)[NN / 2], int value, const EventParams &p) {
  static_assert(NN % 2 == 0, "packed path needs an even stencil");
  const uint32_t a = pk_rep(value, p);
  uint32_t e[2];
#pragma unroll
  for (int j = 0; j < NN / 2; j++) {
    const uint32_t t = __viaddmin_u16x2(a, negp[j], 0x00010001u);
    if (j < 2) e[j] = t;
    else e[j & 1] = t * (uint32_t)p.one + e[j & 1];
  }
  const uint32_t s = NN > 2 ? e[0] + e[1] : e[0];
  return (int)((s & 0xFFFFu) + (s >> 16));
}
// first_occurrence_mask (dedupe_chain.cuh) on packed pairs: slot n is tested against the pairs of
// the earlier slots plus (old, old) [n even] or (v[n-1], old) [n odd]; the chain starts at (1, 1)
// so that it ends as 0x00010001 exactly when every test was "unequal".  The 0/1 halfwords of all
// slots are gathered by IMADs (halfword bit n = slot n, 16 slots per accumulator) and ANDed.
template <int NN>
__device__ __forceinline__ uint32_t first_occurrence_mask16(const int (&v)[NN], const uint32_t (&negp)[NN / 2], int old,
                                                             const EventParams &p) {
  const uint32_t negoo = pk_negpair(old, old, p);
  uint32_t sum[2] = {0u, 0u};
#pragma unroll
  for (int n = 0; n < NN; n++) {
    const uint32_t a = pk_rep(v[n], p);
    uint32_t acc = 0x00010001u;
#pragma unroll
    for (int j = 0; j < n / 2; j++) acc = __viaddmin_u16x2(a, negp[j], acc);
    acc = __viaddmin_u16x2(a, (n & 1) ? pk_negpair(v[n - 1], old, p) : negoo, acc);
    sum[n >> 4] = acc * p.pow2[n & 15] + sum[n >> 4];
  }
  uint32_t first = sum[0] & (sum[0] >> 16);
  if (NN > 16) first |= (sum[1] & (sum[1] >> 16)) << 16;
  return first;
}
// v[pos] from the packed pairs: select tree over NN/2 words, then the halfword
template <int NN> __device__ __forceinline__ int select_slot16(const uint32_t (&negp)[NN / 2], int pos) {
  uint32_t t[16];
#pragma unroll
  for (int n = 0; n < 16; n++) t[n] = negp[n < NN / 2 ? n : NN / 2 - 1];
#pragma unroll
  for (int bit = 1, len = 16; bit < 5; bit++, len >>= 1) {
    const bool hi = (pos >> bit) & 1;
#pragma unroll
    for (int n = 0; n < len / 2; n++) t[n] = hi ? t[2 * n + 1] : t[2 * n];
  }
  const uint32_t pr = 0u - t[0];  // hi << 16 | lo
  return (int)((pos & 1) ? (pr >> 16) : (pr & 0xFFFFu));
}

// position of the pick-th (0-based) set bit of m; pick < popc(m)
__device__ __forceinline__ int nth_set_bit(uint32_t m, int pick) {
  int pos = 0;
#pragma unroll
  for (int width = 16; width >= 1; width >>= 1) {
    const int c = __popc(m & ((1u << width) - 1u));
    if (pick >= c) { pick -= c; m >>= width; pos += width; }
  }
  return pos;
}

// v[pos] for a register-resident array: binary select tree on the bits of pos
template <int NN> __device__ __forceinline__ int select_slot(const int (&v)[NN], int pos) {
  int t[32];
#pragma unroll
  for (int n = 0; n < 32; n++) t[n] = v[n < NN ? n : NN - 1];
#pragma unroll
  for (int bit = 0, len = 32; bit < 5; bit++, len >>= 1) {
    const bool hi = (pos >> bit) & 1;
#pragma unroll
    for (int n = 0; n < len / 2; n++) t[n] = hi ? t[2 * n + 1] : t[2 * n];
  }
  return t[0];
}

// ---------------------------------------------------------------------------------------------
// Candidate elimination for the candidate-list apps on sc/26n (packed 16-bit spins): work
// proportional to the number of DISTINCT neighbour values instead of the 351 pairwise tests of
// first_occurrence_mask16.  After a few sweeps of grain growth a site sees 2-6 distinct values
// (a fresh random lattice: 26), so this is the production path once the lattice has coarsened;
// the host picks the variant per launch from a device-side statistic (spk_cuda.cu: neigh_mode).
//
// Layout: word j (0..12) holds the negated pair (slot j, slot j+13), so that in a result mask
// bit j <-> slot j and bit 16+j <-> slot 13+j: bit order = reference neighbour order, the lowest
// set bit of a mask is its earliest slot.
constexpr uint32_t NB26_VALID = 0x1FFF1FFFu;
__host__ __device__ constexpr int nb26_bit_to_slot(int b) { return b < 16 ? b : b - 3; }

// bit set <=> that slot's value differs from c  (13 VIADDMNMX on the ALU pipe + 13 IMAD gathers
// on the FMA pipe)
__device__ __forceinline__ uint32_t nb26_gather(const uint32_t (&t)[13], const EventParams &p) {
  uint32_t s[2];
#pragma unroll
  for (int j = 0; j < 13; j++) {
    if (j == 0) s[0] = t[0];
    else if (j == 1) s[1] = t[1] * p.pow2[1];
    else s[j & 1] = t[j] * p.pow2[j] + s[j & 1];
  }
  return s[1] * (uint32_t)p.one + s[0];
}
__device__ __forceinline__ uint32_t nb26_unequal_mask(const uint32_t (&w)[13], int c, const EventParams &p) {
  const uint32_t a = pk_rep(c, p);
  uint32_t t[13];
#pragma unroll
  for (int j = 0; j < 13; j++) t[j] = __viaddmin_u16x2(a, w[j], 0x00010001u);
  return nb26_gather(t, p);
}
// bit set <=> that slot's value differs from c1 AND from c2: the second test takes the first
// one's result as its min operand, so two candidates share one gather
__device__ __forceinline__ uint32_t nb26_unequal_mask2(const uint32_t (&w)[13], int c1, int c2, const EventParams &p) {
  const uint32_t a1 = pk_rep(c1, p), a2 = pk_rep(c2, p);
  uint32_t t[13];
#pragma unroll
  for (int j = 0; j < 13; j++) t[j] = __viaddmin_u16x2(a2, w[j], __viaddmin_u16x2(a1, w[j], 0x00010001u));
  return nb26_gather(t, p);
}

// site_event_rejection of potts/neighonly (src/app_potts_neighonly.cpp:71-117) and potts/neigh
// (src/app_potts_neigh.cpp:71-116) by candidate elimination, for the TWO class sites of a chunk at
// once (two independent dependency chains per thread hide the latency of a pass).
// fetch(h, b) returns the neighbour value of site h behind mask bit b (the streaming kernel reads
// it from the shared-memory planes; bit 31 is never a neighbour and must return any valid spin).
// Same draws, same first-occurrence order, hence the same result as site_event<..., P16>.
// *kstat accumulates the number of candidates (for the host's variant choice).
template <int MODEL, class Fetch>
__device__ __forceinline__ void site_event_nb26x2(const EventParams &p, int (&cur)[2], const uint32_t (&w)[2][13],
                                                  const uint4 &blk, Fetch fetch, int *kstat) {
  constexpr int NN = 26;
  constexpr bool NULLBIN = MODEL == SPK_POTTS_NEIGH;
  uint32_t u0[2], rem[2], first[2] = {0u, 0u};
#pragma unroll
  for (int h = 0; h < 2; h++) rem[h] = u0[h] = nb26_unequal_mask(w[h], cur[h], p);
  while (rem[0] | rem[1]) {
    // per site: the two earliest slots still unexplained; the first one opens a new candidate, the
    // second one too unless it repeats the first one's value; both are eliminated in one pass.
    // A site that is done (rem == 0) runs along on bit 31 without effect.
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const uint32_t lo1 = rem[h] & (0u - rem[h]), rest = rem[h] ^ lo1;
      const uint32_t lo2 = rest ? (rest & (0u - rest)) : lo1;
      const int c1 = fetch(h, (31 - __clz((int)lo1)) & 31), c2 = fetch(h, (31 - __clz((int)lo2)) & 31);
      first[h] |= lo1 | (c1 != c2 ? lo2 : 0u);
      rem[h] &= nb26_unequal_mask2(w[h], c1, c2, p);
    }
  }
  int trial[2], nevent[2];
  bool live[2];
#pragma unroll
  for (int h = 0; h < 2; h++) {
    nevent[h] = __popc(first[h]);
    const int pick = (int)__umulhi((uint32_t)(NULLBIN ? NN : nevent[h]), h ? blk.z : blk.x);
    live[h] = nevent[h] > 0 && (!NULLBIN || pick < nevent[h]);
    uint32_t f = first[h];
    for (int i = 0; i < pick && live[h]; i++) f &= f - 1u;  // drop the `pick` earliest candidates
    trial[h] = fetch(h, (31 - __clz((int)(f & (0u - f)))) & 31);
  }
  *kstat += nevent[0] + nevent[1];
  if (!(live[0] || live[1])) return;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int e0 = __popc(u0[h]);
    const int e1 = __popc(nb26_unequal_mask(w[h], trial[h], p));
    const uint32_t w2 = h ? blk.w : blk.y;
    const bool ok = live[h] && (e1 <= e0 || (!p.t_zero && w2 <= p.thr[e1 > e0 ? e1 - e0 : 0]));
    cur[h] = ok ? trial[h] : cur[h];
  }
}

// the same for one site (sequential variant: the two sites of a chunk one after the other)
template <int MODEL, class Fetch>
__device__ __forceinline__ int site_event_nb26(const EventParams &p, int old, const uint32_t (&w)[13],
                                               uint32_t w1, uint32_t w2, Fetch fetch, int *kstat) {
  constexpr int NN = 26;
  constexpr bool NULLBIN = MODEL == SPK_POTTS_NEIGH;
  const uint32_t u0 = nb26_unequal_mask(w, old, p);
  if (u0 == 0) return old;  // nevent == 0
  uint32_t rem = u0, first = 0;
  do {
    const uint32_t lo1 = rem & (0u - rem), rest = rem ^ lo1;
    const uint32_t lo2 = rest ? (rest & (0u - rest)) : lo1;
    const int c1 = fetch(31 - __clz((int)lo1)), c2 = fetch(31 - __clz((int)lo2));
    first |= lo1 | (c1 != c2 ? lo2 : 0u);
    rem &= nb26_unequal_mask2(w, c1, c2, p);
  } while (rem);
  const int nevent = __popc(first);
  *kstat += nevent;
  const int pick = (int)__umulhi((uint32_t)(NULLBIN ? NN : nevent), w1);
  if constexpr (NULLBIN) {
    if (pick >= nevent) return old;
  }
  for (int i = 0; i < pick; i++) first &= first - 1u;  // drop the `pick` earliest candidates
  const int trial = fetch(31 - __clz((int)(first & (0u - first))));
  const int e0 = __popc(u0);
  if (e0 == NN) return trial;  // efinal <= einitial whatever the trial
  const int e1 = __popc(nb26_unequal_mask(w, trial, p));
  if (e1 <= e0) return trial;
  if (p.t_zero) return old;
  return w2 > p.thr[e1 - e0] ? old : trial;
}

// One site_event_rejection.  `old` = spin[i], v[] = neighbour spins in reference order.
// Returns the new spin (== old when rejected or no event).
// P16: every spin of the lattice is <= PK16_MAX_SPIN, compares run on packed pairs (see above).
// kstat (optional): accumulates nevent of the candidate-list apps (host-side variant choice).
// einitial (optional, `sweep ... mask yes`): receives the initial energy when the event ran to its
// end, -1 when it returned early (pinned site, no candidate, null bin): the reference's mask
// update sits after those returns (src/app_potts_neighonly.cpp:90, 107-117).
// fetch (optional, candidate-list apps): fetch(n) returns v[n] for a run-time slot n from wherever the
// caller keeps the neighbourhood addressable (the streaming kernel: its shared-memory planes); without
// it the value is picked out of the registers by a select tree (~20 ALU instructions per site).
struct NoFetch {};
template <int MODEL, int RNG, int NN, bool P16 = false, class Fetch = NoFetch>
__device__ __forceinline__ int site_event(const EventParams &p, int old, const int (&v)[NN], Draws<RNG> &d,
                                          int *kstat = nullptr, int *einitial = nullptr, Fetch fetch = Fetch()) {
  const int Q = p.nspins;
  if (einitial) *einitial = -1;
  if (MODEL == SPK_POTTS_PIN && old > Q) return old;  // pinned: no draw (src/app_potts_pin.cpp:97)

  constexpr bool PK = P16 && !ising_like(MODEL);
  uint32_t negp[PK ? NN / 2 : 1];
  if constexpr (PK) {
#pragma unroll
    for (int j = 0; j < NN / 2; j++) negp[j] = pk_negpair(v[2 * j], v[2 * j + 1], p);
  }
  int e0;
  if constexpr (PK) e0 = unlike16<NN>(negp, old, p);
  else e0 = unlike<NN>(v, old, p.one);
  int trial;

  if constexpr (neigh_like(MODEL)) {
    // candidates = distinct neighbour values != old, first-occurrence order
    // (src/app_potts_neighonly.cpp:81-88, src/app_potts_neigh.cpp:79-87); bit n of `first` marks
    // slot n as a new candidate.  potts/neighonly picks among the nevent candidates and draws
    // nothing when there is none (:90); potts/neigh ALWAYS draws and picks among maxneigh bins,
    // the bins beyond nevent being null events (src/app_potts_neigh.cpp:89-91)
    constexpr bool NULLBIN = MODEL == SPK_POTTS_NEIGH;
    double u = 0.0;
    if constexpr (NULLBIN && RNG != RNG_PHILOX) {
      if constexpr (RNG == RNG_INJECTED) u = d.u1;
      else u = park_uniform(d.seed);
    }
    if (e0 == 0) return old;  // nevent == 0: no event (neighonly: no draw either)
    uint32_t first;
    if constexpr (PK) first = first_occurrence_mask16<NN>(v, negp, old, p);
    else first = first_occurrence_mask<NN>(v, old);
    const int nevent = __popc(first);
    if (kstat) *kstat += nevent;
    int pick;
    if constexpr (RNG == RNG_PHILOX) {
      pick = (int)__umulhi((uint32_t)(NULLBIN ? NN : nevent), d.w1);
    } else if constexpr (NULLBIN) {
      pick = (int)(NN * u);
    } else {
      if constexpr (RNG == RNG_INJECTED) u = d.u1;
      else u = park_uniform(d.seed);
      pick = (int)(nevent * u);
      if (pick >= nevent) pick = nevent - 1;
    }
    if constexpr (NULLBIN) {
      if (pick >= nevent) return old;
    }
    if constexpr (!std::is_same<Fetch, NoFetch>::value) trial = fetch(nth_set_bit(first, pick));
    else if constexpr (PK) trial = select_slot16<NN>(negp, nth_set_bit(first, pick));
    else trial = select_slot<NN>(v, nth_set_bit(first, pick));
  } else if constexpr (MODEL == SPK_ISING_SINGLE) {
    trial = 3 - old;  // spins are 1 or 2 (src/app_ising_single.cpp:46-47)
  } else if constexpr (RNG == RNG_PHILOX) {
    if (MODEL == SPK_ISING) trial = (d.w1 < 0x80000000u) ? 1 : 2;
    else trial = (int)__umulhi((uint32_t)Q, d.w1) + 1;
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

  if (einitial) *einitial = e0;
  // e0 == NN (every neighbour unlike): efinal <= einitial whatever the trial, so the event is
  // accepted without counting; a warp whose lanes all sit in that case (fresh random lattices)
  // skips the second energy altogether
  if constexpr (NN > 8) {
    if (e0 == NN) return trial;
  }
  if constexpr (ising_like(MODEL)) {
    // two-state shortcut: with spins in {1,2} the neighbours unlike `old` are exactly the ones
    // like a different trial, so efinal = NN - einitial and no second count is needed
    // (ising/single: the trial always differs and its acceptance draw is the event's first draw)
    const int dE2 = NN - 2 * e0;
    if constexpr (RNG == RNG_PHILOX) {
      // branch-free: thr[0] = 2^32-1 accepts every dE <= 0; T = 0 has thr[dE>0] = 0 and must
      // not accept on w2 == 0, hence the explicit t_zero term
      uint32_t thr;
      if constexpr (NN <= 8) {
        // dE2 > 0 takes the values 2, 4, ..., NN only: a select chain over constant-bank operands
        // instead of a dynamically indexed constant load (LDC: long-scoreboard latency per site)
        thr = p.thr[NN];
#pragma unroll
        for (int dE = NN - 2; dE >= 2; dE -= 2) thr = dE2 == dE ? p.thr[dE] : thr;
      } else {
        thr = p.thr[dE2 > 0 ? dE2 : 0];
      }
      const bool ok = (dE2 <= 0) || (!p.t_zero && d.w2 <= thr);
      return ok ? trial : old;
    } else {
      if (trial == old) return old;
      if (dE2 <= 0) return trial;
      if (p.t_zero) return old;
      double u;
      if constexpr (RNG == RNG_INJECTED) u = MODEL == SPK_ISING_SINGLE ? d.u1 : d.u2;
      else u = park_uniform(d.seed);
      return u > p.ptab[dE2] ? old : trial;
    }
  }
  int e1;
  if constexpr (PK) e1 = unlike16<NN>(negp, trial, p);
  else e1 = unlike<NN>(v, trial, p.one);
  if (e1 <= e0) return trial;
  if (p.t_zero) return old;
  const int dE = e1 - e0;
  bool reject;
  if constexpr (RNG == RNG_PHILOX) {
    reject = d.w2 > p.thr[dE];
  } else {
    double u;
    if constexpr (RNG == RNG_INJECTED) u = d.u2;
    else u = park_uniform(d.seed);
    reject = u > p.ptab[dE];  // src/app_potts.cpp:130
  }
  return reject ? old : trial;
}

}  // namespace spk
