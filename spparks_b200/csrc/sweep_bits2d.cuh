// Multi-spin-coded colour kernel for the two-state apps (ising, ising/single) on sq/4n.
//
// The two spin values live as ONE BIT per site in a packed copy of the lattice ("bit shadow", kept
// next to the canonical int32 lattice like the 16-bit shadow of sweep_stream16.cuh): one 32-bit
// word holds 32 same-colour sites of a row, so the four neighbour spins of 32 sites are four words
// of the other colour and the unlike counts come out of a bit-sliced adder (a dozen logic
// instructions per 32 sites instead of ~8 per site).  What remains per site is what the random
// stream costs: one Philox4x32-10 word per site, same block / word assignment as every other
// kernel (event.cuh: four same-colour sites per block, counter = (leader, sweep, colour)), the
// top bit is the trial spin of AppIsing::site_event_rejection (src/app_ising.cpp:111-112), the low
// 31 bits the acceptance uniform compared with the host-tabulated thresholds (:119-123) — so the
// result is bit-identical to the int32 kernels and to the oracle.
//
// Layout: bits[c][j][k], c = colour (i + jglobal) & 1, j = local row, k = word; bit b of word k is
// site m = 32 k + b of the colour in that row, at cell i = 2 m + ((c + jglobal) & 1).  Rows of the
// other colour above and below hold the vertical neighbours at the same (k, b); the horizontal
// neighbours are the other colour's word k of the same row and that word shifted by one site.
// A whole 16384^2 lattice is 32 MB: it stays in L2, the kernel is bound by the Philox arithmetic.
#pragma once
#include "event.cuh"

namespace spk {

struct Bits2dArgs {
  uint32_t *bits;         // [2][Y][W]
  int Y, W;               // rows, words per colour row (X / 64)
  int c0;                 // colour - 1
  unsigned long long gX;  // cells per row (global = local: one GPU)
  EventParams ev;
  unsigned long long *naccept;
};

// acc <- 2*acc + (w2 > thr): the carry of w2 + ~thr (= w2 - thr - 1 + 2^32) shifted in, two
// instructions per compare (an add.cc / addc pair: plain extended-precision addition)
__device__ __forceinline__ void shift_in_reject(uint32_t &acc, uint32_t not_thr, uint32_t w2) {
#ifdef SPK_HOST_SHIM  // host build of tests/host_shim: the same carry without PTX
  const uint32_t sum = w2 + not_thr;
  acc = acc + acc + (sum < w2 ? 1u : 0u);
#else
  uint32_t tmp;
  asm("{\n\tadd.cc.u32 %1, %2, %3;\n\taddc.u32 %0, %0, %0;\n\t}" : "+r"(acc), "=r"(tmp) : "r"(w2), "r"(not_thr));
#endif
}

template <int MODEL>
__global__ void __launch_bounds__(128) sweep_bits2d_kernel(const __grid_constant__ Bits2dArgs a) {
  const int W = a.W, Y = a.Y;
  const long long nwords = (long long)W * Y;
  const uint32_t *other = a.bits + (long long)(1 - a.c0) * nwords;
  uint32_t *mine = a.bits + (long long)a.c0 * nwords;
  const uint32_t nthr2 = ~a.ev.thr[2], nthr4 = ~a.ev.thr[4];
  int accepted = 0;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nwords; t += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(t / W), k = (int)(t - (long long)j * W);
    const int px = (a.c0 + j) & 1;  // x parity of this colour's sites in row j
    const int ju = j == 0 ? Y - 1 : j - 1, jd = j == Y - 1 ? 0 : j + 1;
    const uint32_t own = mine[t];
    const uint32_t up = other[(long long)ju * W + k], down = other[(long long)jd * W + k];
    const uint32_t same = other[t];
    // px = 0: the other colour's site m is my right neighbour, its site m-1 my left one;
    // px = 1: site m is my left neighbour, site m+1 my right one (periodic in x)
    const uint32_t adj = other[(long long)j * W + (px ? (k == W - 1 ? 0 : k + 1) : (k == 0 ? W - 1 : k - 1))];
    const uint32_t shifted = px ? __funnelshift_r(same, adj, 1) : __funnelshift_l(adj, same, 1);
    // unlike neighbours per site, bit-sliced: count = s0 + 2 s1 + 4 s2
    const uint32_t n1 = up ^ own, n2 = down ^ own, n3 = same ^ own, n4 = shifted ^ own;
    const uint32_t x12 = n1 ^ n2, c12 = n1 & n2, x34 = n3 ^ n4, c34 = n3 & n4;
    const uint32_t s0 = x12 ^ x34, cc = x12 & x34;
    const uint32_t ge2 = c12 | c34 | cc;        // einitial >= 2: efinal <= einitial, accept (src/app_ising.cpp:117)
    const uint32_t eq1 = s0 & ~ge2, eq0 = ~(s0 | ge2);

    // one Philox word per site: block u serves sites 4u..4u+3 of the colour row (cells 8u + 2v + px)
    uint32_t trial = 0, rej2 = 0, rej4 = 0;
    const unsigned long long rowbase = (unsigned long long)j * a.gX + (unsigned long long)px;
#pragma unroll
    for (int q = 7; q >= 0; q--) {  // masks are built most significant bit first
      const unsigned long long leader = rowbase + 8ull * (unsigned long long)(8 * k + q);
      const uint4 blk = philox4x32_10((uint32_t)leader, (uint32_t)(leader >> 32), a.ev.sweep, (uint32_t)(a.c0 + 1), a.ev.rk);
      const uint32_t w[4] = {blk.x, blk.y, blk.z, blk.w};
#pragma unroll
      for (int v = 3; v >= 0; v--) {
        trial = __funnelshift_l(w[v], trial, 1);  // top bit = trial spin - 1
        const uint32_t w2 = w[v] << 1;            // acceptance uniform (event.cuh: ising)
        shift_in_reject(rej2, nthr2, w2);
        shift_in_reject(rej4, nthr4, w2);
      }
    }
    uint32_t ok = ge2;
    if (!a.ev.t_zero) ok |= (eq1 & ~rej2) | (eq0 & ~rej4);
    const uint32_t differs = MODEL == SPK_ISING_SINGLE ? 0xFFFFFFFFu : trial ^ own;  // ising/single: the trial is the flip
    const uint32_t flip = differs & ok;
    if (flip) mine[t] = own ^ flip;
    accepted += __popc(flip);
  }
  accepted = __reduce_add_sync(0xffffffffu, accepted);
  if ((threadIdx.x & 31) == 0 && accepted) atomicAdd(a.naccept, (unsigned long long)accepted);
}

// int32 lattice (values 1, 2; row-major, X cells per row) <-> bit shadow.  One thread per word.
static __global__ void __launch_bounds__(256) bits2d_pack_kernel(const int32_t *__restrict__ spin, uint32_t *bits, int Y, int W) {
  const long long nwords = (long long)W * Y;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < 2 * nwords; t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t / nwords);
    const long long r = t - (long long)c * nwords;
    const int j = (int)(r / W), k = (int)(r - (long long)j * W);
    const int px = (c + j) & 1;
    const int32_t *row = spin + (long long)j * (64ll * W) + 64 * k + px;
    uint32_t w = 0;
#pragma unroll
    for (int b = 0; b < 32; b += 2) {  // cells 2b+px, 2b+2+px: two sites per 16-byte load
      const int4 v = *reinterpret_cast<const int4 *>(row + 2 * b - px);
      w |= (uint32_t)((px ? v.y : v.x) - 1) << b;
      w |= (uint32_t)((px ? v.w : v.z) - 1) << (b + 1);
    }
    bits[t] = w;
  }
}
static __global__ void __launch_bounds__(256) bits2d_unpack_kernel(const uint32_t *__restrict__ bits, int32_t *spin, int Y, int W) {
  const long long nwords = (long long)W * Y;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < nwords; r += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(r / W), k = (int)(r - (long long)j * W);
    const int p0 = j & 1;  // x parity of colour 0 in row j
    const uint32_t w0 = bits[r], w1 = bits[nwords + r];
    const uint32_t even = p0 ? w1 : w0, odd = p0 ? w0 : w1;  // words of the even-x / odd-x sites
    int4 *row = reinterpret_cast<int4 *>(spin + (long long)j * (64ll * W) + 64 * k);
#pragma unroll
    for (int b = 0; b < 32; b += 2)
      row[b / 2] = make_int4(1 + (int)((even >> b) & 1u), 1 + (int)((odd >> b) & 1u), 1 + (int)((even >> (b + 1)) & 1u),
                             1 + (int)((odd >> (b + 1)) & 1u));
  }
}

}  // namespace spk
