// Streaming colour-set kernel for sc/26n potts / potts/pin on a 16-BIT copy of the lattice.
//
// Same pipeline as sweep_stream.cuh (persistent CTAs, one TMA box per lattice plane into a
// shared-memory ring, mbarrier hand-over, periodic halo patching), but the spins live in HBM as
// uint16 (every spin <= 65534, the condition of the packed compares anyway): a colour launch then
// moves HALF the bytes of the int32 kernel, which is bound by exactly those bytes (5.4 TB/s of DRAM
// traffic at 1024^3).  The library keeps the int32 lattice as the canonical copy and this 16-bit
// shadow next to it, converting lazily at the boundaries of spk_cuda_sweeps (spk_cuda.cu: shadow16).
//
// A thread owns a 16-byte chunk = 8 sites = 4 class sites of a class row.  Two adjacent sites share
// a 32-bit word, which is already the operand format of the packed compare (event.cuh): the
// negated word -w tests BOTH its sites against a replicated value in one VIADDMNMX, and the `min`
// operand masks the half that is not a neighbour (cap 0 instead of 1), so no packing arithmetic is
// needed beyond one negation per word, shared by the sites around it.
#pragma once
#include "sweep_stream.cuh"

namespace spk {

constexpr int S16_TX = 256;          // tile width in sites: 32 lanes x one 16-byte chunk of 8
constexpr int S16_BX = S16_TX + 16;  // TMA box inner extent: an 8-element (16-byte) halo chunk per side
constexpr int S16_CR = 16;           // class rows per tile: 8 warps x 2

struct Stream16Cfg {                 // sc/26n only
  static constexpr int YT = S16_CR * 2;
  static constexpr int RP = YT + 2;
  static constexpr int RING = 5;
  static constexpr int PLANE_BYTES = S16_BX * RP * 2;
  static constexpr int PLANE_STRIDE = (PLANE_BYTES + 127) / 128 * 128;
  static constexpr int BAR_OFF = RING * PLANE_STRIDE;
  static constexpr int SMEM_BYTES = BAR_OFF + 64;
};

struct Sweep16Args {
  uint16_t *spin;       // first owned site of the 16-bit lattice (ghost planes at negative offsets)
  Geo g;
  int c0;               // colour - 1
  EventParams ev;
  unsigned long long *naccept;
  uint16_t *mirror_lo, *mirror_hi;  // the ring neighbours' ghost planes (peer memory), or null
  int minus_one;                      // -1 (negation as IMAD on the FMA pipe)
};

// number of the 26 neighbours of the class site in word `i` of the chunk that differ from `value`.
// N[row][5]: negated words of the stencil rows; P = 0: [0] = word left of the chunk, [1..4] = the
// chunk, the site is the LOW half of word 1+i; P = 1: [0..3] = the chunk, [4] = word right of it,
// the site is the HIGH half of word i.
template <int P>
__device__ __forceinline__ int unlike26_16(const uint32_t (&N)[9][5], int i, int value, const Sweep16Args &a) {
  const uint32_t rep = pk_rep(value, a.ev);
  // the 0/1 halfwords of the 18 compares are summed three at a time (IADD3): half the instructions
  // of a multiply-add chain, and the kernel is bound by issue slots, not by one pipe
  uint32_t t[18];
#pragma unroll
  for (int rr = 0; rr < 9; rr++) {
    // main word: the site's own word (both halves are neighbours, except in the site's own row)
    // side word: holds the one neighbour beyond the own word
    // (the caps are literals: as immediates of the instruction they cost nothing; passed through
    // kernel parameters ptxas re-materialised them with one PRMT per compare)
    constexpr uint32_t CAP_LO = 0x00000001u, CAP_HI = 0x00010000u, CAP_BOTH = 0x00010001u;
    t[2 * rr] = __viaddmin_u16x2(rep, N[rr][P ? i : 1 + i], rr == 4 ? (P ? CAP_LO : CAP_HI) : CAP_BOTH);
    t[2 * rr + 1] = __viaddmin_u16x2(rep, N[rr][P ? i + 1 : i], P ? CAP_LO : CAP_HI);
  }
  uint32_t e[2] = {0u, 0u};
#pragma unroll
  for (int q = 0; q < 18; q += 6) {
    e[0] += t[q] + t[q + 1] + t[q + 2];
    e[1] += t[q + 3] + t[q + 4] + t[q + 5];
  }
  const uint32_t s = e[0] + e[1];
  return (int)((s & 0xFFFFu) + (s >> 16));
}

// AppPotts / AppPottsPin::site_event_rejection in Philox mode (event.cuh: site_event, same draws)
template <int MODEL, int P>
__device__ __forceinline__ int potts16_event(const Sweep16Args &a, const uint32_t (&N)[9][5], int i, int old, uint32_t w1,
                                             uint32_t w2) {
  const int Q = a.ev.nspins;
  if (MODEL == SPK_POTTS_PIN && old > Q) return old;  // pinned: no event (src/app_potts_pin.cpp:97)
  const int e0 = unlike26_16<P>(N, i, old, a);
  const int trial = (int)__umulhi((uint32_t)Q, w1) + 1;
  if (e0 == 26) return trial;
  const int e1 = unlike26_16<P>(N, i, trial, a);
  if (e1 <= e0) return trial;
  if (a.ev.t_zero) return old;
  return w2 > a.ev.thr[e1 - e0] ? old : trial;
}

// one class row of one class plane: the thread's chunk holds four class sites at P, P+2, P+4, P+6
template <int MODEL, int P>
__device__ __forceinline__ int stream16_chunk(const Sweep16Args &a, const uint16_t *const (&pl)[3], int r, int lane,
                                              uint16_t *gchunk, uint16_t *gmirror, unsigned long long gsite0) {
  uint32_t N[9][5];
  uint4 own;
#pragma unroll
  for (int rr = 0; rr < 9; rr++) {
    const int dy = rr % 3 - 1, dz = rr / 3 - 1;
    const uint16_t *rowp = pl[dz + 1] + (r + dy) * S16_BX + 8 + 8 * lane;
    const uint4 t = *reinterpret_cast<const uint4 *>(rowp);
    const uint32_t side = *reinterpret_cast<const uint32_t *>(P ? rowp + 8 : rowp - 2);
    if (rr == 4) own = t;
    const uint32_t w[5] = {P ? t.x : side, P ? t.y : t.x, P ? t.z : t.y, P ? t.w : t.z, P ? side : t.w};
#pragma unroll
    for (int q = 0; q < 5; q++) N[rr][q] = w[q] * (uint32_t)a.minus_one;  // -w: low half -lo, high half -hi-1 (pk_negpair)
  }
  uint32_t ow[4] = {own.x, own.y, own.z, own.w};
  int accepted = 0;
  bool changed = false;
#pragma unroll
  for (int pr = 0; pr < 2; pr++) {  // one Philox block per pair of class sites (cells x, x^2: event.cuh)
    const unsigned long long leader = gsite0 + P + 4 * pr;
    const uint4 blk = philox4x32_10((uint32_t)leader, (uint32_t)(leader >> 32), a.ev.sweep, (uint32_t)(a.c0 + 1), a.ev.rk);
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int i = 2 * pr + h;
      const int old = (int)(P ? ow[i] >> 16 : ow[i] & 0xFFFFu);
      const int nw = potts16_event<MODEL, P>(a, N, i, old, h ? blk.z : blk.x, h ? blk.w : blk.y);
      if (nw != old) {
        accepted++;
        changed = true;
        ow[i] = P ? (ow[i] & 0x0000FFFFu) | ((uint32_t)nw << 16) : (ow[i] & 0xFFFF0000u) | (uint32_t)nw;
      }
    }
  }
  if (changed) {
    const uint4 nv = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    *reinterpret_cast<uint4 *>(gchunk) = nv;
    if (gmirror) *reinterpret_cast<uint4 *>(gmirror) = nv;  // the neighbour's ghost copy of this chunk
  }
  return accepted;
}

// Periodic images of the halo cells outside the lattice (see sweep_stream.cuh: PatchPlan), patched
// as 32-bit words of two sites (tile origin and row length are even): the columns x0-2,x0-1 and
// the first word past the row end (threads 0..RP-1: one row each), the rows y0-1 and Y (threads
// 64..64+BX/2-1: one word each).  Whole words, because the packed compare needs the LOW half of
// every word it touches to be a real spin (>= 1): its borrow is part of the encoding (pk_negpair).
struct Patch16 {
  int col_l, col_r, row_t, row_b;
};

template <int MODEL>
__global__ void __launch_bounds__(256, 2)
sweep_stream16_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Sweep16Args a,
                      const __grid_constant__ StreamMap sm) {
  using C = Stream16Cfg;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + C::BAR_OFF);
  auto plane_ptr = [&](uint32_t slot) { return reinterpret_cast<uint16_t *>(smem_raw + slot * C::PLANE_STRIDE); };

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const Geo &g = a.g;
  const int rowlen = g.X, Y = g.Y;
  const long long sy = rowlen, sz = sy * Y;

  if (tid == 0) {
    for (int i = 0; i < C::RING; i++) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto zmap = [&](int kk) { return kk < 0 ? g.lo_wrap : (kk >= g.Z ? g.hi_wrap : kk); };
  auto adv = [](uint32_t sl, uint32_t d) { sl += d; return sl >= (uint32_t)C::RING ? sl - C::RING : sl; };
  uint32_t slot = 0, phase = 0;
  auto issue = [&](int kk, uint32_t sl, int x0, int y0) {
    uint64_t *bar = &full[sl];
    mbar_expect_tx(bar, C::PLANE_BYTES);
    // the tensor is described in 32-bit words (two sites): box dimensions are limited to 256 elements
    tma_load_3d(plane_ptr(sl), &tmap, bar, (x0 - 8) / 2, y0 - 1, zmap(kk) + 1);  // tensor plane 0 = lower ghost plane
  };
  auto wait = [&](uint32_t sl) {
    mbar_wait(&full[sl], (phase >> sl) & 1u);
    phase ^= 1u << sl;
  };
  // patch values of one plane: [0],[1] = the two columns (row threads) or the two rows' words (column threads)
  auto patch_load = [&](const Patch16 &pp, const uint16_t *gplane, int x0, int y0, uint32_t (&v)[2]) {
    v[0] = v[1] = 0;
    if (tid < C::RP) {
      int y = y0 - 1 + tid;
      y = y < 0 ? Y - 1 : (y >= Y ? y - Y : y);
      if (y < Y) {
        const uint16_t *row = gplane + (long long)y * rowlen;
        if (pp.col_l >= 0) v[0] = *reinterpret_cast<const uint32_t *>(row + rowlen - 2);
        if (pp.col_r >= 0) v[1] = *reinterpret_cast<const uint32_t *>(row);
      }
    } else if (tid >= 64 && tid < 64 + S16_BX / 2) {
      int x = x0 - 8 + 2 * (tid - 64);
      x = x < 0 ? x + rowlen : (x >= rowlen ? x - rowlen : x);
      if (x >= 0 && x < rowlen) {
        if (pp.row_t >= 0) v[0] = *reinterpret_cast<const uint32_t *>(gplane + (long long)(Y - 1) * rowlen + x);
        if (pp.row_b >= 0) v[1] = *reinterpret_cast<const uint32_t *>(gplane + x);
      }
    }
  };
  auto patch_store = [&](const Patch16 &pp, uint16_t *pl, const uint32_t (&v)[2]) {
    if (tid < C::RP) {
      if (pp.col_l >= 0) *reinterpret_cast<uint32_t *>(pl + tid * S16_BX + pp.col_l) = v[0];
      if (pp.col_r >= 0) *reinterpret_cast<uint32_t *>(pl + tid * S16_BX + pp.col_r) = v[1];
    } else if (tid >= 64 && tid < 64 + S16_BX / 2) {
      if (pp.row_t >= 0) *reinterpret_cast<uint32_t *>(pl + pp.row_t * S16_BX + 2 * (tid - 64)) = v[0];
      if (pp.row_b >= 0) *reinterpret_cast<uint32_t *>(pl + pp.row_b * S16_BX + 2 * (tid - 64)) = v[1];
    }
  };

  int accepted = 0;
  const bool mirrors = a.mirror_lo != nullptr || a.mirror_hi != nullptr;
  const unsigned long long plane_sites = (unsigned long long)g.gY * (unsigned long long)g.gX;
  const int P = a.c0 & 1;
  for (int item = blockIdx.x; item < sm.items; item += gridDim.x) {
    const int tx = item % sm.ntx, ty = (item / sm.ntx) % sm.nty, zc = item / (sm.ntx * sm.nty);
    const int x0 = tx * S16_TX, y0 = ty * C::YT;
    const int q0 = zc * sm.qchunk, q1 = min(q0 + sm.qchunk, sm.nq), nsteps = q1 - q0;
    const int kstart = sm.kfirst + q0 * 2;
    const bool edge = x0 == 0 || x0 + S16_TX >= rowlen || y0 == 0 || y0 + C::YT >= Y;

    Patch16 pp;
    pp.col_l = x0 == 0 ? 6 : -1;  // word (x0-2, x0-1)
    const int past = rowlen - x0;  // tile-relative x of the first column past the row end
    pp.col_r = past <= S16_TX ? 8 + past : -1;
    pp.row_t = y0 == 0 ? 0 : -1;
    const int rb = Y - y0 + 1;
    pp.row_b = rb <= C::RP - 1 ? rb : -1;

    if (tid == 0)
      for (int p = 0; p < 3; p++) issue(kstart - 1 + p, adv(slot, p), x0, y0);
    if (edge) {
      uint32_t pv0[3][2];
      for (int p = 0; p < 3; p++) patch_load(pp, a.spin + (long long)zmap(kstart - 1 + p) * sz, x0, y0, pv0[p]);
      for (int p = 0; p < 3; p++) {
        wait(adv(slot, p));
        patch_store(pp, plane_ptr(adv(slot, p)), pv0[p]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
    }

    const int x = x0 + 8 * lane;
    int yrow[2];
    bool live[2];
    uint16_t *gptr[2];
    unsigned long long gsite[2];
#pragma unroll
    for (int t = 0; t < 2; t++) {
      const int cr = warp + 8 * t;
      yrow[t] = y0 + 2 * cr + ((a.c0 >> 1) & 1);
      live[t] = yrow[t] < Y && x < rowlen;
      gptr[t] = a.spin + ((long long)kstart * sz + (long long)yrow[t] * sy + x);
      gsite[t] = ((unsigned long long)(kstart + g.g0) * g.gY + yrow[t]) * (unsigned long long)g.gX + x;
    }

    uint32_t pv[2][2];
    for (int s = 0; s < nsteps; s++) {
      const int k = kstart + 2 * s;
      const uint32_t w0 = slot, w1 = adv(slot, 1), w2 = adv(slot, 2);
      if (s > 0) {
        wait(w1); wait(w2);
        if (edge) {
          patch_store(pp, plane_ptr(w1), pv[0]);
          patch_store(pp, plane_ptr(w2), pv[1]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncthreads();
        }
      } else if (!edge) {
        wait(w0); wait(w1); wait(w2);
      }
      if (s + 1 < nsteps) {
        if (tid == 0)
          for (int p = 0; p < 2; p++) issue(k + 2 + p, adv(slot, 3 + p), x0, y0);
        if (edge)
          for (int p = 0; p < 2; p++) patch_load(pp, a.spin + (long long)zmap(k + 2 + p) * sz, x0, y0, pv[p]);
      }

      const uint16_t *const pl[3] = {plane_ptr(w0), plane_ptr(w1), plane_ptr(w2)};
      long long moff = 0;
      bool mirror = false;
      if (mirrors) {
        const long long here = reinterpret_cast<long long>(a.spin + (long long)k * sz);
        if (k == 0 && a.mirror_lo) { mirror = true; moff = reinterpret_cast<long long>(a.mirror_lo) - here; }
        if (k == g.Z - 1 && a.mirror_hi) { mirror = true; moff = reinterpret_cast<long long>(a.mirror_hi) - here; }
      }
#pragma unroll
      for (int t = 0; t < 2; t++) {
        uint16_t *gchunk = gptr[t];
        uint16_t *gmirror = mirror ? reinterpret_cast<uint16_t *>(reinterpret_cast<long long>(gchunk) + moff) : nullptr;
        if (live[t]) {
          const int r = yrow[t] - y0 + 1;
          if (P) accepted += stream16_chunk<MODEL, 1>(a, pl, r, lane, gchunk, gmirror, gsite[t]);
          else accepted += stream16_chunk<MODEL, 0>(a, pl, r, lane, gchunk, gmirror, gsite[t]);
        }
        gptr[t] += 2 * sz;
        gsite[t] += 2ull * plane_sites;
      }
      __syncthreads();
      if (s + 1 < nsteps) slot = adv(slot, 2);
    }
    slot = adv(slot, 3);
  }
  accepted = __reduce_add_sync(0xffffffffu, accepted);
  if (lane == 0 && accepted) atomicAdd(a.naccept, (unsigned long long)accepted);
}

// int32 lattice <-> 16-bit shadow (whole allocations incl. ghost planes; n a multiple of 8)
static __global__ void __launch_bounds__(256) shadow_pack_kernel(const int4 *__restrict__ src, uint4 *dst, long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const int4 a = src[2 * i], b = src[2 * i + 1];
    dst[i] = make_uint4((uint32_t)a.x | ((uint32_t)a.y << 16), (uint32_t)a.z | ((uint32_t)a.w << 16),
                        (uint32_t)b.x | ((uint32_t)b.y << 16), (uint32_t)b.z | ((uint32_t)b.w << 16));
  }
}
static __global__ void __launch_bounds__(256) shadow_unpack_kernel(const uint4 *__restrict__ src, int4 *dst, long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const uint4 s = src[i];
    dst[2 * i] = make_int4((int)(s.x & 0xFFFFu), (int)(s.x >> 16), (int)(s.y & 0xFFFFu), (int)(s.y >> 16));
    dst[2 * i + 1] = make_int4((int)(s.z & 0xFFFFu), (int)(s.z >> 16), (int)(s.w & 0xFFFFu), (int)(s.w >> 16));
  }
}

}  // namespace spk
