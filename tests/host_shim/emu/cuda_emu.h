// TEST INFRASTRUCTURE.  A small CUDA-on-CPU emulator: just enough of the CUDA C++ language
// extensions, the runtime API and the sm_90+ TMA / mbarrier primitives to compile the UNCHANGED
// sources of libspparks_cuda (spparks_b200/csrc) with g++ and run its kernels on host threads
// (tests/host_shim/build_emu.py; tests/test_emulated_library.py).
//
// Execution model: the CTAs of a launch are handed to a few worker OS threads; the threads of a
// CTA are fibers (ucontext) of its worker, switched cooperatively wherever a CUDA thread can block:
// __syncthreads, warp collectives, mbarrier waits.  A CTA therefore runs deterministically, with
// the blocking points exactly where they are on the GPU; CTAs run concurrently, as on the GPU.
// Streams are synchronous (every call completes before it returns), which is a legal schedule of
// the CUDA stream semantics.  What the emulator cannot show: timing, data races inside a CTA,
// anything that depends on the asynchrony of TMA or of peer-memory stores.
//
// Each primitive is restated from its definition in the CUDA programming guide / PTX ISA.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <type_traits>

#define SPK_HOST_SHIM 1
#define SPK_CUDA_EMU 1

// ---- language extensions ------------------------------------------------------------------------
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static thread_local  // one copy per worker = per running CTA
#define __align__(n) alignas(n)

struct dim3 {
  unsigned x, y, z;
  constexpr dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(8) uint2 { unsigned x, y; };
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }

namespace emu {
struct Tls {
  dim3 tid, bid;
  int lin = 0, lane = 0, warp = 0;  // linear thread index in the CTA, lane, warp
};
extern thread_local Tls *tlsp;  // the running fiber's indices
extern dim3 bdim, gdim;

void syncthreads();
uint64_t warp_exchange(uint64_t mine, int src_lane);  // every lane deposits, then reads src_lane's value
void warp_gather(uint64_t mine, uint64_t (&all)[32]);  // every lane deposits, everybody reads all 32
unsigned char *dynamic_smem();                          // base of the CTA's dynamic shared memory
template <class F> static inline const void *fn_addr(F f) { return reinterpret_cast<const void *>(f); }
void launch(const char *kernel, const void *fn, dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &thread_body);
uint64_t globaltimer();

// mbarrier with transaction count (PTX ISA 9.7.13.15) for arrival count 1, and the bulk tensor copy
void mbar_init(uint64_t *bar, int count);
void mbar_expect_tx(uint64_t *bar, uint32_t bytes);
void mbar_wait(uint64_t *bar, uint32_t parity);
void tma_load(void *dst, const void *tensor_map, uint64_t *bar, const int *coord, int rank);
}  // namespace emu

#define threadIdx (::emu::tlsp->tid)
#define blockIdx (::emu::tlsp->bid)
#define blockDim (::emu::bdim)
#define gridDim (::emu::gdim)

static inline void __syncthreads() { ::emu::syncthreads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { uint64_t a[32]; ::emu::warp_gather(0, a); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
void __nanosleep(unsigned ns);

// ---- integer intrinsics ---------------------------------------------------------------------------
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline uint32_t __viaddmin_u16x2(uint32_t a, uint32_t b, uint32_t c) {  // per halfword min(a + b mod 2^16, c)
  const uint32_t lo = (uint16_t)((a & 0xFFFFu) + (b & 0xFFFFu)), hi = (uint16_t)((a >> 16) + (b >> 16));
  const uint32_t clo = c & 0xFFFFu, chi = c >> 16;
  return (lo < clo ? lo : clo) | ((hi < chi ? hi : chi) << 16);
}
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s) {
  s &= 31;
  return s ? (hi << s) | (lo >> (32 - s)) : hi;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) {
  s &= 31;
  return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline size_t __cvta_generic_to_shared(const void *p) {
  return (size_t)((const unsigned char *)p - ::emu::dynamic_smem());
}

// ---- warp collectives (all 32 lanes take part: the kernels only use full masks) -----------------
template <class T> static inline uint64_t emu_bits(T v) {
  static_assert(sizeof(T) <= 8, "");
  uint64_t b = 0;
  memcpy(&b, &v, sizeof(T));
  return b;
}
template <class T> static inline T emu_unbits(uint64_t b) {
  T v;
  memcpy(&v, &b, sizeof(T));
  return v;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) {
  return emu_unbits<T>(::emu::warp_exchange(emu_bits(v), src & 31));
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d) {
  const int src = ::emu::tlsp->lane + (int)d;
  return emu_unbits<T>(::emu::warp_exchange(emu_bits(v), src < 32 ? src : ::emu::tlsp->lane));
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d) {
  const int src = ::emu::tlsp->lane - (int)d;
  return emu_unbits<T>(::emu::warp_exchange(emu_bits(v), src >= 0 ? src : ::emu::tlsp->lane));
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) {
  return emu_unbits<T>(::emu::warp_exchange(emu_bits(v), (::emu::tlsp->lane ^ m) & 31));
}
template <class T> static inline T emu_reduce_add(T v) {
  uint64_t all[32];
  ::emu::warp_gather(emu_bits(v), all);
  T s = 0;
  for (int l = 0; l < 32; l++) s += emu_unbits<T>(all[l]);
  return s;
}
static inline int __reduce_add_sync(unsigned, int v) { return emu_reduce_add<int>(v); }
static inline unsigned __reduce_add_sync(unsigned, unsigned v) { return emu_reduce_add<unsigned>(v); }
static inline unsigned __ballot_sync(unsigned, int pred) {
  uint64_t all[32];
  ::emu::warp_gather(pred ? 1 : 0, all);
  unsigned m = 0;
  for (int l = 0; l < 32; l++) m |= (unsigned)(all[l] & 1) << l;
  return m;
}
template <class T> static inline unsigned __match_any_sync(unsigned, T v) {
  uint64_t all[32];
  ::emu::warp_gather(emu_bits(v), all);
  unsigned m = 0;
  for (int l = 0; l < 32; l++) m |= (unsigned)(all[l] == all[::emu::tlsp->lane]) << l;
  return m;
}

// ---- atomics (CTAs run one at a time, the threads of a CTA concurrently) ---------------------------
template <class T> static inline T emu_atomic_add(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicAdd(int *p, int v) { return emu_atomic_add(p, v); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return emu_atomic_add(p, v); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return emu_atomic_add(p, v); }
static inline double atomicAdd(double *p, double v) {
  uint64_t *q = reinterpret_cast<uint64_t *>(p);
  uint64_t old = __atomic_load_n(q, __ATOMIC_SEQ_CST);
  for (;;) {
    const uint64_t nw = emu_bits(emu_unbits<double>(old) + v);
    if (__atomic_compare_exchange_n(q, &old, nw, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) return emu_unbits<double>(old);
  }
}
template <class T> static inline T emu_atomic_min(T *p, T v) {
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
  }
  return old;
}
static inline unsigned atomicMin(unsigned *p, unsigned v) { return emu_atomic_min(p, v); }
static inline int atomicMin(int *p, int v) { return emu_atomic_min(p, v); }
static inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) { return emu_atomic_min(p, v); }
