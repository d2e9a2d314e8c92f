// TEST INFRASTRUCTURE.  Host stand-ins for the handful of CUDA qualifiers, vector types and
// integer intrinsics that spparks_b200/csrc/event.cuh uses, so that the DEVICE source of the
// per-site events (site_event: all-pairs chain with 32-bit and packed 16-bit compares,
// site_event_nb26: candidate elimination) can be compiled with g++ and checked against the
// oracle on the CPU (tests/test_device_arithmetic_on_host.py).  Each intrinsic is restated from
// its CUDA definition.
#pragma once
#include <stdint.h>

#define SPK_HOST_SHIM 1
#define __device__
#define __host__
#define __forceinline__ inline

struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }

static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
// per halfword: min(a + b (mod 2^16), c), unsigned  (CUDA: __viaddmin_u16x2)
static inline uint32_t __viaddmin_u16x2(uint32_t a, uint32_t b, uint32_t c) {
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
