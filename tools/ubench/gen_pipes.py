#!/usr/bin/env python
"""Generates pipes.cu: issue-rate microbenchmarks for the compare instructions the potts/neighonly
candidate list can be built from (which pipe, how many cycles per warp instruction per SMSP)."""
UNROLL = 64  # ops per loop iteration, 4 independent chains

def body(kind):
    L = []
    if kind == "isetp":
        for i in range(UNROLL):
            L.append("setp.eq.or.s32 p%d, %%1, %%%d, p%d;" % (i % 4, 2 + i % 8, i % 4))
    elif kind == "fsetp":
        for i in range(UNROLL):
            L.append("setp.neu.and.f32 p%d, %%1, %%%d, p%d;" % (i % 4, 2 + i % 8, i % 4))
    elif kind == "hsetp2":
        for i in range(UNROLL):
            L.append("setp.ne.and.f16x2 p%d|q%d, %%1, %%%d, p%d;" % (i % 4, i % 4, 2 + i % 8, i % 4))
    elif kind == "hsetp2_plop":
        for i in range(UNROLL // 2):
            L.append("setp.ne.and.f16x2 p%d|q%d, %%1, %%%d, p%d;" % (i % 4, i % 4, 2 + i % 8, i % 4))
            L.append("and.pred p%d, p%d, q%d;" % (i % 4, i % 4, i % 4))
    elif kind == "plop":
        for i in range(UNROLL):
            L.append("and.pred p%d, p%d, q%d;" % (i % 4, (i + 1) % 4, i % 4))
            L.append("xor.pred q%d, p%d, q%d;" % (i % 4, (i + 2) % 4, (i + 1) % 4)) if False else None
        L = [x for x in L if x]
    elif kind == "hset2bm":
        for i in range(UNROLL):
            L.append("set.ne.u32.f16x2 acc%d, acc%d, %%%d;" % (i % 4, i % 4, 2 + i % 8))
        L = [x for x in L if x]
    elif kind == "hset2bm_lop":
        for i in range(UNROLL // 3):
            j = i % 4
            L.append("set.ne.u32.f16x2 r%d, %%1, %%%d;" % (j, 2 + (2 * i) % 8))
            L.append("set.ne.u32.f16x2 s%d, %%1, %%%d;" % (j, 2 + (2 * i + 1) % 8))
            L.append("lop3.b32 acc%d, acc%d, r%d, s%d, 0x80;" % (j, j, j, j))
    elif kind == "lop3":
        for i in range(UNROLL):
            L.append("lop3.b32 acc%d, acc%d, %%%d, %%1, 0x96;" % (i % 4, i % 4, 2 + i % 8))
    elif kind == "imad":
        for i in range(UNROLL):
            L.append("mad.lo.s32 acc%d, acc%d, %%%d, %%1;" % (i % 4, i % 4, 2 + i % 8))
    elif kind == "isetp_hsetp2":
        for i in range(UNROLL // 2):
            L.append("setp.eq.or.s32 p%d, %%1, %%%d, p%d;" % (i % 2, 2 + i % 8, i % 2))
            L.append("setp.ne.and.f16x2 p%d|q%d, %%1, %%%d, p%d;" % (2 + i % 2, i % 2, 2 + (i + 3) % 8, 2 + i % 2))
    elif kind == "isetp_fsetp":
        for i in range(UNROLL // 2):
            L.append("setp.eq.or.s32 p%d, %%1, %%%d, p%d;" % (i % 2, 2 + i % 8, i % 2))
            L.append("setp.neu.and.f32 p%d, %%1, %%%d, p%d;" % (2 + i % 2, 2 + (i + 3) % 8, 2 + i % 2))
    elif kind == "isetp_hset2bm":
        for i in range(UNROLL // 2):
            L.append("setp.eq.or.s32 p%d, %%1, %%%d, p%d;" % (i % 4, 2 + i % 8, i % 4))
            L.append("set.ne.u32.f16x2 r%d, r%d, %%%d;" % (i % 4, i % 4, 2 + (i + 3) % 8))
    elif kind == "isetp_imad":
        for i in range(UNROLL // 2):
            L.append("setp.eq.or.s32 p%d, %%1, %%%d, p%d;" % (i % 4, 2 + i % 8, i % 4))
            L.append("mad.lo.s32 acc%d, acc%d, %%%d, %%1;" % (i % 4, i % 4, 2 + i % 8))
    elif kind == "hset2bm_imad":
        for i in range(UNROLL // 2):
            L.append("set.ne.u32.f16x2 r%d, r%d, %%%d;" % (i % 4, i % 4, 2 + (i + 3) % 8))
            L.append("mad.lo.s32 acc%d, acc%d, %%%d, %%1;" % (i % 4, i % 4, 2 + i % 8))
    elif kind == "hsetp2_imad":
        for i in range(UNROLL // 2):
            L.append("setp.ne.and.f16x2 p%d|q%d, %%1, %%%d, p%d;" % (i % 4, i % 4, 2 + i % 8, i % 4))
            L.append("mad.lo.s32 acc%d, acc%d, %%%d, %%1;" % (i % 4, i % 4, 2 + i % 8))
    elif kind == "plop_imad":
        for i in range(UNROLL // 2):
            L.append("and.pred p%d, p%d, q%d;" % (i % 4, (i + 1) % 4, i % 4))
            L.append("mad.lo.s32 acc%d, acc%d, %%%d, %%1;" % (i % 4, i % 4, 2 + i % 8))
    elif kind == "viaddmnmx":
        for i in range(UNROLL):
            L.append("add.u16x2 r%d, %%1, %%%d;" % (i % 4, 2 + i % 8))
            L.append("min.u16x2 acc%d, r%d, acc%d;" % (i % 4, i % 4, i % 4))
    elif kind == "viaddmnmx_imad":
        for i in range(UNROLL // 2):
            L.append("add.u16x2 r%d, %%1, %%%d;" % (i % 4, 2 + i % 8))
            L.append("min.u16x2 acc%d, r%d, acc%d;" % (i % 4, i % 4, i % 4))
            L.append("mad.lo.s32 s%d, s%d, %%%d, %%1;" % (i % 4, i % 4, 2 + i % 8))
    elif kind == "viaddmnmx_isetp":
        for i in range(UNROLL // 2):
            L.append("add.u16x2 r%d, %%1, %%%d;" % (i % 4, 2 + i % 8))
            L.append("min.u16x2 acc%d, r%d, acc%d;" % (i % 4, i % 4, i % 4))
            L.append("setp.eq.or.s32 p%d, %%1, %%%d, p%d;" % (i % 4, 2 + i % 8, i % 4))
    elif kind == "viaddmnmx_imm_imad":   # third VIADDMNMX operand immediate (2 register sources)
        for i in range(UNROLL // 2):
            L.append("add.u16x2 r%d, acc%d, %%%d;" % (i % 4, i % 4, 2 + i % 8))
            L.append("min.u16x2 acc%d, r%d, cst;" % (i % 4, i % 4))
            L.append("mad.lo.s32 s%d, s%d, %%%d, %%1;" % (i % 4, i % 4, 2 + i % 8))
    elif kind == "viaddmnmx_imad2":   # IMAD with two register sources (multiplier immediate-like constant 3)
        for i in range(UNROLL // 2):
            L.append("add.u16x2 r%d, %%1, %%%d;" % (i % 4, 2 + i % 8))
            L.append("min.u16x2 acc%d, r%d, acc%d;" % (i % 4, i % 4, i % 4))
            L.append("mad.lo.s32 s%d, s%d, 3, %%1;" % (i % 4, i % 4))
    elif kind == "lop3_imad":
        for i in range(UNROLL // 2):
            L.append("lop3.b32 acc%d, acc%d, %%%d, %%1, 0x96;" % (i % 4, i % 4, 2 + i % 8))
            L.append("mad.lo.s32 s%d, s%d, %%%d, %%1;" % (i % 4, i % 4, 2 + i % 8))
    elif kind == "viaddmnmx_ffma":
        for i in range(UNROLL // 2):
            L.append("add.u16x2 r%d, %%1, %%%d;" % (i % 4, 2 + i % 8))
            L.append("min.u16x2 acc%d, r%d, acc%d;" % (i % 4, i % 4, i % 4))
            L.append("fma.rn.f32 s%d, s%d, %%%d, %%1;" % (i % 4, i % 4, 2 + i % 8))
    elif kind == "viaddmnmx_2to1":  # 2 VIADDMNMX : 1 IMAD, the ratio of the dedupe code
        for i in range(21):
            L.append("add.u16x2 r%d, %%1, %%%d;" % (i % 4, 2 + i % 8))
            L.append("min.u16x2 acc%d, r%d, acc%d;" % (i % 4, i % 4, i % 4))
            L.append("add.u16x2 r%d, %%1, %%%d;" % ((i+1) % 4, 2 + (i+1) % 8))
            L.append("min.u16x2 acc%d, r%d, acc%d;" % ((i+1) % 4, (i+1) % 4, (i+1) % 4))
            L.append("mad.lo.s32 s%d, s%d, %%%d, %%1;" % (i % 4, i % 4, 2 + i % 8))
    elif kind == "match":
        for i in range(UNROLL // 8):
            L.append("match.any.sync.b32 r%d, acc%d, 0xffffffff;" % (i % 4, i % 4))
            L.append("add.u32 acc%d, acc%d, r%d;" % (i % 4, i % 4, i % 4))
    return L

KINDS = ["isetp", "fsetp", "hsetp2", "hsetp2_plop", "plop", "hset2bm", "hset2bm_lop", "lop3", "imad",
         "isetp_hsetp2", "isetp_fsetp", "isetp_hset2bm", "isetp_imad", "hset2bm_imad", "hsetp2_imad", "plop_imad", "viaddmnmx", "viaddmnmx_imad", "viaddmnmx_isetp", "viaddmnmx_imm_imad", "viaddmnmx_imad2", "lop3_imad", "viaddmnmx_ffma", "viaddmnmx_2to1", "match"]

print('#include <cstdio>\n#include <cstdint>\n#include <cuda_runtime.h>')
for k in KINDS:
    ops = body(k)
    print("__global__ void __launch_bounds__(256) k_%s(const uint32_t *in, uint32_t *out, int iters, long long *cyc) {" % k)
    print("  uint32_t a = in[threadIdx.x], b[8], r = 0;")
    print("  for (int i = 0; i < 8; i++) b[i] = in[256 + i + (threadIdx.x & 1)];")
    print("  long long t0 = clock64();")
    print("  asm volatile(\"{\\n\\t.reg .pred p<4>, q<4>;\\n\\t.reg .b32 r<4>, s<4>, acc<4>;\\n\\t.reg .s32 it;\\n\\t.reg .b32 cst;\\n\\tmov.b32 cst, 0x00010001;\\n\\t\"")
    for j in range(4):
        print('      "setp.eq.s32 p%d, %%1, %d;\\n\\tsetp.eq.s32 q%d, %%2, %d;\\n\\tmov.b32 acc%d, %%%d;\\n\\tmov.b32 r%d, 0;\\n\\tmov.b32 s%d, 0;\\n\\t"' % (j, j + 77, j, j + 55, j, 2 + j, j, j))
    print('      "mov.s32 it, %%10;\\n\\tLOOP_%s:\\n\\t"' % k)
    for o in ops:
        print('      "%s\\n\\t"' % o)
    print('      "add.s32 it, it, -1;\\n\\tsetp.gt.s32 q0, it, 0;\\n\\t@q0 bra LOOP_%s;\\n\\t"' % k)
    print('      "mov.b32 %0, 0;\\n\\t"')
    for j in range(4):
        print('      "@p%d add.u32 %%0, %%0, %d;\\n\\t@q%d add.u32 %%0, %%0, %d;\\n\\tadd.u32 %%0, %%0, acc%d;\\n\\tadd.u32 %%0, %%0, r%d;\\n\\tadd.u32 %%0, %%0, s%d;\\n\\t"' % (j, 1 << j, j, 16 << j, j, j, j))
    print('      "}" : "=r"(r) : "r"(a), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(iters));')
    print("  long long t1 = clock64();")
    print("  out[blockIdx.x * blockDim.x + threadIdx.x] = r;")
    print("  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;")
    print("}")
print("int main() {")
print("  uint32_t *in, *out; long long *cyc; int nb = 148 * 2, iters = 2000;")
print("  cudaMalloc(&in, 4096); cudaMemset(in, 1, 4096); cudaMalloc(&out, nb * 256 * 4); cudaMalloc(&cyc, nb * 8);")
print("  long long *h = new long long[nb];")
for k in KINDS:
    n = len(body(k))
    if k.startswith("viaddmnmx"): n = sum(1 for o in body(k) if not o.startswith("add.u16x2"))
    print("  for (int rep = 0; rep < 2; rep++) { k_%s<<<nb, 256>>>(in, out, iters, cyc); cudaDeviceSynchronize(); }" % k)
    print("  cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);")
    print("  { double s = 0; for (int i = 0; i < nb; i++) s += h[i]; s /= nb;")
    print("    // 512 threads/SM = 16 warps = 4 warps per SMSP")
    print('    printf("%%-16s ops/iter %%3d  cycles/warp-instr/SMSP %%.3f  (err %%s)\\n", "%s", %d, s / ((double)iters * %d * 4), cudaGetErrorString(cudaGetLastError())); }' % (k, n, n))
print("  return 0; }")
