// microbench.cuh -- INT32 multiply issue-rate probes (the roofline denominator of the
// 768-bit path, SURVEY.md section 8d: "P_mac = measured wide-MAC/s of the chip from a
// dependent-chain-free microbenchmark run in the same job").
//
// Every probe keeps 8 independent accumulators per thread, 64 warps per SM, and multiplies by
// DISTINCT multiplicand registers: with loop-invariant identical operands ptxas computes the
// product once and turns the loop into IADD3s (the first version of this probe measured
// exactly that and reported twice the real rate).  tests/test_build_artifacts.py checks in the
// SASS that each probe's loop still contains the multiplies it claims to time.
//   MODE 0  IMAD        mad.lo.u32              32x32 -> low 32 + 32
//   MODE 1  IMAD.HI     mad.hi.u32              32x32 -> high 32 + 32   (multiplicand loop-variant)
//   MODE 2  IMAD.WIDE   mad.lo.cc + madc.hi     32x32 -> 64 + 64, accumulate form, distinct multiplicands
//                       <- "wide MAC" peak: the roofline denominator
//   MODE 3  IMAD.WIDE   same, one shared multiplicand pair (operand-reuse friendly)
//   MODE 4  IMAD.WIDE.U32.X  carry chains of 8 links (the form the CIOS product issues)
//   MODE 5  IADD3.X     add.cc / addc.cc chains of 8
//   MODE 6  DFMA        fma.rz.f64, 16 independent accumulators (what a 52-bit-limb product would issue)
//   MODE 7  IMAD.WIDE + DFMA interleaved 8 + 8: do the two pipes run side by side?
#pragma once
#include <cstdint>

namespace gsn {

constexpr int INT32_PROBE_MODES = 8;
constexpr int INT32_PROBE_UNROLL = 4;  // repetitions of the 8-accumulator group per loop iteration

template <int MODE>
__global__ void __launch_bounds__(256) int32_issue_probe(uint32_t *sink, const uint32_t *in, int iters) {
    uint32_t a[8], lo[8], hi[8];
    uint32_t b0 = in[threadIdx.x & 31], b1 = in[32 + (threadIdx.x & 31)];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = in[64 + i + threadIdx.x]; lo[i] = a[i] * 3u + i; hi[i] = a[i] * 5u + i; }
    double d[16], e[8];
    if (MODE >= 6) {
#pragma unroll
        for (int i = 0; i < 16; ++i) d[i] = (double)(in[128 + i + threadIdx.x] & 0xFFFFF) * 1e-9;
#pragma unroll
        for (int i = 0; i < 8; ++i) e[i] = 1.0 + (double)(a[i] & 0xFFFF) * 1e-12;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < INT32_PROBE_UNROLL; ++rep) {
            const uint32_t b = (rep & 1) ? b1 : b0;
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[i]) : "r"(a[i]), "r"(b));
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi[i]) : "r"(a[i]), "r"(b1));
                }
            } else if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(lo[i]) : "r"(b), "r"(a[i]));
                    asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(hi[i]) : "r"(b1), "r"(a[i]));
                }
            } else if (MODE == 2) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(a[i]), "r"(b));
            } else if (MODE == 3) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(a[0]), "r"(b));
            } else if (MODE == 4) {
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[0]), "+r"(hi[0]) : "r"(a[0]), "r"(b));
#pragma unroll
                for (int i = 1; i < 8; ++i)
                    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(a[i]), "r"(b));
                asm volatile("addc.u32 %0, %0, 0;" : "+r"(b1));
            } else if (MODE == 6) {
#pragma unroll
                for (int i = 0; i < 16; ++i) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(e[i & 7]), "d"(e[(i + rep) & 7]));
            } else if (MODE == 7) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(a[i]), "r"(b));
                    asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(e[i]), "d"(e[(i + rep) & 7]));
                }
            } else {
                asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(lo[0]) : "r"(a[0]));
#pragma unroll
                for (int i = 1; i < 8; ++i) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(lo[i]) : "r"(a[i]));
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(hi[i]) : "r"(a[i]));
                asm volatile("addc.u32 %0, %0, 0;" : "+r"(b1));
            }
        }
    }
    uint32_t x = b1;
#pragma unroll
    for (int i = 0; i < 8; ++i) x ^= lo[i] ^ hi[i];
    if (MODE >= 6) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x ^= (uint32_t)__double2hiint(d[i]) ^ (uint32_t)__double2loint(d[i]);
    }
    if (x == 0x12345678u) sink[0] = x;  // never true in practice; keeps the work alive
}

// timed instructions per thread per loop iteration
__host__ inline int int32_probe_ops_per_iter(int mode) {
    return INT32_PROBE_UNROLL * ((mode == 0 || mode == 1 || mode == 5 || mode == 6 || mode == 7) ? 16 : 8);
}

}  // namespace gsn
