// microbench.cuh -- INT32 multiply issue-rate probes (the roofline denominator of the
// 768-bit path, SURVEY.md section 8d: "P_mac = measured wide-MAC/s of the chip from a
// dependent-chain-free microbenchmark run in the same job").
#pragma once
#include <cstdint>

namespace gsn {

// MODE 0: mad.lo.u32      (IMAD)             16 independent accumulators
// MODE 1: mad.hi.u32      (IMAD.HI.U32)      16 independent accumulators
// MODE 2: mad.wide.u32    (IMAD.WIDE.U32)     8 independent 64-bit accumulators
// MODE 3: mad.lo.cc/madc.hi.cc chains (IMAD.WIDE.U32.X), 2 chains x 4 links, as in fp768.cuh
// MODE 4: MODE 2 with one IADD3 per wide MAC (checks that the ALU pipe issues alongside)
// MODE 5: mad.hi.u32 with a loop-variant multiplicand (IMAD.HI.U32 that cannot be hoisted)
// MODE 6: mad.lo.cc / madc.lo.cc chains of 8 (32-bit IMAD with carry in/out)
// MODE 7: mad.hi.cc / madc.hi.cc chains of 8 (IMAD.HI with carry in/out)
// MODE 8: add.cc / addc.cc chains of 8 (IADD3.X)
// MODE 9: one chain of 8 wide links (long carry chain, as in one row of the CIOS product)
// MODE 10: mad.lo.u32 + mad.hi.u32 on the same operands, no carries (split wide product)
// MODE 11: mad.wide.u32 with loop-variant multiplicand (rules out hoisting in MODE 2)
template <int MODE>
__global__ void __launch_bounds__(256) int32_issue_probe(uint32_t *sink, uint32_t seed, int iters) {
    uint32_t a = seed ^ (threadIdx.x * 2654435761u), b = seed * 40503u + blockIdx.x;
    uint32_t acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = a + i * 977u;
    uint32_t side = b;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b));
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b));
        } else if (MODE == 2 || MODE == 4) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                asm volatile("{ .reg .u64 t; mov.b64 t, {%0, %1}; mad.wide.u32 t, %2, %3, t; mov.b64 {%0, %1}, t; }"
                             : "+r"(acc[i]), "+r"(acc[i + 1]) : "r"(a), "r"(b));
                if (MODE == 4) asm volatile("add.u32 %0, %0, %1;" : "+r"(side) : "r"(a));
            }
        } else if (MODE == 5) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(b), "r"(a));
        } else if (MODE == 6 || MODE == 7 || MODE == 8) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                if (MODE == 6) asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(acc[8 * c]) : "r"(a), "r"(b));
                if (MODE == 7) asm volatile("mad.hi.cc.u32 %0, %1, %2, %0;" : "+r"(acc[8 * c]) : "r"(a), "r"(b));
                if (MODE == 8) asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(acc[8 * c]) : "r"(a));
#pragma unroll
                for (int i = 1; i < 8; ++i) {
                    if (MODE == 6) asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(acc[8 * c + i]) : "r"(a), "r"(b));
                    if (MODE == 7) asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(acc[8 * c + i]) : "r"(a), "r"(b));
                    if (MODE == 8) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(acc[8 * c + i]) : "r"(a));
                }
                asm volatile("addc.u32 %0, %0, 0;" : "+r"(side));
            }
        } else if (MODE == 9) {
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[0]), "+r"(acc[1]) : "r"(a), "r"(b));
#pragma unroll
            for (int i = 2; i < 16; i += 2)
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[i]), "+r"(acc[i + 1]) : "r"(a), "r"(b));
            asm volatile("addc.u32 %0, %0, 0;" : "+r"(side));
        } else if (MODE == 10) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b));
                asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(acc[i + 1]) : "r"(acc[i]), "r"(b));
            }
        } else if (MODE == 11) {
#pragma unroll
            for (int i = 0; i < 16; i += 2)
                asm volatile("{ .reg .u64 t; mov.b64 t, {%0, %1}; mad.wide.u32 t, %0, %2, t; mov.b64 {%0, %1}, t; }"
                             : "+r"(acc[i]), "+r"(acc[i + 1]) : "r"(b));
        } else {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[8 * c]), "+r"(acc[8 * c + 1]) : "r"(a), "r"(b));
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[8 * c + 2]), "+r"(acc[8 * c + 3]) : "r"(a), "r"(b));
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[8 * c + 4]), "+r"(acc[8 * c + 5]) : "r"(a), "r"(b));
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[8 * c + 6]), "+r"(acc[8 * c + 7]) : "r"(a), "r"(b));
            }
        }
    }
    uint32_t x = side;
#pragma unroll
    for (int i = 0; i < 16; ++i) x ^= acc[i];
    if (x == 0x12345678u) sink[0] = x;  // never true in practice; keeps the work alive
}

// multiply-instructions issued per thread per iteration, per mode
__host__ inline int int32_probe_ops_per_iter(int mode) {
    switch (mode) {
        case 0: case 1: case 5: case 6: case 7: case 8: case 10: return 16;
        default: return 8;
    }
}
constexpr int INT32_PROBE_MODES = 12;

}  // namespace gsn
