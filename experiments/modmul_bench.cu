// experiments/modmul_bench.cu -- throughput of one Montgomery product per thread, two formulations:
//   A  24 x 32-bit limbs, CIOS with mad.lo.cc/madc.hi.cc carry chains (fp768.cuh, the shipped one)
//   B  27 x 29-bit limbs, carry-free column accumulation with plain mad.wide.u32 (64-bit accumulators)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I gpusnarks_b200/csrc -I include -o experiments/modmul_bench experiments/modmul_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include "gsn_constants.h"
#include "fp768.cuh"

using namespace gsn;

constexpr int RB = 29, RL = 27;
constexpr uint32_t RMASK = (1u << RB) - 1;
struct RRConst { uint32_t p[RL]; uint32_t np; };
__constant__ RRConst c_rr;

struct Acc64 { uint32_t lo, hi; };
__device__ __forceinline__ void madw(Acc64 &acc, uint32_t a, uint32_t b) {
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(acc.lo), "+r"(acc.hi) : "r"(a), "r"(b));
}

// t = a*b/2^(29*27) mod p, limbs normalized (< 2^29), value < 2p.  a normalized < p; b limbs < 2^29 + 8.
__device__ __forceinline__ void mont_rr(uint32_t *t, const uint32_t *a, const uint32_t *b) {
    Acc64 acc[RL];
#pragma unroll
    for (int j = 0; j < RL; ++j) acc[j].lo = acc[j].hi = 0;
#pragma unroll
    for (int i = 0; i < RL; ++i) {
        const uint32_t bi = b[i];
#pragma unroll
        for (int j = 0; j < RL; ++j) madw(acc[(i + j) % RL], a[j], bi);
        const uint32_t m = (acc[i % RL].lo * c_rr.np) & RMASK;
#pragma unroll
        for (int j = 0; j < RL; ++j) madw(acc[(i + j) % RL], m, c_rr.p[j]);
        {   // acc[i+1] += acc[i] >> 29  (64-bit)
            const uint32_t clo = (acc[i % RL].lo >> RB) | (acc[i % RL].hi << (32 - RB)), chi = acc[i % RL].hi >> RB;
            Acc64 &n = acc[(i + 1) % RL];
            n.lo = add_cc(n.lo, clo);
            n.hi = addc(n.hi, chi);
        }
        acc[i % RL].lo = acc[i % RL].hi = 0;
    }
    uint32_t clo = 0, chi = 0;
#pragma unroll
    for (int j = 0; j < RL; ++j) {
        const uint32_t slo = add_cc(acc[j].lo, clo);
        const uint32_t shi = addc(acc[j].hi, chi);
        t[j] = slo & RMASK;
        clo = (slo >> RB) | (shi << (32 - RB));
        chi = shi >> RB;
    }
}

template <int WHICH>
__global__ void __launch_bounds__(256) chain(uint32_t *out, const uint32_t *in_a, const uint32_t *in_b, int iters) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    constexpr int W = WHICH == 0 ? 24 : RL;
    uint32_t a[W], x[W], t[W];
    for (int i = 0; i < W; ++i) { a[i] = in_a[i]; x[i] = in_b[(tid & 63) * W + i]; }
    for (int it = 0; it < iters; ++it) {
        if (WHICH == 0) mont_mul_lazy(t, a, x);
        else mont_rr(t, a, x);
#pragma unroll
        for (int i = 0; i < W; ++i) x[i] = t[i];
    }
    for (int i = 0; i < W; ++i) out[(size_t)tid * W + i] = x[i];
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

int main(int argc, char **argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 256;
    // constants
    static const uint32_t p32[24] = GSN_FR_MOD, p2[24] = GSN_FR_MOD2, r1[24] = GSN_FR_R1, r2[24] = GSN_FR_R2;
    FieldConstants768 fc;
    memcpy(fc.p, p32, 96); memcpy(fc.p2, p2, 96); memcpy(fc.r1, r1, 96); memcpy(fc.r2, r2, 96); fc.np0 = GSN_FR_NP0;
    CK(cudaMemcpyToSymbol(c_fp, &fc, sizeof(fc)));
    RRConst rc;
    auto to29 = [](uint32_t *d, const uint32_t *s) {
        for (int k = 0; k < RL; ++k) {
            int bit = RB * k, w = bit >> 5, o = bit & 31;
            unsigned long long v = s[w];
            if (w + 1 < 24) v |= (unsigned long long)s[w + 1] << 32;
            d[k] = (uint32_t)(v >> o) & RMASK;
        }
    };
    to29(rc.p, p32);
    uint32_t inv = 1;
    for (int i = 0; i < 5; ++i) inv *= 2 - rc.p[0] * inv;
    rc.np = (0u - inv) & RMASK;
    CK(cudaMemcpyToSymbol(c_rr, &rc, sizeof(rc)));
    // inputs: a = 3 (as raw limbs), b = 64 distinct small-ish values
    for (int which = 0; which < 2; ++which) {
        const int W = which == 0 ? 24 : RL;
        uint32_t ha[32] = {0}, hb[64 * 32];
        memset(hb, 0, sizeof(hb));
        uint32_t a32[24] = {0}, b32[24];
        for (int i = 0; i < 23; ++i) a32[i] = 0x9E3779B9u * (i + 1);
        a32[23] = 0x1234;
        if (which == 0) memcpy(ha, a32, 96); else to29(ha, a32);
        for (int k = 0; k < 64; ++k) {
            for (int i = 0; i < 23; ++i) b32[i] = 0x85EBCA6Bu * (i + 7 * k + 3);
            b32[23] = 0x0FFF;
            if (which == 0) memcpy(hb + k * W, b32, 96); else to29(hb + k * W, b32);
        }
        const int blocks = 148 * 8, threads = 256;
        uint32_t *da, *db, *dout;
        CK(cudaMalloc(&da, 128)); CK(cudaMalloc(&db, sizeof(hb))); CK(cudaMalloc(&dout, (size_t)blocks * threads * W * 4));
        CK(cudaMemcpy(da, ha, 128, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, hb, sizeof(hb), cudaMemcpyHostToDevice));
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaEventRecord(e0));
            if (which == 0) chain<0><<<blocks, threads>>>(dout, da, db, iters);
            else chain<1><<<blocks, threads>>>(dout, da, db, iters);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep) best = best < ms ? best : ms;
        }
        CK(cudaGetLastError());
        const double rate = (double)blocks * threads * iters / (best * 1e-3);
        uint32_t res[32];
        CK(cudaMemcpy(res, dout, W * 4, cudaMemcpyDeviceToHost));
        printf("variant %c: %d iters, %.3f ms, %.4g modmul/s, %.4g wideMAC-equiv/s (x1176); result limb0..2 of thread 0: %08x %08x %08x\n",
               which ? 'B' : 'A', iters, best, rate, rate * 1176, res[0], res[1], res[2]);
        // print full result of thread 0 for checking (iters = 1 gives a single product)
        printf("R%c:", which ? 'B' : 'A');
        for (int i = 0; i < W; ++i) printf(" %u", res[i]);
        printf("\n");
        cudaFree(da); cudaFree(db); cudaFree(dout);
    }
    return 0;
}
