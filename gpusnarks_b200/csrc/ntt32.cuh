// ntt32.cuh -- multi-pass radix-2 NTT over a 32-bit prime field (odd prime p < 2^31), sm_100a.
//
// The reference's 32-bit field is `dummy_fields::Field` (reference fields/dummy_field.h:24-62):
// one uint32_t `im_rep`, plain residues.  Its arithmetic is C wrap-around with mod = 0
// (fields/dummy_field.cpp:25,75-90), which cannot host an NTT; here every operation is a real
// reduction modulo `mod`.  Elements stay plain residues (no Montgomery form), so twiddle
// products use Shoup's precomputed-quotient trick: for a table entry (w, w' = floor(w*2^32/p))
//     q = umulhi(x, w');  r = x*w - q*p  in [0, 2p)   -- two low multiplies and one high.
// Same pass structure and index math as ntt768.cuh (PassGeom, tools/model_passes.py): tile in
// shared memory, bit-reversed placement fused into the load, stages in shared memory, optional
// four-step pre-twiddle on load, digit-reversed store in the last pass.
#pragma once
#include <cstdint>
#include "ntt768.cuh"

namespace gsn {
// geometry of one 32-bit pass (same index math as the 768-bit PassGeom, tools/model_passes.py)
struct PassGeom32 {
    uint32_t log_l;          // stages of this pass (digit width)
    uint32_t log_s;          // log2 stride (elements) of this digit
    uint32_t log_r;          // log2 inner stride
    uint32_t pre_shift;      // pre-twiddle index = (element index >> pre_shift) & pre_mask
    uint32_t log_tile;       // log2 elements per CTA tile (>= log_l)
    uint32_t wloc_shift;     // local table index = (jj << (log_l - s)) << wloc_shift
    uint32_t final_natural;  // last pass: write digits reversed
    uint32_t canonical;
    uint32_t ndig;           // number of digits of the whole transform
    uint32_t dig[4];         // digit widths l_1..l_P
    uint32_t logn;           // sum of digits
    uint32_t has_pre;        // pre-twiddle table present
    uint32_t tile0;
    uint64_t pre_mask;       // 0 => scalar pre-multiply (pre_tw[0])
};
__device__ __forceinline__ uint64_t elem_index(const PassGeom32 &g, uint64_t t, uint32_t j) {
    const uint64_t o = t >> g.log_s, rlow = t & ((1ull << g.log_s) - 1);
    return (((o << g.log_l) | j) << g.log_s) | rlow;
}
__device__ __forceinline__ uint64_t out_index(const PassGeom32 &g, uint64_t t, uint32_t k) {
    if (!g.final_natural) return elem_index(g, t, k);
    const uint64_t o = t >> g.log_s, rlow = t & ((1ull << g.log_s) - 1);
    const uint32_t inner_bits = g.logn - g.log_l;
    const uint64_t batch = o >> inner_bits;
    const uint64_t rest = o & ((1ull << inner_bits) - 1);
    uint64_t out = 0;
    uint32_t shift = 0, pos = inner_bits;
    for (uint32_t q = 0; q + 1 < g.ndig; ++q) {
        pos -= g.dig[q];
        out |= ((rest >> pos) & ((1ull << g.dig[q]) - 1)) << shift;
        shift += g.dig[q];
    }
    out |= (uint64_t)k << shift;
    return (((batch << g.logn) | out) << g.log_r) | rlow;
}
}  // namespace gsn

namespace gsn {

__device__ __forceinline__ uint32_t mulmod_shoup(uint32_t x, uint2 w, uint32_t p) {
    const uint32_t q = __umulhi(x, w.y);
    const uint32_t r = x * w.x - q * p;   // [0, 2p)
    return min(r, r - p);
}
__device__ __forceinline__ uint32_t addmod(uint32_t a, uint32_t b, uint32_t p) {
    const uint32_t s = a + b;             // < 2p < 2^32
    return min(s, s - p);
}
__device__ __forceinline__ uint32_t submod(uint32_t a, uint32_t b, uint32_t p) {
    const uint32_t d = a - b;
    return min(d, d + p);
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
ntt32_pass(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, const uint2 *__restrict__ wloc,
           const uint2 *__restrict__ pre_tw, const PassGeom32 g, const uint32_t p) {
    extern __shared__ uint32_t tile32[];
    const uint32_t T = 1u << g.log_tile;
    const uint32_t lq = g.log_l;
    const uint32_t Lm1 = (1u << lq) - 1;
    const uint32_t log_slots = g.log_tile - lq;
    const uint64_t sub0 = (uint64_t)blockIdx.x << log_slots;
    // lanes run along the memory-contiguous direction: j when this digit has stride 1,
    // otherwise the sub-transform slot (consecutive inner index)
    const bool slot_fast = g.log_s != 0;

    for (uint32_t e = threadIdx.x; e < T; e += THREADS) {
        uint32_t slot, j;
        if (slot_fast) { slot = e & ((1u << log_slots) - 1); j = e >> log_slots; }
        else { slot = e >> lq; j = e & Lm1; }
        const uint64_t gi = elem_index(g, sub0 + slot, j);
        uint32_t x = src[gi];
        if (g.has_pre) x = mulmod_shoup(x, __ldg(pre_tw + ((gi >> g.pre_shift) & g.pre_mask)), p);
        tile32[(slot << lq) | (lq ? (__brev(j) >> (32 - lq)) : 0u)] = x;
    }
    __syncthreads();

    for (uint32_t s = 1; s <= lq; ++s) {
        const uint32_t m = 1u << (s - 1);
        for (uint32_t b = threadIdx.x; b < (T >> 1); b += THREADS) {
            const uint32_t jj = b & (m - 1);
            const uint32_t lo = ((b >> (s - 1)) << s) | jj;
            const uint32_t hi = lo + m;
            uint32_t t = tile32[hi];
            if (s > 1) t = mulmod_shoup(t, __ldg(wloc + ((jj << (lq - s)) << g.wloc_shift)), p);
            const uint32_t u = tile32[lo];
            tile32[lo] = addmod(u, t, p);
            tile32[hi] = submod(u, t, p);
        }
        __syncthreads();
    }

    for (uint32_t e = threadIdx.x; e < T; e += THREADS) {
        uint32_t slot, k;
        if (slot_fast) { slot = e & ((1u << log_slots) - 1); k = e >> log_slots; }
        else { slot = e >> lq; k = e & Lm1; }
        dst[out_index(g, sub0 + slot, k)] = tile32[(slot << lq) | k];
    }
}

__device__ __forceinline__ uint32_t mulmod64(uint32_t a, uint32_t b, uint32_t p) { return (uint32_t)((uint64_t)a * b % p); }
__device__ __forceinline__ uint2 shoup_pair(uint32_t w, uint32_t p) { return make_uint2(w, (uint32_t)(((uint64_t)w << 32) / p)); }

// out[k] = (w, w') for w = base^(k * stride) * scale
__global__ void pow_table32(uint2 *out, uint32_t base, uint64_t count, uint64_t stride, uint32_t scale, uint32_t p) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    uint64_t e = k * stride;
    uint32_t acc = scale % p, b = base;
    while (e) {
        if (e & 1) acc = mulmod64(acc, b, p);
        b = mulmod64(b, b, p);
        e >>= 1;
    }
    out[k] = shoup_pair(acc, p);
}

// out[idx] = (w, w') for w = w_n^(((k * rest) mod N) << exp_shift), idx = (k << rest_bits) | rest
__global__ void build_pretw32(uint2 *out, const uint2 *t_lo, const uint2 *t_hi, uint32_t logN, uint32_t rest_bits,
                              uint32_t exp_shift, uint32_t lo_bits, uint32_t p) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >> logN) return;
    const uint64_t k = idx >> rest_bits, rest = idx & ((1ull << rest_bits) - 1);
    const uint64_t e = ((k * rest) & ((1ull << logN) - 1)) << exp_shift;
    const uint32_t w = mulmod64(t_lo[e & ((1ull << lo_bits) - 1)].x, t_hi[e >> lo_bits].x, p);
    out[idx] = shoup_pair(w, p);
}

}  // namespace gsn
