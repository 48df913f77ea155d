// ntt32_fast.cuh -- register-blocked pass kernel for the 32-bit field (digits of 9..12 stages).
//
// A pass transforms sub-transforms of length L = 2^(A+B+C).  One CTA owns 8 of them (8
// consecutive sub-transform indices = 8 memory-adjacent columns for a strided digit, 8
// adjacent rows for the contiguous last digit, so that every global access is a full 32-byte
// sector), and runs an in-tile four-step with three register-resident rounds:
//   round A  2^A-point NTT over the top index bits, x twiddle w_L^(k1 * rest)      (table tA)
//   round B  2^B-point NTT over the middle bits,    x twiddle w_{2^(B+C)}^(k2 * j3) (table tB)
//   round C  2^C-point NTT over the low bits
// Round A reads straight from global memory (applying the inter-pass four-step twiddle when
// PRE: one two-level table lookup per thread, then a geometric recurrence along the 2^A
// elements it holds) and round C writes straight to global memory, so
// the tile crosses shared memory only twice.  Same index conventions as ntt768.cuh
// (PassGeom / tools/model_passes.py): output digit order k = k1 + 2^A k2 + 2^(A+B) k3.
// Arithmetic: canonical residues in [0, p), p < 2^31; twiddle products by Shoup's method.
// (A persistent one-CTA-per-SM variant that prefetched the next tile with cp.async during rounds
// B and C was measured ~10 % slower than two independent CTAs per SM -- 32.1 vs 29.3 us at 2^22 --
// and was removed: the stalls are pipe contention inside the compute phases, not load latency.)
#pragma once
#include "ntt32.cuh"

namespace gsn {

struct Ntt32Consts {
    uint2 rt[8];   // (w, w') for w_16^e, e = 0..7
    uint32_t p;
    uint32_t pinv;           // p^-1 mod 2^32 (Montgomery reduction of the running inter-pass twiddle)
    // inter-pass twiddle w_N^(k * rest): k = (t >> log_s) mod 2^pre_k_bits (the digit produced by the previous
    // pass), rest = (j << log_s) | (t mod 2^log_s); exponents mod 2^pre_logN, scaled by 2^pre_exp_shift to w_n
    uint32_t pre_k_bits, pre_logN, pre_exp_shift;
    uint32_t pre_lo_bits;    // two-level split of the w_n exponent
    uint32_t slot_shift;     // the 8 sub-transforms of a tile are 2^slot_shift apart (adjacent OUTPUTS for the last pass)
};

// first sub-transform of a tile and the step between its 8 slots
__device__ __forceinline__ uint64_t tile_sub0(uint32_t tile, uint32_t slot_shift) {
    return ((uint64_t)(tile >> slot_shift) << (slot_shift + 3)) | (tile & ((1u << slot_shift) - 1));
}

// a * bM * 2^-32 mod p, canonical, for a < 2^32 and bM < p (bM in Montgomery form => plain product a*b)
__device__ __forceinline__ uint32_t montmul32(uint32_t a, uint32_t bM, uint32_t p, uint32_t pinv) {
    const uint32_t lo = a * bM, hi = __umulhi(a, bM);
    const uint32_t m = lo * pinv;
    const uint32_t r = hi - __umulhi(m, p);   // in (-p, p)
    return min(r, r + p);
}

// natural-order in, natural-order out 2^LOGR-point NTT on registers (DIF + compile-time unscramble)
template <int LOGR>
__device__ __forceinline__ void ntt_reg(uint32_t *r, const Ntt32Consts &c) {
    constexpr int R = 1 << LOGR;
#pragma unroll
    for (int half = R / 2; half >= 1; half >>= 1) {
#pragma unroll
        for (int base = 0; base < R; base += 2 * half) {
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const uint32_t u = r[base + j], v = r[base + j + half];
                r[base + j] = addmod(u, v, c.p);
                const int e = j * (8 / half);  // w_16^(j * 16 / (2 half))
                // the Shoup product accepts any 32-bit operand, so u - v + p (in (0, 2p)) needs no reduction first
                r[base + j + half] = e == 0 ? submod(u, v, c.p) : mulmod_shoup(u - v + c.p, c.rt[e], c.p);
            }
        }
    }
    // outputs are bit-reversed: X[k] sits in r[brev(k)]
    uint32_t t[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
        int rk = 0;
#pragma unroll
        for (int b = 0; b < LOGR; ++b) rk |= ((k >> b) & 1) << (LOGR - 1 - b);
        t[k] = r[rk];
    }
#pragma unroll
    for (int k = 0; k < R; ++k) r[k] = t[k];
}

template <int A, int B, int C>
struct FastTile {
    static constexpr int LOGL = A + B + C;
    static constexpr int L = 1 << LOGL;
    static constexpr int SLOTS = 8;
    static constexpr int QA = 1 << (B + C);                 // elements per top-index value
    static constexpr int PITCH_A = QA + (1 << C);           // k1 * PITCH_A == k1 * 2^C (mod 32): round B conflict free
    static constexpr int PITCH_S = (1 << A) * PITCH_A + 4;  // slot * 4 (mod 32), 16-byte chunks: slot (mod 8)
    static constexpr size_t SMEM_BYTES = (size_t)SLOTS * PITCH_S * 4;
};

template <int A, int B, int C, bool SLOT_FAST, bool PRE, int MIN_BLOCKS>
__global__ void __launch_bounds__(512, MIN_BLOCKS)
ntt32_fast_pass(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, const uint2 *__restrict__ tA,
                const uint2 *__restrict__ tB, const uint2 *__restrict__ t_lo, const uint2 *__restrict__ t_hi,
                const uint2 *__restrict__ tG, const PassGeom32 g, const Ntt32Consts c) {
    using T = FastTile<A, B, C>;
    extern __shared__ uint32_t sm[];
    const uint64_t sub0 = tile_sub0(blockIdx.x, c.slot_shift);
    const uint32_t ss = c.slot_shift;
    const uint32_t p = c.p;

    // ---------------- round A: global -> registers -> shared
#pragma unroll 1
    for (uint32_t item = threadIdx.x; item < T::SLOTS * T::QA; item += 512) {
        uint32_t s, q;
        if (SLOT_FAST) { s = item & 7; q = item >> 3; } else { q = item & (T::QA - 1); s = item >> (B + C); }
        // element (sub-transform t, index j) lives at base + (j << log_s): linear in j
        const uint64_t t = sub0 + ((uint64_t)s << ss);
        const uint32_t *in = src + elem_index(g, t, q);
        const uint64_t jstride = (uint64_t)T::QA << g.log_s;
        uint32_t r[1 << A];
#pragma unroll
        for (int j1 = 0; j1 < (1 << A); ++j1) r[j1] = in[j1 * jstride];
        if (PRE) {
            // inter-pass twiddle w_N^(k * rest(j1)), rest(j1) = ((j1*QA + q) << log_s) | rlow: the first factor
            // w^(k * rest(0)) is one two-level table lookup, the rest a geometric recurrence with the per-k
            // step tG[k] = w_N^(k * (QA << log_s)).  The running twiddle is kept in Montgomery form (t_lo holds
            // w^e * [n^-1] * 2^32), so stepping it is a Shoup product and applying it a Montgomery product.
            const uint32_t k = (uint32_t)(t >> g.log_s) & ((1u << c.pre_k_bits) - 1);
            const uint32_t rest0 = (q << g.log_s) | ((uint32_t)t & ((1u << g.log_s) - 1));
            const uint32_t e0 = ((k * rest0) & ((1u << c.pre_logN) - 1)) << c.pre_exp_shift;
            uint32_t tw = mulmod_shoup(__ldg(t_lo + (e0 & ((1u << c.pre_lo_bits) - 1))).x, __ldg(t_hi + (e0 >> c.pre_lo_bits)), p);
            const uint2 step = __ldg(tG + k);
#pragma unroll
            for (int j1 = 0; j1 < (1 << A); ++j1) {
                r[j1] = montmul32(r[j1], tw, p, c.pinv);
                if (j1 + 1 < (1 << A)) tw = mulmod_shoup(tw, step, p);
            }
        }
        ntt_reg<A>(r, c);
        uint32_t *out = sm + s * T::PITCH_S + q;
        out[0] = r[0];
#pragma unroll
        for (int k1 = 1; k1 < (1 << A); ++k1) out[k1 * T::PITCH_A] = mulmod_shoup(r[k1], __ldg(tA + k1 * T::QA + q), p);
    }
    __syncthreads();

    // ---------------- round B: shared -> registers -> shared (in place)
#pragma unroll 1
    for (uint32_t item = threadIdx.x; item < T::SLOTS * (1u << (A + C)); item += 512) {
        const uint32_t j3 = item & ((1u << C) - 1), k1 = (item >> C) & ((1u << A) - 1), s = item >> (A + C);
        uint32_t *base = sm + s * T::PITCH_S + k1 * T::PITCH_A + j3;
        uint32_t r[1 << B];
#pragma unroll
        for (int j2 = 0; j2 < (1 << B); ++j2) r[j2] = base[j2 << C];
        ntt_reg<B>(r, c);
        base[0] = r[0];
#pragma unroll
        for (int k2 = 1; k2 < (1 << B); ++k2) base[k2 << C] = mulmod_shoup(r[k2], __ldg(tB + (k2 << C) + j3), p);
    }
    __syncthreads();

    // ---------------- round C: shared -> registers -> global
#pragma unroll 1
    for (uint32_t item = threadIdx.x; item < T::SLOTS * (1u << (A + B)); item += 512) {
        const uint32_t s = item & 7, kk = item >> 3;  // slot-fast: 8 adjacent outputs per 32-byte sector
        const uint32_t k1 = kk & ((1u << A) - 1), k2 = kk >> A;
        const uint4 *base = reinterpret_cast<const uint4 *>(sm + s * T::PITCH_S + k1 * T::PITCH_A + (k2 << C));
        uint32_t r[1 << C];
#pragma unroll
        for (int v = 0; v < (1 << C) / 4; ++v) {
            const uint4 x = base[v];
            r[4 * v] = x.x; r[4 * v + 1] = x.y; r[4 * v + 2] = x.z; r[4 * v + 3] = x.w;
        }
        ntt_reg<C>(r, c);
        // the output index is linear in k as well: out(k) = out(kk) + k3 * kstride
        const uint64_t t = sub0 + ((uint64_t)s << ss);
        const uint64_t o0 = out_index(g, t, kk);
        const uint64_t kstride = out_index(g, t, kk + (1u << (A + B))) - o0;
#pragma unroll
        for (int k3 = 0; k3 < (1 << C); ++k3) dst[o0 + k3 * kstride] = r[k3];
    }
}

}  // namespace gsn
