// fp768.cuh -- 768-bit (24 x 32-bit limb) Montgomery field arithmetic for sm_100a.
//
// Replaces the reference's device_field_operators.h (reference
// cuda/device_field_operators.h:92-258: less/_add/_subtract/montyNormalize/
// ciosMontgomeryMultiply/Scalar::add/subtract/mul) with PTX carry-chain code:
//   * add.cc / addc.cc / sub.cc / subc.cc chains for add, sub and the conditional
//     corrections (branch free, select by the borrow word);
//   * a CIOS Montgomery product whose 2*24*24 word products are issued as
//     mad.lo.cc / madc.hi.cc pairs on the same operands, which ptxas fuses into one
//     IMAD.WIDE.U32 with a predicate carry.  To keep every 64-bit product on an
//     aligned (even, odd) register pair the running sum is held in two interleaved
//     accumulators: `ev` collects the products that start on an even limb of the
//     current frame, `od` those that start on an odd limb.  Each outer iteration
//     ends with a one-limb right shift (the Montgomery division by 2^32), which
//     swaps the alignment of the two accumulators -- so the loop is unrolled by two
//     with the roles exchanged, and no data ever moves.
// Values are kept "lazy" in [0, 2p) between butterfly stages (p has 753 bits, the
// container 768, so there are 15 bits of headroom) and are made canonical, [0, p),
// only where they leave the transform.
//
// The modulus is a kernel parameter (FieldConstants768, `const __grid_constant__`), so one
// binary serves MNT4-753 Fr (default) and Fq (the reference's literal `_mod`) per launch;
// with fully unrolled loops every modulus word is a constant-bank operand of the IMAD and
// costs no register.
#pragma once
#include <cstdint>

namespace gsn {

constexpr int NL = 24;  // reference: #define SIZE (768 / 32), cuda/device_field.h:35

struct FieldConstants768 {
    uint32_t p[NL];    // modulus
    uint32_t p2[NL];   // 2 * modulus
    uint32_t p3[NL];   // 3 * modulus (lazy subtraction: y = u + 3p - t for t in [0, 3p))
    uint32_t p6[NL];   // 6 * modulus (the same for the unit-twiddle butterflies of stage 2, whose t is < 6p)
    uint32_t r1[NL];   // R mod p  (Montgomery one)
    uint32_t r2[NL];   // R^2 mod p
    uint32_t np0;      // -p^-1 mod 2^32
    uint32_t qmagic;   // floor(2^32 / ((p >> 736) + 1)): quotient estimate of reduce_small
    uint32_t pad[2];
    uint32_t nprime[NL];  // -p^-1 mod 2^768 (builds the quotient constants of the fixed-operand product)
};
// The constants travel with every launch as a `const __grid_constant__` kernel parameter (param space is a constant
// bank: with fully unrolled loops each modulus word is an immediate-offset constant operand of the IMAD, no register,
// no upload, and -- unlike one __constant__ symbol per device -- nothing that a second context could overwrite).
#define GSN_FC const FieldConstants768 &fc

// ------------------------------------------------------------------ carry-chain primitives
__device__ __forceinline__ uint32_t add_cc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t addc_cc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t addc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t sub_cc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t subc_cc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t subc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
// (hi:lo) = a*b            (no carry in, no carry out)
__device__ __forceinline__ void mul_wide(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b) {
    asm volatile("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
// (hi:lo) += a*b           starts a chain: carry out in CC
__device__ __forceinline__ void mad_wide_cc(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b) {
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
// (hi:lo) += a*b + CC      continues a chain
__device__ __forceinline__ void madc_wide_cc(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b) {
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
// (dhi:dlo) = (shi:slo) + a*b + CC   continues a chain, separate source pair (the shifted accumulate)
__device__ __forceinline__ void madc_wide_cc_from(uint32_t &dlo, uint32_t &dhi, uint32_t a, uint32_t b, uint32_t slo, uint32_t shi) {
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;" : "=r"(dlo), "=r"(dhi) : "r"(a), "r"(b), "r"(slo), "r"(shi));
}

// ------------------------------------------------------------------ add / sub
// r = a + b            (no reduction; caller guarantees no overflow of 768 bits)
__device__ __forceinline__ void add_raw(uint32_t *r, const uint32_t *a, const uint32_t *b) {
    r[0] = add_cc(a[0], b[0]);
#pragma unroll
    for (int i = 1; i < NL - 1; ++i) r[i] = addc_cc(a[i], b[i]);
    r[NL - 1] = addc(a[NL - 1], b[NL - 1]);
}
// r = a - b, returns the borrow as an all-ones / all-zeros mask
__device__ __forceinline__ uint32_t sub_raw(uint32_t *r, const uint32_t *a, const uint32_t *b) {
    r[0] = sub_cc(a[0], b[0]);
#pragma unroll
    for (int i = 1; i < NL; ++i) r[i] = subc_cc(a[i], b[i]);
    return subc(0u, 0u);  // 0 - 0 - borrow
}
// if a >= m: a -= m      (m = fc.p or fc.p2), branch free
__device__ __forceinline__ void cond_sub(uint32_t *a, const uint32_t *m) {
    uint32_t d[NL];
    uint32_t borrow = sub_raw(d, a, m);
#pragma unroll
    for (int i = 0; i < NL; ++i) a[i] = borrow ? a[i] : d[i];
}
// lazy butterfly outputs: inputs u, t in [0, 2p)  ->  x = u + t, y = u - t, both in [0, 2p)
__device__ __forceinline__ void add_lazy(GSN_FC, uint32_t *x, const uint32_t *u, const uint32_t *t) {
    add_raw(x, u, t);          // < 4p < 2^768
    cond_sub(x, fc.p2);
}
__device__ __forceinline__ void sub_lazy(GSN_FC, uint32_t *y, const uint32_t *u, const uint32_t *t) {
    uint32_t borrow = sub_raw(y, u, t);   // in (-2p, 2p)
    // y += borrow ? 2p : 0
    y[0] = add_cc(y[0], fc.p2[0] & borrow);
#pragma unroll
    for (int i = 1; i < NL - 1; ++i) y[i] = addc_cc(y[i], fc.p2[i] & borrow);
    y[NL - 1] = addc(y[NL - 1], fc.p2[NL - 1] & borrow);
}
// [0, 2p) -> [0, p)
__device__ __forceinline__ void canonicalize(GSN_FC, uint32_t *a) { cond_sub(a, fc.p); }

// ------------------------------------------------------------------ Montgomery product
// One outer CIOS step for multiplier word `bi` (reference: one pass of the `i` loop of
// ciosMontgomeryMultiply, device_field_operators.h:156-184).  `ev`/`od` are the accumulators
// aligned to even/odd limbs of the frame on entry to this step (see file header).
template <bool FIRST>
__device__ __forceinline__ void cios_step(GSN_FC, uint32_t *ev, uint32_t *od, const uint32_t *a, uint32_t bi) {
    if (FIRST) {
#pragma unroll
        for (int j = 0; j < NL; j += 2) mul_wide(od[j], od[j + 1], a[j + 1], bi);
#pragma unroll
        for (int j = 0; j < NL; j += 2) mul_wide(ev[j], ev[j + 1], a[j], bi);
    } else {
        // After the previous step's shift `od` (old even accumulator) still carries its
        // limb 1 below this frame's first odd slot: fold it into ev[0], then accumulate
        // the odd products while sliding `od` down by two limbs.
        ev[0] = add_cc(ev[0], od[1]);
#pragma unroll
        for (int j = 0; j < NL - 2; j += 2) madc_wide_cc_from(od[j], od[j + 1], a[j + 1], bi, od[j + 2], od[j + 3]);
        madc_wide_cc_from(od[NL - 2], od[NL - 1], a[NL - 1], bi, 0u, 0u);
        mad_wide_cc(ev[0], ev[1], a[0], bi);
#pragma unroll
        for (int j = 2; j < NL; j += 2) madc_wide_cc(ev[j], ev[j + 1], a[j], bi);
        od[NL - 1] = addc(od[NL - 1], 0u);
    }
    const uint32_t m = ev[0] * fc.np0;
    mad_wide_cc(od[0], od[1], fc.p[1], m);
#pragma unroll
    for (int j = 2; j < NL; j += 2) madc_wide_cc(od[j], od[j + 1], fc.p[j + 1], m);
    mad_wide_cc(ev[0], ev[1], fc.p[0], m);
#pragma unroll
    for (int j = 2; j < NL; j += 2) madc_wide_cc(ev[j], ev[j + 1], fc.p[j], m);
    od[NL - 1] = addc(od[NL - 1], 0u);
    // ev[0] is now 0 mod 2^32: the frame shifts right one limb, ev <-> od swap roles.
}

// r = a * b * 2^-768 mod p, result in [0, 2p) for a in [0, 2p), b < 2^768 (lazy: no final
// subtraction).  `B` is any callable  uint32_t B(int i)  giving word i of b, so the
// multiplier can stream from shared memory without occupying 24 registers.
template <typename BWord>
__device__ __forceinline__ void mont_mul_lazy_w(GSN_FC, uint32_t *r, const uint32_t *a, BWord b) {
    uint32_t ev[NL], od[NL];
    cios_step<true>(fc, ev, od, a, b(0));
    cios_step<false>(fc, od, ev, a, b(1));
#pragma unroll
    for (int i = 2; i < NL; i += 2) {
        cios_step<false>(fc, ev, od, a, b(i));
        cios_step<false>(fc, od, ev, a, b(i + 1));
    }
    // The last step ran with the roles exchanged (even accumulator = `od`) and left its
    // one-limb shift pending: value = (od >> 32) + ev.
    r[0] = add_cc(od[1], ev[0]);
#pragma unroll
    for (int k = 1; k < NL - 1; ++k) r[k] = addc_cc(od[k + 1], ev[k]);
    r[NL - 1] = addc(ev[NL - 1], 0u);
}

struct RegWords {
    const uint32_t *w;
    __device__ __forceinline__ uint32_t operator()(int i) const { return w[i]; }
};

__device__ __forceinline__ void mont_mul_lazy(GSN_FC, uint32_t *r, const uint32_t *a, const uint32_t *b) {
    mont_mul_lazy_w(fc, r, a, RegWords{b});
}
// canonical product, [0, p)
__device__ __forceinline__ void mont_mul(GSN_FC, uint32_t *r, const uint32_t *a, const uint32_t *b) {
    mont_mul_lazy_w(fc, r, a, RegWords{b});
    canonicalize(fc, r);
}

// ------------------------------------------------------------------ fixed-operand ("Shoup") product
// Every product of the transform multiplies data by a twiddle that is known when the plan is built.  For such a
// fixed operand w (< p, plain integer) the quotient of x*w by p can be estimated from a precomputed constant
// w'' = floor(w * 2^768 / p):  q = floor(x * w'' / 2^768) is within 2 of floor(x*w/p) even when only the partial
// products at limb positions >= 22 are summed, so  t = x*w - q*p  lies in [0, 3p) and needs only the LOW 768 bits of
// x*w and of q*p.  Three truncated half products replace the full 24x24 products of the CIOS loop:
//   876 wide + 48 low multiplies instead of 1152 + 24  (measured stand-alone: 9.85e9 vs 7.93e9 products/s).
// The data keeps its Montgomery form: x = X*R, w plain  =>  t = (X*w)*R.  The CIOS product above remains the general
// one (table building, element-wise API); this one is used wherever the second operand comes from a twiddle table.
//
// Row primitive: adds sum_{j in [jlo, jhi)} a[j] * b * 2^(32 (i + j)) into the interleaved accumulators (EV: products
// starting on an even limb position, OD: odd, shifted by one limb), positions relative to BASE; positions >= TOP fall
// outside the kept range (their carries are dropped) and at position LO_POS only the low word is kept.
// FRESH_TOP: the chain that ends with the product a[NL-1] * b (limb position i + NL - 1) propagates no carry: that
// register pair is touched by no earlier product of the row sequence i = 0, 1, 2, ... and holds at most a propagated
// carry bit in its low limb, so lo + product + carry-in < 2^64 (asserted limb for limb in tools/model_products.py).
// FIRST_ROW: the accumulators hold nothing yet: every product is WRITTEN (mul.lo / mul.hi, no carry chain, and no
// zero-initialisation of the accumulators before it).
template <int BASE, int TOP, int LO_POS, bool FRESH_TOP = false, bool FIRST_ROW = false>
__device__ __forceinline__ void row_mac(uint32_t *ev, uint32_t *od, const uint32_t *a, uint32_t b, int i, int jlo, int jhi) {
#pragma unroll
    for (int parity = 0; parity < 2; ++parity) {
        uint32_t *arr = parity ? od : ev;
        bool started = false;
        int last = -1, last_j = -1;
#pragma unroll
        for (int j = 0; j < NL; ++j) {
            if (j < jlo || j >= jhi) continue;
            const int pos = i + j;
            if ((pos & 1) != parity) continue;
            const int k = pos - BASE - parity;
            if (FIRST_ROW) {
                if (pos == LO_POS) arr[k] = a[j] * b;
                else mul_wide(arr[k], arr[k + 1], a[j], b);
                continue;
            }
            if (pos == LO_POS) {  // only the low word lands inside the kept range; it ends the chain
                if (started) asm volatile("madc.lo.u32 %0, %1, %2, %0;" : "+r"(arr[k]) : "r"(a[j]), "r"(b));
                else asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(arr[k]) : "r"(a[j]), "r"(b));
                started = false;
                last = -1;
            } else {
                if (!started) mad_wide_cc(arr[k], arr[k + 1], a[j], b);
                else madc_wide_cc(arr[k], arr[k + 1], a[j], b);
                started = true;
                last = k + 1;
                last_j = j;
            }
        }
        if (started && (BASE + last + 1 + parity) < TOP && !(FRESH_TOP && last_j == NL - 1)) arr[last + 1] = addc(arr[last + 1], 0u);
    }
}
// out[k] = EV[k] + OD[k-1] (+ carry), k < N
template <int N>
__device__ __forceinline__ void merge_evod(uint32_t *out, const uint32_t *ev, const uint32_t *od) {
    out[0] = ev[0];
    out[1] = add_cc(ev[1], od[0]);
#pragma unroll
    for (int k = 2; k < N - 1; ++k) out[k] = addc_cc(ev[k], od[k - 1]);
    out[N - 1] = addc(ev[N - 1], od[N - 2]);
}
// low 768 bits of a * b, b given word by word (callable)
template <typename BWord>
__device__ __forceinline__ void mul_lo768(uint32_t *r, const uint32_t *a, BWord b) {
    uint32_t ev[NL + 2], od[NL + 2];
    // row 0 covers limb positions 0..23, i.e. every accumulator limb merge_evod<NL> reads: it writes them, nothing is zeroed
    row_mac<0, NL, NL - 1, false, true>(ev, od, a, b(0), 0, 0, NL);
#pragma unroll
    for (int i = 1; i < NL; ++i) row_mac<0, NL, NL - 1>(ev, od, a, b(i), i, 0, NL - i);
    merge_evod<NL>(r, ev, od);
}
struct ConstModulus { const FieldConstants768 &fc; __device__ __forceinline__ uint32_t operator()(int i) const { return fc.p[i]; } };
struct ConstNprime { const FieldConstants768 &fc; __device__ __forceinline__ uint32_t operator()(int i) const { return fc.nprime[i]; } };

// 96-byte table half -> registers (read-only path; table entries are 32-byte aligned)
__device__ __forceinline__ void load_tw_half(uint32_t *r, const uint32_t *g) {
    const uint4 *pw = reinterpret_cast<const uint4 *>(g);
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        const uint4 v = __ldg(pw + c);
        r[4 * c] = v.x; r[4 * c + 1] = v.y; r[4 * c + 2] = v.z; r[4 * c + 3] = v.w;
    }
}

// t = x * w mod p, unreduced: t = x*w - q*p in [0, 3p) for ANY x < 2^768.  x is given word by word, twice (two passes
// over it); w2 = w'' (registers: the caller may have prefetched it), tw points at the table entry whose first half
// is w (plain, < p); it is fetched while the quotient product runs.
template <typename XWords>
__device__ __forceinline__ void shoup_mul_3p(GSN_FC, uint32_t *t, XWords x1, XWords x2, const uint32_t *w2, const uint32_t *tw) {
    uint32_t q[NL];
    {
        uint32_t ev[NL + 4], od[NL + 4], hi[NL + 2];
#pragma unroll
        for (int k = 0; k < NL + 4; ++k) ev[k] = od[k] = 0;
#pragma unroll
        for (int i = 0; i < NL; ++i) row_mac<22, 1000, -1, true>(ev, od, w2, x1(i), i, (22 - i) > 0 ? (22 - i) : 0, NL);
        merge_evod<NL + 2>(hi, ev, od);  // limbs 22..47 of x * w''
#pragma unroll
        for (int k = 0; k < NL; ++k) q[k] = hi[k + 2];
    }
    uint32_t p3[NL];
    mul_lo768(p3, q, ConstModulus{fc});
    uint32_t p2v[NL];
    {
        uint32_t w[NL];
        load_tw_half(w, tw);
        mul_lo768(p2v, w, x2);
    }
    t[0] = sub_cc(p2v[0], p3[0]);
#pragma unroll
    for (int k = 1; k < NL - 1; ++k) t[k] = subc_cc(p2v[k], p3[k]);
    t[NL - 1] = subc(p2v[NL - 1], p3[NL - 1]);
}

// t = x * w mod p in [0, 2p): the same with one conditional subtraction (strict callers)
template <typename XWords>
__device__ __forceinline__ void shoup_mul_lazy(GSN_FC, uint32_t *t, XWords x1, XWords x2, const uint32_t *tw) {
    uint32_t w2[NL];
    load_tw_half(w2, tw + NL);
    shoup_mul_3p(fc, t, x1, x2, w2, tw);
    cond_sub(t, fc.p2);
}

// ------------------------------------------------------------------ wide lazy ranges
// Inside a pass nothing but the products reduces: a butterfly with t in [0, K p) maps u -> (u + t, u + K p - t).
// K = 3 wherever t comes out of a product; a unit-twiddle butterfly of stage s (s <= 4) takes t as it is, below
// 3p 2^(s-1), and uses K = 3 2^(s-1).  Bounds: < 6p after stage 1, < 12p after stage 2, then + 3p per stage (< 36p after ten
// stages) or, with unit butterflies in stages 3 and 4 as well, < 24p, < 48p and + 3p per later stage (< 66p); p < 2^753,
// container 2^768.  The product accepts any x < 2^768, and values are brought back to [0, p) once, by reduce_small,
// where they leave the transform.  That removes three carry-chain passes and two selects from every butterfly.
// d = K p - t
__device__ __forceinline__ void neg_wide(uint32_t *d, const uint32_t *kp, const uint32_t *t) {
    d[0] = sub_cc(kp[0], t[0]);
#pragma unroll
    for (int i = 1; i < NL - 1; ++i) d[i] = subc_cc(kp[i], t[i]);
    d[NL - 1] = subc(kp[NL - 1], t[NL - 1]);
}
// v < 1024p  ->  [0, p).  Quotient estimate from the top limb: with d = (p >> 736) + 1, q = floor(v[23] * floor(2^32/d) / 2^32)
// satisfies floor(v/p) - 2 <= q <= floor(v/p) (tools/model_products.py checks the bound), so v - q*p is in [0, 3p).
__device__ __forceinline__ void reduce_small(GSN_FC, uint32_t *v) {
    const uint32_t q = __umulhi(v[NL - 1], fc.qmagic);
    uint32_t qp[NL];
    {
        uint32_t lo, hi, carry = 0;
#pragma unroll
        for (int i = 0; i < NL; ++i) {  // q < 1024, p[i] < 2^32: q*p[i] + carry fits 64 bits
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %4; madc.hi.u32 %1, %2, %3, 0;" : "=r"(lo), "=r"(hi) : "r"(fc.p[i]), "r"(q), "r"(carry));
            qp[i] = lo;
            carry = hi;
        }
    }
    v[0] = sub_cc(v[0], qp[0]);
#pragma unroll
    for (int i = 1; i < NL - 1; ++i) v[i] = subc_cc(v[i], qp[i]);
    v[NL - 1] = subc(v[NL - 1], qp[NL - 1]);
    cond_sub(v, fc.p2);
    cond_sub(v, fc.p);
}

// ------------------------------------------------------------------ global memory access
// elements are AoS, 96 bytes = 6 x 16 B, 32-byte aligned (reference layout: raw im_rep[24])
__device__ __forceinline__ void load_elem(uint32_t *r, const uint32_t *g) {
    const uint4 *p = reinterpret_cast<const uint4 *>(g);
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        uint4 v = p[c];
        r[4 * c] = v.x; r[4 * c + 1] = v.y; r[4 * c + 2] = v.z; r[4 * c + 3] = v.w;
    }
}
__device__ __forceinline__ void load_elem_nc(uint32_t *r, const uint32_t *g) {
    const uint4 *p = reinterpret_cast<const uint4 *>(g);
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        uint4 v = __ldg(p + c);
        r[4 * c] = v.x; r[4 * c + 1] = v.y; r[4 * c + 2] = v.z; r[4 * c + 3] = v.w;
    }
}
__device__ __forceinline__ void store_elem(uint32_t *g, const uint32_t *r) {
    uint4 *p = reinterpret_cast<uint4 *>(g);
#pragma unroll
    for (int c = 0; c < 6; ++c) p[c] = make_uint4(r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
}

}  // namespace gsn
