// g1.cuh -- MNT4-753 G1 multi-exponentiation sum_i s_i * P_i, the reference's second kernel family
// (reference cuda/multi_exp.cu:86-137 over `mnt4753_G1`, cuda/device_field.h:296-437; CPU form test/multiexp.h:3-13).
// SURVEY.md section 8f rank 2 ("next"): built on the same Montgomery arithmetic as the NTT, with the context's
// 768-bit field set to MNT4-753 Fq (the curve's base field = the reference's literal `_mod`).
//
// Same coordinates and formulas as the reference: homogeneous projective (X : Y : Z), a = 2,
//   add  = add-1998-cmo-2 (12M + 2S)   reference mnt4753_G1::operator+ (:327-348)
//   dbl  = dbl-2007-bl                 reference mnt4753_G1::dbl       (:350-372)
//   s*P  = MSB-first double-and-add    reference mnt4753_G1::operator* (:394-411)
// What differs: the reference's operator+ has no identity / doubling / inverse cases, so its `zero() + P` is
// (0,0,0) and every multiple it computes is zero; here the identity is Z == 0 and those cases are handled.
// The algorithm is the reference's (one scalar multiplication per point, then a tree reduction), not a bucket
// method: this row is about coverage and bit-exactness, not MSM performance.
//
// The Montgomery product is ~1.2k instructions, and the group law calls it 25 times: it is kept out of line
// (`fq_mul`, operands in local memory) so the kernels stay small enough for the instruction cache.
#pragma once
#include "../../include/gsn_constants.h"
#include "fp768.cuh"

namespace gsn {

// The curve lives over MNT4-753 Fq whatever field the context's transforms use: its constants are a compile-time
// __constant__ object (never written at run time), so G1 calls need no gsn_set_field768 and cannot race with it.
__constant__ FieldConstants768 c_fq = {GSN_FQ_MOD, GSN_FQ_MOD2, GSN_FQ_MOD3, GSN_FQ_MOD6, GSN_FQ_R1, GSN_FQ_R2,
                                       GSN_FQ_NP0, GSN_FQ_QMAGIC, {0, 0}, GSN_FQ_NPRIME768};

struct G1 {
    uint32_t x[NL], y[NL], z[NL];
};

__device__ __noinline__ void fq_mul(uint32_t *r, const uint32_t *a, const uint32_t *b) {
    uint32_t x[NL], y[NL], t[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) { x[i] = a[i]; y[i] = b[i]; }
    mont_mul_lazy(c_fq, t, x, y);
#pragma unroll
    for (int i = 0; i < NL; ++i) r[i] = t[i];
}
__device__ __noinline__ void fq_add(uint32_t *r, const uint32_t *a, const uint32_t *b) {
    uint32_t x[NL], y[NL], t[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) { x[i] = a[i]; y[i] = b[i]; }
    add_lazy(c_fq, t, x, y);
#pragma unroll
    for (int i = 0; i < NL; ++i) r[i] = t[i];
}
__device__ __noinline__ void fq_sub(uint32_t *r, const uint32_t *a, const uint32_t *b) {
    uint32_t x[NL], y[NL], t[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) { x[i] = a[i]; y[i] = b[i]; }
    sub_lazy(c_fq, t, x, y);
#pragma unroll
    for (int i = 0; i < NL; ++i) r[i] = t[i];
}
// lazy values live in [0, 2p): zero is 0 or p
__device__ __forceinline__ bool fq_is_zero(const uint32_t *a) {
    bool all0 = true, allp = true;
    for (int i = 0; i < NL; ++i) { all0 = all0 && a[i] == 0; allp = allp && a[i] == c_fq.p[i]; }
    return all0 || allp;
}
__device__ __forceinline__ void fq_copy(uint32_t *r, const uint32_t *a) {
    for (int i = 0; i < NL; ++i) r[i] = a[i];
}
__device__ __forceinline__ void g1_set_identity(G1 &r) {
    for (int i = 0; i < NL; ++i) { r.x[i] = 0; r.y[i] = c_fq.r1[i]; r.z[i] = 0; }
}
__device__ __forceinline__ void g1_copy(G1 &r, const G1 &a) { fq_copy(r.x, a.x); fq_copy(r.y, a.y); fq_copy(r.z, a.z); }

// reference mnt4753_G1::dbl (device_field.h:350-372); Z1 == 0 (or Y1 == 0) gives Z3 == 0, the identity
__device__ __noinline__ void g1_dbl(G1 &r, const G1 &p) {
    uint32_t XX[NL], ZZ[NL], w[NL], s[NL], ss[NL], sss[NL], R[NL], RR[NL], T[NL], B[NL], h[NL], t[NL];
    fq_mul(XX, p.x, p.x);
    fq_mul(ZZ, p.z, p.z);
    fq_add(w, ZZ, ZZ);            // a * ZZ, a = 2
    fq_add(t, XX, XX);
    fq_add(t, t, XX);
    fq_add(w, w, t);              // w = a*ZZ + 3*XX
    fq_mul(s, p.y, p.z);
    fq_add(s, s, s);              // s = 2*Y1*Z1
    fq_mul(ss, s, s);
    fq_mul(sss, s, ss);
    fq_mul(R, p.y, s);
    fq_mul(RR, R, R);
    fq_add(T, p.x, R);
    fq_mul(T, T, T);              // TT
    fq_sub(B, T, XX);
    fq_sub(B, B, RR);             // B = (X1+R)^2 - XX - RR
    fq_mul(h, w, w);
    fq_sub(h, h, B);
    fq_sub(h, h, B);              // h = w^2 - 2B
    fq_mul(r.x, h, s);            // X3 = h*s
    fq_sub(t, B, h);
    fq_mul(t, w, t);
    fq_sub(t, t, RR);
    fq_sub(r.y, t, RR);           // Y3 = w*(B-h) - 2*RR
    fq_copy(r.z, sss);            // Z3 = sss
}

// reference mnt4753_G1::operator+ (device_field.h:327-348) plus the cases it lacks
__device__ __noinline__ void g1_add(G1 &r, const G1 &p, const G1 &q) {
    if (fq_is_zero(p.z)) { g1_copy(r, q); return; }
    if (fq_is_zero(q.z)) { g1_copy(r, p); return; }
    uint32_t X1Z2[NL], Y1Z2[NL], Z1Z2[NL], u[NL], v[NL], uu[NL], vv[NL], vvv[NL], R[NL], A[NL], t[NL];
    fq_mul(X1Z2, p.x, q.z);
    fq_mul(Y1Z2, p.y, q.z);
    fq_mul(Z1Z2, p.z, q.z);
    fq_mul(u, q.y, p.z);
    fq_sub(u, u, Y1Z2);
    fq_mul(v, q.x, p.z);
    fq_sub(v, v, X1Z2);
    if (fq_is_zero(v)) {
        if (fq_is_zero(u)) { G1 c; g1_copy(c, p); g1_dbl(r, c); }
        else g1_set_identity(r);
        return;
    }
    fq_mul(uu, u, u);
    fq_mul(vv, v, v);
    fq_mul(vvv, vv, v);
    fq_mul(R, vv, X1Z2);
    fq_mul(A, uu, Z1Z2);
    fq_add(t, R, R);
    fq_add(t, t, vvv);
    fq_sub(A, A, t);              // A = uu*Z1Z2 - (vvv + 2R)
    fq_sub(t, R, A);
    fq_mul(t, u, t);
    fq_mul(Y1Z2, vvv, Y1Z2);
    uint32_t X3[NL];
    fq_mul(X3, v, A);
    fq_sub(r.y, t, Y1Z2);         // Y3 = u*(R-A) - vvv*Y1Z2
    fq_mul(r.z, vvv, Z1Z2);       // Z3 = vvv*Z1Z2
    fq_copy(r.x, X3);
}

__device__ __forceinline__ void g1_load(G1 &p, const uint32_t *g) { load_elem(p.x, g); load_elem(p.y, g + NL); load_elem(p.z, g + 2 * NL); }
__device__ __forceinline__ void g1_store(uint32_t *g, const G1 &p) { store_elem(g, p.x); store_elem(g + NL, p.y); store_elem(g + 2 * NL, p.z); }

// out[i] = scalars[i] * points[i]; scalars are raw 768-bit little-endian integers (reference: hasBitAt on im_rep)
__global__ void __launch_bounds__(128) g1_scalar_mul_kernel(uint32_t *out, const uint32_t *points, const uint32_t *scalars, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1 P, R, T;
    g1_load(P, points + i * 3 * NL);
    uint32_t k[NL];
    load_elem(k, scalars + i * NL);
    g1_set_identity(R);
    int top = -1;
    for (int w = NL - 1; w >= 0 && top < 0; --w)
        if (k[w]) top = 32 * w + 31 - __clz(k[w]);
    for (int bit = top; bit >= 0; --bit) {
        g1_dbl(T, R);
        if ((k[bit >> 5] >> (bit & 31)) & 1) g1_add(R, T, P);
        else g1_copy(R, T);
    }
    g1_store(out + i * 3 * NL, R);
}

// one block: out = sum of in[0..count) (grid-stride accumulation, shared-memory tree); canonical coordinates
template <int THREADS>
__global__ void __launch_bounds__(THREADS) g1_reduce_kernel(uint32_t *out, const uint32_t *in, uint64_t count) {
    extern __shared__ uint32_t g1_red[];  // THREADS * 72 words
    G1 acc, t, nxt;
    g1_set_identity(acc);
    for (uint64_t i = threadIdx.x; i < count; i += THREADS) {
        g1_load(nxt, in + i * 3 * NL);
        g1_add(t, acc, nxt);
        g1_copy(acc, t);
    }
    uint32_t *mine = g1_red + threadIdx.x * 3 * NL;
    for (int i = 0; i < NL; ++i) { mine[i] = acc.x[i]; mine[NL + i] = acc.y[i]; mine[2 * NL + i] = acc.z[i]; }
    __syncthreads();
    for (int half = THREADS / 2; half >= 1; half >>= 1) {
        if ((int)threadIdx.x < half) {
            const uint32_t *o = g1_red + (threadIdx.x + half) * 3 * NL;
            for (int i = 0; i < NL; ++i) { acc.x[i] = mine[i]; acc.y[i] = mine[NL + i]; acc.z[i] = mine[2 * NL + i]; }
            for (int i = 0; i < NL; ++i) { nxt.x[i] = o[i]; nxt.y[i] = o[NL + i]; nxt.z[i] = o[2 * NL + i]; }
            g1_add(t, acc, nxt);
            for (int i = 0; i < NL; ++i) { mine[i] = t.x[i]; mine[NL + i] = t.y[i]; mine[2 * NL + i] = t.z[i]; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < NL; ++i) { acc.x[i] = mine[i]; acc.y[i] = mine[NL + i]; acc.z[i] = mine[2 * NL + i]; }
        canonicalize(c_fq, acc.x);
        canonicalize(c_fq, acc.y);
        canonicalize(c_fq, acc.z);
        g1_store(out, acc);
    }
}

}  // namespace gsn
