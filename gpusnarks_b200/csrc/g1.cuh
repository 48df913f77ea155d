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
// Two algorithms: the reference's own (one double-and-add per point, then a tree reduction: g1_scalar_mul_kernel +
// g1_reduce_kernel, kept for very small inputs) and the bucket method (Pippenger) below, which is what a real
// multi-exponentiation uses: signed c-bit windows, one bucket per (window, |digit|), points sorted by bucket, one thread
// per bucket, a chunked running-sum reduction per window, and Horner over the window sums on the host (g1_host.h).
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
// exactly the Montgomery form of 1 (canonical): the Z of an affine input point
__device__ __forceinline__ bool fq_is_one(const uint32_t *a) {
    bool one = true;
    for (int i = 0; i < NL; ++i) one = one && a[i] == c_fq.r1[i];
    return one;
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
    if (fq_is_one(q.z)) {   // affine second operand (the usual input point): mixed addition, three products fewer
        fq_copy(X1Z2, p.x);
        fq_copy(Y1Z2, p.y);
        fq_copy(Z1Z2, p.z);
    } else {
        fq_mul(X1Z2, p.x, q.z);
        fq_mul(Y1Z2, p.y, q.z);
        fq_mul(Z1Z2, p.z, q.z);
    }
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

// ------------------------------------------------------------------ Fq2 = Fq[u] / (u^2 - 13)
// The reference's `fp2` (cuda/device_field.h:220-294): elements x + y u stored as (x, y), 2 x 24 limbs, Montgomery
// form; product by Karatsuba, aA + 13 bB and (a + b)(A + B) - aA - bB (device_field.h:253-262).  The reference multiplies
// by the raw integer 13 through its Montgomery routine (which scales by R^-1); here 13 bB is the field multiple, formed
// by additions (13 = 8 + 4 + 1).  op: 0 mul, 1 add, 2 sub; canonical outputs.
__global__ void fp2_binop768(uint32_t *out, const uint32_t *a, const uint32_t *b, uint64_t count, int op) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t ax[NL], ay[NL], bx[NL], by[NL], rx[NL], ry[NL];
    load_elem(ax, a + i * 2 * NL);
    load_elem(ay, a + i * 2 * NL + NL);
    load_elem(bx, b + i * 2 * NL);
    load_elem(by, b + i * 2 * NL + NL);
    if (op == 1) { fq_add(rx, ax, bx); fq_add(ry, ay, by); }
    else if (op == 2) { fq_sub(rx, ax, bx); fq_sub(ry, ay, by); }
    else {
        uint32_t aA[NL], bB[NL], s1[NL], s2[NL], t[NL], t4[NL];
        fq_mul(aA, ax, bx);
        fq_mul(bB, ay, by);
        fq_add(s1, ax, ay);
        fq_add(s2, bx, by);
        fq_mul(ry, s1, s2);
        fq_sub(ry, ry, aA);
        fq_sub(ry, ry, bB);
        fq_add(t, bB, bB);        // 2 bB
        fq_add(t4, t, t);         // 4 bB
        fq_add(t, t4, t4);        // 8 bB
        fq_add(t, t, t4);         // 12 bB
        fq_add(t, t, bB);         // 13 bB
        fq_add(rx, aA, t);
    }
    canonicalize(c_fq, rx);
    canonicalize(c_fq, ry);
    store_elem(out + i * 2 * NL, rx);
    store_elem(out + i * 2 * NL + NL, ry);
}

// ------------------------------------------------------------------ bucket method (Pippenger)
// Signed window digits of the raw 768-bit scalars.  Window w covers bits [w c, (w+1) c); with the carry of the previous
// window v = bits + carry, and v > 2^(c-1) is recoded as v - 2^c with a carry into the next window, so |digit| <= 2^(c-1).
// keys[w * n + i] = w * bs + |digit|, bs = 2^(c-1) + 1 buckets per window (|digit| = 0: nothing to add, bucket 0 is ignored),
// vals[w * n + i] = 2 i + (digit < 0).
__global__ void g1_digits_kernel(uint32_t *keys, uint32_t *vals, const uint32_t *scalars, uint64_t n, uint32_t c, uint32_t windows, uint32_t bs) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k[NL + 1];
    load_elem(k, scalars + i * NL);
    k[NL] = 0;
    uint32_t carry = 0;
    for (uint32_t w = 0; w < windows; ++w) {
        const uint32_t bit = w * c, word = bit >> 5, off = bit & 31;
        uint32_t v = 0;
        if (word <= NL - 1) {
            const uint64_t two = (uint64_t)k[word] | ((uint64_t)k[word + 1] << 32);
            v = (uint32_t)(two >> off) & ((1u << c) - 1);
        }
        v += carry;
        uint32_t neg = 0;
        carry = 0;
        if (v > (1u << (c - 1))) { v = (1u << c) - v; neg = 1; carry = 1; }
        keys[(uint64_t)w * n + i] = w * bs + v;       // v <= 2^(c-1)
        vals[(uint64_t)w * n + i] = (uint32_t)(2 * i + neg);
    }
}

// first index of the sorted key array that is >= key
__device__ __forceinline__ uint64_t g1_lower_bound(const uint32_t *keys, uint64_t count, uint32_t key) {
    uint64_t lo = 0, hi = count;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (keys[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ void g1_load_signed(G1 &P, const uint32_t *points, uint32_t v) {
    g1_load(P, points + (uint64_t)(v >> 1) * 3 * NL);
    if (v & 1) {  // -P = (X, -Y, Z)
        uint32_t zero[NL];
        for (int i = 0; i < NL; ++i) zero[i] = 0;
        fq_sub(P.y, zero, P.y);
    }
}

// Work list of the buckets one thread cannot take.  A bucket with more than `limit` points becomes one work item (one
// block sums it); with more than G1_SPLIT_POINTS points it is cut into `parts` items whose partial sums land in `partial`
// slots and are added by g1_heavy_combine_kernel.  (Scalars below 2^768 leave the top windows with a handful of digit
// values -- the carry of the signed recoding alone puts about half of all points into ONE bucket of the last window.)
struct G1HeavyItem { uint32_t key, part, parts, slot; };
struct G1HeavySplit { uint32_t key, slot0, parts; };
struct G1HeavyLists {
    uint32_t *counters;      // [0] items, [1] partial slots, [2] split buckets
    G1HeavyItem *items;
    G1HeavySplit *splits;
    uint32_t *partial;       // slots x 72 words
};
constexpr uint32_t G1_SPLIT_POINTS = 4096, G1_MAX_PARTS = 256;

// one thread per (window, bucket): buckets[key] = sum of the (signed) points whose sorted key equals `key`
__global__ void __launch_bounds__(128) g1_bucket_kernel(uint32_t *buckets, const uint32_t *points, const uint32_t *keys, const uint32_t *vals,
                                                        uint64_t pairs, uint32_t nbuckets, uint32_t bs, uint32_t limit, G1HeavyLists hl) {
    const uint32_t key = blockIdx.x * blockDim.x + threadIdx.x;
    if (key >= nbuckets) return;
    G1 acc, P, t;
    g1_set_identity(acc);
    if (key % bs != 0) {
        const uint64_t lo = g1_lower_bound(keys, pairs, key), hi = g1_lower_bound(keys, pairs, key + 1);
        const uint64_t size = hi - lo;
        if (size > limit) {
            uint32_t parts = 1;
            if (size > G1_SPLIT_POINTS) parts = (uint32_t)min((uint64_t)G1_MAX_PARTS, (size + G1_SPLIT_POINTS - 1) / G1_SPLIT_POINTS);
            const uint32_t j = atomicAdd(hl.counters, parts);
            uint32_t slot0 = 0;
            if (parts > 1) {
                slot0 = atomicAdd(hl.counters + 1, parts);
                hl.splits[atomicAdd(hl.counters + 2, 1u)] = G1HeavySplit{key, slot0, parts};
            }
            for (uint32_t k = 0; k < parts; ++k) hl.items[j + k] = G1HeavyItem{key, k, parts, slot0 + k};
        } else {
            for (uint64_t i = lo; i < hi; ++i) {
                g1_load_signed(P, points, vals[i]);
                g1_add(t, acc, P);
                g1_copy(acc, t);
            }
        }
    }
    g1_store(buckets + (uint64_t)key * 3 * NL, acc);
}

// shared-memory tree over the THREADS partial sums of a block; the total ends in thread 0's `acc`
template <int THREADS>
__device__ __forceinline__ void g1_block_sum(G1 &acc, uint32_t *g1_red) {
    G1 o, t;
    uint32_t *mine = g1_red + threadIdx.x * 3 * NL;
    for (int i = 0; i < NL; ++i) { mine[i] = acc.x[i]; mine[NL + i] = acc.y[i]; mine[2 * NL + i] = acc.z[i]; }
    __syncthreads();
    for (int half = THREADS / 2; half >= 1; half >>= 1) {
        if ((int)threadIdx.x < half) {
            const uint32_t *op = g1_red + (threadIdx.x + half) * 3 * NL;
            for (int i = 0; i < NL; ++i) { acc.x[i] = mine[i]; acc.y[i] = mine[NL + i]; acc.z[i] = mine[2 * NL + i]; }
            for (int i = 0; i < NL; ++i) { o.x[i] = op[i]; o.y[i] = op[NL + i]; o.z[i] = op[2 * NL + i]; }
            g1_add(t, acc, o);
            for (int i = 0; i < NL; ++i) { mine[i] = t.x[i]; mine[NL + i] = t.y[i]; mine[2 * NL + i] = t.z[i]; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0)
        for (int i = 0; i < NL; ++i) { acc.x[i] = mine[i]; acc.y[i] = mine[NL + i]; acc.z[i] = mine[2 * NL + i]; }
    __syncthreads();
}

// one block per work item: the threads stride over the item's share of the bucket, then a tree adds their partial sums
template <int THREADS>
__global__ void __launch_bounds__(THREADS) g1_heavy_bucket_kernel(uint32_t *buckets, const uint32_t *points, const uint32_t *keys, const uint32_t *vals,
                                                                  uint64_t pairs, G1HeavyLists hl) {
    extern __shared__ uint32_t g1_red[];  // THREADS * 72 words
    const uint32_t count = hl.counters[0];
    for (uint32_t h = blockIdx.x; h < count; h += gridDim.x) {
        const G1HeavyItem it = hl.items[h];
        const uint64_t lo = g1_lower_bound(keys, pairs, it.key), hi = g1_lower_bound(keys, pairs, it.key + 1), size = hi - lo;
        const uint64_t a = lo + size * it.part / it.parts, b = lo + size * (it.part + 1) / it.parts;
        G1 acc, P, t;
        g1_set_identity(acc);
        for (uint64_t i = a + threadIdx.x; i < b; i += THREADS) {
            g1_load_signed(P, points, vals[i]);
            g1_add(t, acc, P);
            g1_copy(acc, t);
        }
        g1_block_sum<THREADS>(acc, g1_red);
        if (threadIdx.x == 0) g1_store(it.parts == 1 ? buckets + (uint64_t)it.key * 3 * NL : hl.partial + (uint64_t)it.slot * 3 * NL, acc);
    }
}

// one block per split bucket: buckets[key] = sum of its `parts` partial sums
template <int THREADS>
__global__ void __launch_bounds__(THREADS) g1_heavy_combine_kernel(uint32_t *buckets, G1HeavyLists hl) {
    static_assert(THREADS >= (int)G1_MAX_PARTS, "one partial per thread");
    extern __shared__ uint32_t g1_red[];
    const uint32_t count = hl.counters[2];
    for (uint32_t h = blockIdx.x; h < count; h += gridDim.x) {
        const G1HeavySplit sp = hl.splits[h];
        G1 acc;
        if (threadIdx.x < sp.parts) g1_load(acc, hl.partial + (uint64_t)(sp.slot0 + threadIdx.x) * 3 * NL);
        else g1_set_identity(acc);
        g1_block_sum<THREADS>(acc, g1_red);
        if (threadIdx.x == 0) g1_store(buckets + (uint64_t)sp.key * 3 * NL, acc);
    }
}

// small multiple k * P by double-and-add (k < 2^16)
__device__ __noinline__ void g1_mul_small(G1 &r, const G1 &p, uint32_t k) {
    G1 acc, t;
    g1_set_identity(acc);
    for (int bit = 31 - __clz(k | 1); bit >= 0; --bit) {
        g1_dbl(t, acc);
        if ((k >> bit) & 1) g1_add(acc, t, p);
        else g1_copy(acc, t);
    }
    g1_copy(r, acc);
}

// gridDim.y blocks per window: S_w = sum_{b=1}^{nb} b * B[w][b], nb = 2^(c-1).  Block (w, j) takes the buckets
// (j nb/P, (j+1) nb/P], thread t of it the m = nb / (P T) buckets above b0 = j nb/P + t m: a descending running sum gives
// sum (b - b0) B_b and the chunk total T_t, the chunk contributes that plus b0 * T_t; the THREADS contributions are
// added by a shared-memory tree.  out[w * P + j] (canonical); the host adds the P parts of a window.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) g1_window_reduce_kernel(uint32_t *out, const uint32_t *buckets, uint32_t c, uint32_t bs) {
    extern __shared__ uint32_t g1_red[];  // THREADS * 72 words
    const uint32_t w = blockIdx.x, nb = (1u << (c - 1)) / gridDim.y, base = blockIdx.y * nb;
    const uint32_t T = nb < (uint32_t)THREADS ? nb : (uint32_t)THREADS, m = nb / T;
    const uint32_t *B = buckets + (uint64_t)w * bs * 3 * NL;
    G1 run, sum, t, nxt;
    g1_set_identity(run);
    g1_set_identity(sum);
    if (threadIdx.x < T) {
        const uint32_t b0 = base + threadIdx.x * m;
        for (uint32_t b = b0 + m; b > b0; --b) {
            g1_load(nxt, B + (uint64_t)b * 3 * NL);
            g1_add(t, run, nxt);
            g1_copy(run, t);
            g1_add(t, sum, run);
            g1_copy(sum, t);
        }
        if (b0) {
            g1_mul_small(nxt, run, b0);
            g1_add(t, sum, nxt);
            g1_copy(sum, t);
        }
    }
    g1_block_sum<THREADS>(sum, g1_red);
    if (threadIdx.x == 0) {
        canonicalize(c_fq, sum.x);
        canonicalize(c_fq, sum.y);
        canonicalize(c_fq, sum.z);
        g1_store(out + ((uint64_t)w * gridDim.y + blockIdx.y) * 3 * NL, sum);
    }
}

}  // namespace gsn
