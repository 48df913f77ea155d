// ntt768.cuh -- multi-pass radix-2 NTT over the 768-bit field, sm_100a.
//
// Replaces reference cuda/fft_kernel.cu:52-115 (`cuda_fft`, an O(n * 1024) routine hard
// wired to n = 2^16 that ignores omega) with an O(n log n) transform for any power of two.
// The index math is the one modelled and checked in tools/model_passes.py (`run_pass`).
//
// Decomposition.  log2(n) is split into P "digits" l_1..l_P (each <= 10); the input index is
// read as mixed radix (d_1 .. d_P | r) with d_1 most significant and an optional inner
// stride 2^log_r, pass q transforms digit d_q into k_q, and the last pass writes the digits
// in reversed significance (k_P .. k_1 | r): natural order in, natural order out.  Between
// passes every element is multiplied by w_{N_q}^(k_{q-1} * (d_q..d_P)) from a device table
// (four-step / Bailey twiddles, built once per (n, omega) by build_pretw768).
//
// One CTA owns one tile of 2^log_tile (<= 1024) elements = 2^(log_tile - l_q) sub-transforms:
//   load   global -> shared, bit-reversed placement (the reference's swap loop,
//          test/fft_host.h:24-29, fused into the load)
//   mul    optional pre-twiddle, one Montgomery product per element
//   stages l_q radix-2 DIT butterfly stages in shared memory (test/fft_host.h:31-53 with the
//          twiddle read from a table instead of the running product w *= w_m)
//   store  shared -> global (in-place position, or digit-reversed for the last pass)
// All twiddle products of the kernel go through ONE inlined call site (the `ph` loop):
// the product is ~1k instructions, and a single copy keeps the kernel inside the
// instruction cache.  Butterflies whose twiddle is 1 skip the product: all of stage 1, and in
// stages 2-4 the jj == 0 butterflies, which a twiddle-major enumeration packs into whole warps
// (about 10 % of the products of a 10-stage pass).
//
// Shared memory layout: AoS with a 112-byte pitch (96 B of limbs + 16 B pad).  With a
// 28-word pitch the eight lanes of a quarter-warp that read the same 16-byte chunk of eight
// consecutive elements hit eight distinct bank groups, so LDS.128/STS.128 are conflict
// free for every stage with half-distance >= 8.
#pragma once
#include "fp768.cuh"

namespace gsn {

constexpr int SMEM_PITCH4 = 7;  // uint4 per element in shared memory (112 B)
constexpr int TW_WORDS = 2 * NL;  // twiddle table entry: w[24] (plain) followed by w''[24] = floor(w * 2^768 / p)

struct PassGeom {
    uint32_t log_l;          // stages of this pass (digit width)
    uint32_t log_s;          // log2 stride (elements) of this digit
    uint32_t log_r;          // log2 inner stride (elements below the transform index)
    uint32_t pre_shift;      // pre-twiddle index = (element index >> pre_shift) & pre_mask
    uint32_t log_tile;       // log2 elements per CTA tile (>= log_l)
    uint32_t wloc_shift;     // local table index = (jj << (log_l - s)) << wloc_shift
    uint32_t final_natural;  // last pass: write digits reversed
    uint32_t canonical;      // outputs reduced to [0, p) (else lazy [0, 2p))
    uint32_t ndig;           // number of digits of the whole transform
    uint32_t dig[4];         // digit widths l_1..l_P
    uint32_t logn;           // sum of digits
    uint32_t has_pre;        // pre-twiddle table present
    uint32_t tile0;          // first tile of this launch (chunked launches of one pass)
    uint64_t pre_mask;       // 0 => scalar pre-multiply (pre_tw[0])
};

// Fused exchange (multi-GPU four-step): the last pass can store each element straight into a
// peer GPU's buffer over NVLink instead of its own.  The local natural output index is split into
// three bit fields; the field at [rank_shift, rank_shift + rank_bits) names the destination rank and
// is replaced, in the destination index, by this rank's id inserted at bit ins_shift of the rest.
struct ScatterDesc {
    uint32_t *peers[8];   // peers[r] = base of rank r's receive buffer (own entry = local pointer)
    uint32_t enabled, rank_shift, rank_bits, ins_shift, my_rank, pad[3];
};

__device__ __forceinline__ uint32_t *scatter_target(const ScatterDesc &sc, uint64_t go) {
    const uint32_t dest = (uint32_t)(go >> sc.rank_shift) & ((1u << sc.rank_bits) - 1);
    const uint64_t rem = ((go >> (sc.rank_shift + sc.rank_bits)) << sc.rank_shift) | (go & ((1ull << sc.rank_shift) - 1));
    const uint64_t idx = ((rem >> sc.ins_shift) << (sc.ins_shift + sc.rank_bits)) | ((uint64_t)sc.my_rank << sc.ins_shift) |
                         (rem & ((1ull << sc.ins_shift) - 1));
    return sc.peers[dest] + idx * NL;
}

// Barrier across the GPUs of a four-step transform, in peer memory: lane r publishes `epoch` into slot
// [my_rank] of rank r's flag array (release at system scope: everything this GPU stored before the kernel
// boundary -- the scattered tiles -- is visible to whoever acquires the flag), then waits until rank r has
// published the same epoch into slot [r] of the local array.  A peer that never arrives traps after
// ~10 s instead of hanging the device.
struct PeerFlags {
    uint32_t *flags[8];   // flags[r] = rank r's array of 8 slots (own entry = local array)
};

__global__ void peer_barrier_kernel(const PeerFlags pf, uint32_t n_peers, uint32_t my_rank, uint32_t epoch) {
    const uint32_t r = threadIdx.x;
    if (r >= n_peers) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pf.flags[r] + my_rank), "r"(epoch) : "memory");
    const uint32_t *mine = pf.flags[my_rank] + r;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if ((int32_t)(v - epoch) >= 0) break;
        __nanosleep(200);
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 10000000000ull) __trap();
    }
}

__device__ __forceinline__ uint64_t elem_index(const PassGeom &g, uint64_t t, uint32_t j) {
    const uint64_t o = t >> g.log_s, rlow = t & ((1ull << g.log_s) - 1);
    return (((o << g.log_l) | j) << g.log_s) | rlow;
}

__device__ __forceinline__ uint64_t out_index(const PassGeom &g, uint64_t t, uint32_t k) {
    if (!g.final_natural) return elem_index(g, t, k);
    // t = (o | r) with o = (batch, k_1 .. k_{P-1}), log_s == log_r
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

// multiplier words streamed from shared memory, one LDS.128 per four limbs
struct SmemWords {
    const uint4 *p;
    uint4 cur;
    __device__ __forceinline__ uint32_t operator()(int i) {
        if ((i & 3) == 0) cur = p[i >> 2];
        return (i & 3) == 0 ? cur.x : (i & 3) == 1 ? cur.y : (i & 3) == 2 ? cur.z : cur.w;
    }
};

// Shared-memory slot of tile element e.  XOR-ing the low three bits with the next three keeps
// eight consecutive elements on eight distinct 16-byte bank groups (the common case) and also
// makes the stride-4 and stride-8 element patterns of the remapped early stages conflict free.
__device__ __forceinline__ uint32_t slot_of(uint32_t e) { return e ^ ((e >> 3) & 7u); }

__device__ __forceinline__ void lds_elem(uint32_t *r, const uint4 *s) {
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        uint4 v = s[c];
        r[4 * c] = v.x; r[4 * c + 1] = v.y; r[4 * c + 2] = v.z; r[4 * c + 3] = v.w;
    }
}
__device__ __forceinline__ void sts_elem(uint4 *s, const uint32_t *r) {
#pragma unroll
    for (int c = 0; c < 6; ++c) s[c] = make_uint4(r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
}

template <int THREADS, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
ntt768_pass(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, const uint32_t *__restrict__ wloc,
            const uint32_t *__restrict__ pre_tw, const PassGeom g, const ScatterDesc sc) {
    extern __shared__ uint4 tile[];
    const uint32_t T = 1u << g.log_tile;
    const uint32_t lq = g.log_l;
    const uint32_t Lm1 = (1u << lq) - 1;
    const uint64_t sub0 = (uint64_t)(blockIdx.x + g.tile0) << (g.log_tile - lq);  // first sub-transform of this tile

    // ---- load: lane per 16-byte chunk (coalesced 96-byte elements), bit-reversed placement inside the sub-transform
    for (uint32_t idx = threadIdx.x; idx < T * 6; idx += THREADS) {
        const uint32_t e = idx / 6, c = idx - e * 6;
        const uint32_t slot = e >> lq, j = e & Lm1;
        const uint64_t gi = elem_index(g, sub0 + slot, j);
        const uint32_t pos = (slot << lq) | (lq ? (__brev(j) >> (32 - lq)) : 0u);
        tile[slot_of(pos) * SMEM_PITCH4 + c] = reinterpret_cast<const uint4 *>(src + gi * NL)[c];
    }
    __syncthreads();

    // ---- ph = 0: pre-twiddle (one product per element); ph = s >= 1: butterfly stage s
    for (uint32_t ph = g.has_pre ? 0u : 1u; ph <= lq; ++ph) {
        const uint32_t work = ph == 0 ? T : (T >> 1);
        // work item b -> (lo, hi, twiddle pointer, unit flag)
        auto locate = [&](uint32_t b, uint32_t &lo, uint32_t &hi, const uint32_t *&wp, bool &unit) {
            lo = 0;
            unit = false;  // twiddle == 1: no product (warp uniform by construction)
            if (ph == 0) {
                hi = b;
                const uint32_t slot = b >> lq;
                const uint32_t j = lq ? (__brev(b & Lm1) >> (32 - lq)) : 0u;
                const uint64_t gi = elem_index(g, sub0 + slot, j);
                wp = pre_tw + ((gi >> g.pre_shift) & g.pre_mask) * TW_WORDS;
            } else {
                const uint32_t m = 1u << (ph - 1);
                uint32_t jj, grp;
                const bool twiddle_major = ph >= 2 && ph <= 4 && g.log_tile >= ph + 5;
                if (twiddle_major) {
                    // early stages: enumerate butterflies twiddle-major, so that the T/2m butterflies
                    // with jj == 0 (unit twiddle) fill whole warps and skip the product
                    jj = b >> (g.log_tile - ph);
                    grp = b & ((1u << (g.log_tile - ph)) - 1);
                } else {
                    jj = b & (m - 1);
                    grp = b >> (ph - 1);
                }
                unit = jj == 0 && (ph == 1 || twiddle_major);  // elsewhere unit lanes are scattered: keep warps converged
                lo = (grp << ph) | jj;
                hi = lo + m;
                wp = wloc + (size_t)((jj << (lq - ph)) << g.wloc_shift) * TW_WORDS;
            }
        };
        for (uint32_t b = threadIdx.x; b < work; b += THREADS) {
            uint32_t lo, hi;
            const uint32_t *wp;
            bool unit;
            locate(b, lo, hi, wp, unit);
            uint4 *sh = tile + slot_of(hi) * SMEM_PITCH4;
            uint32_t t[NL];
            if (unit) {
                lds_elem(t, sh);  // unit twiddle
            } else {
                // fixed-operand product: the data streams from shared memory (twice), the twiddle's (w, w'') from the table
                SmemWords x1{sh, make_uint4(0, 0, 0, 0)}, x2{sh, make_uint4(0, 0, 0, 0)};
                shoup_mul_lazy(t, x1, x2, wp);
            }
            if (ph == 0) {
                sts_elem(sh, t);
            } else {
                uint4 *sl = tile + slot_of(lo) * SMEM_PITCH4;
                uint32_t u[NL], x[NL];
                lds_elem(u, sl);
                add_lazy(x, u, t);
                if (g.canonical && ph == lq) canonicalize(x);
                sts_elem(sl, x);
                sub_lazy(x, u, t);
                if (g.canonical && ph == lq) canonicalize(x);
                sts_elem(sh, x);
            }
        }
        __syncthreads();
    }

    // ---- store: consecutive lanes write consecutive 16-byte chunks of an element, so every warp
    // store covers whole 96-byte elements (full sectors locally, full packets over NVLink)
    if (lq == 0 && g.canonical) {  // degenerate n = 1 transforms still leave canonical values
        for (uint32_t e = threadIdx.x; e < T; e += THREADS) {
            const uint64_t go = out_index(g, sub0 + e, 0);
            uint32_t x[NL];
            lds_elem(x, tile + slot_of(e) * SMEM_PITCH4);
            canonicalize(x);
            store_elem(sc.enabled ? scatter_target(sc, go) : dst + go * NL, x);
        }
        return;
    }
    for (uint32_t idx = threadIdx.x; idx < T * 6; idx += THREADS) {
        const uint32_t e = idx / 6, c = idx - e * 6;
        const uint32_t slot = e >> lq, k = e & Lm1;
        const uint64_t go = out_index(g, sub0 + slot, k);
        uint32_t *target = sc.enabled ? scatter_target(sc, go) : dst + go * NL;  // peer memory: plain st.global over NVLink
        reinterpret_cast<uint4 *>(target)[c] = tile[slot_of(e) * SMEM_PITCH4 + c];
    }
}

// ------------------------------------------------------------------ table builders
// out[k] = base^(k * stride), k < count   (square and multiply per thread; tables are tiny
// compared with the transform and are cached per (n, omega))
__global__ void pow_table768(uint32_t *out, const uint32_t *base, uint64_t count, uint64_t stride) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    uint32_t b[NL], acc[NL];
    load_elem(b, base);
#pragma unroll
    for (int i = 0; i < NL; ++i) acc[i] = c_fp.r1[i];
    const uint64_t e = k * stride;
    const int top = 63 - __clzll((long long)(e | 1));
    for (int bit = top; bit >= 0; --bit) {
        uint32_t t[NL];
        mont_mul_lazy(t, acc, acc);
        if ((e >> bit) & 1) mont_mul_lazy(acc, t, b);
        else {
#pragma unroll
            for (int i = 0; i < NL; ++i) acc[i] = t[i];
        }
    }
    canonicalize(acc);
    store_elem(out + k * NL, acc);
}

// Pre-twiddle table of one pass boundary: out[idx] = w_n^(((k * rest) mod N) << exp_shift),
// idx = (k << rest_bits) | rest, assembled from the two-level tables
//   t_lo[e] = w_n^e (e < 2^lo_bits),  t_hi[e] = w_n^(e << lo_bits).
__global__ void build_pretw768(uint32_t *out, const uint32_t *t_lo, const uint32_t *t_hi, uint32_t logN,
                               uint32_t rest_bits, uint32_t exp_shift, uint32_t lo_bits) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >> logN) return;
    const uint64_t k = idx >> rest_bits, rest = idx & ((1ull << rest_bits) - 1);
    const uint64_t e = ((k * rest) & ((1ull << logN) - 1)) << exp_shift;
    uint32_t a[NL], b[NL], r[NL];
    load_elem(a, t_lo + (e & ((1ull << lo_bits) - 1)) * NL);
    load_elem(b, t_hi + (e >> lo_bits) * NL);
    mont_mul(r, a, b);
    store_elem(out + idx * NL, r);
}

// Four-step (Bailey) twiddles of one shard: out[r * cols + c] = w_n^((row0 + r) * (col0 + c) mod n)
// from the two-level tables t_lo[e] = w^e (e < 2^lo_bits), t_hi[e] = w^(e << lo_bits).
__global__ void build_fourstep768(uint32_t *out, const uint32_t *t_lo, const uint32_t *t_hi, uint64_t rows, uint64_t cols,
                                  uint64_t row0, uint64_t col0, uint32_t logn, uint32_t lo_bits) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols) return;
    const uint64_t r = idx / cols, c = idx - r * cols;
    const uint64_t e = ((row0 + r) * (col0 + c)) & ((1ull << logn) - 1);
    uint32_t a[NL], b[NL], o[NL];
    load_elem(a, t_lo + (e & ((1ull << lo_bits) - 1)) * NL);
    load_elem(b, t_hi + (e >> lo_bits) * NL);
    mont_mul(o, a, b);
    store_elem(out + idx * NL, o);
}

// Twiddle tables are built in Montgomery form (the kernels above) and converted once into the fixed-operand
// format of the transform kernels: out[i] = ( w = in[i] * R^-1 (plain, canonical),  w'' = floor(w * 2^768 / p) ).
// Since w*R = w''*p + in[i] exactly,  w'' = lo768(in[i] * (-p^-1 mod 2^768)).
__global__ void to_shoup_table768(uint32_t *out, const uint32_t *in, uint64_t count) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t wm[NL], one[NL], w[NL], w2[NL];
    load_elem(wm, in + i * NL);
#pragma unroll
    for (int k = 0; k < NL; ++k) one[k] = k == 0 ? 1u : 0u;
    mont_mul(w, wm, one);
    mul_lo768(w2, wm, ConstNprime{});
    store_elem(out + i * TW_WORDS, w);
    store_elem(out + i * TW_WORDS + NL, w2);
}

// out[i] = a[i] * s   (canonical); used to fold n^-1 into a table
__global__ void scale_table768(uint32_t *out, const uint32_t *a, const uint32_t *s, uint64_t count) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t x[NL], y[NL], r[NL];
    load_elem(x, a + i * NL);
    load_elem(y, s);
    mont_mul(r, x, y);
    store_elem(out + i * NL, r);
}

// out[i] = scale * base^i (canonical), i < count: coset shifts g^i for coset transforms.
// Thread i computes base^i by square-and-multiply (tables are built once per domain).
__global__ void powers768(uint32_t *out, const uint32_t *base, const uint32_t *scale, uint64_t count) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t b[NL], acc[NL];
    load_elem(b, base);
    load_elem(acc, scale);
    const int top = 63 - __clzll((long long)(i | 1));
    uint32_t r[NL];
#pragma unroll
    for (int k = 0; k < NL; ++k) r[k] = c_fp.r1[k];
    for (int bit = top; bit >= 0; --bit) {
        uint32_t t[NL];
        mont_mul_lazy(t, r, r);
        if ((i >> bit) & 1) mont_mul_lazy(r, t, b);
        else {
#pragma unroll
            for (int k = 0; k < NL; ++k) r[k] = t[k];
        }
    }
    uint32_t o[NL];
    mont_mul(o, r, acc);
    store_elem(out + i * NL, o);
}

// Field inner product sum_i a[i] * b[i]: the reference's multiexp<Scalar, Scalar> (reference
// cuda/multi_exp.cu:86-137 `deviceReduceKernel` + `deviceReduceKernelSecond`, CPU form test/multiexp.h:3-13).
// Stage 1: grid-stride products accumulated lazily per thread, block tree reduction in shared memory,
// one partial per block.  Stage 2 (same kernel, one block, b == nullptr): sums the partials.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) inner_product768(uint32_t *out, const uint32_t *a, const uint32_t *b, uint64_t count) {
    __shared__ uint4 red[THREADS * 6];
    uint32_t acc[NL];
#pragma unroll
    for (int k = 0; k < NL; ++k) acc[k] = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < count; i += (uint64_t)gridDim.x * THREADS) {
        uint32_t x[NL], t[NL];
        load_elem(x, a + i * NL);
        if (b) {
            uint32_t y[NL];
            load_elem(y, b + i * NL);
            mont_mul_lazy(t, x, y);
        } else {
#pragma unroll
            for (int k = 0; k < NL; ++k) t[k] = x[k];
        }
        uint32_t s[NL];
        add_lazy(s, acc, t);
#pragma unroll
        for (int k = 0; k < NL; ++k) acc[k] = s[k];
    }
    sts_elem(red + threadIdx.x * 6, acc);
    __syncthreads();
    for (int half = THREADS / 2; half >= 1; half >>= 1) {
        if ((int)threadIdx.x < half) {
            uint32_t x[NL], y[NL], s[NL];
            lds_elem(x, red + threadIdx.x * 6);
            lds_elem(y, red + (threadIdx.x + half) * 6);
            add_lazy(s, x, y);
            sts_elem(red + threadIdx.x * 6, s);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        uint32_t x[NL];
        lds_elem(x, red);
        canonicalize(x);
        store_elem(out + (uint64_t)blockIdx.x * NL, x);
    }
}

// element-wise field ops for the arithmetic parity tests (canonical results)
// op: 0 mul, 1 add, 2 sub
__global__ void binop768(uint32_t *out, const uint32_t *a, const uint32_t *b, uint64_t count, int op) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t x[NL], y[NL], r[NL];
    load_elem(x, a + i * NL);
    load_elem(y, b + i * NL);
    if (op == 0) mont_mul(r, x, y);
    else if (op == 1) { add_lazy(r, x, y); canonicalize(r); }
    else { sub_lazy(r, x, y); canonicalize(r); }
    store_elem(out + i * NL, r);
}

}  // namespace gsn
