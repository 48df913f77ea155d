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
// stages 2-5 the jj == 0 butterflies, which a twiddle-major enumeration packs into whole warps
// (about 10 % of the products of a 10-stage pass).
//
// Shared memory layout: AoS with a 112-byte pitch (96 B of limbs + 16 B pad).  With a
// 28-word pitch the eight lanes of a quarter-warp that read the same 16-byte chunk of eight
// consecutive elements hit eight distinct bank groups, so LDS.128/STS.128 are conflict
// free for every stage with half-distance >= 8.
#pragma once
#include "fp768.cuh"
#include "v2_index.h"

namespace gsn {

constexpr int SMEM_PITCH4 = 7;  // uint4 per element in shared memory (112 B)
constexpr int TW_WORDS = 2 * NL;  // twiddle table entry: w[24] (plain) followed by w''[24] = floor(w * 2^768 / p)

struct PassGeom {
    uint32_t log_l;          // stages of this pass (digit width)
    uint32_t log_s;          // log2 stride (elements) of this digit
    uint32_t log_r;          // log2 inner stride (elements below the transform index)
    uint32_t log_tile;       // log2 elements per CTA tile (>= log_l)
    uint32_t wloc_shift;     // local table index = (jj << (log_l - s)) << wloc_shift
    uint32_t final_natural;  // last pass: write digits reversed
    uint32_t canonical;      // outputs reduced to [0, p) (else lazy)
    uint32_t logn;           // sum of digits
    uint32_t tile0;          // first tile of this launch (chunked launches of one pass)
    // last pass: output digit q (q < ndig - 1) = (rest >> dpos[q]) & dmask[q], placed at bit dshift[q]; k at kshift
    uint32_t dpos[3], dmask[3], dshift[3], kshift;
    // multi-GPU arrival flags (fused four-step): the tiles of the launch are visited source rank by source rank,
    // starting with wait_first (the local rank: its columns are there first); before loading, a tile whose source
    // rank is r = (tile >> wait_shift) & wait_mask spins until wait_flags[r] >= wait_epoch (null = no wait)
    const uint32_t *wait_flags;
    uint32_t wait_epoch, wait_shift, wait_mask, wait_first;
};

// Pre-twiddle of a pass: every element is multiplied by a table entry before the first stage.
//   mode 1  flat:       entry (gi >> flat_shift) & flat_mask of `tab`  (192 B per element index: one product)
//   mode 2  two-level:  w^e = tab[e & lomask] * tab_hi[e >> lo_bits] with e = ((k * r) mod 2^logN) << exp_shift,
//                       k = ((gi >> k_shift) & k_mask) + k_add,  r = gap((gi >> r_shift) & r_mask) + r_add
//                       (two products per element, tables of 2^lo_bits + 2^(logn - lo_bits) entries)
// gi is the element's global index in the launch's buffer.
struct PreDesc {
    const uint32_t *tab, *tab_hi;
    uint64_t flat_mask, k_mask, k_add, r_mask, r_add;
    uint32_t mode, flat_shift, k_shift, r_shift, lo_bits, exp_shift, logN;
    uint32_t gap_shift, gap_bits, pad;   // r gets gap_bits zero bits inserted at bit gap_shift before r_add (block-cyclic column ownership)
};

// Fused exchange (multi-GPU four-step): the last pass can store each element straight into a
// peer GPU's buffer over NVLink instead of its own.  The local natural output index `go` names its destination rank
// in the bit field [rank_shift, rank_shift + rank_bits); the destination index is what remains of `go` (rem_bits
// bits), rotated right by rot_bits (the low rot_bits bits move to the top: row index <-> column index), with this
// rank's id inserted at bit ins_shift.
struct ScatterDesc {
    uint32_t *peers[8];   // peers[r] = base of rank r's receive buffer (own entry = local pointer)
    uint32_t enabled, rank_shift, rank_bits, ins_shift, my_rank, rot_bits, rem_bits, pad;
};

__device__ __forceinline__ uint32_t *scatter_target(const ScatterDesc &sc, uint64_t go) {
    const uint32_t dest = (uint32_t)(go >> sc.rank_shift) & ((1u << sc.rank_bits) - 1);
    uint64_t rem = ((go >> (sc.rank_shift + sc.rank_bits)) << sc.rank_shift) | (go & ((1ull << sc.rank_shift) - 1));
    rem = (rem >> sc.rot_bits) | ((rem & ((1ull << sc.rot_bits) - 1)) << (sc.rem_bits - sc.rot_bits));
    const uint64_t idx = ((rem >> sc.ins_shift) << (sc.ins_shift + sc.rank_bits)) | ((uint64_t)sc.my_rank << sc.ins_shift) |
                         (rem & ((1ull << sc.ins_shift) - 1));
    return sc.peers[dest] + idx * NL;
}

// Flags of the multi-GPU exchange, in peer memory.  flags[r] = rank r's array of 8 slots (own entry = local array).
struct PeerFlags {
    uint32_t *flags[8];
};

__device__ __forceinline__ void spin_until(const uint32_t *flag, uint32_t epoch) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int32_t)(v - epoch) >= 0) break;
        __nanosleep(100);
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 10000000000ull) __trap();  // a peer that never arrives traps after ~10 s instead of hanging the device
    }
}

// Lane r publishes `epoch` into slot [my_rank] of rank r's flag array (release at system scope: everything this GPU
// stored before the kernel boundary -- the scattered tiles -- is visible to whoever acquires the flag) and, when
// `wait` is set, waits until rank r has published the same epoch into slot [r] of the local array (a full barrier).
// With wait == 0 it only signals; the consumer kernel waits per source rank (PassGeom::wait_flags).
__global__ void peer_barrier_kernel(const PeerFlags pf, uint32_t n_peers, uint32_t my_rank, uint32_t epoch, uint32_t wait) {
    const uint32_t r = threadIdx.x;
    if (r >= n_peers) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pf.flags[r] + my_rank), "r"(epoch) : "memory");
    if (wait) spin_until(pf.flags[my_rank] + r, epoch);
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
    uint64_t out = (uint64_t)k << g.kshift;
#pragma unroll
    for (int q = 0; q < 3; ++q) out |= ((rest >> g.dpos[q]) & g.dmask[q]) << g.dshift[q];
    return (((batch << g.logn) | out) << g.log_r) | rlow;
}

// exponent of the two-level twiddle of element gi
__device__ __forceinline__ uint64_t pre_exponent(const PreDesc &pd, uint64_t gi) {
    const uint64_t k = ((gi >> pd.k_shift) & pd.k_mask) + pd.k_add;
    uint64_t r = (gi >> pd.r_shift) & pd.r_mask;
    r = ((r >> pd.gap_shift) << (pd.gap_shift + pd.gap_bits)) | (r & ((1ull << pd.gap_shift) - 1));
    uint64_t e = k * (r + pd.r_add);
    if (pd.logN < 64) e &= (1ull << pd.logN) - 1;
    return e << pd.exp_shift;
}
// table entry multiplied in round `first` (two-level: low table first, then the high table)
__device__ __forceinline__ const uint32_t *pre_entry(const PreDesc &pd, uint64_t gi, bool first) {
    if (pd.mode == 1) return pd.tab + ((gi >> pd.flat_shift) & pd.flat_mask) * TW_WORDS;
    const uint64_t e = pre_exponent(pd, gi);
    return first ? pd.tab + (e & ((1ull << pd.lo_bits) - 1)) * TW_WORDS : pd.tab_hi + (e >> pd.lo_bits) * TW_WORDS;
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

// ------------------------------------------------------------------ CTA-wide kernel (any tile size)
// CTA-wide enumeration of the butterflies of every stage, one __syncthreads per stage, values in [0, 2p) between stages.
// Stages 2..5 are enumerated twiddle-major, so the butterflies with a unit twiddle fill whole warps and skip the
// product (1.9375 of 10 stages' worth).  (A wide-lazy-range form of this kernel measured 2 % slower and was removed.)
template <int THREADS, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
ntt768_pass(const uint32_t *src, uint32_t *dst, const uint32_t *__restrict__ wloc, const __grid_constant__ PassGeom g,
            const __grid_constant__ PreDesc pd, const __grid_constant__ PreDesc post, const __grid_constant__ ScatterDesc sc,
            const __grid_constant__ FieldConstants768 fc) {
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

    // ---- ph <= 0: pre-twiddle (one product per element and round); ph = s in 1..lq: butterfly stage s; ph > lq: post-twiddle
    const int ph_last = (int)lq + (post.mode == 0 ? 0 : (post.mode == 2 ? 2 : 1));
    for (int ph = pd.mode == 0 ? 1 : (pd.mode == 2 ? -1 : 0); ph <= ph_last; ++ph) {
        const bool elementwise = ph <= 0 || ph > (int)lq;
        const uint32_t work = elementwise ? T : (T >> 1);
        for (uint32_t b = threadIdx.x; b < work; b += THREADS) {
            uint32_t lo = 0, hi;
            const uint32_t *wp;
            bool unit = false;  // twiddle == 1: no product (warp uniform by construction)
            if (ph <= 0) {
                hi = b;
                const uint32_t slot = b >> lq;
                const uint32_t j = lq ? (__brev(b & Lm1) >> (32 - lq)) : 0u;
                wp = pre_entry(pd, elem_index(g, sub0 + slot, j), ph < 0);
            } else if (ph > (int)lq) {
                hi = b;
                wp = pre_entry(post, out_index(g, sub0 + (b >> lq), b & Lm1), post.mode == 2 && ph == (int)lq + 1);
            } else {
                const uint32_t m = 1u << (ph - 1);
                uint32_t jj, grp;
                const bool twiddle_major = ph >= 2 && ph <= 5 && g.log_tile >= (uint32_t)ph + 5;
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
            uint4 *sh = tile + slot_of(hi) * SMEM_PITCH4;
            uint32_t t[NL];
            if (unit) {
                lds_elem(t, sh);  // unit twiddle
            } else {
                // fixed-operand product: the data streams from shared memory (twice), the twiddle's (w, w'') from the table
                SmemWords x1{sh, make_uint4(0, 0, 0, 0)}, x2{sh, make_uint4(0, 0, 0, 0)};
                shoup_mul_lazy(fc, t, x1, x2, wp);
            }
            if (elementwise) {
                if (g.canonical && ph == ph_last) canonicalize(fc, t);
                sts_elem(sh, t);
            } else {
                uint4 *sl = tile + slot_of(lo) * SMEM_PITCH4;
                uint32_t u[NL], x[NL];
                lds_elem(u, sl);
                const bool last = g.canonical && ph == ph_last;
                add_lazy(fc, x, u, t);
                if (last) canonicalize(fc, x);
                sts_elem(sl, x);
                sub_lazy(fc, x, u, t);
                if (last) canonicalize(fc, x);
                sts_elem(sh, x);
            }
        }
        __syncthreads();
    }

    // ---- store: consecutive lanes write consecutive 16-byte chunks of an element, so every warp
    // store covers whole 96-byte elements (full sectors locally, full packets over NVLink)
    if (lq == 0 && g.canonical && post.mode == 0) {  // degenerate n = 1 transforms still leave canonical values
        for (uint32_t e = threadIdx.x; e < T; e += THREADS) {
            const uint64_t go = out_index(g, sub0 + e, 0);
            uint32_t x[NL];
            lds_elem(x, tile + slot_of(e) * SMEM_PITCH4);
            canonicalize(fc, x);
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

// ------------------------------------------------------------------ large-tile kernel: 1024 elements, 8 warps, warp-owned blocks
// Every warp OWNS 128 elements of the tile and runs their stages with no CTA-wide synchronisation:
//   phase A   pos in [128 W, 128 W + 128): the warp loads these elements itself (bit-reversed placement), applies the
//             pre-twiddle and runs stages 1..7 -- all of them stay inside the block -- behind __syncwarp()
//   (one __syncthreads, only if the pass has more than 7 stages)
//   phase B   pos = (h << 7) | (W << 4) | l, h < 8, l < 16: closed under stages 8..10, again warp local; the warp then
//             stores the elements it owns.
// One CTA barrier per tile instead of one per stage: the eight warps drift apart, so one warp's global loads, twiddle
// fetches and add/sub carry chains overlap the other warps' products instead of lining up behind a barrier.
// Wide lazy ranges throughout (no conditional subtractions inside the pass, see fp768.cuh).
// Unit twiddles: all of stage 1 and the jj == 0 half of stage 2 (enumerated twiddle-major: one whole iteration) skip the
// product; in later stages the jj == 0 lanes multiply by table entry 0 = (1, floor(2^768/p)).
// FLAGS & V2_EARLY_CTA: stages 3 and 4 are enumerated CTA-wide and twiddle-major instead (three more CTA barriers), so
// that their unit-twiddle butterflies fill whole warps and skip the product as in ntt768_pass: 8.125 instead of 8.5
// stages' worth of products per ten-stage pass.
constexpr int V2_LAZY = 1, V2_EARLY_CTA = 4;

struct WorkItem {
    const uint32_t *wp;
    uint32_t lo, hi;
    bool unit;
};

// stage ph of the tile is enumerated CTA-wide (not by the owning warp)
template <int FLAGS>
__device__ __forceinline__ bool v2_cta_stage(int ph, uint32_t lq) { return (FLAGS & V2_EARLY_CTA) && (ph == 3 || ph == 4) && ph <= (int)lq; }

template <int FLAGS>
__device__ __forceinline__ WorkItem v2_locate(const PassGeom &g, const PreDesc &pd, const PreDesc &post, const uint32_t *wloc, uint64_t sub0,
                                              uint32_t W, uint32_t lane, int ph, int it) {
    WorkItem w;
    const uint32_t lq = g.log_l;
    if (ph > (int)lq) {  // post-twiddle of the elements the warp owns after the last stage
        const uint32_t pos = lq > 7 ? own_b(W, lane + 32 * it) : own_a(W, lane + 32 * it);
        w.lo = 0;
        w.hi = pos;
        w.unit = false;
        w.wp = pre_entry(post, out_index(g, sub0 + (pos >> lq), pos & ((1u << lq) - 1)), post.mode == 2 && ph == (int)lq + 1);
    } else if (ph <= 0) {
        const uint32_t pos = own_a(W, lane + 32 * it);
        const uint32_t j = __brev(pos & ((1u << lq) - 1)) >> (32 - lq);
        w.lo = 0;
        w.hi = pos;
        w.unit = false;
        w.wp = pre_entry(pd, elem_index(g, sub0 + (pos >> lq), j), ph < 0);
    } else {
        uint32_t jj;
        if (v2_cta_stage<FLAGS>(ph, lq)) {
            cta_butterfly(ph, threadIdx.x + 256 * it, w.lo, jj);
            w.unit = jj == 0;   // warp uniform: jj = (threadIdx.x + 256 it) >> (10 - ph)
        } else {
            v2_butterfly(W, ph, lane + 32 * it, w.lo, jj);
            w.unit = v2_unit(ph, it);
        }
        w.hi = w.lo + (1u << (ph - 1));
        w.wp = wloc + (size_t)((jj << (lq - ph)) << g.wloc_shift) * TW_WORDS;
    }
    return w;
}

template <int FLAGS>
__global__ void __launch_bounds__(256, 2)
ntt768_pass2(const uint32_t *src, uint32_t *dst, const uint32_t *__restrict__ wloc, const __grid_constant__ PassGeom g,
             const __grid_constant__ PreDesc pd, const __grid_constant__ PreDesc post, const __grid_constant__ ScatterDesc sc,
             const __grid_constant__ FieldConstants768 fc) {
    extern __shared__ uint4 tile[];
    const uint32_t lq = g.log_l;  // 1..10, log_tile == 10
    const uint32_t Lm1 = (1u << lq) - 1;
    const uint32_t W = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t tile_id = blockIdx.x + g.tile0;
    if (g.wait_flags) {  // fused four-step: wait until this tile's source rank has delivered its columns
        // block b -> source rank (wait_first + b / per_src) mod G, tile = the (b mod per_src)-th tile of that source
        const uint32_t logG = 32 - __clz(g.wait_mask), per_src = gridDim.x >> logG;
        const uint32_t src_rank = (g.wait_first + blockIdx.x / per_src) & g.wait_mask, idx = blockIdx.x % per_src;
        tile_id = ((idx >> g.wait_shift) << (g.wait_shift + logG)) | (src_rank << g.wait_shift) | (idx & ((1u << g.wait_shift) - 1));
        if (lane == 0) spin_until(g.wait_flags + src_rank, g.wait_epoch);
        __syncwarp();
    }
    const uint64_t sub0 = (uint64_t)tile_id << (10 - lq);
    // ---- load the warp's own block: lane per 16-byte chunk (whole 96-byte elements per warp load)
#pragma unroll 4
    for (uint32_t idx = lane; idx < 128 * 6; idx += 32) {
        const uint32_t i = idx / 6, c = idx - i * 6;
        const uint32_t pos = own_a(W, i);
        const uint32_t j = __brev(pos & Lm1) >> (32 - lq);
        const uint64_t gi = elem_index(g, sub0 + (pos >> lq), j);
        tile[slot_of(pos) * SMEM_PITCH4 + c] = reinterpret_cast<const uint4 *>(src + gi * NL)[c];
    }
    __syncwarp();

    const int ph_last = (int)lq + (post.mode == 0 ? 0 : (post.mode == 2 ? 2 : 1));
    // ONE inlined copy of the product (about 1300 instructions): neither loop may be unrolled
#pragma unroll 1
    for (int ph = pd.mode == 0 ? 1 : (pd.mode == 2 ? -1 : 0); ph <= ph_last; ++ph) {
        const bool elementwise = ph <= 0 || ph > (int)lq;
        const bool last = g.canonical && ph == ph_last;
        const int its = elementwise ? 4 : 2;
#pragma unroll 1
        for (int it = 0; it < its; ++it) {
            const WorkItem cur = v2_locate<FLAGS>(g, pd, post, wloc, sub0, W, lane, ph, it);
            uint4 *sh = tile + slot_of(cur.hi) * SMEM_PITCH4;
            uint32_t t[NL], d[NL];  // d = K p - t
            if (cur.unit) {         // unit twiddle: t is taken as it is, < 3p 2^(ph-1) (ph <= 4), and subtracted from that bound
                lds_elem(t, sh);
                uint32_t kp[NL];
                const uint32_t sft = ph - 1;
                kp[0] = fc.p3[0] << sft;
#pragma unroll
                for (int k = 1; k < NL; ++k) kp[k] = __funnelshift_l(fc.p3[k - 1], fc.p3[k], sft);
                neg_wide(d, kp, t);
            } else {
                SmemWords x1{sh, make_uint4(0, 0, 0, 0)}, x2{sh, make_uint4(0, 0, 0, 0)};
                uint32_t w2[NL];
                load_tw_half(w2, cur.wp + NL);
                shoup_mul_3p(fc, t, x1, x2, w2, cur.wp);
                if (!elementwise) neg_wide(d, fc.p3, t);
            }
            if (elementwise) {
                if (last) {  // last post-twiddle round: t in [0, 3p) -> [0, p)
                    cond_sub(t, fc.p2);
                    cond_sub(t, fc.p);
                }
                sts_elem(sh, t);
            } else {
                uint4 *sl = tile + slot_of(cur.lo) * SMEM_PITCH4;
                uint32_t u[NL], x[NL];
                lds_elem(u, sl);
                add_raw(x, u, t);
                if (last) reduce_small(fc, x);
                sts_elem(sl, x);
                add_raw(x, u, d);
                if (last) reduce_small(fc, x);
                sts_elem(sh, x);
            }
        }
        if (ph < ph_last) {
            // a CTA barrier where the next stage reads elements another warp wrote: into / out of a CTA-wide stage, and
            // where ownership changes from phase A blocks to phase B sets; otherwise the warp only waits for itself
            const bool cta = v2_cta_stage<FLAGS>(ph, lq) || v2_cta_stage<FLAGS>(ph + 1, lq) || (ph == 7 && lq > 7);
            if (cta) __syncthreads();
            else __syncwarp();
        }
    }
    __syncwarp();

    // ---- store the elements this warp owns (phase B set if the pass had more than 7 stages)
#pragma unroll 4
    for (uint32_t idx = lane; idx < 128 * 6; idx += 32) {
        const uint32_t i = idx / 6, c = idx - i * 6;
        const uint32_t pos = lq > 7 ? own_b(W, i) : own_a(W, i);
        const uint64_t go = out_index(g, sub0 + (pos >> lq), pos & Lm1);
        uint32_t *target = sc.enabled ? scatter_target(sc, go) : dst + go * NL;  // peer memory: plain st.global over NVLink
        reinterpret_cast<uint4 *>(target)[c] = tile[slot_of(pos) * SMEM_PITCH4 + c];
    }
}

// ------------------------------------------------------------------ table builders
// out[k] = base^(k * stride), k < count   (square and multiply per thread; tables are tiny
// compared with the transform and are cached per (n, omega))
__global__ void pow_table768(const __grid_constant__ FieldConstants768 fc, uint32_t *out, const uint32_t *base, uint64_t count, uint64_t stride) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    uint32_t b[NL], acc[NL];
    load_elem(b, base);
#pragma unroll
    for (int i = 0; i < NL; ++i) acc[i] = fc.r1[i];
    const uint64_t e = k * stride;
    const int top = 63 - __clzll((long long)(e | 1));
    for (int bit = top; bit >= 0; --bit) {
        uint32_t t[NL];
        mont_mul_lazy(fc, t, acc, acc);
        if ((e >> bit) & 1) mont_mul_lazy(fc, acc, t, b);
        else {
#pragma unroll
            for (int i = 0; i < NL; ++i) acc[i] = t[i];
        }
    }
    canonicalize(fc, acc);
    store_elem(out + k * NL, acc);
}

// Pre-twiddle table of one pass boundary: out[idx] = w_n^(((k * rest) mod N) << exp_shift),
// idx = (k << rest_bits) | rest, assembled from the two-level tables
//   t_lo[e] = w_n^e (e < 2^lo_bits),  t_hi[e] = w_n^(e << lo_bits).
__global__ void build_pretw768(const __grid_constant__ FieldConstants768 fc, uint32_t *out, const uint32_t *t_lo, const uint32_t *t_hi, uint32_t logN,
                               uint32_t rest_bits, uint32_t exp_shift, uint32_t lo_bits) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >> logN) return;
    const uint64_t k = idx >> rest_bits, rest = idx & ((1ull << rest_bits) - 1);
    const uint64_t e = ((k * rest) & ((1ull << logN) - 1)) << exp_shift;
    uint32_t a[NL], b[NL], r[NL];
    load_elem(a, t_lo + (e & ((1ull << lo_bits) - 1)) * NL);
    load_elem(b, t_hi + (e >> lo_bits) * NL);
    mont_mul(fc, r, a, b);
    store_elem(out + idx * NL, r);
}

// Flat pre-twiddle table from a two-level descriptor: out[gi] = t_lo[e & lomask] * t_hi[e >> lo_bits], e = pre_exponent(pd, gi),
// gi < count.  Here pd.tab / pd.tab_hi point at MONTGOMERY-form tables (96 B per entry); the result is Montgomery form,
// canonical, and goes through to_shoup_table768 afterwards.
__global__ void materialize_pre768(const __grid_constant__ FieldConstants768 fc, uint32_t *out, const __grid_constant__ PreDesc pd, uint64_t count) {
    const uint64_t gi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= count) return;
    const uint64_t e = pre_exponent(pd, gi);
    uint32_t a[NL], b[NL], r[NL];
    load_elem(a, pd.tab + (e & ((1ull << pd.lo_bits) - 1)) * NL);
    load_elem(b, pd.tab_hi + (e >> pd.lo_bits) * NL);
    mont_mul(fc, r, a, b);
    store_elem(out + gi * NL, r);
}

// Four-step (Bailey) twiddles of one shard: out[r * cols + c] = w_n^((row0 + r) * (col0 + c) mod n)
// from the two-level tables t_lo[e] = w^e (e < 2^lo_bits), t_hi[e] = w^(e << lo_bits).
__global__ void build_fourstep768(const __grid_constant__ FieldConstants768 fc, uint32_t *out, const uint32_t *t_lo, const uint32_t *t_hi, uint64_t rows, uint64_t cols,
                                  uint64_t row0, uint64_t col0, uint32_t logn, uint32_t lo_bits) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols) return;
    const uint64_t r = idx / cols, c = idx - r * cols;
    const uint64_t e = ((row0 + r) * (col0 + c)) & ((1ull << logn) - 1);
    uint32_t a[NL], b[NL], o[NL];
    load_elem(a, t_lo + (e & ((1ull << lo_bits) - 1)) * NL);
    load_elem(b, t_hi + (e >> lo_bits) * NL);
    mont_mul(fc, o, a, b);
    store_elem(out + idx * NL, o);
}

// Twiddle tables are built in Montgomery form (the kernels above) and converted once into the fixed-operand
// format of the transform kernels: out[i] = ( w = in[i] * R^-1 (plain, canonical),  w'' = floor(w * 2^768 / p) ).
// Since w*R = w''*p + in[i] exactly,  w'' = lo768(in[i] * (-p^-1 mod 2^768)).
__global__ void to_shoup_table768(const __grid_constant__ FieldConstants768 fc, uint32_t *out, const uint32_t *in, uint64_t count) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t wm[NL], one[NL], w[NL], w2[NL];
    load_elem(wm, in + i * NL);
#pragma unroll
    for (int k = 0; k < NL; ++k) one[k] = k == 0 ? 1u : 0u;
    mont_mul(fc, w, wm, one);
    mul_lo768(w2, wm, ConstNprime{fc});
    store_elem(out + i * TW_WORDS, w);
    store_elem(out + i * TW_WORDS + NL, w2);
}

// out[i] = a[i] * s   (canonical); used to fold n^-1 into a table
__global__ void scale_table768(const __grid_constant__ FieldConstants768 fc, uint32_t *out, const uint32_t *a, const uint32_t *s, uint64_t count) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t x[NL], y[NL], r[NL];
    load_elem(x, a + i * NL);
    load_elem(y, s);
    mont_mul(fc, r, x, y);
    store_elem(out + i * NL, r);
}

// out[i] = scale * base^i (canonical), i < count: coset shifts g^i for coset transforms.
// Thread i computes base^i by square-and-multiply (tables are built once per domain).
struct Elem768 { uint32_t v[NL]; };   // one field element by value (kernel parameter: no device allocation, no copy to wait for)

__global__ void powers768(const __grid_constant__ FieldConstants768 fc, uint32_t *out, const __grid_constant__ Elem768 base,
                          const __grid_constant__ Elem768 scale, uint64_t count) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t b[NL], acc[NL];
#pragma unroll
    for (int k = 0; k < NL; ++k) { b[k] = base.v[k]; acc[k] = scale.v[k]; }
    const int top = 63 - __clzll((long long)(i | 1));
    uint32_t r[NL];
#pragma unroll
    for (int k = 0; k < NL; ++k) r[k] = fc.r1[k];
    for (int bit = top; bit >= 0; --bit) {
        uint32_t t[NL];
        mont_mul_lazy(fc, t, r, r);
        if ((i >> bit) & 1) mont_mul_lazy(fc, r, t, b);
        else {
#pragma unroll
            for (int k = 0; k < NL; ++k) r[k] = t[k];
        }
    }
    uint32_t o[NL];
    mont_mul(fc, o, r, acc);
    store_elem(out + i * NL, o);
}

// Field inner product sum_i a[i] * b[i]: the reference's multiexp<Scalar, Scalar> (reference
// cuda/multi_exp.cu:86-137 `deviceReduceKernel` + `deviceReduceKernelSecond`, CPU form test/multiexp.h:3-13).
// Stage 1: grid-stride products accumulated lazily per thread, block tree reduction in shared memory,
// one partial per block.  Stage 2 (same kernel, one block, b == nullptr): sums the partials.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) inner_product768(const __grid_constant__ FieldConstants768 fc, uint32_t *out, const uint32_t *a, const uint32_t *b, uint64_t count) {
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
            mont_mul_lazy(fc, t, x, y);
        } else {
#pragma unroll
            for (int k = 0; k < NL; ++k) t[k] = x[k];
        }
        uint32_t s[NL];
        add_lazy(fc, s, acc, t);
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
            add_lazy(fc, s, x, y);
            sts_elem(red + threadIdx.x * 6, s);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        uint32_t x[NL];
        lds_elem(x, red);
        canonicalize(fc, x);
        store_elem(out + (uint64_t)blockIdx.x * NL, x);
    }
}

// element-wise field ops for the arithmetic parity tests (canonical results)
// op: 0 mul, 1 add, 2 sub
__global__ void binop768(const __grid_constant__ FieldConstants768 fc, uint32_t *out, const uint32_t *a, const uint32_t *b, uint64_t count, int op) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t x[NL], y[NL], r[NL];
    load_elem(x, a + i * NL);
    load_elem(y, b + i * NL);
    if (op == 0) mont_mul(fc, r, x, y);
    else if (op == 1) { add_lazy(fc, r, x, y); canonicalize(fc, r); }
    else { sub_lazy(fc, r, x, y); canonicalize(fc, r); }
    store_elem(out + i * NL, r);
}

}  // namespace gsn
