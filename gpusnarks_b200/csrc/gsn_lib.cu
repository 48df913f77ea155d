// gsn_lib.cu -- host planner + C ABI of libgpusnarks_b200.so (see include/gpusnarks_b200.h).
//
// Replaces the reference's host launcher best_fft (reference cuda/fft_kernel.cu:117-147):
// where the reference cudaMallocs two buffers per call, copies, launches one kernel of
// 5 x 256 threads and never frees, this keeps a context with a stream, a reusable
// workspace and per-(n, omega) plans whose twiddle tables are computed on the device.
// There is no CPU fallback: without a CUDA device every entry point fails with
// GSN_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

#include "../../include/gpusnarks_b200.h"
#include "../../include/gsn_constants.h"
#include "fp768.cuh"
#include "host_pool.h"
#include "g1.cuh"
#include "../../include/fields/g1_host.h"
#include "../../include/fields/fp768_host.h"
#include "microbench.cuh"
#include "ntt32.cuh"
#include "ntt32_fast.cuh"
#include "ntt768.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) return fail(GSN_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

constexpr int MAX_PASS_LOG = 10;  // stages per shared-memory pass (tile of 1024 x 96 B)
constexpr int NTT768_THREADS = 256;

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};

struct Plan768 {
    int field = 0;
    uint32_t logn = 0;
    int inverse = 0;              // use omega^-1
    int scale = 0;                // fold n^-1 into the transform
    uint32_t omega[24];           // as passed by the caller (forward root)
    std::vector<uint32_t> digits; // l_1..l_P
    uint32_t lmax = 0;
    DevBuf wloc;                  // w_T^k, k < T/2, T = 2^lmax
    // pre-twiddles of pass q (boundary q-1 | q): flat table (192 B per element index of the boundary's sub-problem)
    // or, when that table would exceed the context's flat-table limit, the two-level tables below (two products)
    std::vector<std::unique_ptr<DevBuf>> pre;  // pre[q]: flat table read by pass q (nullptr if none / two-level)
    std::vector<uint64_t> pre_mask;
    std::vector<uint8_t> pre_two_level;        // pass q multiplies by t_lo[e & lomask] * t_hi[e >> lo_bits]
    DevBuf t_lo, t_lo_scaled, t_hi;            // fixed-operand format; t_lo_scaled = n^-1 * t_lo (boundary 1 of a scaled plan)
    uint32_t lo_bits = 0;
    size_t table_bytes = 0;
    uint64_t last_use = 0;
};

struct Plan32 {
    uint32_t mod = 0, omega = 0, logn = 0;
    int inverse = 0;
    std::vector<uint32_t> digits;
    uint32_t lmax = 0;
    DevBuf wloc;                  // pairs (w, w') Shoup form
    std::vector<std::unique_ptr<DevBuf>> pre;
    std::vector<uint64_t> pre_mask;
    // fast path (two passes, digits of 9..12 stages): in-tile four-step tables per pass and the
    // two-level power tables for the inter-pass twiddle
    bool fast = false;
    std::vector<std::unique_ptr<DevBuf>> tA, tB, tG;  // per pass
    DevBuf t_lo, t_lo_scaled, t_hi;                    // two-level powers of w_n (Montgomery-form low tables)
    uint32_t lo_bits = 0;
    std::vector<gsn::Ntt32Consts> consts;              // per pass
};

std::vector<uint32_t> plan_digits(uint32_t logn, uint32_t max_log) {
    if (logn == 0) return {0};
    const uint32_t npass = (logn + max_log - 1) / max_log;
    const uint32_t base = logn / npass, extra = logn % npass;
    std::vector<uint32_t> d(npass);
    for (uint32_t i = 0; i < npass; ++i) d[i] = base + (i < extra ? 1 : 0);
    return d;
}

bool is_pow2(size_t n) { return n && !(n & (n - 1)); }
uint32_t ilog2(size_t n) { uint32_t l = 0; while (((size_t)1 << l) < n) ++l; return l; }

struct StreamWork {   // scratch buffer of the multi-pass transforms, one per stream that has used the context
    cudaStream_t stream = nullptr;
    DevBuf buf;
    uint64_t last_use = 0;
};

}  // namespace

struct gsn_coset_entry;

struct gsn_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    int field = GSN_FIELD_MNT4753_FR;
    gsn::FieldConstants768 fc;     // travels with every launch as a kernel parameter: per context, never uploaded
    gsn::host::Field768 hf;
    int two_adicity = 30;
    std::vector<std::unique_ptr<StreamWork>> works;  // keyed by stream: concurrent streams never share scratch
    DevBuf io;   // device staging buffer of the host-pointer entry points (grown on demand, reused)
    DevBuf io2;  // second staging buffer: the batch entry point alternates between the two
    cudaEvent_t ev_io_free[2] = {nullptr, nullptr};  // D2H of the transform that last used io / io2 has finished
    std::vector<std::unique_ptr<Plan768>> plans768;
    std::vector<std::unique_ptr<Plan32>> plans32;
    std::vector<std::unique_ptr<gsn_coset_entry>> cosets;   // cached coset shift tables (fourstep_host.inl)
    size_t flat_table_limit = (size_t)4 << 30;   // a pre-twiddle table larger than this becomes two-level
    size_t plan_cache_bytes = (size_t)16 << 30;  // least-recently-used 768-bit plans are dropped beyond this (and beyond 16 plans)
    uint64_t use_clock = 0;
    int v2_flags = 4;  // kernel for 1024-element tiles: 4 / -1 = CTA-wide, 1 = warp-owned tiles (lazy ranges), 5 = + CTA-wide stages 3-4
    uint64_t launches = 0;
    int sm_count = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;  // copy streams of the pipelined host-pointer path
    cudaEvent_t ev_chunk[2][16] = {{nullptr}};
    std::unordered_set<const void *> smem_configured;  // kernels whose dynamic shared-memory limit was raised on this device
    // pinned bounce buffers of the pageable host path (two per direction) and the host threads that fill / drain them
    void *bounce[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t bounce_bytes = 0;
    cudaEvent_t ev_bounce[2] = {nullptr, nullptr};
    std::unique_ptr<HostPool> pool;
    gsn_ctx();
    ~gsn_ctx();
};

namespace {

int set_field(gsn_ctx *ctx, int field) {
    static const uint32_t fr_p[24] = GSN_FR_MOD, fr_p2[24] = GSN_FR_MOD2, fr_p3[24] = GSN_FR_MOD3, fr_p6[24] = GSN_FR_MOD6, fr_r1[24] = GSN_FR_R1, fr_r2[24] = GSN_FR_R2;
    static const uint32_t fq_p[24] = GSN_FQ_MOD, fq_p2[24] = GSN_FQ_MOD2, fq_p3[24] = GSN_FQ_MOD3, fq_p6[24] = GSN_FQ_MOD6, fq_r1[24] = GSN_FQ_R1, fq_r2[24] = GSN_FQ_R2;
    static const uint32_t fr_np[24] = GSN_FR_NPRIME768, fq_np[24] = GSN_FQ_NPRIME768;
    const bool fr = field == GSN_FIELD_MNT4753_FR;
    memset(&ctx->fc, 0, sizeof(ctx->fc));
    memcpy(ctx->fc.p, fr ? fr_p : fq_p, 96);
    memcpy(ctx->fc.p2, fr ? fr_p2 : fq_p2, 96);
    memcpy(ctx->fc.p3, fr ? fr_p3 : fq_p3, 96);
    memcpy(ctx->fc.p6, fr ? fr_p6 : fq_p6, 96);
    memcpy(ctx->fc.r1, fr ? fr_r1 : fq_r1, 96);
    memcpy(ctx->fc.r2, fr ? fr_r2 : fq_r2, 96);
    ctx->fc.np0 = fr ? GSN_FR_NP0 : GSN_FQ_NP0;
    ctx->fc.qmagic = fr ? GSN_FR_QMAGIC : GSN_FQ_QMAGIC;
    memcpy(ctx->fc.nprime, fr ? fr_np : fq_np, 96);
    ctx->two_adicity = fr ? GSN_FR_TWO_ADICITY : GSN_FQ_TWO_ADICITY;
    ctx->hf.init(ctx->fc.p, ctx->fc.r1);
    ctx->field = field;
    return GSN_OK;
}

// scratch of `bytes` for work enqueued on stream st (kernels of different streams may run concurrently)
int ensure_work(gsn_ctx *ctx, cudaStream_t st, size_t bytes, uint32_t **out) {
    StreamWork *w = nullptr;
    for (auto &sw : ctx->works) if (sw->stream == st) w = sw.get();
    if (!w) {
        if (ctx->works.size() >= 8) {  // drop the least recently used stream's scratch (cudaFree waits for its kernels)
            size_t lru = 0;
            for (size_t i = 1; i < ctx->works.size(); ++i) if (ctx->works[i]->last_use < ctx->works[lru]->last_use) lru = i;
            ctx->works.erase(ctx->works.begin() + lru);
        }
        ctx->works.push_back(std::make_unique<StreamWork>());
        w = ctx->works.back().get();
        w->stream = st;
    }
    w->last_use = ++ctx->use_clock;
    if (w->buf.bytes < bytes) {
        if (w->buf.p) { cudaFree(w->buf.p); w->buf.p = nullptr; w->buf.bytes = 0; }
        cudaError_t e = cudaMalloc(&w->buf.p, bytes);
        if (e != cudaSuccess) { cudaGetLastError(); w->buf.p = nullptr; return fail(GSN_ERR_TOO_LARGE, "workspace of %zu bytes: %s", bytes, cudaGetErrorString(e)); }
        w->buf.bytes = bytes;
    }
    *out = (uint32_t *)w->buf.p;
    return GSN_OK;
}

int ensure_io(gsn_ctx *ctx, size_t bytes) {
    if (ctx->io.bytes >= bytes) return GSN_OK;
    if (ctx->io.p) { cudaFree(ctx->io.p); ctx->io.p = nullptr; ctx->io.bytes = 0; }
    cudaError_t e = cudaMalloc(&ctx->io.p, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(GSN_ERR_TOO_LARGE, "staging buffer of %zu bytes: %s", bytes, cudaGetErrorString(e)); }
    ctx->io.bytes = bytes;
    return GSN_OK;
}

int dev_alloc(DevBuf &b, size_t bytes) {
    cudaError_t e = cudaMalloc(&b.p, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); b.p = nullptr; return fail(GSN_ERR_TOO_LARGE, "table of %zu bytes: %s", bytes, cudaGetErrorString(e)); }
    b.bytes = bytes;
    return GSN_OK;
}

// Montgomery-form table (count x 96 B) -> fixed-operand format (count x 192 B) used by the transform kernels
int convert_to_shoup(gsn_ctx *ctx, uint32_t *d_out, const uint32_t *d_in, uint64_t count, cudaStream_t st) {
    gsn::to_shoup_table768<<<(unsigned)((count + 127) / 128), 128, 0, st>>>(ctx->fc, d_out, d_in, count);
    ctx->launches++;
    CU(cudaGetLastError());
    return GSN_OK;
}

// Two-level power tables of `base` (Montgomery form, on the host) in fixed-operand format:
//   t_lo[e] = scale * base^e, e < 2^lo_bits;   t_hi[e] = base^(e << lo_bits), e < 2^hi_bits     (scale may be null)
// built on the device: per-thread square-and-multiply for the 2^lo_bits + 2^hi_bits entries.
int build_two_level(gsn_ctx *ctx, const uint64_t *base_h, const uint64_t *scale_h, uint32_t lo_bits, uint32_t hi_bits, DevBuf &t_lo, DevBuf *t_lo_scaled,
                    DevBuf &t_hi, cudaStream_t st) {
    int rc;
    DevBuf d_w, d_sc, lo_m, hi_m, lo_s;
    const uint64_t nlo = 1ull << lo_bits, nhi = 1ull << hi_bits;
    if ((rc = dev_alloc(d_w, 96)) || (rc = dev_alloc(lo_m, nlo * 96)) || (rc = dev_alloc(hi_m, nhi * 96))) return rc;
    if ((rc = dev_alloc(t_lo, nlo * 192)) || (rc = dev_alloc(t_hi, nhi * 192))) return rc;
    CU(cudaMemcpyAsync(d_w.p, base_h, 96, cudaMemcpyHostToDevice, st));
    gsn::pow_table768<<<(unsigned)((nlo + 127) / 128), 128, 0, st>>>(ctx->fc, (uint32_t *)lo_m.p, (const uint32_t *)d_w.p, nlo, 1);
    gsn::pow_table768<<<(unsigned)((nhi + 127) / 128), 128, 0, st>>>(ctx->fc, (uint32_t *)hi_m.p, (const uint32_t *)d_w.p, nhi, nlo);
    ctx->launches += 2;
    if ((rc = convert_to_shoup(ctx, (uint32_t *)t_lo.p, (const uint32_t *)lo_m.p, nlo, st))) return rc;
    if ((rc = convert_to_shoup(ctx, (uint32_t *)t_hi.p, (const uint32_t *)hi_m.p, nhi, st))) return rc;
    if (scale_h && t_lo_scaled) {
        if ((rc = dev_alloc(d_sc, 96)) || (rc = dev_alloc(lo_s, nlo * 96)) || (rc = dev_alloc(*t_lo_scaled, nlo * 192))) return rc;
        CU(cudaMemcpyAsync(d_sc.p, scale_h, 96, cudaMemcpyHostToDevice, st));
        gsn::scale_table768<<<(unsigned)((nlo + 127) / 128), 128, 0, st>>>(ctx->fc, (uint32_t *)lo_s.p, (const uint32_t *)lo_m.p, (const uint32_t *)d_sc.p, nlo);
        ctx->launches++;
        if ((rc = convert_to_shoup(ctx, (uint32_t *)t_lo_scaled->p, (const uint32_t *)lo_s.p, nlo, st))) return rc;
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));  // the Montgomery-form temporaries are freed on return
    return GSN_OK;
}

// ------------------------------------------------------------------------------ 768-bit plans
int validate_omega768(gsn_ctx *ctx, const uint32_t *omega, uint32_t logn) {
    uint64_t w[12];
    memcpy(w, omega, 96);
    if (!gsn::host::Field768::geq(ctx->hf.p, w) || memcmp(w, ctx->hf.p, 96) == 0)
        return fail(GSN_ERR_BAD_OMEGA, "omega is not reduced modulo p");
    if (logn == 0) return ctx->hf.is_one(w) ? GSN_OK : fail(GSN_ERR_BAD_OMEGA, "n = 1 needs omega = one()");
    for (uint32_t i = 0; i + 1 < logn; ++i) ctx->hf.mul(w, w, w);  // w^(n/2)
    if (!ctx->hf.is_minus_one(w)) return fail(GSN_ERR_BAD_OMEGA, "omega^(n/2) != -1: not a primitive 2^%u-th root of unity", logn);
    return GSN_OK;
}

void evict_plans768(gsn_ctx *ctx, const Plan768 *keep) {
    for (;;) {
        size_t total = 0;
        for (auto &pl : ctx->plans768) total += pl->table_bytes;
        if (ctx->plans768.size() <= 16 && total <= ctx->plan_cache_bytes) return;
        size_t lru = ctx->plans768.size();
        for (size_t i = 0; i < ctx->plans768.size(); ++i)
            if (ctx->plans768[i].get() != keep && (lru == ctx->plans768.size() || ctx->plans768[i]->last_use < ctx->plans768[lru]->last_use)) lru = i;
        if (lru == ctx->plans768.size()) return;
        cudaDeviceSynchronize();  // kernels in flight may still read the tables
        ctx->plans768.erase(ctx->plans768.begin() + lru);
    }
}

int get_plan768(gsn_ctx *ctx, uint32_t logn, const uint32_t *omega, int inverse, int scale, Plan768 **out) {
    for (auto &pl : ctx->plans768)
        if (pl->field == ctx->field && pl->logn == logn && pl->inverse == (inverse != 0) && pl->scale == (scale != 0) && memcmp(pl->omega, omega, 96) == 0) {
            pl->last_use = ++ctx->use_clock;
            *out = pl.get();
            return GSN_OK;
        }
    if ((int)logn > ctx->two_adicity) return fail(GSN_ERR_TOO_LARGE, "n = 2^%u exceeds the field's 2-adicity %d", logn, ctx->two_adicity);
    int rc = validate_omega768(ctx, omega, logn);
    if (rc) return rc;

    auto pl = std::make_unique<Plan768>();
    pl->field = ctx->field;
    pl->logn = logn;
    pl->inverse = inverse != 0;
    pl->scale = scale != 0;
    memcpy(pl->omega, omega, 96);
    pl->digits = plan_digits(logn, MAX_PASS_LOG);
    pl->lmax = *std::max_element(pl->digits.begin(), pl->digits.end());
    const size_t P = pl->digits.size();
    pl->pre.resize(P);
    pl->pre_mask.assign(P, 0);
    pl->pre_two_level.assign(P, 0);

    const uint64_t n = 1ull << logn;
    // effective root (omega or omega^-1 = omega^(n-1)) and n^-1, Montgomery form, on the host
    uint64_t w_eff[12], n_inv[12];
    memcpy(w_eff, omega, 96);
    if (inverse) ctx->hf.pow(w_eff, w_eff, n - 1);
    if (scale) {
        memcpy(n_inv, ctx->hf.r1, 96);
        for (uint32_t i = 0; i < logn; ++i) ctx->hf.halve(n_inv, n_inv);
    }
    cudaStream_t st = ctx->stream;
    DevBuf d_w, d_ninv, t_lo, t_hi;
    if ((rc = dev_alloc(d_w, 96))) return rc;
    CU(cudaMemcpyAsync(d_w.p, w_eff, 96, cudaMemcpyHostToDevice, st));
    if (scale) {
        if ((rc = dev_alloc(d_ninv, 96))) return rc;
        CU(cudaMemcpyAsync(d_ninv.p, n_inv, 96, cudaMemcpyHostToDevice, st));
    }
    // local table: w_T^k, k < T/2
    const uint64_t half = pl->lmax ? (1ull << (pl->lmax - 1)) : 1;
    DevBuf wloc_m;  // Montgomery form, converted below
    if ((rc = dev_alloc(wloc_m, half * 96)) || (rc = dev_alloc(pl->wloc, half * 192))) return rc;
    pl->table_bytes += half * 192;
    gsn::pow_table768<<<(unsigned)((half + 127) / 128), 128, 0, st>>>(ctx->fc, (uint32_t *)wloc_m.p, (const uint32_t *)d_w.p, half,
                                                                        pl->lmax ? (n >> pl->lmax) : 0);
    ctx->launches++;
    if ((rc = convert_to_shoup(ctx, (uint32_t *)pl->wloc.p, (const uint32_t *)wloc_m.p, half, st))) return rc;
    if (P > 1) {
        // which boundaries get a flat table (one product per element) and which the two-level tables (two products)
        bool any_two_level = false, any_flat = false;
        for (size_t q = 1; q < P; ++q) {
            uint32_t logN = 0;
            for (size_t i = q - 1; i < P; ++i) logN += pl->digits[i];
            pl->pre_two_level[q] = ((size_t)192 << logN) > ctx->flat_table_limit;
            (pl->pre_two_level[q] ? any_two_level : any_flat) = true;
        }
        if (any_two_level) {
            pl->lo_bits = (logn + 1) / 2;
            if ((rc = build_two_level(ctx, w_eff, scale ? n_inv : nullptr, pl->lo_bits, logn - pl->lo_bits, pl->t_lo, &pl->t_lo_scaled, pl->t_hi, st))) return rc;
            pl->table_bytes += pl->t_lo.bytes + pl->t_lo_scaled.bytes + pl->t_hi.bytes;
        }
        if (any_flat) {
            const uint32_t lo_bits = std::min<uint32_t>(10, logn);
            if ((rc = dev_alloc(t_lo, (1ull << lo_bits) * 96))) return rc;
            if ((rc = dev_alloc(t_hi, (n >> lo_bits) * 96))) return rc;
            gsn::pow_table768<<<(unsigned)(((1ull << lo_bits) + 127) / 128), 128, 0, st>>>(ctx->fc, (uint32_t *)t_lo.p, (const uint32_t *)d_w.p, 1ull << lo_bits, 1);
            gsn::pow_table768<<<(unsigned)(((n >> lo_bits) + 127) / 128), 128, 0, st>>>(ctx->fc, (uint32_t *)t_hi.p, (const uint32_t *)d_w.p, n >> lo_bits, 1ull << lo_bits);
            ctx->launches += 2;
            // boundaries are built last-to-first so that, for an inverse plan, the low table can be
            // scaled by n^-1 in place just before boundary 1 (which thereby carries the scaling)
            for (size_t q = P - 1; q >= 1; --q) {
                if (pl->pre_two_level[q]) continue;
                uint32_t logN = 0;
                for (size_t i = q - 1; i < P; ++i) logN += pl->digits[i];
                const uint32_t rest_bits = logN - pl->digits[q - 1];
                if (q == 1 && scale) {
                    const uint64_t cnt = 1ull << lo_bits;
                    gsn::scale_table768<<<(unsigned)((cnt + 127) / 128), 128, 0, st>>>(ctx->fc, (uint32_t *)t_lo.p, (const uint32_t *)t_lo.p, (const uint32_t *)d_ninv.p, cnt);
                    ctx->launches++;
                }
                pl->pre[q] = std::make_unique<DevBuf>();
                DevBuf pre_m;  // Montgomery form (transient), converted into the plan's table
                if ((rc = dev_alloc(pre_m, (1ull << logN) * 96)) || (rc = dev_alloc(*pl->pre[q], (1ull << logN) * 192))) return rc;
                pl->table_bytes += pl->pre[q]->bytes;
                pl->pre_mask[q] = (1ull << logN) - 1;
                const uint64_t cnt = 1ull << logN;
                gsn::build_pretw768<<<(unsigned)((cnt + 127) / 128), 128, 0, st>>>(ctx->fc, (uint32_t *)pre_m.p, (const uint32_t *)t_lo.p, (const uint32_t *)t_hi.p,
                                                                                    logN, rest_bits, logn - logN, lo_bits);
                ctx->launches++;
                if ((rc = convert_to_shoup(ctx, (uint32_t *)pl->pre[q]->p, (const uint32_t *)pre_m.p, cnt, st))) return rc;
                CU(cudaStreamSynchronize(st));  // pre_m is freed at the end of this iteration
            }
        }
    } else if (scale) {
        pl->pre[0] = std::make_unique<DevBuf>();
        if ((rc = dev_alloc(*pl->pre[0], 192))) return rc;
        pl->table_bytes += 192;
        if ((rc = convert_to_shoup(ctx, (uint32_t *)pl->pre[0]->p, (const uint32_t *)d_ninv.p, 1, st))) return rc;
        pl->pre_mask[0] = 0;
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    pl->last_use = ++ctx->use_clock;
    *out = pl.get();
    ctx->plans768.push_back(std::move(pl));
    evict_plans768(ctx, *out);
    return GSN_OK;
}

// Tile size of the 768-bit pass kernel for `total` elements: 2^10 (or the 2-adic valuation of total if smaller), but
// small transforms prefer more, smaller tiles (never below the longest digit) until the grid covers every CTA slot
// of the device -- 2^16 is 64 tiles of 1024 but 256 tiles of 256.
uint32_t choose_log_tile768(const gsn_ctx *ctx, const Plan768 *pl, uint64_t total) {
    uint32_t log_tile = 0;
    while (log_tile < 10 && !((total >> log_tile) & 1)) ++log_tile;
    while (log_tile > pl->lmax && (total >> log_tile) < 2ull * (uint64_t)ctx->sm_count) --log_tile;
    return log_tile;
}

// Pre-twiddle supplied by the caller of a transform (first pass): a flat table indexed by the element index, or
// two-level tables of w^(k * r) with (k, r) cut out of the element index (four-step twiddles, coset shifts).
struct ExtPre {
    gsn::PreDesc pd;
    bool present = false;
};

ExtPre ext_flat(const uint32_t *table) {
    ExtPre e;
    memset(&e.pd, 0, sizeof(e.pd));
    if (table) {
        e.present = true;
        e.pd.mode = 1;
        e.pd.tab = table;
        e.pd.flat_mask = ~0ull;
    }
    return e;
}

struct WaitDesc {  // arrival flags of the fused four-step (see PassGeom::wait_flags)
    const uint32_t *flags = nullptr;
    uint32_t epoch = 0, shift = 0, mask = 0, first = 0;
};

template <int FLAGS>
int launch_pass2(gsn_ctx *ctx, unsigned grid, size_t smem, cudaStream_t st, const uint32_t *src, uint32_t *dst, const uint32_t *wloc, const gsn::PassGeom &g,
                 const gsn::PreDesc &pd, const gsn::PreDesc &post, const gsn::ScatterDesc &sc) {
    auto kern = gsn::ntt768_pass2<FLAGS>;
    if (!ctx->smem_configured.count((const void *)kern)) {
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << MAX_PASS_LOG) * gsn::SMEM_PITCH4 * 16));
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        ctx->smem_configured.insert((const void *)kern);
    }
    kern<<<grid, 256, smem, st>>>(src, dst, wloc, g, pd, post, sc, ctx->fc);
    return GSN_OK;
}

// Launches passes [q_begin, q_end) of the plan; tile range [tile0, tile0 + ntiles) of each (ntiles == 0: all).
int launch_ntt768_range(gsn_ctx *ctx, Plan768 *pl, uint32_t *d_data, size_t batch, uint32_t log_r, const ExtPre &ext, cudaStream_t st,
                        size_t q_begin, size_t q_end, uint64_t tile0, uint64_t ntiles, const gsn::ScatterDesc *scatter = nullptr,
                        const WaitDesc *wait = nullptr, const ExtPre *ext_post = nullptr) {
    const size_t P = pl->digits.size();
    const uint64_t total = (uint64_t)batch << (pl->logn + log_r);
    const uint32_t log_tile = choose_log_tile768(ctx, pl, total);
    int rc;
    uint32_t *work = nullptr;
    if (P > 1 && (rc = ensure_work(ctx, st, total * 96, &work))) return rc;
    gsn::ScatterDesc no_scatter;
    memset(&no_scatter, 0, sizeof(no_scatter));

    // kernel variant for 1024-element tiles: 1 = warp-owned tiles with wide lazy ranges, 5 = the same with stages 3-4
    // enumerated CTA-wide (unit-twiddle skips), 4 (or -1) = CTA-wide kernel.  Smaller tiles always use the CTA-wide kernel.
    // (Measured on B200 at 2^20: CTA-wide 1.280 ms, warp-owned lazy 1.297-1.313 ms, with CTA-wide stages 3-4 1.281 ms; a
    // register-prefetch form 1.337 ms, a strict-range warp-owned form 1.408 ms and a lazy CTA-wide kernel 1.308 ms were
    // built, measured and removed -- profiles/variants_r02.jsonl.)
    const int variant = ctx->v2_flags < 0 ? 4 : ctx->v2_flags;
    auto kern = gsn::ntt768_pass<NTT768_THREADS, 2>;
    if (!ctx->smem_configured.count((const void *)kern)) {
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << MAX_PASS_LOG) * gsn::SMEM_PITCH4 * 16));
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        ctx->smem_configured.insert((const void *)kern);
    }
    uint32_t below = log_r;
    for (size_t i = 0; i < P; ++i) below += pl->digits[i];
    for (size_t q = 0; q < P; ++q) {
        below -= pl->digits[q];
        if (q < q_begin || q >= q_end) continue;
        gsn::PassGeom g;
        memset(&g, 0, sizeof(g));
        g.log_l = pl->digits[q];
        g.log_s = below;
        g.log_r = log_r;
        g.tile0 = (uint32_t)tile0;
        g.log_tile = log_tile;
        g.wloc_shift = pl->lmax - pl->digits[q];
        g.final_natural = q + 1 == P;
        g.canonical = q + 1 == P;
        g.logn = pl->logn;
        if (g.final_natural) {  // output digits in reversed significance: (k_P .. k_1)
            uint32_t shift = 0, pos = pl->logn - pl->digits[P - 1];
            for (size_t i = 0; i + 1 < P; ++i) {
                pos -= pl->digits[i];
                g.dpos[i] = pos;
                g.dmask[i] = (1u << pl->digits[i]) - 1;
                g.dshift[i] = shift;
                shift += pl->digits[i];
            }
            g.kshift = shift;
        }
        if (wait && q == 0 && wait->flags) {
            g.wait_flags = wait->flags;
            g.wait_epoch = wait->epoch;
            g.wait_shift = wait->shift;
            g.wait_mask = wait->mask;
            g.wait_first = wait->first;
        }
        gsn::PreDesc pd;
        memset(&pd, 0, sizeof(pd));
        if (q == 0 && ext.present) {
            pd = ext.pd;  // caller-supplied pre-twiddle (four-step twiddles, coset shifts)
        } else if (pl->pre[q]) {
            pd.mode = 1;
            pd.tab = (const uint32_t *)pl->pre[q]->p;
            pd.flat_shift = log_r;
            pd.flat_mask = pl->pre_mask[q];
        } else if (q > 0 && pl->pre_two_level[q]) {
            // boundary q-1 | q of the sub-problem of size N = 2^(l_{q-1} + ... + l_P): w_N^(k * rest)
            uint32_t logN = 0;
            for (size_t i = q - 1; i < P; ++i) logN += pl->digits[i];
            const uint32_t rest_bits = logN - pl->digits[q - 1];
            pd.mode = 2;
            pd.tab = (const uint32_t *)((q == 1 && pl->scale) ? pl->t_lo_scaled.p : pl->t_lo.p);
            pd.tab_hi = (const uint32_t *)pl->t_hi.p;
            pd.k_shift = log_r + rest_bits;
            pd.k_mask = (1ull << pl->digits[q - 1]) - 1;
            pd.r_shift = log_r;
            pd.r_mask = (1ull << rest_bits) - 1;
            pd.logN = logN;
            pd.exp_shift = pl->logn - logN;
            pd.lo_bits = pl->lo_bits;
        }
        gsn::PreDesc post;
        memset(&post, 0, sizeof(post));
        if (q + 1 == P && ext_post && ext_post->present) post = ext_post->pd;  // caller-supplied post-twiddle, indexed by the output index
        // pass 1 reads the caller's buffer, the last pass writes it; middle passes run in
        // place in the workspace (a tile reads and writes the same index set).
        const uint32_t *src = q == 0 ? d_data : work;
        uint32_t *dst = (q + 1 == P) ? d_data : work;
        const size_t smem = ((size_t)1 << log_tile) * gsn::SMEM_PITCH4 * 16;
        const unsigned grid = (unsigned)(ntiles ? ntiles : (total >> log_tile));
        const gsn::ScatterDesc &sc = (scatter && q + 1 == P) ? *scatter : no_scatter;
        const uint32_t *wloc = (const uint32_t *)pl->wloc.p;
        if (log_tile == 10 && g.log_l >= 1 && (variant != 4 || g.wait_flags)) {   // a pass that waits on arrival flags needs the warp-owned kernel
            if (variant == 1) rc = launch_pass2<1>(ctx, grid, smem, st, src, dst, wloc, g, pd, post, sc);
            else rc = launch_pass2<5>(ctx, grid, smem, st, src, dst, wloc, g, pd, post, sc);   // variant 5, and every pass that waits on arrival flags
            if (rc) return rc;
        } else {
            if (g.wait_flags) return fail(GSN_ERR_INVALID_ARG, "arrival flags need the large-tile kernel");
            kern<<<grid, NTT768_THREADS, smem, st>>>(src, dst, wloc, g, pd, post, sc, ctx->fc);
        }
        ctx->launches++;
    }
    CU(cudaGetLastError());
    return GSN_OK;
}

int launch_ntt768(gsn_ctx *ctx, Plan768 *pl, uint32_t *d_data, size_t batch, uint32_t log_r, const uint32_t *ext_pre, cudaStream_t st) {
    return launch_ntt768_range(ctx, pl, d_data, batch, log_r, ext_flat(ext_pre), st, 0, pl->digits.size(), 0, 0);
}

int check_n(size_t n, size_t batch) {
    if (!is_pow2(n)) return fail(GSN_ERR_NOT_POW2, "n = %zu is not a power of two", n);
    if (batch == 0) return fail(GSN_ERR_INVALID_ARG, "batch = 0");
    return GSN_OK;
}

}  // namespace

// host-side staging helpers of the pageable paths (defined with the host-pointer entry points below)
extern "C" {
static bool is_pageable(const void *p);
static HostPool &host_pool(gsn_ctx *ctx);
static int ensure_bounce(gsn_ctx *ctx, size_t bytes);
}

#include "ntt32_host.inl"
#include "fourstep_host.inl"

gsn_ctx::gsn_ctx() = default;
gsn_ctx::~gsn_ctx() = default;

extern "C" {

const char *gsn_last_error(void) { return g_err.c_str(); }

int gsn_device_count(int *count) {
    if (!count) return fail(GSN_ERR_INVALID_ARG, "null count");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) { *count = 0; cudaGetLastError(); return fail(GSN_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    return GSN_OK;
}

int gsn_ctx_create(gsn_ctx **out, int device) {
    if (!out) return fail(GSN_ERR_INVALID_ARG, "null ctx pointer");
    *out = nullptr;
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt == 0) {
        cudaGetLastError();
        return fail(GSN_ERR_NO_DEVICE, "no CUDA device (%s); this library has no CPU fallback", e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= cnt) return fail(GSN_ERR_INVALID_ARG, "device %d out of range (%d devices)", device, cnt);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(GSN_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    std::unique_ptr<gsn_ctx> ctx(new gsn_ctx());
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&ctx->ev0));
    CU(cudaEventCreate(&ctx->ev1));
    CU(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
    for (int d = 0; d < 2; ++d)
        for (int i = 0; i < 16; ++i) CU(cudaEventCreateWithFlags(&ctx->ev_chunk[d][i], cudaEventDisableTiming));
    for (int d = 0; d < 2; ++d) CU(cudaEventCreateWithFlags(&ctx->ev_io_free[d], cudaEventDisableTiming));
    for (int d = 0; d < 2; ++d) CU(cudaEventCreateWithFlags(&ctx->ev_bounce[d], cudaEventDisableTiming));
    int rc = set_field(ctx.get(), GSN_FIELD_MNT4753_FR);
    if (rc) return rc;
    if (const char *v = getenv("GSN_NTT768_VARIANT")) { const int k = atoi(v); if (k == -1 || k == 1 || k == 5 || k == 4) ctx->v2_flags = k; }
    if (const char *v = getenv("GSN_FLAT_TABLE_LIMIT")) ctx->flat_table_limit = strtoull(v, nullptr, 10);
    *out = ctx.release();
    return GSN_OK;
}

int gsn_ctx_destroy(gsn_ctx *ctx) {
    if (!ctx) return GSN_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaDeviceSynchronize();
    ctx->plans768.clear();
    ctx->plans32.clear();
    ctx->cosets.clear();
    ctx->works.clear();
    ctx->pool.reset();
    for (int b = 0; b < 4; ++b) if (ctx->bounce[b]) cudaFreeHost(ctx->bounce[b]);
    for (int d = 0; d < 2; ++d) if (ctx->ev_bounce[d]) cudaEventDestroy(ctx->ev_bounce[d]);
    for (int d = 0; d < 2; ++d)
        for (int i = 0; i < 16; ++i) if (ctx->ev_chunk[d][i]) cudaEventDestroy(ctx->ev_chunk[d][i]);
    for (int d = 0; d < 2; ++d) if (ctx->ev_io_free[d]) cudaEventDestroy(ctx->ev_io_free[d]);
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return GSN_OK;
}

int gsn_ctx_trim(gsn_ctx *ctx) {
    if (!ctx) return fail(GSN_ERR_INVALID_ARG, "null ctx");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(cudaDeviceSynchronize());  // caller streams may still be running transforms that use the tables / scratch
    ctx->plans768.clear();
    ctx->plans32.clear();
    ctx->cosets.clear();
    ctx->works.clear();
    if (ctx->io.p) { cudaFree(ctx->io.p); ctx->io.p = nullptr; ctx->io.bytes = 0; }
    if (ctx->io2.p) { cudaFree(ctx->io2.p); ctx->io2.p = nullptr; ctx->io2.bytes = 0; }
    return GSN_OK;
}

int gsn_launch_count(gsn_ctx *ctx, uint64_t *count) {
    if (!ctx || !count) return fail(GSN_ERR_INVALID_ARG, "null argument");
    *count = ctx->launches;
    return GSN_OK;
}

int gsn_set_field768(gsn_ctx *ctx, int field) {
    if (!ctx) return fail(GSN_ERR_INVALID_ARG, "null ctx");
    if (field != GSN_FIELD_MNT4753_FR && field != GSN_FIELD_MNT4753_FQ) return fail(GSN_ERR_INVALID_ARG, "unknown field %d", field);
    std::lock_guard<std::mutex> lk(ctx->mu);
    return set_field(ctx, field);  // host-side only: the constants travel with each launch, other contexts are unaffected
}

int gsn_ctx_set_option(gsn_ctx *ctx, int option, uint64_t value) {
    if (!ctx) return fail(GSN_ERR_INVALID_ARG, "null ctx");
    std::lock_guard<std::mutex> lk(ctx->mu);
    switch (option) {
        case GSN_OPT_FLAT_TABLE_LIMIT: ctx->flat_table_limit = (size_t)value; break;
        case GSN_OPT_PLAN_CACHE_BYTES:
            ctx->plan_cache_bytes = (size_t)value;
            CU(cudaSetDevice(ctx->device));
            evict_plans768(ctx, nullptr);
            break;
        case GSN_OPT_KERNEL_VARIANT:
            if ((int64_t)value != -1 && value != 1 && value != 5 && value != 4) return fail(GSN_ERR_INVALID_ARG, "kernel variant %lld (1, 4, 5 or -1)", (long long)value);
            ctx->v2_flags = (int)(int64_t)value;
            break;
        default: return fail(GSN_ERR_INVALID_ARG, "unknown option %d", option);
    }
    return GSN_OK;
}

int gsn_ntt768_plan_info(gsn_ctx *ctx, size_t n, const uint32_t *omega, int inverse, uint64_t *table_bytes, unsigned *passes,
                         unsigned *two_level_boundaries, uint64_t *cached_plans, uint64_t *cached_bytes) {
    if (!ctx || !omega) return fail(GSN_ERR_INVALID_ARG, "null argument");
    int rc = check_n(n, 1);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    Plan768 *pl;
    if ((rc = get_plan768(ctx, ilog2(n), omega, inverse, inverse, &pl))) return rc;
    if (table_bytes) *table_bytes = pl->table_bytes;
    if (passes) *passes = (unsigned)pl->digits.size();
    if (two_level_boundaries) { *two_level_boundaries = 0; for (auto f : pl->pre_two_level) *two_level_boundaries += f; }
    if (cached_plans) *cached_plans = ctx->plans768.size();
    if (cached_bytes) { *cached_bytes = 0; for (auto &q : ctx->plans768) *cached_bytes += q->table_bytes; }
    return GSN_OK;
}

int gsn_ntt768_prepare(gsn_ctx *ctx, size_t n, size_t batch, const uint32_t *omega, int inverse) {
    if (!ctx || !omega) return fail(GSN_ERR_INVALID_ARG, "null argument");
    int rc = check_n(n, batch);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    Plan768 *pl;
    if ((rc = get_plan768(ctx, ilog2(n), omega, inverse, inverse, &pl))) return rc;
    uint32_t *work;
    if (pl->digits.size() > 1) return ensure_work(ctx, ctx->stream, (size_t)batch * n * 96, &work);
    return GSN_OK;
}

int gsn_ntt768_device_ex(gsn_ctx *ctx, uint32_t *d_limbs, size_t n, size_t batch, unsigned log_r, const uint32_t *omega,
                         unsigned flags, const uint32_t *d_pre_table, void *stream) {
    if (!ctx || !d_limbs || !omega) return fail(GSN_ERR_INVALID_ARG, "null argument");
    int rc = check_n(n, batch);
    if (rc) return rc;
    if (log_r > 40) return fail(GSN_ERR_INVALID_ARG, "log_r = %u", log_r);
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    Plan768 *pl;
    const int inv = (flags & GSN_FLAG_INVERSE_ROOT) != 0, scale = inv && !(flags & GSN_FLAG_NO_SCALE);
    if (d_pre_table && scale && n <= 1024)
        return fail(GSN_ERR_INVALID_ARG, "a pre-twiddle table and n^-1 scaling cannot share a one-pass transform: fold the scale into the table and pass GSN_FLAG_NO_SCALE");
    if ((rc = get_plan768(ctx, ilog2(n), omega, inv, scale, &pl))) return rc;
    return launch_ntt768(ctx, pl, d_limbs, batch, log_r, d_pre_table, stream ? (cudaStream_t)stream : ctx->stream);
}

int gsn_ntt768_strided_device(gsn_ctx *ctx, uint32_t *d_limbs, size_t n, size_t batch, unsigned log_r, const uint32_t *omega,
                              int inverse, void *stream) {
    return gsn_ntt768_device_ex(ctx, d_limbs, n, batch, log_r, omega, inverse ? GSN_FLAG_INVERSE_ROOT : 0, nullptr, stream);
}

int gsn_ntt768_device(gsn_ctx *ctx, uint32_t *d_limbs, size_t n, size_t batch, const uint32_t *omega, int inverse, void *stream) {
    return gsn_ntt768_strided_device(ctx, d_limbs, n, batch, 0, omega, inverse, stream);
}

// Enqueue one host-pointer transform on the context's three streams (copy-in, compute, copy-out) without
// waiting for it: H2D in column blocks, pass 1 per block as it lands, middle passes, last pass per block of
// output columns, D2H per block.  `io` is the device staging buffer to use; ev_free (may be null) is
// recorded on the copy-out stream when the last D2H of this transform has been issued.
static int enqueue_ntt768_host(gsn_ctx *ctx, Plan768 *pl, uint32_t *limbs, size_t n, uint32_t *io, cudaEvent_t wait_before_h2d,
                               cudaEvent_t ev_free, bool first) {
    int rc;
    const size_t P = pl->digits.size();
    cudaStream_t st = ctx->stream;
    if (wait_before_h2d) {
        CU(cudaStreamWaitEvent(ctx->s_in, wait_before_h2d, 0));
        CU(cudaStreamWaitEvent(st, wait_before_h2d, 0));
    }
    if (P < 2 || n * 96 < (8u << 20)) {  // small: copy in, transform, copy out on the compute stream
        CU(cudaMemcpyAsync(io, limbs, n * 96, cudaMemcpyHostToDevice, st));
        if ((rc = launch_ntt768(ctx, pl, io, 1, 0, nullptr, st))) return rc;
        CU(cudaMemcpyAsync(limbs, io, n * 96, cudaMemcpyDeviceToHost, st));
        if (ev_free) {
            CU(cudaEventRecord(ev_free, st));
        }
        return GSN_OK;
    }
    // The first pass works on columns of the (2^l_1 x n/2^l_1) view of the input, so the H2D copy is cut into
    // column blocks (2-D copies) and each block's tiles start as soon as it has landed; the last pass produces
    // column blocks of the (n/2^l_1 x 2^l_1) view of the output, copied back while later blocks are computed.
    const uint32_t l1 = pl->digits[0], lP = pl->digits[P - 1];
    const uint64_t cols_in = n >> l1, rows_in = 1ull << l1;     // input view: rows_in x cols_in
    const uint64_t cols_out = 1ull << l1, rows_out = n >> l1;   // output view: rows_out x cols_out (k1 fastest)
    const uint64_t tiles = n >> choose_log_tile768(ctx, pl, n);
    int chunks = 8;
    while (chunks > 1 && (cols_in % chunks || cols_out % chunks || tiles % chunks || (cols_in / chunks) * rows_in < 1024 ||
                          (cols_out / chunks) * (1ull << lP) < 1024)) chunks >>= 1;
    uint32_t *work_unused;
    if ((rc = ensure_work(ctx, st, n * 96, &work_unused))) return rc;
    if (first) {  // order the copy-in stream after earlier (asynchronous, device-pointer) work on the context; later
                  // transforms of a batch must NOT wait for the compute stream, or their H2D could not overlap it
        CU(cudaEventRecord(ctx->ev0, st));
        CU(cudaStreamWaitEvent(ctx->s_in, ctx->ev0, 0));
    }
    const uint64_t cw_in = cols_in / chunks, cw_out = cols_out / chunks, tiles_per_chunk = tiles / chunks;
    for (int c = 0; c < chunks; ++c) {
        CU(cudaMemcpy2DAsync(io + c * cw_in * 24, cols_in * 96, limbs + c * cw_in * 24, cols_in * 96, cw_in * 96, rows_in, cudaMemcpyHostToDevice, ctx->s_in));
        CU(cudaEventRecord(ctx->ev_chunk[0][c], ctx->s_in));
        CU(cudaStreamWaitEvent(st, ctx->ev_chunk[0][c], 0));
        // pass 1 tiles of this column block: tile t covers sub-transforms (= columns) [t * 2^(log_tile-l1), ...)
        if ((rc = launch_ntt768_range(ctx, pl, io, 1, 0, ext_flat(nullptr), st, 0, 1, c * tiles_per_chunk, tiles_per_chunk))) return rc;
    }
    if (P > 2 && (rc = launch_ntt768_range(ctx, pl, io, 1, 0, ext_flat(nullptr), st, 1, P - 1, 0, 0))) return rc;
    for (int c = 0; c < chunks; ++c) {
        // last pass: sub-transform index t = (k_1, k_2, ...) with k_1 most significant, so a contiguous tile
        // range is a k_1 range = a column block of the output view
        if ((rc = launch_ntt768_range(ctx, pl, io, 1, 0, ext_flat(nullptr), st, P - 1, P, c * tiles_per_chunk, tiles_per_chunk))) return rc;
        CU(cudaEventRecord(ctx->ev_chunk[1][c], st));
        CU(cudaStreamWaitEvent(ctx->s_out, ctx->ev_chunk[1][c], 0));
        CU(cudaMemcpy2DAsync(limbs + c * cw_out * 24, cols_out * 96, io + c * cw_out * 24, cols_out * 96, cw_out * 96, rows_out, cudaMemcpyDeviceToHost, ctx->s_out));
    }
    if (ev_free) {
        CU(cudaEventRecord(ev_free, ctx->s_out));
    }
    return GSN_OK;
}

// Pageable host memory (what a std::vector hands us): a DMA from it is staged by the driver through one small pinned
// buffer, at a fraction of the PCIe rate.  Here the staging is ours: two pinned bounce buffers per direction, filled /
// drained by a few host threads (memcpy at memory speed) while the DMA engine moves the other one.
static bool is_pageable(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

static HostPool &host_pool(gsn_ctx *ctx) {
    if (!ctx->pool) {
        unsigned nt = std::min(8u, std::max(1u, std::thread::hardware_concurrency() / 2));
        if (const char *v = getenv("GSN_HOST_THREADS")) nt = (unsigned)std::max(1, std::min(64, atoi(v)));
        ctx->pool.reset(new HostPool(nt));
    }
    return *ctx->pool;
}

// rows x width bytes between two pitched host buffers, rows split over the pool
static void copy_rows(gsn_ctx *ctx, void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows) {
    static const bool streaming = !(getenv("GSN_STREAM_COPY") && atoi(getenv("GSN_STREAM_COPY")) == 0);   // non-temporal stores (default on)
    if (!streaming) {
        host_pool(ctx).run([=](unsigned t, unsigned nt) {
            if (dpitch == width && spitch == width) {
                const size_t bytes = width * rows, per = ((bytes / nt) + 4095) & ~(size_t)4095, off = t * per;
                if (off < bytes) memcpy((char *)dst + off, (const char *)src + off, std::min(per, bytes - off));
                return;
            }
            for (size_t r = rows * t / nt; r < rows * (t + 1) / nt; ++r) memcpy((char *)dst + r * dpitch, (const char *)src + r * spitch, width);
        });
        return;
    }
    host_pool(ctx).run([=](unsigned t, unsigned nt) {
        if (dpitch == width && spitch == width) {   // contiguous: split by bytes, page aligned
            const size_t bytes = width * rows, per = ((bytes / nt) + 4095) & ~(size_t)4095;
            const size_t off = t * per;
            if (off < bytes) stream_copy((char *)dst + off, (const char *)src + off, std::min(per, bytes - off));
            return;
        }
        const size_t r0 = rows * t / nt, r1 = rows * (t + 1) / nt;
        for (size_t r = r0; r < r1; ++r) stream_copy((char *)dst + r * dpitch, (const char *)src + r * spitch, width);
    });
}

static int ensure_bounce(gsn_ctx *ctx, size_t bytes) {
    if (ctx->bounce_bytes >= bytes) return GSN_OK;
    for (int b = 0; b < 4; ++b) {
        if (ctx->bounce[b]) { cudaFreeHost(ctx->bounce[b]); ctx->bounce[b] = nullptr; }
    }
    ctx->bounce_bytes = 0;
    for (int b = 0; b < 4; ++b) CU(cudaHostAlloc(&ctx->bounce[b], bytes, cudaHostAllocPortable));   // the multi-GPU host entry DMAs from them on every device
    ctx->bounce_bytes = bytes;
    return GSN_OK;
}

// Pageable vector, multi-pass plan, blocks of at most 64 MB: the same column-block pipeline as enqueue_ntt768_host,
// with the host threads gathering each input block into a pinned buffer (and scattering each output block from one)
// while the DMA engines and the passes work on the neighbouring blocks.  *done stays false when the shape does not fit
// (single pass, too few or too large blocks); the caller then uses the contiguous-chunk path below.
static int ntt768_host_pageable_blocks(gsn_ctx *ctx, Plan768 *pl, uint32_t *limbs, size_t n, bool *done) {
    *done = false;
    int rc;
    const size_t P = pl->digits.size();
    if (P < 2) return GSN_OK;
    const uint32_t l1 = pl->digits[0], lP = pl->digits[P - 1];
    const uint64_t cols_in = n >> l1, rows_in = 1ull << l1, cols_out = 1ull << l1, rows_out = n >> l1;
    const uint64_t tiles = n >> choose_log_tile768(ctx, pl, n);
    int chunks = 8;
    while (chunks > 1 && (cols_in % chunks || cols_out % chunks || tiles % chunks || (cols_in / chunks) * rows_in < 1024 ||
                          (cols_out / chunks) * (1ull << lP) < 1024)) chunks >>= 1;
    const size_t blk = n * 96 / chunks;
    if (chunks < 4 || blk > ((size_t)64 << 20)) return GSN_OK;
    if ((rc = ensure_io(ctx, n * 96)) || (rc = ensure_bounce(ctx, std::max(blk, (size_t)16 << 20)))) return rc;
    cudaStream_t st = ctx->stream;
    uint32_t *work_unused;
    if ((rc = ensure_work(ctx, st, n * 96, &work_unused))) return rc;
    uint32_t *io = (uint32_t *)ctx->io.p;
    char *host = (char *)limbs;
    CU(cudaEventRecord(ctx->ev0, st));   // the copy streams start after earlier work on the context
    CU(cudaStreamWaitEvent(ctx->s_in, ctx->ev0, 0));
    CU(cudaStreamWaitEvent(ctx->s_out, ctx->ev0, 0));
    const uint64_t cw_in = cols_in / chunks, cw_out = cols_out / chunks, tiles_per_chunk = tiles / chunks;
    for (int c = 0; c < chunks; ++c) {
        const int b = c & 1;
        if (c >= 2) CU(cudaEventSynchronize(ctx->ev_chunk[0][c - 2]));   // the DMA that last read this bounce buffer is done
        copy_rows(ctx, ctx->bounce[b], cw_in * 96, host + c * cw_in * 96, cols_in * 96, cw_in * 96, rows_in);
        CU(cudaMemcpy2DAsync(io + c * cw_in * 24, cols_in * 96, ctx->bounce[b], cw_in * 96, cw_in * 96, rows_in, cudaMemcpyHostToDevice, ctx->s_in));
        CU(cudaEventRecord(ctx->ev_chunk[0][c], ctx->s_in));
        CU(cudaStreamWaitEvent(st, ctx->ev_chunk[0][c], 0));
        if ((rc = launch_ntt768_range(ctx, pl, io, 1, 0, ext_flat(nullptr), st, 0, 1, c * tiles_per_chunk, tiles_per_chunk))) return rc;
    }
    if (P > 2 && (rc = launch_ntt768_range(ctx, pl, io, 1, 0, ext_flat(nullptr), st, 1, P - 1, 0, 0))) return rc;
    for (int c = 0; c < chunks; ++c) {
        if ((rc = launch_ntt768_range(ctx, pl, io, 1, 0, ext_flat(nullptr), st, P - 1, P, c * tiles_per_chunk, tiles_per_chunk))) return rc;
        CU(cudaEventRecord(ctx->ev_chunk[1][c], st));
    }
    auto copy_out = [&](int c) -> int {
        CU(cudaStreamWaitEvent(ctx->s_out, ctx->ev_chunk[1][c], 0));
        CU(cudaMemcpy2DAsync(ctx->bounce[2 + (c & 1)], cw_out * 96, io + c * cw_out * 24, cols_out * 96, cw_out * 96, rows_out, cudaMemcpyDeviceToHost, ctx->s_out));
        CU(cudaEventRecord(ctx->ev_bounce[c & 1], ctx->s_out));
        return GSN_OK;
    };
    for (int c = 0; c < std::min(2, chunks); ++c)
        if ((rc = copy_out(c))) return rc;
    for (int c = 0; c < chunks; ++c) {
        CU(cudaEventSynchronize(ctx->ev_bounce[c & 1]));
        copy_rows(ctx, host + c * cw_out * 96, cols_out * 96, ctx->bounce[2 + (c & 1)], cw_out * 96, cw_out * 96, rows_out);
        if (c + 2 < chunks && (rc = copy_out(c + 2))) return rc;
    }
    CU(cudaStreamSynchronize(ctx->s_out));
    CU(cudaStreamSynchronize(st));
    CU(cudaStreamSynchronize(ctx->s_in));
    *done = true;
    return GSN_OK;
}

// host -> device copy of a possibly pageable buffer: a large pageable source goes through the pinned bounce buffers
// (filled by the host threads while the DMA engine moves the previous chunk); returns with the copy complete
static int h2d_staged(gsn_ctx *ctx, void *dst, const void *src, size_t bytes, cudaStream_t st) {
    if (bytes < ((size_t)4 << 20) || !is_pageable(src)) {
        CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));
        return GSN_OK;
    }
    int rc;
    const size_t chunk = (size_t)16 << 20;
    if ((rc = ensure_bounce(ctx, chunk))) return rc;
    size_t k = 0;
    for (size_t off = 0; off < bytes; off += chunk, ++k) {
        const size_t len = std::min(chunk, bytes - off);
        const int b = (int)(k & 1);
        if (k >= 2) CU(cudaEventSynchronize(ctx->ev_chunk[0][b]));
        copy_rows(ctx, ctx->bounce[b], len, (const char *)src + off, len, len, 1);
        CU(cudaMemcpyAsync((char *)dst + off, ctx->bounce[b], len, cudaMemcpyHostToDevice, st));
        CU(cudaEventRecord(ctx->ev_chunk[0][b], st));
    }
    CU(cudaStreamSynchronize(st));
    return GSN_OK;
}

static int ntt768_host_pageable(gsn_ctx *ctx, Plan768 *pl, uint32_t *limbs, size_t n) {
    int rc;
    bool done = false;
    if ((rc = ntt768_host_pageable_blocks(ctx, pl, limbs, n, &done)) || done) return rc;
    const size_t bytes = n * 96, chunk = (size_t)16 << 20;
    if ((rc = ensure_io(ctx, bytes)) || (rc = ensure_bounce(ctx, chunk))) return rc;
    char *dev = (char *)ctx->io.p, *host = (char *)limbs;
    cudaStream_t st = ctx->stream;
    size_t k = 0;
    for (size_t off = 0; off < bytes; off += chunk, ++k) {
        const size_t len = std::min(chunk, bytes - off);
        const int b = (int)(k & 1);
        if (k >= 2) CU(cudaEventSynchronize(ctx->ev_chunk[0][b]));   // the DMA that last read this bounce buffer is done
        copy_rows(ctx, ctx->bounce[b], len, host + off, len, len, 1);
        CU(cudaMemcpyAsync(dev + off, ctx->bounce[b], len, cudaMemcpyHostToDevice, st));
        CU(cudaEventRecord(ctx->ev_chunk[0][b], st));
    }
    if ((rc = launch_ntt768(ctx, pl, (uint32_t *)dev, 1, 0, nullptr, st))) return rc;
    // copy out: DMA of chunk k+1 overlaps the host memcpy of chunk k
    const size_t nchunks = (bytes + chunk - 1) / chunk;
    for (size_t c = 0; c < std::min<size_t>(2, nchunks); ++c) {
        CU(cudaMemcpyAsync(ctx->bounce[c & 1], dev + c * chunk, std::min(chunk, bytes - c * chunk), cudaMemcpyDeviceToHost, st));
        CU(cudaEventRecord(ctx->ev_chunk[1][c & 1], st));
    }
    for (size_t c = 0; c < nchunks; ++c) {
        const int b = (int)(c & 1);
        const size_t len = std::min(chunk, bytes - c * chunk);
        CU(cudaEventSynchronize(ctx->ev_chunk[1][b]));
        copy_rows(ctx, host + c * chunk, len, ctx->bounce[b], len, len, 1);
        if (c + 2 < nchunks) {
            CU(cudaMemcpyAsync(ctx->bounce[b], dev + (c + 2) * chunk, std::min(chunk, bytes - (c + 2) * chunk), cudaMemcpyDeviceToHost, st));
            CU(cudaEventRecord(ctx->ev_chunk[1][b], st));
        }
    }
    CU(cudaStreamSynchronize(st));
    return GSN_OK;
}

int gsn_ntt768_host_batch(gsn_ctx *ctx, uint32_t *const *limbs, size_t count, size_t n, const uint32_t *omega, int inverse) {
    if (!ctx || !limbs || !omega) return fail(GSN_ERR_INVALID_ARG, "null argument");
    for (size_t i = 0; i < count; ++i)
        if (!limbs[i]) return fail(GSN_ERR_INVALID_ARG, "limbs[%zu] is null", i);
    int rc = check_n(n, 1);
    if (rc) return rc;
    if (count == 0) return GSN_OK;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    Plan768 *pl;
    if ((rc = get_plan768(ctx, ilog2(n), omega, inverse, inverse, &pl))) return rc;
    if (n * 96 >= ((size_t)4 << 20)) {   // large pageable vectors go through our own pinned bounce buffers
        bool pageable = false;
        for (size_t i = 0; i < count; ++i) pageable = pageable || is_pageable(limbs[i]);
        if (pageable) {
            for (size_t i = 0; i < count; ++i)
                if ((rc = ntt768_host_pageable(ctx, pl, limbs[i], n))) return rc;
            return GSN_OK;
        }
    }
    if ((rc = ensure_io(ctx, n * 96))) return rc;
    if (count > 1) {
        if (ctx->io2.bytes < n * 96) {
            if (ctx->io2.p) { cudaFree(ctx->io2.p); ctx->io2.p = nullptr; ctx->io2.bytes = 0; }
            if ((rc = dev_alloc(ctx->io2, n * 96))) return rc;
        }
    }
    // Transform i uses staging buffer i % 2; its copy-in waits until the copy-out of transform i - 2 is done.
    // The three streams are in order, so H2D of i+1 overlaps the passes and the D2H of i (full-duplex PCIe).
    for (size_t i = 0; i < count; ++i) {
        uint32_t *io = (uint32_t *)((i & 1) ? ctx->io2.p : ctx->io.p);
        if ((rc = enqueue_ntt768_host(ctx, pl, limbs[i], n, io, i >= 2 ? ctx->ev_io_free[i & 1] : nullptr, ctx->ev_io_free[i & 1], i == 0))) {
            cudaStreamSynchronize(ctx->s_out);
            cudaStreamSynchronize(ctx->stream);
            return rc;
        }
    }
    CU(cudaStreamSynchronize(ctx->s_out));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaStreamSynchronize(ctx->s_in));
    return GSN_OK;
}

int gsn_ntt768_host(gsn_ctx *ctx, uint32_t *limbs, size_t n, const uint32_t *omega, int inverse) {
    uint32_t *one[1] = {limbs};
    if (!limbs) return fail(GSN_ERR_INVALID_ARG, "null argument");
    return gsn_ntt768_host_batch(ctx, one, 1, n, omega, inverse);
}

int gsn_ntt768_time_device(gsn_ctx *ctx, uint32_t *d_limbs, size_t n, size_t batch, const uint32_t *omega, int inverse, int reps,
                           float *ms_each) {
    if (!ctx || !d_limbs || !omega || !ms_each || reps <= 0) return fail(GSN_ERR_INVALID_ARG, "bad argument");
    int rc = gsn_ntt768_prepare(ctx, n, batch, omega, inverse);
    if (rc) return rc;
    for (int i = 0; i < reps; ++i) {
        CU(cudaEventRecord(ctx->ev0, ctx->stream));
        if ((rc = gsn_ntt768_device(ctx, d_limbs, n, batch, omega, inverse, nullptr))) return rc;
        CU(cudaEventRecord(ctx->ev1, ctx->stream));
        CU(cudaEventSynchronize(ctx->ev1));
        CU(cudaEventElapsedTime(&ms_each[i], ctx->ev0, ctx->ev1));
    }
    return GSN_OK;
}

int gsn_fp768_binop_device(gsn_ctx *ctx, int op, uint32_t *d_out, const uint32_t *d_a, const uint32_t *d_b, size_t count, void *stream) {
    if (!ctx || !d_out || !d_a || !d_b) return fail(GSN_ERR_INVALID_ARG, "null argument");
    if (op < 0 || op > 2) return fail(GSN_ERR_INVALID_ARG, "op %d", op);
    if (count == 0) return GSN_OK;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    gsn::binop768<<<(unsigned)((count + 127) / 128), 128, 0, stream ? (cudaStream_t)stream : ctx->stream>>>(ctx->fc, d_out, d_a, d_b, count, op);
    ctx->launches++;
    CU(cudaGetLastError());
    return GSN_OK;
}

int gsn_fp768_powers_device(gsn_ctx *ctx, uint32_t *d_table, size_t count, const uint32_t *base, const uint32_t *scale, void *stream) {
    if (!ctx || !d_table || !base) return fail(GSN_ERR_INVALID_ARG, "null argument");
    if (count == 0) return GSN_OK;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    gsn::Elem768 b, sc;   // by value in the launch parameters: stream ordered, nothing to allocate or wait for
    memcpy(b.v, base, 96);
    memcpy(sc.v, scale ? scale : ctx->fc.r1, 96);
    gsn::powers768<<<(unsigned)((count + 127) / 128), 128, 0, st>>>(ctx->fc, d_table, b, sc, count);
    ctx->launches++;
    CU(cudaGetLastError());
    return GSN_OK;
}

int gsn_fp768_twiddle_table_device(gsn_ctx *ctx, uint32_t *d_table, const uint32_t *d_elems, size_t count, void *stream) {
    if (!ctx || !d_table || !d_elems) return fail(GSN_ERR_INVALID_ARG, "null argument");
    if (count == 0) return GSN_OK;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    return convert_to_shoup(ctx, d_table, d_elems, count, stream ? (cudaStream_t)stream : ctx->stream);
}

int gsn_fp768_inner_product_device(gsn_ctx *ctx, uint32_t *d_out, const uint32_t *d_a, const uint32_t *d_b, size_t count, void *stream) {
    if (!ctx || !d_out || !d_a || !d_b) return fail(GSN_ERR_INVALID_ARG, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    const unsigned blocks = (unsigned)std::min<size_t>((count + 127) / 128, (size_t)ctx->sm_count * 8);
    int rc;
    uint32_t *work;
    if ((rc = ensure_work(ctx, st, (size_t)std::max(1u, blocks) * 96, &work))) return rc;
    if (count == 0) { CU(cudaMemsetAsync(d_out, 0, 96, st)); return GSN_OK; }
    gsn::inner_product768<128><<<blocks, 128, 0, st>>>(ctx->fc, work, d_a, d_b, count);
    gsn::inner_product768<128><<<1, 128, 0, st>>>(ctx->fc, d_out, work, nullptr, blocks);
    ctx->launches += 2;
    CU(cudaGetLastError());
    return GSN_OK;
}

int gsn_fp768_inner_product_host(gsn_ctx *ctx, uint32_t *out, const uint32_t *a, const uint32_t *b, size_t count) {
    if (!ctx || !out || !a || !b) return fail(GSN_ERR_INVALID_ARG, "null argument");
    DevBuf da, db, dc;
    int rc;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        CU(cudaSetDevice(ctx->device));
        if ((rc = dev_alloc(da, std::max<size_t>(count, 1) * 96)) || (rc = dev_alloc(db, std::max<size_t>(count, 1) * 96)) || (rc = dev_alloc(dc, 96))) return rc;
        CU(cudaMemcpyAsync(da.p, a, count * 96, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(db.p, b, count * 96, cudaMemcpyHostToDevice, ctx->stream));
    }
    if ((rc = gsn_fp768_inner_product_device(ctx, (uint32_t *)dc.p, (const uint32_t *)da.p, (const uint32_t *)db.p, count, nullptr))) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaMemcpyAsync(out, dc.p, 96, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return GSN_OK;
}

// the reference's own algorithm: one double-and-add per point, then a tree reduction (tiny inputs, and the A/B of the tests)
static int g1_multiexp_naive(gsn_ctx *ctx, uint32_t *d_out, const uint32_t *d_points, const uint32_t *d_scalars, size_t n, cudaStream_t st) {
    int rc;
    uint32_t *work;
    if ((rc = ensure_work(ctx, st, std::max<size_t>(n, 1) * 288, &work))) return rc;
    constexpr int RT = 128;
    auto red = gsn::g1_reduce_kernel<RT>;
    if (!ctx->smem_configured.count((const void *)red)) {
        CU(cudaFuncSetAttribute(red, cudaFuncAttributeMaxDynamicSharedMemorySize, RT * 288));
        ctx->smem_configured.insert((const void *)red);
    }
    if (n) {
        gsn::g1_scalar_mul_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(work, d_points, d_scalars, n);
        ctx->launches++;
    }
    red<<<1, RT, RT * 288, st>>>(d_out, work, n);
    ctx->launches++;
    CU(cudaGetLastError());
    return GSN_OK;
}

// bucket method: signed c-bit windows -> (bucket, point) pairs sorted by bucket (CUB radix sort) -> one thread per
// bucket -> per-window running-sum reduction -> Horner over the window sums on the host (g1_host.h).  Blocking.
static int g1_multiexp_pippenger(gsn_ctx *ctx, uint32_t *d_out, const uint32_t *d_points, const uint32_t *d_scalars, size_t n, cudaStream_t st,
                                 unsigned c_override) {
    int rc;
    uint32_t c = c_override ? c_override : (uint32_t)std::min<int>(16, std::max<int>(2, (int)ilog2(n) - 3));
    const uint32_t windows = (768 + 1 + c - 1) / c, nb = 1u << (c - 1), bs = nb + 1;
    const uint64_t pairs = (uint64_t)n * windows, nbuckets = (uint64_t)windows * bs;
    if (pairs >= (1ull << 31)) return fail(GSN_ERR_TOO_LARGE, "multiexp of %zu points: %llu (point, window) pairs", n, (unsigned long long)pairs);
    // one arena from the stream's scratch buffer (no cudaMalloc / cudaFree per call)
    int key_bits = 1;
    while ((1ull << key_bits) < nbuckets) ++key_bits;
    size_t tmp_bytes = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                       (int)pairs, 0, key_bits, st));
    // a bucket goes to the block-per-item kernels when it holds more than `limit` points: four times the average bucket
    // (n / 2^(c-1) points; the window width is capped at 16 bits, so the average itself passes 128 from 2^22 points), at
    // least 128 and at most 1024 -- no single thread walks more than that, however few buckets a narrow window leaves
    const uint32_t limit = (uint32_t)std::min<uint64_t>(1024, std::max<uint64_t>(128, 4 * ((uint64_t)n >> (c - 1))));
    const uint32_t wparts = nb >= 8192 ? 4 : nb >= 2048 ? 2 : 1;   // blocks per window of the running-sum reduction
    // bounds of the heavy work list: every item but the split ones covers > limit points; a split bucket of s points has at most 2 s / 4096 parts
    const size_t split_cap = pairs / gsn::G1_SPLIT_POINTS + 2, slot_cap = 2 * split_cap + gsn::G1_MAX_PARTS, item_cap = pairs / limit + slot_cap + 2;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_keys = 0, o_vals = o_keys + up(pairs * 4), o_keys2 = o_vals + up(pairs * 4), o_vals2 = o_keys2 + up(pairs * 4),
                 o_tmp = o_vals2 + up(pairs * 4), o_cnt = o_tmp + up(std::max<size_t>(tmp_bytes, 16)), o_items = o_cnt + 256,
                 o_splits = o_items + up(item_cap * sizeof(gsn::G1HeavyItem)), o_partial = o_splits + up(split_cap * sizeof(gsn::G1HeavySplit)),
                 o_buckets = o_partial + up(slot_cap * 288), o_wsum = o_buckets + up(nbuckets * 288), total_bytes = o_wsum + up((size_t)windows * wparts * 288);
    uint32_t *arena_w;
    if ((rc = ensure_work(ctx, st, total_bytes, &arena_w))) return rc;
    char *arena = (char *)arena_w;
    uint32_t *keys = (uint32_t *)(arena + o_keys), *vals = (uint32_t *)(arena + o_vals), *keys2 = (uint32_t *)(arena + o_keys2), *vals2 = (uint32_t *)(arena + o_vals2);
    uint32_t *buckets = (uint32_t *)(arena + o_buckets), *wsum = (uint32_t *)(arena + o_wsum);
    gsn::G1HeavyLists hl{(uint32_t *)(arena + o_cnt), (gsn::G1HeavyItem *)(arena + o_items), (gsn::G1HeavySplit *)(arena + o_splits),
                         (uint32_t *)(arena + o_partial)};
    CU(cudaMemsetAsync(hl.counters, 0, 16, st));
    gsn::g1_digits_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(keys, vals, d_scalars, n, c, windows, bs);
    CU(cub::DeviceRadixSort::SortPairs(arena + o_tmp, tmp_bytes, (const uint32_t *)keys, keys2, (const uint32_t *)vals, vals2, (int)pairs, 0, key_bits, st));
    gsn::g1_bucket_kernel<<<(unsigned)((nbuckets + 127) / 128), 128, 0, st>>>(buckets, d_points, keys2, vals2, pairs, (uint32_t)nbuckets, bs, limit, hl);
    constexpr int HT = 128, WT = 256;
    auto hv = gsn::g1_heavy_bucket_kernel<HT>;
    auto cmb = gsn::g1_heavy_combine_kernel<WT>;
    auto red = gsn::g1_window_reduce_kernel<WT>;
    if (!ctx->smem_configured.count((const void *)red)) {
        CU(cudaFuncSetAttribute(red, cudaFuncAttributeMaxDynamicSharedMemorySize, WT * 288));
        CU(cudaFuncSetAttribute(cmb, cudaFuncAttributeMaxDynamicSharedMemorySize, WT * 288));
        CU(cudaFuncSetAttribute(hv, cudaFuncAttributeMaxDynamicSharedMemorySize, HT * 288));
        ctx->smem_configured.insert((const void *)red);
    }
    hv<<<(unsigned)std::min<size_t>(item_cap, (size_t)ctx->sm_count * 4), HT, HT * 288, st>>>(buckets, d_points, keys2, vals2, pairs, hl);
    cmb<<<(unsigned)std::min<size_t>(split_cap, (size_t)ctx->sm_count), WT, WT * 288, st>>>(buckets, hl);
    red<<<dim3(windows, wparts), WT, WT * 288, st>>>(wsum, buckets, c, bs);
    ctx->launches += 5;
    CU(cudaGetLastError());
    std::vector<gsn::host::G1Host> S((size_t)windows * wparts);
    static_assert(sizeof(gsn::host::G1Host) == 288, "three 96-byte coordinates");
    CU(cudaMemcpyAsync(S.data(), wsum, (size_t)windows * wparts * 288, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    // Horner over the windows, most significant first: acc = 2^c acc + S_w
    gsn::host::G1Ops ops(gsn::host::fq_field());
    gsn::host::G1Host acc;
    ops.identity(acc);
    for (uint32_t w = windows; w-- > 0;) {
        if (!ops.is_identity(acc))
            for (uint32_t k = 0; k < c; ++k) ops.dbl(acc, acc);
        for (uint32_t j = 0; j < wparts; ++j) ops.add(acc, acc, S[(size_t)w * wparts + j]);
    }
    CU(cudaMemcpyAsync(d_out, &acc, 288, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    return GSN_OK;
}

int gsn_g1_multiexp_device_ex(gsn_ctx *ctx, uint32_t *d_out, const uint32_t *d_points, const uint32_t *d_scalars, size_t n, unsigned method,
                              unsigned window_bits, void *stream) {
    if (!ctx || !d_out || !d_points || !d_scalars) return fail(GSN_ERR_INVALID_ARG, "null argument");
    if (method > 2 || window_bits > 16 || window_bits == 1) return fail(GSN_ERR_INVALID_ARG, "method %u, window_bits %u", method, window_bits);
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    if (method == 1 || (method == 0 && n < 8)) return g1_multiexp_naive(ctx, d_out, d_points, d_scalars, n, st);
    return g1_multiexp_pippenger(ctx, d_out, d_points, d_scalars, n, st, window_bits);
}

int gsn_g1_multiexp_device(gsn_ctx *ctx, uint32_t *d_out, const uint32_t *d_points, const uint32_t *d_scalars, size_t n, void *stream) {
    return gsn_g1_multiexp_device_ex(ctx, d_out, d_points, d_scalars, n, 0, 0, stream);
}

int gsn_fp2_binop_device(gsn_ctx *ctx, int op, uint32_t *d_out, const uint32_t *d_a, const uint32_t *d_b, size_t count, void *stream) {
    if (!ctx || !d_out || !d_a || !d_b) return fail(GSN_ERR_INVALID_ARG, "null argument");
    if (op < 0 || op > 2) return fail(GSN_ERR_INVALID_ARG, "op %d", op);
    if (count == 0) return GSN_OK;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    gsn::fp2_binop768<<<(unsigned)((count + 63) / 64), 64, 0, stream ? (cudaStream_t)stream : ctx->stream>>>(d_out, d_a, d_b, count, op);
    ctx->launches++;
    CU(cudaGetLastError());
    return GSN_OK;
}

int gsn_fp2_binop_host(gsn_ctx *ctx, int op, uint32_t *out, const uint32_t *a, const uint32_t *b, size_t count) {
    if (!ctx || !out || !a || !b) return fail(GSN_ERR_INVALID_ARG, "null argument");
    if (count == 0) return GSN_OK;
    DevBuf da, db, dc;
    int rc;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        CU(cudaSetDevice(ctx->device));
        if ((rc = dev_alloc(da, count * 192)) || (rc = dev_alloc(db, count * 192)) || (rc = dev_alloc(dc, count * 192))) return rc;
        CU(cudaMemcpyAsync(da.p, a, count * 192, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(db.p, b, count * 192, cudaMemcpyHostToDevice, ctx->stream));
    }
    if ((rc = gsn_fp2_binop_device(ctx, op, (uint32_t *)dc.p, (const uint32_t *)da.p, (const uint32_t *)db.p, count, nullptr))) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaMemcpyAsync(out, dc.p, count * 192, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return GSN_OK;
}

int gsn_g1_multiexp_host(gsn_ctx *ctx, uint32_t *out, const uint32_t *points, const uint32_t *scalars, size_t n) {
    if (!ctx || !out || !points || !scalars) return fail(GSN_ERR_INVALID_ARG, "null argument");
    DevBuf dp, ds, dout;
    int rc;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        CU(cudaSetDevice(ctx->device));
        if ((rc = dev_alloc(dp, std::max<size_t>(n, 1) * 288)) || (rc = dev_alloc(ds, std::max<size_t>(n, 1) * 96)) || (rc = dev_alloc(dout, 288))) return rc;
        if ((rc = h2d_staged(ctx, dp.p, points, n * 288, ctx->stream)) || (rc = h2d_staged(ctx, ds.p, scalars, n * 96, ctx->stream))) return rc;
    }
    if ((rc = gsn_g1_multiexp_device(ctx, (uint32_t *)dout.p, (const uint32_t *)dp.p, (const uint32_t *)ds.p, n, nullptr))) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaMemcpyAsync(out, dout.p, 288, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return GSN_OK;
}

int gsn_g1_multiexp_multi_host(const int *devices, unsigned n_devices, uint32_t *out, const uint32_t *points, const uint32_t *scalars, size_t n) {
    if (!out || (n && (!points || !scalars))) return fail(GSN_ERR_INVALID_ARG, "null argument");
    int visible = 0;
    CU(cudaGetDeviceCount(&visible));
    std::vector<int> devs;
    if (n_devices == 0) for (int d = 0; d < visible; ++d) devs.push_back(d);
    else {
        if (!devices) return fail(GSN_ERR_INVALID_ARG, "null device list");
        devs.assign(devices, devices + n_devices);
    }
    for (int d : devs) if (d < 0 || d >= visible) return fail(GSN_ERR_INVALID_ARG, "device %d of %d", d, visible);
    if (devs.empty()) return fail(GSN_ERR_INVALID_ARG, "no device");
    // one context per device, created on first use and kept for the life of the process
    static std::mutex mu;
    static std::vector<gsn_ctx *> cached;
    std::vector<gsn_ctx *> ctxs;
    {
        std::lock_guard<std::mutex> lk(mu);
        if ((int)cached.size() < visible) cached.resize(visible, nullptr);
        for (int d : devs) {
            if (!cached[d]) {
                int rc = gsn_ctx_create(&cached[d], d);
                if (rc) return rc;
            }
            ctxs.push_back(cached[d]);
        }
    }
    const size_t G = devs.size(), per = (n + G - 1) / G;
    std::vector<gsn::host::G1Host> part(G);
    std::vector<int> rcs(G, GSN_OK);
    std::vector<std::string> errs(G);
    std::vector<std::thread> threads;
    for (size_t g = 0; g < G; ++g) {
        const size_t lo = std::min(n, g * per), cnt = std::min(n, lo + per) - lo;
        threads.emplace_back([&, g, lo, cnt] {
            if (cnt == 0) { gsn::host::G1Ops(gsn::host::fq_field()).identity(part[g]); return; }   // an empty slice
            rcs[g] = gsn_g1_multiexp_host(ctxs[g], (uint32_t *)&part[g], points + lo * 72, scalars + lo * 24, cnt);
            if (rcs[g]) errs[g] = gsn_last_error();   // the message is thread local
        });
    }
    for (auto &t : threads) t.join();
    for (size_t g = 0; g < G; ++g)
        if (rcs[g]) return fail(rcs[g], "device %d: %s", devs[g], errs[g].c_str());
    gsn::host::G1Ops ops(gsn::host::fq_field());
    gsn::host::G1Host acc;
    ops.identity(acc);
    for (size_t g = 0; g < G; ++g) ops.add(acc, acc, part[g]);
    memcpy(out, &acc, 288);
    return GSN_OK;
}

int gsn_fp768_binop_host(gsn_ctx *ctx, int op, uint32_t *out, const uint32_t *a, const uint32_t *b, size_t count) {
    if (!ctx || !out || !a || !b) return fail(GSN_ERR_INVALID_ARG, "null argument");
    if (op < 0 || op > 2) return fail(GSN_ERR_INVALID_ARG, "op %d", op);
    if (count == 0) return GSN_OK;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    DevBuf da, db, dc;
    int rc;
    if ((rc = dev_alloc(da, count * 96)) || (rc = dev_alloc(db, count * 96)) || (rc = dev_alloc(dc, count * 96))) return rc;
    CU(cudaMemcpyAsync(da.p, a, count * 96, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(db.p, b, count * 96, cudaMemcpyHostToDevice, ctx->stream));
    gsn::binop768<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(ctx->fc, (uint32_t *)dc.p, (const uint32_t *)da.p, (const uint32_t *)db.p, count, op);
    ctx->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, dc.p, count * 96, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return GSN_OK;
}

int gsn_ntt768_device_scatter(gsn_ctx *ctx, const uint32_t *d_limbs, size_t n, size_t batch, unsigned log_r, const uint32_t *omega,
                              unsigned flags, const uint32_t *d_pre_table, uint32_t *const *peers, unsigned n_peers, unsigned my_rank,
                              unsigned rank_shift, unsigned ins_shift, void *stream) {
    if (!ctx || !d_limbs || !omega || !peers) return fail(GSN_ERR_INVALID_ARG, "null argument");
    if (n_peers < 1 || n_peers > 8 || (n_peers & (n_peers - 1)) || my_rank >= n_peers) return fail(GSN_ERR_INVALID_ARG, "n_peers = %u, my_rank = %u", n_peers, my_rank);
    int rc = check_n(n, batch);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    Plan768 *pl;
    const int inv = (flags & GSN_FLAG_INVERSE_ROOT) != 0, scale = inv && !(flags & GSN_FLAG_NO_SCALE);
    if (d_pre_table && scale && n <= 1024) return fail(GSN_ERR_INVALID_ARG, "fold the scale into the pre-twiddle table and pass GSN_FLAG_NO_SCALE");
    if ((rc = get_plan768(ctx, ilog2(n), omega, inv, scale, &pl))) return rc;
    gsn::ScatterDesc sc;
    memset(&sc, 0, sizeof(sc));
    for (unsigned r = 0; r < n_peers; ++r) {
        if (!peers[r]) return fail(GSN_ERR_INVALID_ARG, "peers[%u] is null", r);
        sc.peers[r] = peers[r];
    }
    sc.enabled = 1;
    sc.rank_shift = rank_shift;
    sc.rank_bits = ilog2(n_peers);
    sc.ins_shift = ins_shift;
    sc.my_rank = my_rank;
    // the source is only read: pass 1 goes to the workspace (or, for a one-pass plan, straight to the peers)
    return launch_ntt768_range(ctx, pl, const_cast<uint32_t *>(d_limbs), batch, log_r, ext_flat(d_pre_table), stream ? (cudaStream_t)stream : ctx->stream, 0,
                               pl->digits.size(), 0, 0, &sc);
}

int gsn_peer_barrier(gsn_ctx *ctx, uint32_t *const *peer_flags, unsigned n_peers, unsigned my_rank, unsigned epoch, void *stream) {
    if (!ctx || !peer_flags) return fail(GSN_ERR_INVALID_ARG, "null argument");
    if (n_peers < 1 || n_peers > 8 || my_rank >= n_peers) return fail(GSN_ERR_INVALID_ARG, "n_peers = %u, my_rank = %u", n_peers, my_rank);
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    gsn::PeerFlags pf;
    memset(&pf, 0, sizeof(pf));
    for (unsigned r = 0; r < n_peers; ++r) {
        if (!peer_flags[r]) return fail(GSN_ERR_INVALID_ARG, "peer_flags[%u] is null", r);
        pf.flags[r] = peer_flags[r];
    }
    gsn::peer_barrier_kernel<<<1, 32, 0, stream ? (cudaStream_t)stream : ctx->stream>>>(pf, n_peers, my_rank, epoch, 1u);
    ctx->launches++;
    CU(cudaGetLastError());
    return GSN_OK;
}

int gsn_ipc_export(gsn_ctx *ctx, void *dptr, unsigned char handle[64]) {
    if (!ctx || !dptr || !handle) return fail(GSN_ERR_INVALID_ARG, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    CU(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, dptr));
    memcpy(handle, &h, 64);
    return GSN_OK;
}

int gsn_ipc_import(gsn_ctx *ctx, const unsigned char handle[64], void **dptr) {
    if (!ctx || !dptr || !handle) return fail(GSN_ERR_INVALID_ARG, "null argument");
    CU(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CU(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return GSN_OK;
}

int gsn_ipc_close(gsn_ctx *ctx, void *dptr) {
    if (!ctx) return fail(GSN_ERR_INVALID_ARG, "null ctx");
    CU(cudaSetDevice(ctx->device));
    CU(cudaIpcCloseMemHandle(dptr));
    return GSN_OK;
}

int gsn_fourstep_table768(gsn_ctx *ctx, uint32_t *d_table, size_t rows, size_t cols, size_t row0, size_t col0, size_t n_total,
                          const uint32_t *omega, unsigned flags, void *stream) {
    if (!ctx || !d_table || !omega) return fail(GSN_ERR_INVALID_ARG, "null argument");
    if (!is_pow2(n_total) || rows == 0 || cols == 0) return fail(GSN_ERR_NOT_POW2, "n_total = %zu", n_total);
    const uint32_t logn = ilog2(n_total);
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    if ((int)logn > ctx->two_adicity) return fail(GSN_ERR_TOO_LARGE, "n = 2^%u exceeds the field's 2-adicity %d", logn, ctx->two_adicity);
    int rc = validate_omega768(ctx, omega, logn);
    if (rc) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    uint64_t w_eff[12], sc[12];
    memcpy(w_eff, omega, 96);
    if (flags & GSN_FLAG_INVERSE_ROOT) ctx->hf.pow(w_eff, w_eff, n_total - 1);
    const bool scale = (flags & GSN_FLAG_SCALE_TABLE) != 0;
    if (scale) {
        memcpy(sc, ctx->hf.r1, 96);
        for (uint32_t i = 0; i < logn; ++i) ctx->hf.halve(sc, sc);
    }
    const uint32_t lo_bits = (logn + 1) / 2;
    DevBuf d_w, d_sc, t_lo, t_hi;
    if ((rc = dev_alloc(d_w, 96)) || (rc = dev_alloc(t_lo, (1ull << lo_bits) * 96)) || (rc = dev_alloc(t_hi, (n_total >> lo_bits) * 96))) return rc;
    CU(cudaMemcpyAsync(d_w.p, w_eff, 96, cudaMemcpyHostToDevice, st));
    gsn::pow_table768<<<(unsigned)(((1ull << lo_bits) + 127) / 128), 128, 0, st>>>(ctx->fc, (uint32_t *)t_lo.p, (const uint32_t *)d_w.p, 1ull << lo_bits, 1);
    gsn::pow_table768<<<(unsigned)(((n_total >> lo_bits) + 127) / 128), 128, 0, st>>>(ctx->fc, (uint32_t *)t_hi.p, (const uint32_t *)d_w.p, n_total >> lo_bits, 1ull << lo_bits);
    ctx->launches += 2;
    if (scale) {
        if ((rc = dev_alloc(d_sc, 96))) return rc;
        CU(cudaMemcpyAsync(d_sc.p, sc, 96, cudaMemcpyHostToDevice, st));
        const uint64_t cnt = 1ull << lo_bits;
        gsn::scale_table768<<<(unsigned)((cnt + 127) / 128), 128, 0, st>>>(ctx->fc, (uint32_t *)t_lo.p, (const uint32_t *)t_lo.p, (const uint32_t *)d_sc.p, cnt);
        ctx->launches++;
    }
    const uint64_t cnt = (uint64_t)rows * cols;
    DevBuf tab_m;  // Montgomery form, converted into the caller's table (fixed-operand format, 192 B per entry)
    if ((rc = dev_alloc(tab_m, cnt * 96))) return rc;
    gsn::build_fourstep768<<<(unsigned)((cnt + 127) / 128), 128, 0, st>>>(ctx->fc, (uint32_t *)tab_m.p, (const uint32_t *)t_lo.p, (const uint32_t *)t_hi.p, rows, cols,
                                                                           row0, col0, logn, lo_bits);
    ctx->launches++;
    if ((rc = convert_to_shoup(ctx, d_table, (const uint32_t *)tab_m.p, cnt, st))) return rc;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));  // the temporaries above are freed on return
    return GSN_OK;
}

// ---- helpers
int gsn_host_alloc(void **ptr, size_t bytes) {
    if (!ptr) return fail(GSN_ERR_INVALID_ARG, "null ptr");
    CU(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
    return GSN_OK;
}
int gsn_host_free(void *ptr) { CU(cudaFreeHost(ptr)); return GSN_OK; }
int gsn_device_alloc(gsn_ctx *ctx, void **dptr, size_t bytes) {
    if (!ctx || !dptr) return fail(GSN_ERR_INVALID_ARG, "null argument");
    CU(cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(dptr, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(GSN_ERR_TOO_LARGE, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); }
    return GSN_OK;
}
int gsn_device_free(gsn_ctx *ctx, void *dptr) {
    if (!ctx) return fail(GSN_ERR_INVALID_ARG, "null ctx");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaFree(dptr));
    return GSN_OK;
}
int gsn_memcpy_h2d(gsn_ctx *ctx, void *dptr, const void *hptr, size_t bytes) {
    if (!ctx) return fail(GSN_ERR_INVALID_ARG, "null ctx");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(dptr, hptr, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return GSN_OK;
}
int gsn_memcpy_d2h(gsn_ctx *ctx, void *hptr, const void *dptr, size_t bytes) {
    if (!ctx) return fail(GSN_ERR_INVALID_ARG, "null ctx");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(hptr, dptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return GSN_OK;
}
int gsn_ctx_synchronize(gsn_ctx *ctx) {
    if (!ctx) return fail(GSN_ERR_INVALID_ARG, "null ctx");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return GSN_OK;
}

// ---- INT32 issue-rate probe
}  // extern "C"

template <int MODE>
static void launch_probe(int blocks, int threads, cudaStream_t st, uint32_t *sink, const uint32_t *in, int iters) {
    gsn::int32_issue_probe<MODE><<<blocks, threads, 0, st>>>(sink, in, iters);
}

extern "C" {

int gsn_int32_issue_rates(gsn_ctx *ctx, double *rates, int max_modes, int *n_modes, int *sm_count, int *sm_clock_khz) {
    if (!ctx || !rates || max_modes <= 0) return fail(GSN_ERR_INVALID_ARG, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    DevBuf sink, in;
    int rc;
    if ((rc = dev_alloc(sink, 256)) || (rc = dev_alloc(in, 4096))) return rc;
    {
        uint32_t h[1024];
        for (int i = 0; i < 1024; ++i) h[i] = 0x9E3779B9u * (uint32_t)(i + 1) | 1u;
        CU(cudaMemcpyAsync(in.p, h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    const int iters = 2048, blocks = ctx->sm_count * 8, threads = 256;
    typedef void (*launch_fn)(int, int, cudaStream_t, uint32_t *, const uint32_t *, int);
    static const launch_fn fns[gsn::INT32_PROBE_MODES] = {launch_probe<0>, launch_probe<1>, launch_probe<2>,
                                                         launch_probe<3>, launch_probe<4>, launch_probe<5>,
                                                         launch_probe<6>, launch_probe<7>};
    const int modes = std::min(max_modes, gsn::INT32_PROBE_MODES);
    for (int mode = 0; mode < modes; ++mode) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            CU(cudaEventRecord(ctx->ev0, ctx->stream));
            fns[mode](blocks, threads, ctx->stream, (uint32_t *)sink.p, (const uint32_t *)in.p, iters);
            ctx->launches++;
            CU(cudaEventRecord(ctx->ev1, ctx->stream));
            CU(cudaEventSynchronize(ctx->ev1));
            float ms;
            CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
            if (rep > 0) best = std::min(best, ms);
        }
        const double ops = (double)blocks * threads * iters * gsn::int32_probe_ops_per_iter(mode);
        rates[mode] = ops / (best * 1e-3);
    }
    CU(cudaGetLastError());
    if (n_modes) *n_modes = modes;
    if (sm_count) *sm_count = ctx->sm_count;
    if (sm_clock_khz) { int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device); *sm_clock_khz = khz; }
    return GSN_OK;
}

}  // extern "C"
