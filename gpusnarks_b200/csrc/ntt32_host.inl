// ntt32_host.inl -- host planner + C ABI for the 32-bit field (included by gsn_lib.cu).
// Replaces best_fft for the reference's 32-bit field sketch (fields/dummy_field.h:24-62).
namespace {

constexpr int MAX_PASS_LOG32 = 11;   // stages per pass (2^22 = 11 + 11)
constexpr int MAX_TILE_LOG32 = 13;   // elements per CTA tile (32 KiB of shared memory)
constexpr int NTT32_THREADS = 512;

uint32_t mulmod_h(uint32_t a, uint32_t b, uint32_t p) { return (uint32_t)((uint64_t)a * b % p); }
uint32_t powmod_h(uint32_t a, uint64_t e, uint32_t p) {
    uint32_t acc = 1 % p;
    while (e) { if (e & 1) acc = mulmod_h(acc, a, p); a = mulmod_h(a, a, p); e >>= 1; }
    return acc;
}
bool is_prime32(uint32_t n) {  // deterministic Miller-Rabin for 32-bit integers
    if (n < 2) return false;
    for (uint32_t q : {2u, 3u, 5u, 7u, 11u, 13u}) if (n % q == 0) return n == q;
    uint32_t d = n - 1; int s = 0;
    while (!(d & 1)) { d >>= 1; ++s; }
    for (uint32_t a : {2u, 7u, 61u}) {
        uint32_t x = powmod_h(a % n, d, n);
        if (x == 1 || x == n - 1 || a % n == 0) continue;
        bool comp = true;
        for (int i = 1; i < s && comp; ++i) { x = mulmod_h(x, x, n); if (x == n - 1) comp = false; }
        if (comp) return false;
    }
    return true;
}

int get_plan32(gsn_ctx *ctx, uint32_t logn, uint32_t omega, uint32_t mod, int inverse, Plan32 **out) {
    for (auto &pl : ctx->plans32)
        if (pl->mod == mod && pl->logn == logn && pl->omega == omega && pl->inverse == (inverse != 0)) { *out = pl.get(); return GSN_OK; }
    if (!(mod & 1) || mod >= (1u << 31) || !is_prime32(mod)) return fail(GSN_ERR_BAD_MODULUS, "mod = %u is not an odd prime below 2^31", mod);
    const uint64_t n = 1ull << logn;
    if (logn > 31 || (mod - 1) % n) return fail(GSN_ERR_TOO_LARGE, "n = 2^%u does not divide mod - 1 = %u", logn, mod - 1);
    if (omega >= mod) return fail(GSN_ERR_BAD_OMEGA, "omega = %u is not reduced modulo %u", omega, mod);
    if (logn == 0 ? omega != 1 : powmod_h(omega, n / 2, mod) != mod - 1)
        return fail(GSN_ERR_BAD_OMEGA, "omega = %u is not a primitive 2^%u-th root of unity modulo %u", omega, logn, mod);

    auto pl = std::make_unique<Plan32>();
    pl->mod = mod; pl->omega = omega; pl->logn = logn; pl->inverse = inverse != 0;
    pl->digits = plan_digits(logn, MAX_PASS_LOG32);
    pl->lmax = *std::max_element(pl->digits.begin(), pl->digits.end());
    const size_t P = pl->digits.size();
    pl->pre.resize(P);
    pl->pre_mask.assign(P, 0);
    const uint32_t w_eff = inverse ? powmod_h(omega, n - 1, mod) : omega;
    const uint32_t n_inv = inverse ? powmod_h((uint32_t)(n % mod), mod - 2, mod) : 1;
    cudaStream_t st = ctx->stream;
    int rc;
    const uint64_t half = pl->lmax ? (1ull << (pl->lmax - 1)) : 1;
    if ((rc = dev_alloc(pl->wloc, half * 8))) return rc;
    gsn::pow_table32<<<(unsigned)((half + 255) / 256), 256, 0, st>>>((uint2 *)pl->wloc.p, w_eff, half, pl->lmax ? (n >> pl->lmax) : 0, 1, mod);
    ctx->launches++;
    DevBuf t_lo, t_hi;
    if (P > 1) {
        const uint32_t lo_bits = std::min<uint32_t>(11, logn);
        if ((rc = dev_alloc(t_lo, (1ull << lo_bits) * 8)) || (rc = dev_alloc(t_hi, (n >> lo_bits) * 8))) return rc;
        gsn::pow_table32<<<(unsigned)((n >> lo_bits) + 255) / 256, 256, 0, st>>>((uint2 *)t_hi.p, w_eff, n >> lo_bits, 1ull << lo_bits, 1, mod);
        ctx->launches++;
        for (size_t q = P - 1; q >= 1; --q) {
            uint32_t logN = 0;
            for (size_t i = q - 1; i < P; ++i) logN += pl->digits[i];
            // the low table carries n^-1 for boundary 1 of an inverse plan
            gsn::pow_table32<<<(unsigned)(((1ull << lo_bits) + 255) / 256), 256, 0, st>>>((uint2 *)t_lo.p, w_eff, 1ull << lo_bits, 1, q == 1 ? n_inv : 1, mod);
            pl->pre[q] = std::make_unique<DevBuf>();
            if ((rc = dev_alloc(*pl->pre[q], (1ull << logN) * 8))) return rc;
            pl->pre_mask[q] = (1ull << logN) - 1;
            const uint64_t cnt = 1ull << logN;
            gsn::build_pretw32<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>((uint2 *)pl->pre[q]->p, (const uint2 *)t_lo.p, (const uint2 *)t_hi.p, logN,
                                                                               logN - pl->digits[q - 1], logn - logN, lo_bits, mod);
            ctx->launches += 2;
        }
    } else if (inverse) {
        pl->pre[0] = std::make_unique<DevBuf>();
        if ((rc = dev_alloc(*pl->pre[0], 8))) return rc;
        gsn::pow_table32<<<1, 32, 0, st>>>((uint2 *)pl->pre[0]->p, 1, 1, 0, n_inv, mod);
        ctx->launches++;
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    *out = pl.get();
    ctx->plans32.push_back(std::move(pl));
    return GSN_OK;
}

int launch_ntt32(gsn_ctx *ctx, Plan32 *pl, uint32_t *d_data, size_t batch, cudaStream_t st) {
    const size_t P = pl->digits.size();
    const uint64_t total = (uint64_t)batch << pl->logn;
    uint32_t v2 = 0;
    while (v2 < (uint32_t)MAX_TILE_LOG32 && !((total >> v2) & 1)) ++v2;
    int rc;
    if (P > 1 && (rc = ensure_work(ctx, total * 4))) return rc;
    uint32_t *work = (uint32_t *)ctx->work.p;
    auto kern = gsn::ntt32_pass<NTT32_THREADS>;
    if (!ctx->attr32_set) {
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << MAX_TILE_LOG32) * 4));
        ctx->attr32_set = true;
    }
    uint32_t below = pl->logn;
    for (size_t q = 0; q < P; ++q) {
        below -= pl->digits[q];
        gsn::PassGeom g;
        memset(&g, 0, sizeof(g));
        g.log_l = pl->digits[q];
        g.log_s = below;
        g.log_r = 0;
        // strided digits take the widest tile (more consecutive columns per row segment);
        // the contiguous last digit takes one or two rows per tile
        g.log_tile = below ? v2 : std::min<uint32_t>(v2, std::max<uint32_t>(g.log_l, 11));
        g.wloc_shift = pl->lmax - pl->digits[q];
        g.final_natural = q + 1 == P;
        g.canonical = 1;
        g.ndig = (uint32_t)P;
        for (size_t i = 0; i < P; ++i) g.dig[i] = pl->digits[i];
        g.logn = pl->logn;
        g.has_pre = pl->pre[q] != nullptr;
        g.pre_mask = pl->pre_mask[q];
        const uint32_t *src = q == 0 ? d_data : work;
        uint32_t *dst = (q + 1 == P) ? d_data : work;
        kern<<<(unsigned)(total >> g.log_tile), NTT32_THREADS, ((size_t)4) << g.log_tile, st>>>(
            src, dst, (const uint2 *)pl->wloc.p, g.has_pre ? (const uint2 *)pl->pre[q]->p : nullptr, g, pl->mod);
        ctx->launches++;
    }
    CU(cudaGetLastError());
    return GSN_OK;
}

}  // namespace

extern "C" {

int gsn_ntt32_device(gsn_ctx *ctx, uint32_t *d_a, size_t n, size_t batch, uint32_t omega, uint32_t mod, int inverse, void *stream) {
    if (!ctx || !d_a) return fail(GSN_ERR_INVALID_ARG, "null argument");
    int rc = check_n(n, batch);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    Plan32 *pl;
    if ((rc = get_plan32(ctx, ilog2(n), omega, mod, inverse, &pl))) return rc;
    return launch_ntt32(ctx, pl, d_a, batch, stream ? (cudaStream_t)stream : ctx->stream);
}

int gsn_ntt32_host(gsn_ctx *ctx, uint32_t *a, size_t n, uint32_t omega, uint32_t mod, int inverse) {
    if (!ctx || !a) return fail(GSN_ERR_INVALID_ARG, "null argument");
    int rc = check_n(n, 1);
    if (rc) return rc;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        CU(cudaSetDevice(ctx->device));
        if ((rc = ensure_io(ctx, n * 4))) return rc;
        CU(cudaMemcpyAsync(ctx->io.p, a, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    if ((rc = gsn_ntt32_device(ctx, (uint32_t *)ctx->io.p, n, 1, omega, mod, inverse, nullptr))) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaMemcpyAsync(a, ctx->io.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return GSN_OK;
}

int gsn_ntt32_time_device(gsn_ctx *ctx, uint32_t *d_a, size_t n, size_t batch, uint32_t omega, uint32_t mod, int inverse, int reps,
                          float *ms_each) {
    if (!ctx || !d_a || !ms_each || reps <= 0) return fail(GSN_ERR_INVALID_ARG, "bad argument");
    int rc = gsn_ntt32_device(ctx, d_a, n, batch, omega, mod, inverse, nullptr);  // builds the plan, warms up
    if (rc) return rc;
    for (int i = 0; i < reps; ++i) {
        CU(cudaEventRecord(ctx->ev0, ctx->stream));
        if ((rc = gsn_ntt32_device(ctx, d_a, n, batch, omega, mod, inverse, nullptr))) return rc;
        CU(cudaEventRecord(ctx->ev1, ctx->stream));
        CU(cudaEventSynchronize(ctx->ev1));
        CU(cudaEventElapsedTime(&ms_each[i], ctx->ev0, ctx->ev1));
    }
    return GSN_OK;
}

}  // extern "C"
