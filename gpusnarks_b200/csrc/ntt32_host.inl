// ntt32_host.inl -- host planner + C ABI for the 32-bit field (included by gsn_lib.cu).
// Replaces best_fft for the reference's 32-bit field sketch (fields/dummy_field.h:24-62).
namespace {

constexpr int MAX_PASS_LOG32 = 12;   // stages per pass (2^22 = 11 + 11, 2^24 = 12 + 12)
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
    // fast path: every digit has 8..12 stages (n >= 2^16) -> register-blocked in-tile four-step kernels
    pl->fast = P >= 2 && P <= 4;
    for (size_t q = 0; q < P; ++q) pl->fast = pl->fast && pl->digits[q] >= 8 && pl->digits[q] <= 12;
    if (P > 1) {
        const uint32_t lo_bits = std::min<uint32_t>(11, logn);
        pl->lo_bits = lo_bits;
        DevBuf &t_lo = pl->t_lo, &t_hi = pl->t_hi;
        if ((rc = dev_alloc(t_lo, (1ull << lo_bits) * 8)) || (rc = dev_alloc(t_hi, (n >> lo_bits) * 8))) return rc;
        gsn::pow_table32<<<(unsigned)((n >> lo_bits) + 255) / 256, 256, 0, st>>>((uint2 *)t_hi.p, w_eff, n >> lo_bits, 1ull << lo_bits, 1, mod);
        ctx->launches++;
        if (pl->fast) {
            auto split = [](uint32_t L, uint32_t &a, uint32_t &b, uint32_t &c) {
                a = L >= 10 ? 4 : 3;
                c = L == 12 ? 4 : (L == 8 ? 2 : 3);
                b = L - a - c;
            };
            // in-tile four-step tables from the plain low table
            gsn::pow_table32<<<(unsigned)(((1ull << lo_bits) + 255) / 256), 256, 0, st>>>((uint2 *)t_lo.p, w_eff, 1ull << lo_bits, 1, 1, mod);
            ctx->launches++;
            pl->tA.resize(P);
            pl->tB.resize(P);
            pl->tG.resize(P);
            pl->consts.resize(P);
            uint32_t inv = 1;  // Newton: p^-1 mod 2^32
            for (int i = 0; i < 5; ++i) inv *= 2 - mod * inv;
            const uint32_t w16 = powmod_h(w_eff, n >> 4, mod);
            uint32_t below = logn;
            for (size_t q = 0; q < P; ++q) {
                const uint32_t L = pl->digits[q];
                below -= L;  // log2 stride of this digit
                uint32_t a, b, c;
                split(L, a, b, c);
                pl->tA[q] = std::make_unique<DevBuf>();
                pl->tB[q] = std::make_unique<DevBuf>();
                if ((rc = dev_alloc(*pl->tA[q], (1ull << L) * 8)) || (rc = dev_alloc(*pl->tB[q], (1ull << (b + c)) * 8))) return rc;
                gsn::build_pretw32<<<(unsigned)(((1ull << L) + 255) / 256), 256, 0, st>>>((uint2 *)pl->tA[q]->p, (const uint2 *)t_lo.p, (const uint2 *)t_hi.p, L,
                                                                                         b + c, logn - L, lo_bits, mod);
                gsn::build_pretw32<<<(unsigned)(((1ull << (b + c)) + 255) / 256), 256, 0, st>>>((uint2 *)pl->tB[q]->p, (const uint2 *)t_lo.p, (const uint2 *)t_hi.p,
                                                                                               b + c, c, logn - (b + c), lo_bits, mod);
                ctx->launches += 2;
                gsn::Ntt32Consts &k = pl->consts[q];
                memset(&k, 0, sizeof(k));
                uint32_t acc = 1;
                for (int e = 0; e < 8; ++e) {
                    k.rt[e] = make_uint2(acc, (uint32_t)(((uint64_t)acc << 32) / mod));
                    acc = mulmod_h(acc, w16, mod);
                }
                k.p = mod;
                k.pinv = inv;
                k.pre_lo_bits = lo_bits;
                // the last pass groups sub-transforms whose OUTPUTS are adjacent: k_1 varies fastest in the output
                k.slot_shift = (q + 1 == P && P > 2) ? (logn - pl->digits[0] - L) : 0;
                if (q >= 1) {
                    // boundary q: w_N^(k * rest), N = 2^(l_{q-1} + ... + l_P), k = digit q-1, rest = (digit q.. | )
                    uint32_t logN = 0;
                    for (size_t i = q - 1; i < P; ++i) logN += pl->digits[i];
                    k.pre_k_bits = pl->digits[q - 1];
                    k.pre_logN = logN;
                    k.pre_exp_shift = logn - logN;
                    pl->pre_mask[q] = (1ull << logN) - 1;
                    // per-k step of the recurrence along the 2^a elements a thread holds: w_N^(k * (QA << log_s))
                    const uint64_t rows = 1ull << k.pre_k_bits;
                    pl->tG[q] = std::make_unique<DevBuf>();
                    if ((rc = dev_alloc(*pl->tG[q], rows * 8))) return rc;
                    const uint64_t step_e = ((1ull << (L - a)) << below) << k.pre_exp_shift;  // exponent of w_n
                    gsn::pow_table32<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>((uint2 *)pl->tG[q]->p, powmod_h(w_eff, step_e, mod), rows, 1, 1, mod);
                    ctx->launches++;
                }
            }
            // low tables of the inter-pass twiddles in Montgomery form (x 2^32); boundary 1 also carries n^-1
            const uint32_t r32 = (uint32_t)((1ull << 32) % mod);
            if ((rc = dev_alloc(pl->t_lo_scaled, (1ull << lo_bits) * 8))) return rc;
            gsn::pow_table32<<<(unsigned)(((1ull << lo_bits) + 255) / 256), 256, 0, st>>>((uint2 *)pl->t_lo_scaled.p, w_eff, 1ull << lo_bits, 1, mulmod_h(n_inv, r32, mod), mod);
            gsn::pow_table32<<<(unsigned)(((1ull << lo_bits) + 255) / 256), 256, 0, st>>>((uint2 *)t_lo.p, w_eff, 1ull << lo_bits, 1, r32, mod);
            ctx->launches += 2;
        } else {
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

template <int A, int B, int C, bool SLOT_FAST, bool PRE>
int launch_fast32(gsn_ctx *ctx, unsigned grid, cudaStream_t st, const uint32_t *src, uint32_t *dst, const uint2 *tA, const uint2 *tB,
                  const uint2 *t_lo, const uint2 *t_hi, const uint2 *tG, const gsn::PassGeom32 &g, const gsn::Ntt32Consts &k) {
    constexpr int MINB = (A + B + C) == 12 ? 1 : 2;
    auto kern = gsn::ntt32_fast_pass<A, B, C, SLOT_FAST, PRE, MINB>;
    if (!ctx->smem_configured.count((const void *)kern)) {  // function attributes are per device, i.e. per context
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsn::FastTile<A, B, C>::SMEM_BYTES));
        ctx->smem_configured.insert((const void *)kern);
    }
    kern<<<grid, 512, gsn::FastTile<A, B, C>::SMEM_BYTES, st>>>(src, dst, tA, tB, t_lo, t_hi, tG, g, k);
    ctx->launches++;
    return GSN_OK;
}

template <bool SLOT_FAST, bool PRE>
int dispatch_fast32(uint32_t L, gsn_ctx *ctx, unsigned grid, cudaStream_t st, const uint32_t *src, uint32_t *dst, const uint2 *tA,
                    const uint2 *tB, const uint2 *t_lo, const uint2 *t_hi, const uint2 *tG, const gsn::PassGeom32 &g, const gsn::Ntt32Consts &k) {
    switch (L) {
        case 8: return launch_fast32<3, 3, 2, SLOT_FAST, PRE>(ctx, grid, st, src, dst, tA, tB, t_lo, t_hi, tG, g, k);
        case 9: return launch_fast32<3, 3, 3, SLOT_FAST, PRE>(ctx, grid, st, src, dst, tA, tB, t_lo, t_hi, tG, g, k);
        case 10: return launch_fast32<4, 3, 3, SLOT_FAST, PRE>(ctx, grid, st, src, dst, tA, tB, t_lo, t_hi, tG, g, k);
        case 11: return launch_fast32<4, 4, 3, SLOT_FAST, PRE>(ctx, grid, st, src, dst, tA, tB, t_lo, t_hi, tG, g, k);
        default: return launch_fast32<4, 4, 4, SLOT_FAST, PRE>(ctx, grid, st, src, dst, tA, tB, t_lo, t_hi, tG, g, k);
    }
}

int launch_ntt32_fast(gsn_ctx *ctx, Plan32 *pl, uint32_t *d_data, size_t batch, cudaStream_t st) {
    const size_t P = pl->digits.size();
    const uint64_t total = (uint64_t)batch << pl->logn;
    int rc;
    uint32_t *work;
    if ((rc = ensure_work(ctx, st, total * 4, &work))) return rc;
    uint32_t below = pl->logn;
    for (size_t q = 0; q < P; ++q) {
        below -= pl->digits[q];
        gsn::PassGeom32 g;
        memset(&g, 0, sizeof(g));
        g.log_l = pl->digits[q];
        g.log_s = below;
        g.final_natural = q + 1 == P;
        g.canonical = 1;
        g.ndig = (uint32_t)P;
        for (size_t i = 0; i < P; ++i) g.dig[i] = pl->digits[i];
        g.logn = pl->logn;
        g.has_pre = q >= 1;
        g.pre_mask = pl->pre_mask[q];
        const unsigned grid = (unsigned)((total >> g.log_l) / 8);
        const uint2 *tA = (const uint2 *)pl->tA[q]->p, *tB = (const uint2 *)pl->tB[q]->p;
        const uint2 *t_lo = (const uint2 *)(q == 1 ? pl->t_lo_scaled.p : pl->t_lo.p), *t_hi = (const uint2 *)pl->t_hi.p;
        const uint2 *tG = q >= 1 ? (const uint2 *)pl->tG[q]->p : nullptr;
        // pass 1 reads the caller's buffer, the last pass writes it, middle passes run in place in the workspace
        const uint32_t *src = q == 0 ? d_data : work;
        uint32_t *dst = (q + 1 == P) ? d_data : work;
        if (q == 0) rc = dispatch_fast32<true, false>(g.log_l, ctx, grid, st, src, dst, tA, tB, t_lo, t_hi, tG, g, pl->consts[q]);
        else if (q + 1 < P) rc = dispatch_fast32<true, true>(g.log_l, ctx, grid, st, src, dst, tA, tB, t_lo, t_hi, tG, g, pl->consts[q]);
        else rc = dispatch_fast32<false, true>(g.log_l, ctx, grid, st, src, dst, tA, tB, t_lo, t_hi, tG, g, pl->consts[q]);
        if (rc) return rc;
    }
    CU(cudaGetLastError());
    return GSN_OK;
}

int launch_ntt32(gsn_ctx *ctx, Plan32 *pl, uint32_t *d_data, size_t batch, cudaStream_t st) {
    if (pl->fast) return launch_ntt32_fast(ctx, pl, d_data, batch, st);
    const size_t P = pl->digits.size();
    const uint64_t total = (uint64_t)batch << pl->logn;
    uint32_t v2 = 0;
    while (v2 < (uint32_t)MAX_TILE_LOG32 && !((total >> v2) & 1)) ++v2;
    int rc;
    uint32_t *work = nullptr;
    if (P > 1 && (rc = ensure_work(ctx, st, total * 4, &work))) return rc;
    auto kern = gsn::ntt32_pass<NTT32_THREADS>;
    if (!ctx->smem_configured.count((const void *)kern)) {
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << MAX_TILE_LOG32) * 4));
        ctx->smem_configured.insert((const void *)kern);
    }
    uint32_t below = pl->logn;
    for (size_t q = 0; q < P; ++q) {
        below -= pl->digits[q];
        gsn::PassGeom32 g;
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
