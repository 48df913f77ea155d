// fourstep_host.inl -- element-wise twists (coset transforms), the four-step multi-GPU plan object and the
// single-process multi-GPU entry point (included by gsn_lib.cu).
//
// Reference structure: _basic_parallel_radix2_FFT_inner (reference test/fft_host.h:56-117) is the same four-step
// with the short side done naively on OpenMP threads; here the short side is a fast column transform whose last pass
// stores straight into the peer GPUs' row buffers, and the long side starts per source rank as its columns arrive.
namespace {

// ------------------------------------------------------------------------------ element-wise twists w^(k*r)
// A twist is described by a PreDesc "shape" (how k and r are cut out of the element index); it is realised either
// as a flat table of `count` entries (one product per element) or as two-level tables (two products, 2 sqrt entries).
struct Twist768 {
    DevBuf t_lo, t_hi, flat;
    gsn::PreDesc pd;
    size_t bytes = 0;
    bool present = false;
};

int build_twist(gsn_ctx *ctx, Twist768 &tw, const uint64_t *base_h, const uint64_t *scale_h, uint32_t exp_bits, gsn::PreDesc shape, uint64_t count,
                cudaStream_t st) {
    int rc;
    const uint32_t lo_bits = (exp_bits + 1) / 2, hi_bits = exp_bits - lo_bits;
    const uint64_t nlo = 1ull << lo_bits, nhi = 1ull << hi_bits;
    DevBuf d_w, d_sc, lo_m, hi_m;
    if ((rc = dev_alloc(d_w, 96)) || (rc = dev_alloc(lo_m, nlo * 96)) || (rc = dev_alloc(hi_m, nhi * 96))) return rc;
    CU(cudaMemcpyAsync(d_w.p, base_h, 96, cudaMemcpyHostToDevice, st));
    gsn::pow_table768<<<(unsigned)((nlo + 127) / 128), 128, 0, st>>>(ctx->fc, (uint32_t *)lo_m.p, (const uint32_t *)d_w.p, nlo, 1);
    gsn::pow_table768<<<(unsigned)((nhi + 127) / 128), 128, 0, st>>>(ctx->fc, (uint32_t *)hi_m.p, (const uint32_t *)d_w.p, nhi, nlo);
    ctx->launches += 2;
    if (scale_h) {
        if ((rc = dev_alloc(d_sc, 96))) return rc;
        CU(cudaMemcpyAsync(d_sc.p, scale_h, 96, cudaMemcpyHostToDevice, st));
        gsn::scale_table768<<<(unsigned)((nlo + 127) / 128), 128, 0, st>>>(ctx->fc, (uint32_t *)lo_m.p, (const uint32_t *)lo_m.p, (const uint32_t *)d_sc.p, nlo);
        ctx->launches++;
    }
    shape.lo_bits = lo_bits;
    tw.pd = shape;
    if ((size_t)count * 192 <= ctx->flat_table_limit) {
        DevBuf flat_m;
        if ((rc = dev_alloc(flat_m, count * 96)) || (rc = dev_alloc(tw.flat, count * 192))) return rc;
        gsn::PreDesc src = shape;
        src.mode = 2;
        src.tab = (const uint32_t *)lo_m.p;
        src.tab_hi = (const uint32_t *)hi_m.p;
        gsn::materialize_pre768<<<(unsigned)((count + 127) / 128), 128, 0, st>>>(ctx->fc, (uint32_t *)flat_m.p, src, count);
        ctx->launches++;
        if ((rc = convert_to_shoup(ctx, (uint32_t *)tw.flat.p, (const uint32_t *)flat_m.p, count, st))) return rc;
        CU(cudaStreamSynchronize(st));
        memset(&tw.pd, 0, sizeof(tw.pd));
        tw.pd.mode = 1;
        tw.pd.tab = (const uint32_t *)tw.flat.p;
        tw.pd.flat_mask = count - 1;  // count is a power of two: batches of transforms wrap around
        tw.bytes = tw.flat.bytes;
    } else {
        if ((rc = dev_alloc(tw.t_lo, nlo * 192)) || (rc = dev_alloc(tw.t_hi, nhi * 192))) return rc;
        if ((rc = convert_to_shoup(ctx, (uint32_t *)tw.t_lo.p, (const uint32_t *)lo_m.p, nlo, st))) return rc;
        if ((rc = convert_to_shoup(ctx, (uint32_t *)tw.t_hi.p, (const uint32_t *)hi_m.p, nhi, st))) return rc;
        tw.pd.mode = 2;
        tw.pd.tab = (const uint32_t *)tw.t_lo.p;
        tw.pd.tab_hi = (const uint32_t *)tw.t_hi.p;
        tw.bytes = tw.t_lo.bytes + tw.t_hi.bytes;
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));  // the Montgomery-form temporaries are freed on return
    tw.present = true;
    return GSN_OK;
}

ExtPre ext_of(const Twist768 &tw) {
    ExtPre e;
    e.pd = tw.pd;
    e.present = tw.present;
    return e;
}

}  // namespace

// coset shift tables g^i (forward) / n^-1 g^-i (inverse), cached per context
struct gsn_coset_entry {
    uint32_t logn = 0, shift[24];
    int field = 0, inverse = 0;
    Twist768 tw;
    uint64_t last_use = 0;
};

namespace {

int get_coset_twist(gsn_ctx *ctx, uint32_t logn, const uint32_t *shift, int inverse, Twist768 **out) {
    for (auto &e : ctx->cosets)
        if (e->field == ctx->field && e->logn == logn && e->inverse == inverse && memcmp(e->shift, shift, 96) == 0) {
            e->last_use = ++ctx->use_clock;
            *out = &e->tw;
            return GSN_OK;
        }
    uint64_t g[12], zero[12] = {0};
    memcpy(g, shift, 96);
    if (!gsn::host::Field768::geq(ctx->hf.p, g) || memcmp(g, ctx->hf.p, 96) == 0 || memcmp(g, zero, 96) == 0)
        return fail(GSN_ERR_INVALID_ARG, "coset shift must be a non-zero reduced field element");
    uint64_t n_inv[12];
    if (inverse) {
        ctx->hf.inv(g, g);
        memcpy(n_inv, ctx->hf.r1, 96);
        for (uint32_t i = 0; i < logn; ++i) ctx->hf.halve(n_inv, n_inv);
    }
    auto e = std::make_unique<gsn_coset_entry>();
    e->logn = logn;
    e->field = ctx->field;
    e->inverse = inverse;
    memcpy(e->shift, shift, 96);
    gsn::PreDesc shape;
    memset(&shape, 0, sizeof(shape));
    shape.k_add = 1;                      // exponent = 1 * (index mod n)
    shape.r_mask = (1ull << logn) - 1;
    shape.logN = 64;                      // g has no small order: no wrap-around
    int rc = build_twist(ctx, e->tw, g, inverse ? n_inv : nullptr, std::max(logn, 1u), shape, 1ull << logn, ctx->stream);
    if (rc) return rc;
    if (ctx->cosets.size() >= 8) {
        size_t lru = 0;
        for (size_t i = 1; i < ctx->cosets.size(); ++i) if (ctx->cosets[i]->last_use < ctx->cosets[lru]->last_use) lru = i;
        cudaDeviceSynchronize();
        ctx->cosets.erase(ctx->cosets.begin() + lru);
    }
    e->last_use = ++ctx->use_clock;
    *out = &e->tw;
    ctx->cosets.push_back(std::move(e));
    return GSN_OK;
}

}  // namespace

// ------------------------------------------------------------------------------ four-step plan (one rank)
struct gsn_fourstep {
    gsn_ctx *ctx = nullptr;
    uint32_t logn = 0, log_n1 = 0, log_n2 = 0, G = 1, logG = 0, rank = 0, logC = 0, logR = 0, cl = 0;
    uint32_t omega[24], w_col[24], w_row[24];
    unsigned directions = 0;
    Twist768 tw_fwd, tw_inv;
    DevBuf x, y[2], flags;          // column-layout buffer, two row-layout receive buffers, 16 flag words
    uint32_t *peer_x[8] = {nullptr}, *peer_y[2][8] = {{nullptr}}, *peer_flags[8] = {nullptr};
    bool connected = false;
    int flip = 1;
    uint32_t epoch = 0;             // barrier / arrival epoch (flags slots 0..7), same sequence on every rank
    bool per_source = false;        // forward row pass starts per source rank (needs >= 2 row passes)
    uint32_t wait_shift = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    float phase_ms[3] = {0, 0, 0};
    uint64_t phase_calls = 0;
    bool timing = false;
};

namespace {

uint32_t split_log_n1(uint32_t logn) { return std::min<uint32_t>(10, logn / 2); }

int fs_signal(gsn_fourstep *fs, cudaStream_t st, bool wait) {
    gsn::PeerFlags pf;
    memset(&pf, 0, sizeof(pf));
    for (uint32_t r = 0; r < fs->G; ++r) pf.flags[r] = fs->peer_flags[r];
    ++fs->epoch;
    gsn::peer_barrier_kernel<<<1, 32, 0, st>>>(pf, fs->G, fs->rank, fs->epoch, wait ? 1u : 0u);
    fs->ctx->launches++;
    CU(cudaGetLastError());
    return GSN_OK;
}

}  // namespace

extern "C" {

int gsn_coset_ntt768_device(gsn_ctx *ctx, uint32_t *d_limbs, size_t n, size_t batch, const uint32_t *omega, const uint32_t *shift, int inverse,
                            void *stream) {
    if (!ctx || !d_limbs || !omega || !shift) return fail(GSN_ERR_INVALID_ARG, "null argument");
    int rc = check_n(n, batch);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    const uint32_t logn = ilog2(n);
    Plan768 *pl;
    // forward: multiply by g^i, then transform.  inverse: transform with omega^-1 (unscaled), then multiply by n^-1 g^-i
    if ((rc = get_plan768(ctx, logn, omega, inverse, 0, &pl))) return rc;
    Twist768 *tw;
    if ((rc = get_coset_twist(ctx, logn, shift, inverse != 0, &tw))) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    ExtPre none = ext_flat(nullptr), e = ext_of(*tw);
    if (!inverse) return launch_ntt768_range(ctx, pl, d_limbs, batch, 0, e, st, 0, pl->digits.size(), 0, 0);
    return launch_ntt768_range(ctx, pl, d_limbs, batch, 0, none, st, 0, pl->digits.size(), 0, 0, nullptr, nullptr, &e);
}

int gsn_coset_ntt768_host(gsn_ctx *ctx, uint32_t *limbs, size_t n, const uint32_t *omega, const uint32_t *shift, int inverse) {
    if (!ctx || !limbs || !omega || !shift) return fail(GSN_ERR_INVALID_ARG, "null argument");
    int rc = check_n(n, 1);
    if (rc) return rc;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        CU(cudaSetDevice(ctx->device));
        if ((rc = ensure_io(ctx, n * 96))) return rc;
        CU(cudaMemcpyAsync(ctx->io.p, limbs, n * 96, cudaMemcpyHostToDevice, ctx->stream));
    }
    if ((rc = gsn_coset_ntt768_device(ctx, (uint32_t *)ctx->io.p, n, 1, omega, shift, inverse, nullptr))) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaMemcpyAsync(limbs, ctx->io.p, n * 96, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return GSN_OK;
}

int gsn_fourstep_create(gsn_ctx *ctx, gsn_fourstep **out, unsigned logn, const uint32_t *omega, unsigned n_ranks, unsigned my_rank, unsigned directions) {
    if (!ctx || !out || !omega) return fail(GSN_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (n_ranks < 1 || n_ranks > 8 || (n_ranks & (n_ranks - 1)) || my_rank >= n_ranks) return fail(GSN_ERR_INVALID_ARG, "n_ranks = %u, my_rank = %u", n_ranks, my_rank);
    if (!(directions & 3)) return fail(GSN_ERR_INVALID_ARG, "directions = %u", directions);
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    if ((int)logn > ctx->two_adicity) return fail(GSN_ERR_TOO_LARGE, "n = 2^%u exceeds the field's 2-adicity %d", logn, ctx->two_adicity);
    int rc = validate_omega768(ctx, omega, logn);
    if (rc) return rc;
    std::unique_ptr<gsn_fourstep> fs(new gsn_fourstep());
    fs->ctx = ctx;
    fs->logn = logn;
    fs->G = n_ranks;
    fs->logG = ilog2(n_ranks);
    fs->rank = my_rank;
    fs->log_n1 = split_log_n1(logn);
    fs->log_n2 = logn - fs->log_n1;
    if (fs->log_n1 < fs->logG || fs->log_n2 < fs->logG) return fail(GSN_ERR_INVALID_ARG, "2^%u is too small to shard over %u ranks", logn, n_ranks);
    fs->logC = fs->log_n2 - fs->logG;
    fs->logR = fs->log_n1 - fs->logG;
    fs->directions = directions;
    memcpy(fs->omega, omega, 96);
    uint64_t w[12], wc[12], wr[12], w_inv[12], n_inv[12];
    memcpy(w, omega, 96);
    ctx->hf.pow(wc, w, 1ull << fs->log_n2);   // n1-th root (column transforms)
    ctx->hf.pow(wr, w, 1ull << fs->log_n1);   // n2-th root (row transforms)
    memcpy(fs->w_col, wc, 96);
    memcpy(fs->w_row, wr, 96);
    // Column ownership: the rank field sits at the top of the LOW digit of the row transform's first pass, so that each
    // sub-transform (and each 1024-element tile) of that pass reads columns of ONE source rank.  With a one-pass row
    // transform this is the plain block layout.
    const std::vector<uint32_t> row_digits = plan_digits(fs->log_n2, MAX_PASS_LOG);
    uint32_t low = 0;
    for (size_t i = 1; i < row_digits.size(); ++i) low += row_digits[i];
    if (row_digits.size() < 2 || low < fs->logG) low = fs->log_n2;
    fs->cl = low - fs->logG;
    if (row_digits.size() >= 2 && fs->G > 1) {
        // first row pass: sub-transform index t = (d_b | r) (low digit of i2, then the row), 2^(10 - l_A) consecutive t per tile
        const int sh = (int)fs->logR + (int)fs->cl - (10 - (int)row_digits[0]);
        const uint64_t row_total = (uint64_t)1 << (fs->logR + fs->log_n2);
        fs->per_source = sh >= 0 && row_total >= ((uint64_t)2 * ctx->sm_count << 10);  // the row pass must use 1024-element tiles
        fs->wait_shift = sh >= 0 ? (uint32_t)sh : 0;
    }
    cudaStream_t st = ctx->stream;
    const uint64_t local = 1ull << (logn - fs->logG);
    if (directions & 1) {   // forward twiddles of the row transforms: w^((rank R + r) * i2), element index i2 * R + r
        gsn::PreDesc shape;
        memset(&shape, 0, sizeof(shape));
        shape.k_mask = (1ull << fs->logR) - 1;
        shape.k_add = (uint64_t)my_rank << fs->logR;
        shape.r_shift = fs->logR;
        shape.r_mask = (1ull << fs->log_n2) - 1;
        shape.logN = logn;
        if ((rc = build_twist(ctx, fs->tw_fwd, w, nullptr, std::max(logn, 1u), shape, local, st))) return rc;
    }
    if (directions & 2) {   // inverse twiddles of the column transforms: n^-1 w^-(k1 * i2(c)), element index k1 * C + c
        memcpy(w_inv, w, 96);
        ctx->hf.pow(w_inv, w_inv, (1ull << logn) - 1);
        memcpy(n_inv, ctx->hf.r1, 96);
        for (uint32_t i = 0; i < logn; ++i) ctx->hf.halve(n_inv, n_inv);
        gsn::PreDesc shape;
        memset(&shape, 0, sizeof(shape));
        shape.k_shift = fs->logC;
        shape.k_mask = ~0ull;
        shape.r_mask = (1ull << fs->logC) - 1;
        shape.gap_shift = fs->cl;
        shape.gap_bits = fs->logG;
        shape.r_add = (uint64_t)my_rank << fs->cl;
        shape.logN = logn;
        if ((rc = build_twist(ctx, fs->tw_inv, w_inv, n_inv, std::max(logn, 1u), shape, local, st))) return rc;
    }
    if ((rc = dev_alloc(fs->x, local * 96)) || (rc = dev_alloc(fs->y[0], local * 96)) || (rc = dev_alloc(fs->y[1], local * 96)) || (rc = dev_alloc(fs->flags, 256))) return rc;
    CU(cudaMemsetAsync(fs->flags.p, 0, 256, st));
    CU(cudaStreamSynchronize(st));
    for (auto &e : fs->ev) CU(cudaEventCreate(&e));
    fs->timing = getenv("GSN_FOURSTEP_TIMING") != nullptr;
    if (n_ranks == 1) {
        fs->peer_x[0] = (uint32_t *)fs->x.p;
        fs->peer_y[0][0] = (uint32_t *)fs->y[0].p;
        fs->peer_y[1][0] = (uint32_t *)fs->y[1].p;
        fs->peer_flags[0] = (uint32_t *)fs->flags.p;
        fs->connected = true;
    }
    *out = fs.release();
    return GSN_OK;
}

int gsn_fourstep_destroy(gsn_fourstep *fs) {
    if (!fs) return GSN_OK;
    cudaSetDevice(fs->ctx->device);
    cudaDeviceSynchronize();
    for (auto &e : fs->ev) if (e) cudaEventDestroy(e);
    delete fs;
    return GSN_OK;
}

int gsn_fourstep_info(gsn_fourstep *fs, unsigned *log_n1, unsigned *log_n2, unsigned *rank_bit, uint64_t *table_bytes, unsigned *per_source) {
    if (!fs) return fail(GSN_ERR_INVALID_ARG, "null plan");
    if (log_n1) *log_n1 = fs->log_n1;
    if (log_n2) *log_n2 = fs->log_n2;
    if (rank_bit) *rank_bit = fs->cl;
    if (table_bytes) *table_bytes = fs->tw_fwd.bytes + fs->tw_inv.bytes;
    if (per_source) *per_source = fs->per_source ? 1 : 0;
    return GSN_OK;
}

int gsn_fourstep_buffers(gsn_fourstep *fs, void **x, void **y0, void **y1, void **flags) {
    if (!fs) return fail(GSN_ERR_INVALID_ARG, "null plan");
    if (x) *x = fs->x.p;
    if (y0) *y0 = fs->y[0].p;
    if (y1) *y1 = fs->y[1].p;
    if (flags) *flags = fs->flags.p;
    return GSN_OK;
}

int gsn_fourstep_connect(gsn_fourstep *fs, void *const *peer_x, void *const *peer_y0, void *const *peer_y1, void *const *peer_flags) {
    if (!fs || !peer_x || !peer_y0 || !peer_y1 || !peer_flags) return fail(GSN_ERR_INVALID_ARG, "null argument");
    for (uint32_t r = 0; r < fs->G; ++r) {
        if (!peer_x[r] || !peer_y0[r] || !peer_y1[r] || !peer_flags[r]) return fail(GSN_ERR_INVALID_ARG, "peer %u: null pointer", r);
        fs->peer_x[r] = (uint32_t *)peer_x[r];
        fs->peer_y[0][r] = (uint32_t *)peer_y0[r];
        fs->peer_y[1][r] = (uint32_t *)peer_y1[r];
        fs->peer_flags[r] = (uint32_t *)peer_flags[r];
    }
    fs->connected = true;
    return GSN_OK;
}

int gsn_fourstep_forward(gsn_fourstep *fs, void *stream, void **y_out) {
    if (!fs) return fail(GSN_ERR_INVALID_ARG, "null plan");
    if (!(fs->directions & 1)) return fail(GSN_ERR_INVALID_ARG, "plan was created without the forward direction");
    if (!fs->connected) return fail(GSN_ERR_INVALID_ARG, "gsn_fourstep_connect has not been called");
    gsn_ctx *ctx = fs->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    int rc;
    // Receive buffers alternate between calls, so the only hazard left is "stores have landed": a rank that has seen
    // every flag of call i+1 knows every rank finished the row pass of call i (signals are stream ordered behind it),
    // hence its buffer of call i may be overwritten by call i+2.
    fs->flip ^= 1;
    const int b = fs->flip;
    Plan768 *pc, *pr;
    if ((rc = get_plan768(ctx, fs->log_n1, fs->w_col, 0, 0, &pc)) || (rc = get_plan768(ctx, fs->log_n2, fs->w_row, 0, 0, &pr))) return rc;
    if (fs->timing) CU(cudaEventRecord(fs->ev[0], st));
    // column transforms; output element (k1 = (h, r), c) goes to rank h, position [i2 = c with this rank inserted at bit cl][r]
    gsn::ScatterDesc sc;
    memset(&sc, 0, sizeof(sc));
    for (uint32_t r = 0; r < fs->G; ++r) sc.peers[r] = fs->peer_y[b][r];
    sc.enabled = 1;
    sc.rank_shift = fs->logR + fs->logC;
    sc.rank_bits = fs->logG;
    sc.rem_bits = fs->logR + fs->logC;   // (r, c) -> (c, r)
    sc.rot_bits = fs->logC;
    sc.ins_shift = fs->logR + fs->cl;
    sc.my_rank = fs->rank;
    if ((rc = launch_ntt768_range(ctx, pc, (uint32_t *)fs->x.p, 1, fs->logC, ext_flat(nullptr), st, 0, pc->digits.size(), 0, 0, &sc))) return rc;
    if (fs->timing) CU(cudaEventRecord(fs->ev[1], st));
    WaitDesc wd;
    const bool per_source = fs->per_source;   // the waiting pass always runs on the warp-owned kernel (launch_ntt768_range)
    if (fs->G > 1) {
        if ((rc = fs_signal(fs, st, !per_source))) return rc;   // per-source: signal only, the row pass waits per tile
        if (per_source) {
            wd.flags = (const uint32_t *)fs->flags.p;
            wd.epoch = fs->epoch;
            wd.shift = fs->wait_shift;
            wd.mask = fs->G - 1;
            wd.first = fs->rank;
        }
    }
    if (fs->timing) CU(cudaEventRecord(fs->ev[2], st));
    // row transforms along i2 with the row index as inner stride: the result A[(rank R + r) + n1 k2] lands at [k2][r]
    if ((rc = launch_ntt768_range(ctx, pr, (uint32_t *)fs->y[b].p, 1, fs->logR, ext_of(fs->tw_fwd), st, 0, pr->digits.size(), 0, 0, nullptr,
                                  per_source ? &wd : nullptr))) return rc;
    if (fs->timing) {
        CU(cudaEventRecord(fs->ev[3], st));
        CU(cudaEventSynchronize(fs->ev[3]));
        for (int i = 0; i < 3; ++i) { float ms; CU(cudaEventElapsedTime(&ms, fs->ev[i], fs->ev[i + 1])); fs->phase_ms[i] += ms; }
        fs->phase_calls++;
    }
    if (y_out) *y_out = fs->y[b].p;
    return GSN_OK;
}

int gsn_fourstep_inverse(gsn_fourstep *fs, void *stream, void **x_out) {
    if (!fs) return fail(GSN_ERR_INVALID_ARG, "null plan");
    if (!(fs->directions & 2)) return fail(GSN_ERR_INVALID_ARG, "plan was created without the inverse direction");
    if (!fs->connected) return fail(GSN_ERR_INVALID_ARG, "gsn_fourstep_connect has not been called");
    gsn_ctx *ctx = fs->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    int rc;
    const int b = fs->flip;   // the row buffer the last forward() filled (or the caller wrote)
    Plan768 *pc, *pr;
    if ((rc = get_plan768(ctx, fs->log_n1, fs->w_col, 1, 0, &pc)) || (rc = get_plan768(ctx, fs->log_n2, fs->w_row, 1, 0, &pr))) return rc;
    // every rank must be done READING x (previous inverse's consumers / the caller's forward) before peers overwrite it
    if (fs->G > 1 && (rc = fs_signal(fs, st, true))) return rc;
    // inverse row transforms; output element (i2, r) goes to rank i2[cl, cl + logG), position [k1 = rank R + r][c = i2 without that field]
    gsn::ScatterDesc sc;
    memset(&sc, 0, sizeof(sc));
    for (uint32_t r = 0; r < fs->G; ++r) sc.peers[r] = fs->peer_x[r];
    sc.enabled = 1;
    sc.rank_shift = fs->logR + fs->cl;
    sc.rank_bits = fs->logG;
    sc.rem_bits = fs->logR + fs->logC;   // (c, r) -> (r, c)
    sc.rot_bits = fs->logR;
    sc.ins_shift = fs->logR + fs->logC;
    sc.my_rank = fs->rank;
    if ((rc = launch_ntt768_range(ctx, pr, (uint32_t *)fs->y[b].p, 1, fs->logR, ext_flat(nullptr), st, 0, pr->digits.size(), 0, 0, &sc))) return rc;
    if (fs->G > 1 && (rc = fs_signal(fs, st, true))) return rc;
    if ((rc = launch_ntt768_range(ctx, pc, (uint32_t *)fs->x.p, 1, fs->logC, ext_of(fs->tw_inv), st, 0, pc->digits.size(), 0, 0))) return rc;
    if (x_out) *x_out = fs->x.p;
    return GSN_OK;
}

int gsn_fourstep_set_timing(gsn_fourstep *fs, int on) {
    if (!fs) return fail(GSN_ERR_INVALID_ARG, "null plan");
    fs->timing = on != 0;
    fs->phase_calls = 0;
    for (float &m : fs->phase_ms) m = 0.f;
    return GSN_OK;
}

int gsn_fourstep_phase_ms(gsn_fourstep *fs, float ms[3], uint64_t *calls) {
    if (!fs || !ms) return fail(GSN_ERR_INVALID_ARG, "null argument");
    for (int i = 0; i < 3; ++i) ms[i] = fs->phase_calls ? fs->phase_ms[i] / (float)fs->phase_calls : 0.f;
    if (calls) *calls = fs->phase_calls;
    return GSN_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------ one process, several GPUs
struct gsn_multi {
    unsigned G = 0, logn = 0;
    std::vector<gsn_ctx *> ctxs;
    std::vector<gsn_fourstep *> plans;
    std::vector<int> devices;
    std::vector<cudaEvent_t> done;
};

extern "C" {

int gsn_multi_destroy(gsn_multi *m) {
    if (!m) return GSN_OK;
    for (auto p : m->plans) gsn_fourstep_destroy(p);
    for (size_t i = 0; i < m->done.size(); ++i) { cudaSetDevice(m->devices[i]); cudaEventDestroy(m->done[i]); }
    for (auto c : m->ctxs) gsn_ctx_destroy(c);
    delete m;
    return GSN_OK;
}

int gsn_multi_create(gsn_multi **out, const int *devices, unsigned n_devices, size_t n, const uint32_t *omega, unsigned directions) {
    if (!out || !devices || !omega) return fail(GSN_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (n_devices < 1 || n_devices > 8 || (n_devices & (n_devices - 1))) return fail(GSN_ERR_INVALID_ARG, "n_devices = %u (1, 2, 4 or 8)", n_devices);
    int rc = check_n(n, 1);
    if (rc) return rc;
    std::unique_ptr<gsn_multi> m(new gsn_multi());
    m->G = n_devices;
    m->logn = ilog2(n);
    for (unsigned i = 0; i < n_devices; ++i) {
        gsn_ctx *c = nullptr;
        if ((rc = gsn_ctx_create(&c, devices[i]))) { gsn_multi_destroy(m.release()); return rc; }
        m->ctxs.push_back(c);
        m->devices.push_back(devices[i]);
    }
    for (unsigned i = 0; i < n_devices; ++i) {   // every device maps every other one (NVLink peer access)
        cudaSetDevice(devices[i]);
        for (unsigned j = 0; j < n_devices; ++j) {
            if (i == j) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[i], devices[j]);
            if (!can) { gsn_multi_destroy(m.release()); return fail(GSN_ERR_CUDA, "device %d cannot access device %d", devices[i], devices[j]); }
            cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { gsn_multi_destroy(m.release()); return fail(GSN_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)); }
            cudaGetLastError();
        }
        cudaEvent_t ev;
        cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        m->done.push_back(ev);
    }
    for (unsigned i = 0; i < n_devices; ++i) {
        gsn_fourstep *fs = nullptr;
        if ((rc = gsn_fourstep_create(m->ctxs[i], &fs, m->logn, omega, n_devices, i, directions))) { gsn_multi_destroy(m.release()); return rc; }
        fs->timing = false;   // phase timing synchronises the host inside forward(): impossible with all ranks on one thread
        m->plans.push_back(fs);
    }
    void *px[8], *py0[8], *py1[8], *pf[8];
    for (unsigned i = 0; i < n_devices; ++i) gsn_fourstep_buffers(m->plans[i], &px[i], &py0[i], &py1[i], &pf[i]);
    for (unsigned i = 0; i < n_devices; ++i)
        if ((rc = gsn_fourstep_connect(m->plans[i], px, py0, py1, pf))) { gsn_multi_destroy(m.release()); return rc; }
    *out = m.release();
    return GSN_OK;
}

}  // extern "C"

// One rank's buffer as rows of runs inside the natural-order host vector:
//   column layout (rank g)  x_g[i1][c], c = (c_hi, c_lo): a[i1 n2 + (c_hi, g, c_lo)]   n1 rows of C/2^cl runs of 2^cl elements
//   row layout (rank h)     y_h[k2][r] = A[(h R + r) + n1 k2]                          n2 rows of one run of R elements
struct MultiLayout {
    uint64_t rows, row_bytes;        // the device buffer: rows x row_bytes, contiguous
    uint64_t run_bytes, runs;        // a row is `runs` runs of run_bytes
    uint64_t row_stride, run_stride; // bytes between rows / runs in the host vector
    uint64_t rank_offset;            // bytes: host position of rank g's first run = g * rank_offset
};

// Pageable host vector <-> the ranks' device buffers through the pinned bounce buffers of context 0: host threads
// gather (scatter) 16 MB pieces while the DMA engines of up to four devices move the previous ones.
static int multi_staged_copy(gsn_multi *m, bool to_device, const MultiLayout &L, char *host, const std::vector<char *> &dev) {
    gsn_ctx *c0 = m->ctxs[0];
    int rc;
    const uint64_t piece_rows = std::max<uint64_t>(1, std::min<uint64_t>(L.rows, ((uint64_t)16 << 20) / L.row_bytes));
    if ((rc = ensure_bounce(c0, std::max<size_t>(piece_rows * L.row_bytes, (size_t)16 << 20)))) return rc;
    const uint64_t pieces = (L.rows + piece_rows - 1) / piece_rows, items = pieces * m->G;
    auto rows_of = [&](uint64_t item, uint64_t &r0, uint64_t &nr, unsigned &g) {
        g = (unsigned)(item % m->G);
        r0 = (item / m->G) * piece_rows;
        nr = std::min(piece_rows, L.rows - r0);
    };
    auto host_side = [&](uint64_t item, bool gather) {
        uint64_t r0, nr;
        unsigned g;
        rows_of(item, r0, nr, g);
        char *bounce = (char *)c0->bounce[item & 3];
        char *base = host + g * L.rank_offset;
        const MultiLayout l = L;
        host_pool(c0).run([=](unsigned t, unsigned nt) {
            for (uint64_t r = nr * t / nt; r < nr * (t + 1) / nt; ++r)
                for (uint64_t k = 0; k < l.runs; ++k) {
                    char *h = base + (r0 + r) * l.row_stride + k * l.run_stride, *b = bounce + r * l.row_bytes + k * l.run_bytes;
                    if (gather) stream_copy(b, h, l.run_bytes);
                    else stream_copy(h, b, l.run_bytes);
                }
        });
    };
    auto dma = [&](uint64_t item) -> int {
        uint64_t r0, nr;
        unsigned g;
        rows_of(item, r0, nr, g);
        gsn_ctx *c = m->ctxs[g];
        CU(cudaSetDevice(c->device));
        char *d = dev[g] + r0 * L.row_bytes, *b = (char *)c0->bounce[item & 3];
        if (to_device) CU(cudaMemcpyAsync(d, b, nr * L.row_bytes, cudaMemcpyHostToDevice, c->stream));
        else CU(cudaMemcpyAsync(b, d, nr * L.row_bytes, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaEventRecord(c->ev_chunk[0][item & 3], c->stream));
        return GSN_OK;
    };
    auto wait = [&](uint64_t item) -> int {   // the DMA of `item` (on its device's stream) has finished
        gsn_ctx *c = m->ctxs[item % m->G];
        CU(cudaSetDevice(c->device));
        CU(cudaEventSynchronize(c->ev_chunk[0][item & 3]));
        return GSN_OK;
    };
    if (to_device) {
        for (uint64_t it = 0; it < items; ++it) {
            if (it >= 4 && (rc = wait(it - 4))) return rc;
            host_side(it, true);
            if ((rc = dma(it))) return rc;
        }
        for (uint64_t it = items > 4 ? items - 4 : 0; it < items; ++it)
            if ((rc = wait(it))) return rc;     // the bounce buffers are free again when this returns
    } else {
        for (uint64_t it = 0; it < std::min<uint64_t>(4, items); ++it)
            if ((rc = dma(it))) return rc;
        for (uint64_t it = 0; it < items; ++it) {
            if ((rc = wait(it))) return rc;
            host_side(it, false);
            if (it + 4 < items && (rc = dma(it + 4))) return rc;
        }
    }
    return GSN_OK;
}

extern "C" {

// natural-order host vector -> column layouts (forward) / row layouts -> natural order: strided 2-D copies straight from
// pinned memory, pieces staged through pinned bounce buffers by host threads when the vector is pageable (a std::vector)
int gsn_multi_ntt768_host(gsn_multi *m, uint32_t *limbs, int inverse) {
    if (!m || !limbs) return fail(GSN_ERR_INVALID_ARG, "null argument");
    const unsigned G = m->G;
    gsn_fourstep *f0 = m->plans[0];
    const uint64_t n1 = 1ull << f0->log_n1, n2 = 1ull << f0->log_n2, C = n2 / G, R = n1 / G;
    const uint64_t runw = 1ull << f0->cl;                 // columns of one rank come in runs of 2^cl, every G * 2^cl
    int rc;
    const bool staged = ((uint64_t)96 << m->logn) >= ((uint64_t)4 << 20) && is_pageable(limbs);
    const MultiLayout col{n1, C * 96, runw * 96, C >> f0->cl, n2 * 96, G * runw * 96, runw * 96};
    const MultiLayout row{n2, R * 96, R * 96, 1, n1 * 96, 0, R * 96};
    std::vector<char *> xs(G), ys(G);
    auto buffers = [&] {
        for (unsigned g = 0; g < G; ++g) { xs[g] = (char *)m->plans[g]->x.p; ys[g] = (char *)m->plans[g]->y[m->plans[g]->flip].p; }
    };
    auto sync_all = [&]() -> int {
        for (unsigned g = 0; g < G; ++g) {
            CU(cudaSetDevice(m->ctxs[g]->device));
            CU(cudaStreamSynchronize(m->ctxs[g]->stream));
        }
        return GSN_OK;
    };
    if (staged) {
        buffers();
        if ((rc = multi_staged_copy(m, true, inverse ? row : col, (char *)limbs, inverse ? ys : xs))) return rc;
        for (unsigned g = 0; g < G; ++g)
            if ((rc = inverse ? gsn_fourstep_inverse(m->plans[g], nullptr, nullptr) : gsn_fourstep_forward(m->plans[g], nullptr, nullptr))) return rc;
        if ((rc = sync_all())) return rc;
        buffers();   // forward flips the y buffers
        if ((rc = multi_staged_copy(m, false, inverse ? col : row, (char *)limbs, inverse ? xs : ys))) return rc;
        return sync_all();
    }
    if (!inverse) {
        // rank g, local column c = (c_hi, c_lo): a[i1 * n2 + (c_hi, g, c_lo)]  ->  x_g[i1][c]
        for (unsigned g = 0; g < G; ++g) {
            gsn_ctx *c = m->ctxs[g];
            cudaSetDevice(c->device);
            for (uint64_t ch = 0; ch < (C >> f0->cl); ++ch)
                CU(cudaMemcpy2DAsync((uint32_t *)m->plans[g]->x.p + ch * runw * 24, C * 96, limbs + ((ch * G + g) * runw) * 24, n2 * 96, runw * 96, n1,
                                     cudaMemcpyHostToDevice, c->stream));
        }
        for (unsigned g = 0; g < G; ++g) if ((rc = gsn_fourstep_forward(m->plans[g], nullptr, nullptr))) return rc;
        // y_h[k2][r] = A[(h R + r) + n1 k2]: runs of R elements, n1 apart in the natural-order vector
        for (unsigned h = 0; h < G; ++h) {
            gsn_ctx *c = m->ctxs[h];
            cudaSetDevice(c->device);
            const uint32_t *y = (const uint32_t *)m->plans[h]->y[m->plans[h]->flip].p;
            CU(cudaMemcpy2DAsync(limbs + h * R * 24, n1 * 96, y, R * 96, R * 96, n2, cudaMemcpyDeviceToHost, c->stream));
        }
    } else {
        for (unsigned h = 0; h < G; ++h) {
            gsn_ctx *c = m->ctxs[h];
            cudaSetDevice(c->device);
            gsn_fourstep *fs = m->plans[h];
            uint32_t *y = (uint32_t *)fs->y[fs->flip].p;
            CU(cudaMemcpy2DAsync(y, R * 96, limbs + h * R * 24, n1 * 96, R * 96, n2, cudaMemcpyHostToDevice, c->stream));
        }
        for (unsigned g = 0; g < G; ++g) if ((rc = gsn_fourstep_inverse(m->plans[g], nullptr, nullptr))) return rc;
        for (unsigned g = 0; g < G; ++g) {
            gsn_ctx *c = m->ctxs[g];
            cudaSetDevice(c->device);
            for (uint64_t ch = 0; ch < (C >> f0->cl); ++ch)
                CU(cudaMemcpy2DAsync(limbs + ((ch * G + g) * runw) * 24, n2 * 96, (const uint32_t *)m->plans[g]->x.p + ch * runw * 24, C * 96, runw * 96, n1,
                                     cudaMemcpyDeviceToHost, c->stream));
        }
    }
    return sync_all();
}

int gsn_multi_device_buffers(gsn_multi *m, unsigned rank, void **x, void **y) {
    if (!m || rank >= m->G) return fail(GSN_ERR_INVALID_ARG, "bad argument");
    if (x) *x = m->plans[rank]->x.p;
    if (y) *y = m->plans[rank]->y[m->plans[rank]->flip].p;
    return GSN_OK;
}

// device-resident: x buffers of all ranks -> y buffers (forward) or back (inverse); returns after enqueueing
int gsn_multi_ntt768_device(gsn_multi *m, int inverse) {
    if (!m) return fail(GSN_ERR_INVALID_ARG, "null argument");
    int rc;
    for (unsigned g = 0; g < m->G; ++g)
        if ((rc = inverse ? gsn_fourstep_inverse(m->plans[g], nullptr, nullptr) : gsn_fourstep_forward(m->plans[g], nullptr, nullptr))) return rc;
    return GSN_OK;
}

int gsn_multi_synchronize(gsn_multi *m) {
    if (!m) return fail(GSN_ERR_INVALID_ARG, "null argument");
    for (unsigned g = 0; g < m->G; ++g) {
        cudaSetDevice(m->ctxs[g]->device);
        CU(cudaStreamSynchronize(m->ctxs[g]->stream));
    }
    return GSN_OK;
}

}  // extern "C"
