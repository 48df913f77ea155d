/* gpusnarks_b200.h -- C ABI of libgpusnarks_b200.so (B200 / sm_100a NTT library).
 *
 * This is the drop-in boundary for the reference's FFT hot path.  The reference has no C
 * ABI: its entry point is a C++ template explicitly instantiated inside the CUDA static
 * library,
 *     template <typename FieldT> void best_fft(std::vector<FieldT>& a, const FieldT& omg);
 *         (reference cuda/fft_kernel.h:24-25, defined cuda/fft_kernel.cu:117-147,
 *          instantiated for fields::Scalar at cuda/fft_kernel.cu:150, called from
 *          test/main.cpp:57)
 * include/cuda/fft_kernel.h in this repository keeps that template (same name, same
 * signature, same header path) and forwards to the functions below; INTEGRATION.md shows
 * the binding.  Plain pointers and sizes only; no C++/torch types cross this boundary.
 *
 * Data format (identical to the reference): an element of the 768-bit field is 24
 * little-endian uint32_t limbs (`fields::Scalar::im_rep`, reference cuda/device_field.h:75),
 * elements are stored AoS, 96 bytes each, in Montgomery form (whatever domain the caller's
 * operator* implies -- the transform is linear, so Montgomery in => Montgomery out).  Inputs
 * must be canonical (< p); outputs are canonical.  An element of the 32-bit field is one
 * uint32_t plain residue in [0, mod) (`dummy_fields::Field::im_rep`, reference
 * fields/dummy_field.h:28).
 *
 * Semantics: out[i] = sum_j a[j] * omega^(i*j) mod p, natural order in and out, in place.
 * inverse != 0 computes the same with omega^-1 and scales by n^-1 (libff convention).
 *
 * Every function returns GSN_OK (0) or a GSN_ERR_* code and never exits the process (the
 * reference prints and exit(-1)s, cuda/fft_kernel.cu:34-39,137-142).  gsn_last_error()
 * returns a thread-local message for the last failing call.
 */
#ifndef GPUSNARKS_B200_H
#define GPUSNARKS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSN_OK 0
#define GSN_ERR_INVALID_ARG 1   /* null pointer, bad enum */
#define GSN_ERR_NOT_POW2 2      /* n is not a power of two (reference: assert(a.size()==CONSTRAINTS), fft_kernel.cu:123) */
#define GSN_ERR_TOO_LARGE 3     /* n exceeds 2^two_adicity of the field, or device memory */
#define GSN_ERR_BAD_OMEGA 4     /* omega is not a primitive n-th root of unity */
#define GSN_ERR_CUDA 5          /* CUDA runtime error (message in gsn_last_error) */
#define GSN_ERR_NO_DEVICE 6     /* no CUDA device: there is NO CPU fallback */
#define GSN_ERR_BAD_MODULUS 7   /* 32-bit modulus not an odd prime < 2^31 */

#define GSN_FIELD_MNT4753_FR 0  /* MNT4-753 scalar field (default; the north-star field) */
#define GSN_FIELD_MNT4753_FQ 1  /* MNT4-753 base field = the reference's literal `_mod` (device_field.h:62-65) */

#define GSN_FP768_LIMBS 24

typedef struct gsn_ctx gsn_ctx;

/* ---- context: owns a stream, the device workspace and the cached twiddle tables.
 * Replaces the per-call cudaMalloc pair that the reference never frees (fft_kernel.cu:129-134).
 * Calls on one context are serialised by an internal mutex; the device-pointer entry points may be used with
 * several caller streams at once (each stream gets its own scratch buffer). */
int gsn_ctx_create(gsn_ctx **ctx, int device);
int gsn_ctx_destroy(gsn_ctx *ctx);
const char *gsn_last_error(void);
/* choose the 768-bit modulus of THIS context (default GSN_FIELD_MNT4753_FR).  The field constants travel with every
 * kernel launch as a parameter, so contexts with different fields can share a device, and kernels already enqueued
 * keep the field they were launched with. */
int gsn_set_field768(gsn_ctx *ctx, int field);
/* tuning knobs.  GSN_OPT_FLAT_TABLE_LIMIT (bytes, default 4 GiB): a pass-boundary twiddle table (192 B per element of
 * the boundary's sub-problem, one product per element) larger than this is replaced by two-level tables of
 * 2 * sqrt(n) entries and two products per element -- 2^26 needs 3 MB instead of 13 GB.  GSN_OPT_PLAN_CACHE_BYTES
 * (default 16 GiB, and at most 16 plans): least-recently-used plans are dropped beyond it.  GSN_OPT_KERNEL_VARIANT:
 * kernel used for 1024-element tiles: 4 (or -1) = CTA-wide stages, 1 = warp-owned tiles with wide lazy ranges,
 * 5 = the same with stages 3-4 enumerated CTA-wide (see DESIGN.md for the measurements). */
#define GSN_OPT_FLAT_TABLE_LIMIT 1
#define GSN_OPT_PLAN_CACHE_BYTES 2
#define GSN_OPT_KERNEL_VARIANT 3
int gsn_ctx_set_option(gsn_ctx *ctx, int option, uint64_t value);
/* drop cached plans (twiddle tables) and the workspace */
int gsn_ctx_trim(gsn_ctx *ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int gsn_launch_count(gsn_ctx *ctx, uint64_t *count);

/* ---- 768-bit NTT.  Replaces best_fft<fields::Scalar> (reference cuda/fft_kernel.cu:117-150).
 * host variant: `limbs` is host memory (n * 24 words), copied H2D, transformed, copied back
 * (blocking, like the reference's two cudaMemcpy, fft_kernel.cu:131,144).  Pinned or pageable: multi-pass sizes are
 * pipelined in column blocks (copy-in, first pass, last pass and copy-out of neighbouring blocks overlap); a pageable
 * vector (std::vector) of 4 MiB or more is staged through pinned bounce buffers by a few host threads inside the
 * call (environment GSN_HOST_THREADS, default min(8, cores / 2)). */
int gsn_ntt768_host(gsn_ctx *ctx, uint32_t *limbs, size_t n, const uint32_t omega[GSN_FP768_LIMBS], int inverse);
/* `count` independent host vectors of n elements each, transformed in place.  Same result as `count`
 * calls of gsn_ntt768_host, but the copy-in of vector i+1 overlaps the passes and the copy-out of vector i
 * (two staging buffers, three streams), so a sequence of transforms runs at PCIe speed in both
 * directions at once -- what a prover's back-to-back FFTs want.  Blocking. */
int gsn_ntt768_host_batch(gsn_ctx *ctx, uint32_t *const *limbs, size_t count, size_t n, const uint32_t omega[GSN_FP768_LIMBS],
                          int inverse);
/* device variant: `d_limbs` is device memory holding `batch` consecutive transforms of n
 * elements; stream-ordered on `stream` (a cudaStream_t, NULL = the context's stream);
 * returns after enqueueing. */
int gsn_ntt768_device(gsn_ctx *ctx, uint32_t *d_limbs, size_t n, size_t batch, const uint32_t omega[GSN_FP768_LIMBS],
                      int inverse, void *stream);
/* Build (or fetch) the plan for (n, omega, inverse) without running it: twiddle tables are
 * computed on the device once and cached.  Optional; the NTT calls do it on first use. */
int gsn_ntt768_prepare(gsn_ctx *ctx, size_t n, size_t batch, const uint32_t omega[GSN_FP768_LIMBS], int inverse);

/* Plan of (n, omega, inverse) (built if needed): bytes of twiddle tables it holds, number of passes, how many pass
 * boundaries use the two-level tables, and the plan cache's population.  Any output pointer may be NULL. */
int gsn_ntt768_plan_info(gsn_ctx *ctx, size_t n, const uint32_t omega[GSN_FP768_LIMBS], int inverse, uint64_t *table_bytes,
                         unsigned *passes, unsigned *two_level_boundaries, uint64_t *cached_plans, uint64_t *cached_bytes);

/* Partial transform used by the multi-GPU four-step driver: `d_limbs` holds
 * batch x n x 2^log_r elements indexed (batch | i | r); transforms along i only. */
int gsn_ntt768_strided_device(gsn_ctx *ctx, uint32_t *d_limbs, size_t n, size_t batch, unsigned log_r,
                              const uint32_t omega[GSN_FP768_LIMBS], int inverse, void *stream);
/* General form.  flags: GSN_FLAG_INVERSE_ROOT uses omega^-1 and (unless GSN_FLAG_NO_SCALE)
 * scales by n^-1.  d_pre_table (device, batch*n*2^log_r entries of GSN_TWIDDLE_TABLE_WORDS words, may be NULL): every input
 * element is first multiplied by the table entry of the same index -- the four-step twiddles
 * produced by gsn_fourstep_table768, fused into the first pass of the transform. */
#define GSN_FLAG_INVERSE_ROOT 1u
#define GSN_FLAG_NO_SCALE 2u
#define GSN_FLAG_SCALE_TABLE 4u
int gsn_ntt768_device_ex(gsn_ctx *ctx, uint32_t *d_limbs, size_t n, size_t batch, unsigned log_r,
                         const uint32_t omega[GSN_FP768_LIMBS], unsigned flags, const uint32_t *d_pre_table, void *stream);
/* Fused transform + exchange for the multi-GPU four-step: like gsn_ntt768_device_ex, but the LAST
 * pass stores every output element straight into a peer GPU's buffer (plain stores to CUDA-IPC
 * mapped peer memory, over NVLink) instead of d_limbs, which is only read.  The local natural
 * output index `go` (over batch x n x 2^log_r) names its destination rank in the bit field
 * [rank_shift, rank_shift + log2 n_peers); the destination index is `go` with that field removed
 * and my_rank inserted at bit ins_shift.  peers[r] = base of rank r's receive buffer as mapped
 * in THIS process (gsn_ipc_import; peers[my_rank] = the local buffer).  n_peers = 1, 2, 4 or 8.
 * Callers order it against the consumers of the peer buffers (e.g. a stream-ordered barrier). */
int gsn_ntt768_device_scatter(gsn_ctx *ctx, const uint32_t *d_limbs, size_t n, size_t batch, unsigned log_r,
                              const uint32_t omega[GSN_FP768_LIMBS], unsigned flags, const uint32_t *d_pre_table,
                              uint32_t *const *peers, unsigned n_peers, unsigned my_rank, unsigned rank_shift,
                              unsigned ins_shift, void *stream);
/* ---- four-step plan object: one rank's share of a transform of n = 2^logn elements sharded over n_ranks GPUs
 * (1, 2, 4 or 8), with the exchange fused into the transform kernels (peer stores over NVLink) and the row pass
 * starting per source rank as its columns arrive (arrival flags in peer memory, no NCCL on the data path).
 * n = n1 * n2 (n1 = 2^min(10, logn/2)), input index i = i1*n2 + i2, output index k = k1 + n1*k2, C = n2/G, R = n1/G.
 *   column layout  x[i1][c]  = a[i1*n2 + i2(c)],  i2(c) = c with the rank inserted at bit `rank_bit` (gsn_fourstep_info):
 *                  the rank's columns come in runs of 2^rank_bit, every G * 2^rank_bit (plain block layout when the
 *                  row transform is a single pass)                                              shape (n1, C) elements
 *   row layout     y[k2][r]  = A[(rank*R + r) + n1*k2]                                          shape (n2, R) elements
 * forward: x -> y (the plan alternates between two y buffers; *y_out receives the one just written);
 * inverse: y (the buffer of the last forward, or gsn_fourstep_buffers' y0 on a fresh plan) -> x, with omega^-1 and n^-1.
 * The plan owns its buffers; peers are connected once: in one process by passing the other plans' pointers (with peer
 * access enabled), across processes through gsn_ipc_export / gsn_ipc_import.  directions: bit 0 forward, bit 1 inverse.
 * All ranks must issue the same sequence of forward / inverse calls.  Reference structure: the four-step
 * _basic_parallel_radix2_FFT_inner, test/fft_host.h:56-117. */
typedef struct gsn_fourstep gsn_fourstep;
int gsn_fourstep_create(gsn_ctx *ctx, gsn_fourstep **plan, unsigned logn, const uint32_t omega[GSN_FP768_LIMBS], unsigned n_ranks,
                        unsigned my_rank, unsigned directions);
int gsn_fourstep_destroy(gsn_fourstep *plan);
int gsn_fourstep_info(gsn_fourstep *plan, unsigned *log_n1, unsigned *log_n2, unsigned *rank_bit, uint64_t *table_bytes, unsigned *per_source);
int gsn_fourstep_buffers(gsn_fourstep *plan, void **x, void **y0, void **y1, void **flags);
int gsn_fourstep_connect(gsn_fourstep *plan, void *const *peer_x, void *const *peer_y0, void *const *peer_y1, void *const *peer_flags);
int gsn_fourstep_forward(gsn_fourstep *plan, void *stream, void **y_out);
int gsn_fourstep_inverse(gsn_fourstep *plan, void *stream, void **x_out);
/* mean milliseconds of the three forward phases (column transforms + scatter, signal/barrier, row transforms) over the
 * calls since gsn_fourstep_set_timing(plan, 1).  Timing synchronises the host inside forward(): switch it on for a
 * few calls outside the measured region (one process per GPU only). */
int gsn_fourstep_set_timing(gsn_fourstep *plan, int on);
int gsn_fourstep_phase_ms(gsn_fourstep *plan, float ms[3], uint64_t *calls);

/* ---- multi-GPU transform from ONE process: what a C++ caller of best_fft (reference cuda/fft_kernel.h:24-25) reaches
 * when several devices are visible.  gsn_multi_create builds one context and one four-step plan per device, enables
 * peer access between them and connects the plans.  gsn_multi_ntt768_host transforms a natural-order host vector in
 * place (strided copies distribute it in the column layout and collect the row layouts; blocking).
 * gsn_multi_ntt768_device runs the transform on the plans' device buffers (gsn_multi_device_buffers; asynchronous,
 * gsn_multi_synchronize waits). */
typedef struct gsn_multi gsn_multi;
int gsn_multi_create(gsn_multi **m, const int *devices, unsigned n_devices, size_t n, const uint32_t omega[GSN_FP768_LIMBS], unsigned directions);
int gsn_multi_destroy(gsn_multi *m);
int gsn_multi_ntt768_host(gsn_multi *m, uint32_t *limbs, int inverse);
int gsn_multi_device_buffers(gsn_multi *m, unsigned rank, void **x, void **y);
int gsn_multi_ntt768_device(gsn_multi *m, int inverse);
int gsn_multi_synchronize(gsn_multi *m);

/* ---- coset transforms (the prover pipeline iFFT -> coset FFT -> pointwise -> coset iFFT).
 * forward: evaluations of the polynomial with coefficients a on the coset shift * <omega>: a[i] *= shift^i fused into
 * the first pass as its pre-twiddle.  inverse: coefficients from such evaluations: omega^-1 transform whose last pass
 * multiplies output i by n^-1 * shift^-i (post-twiddle), no separate scaling pass.  The shift tables are built once
 * per (n, shift, direction) from two-level power tables and cached in the context. */
int gsn_coset_ntt768_device(gsn_ctx *ctx, uint32_t *d_limbs, size_t n, size_t batch, const uint32_t omega[GSN_FP768_LIMBS],
                            const uint32_t shift[GSN_FP768_LIMBS], int inverse, void *stream);
int gsn_coset_ntt768_host(gsn_ctx *ctx, uint32_t *limbs, size_t n, const uint32_t omega[GSN_FP768_LIMBS],
                          const uint32_t shift[GSN_FP768_LIMBS], int inverse);

/* Stream-ordered barrier across the ranks of a four-step transform, in peer memory (no NCCL on the
 * hot path): peer_flags[r] = rank r's array of 8 zero-initialised uint32 slots as mapped in this
 * process.  The kernel publishes `epoch` (must increase by one per call, same on all ranks) with
 * release semantics at system scope and waits for every peer's epoch; it traps after ~10 s. */
int gsn_peer_barrier(gsn_ctx *ctx, uint32_t *const *peer_flags, unsigned n_peers, unsigned my_rank, unsigned epoch, void *stream);
/* CUDA IPC plumbing for the peer buffers (one process per GPU): export a handle for memory from
 * gsn_device_alloc, import a peer's handle (enables peer access), close it. */
int gsn_ipc_export(gsn_ctx *ctx, void *dptr, unsigned char handle[64]);
int gsn_ipc_import(gsn_ctx *ctx, const unsigned char handle[64], void **dptr);
int gsn_ipc_close(gsn_ctx *ctx, void *dptr);
/* d_table (rows * cols entries of GSN_TWIDDLE_TABLE_WORDS words) <- omega^(+-(row0 + r) * (col0 + c))  [* n_total^-1 with GSN_FLAG_SCALE_TABLE],
 * r < rows, c < cols, omega a primitive n_total-th root (GSN_FLAG_INVERSE_ROOT: omega^-1):
 * the twiddles between the column and the row transforms of a four-step (Bailey) NTT for
 * the shard that owns rows row0.. and columns col0.. (reference structure: the omega_j /
 * omega_step factors of _basic_parallel_radix2_FFT_inner, test/fft_host.h:79-101). */
int gsn_fourstep_table768(gsn_ctx *ctx, uint32_t *d_table, size_t rows, size_t cols, size_t row0, size_t col0, size_t n_total,
                          const uint32_t omega[GSN_FP768_LIMBS], unsigned flags, void *stream);

/* ---- field arithmetic on arrays (parity tests of the device Montgomery code against the
 * oracle; reference device_field_operators.h:190-214).  op: 0 mul (a*b*R^-1), 1 add, 2 sub. */
int gsn_fp768_binop_host(gsn_ctx *ctx, int op, uint32_t *out, const uint32_t *a, const uint32_t *b, size_t count);

/* device-resident form of the same element-wise operations (point-wise products between the
 * transforms of a prover pipeline), stream-ordered */
int gsn_fp768_binop_device(gsn_ctx *ctx, int op, uint32_t *d_out, const uint32_t *d_a, const uint32_t *d_b, size_t count, void *stream);
/* d_table[i] = scale * base^i, i < count (scale NULL = one()).  Coset transforms: evaluate on the
 * coset g<omega> with gsn_ntt768_device_ex(..., d_pre_table = powers(g)); interpolate from it with an
 * inverse transform followed by gsn_fp768_binop_device(mul, powers(g^-1)). */
int gsn_fp768_powers_device(gsn_ctx *ctx, uint32_t *d_table, size_t count, const uint32_t base[GSN_FP768_LIMBS],
                            const uint32_t *scale, void *stream);
/* Pre-twiddle tables of gsn_ntt768_device_ex / _scatter are in the transform kernels' fixed-operand format:
 * 48 words per entry, w (plain integer) followed by w'' = floor(w * 2^768 / p).  This converts `count` field
 * elements (24 words each, Montgomery form, canonical) into such a table (count * 192 bytes).
 * gsn_fourstep_table768 produces this format directly. */
#define GSN_TWIDDLE_TABLE_WORDS 48
int gsn_fp768_twiddle_table_device(gsn_ctx *ctx, uint32_t *d_table, const uint32_t *d_elems, size_t count, void *stream);
/* sum_i a[i] * b[i] in the field (Montgomery products): the reference's multiexp<Scalar, Scalar>
 * (reference cuda/multi_exp.h:24-25, cuda/multi_exp.cu:104-137; CPU form test/multiexp.h:3-13). */
int gsn_fp768_inner_product_device(gsn_ctx *ctx, uint32_t *d_out, const uint32_t *d_a, const uint32_t *d_b, size_t count, void *stream);
int gsn_fp768_inner_product_host(gsn_ctx *ctx, uint32_t out[GSN_FP768_LIMBS], const uint32_t *a, const uint32_t *b, size_t count);

/* ---- MNT4-753 G1 multi-exponentiation sum_i s_i * P_i: the reference's multiexp<mnt4753_G1, Scalar>
 * (reference cuda/multi_exp.h:24-25, cuda/multi_exp.cu:104-142; group law cuda/device_field.h:296-437).
 * The curve's base field MNT4-753 Fq is built in: these calls do not depend on gsn_set_field768.  Points are
 * homogeneous projective (X, Y, Z), 3 x 24 limbs each, Montgomery form, identity = any point with Z = 0;
 * scalars are raw 768-bit little-endian integers (the reference reads their bits with hasBitAt).  The result
 * is projective with canonical coordinates (any representative of the point: compare as affine points).
 * Algorithm: the bucket method (Pippenger) -- signed c-bit windows, points sorted by (window, bucket), one thread per
 * bucket, a running-sum reduction per window, Horner over the window sums on the host; blocking.  _ex selects the
 * method (0 automatic, 1 the reference's own algorithm: one double-and-add per point then a tree reduction, 2 bucket
 * method) and the window width (0 automatic, else 2..16). */
int gsn_g1_multiexp_host(gsn_ctx *ctx, uint32_t out[72], const uint32_t *points, const uint32_t *scalars, size_t n);
/* the same over several devices of this process: the points are cut into n_devices contiguous slices, every device
 * runs the bucket method on its slice (one host thread per device, contexts cached inside the library) and the partial
 * sums are added on the host -- independent work, no exchange between the devices.  n_devices = 0: every visible device. */
int gsn_g1_multiexp_multi_host(const int *devices, unsigned n_devices, uint32_t out[72], const uint32_t *points, const uint32_t *scalars, size_t n);
int gsn_g1_multiexp_device(gsn_ctx *ctx, uint32_t *d_out, const uint32_t *d_points, const uint32_t *d_scalars, size_t n, void *stream);
int gsn_g1_multiexp_device_ex(gsn_ctx *ctx, uint32_t *d_out, const uint32_t *d_points, const uint32_t *d_scalars, size_t n, unsigned method,
                              unsigned window_bits, void *stream);

/* ---- Fq2 = Fq[u]/(u^2 - 13) element-wise arithmetic: the reference's `fp2` (cuda/device_field.h:220-294; the
 * extension field of MNT4-753's G2).  Elements are (x, y) = x + y u, 2 x 24 limbs, Montgomery form over Fq (built in,
 * independent of gsn_set_field768).  op: 0 mul (Karatsuba, as the reference), 1 add, 2 sub; canonical outputs. */
int gsn_fp2_binop_host(gsn_ctx *ctx, int op, uint32_t *out, const uint32_t *a, const uint32_t *b, size_t count);
int gsn_fp2_binop_device(gsn_ctx *ctx, int op, uint32_t *d_out, const uint32_t *d_a, const uint32_t *d_b, size_t count, void *stream);

/* ---- 32-bit NTT over Z/mod (mod an odd prime < 2^31 with n | mod-1).
 * Replaces best_fft for the reference's 32-bit field sketch (fields/dummy_field.h:24-62). */
int gsn_ntt32_host(gsn_ctx *ctx, uint32_t *a, size_t n, uint32_t omega, uint32_t mod, int inverse);
int gsn_ntt32_device(gsn_ctx *ctx, uint32_t *d_a, size_t n, size_t batch, uint32_t omega, uint32_t mod, int inverse,
                     void *stream);

/* ---- host helpers (so callers need not link the CUDA runtime themselves) */
int gsn_device_count(int *count);
int gsn_host_alloc(void **ptr, size_t bytes);   /* pinned host memory */
int gsn_host_free(void *ptr);
int gsn_device_alloc(gsn_ctx *ctx, void **dptr, size_t bytes);
int gsn_device_free(gsn_ctx *ctx, void *dptr);
int gsn_memcpy_h2d(gsn_ctx *ctx, void *dptr, const void *hptr, size_t bytes);
int gsn_memcpy_d2h(gsn_ctx *ctx, void *hptr, const void *dptr, size_t bytes);
int gsn_ctx_synchronize(gsn_ctx *ctx);

/* ---- measurement: INT32 multiply issue rates of the device (roofline denominator for the
 * 768-bit path).  rates[k] = thread-level instructions of kind k per second, chip wide, from
 * loops with 8 independent accumulators and distinct multiplicand registers:
 *  0 IMAD (mad.lo)   1 IMAD.HI (mad.hi)   2 IMAD.WIDE.U32 accumulate form (32x32+64 -> 64): the
 *  "wide MAC" peak used as roofline denominator   3 IMAD.WIDE.U32 with shared multiplicands
 *  4 IMAD.WIDE.U32.X carry chains (the form the CIOS product issues)   5 IADD3.X carry chains.
 * At most max_modes entries are written; *n_modes receives the count. */
int gsn_int32_issue_rates(gsn_ctx *ctx, double *rates, int max_modes, int *n_modes, int *sm_count, int *sm_clock_khz);
/* time `reps` back-to-back device transforms with CUDA events on the context's stream;
 * ms_each[i] receives the i-th duration in milliseconds (data stays resident). */
int gsn_ntt768_time_device(gsn_ctx *ctx, uint32_t *d_limbs, size_t n, size_t batch, const uint32_t omega[GSN_FP768_LIMBS],
                           int inverse, int reps, float *ms_each);
int gsn_ntt32_time_device(gsn_ctx *ctx, uint32_t *d_a, size_t n, size_t batch, uint32_t omega, uint32_t mod, int inverse,
                          int reps, float *ms_each);

#ifdef __cplusplus
}
#endif
#endif /* GPUSNARKS_B200_H */
