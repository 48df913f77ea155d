// oracle/oracle_capi.cpp -- TEST INFRASTRUCTURE ONLY.
// extern "C" face of the CPU oracle so that tests/ and bench.py (cpu_baseline leg) can
// drive it through ctypes.  Nothing in gpusnarks_b200/ or include/ may link this.
#include <omp.h>
#include <algorithm>
#include <cstdio>
#include <chrono>
#include "field768.h"
#include "field32.h"
#include "fft_host_oracle.h"

using oracle::Fp32;
using oracle::Fp768;

static std::vector<Fp768> load768(const uint32_t *limbs, size_t n) {
    std::vector<Fp768> v(n);
    for (size_t i = 0; i < n; ++i) memcpy(v[i].im_rep, limbs + i * oracle::NL, sizeof(v[i].im_rep));
    return v;
}
static void store768(uint32_t *limbs, const std::vector<Fp768> &v) {
    for (size_t i = 0; i < v.size(); ++i) memcpy(limbs + i * oracle::NL, v[i].im_rep, sizeof(v[i].im_rep));
}

extern "C" {

// field: 0 = MNT4-753 Fr (product field), 1 = MNT4-753 Fq (the reference's literal _mod)
int oracle_set_field768(int field) {
    if (field == 0) oracle::current_modulus() = &oracle::modulus_fr();
    else if (field == 1) oracle::current_modulus() = &oracle::modulus_fq();
    else return -1;
    return 0;
}
void oracle_set_mod32(uint32_t mod) { Fp32::mod() = mod; }
void oracle_set_threads(int t) { omp_set_num_threads(t); }
int oracle_max_threads(void) { return omp_get_max_threads(); }

void oracle_fp768_constants(uint32_t *p, uint32_t *r1, uint32_t *r2, uint32_t *root, uint32_t *np0, int *two_adicity) {
    const oracle::Modulus768 &M = *oracle::current_modulus();
    memcpy(p, M.p, 96); memcpy(r1, M.r1, 96); memcpy(r2, M.r2, 96); memcpy(root, M.root, 96);
    *np0 = M.np0; *two_adicity = M.two_adicity;
}

// element-wise ops on arrays of count elements (24 limbs each); op: 0 mul, 1 add, 2 sub
void oracle_fp768_binop(int op, uint32_t *out, const uint32_t *a, const uint32_t *b, size_t count) {
    for (size_t i = 0; i < count; ++i) {
        Fp768 x(a + i * 24), y(b + i * 24), z;
        z = op == 0 ? x * y : op == 1 ? x + y : x - y;
        memcpy(out + i * 24, z.im_rep, 96);
    }
}
void oracle_fp768_pow(uint32_t *out, const uint32_t *a, uint64_t e) { Fp768 z = Fp768(a) ^ e; memcpy(out, z.im_rep, 96); }
void oracle_fp768_inverse(uint32_t *out, const uint32_t *a) { Fp768 z = Fp768(a).inverse(); memcpy(out, z.im_rep, 96); }

// forward transform, reference test/fft_host.h entry points.
// log_cpus < 0 -> _basic_serial_radix2_FFT, else _basic_parallel_radix2_FFT_inner(log_cpus)
void oracle_fft768(uint32_t *limbs, size_t n, const uint32_t *omega, int log_cpus) {
    std::vector<Fp768> v = load768(limbs, n);
    if (log_cpus < 0) oracle::_basic_serial_radix2_FFT(v, Fp768(omega), Fp768::one());
    else oracle::_basic_parallel_radix2_FFT_inner(v, Fp768(omega), (size_t)log_cpus, Fp768::one());
    store768(limbs, v);
}
void oracle_ifft768(uint32_t *limbs, size_t n, const uint32_t *omega, int log_cpus) {
    std::vector<Fp768> v = load768(limbs, n);
    oracle::inverse_FFT(v, Fp768(omega), (size_t)(log_cpus < 0 ? 0 : log_cpus));
    store768(limbs, v);
}
void oracle_naive_dft768(uint32_t *limbs, size_t n, const uint32_t *omega) {
    std::vector<Fp768> v = load768(limbs, n);
    store768(limbs, oracle::naive_dft(v, Fp768(omega)));
}
// out[i] = A[ks[i]] for count spot indices (Horner, n multiplies each)
void oracle_dft_points768(uint32_t *out, const uint32_t *limbs, size_t n, const uint32_t *omega, const uint64_t *ks, size_t count) {
    const Fp768 *a = reinterpret_cast<const Fp768 *>(limbs);
    static_assert(sizeof(Fp768) == 96, "layout");
#pragma omp parallel for
    for (size_t i = 0; i < count; ++i) {
        Fp768 z = oracle::dft_point(a, n, Fp768(omega), ks[i]);
        memcpy(out + i * 24, z.im_rep, 96);
    }
}

// The same spot values with every host thread busy: the index range of each Horner evaluation is cut into chunks,
// chunk c covers j in [j0, j1) and yields P_c = sum_j a[j] x^(j - j0); the value is sum_c P_c x^j0 (x = omega^k).
// (bench.py's parity leg at 2^24 / 2^26, where one evaluation is 10^7..10^8 products.)
void oracle_dft_points768_mt(uint32_t *out, const uint32_t *limbs, size_t n, const uint32_t *omega, const uint64_t *ks, size_t count) {
    const Fp768 *a = reinterpret_cast<const Fp768 *>(limbs);
    const size_t chunks = std::max<size_t>(1, std::min<size_t>((size_t)omp_get_max_threads() * 4, n / 4096 + 1));
    const size_t per = (n + chunks - 1) / chunks;
    std::vector<Fp768> partial(count * chunks, Fp768::zero());
#pragma omp parallel for collapse(2) schedule(dynamic)
    for (size_t i = 0; i < count; ++i)
        for (size_t c = 0; c < chunks; ++c) {
            const size_t j0 = c * per, j1 = std::min(n, j0 + per);
            if (j0 >= j1) continue;
            const Fp768 x = Fp768(omega) ^ ks[i];
            Fp768 acc = Fp768::zero();
            for (size_t j = j1; j-- > j0;) acc = acc * x + a[j];
            partial[i * chunks + c] = acc * (x ^ (uint64_t)j0);
        }
    for (size_t i = 0; i < count; ++i) {
        Fp768 z = Fp768::zero();
        for (size_t c = 0; c < chunks; ++c) z = z + partial[i * chunks + c];
        memcpy(out + i * 24, z.im_rep, 96);
    }
}

void oracle_fft32(uint32_t *a, size_t n, uint32_t omega, int log_cpus) {
    std::vector<Fp32> v(n);
    for (size_t i = 0; i < n; ++i) v[i].im_rep = a[i];
    if (log_cpus < 0) oracle::_basic_serial_radix2_FFT(v, Fp32(omega), Fp32::one());
    else oracle::_basic_parallel_radix2_FFT_inner(v, Fp32(omega), (size_t)log_cpus, Fp32::one());
    for (size_t i = 0; i < n; ++i) a[i] = v[i].im_rep;
}
void oracle_ifft32(uint32_t *a, size_t n, uint32_t omega, int log_cpus) {
    std::vector<Fp32> v(n);
    for (size_t i = 0; i < n; ++i) v[i].im_rep = a[i];
    oracle::inverse_FFT(v, Fp32(omega), (size_t)(log_cpus < 0 ? 0 : log_cpus));
    for (size_t i = 0; i < n; ++i) a[i] = v[i].im_rep;
}
void oracle_naive_dft32(uint32_t *a, size_t n, uint32_t omega) {
    std::vector<Fp32> v(n);
    for (size_t i = 0; i < n; ++i) v[i].im_rep = a[i];
    std::vector<Fp32> o = oracle::naive_dft(v, Fp32(omega));
    for (size_t i = 0; i < n; ++i) a[i] = o[i].im_rep;
}
void oracle_dft_points32(uint32_t *out, const uint32_t *a, size_t n, uint32_t omega, const uint64_t *ks, size_t count) {
    const Fp32 *v = reinterpret_cast<const Fp32 *>(a);
#pragma omp parallel for
    for (size_t i = 0; i < count; ++i) out[i] = oracle::dft_point(v, n, Fp32(omega), ks[i]).im_rep;
}

// wall-clock seconds of one forward transform on the host cores (CPU baseline leg)
double oracle_time_fft768(uint32_t *limbs, size_t n, const uint32_t *omega, int log_cpus) {
    std::vector<Fp768> v = load768(limbs, n);
    auto t0 = std::chrono::steady_clock::now();
    if (log_cpus < 0) oracle::_basic_serial_radix2_FFT(v, Fp768(omega), Fp768::one());
    else oracle::_basic_parallel_radix2_FFT_inner(v, Fp768(omega), (size_t)log_cpus, Fp768::one());
    auto t1 = std::chrono::steady_clock::now();
    store768(limbs, v);
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"

extern "C" double oracle_time_fft32(uint32_t *a, size_t n, uint32_t omega, int log_cpus) {
    std::vector<Fp32> v(n);
    for (size_t i = 0; i < n; ++i) v[i].im_rep = a[i];
    auto t0 = std::chrono::steady_clock::now();
    if (log_cpus < 0) oracle::_basic_serial_radix2_FFT(v, Fp32(omega), Fp32::one());
    else oracle::_basic_parallel_radix2_FFT_inner(v, Fp32(omega), (size_t)log_cpus, Fp32::one());
    auto t1 = std::chrono::steady_clock::now();
    for (size_t i = 0; i < n; ++i) a[i] = v[i].im_rep;
    return std::chrono::duration<double>(t1 - t0).count();
}
