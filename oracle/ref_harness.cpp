// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Thin extern "C" harness around the REFERENCE'S OWN host sources, compiled from where
// they lie (/root/reference, or oracle/_ref/patched for the sed-corrected copy that the
// Makefile creates and deletes at build time).  No reference code is restated here: this
// file only includes the reference headers and calls them.
//
//   reference cuda/device_field.h + cuda/device_field_operators.h -> fields::Scalar
//   reference test/fft_host.h -> _basic_serial_radix2_FFT, _basic_parallel_radix2_FFT_inner
//
// Built twice (see Makefile): libref_verbatim.so (Oracle-V: characterisation + CPU
// baseline timing, mirrors the host half of reference test/main.cpp:38-76) and
// libref_patched.so (documented one-line corrections; pins the oracle restatement).
#include <omp.h>
#include <chrono>
#include <cstring>
#include <vector>

#include <cuda/device_field.h>
#include <cuda/device_field_operators.h>
#include "fft_host.h"

// The reference templates instantiated over the ORACLE's (correct) field types: this runs
// the reference's control flow with arithmetic that is known good.
#include "field768.h"
#include "field32.h"

extern "C" {

int ref_is_patched(void) {
#ifdef REF_PATCHED
    return 1;
#else
    return 0;
#endif
}

// op: 0 mul (ciosMontgomeryMultiply via Scalar::operator*), 1 add, 2 sub -- modulus is the
// reference's literal fields::_mod (MNT4-753 Fq)
void ref_scalar_binop(int op, uint32_t *out, const uint32_t *a, const uint32_t *b, size_t count) {
    for (size_t i = 0; i < count; ++i) {
        fields::Scalar x(a + i * SIZE), y(b + i * SIZE);
        fields::Scalar z = op == 0 ? x * y : op == 1 ? x + y : x - y;
        memcpy(out + i * SIZE, z.im_rep, sizeof(z.im_rep));
    }
}

void ref_scalar_pow(uint32_t *out, const uint32_t *a, uint32_t e) {
    fields::Scalar x(a);
    fields::Scalar z = x ^ e;
    memcpy(out, z.im_rep, sizeof(z.im_rep));
}

void ref_mod(uint32_t *out) { memcpy(out, fields::_mod, sizeof(fields::_mod)); }

// reference host FFT over the reference's own Scalar (what test/main.cpp:71 calls)
double ref_fft_scalar(uint32_t *limbs, size_t n, const uint32_t *omega, int log_cpus, int threads) {
    std::vector<fields::Scalar> v;
    v.reserve(n);
    for (size_t i = 0; i < n; ++i) v.push_back(fields::Scalar(limbs + i * SIZE));
    if (threads > 0) omp_set_num_threads(threads);
    auto t0 = std::chrono::steady_clock::now();
    if (log_cpus < 0) _basic_serial_radix2_FFT<fields::Scalar>(v, fields::Scalar(omega), fields::Scalar::one());
    else _basic_parallel_radix2_FFT_inner<fields::Scalar>(v, fields::Scalar(omega), (size_t)log_cpus, fields::Scalar::one());
    auto t1 = std::chrono::steady_clock::now();
    for (size_t i = 0; i < n; ++i) memcpy(limbs + i * SIZE, v[i].im_rep, sizeof(v[i].im_rep));
    return std::chrono::duration<double>(t1 - t0).count();
}

// reference host FFT templates over the oracle's field types
void ref_fft_over_oracle768(uint32_t *limbs, size_t n, const uint32_t *omega, int log_cpus, int field) {
    oracle::current_modulus() = field == 1 ? &oracle::modulus_fq() : &oracle::modulus_fr();
    std::vector<oracle::Fp768> v(n);
    for (size_t i = 0; i < n; ++i) memcpy(v[i].im_rep, limbs + i * 24, 96);
    if (log_cpus < 0) _basic_serial_radix2_FFT<oracle::Fp768>(v, oracle::Fp768(omega), oracle::Fp768::one());
    else _basic_parallel_radix2_FFT_inner<oracle::Fp768>(v, oracle::Fp768(omega), (size_t)log_cpus, oracle::Fp768::one());
    for (size_t i = 0; i < n; ++i) memcpy(limbs + i * 24, v[i].im_rep, 96);
}

void ref_fft_over_oracle32(uint32_t *a, size_t n, uint32_t omega, uint32_t mod, int log_cpus) {
    oracle::Fp32::mod() = mod;
    std::vector<oracle::Fp32> v(n);
    for (size_t i = 0; i < n; ++i) v[i].im_rep = a[i];
    if (log_cpus < 0) _basic_serial_radix2_FFT<oracle::Fp32>(v, oracle::Fp32(omega), oracle::Fp32::one());
    else _basic_parallel_radix2_FFT_inner<oracle::Fp32>(v, oracle::Fp32(omega), (size_t)log_cpus, oracle::Fp32::one());
    for (size_t i = 0; i < n; ++i) a[i] = v[i].im_rep;
}

double ref_time_fft_over_oracle32(uint32_t *a, size_t n, uint32_t omega, uint32_t mod, int log_cpus, int threads) {
    oracle::Fp32::mod() = mod;
    std::vector<oracle::Fp32> v(n);
    for (size_t i = 0; i < n; ++i) v[i].im_rep = a[i];
    if (threads > 0) omp_set_num_threads(threads);
    auto t0 = std::chrono::steady_clock::now();
    _basic_parallel_radix2_FFT_inner<oracle::Fp32>(v, oracle::Fp32(omega), (size_t)log_cpus, oracle::Fp32::one());
    auto t1 = std::chrono::steady_clock::now();
    for (size_t i = 0; i < n; ++i) a[i] = v[i].im_rep;
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
