// tests/cpp/test_host_fields.cpp -- CPU-only: the product's host field types (include/fields/field.h,
// include/fields/dummy_field.h, include/cuda/device_field.h) against the oracle's, and the host FFT
// templates (oracle/fft_host_oracle.h = reference test/fft_host.h restated) instantiated over the
// product types -- the call reference test/main.cpp:71 makes.  Exit code 0 = all good.
#include <cstdio>
#include <random>
#include <vector>

#include <cuda/device_field.h>
#include <fields/dummy_field.h>
#include <fields/field.h>

#include "fft_host_oracle.h"
#include "field32.h"
#include "field768.h"

static int fails = 0;
#define CHECK(c) do { if (!(c)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); ++fails; } } while (0)

static void rand_limbs(std::mt19937_64 &rng, uint32_t *l) {
    for (int k = 0; k < 24; ++k) l[k] = (uint32_t)rng();
    l[23] &= 0xFFFF;
}

int main() {
    std::mt19937_64 rng(42);
    for (int which = 0; which < 2; ++which) {
        cpu_fields::modulus().select(which);
        oracle::current_modulus() = which ? &oracle::modulus_fq() : &oracle::modulus_fr();
        for (int it = 0; it < 2000; ++it) {
            uint32_t a[24], b[24];
            rand_limbs(rng, a); rand_limbs(rng, b);
            cpu_fields::Field x(a), y(b);
            oracle::Fp768 ox(a), oy(b);
            CHECK(memcmp((x * y).im_rep, (ox * oy).im_rep, 96) == 0);
            CHECK(memcmp((x + y).im_rep, (ox + oy).im_rep, 96) == 0);
            CHECK(memcmp((x - y).im_rep, (ox - oy).im_rep, 96) == 0);
            CHECK(memcmp((-x).im_rep, (-ox).im_rep, 96) == 0);
            if (it < 20) {
                CHECK(memcmp((x ^ (size_t)(it * 977 + 3)).im_rep, (ox ^ (uint64_t)(it * 977 + 3)).im_rep, 96) == 0);
                cpu_fields::Field inv = x; cpu_fields::mul_inv(inv);
                CHECK(inv * x == cpu_fields::Field::one());
                fields::Scalar s(a), t(b);
                CHECK(memcmp((s * t).im_rep, (ox * oy).im_rep, 96) == 0);
                CHECK((s * fields::Scalar::one()) == s);
            }
        }
        CHECK(cpu_fields::Field::from_uint(1) == cpu_fields::Field::one());
        const int s = cpu_fields::modulus().two_adicity;
        cpu_fields::Field w = cpu_fields::Field::root_of_unity((size_t)1 << 10);
        CHECK((w ^ (size_t)1024) == cpu_fields::Field::one());
        CHECK(!((w ^ (size_t)512) == cpu_fields::Field::one()));
        (void)s;
    }
    cpu_fields::modulus().select(0);
    oracle::current_modulus() = &oracle::modulus_fr();
    {   // host FFT templates over the product types == over the oracle types
        const size_t n = 256;
        std::vector<fields::Scalar> v; std::vector<oracle::Fp768> ov;
        for (size_t i = 0; i < n; ++i) { uint32_t a[24]; rand_limbs(rng, a); v.push_back(fields::Scalar(a)); ov.push_back(oracle::Fp768(a)); }
        fields::Scalar w = fields::Scalar::root_of_unity(n);
        oracle::_basic_parallel_radix2_FFT_inner<fields::Scalar>(v, w, 2, fields::Scalar::one());
        oracle::_basic_serial_radix2_FFT<oracle::Fp768>(ov, oracle::Fp768(w.im_rep), oracle::Fp768::one());
        for (size_t i = 0; i < n; ++i) CHECK(memcmp(v[i].im_rep, ov[i].im_rep, 96) == 0);
    }
    {
        const size_t n = 4096;
        std::vector<dummy_fields::Field> v; std::vector<oracle::Fp32> ov;
        for (size_t i = 0; i < n; ++i) { uint32_t x = (uint32_t)(rng() % dummy_fields::Field::mod); v.push_back(dummy_fields::Field(x)); ov.push_back(oracle::Fp32(x)); }
        dummy_fields::Field w = dummy_fields::Field::root_of_unity(n);
        oracle::_basic_parallel_radix2_FFT_inner<dummy_fields::Field>(v, w, 3, dummy_fields::Field::one());
        oracle::_basic_serial_radix2_FFT<oracle::Fp32>(ov, oracle::Fp32(w.im_rep), oracle::Fp32::one());
        for (size_t i = 0; i < n; ++i) CHECK(v[i].im_rep == ov[i].im_rep);
        dummy_fields::Field a(5), b(7);
        dummy_fields::Field::mul(a, b); CHECK(a.im_rep == 35);
        dummy_fields::Field::mul_inv(b); CHECK((b * dummy_fields::Field(7)).im_rep == 1);
    }
    printf(fails ? "%d FAILURES\n" : "host fields ok\n", fails);
    return fails ? 1 : 0;
}
