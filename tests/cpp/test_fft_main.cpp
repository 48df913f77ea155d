// tests/cpp/test_fft_main.cpp -- the reference's FFT test driver shape (reference
// test/main.cpp:34-87, `test_fft`): fill two vectors, run best_fft on one (device) and the
// host FFT on the other, compare element by element.  Differences from the reference driver:
// random inputs (the reference uses the constant 1234), a real root of unity (the reference
// passes Scalar(123) to the device and the modulus to the host), both field types, and the
// host side is the oracle restatement of test/fft_host.h (oracle/fft_host_oracle.h) running
// over THIS repo's host field types -- i.e. exactly the call test/main.cpp:71 makes.
//   usage: test_fft_main [log2_n_768] [log2_n_32]
#include <omp.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include <cuda/device_field.h>
#include <cuda/fft_kernel.h>
#include <fields/dummy_field.h>

#include "fft_host_oracle.h"

typedef std::chrono::high_resolution_clock Clock;

template <typename F>
static long ms_since(const F &t1) { return std::chrono::duration_cast<std::chrono::milliseconds>(Clock::now() - t1).count(); }

static int test_fft768(size_t log_n) {
    printf("\nTEST FFT (768-bit, MNT4-753 Fr), 2^%zu\n", log_n);
    const size_t _size = (size_t)1 << log_n;
    std::vector<fields::Scalar> v1, v2;
    v1.reserve(_size);
    v2.reserve(_size);
    std::mt19937_64 rng(1);
    for (size_t i = 0; i < _size; i++) {
        uint32_t limbs[SIZE];
        for (int k = 0; k < SIZE; ++k) limbs[k] = (uint32_t)rng();
        limbs[SIZE - 1] &= 0xFFFF;  // < 2^752 < r
        v1.push_back(fields::Scalar(limbs));
        v2.push_back(fields::Scalar(limbs));
    }
    const fields::Scalar omega = fields::Scalar::root_of_unity(_size);
    omp_set_num_threads(8);
    printf("Field size: %lu, Field count: %lu\n", sizeof(fields::Scalar), v1.size());
    auto t1 = Clock::now();
    best_fft<fields::Scalar>(v1, omega);
    printf("Device FFT took %ld \n", ms_since(t1));
    t1 = Clock::now();
    oracle::_basic_parallel_radix2_FFT_inner<fields::Scalar>(v2, omega, 3, fields::Scalar::one());
    printf("Host FFT took %ld \n", ms_since(t1));
    for (size_t i = 0; i < _size; i++) fields::Scalar::testEquality(v1[i], v2[i]);
    if (!(v1 == v2)) return 1;
    // batch entry point: two copies of the spectrum back through the inverse, copies overlapped
    std::vector<std::vector<fields::Scalar>> batch = {v1, v1};
    best_fft_batch(batch, omega, /*inverse=*/true);
    best_ifft<fields::Scalar>(v1, omega);
    if (!(batch[0] == v1) || !(batch[1] == v1)) { printf("BATCH MISMATCH\n"); return 1; }
    printf("forward == host FFT, batch inverse == inverse, DONE\n");
    return 0;
}

static int test_fft32(size_t log_n) {
    printf("\nTEST FFT (32-bit, p = %u), 2^%zu\n", dummy_fields::Field::mod, log_n);
    const size_t _size = (size_t)1 << log_n;
    std::vector<dummy_fields::Field> v1, v2, v0;
    std::mt19937_64 rng(1);
    for (size_t i = 0; i < _size; i++) {
        dummy_fields::Field x((uint32_t)(rng() % dummy_fields::Field::mod));
        v1.push_back(x);
        v2.push_back(x);
    }
    v0 = v1;
    const dummy_fields::Field omega = dummy_fields::Field::root_of_unity(_size);
    auto t1 = Clock::now();
    best_fft<dummy_fields::Field>(v1, omega);
    printf("Device FFT took %ld \n", ms_since(t1));
    t1 = Clock::now();
    oracle::_basic_parallel_radix2_FFT_inner<dummy_fields::Field>(v2, omega, 3, dummy_fields::Field::one());
    printf("Host FFT took %ld \n", ms_since(t1));
    if (!(v1 == v2)) { printf("MISMATCH\n"); return 1; }
    best_ifft<dummy_fields::Field>(v1, omega);
    if (!(v1 == v0)) { printf("INVERSE MISMATCH\n"); return 1; }
    printf("forward == host FFT, inverse(forward) == input, DONE\n");
    return 0;
}

// cpu_fields::Field over the reference's literal modulus (MNT4-753 Fq, 2-adicity 15): the shim must follow the host-side
// modulus selection and use a context of that field
static int test_fft768_fq(size_t log_n) {
    printf("\nTEST FFT (768-bit, cpu_fields::Field over MNT4-753 Fq), 2^%zu\n", log_n);
    cpu_fields::modulus().select(1);
    const size_t _size = (size_t)1 << log_n;
    std::vector<cpu_fields::Field> v1, v2;
    std::mt19937_64 rng(2);
    for (size_t i = 0; i < _size; i++) {
        uint32_t limbs[SIZE];
        for (int k = 0; k < SIZE; ++k) limbs[k] = (uint32_t)rng();
        limbs[SIZE - 1] &= 0xFFFF;
        v1.push_back(cpu_fields::Field(limbs));
        v2.push_back(cpu_fields::Field(limbs));
    }
    const cpu_fields::Field omega = cpu_fields::Field::root_of_unity(_size);
    best_fft<cpu_fields::Field>(v1, omega);
    oracle::_basic_parallel_radix2_FFT_inner<cpu_fields::Field>(v2, omega, 3, cpu_fields::Field::one());
    cpu_fields::modulus().select(0);
    if (!(v1 == v2)) { printf("FQ MISMATCH\n"); return 1; }
    printf("forward == host FFT over Fq, DONE\n");
    return 0;
}

int main(int argc, char **argv) {
    const size_t l768 = argc > 1 ? (size_t)atoi(argv[1]) : 16;  // the reference's shape: 1 << 16
    const size_t l32 = argc > 2 ? (size_t)atoi(argv[2]) : 16;
    try {
        if (test_fft768(l768)) return 1;
        if (test_fft32(l32)) return 1;
        if (test_fft768_fq(12)) return 1;
    } catch (const std::exception &e) {
        printf("ERROR: %s\n", e.what());
        return 2;
    }
    return 0;
}
