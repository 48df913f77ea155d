// tests/cpp/test_multiexp_main.cpp -- the reference's multi-exponentiation test drivers (reference test/main.cpp:89-179,
// `test_multiexp` and `test_multiexp_mnt4753_G1`) against this repo's headers: device multiexp<> through
// cuda/multi_exp.h and the C ABI, host side = the reference's own loop (test/multiexp.h:3-13: result = result + a[i]*b[i])
// over this repo's host types.  Differences from the reference driver: random field elements instead of the constant
// 1234; curve points that ARE on the curve (read from a file the pytest wrapper generates: the reference's
// (1234, 1234, 1234) is not a point of MNT4-753 G1); sizes that a host double-and-add finishes in seconds.
//   usage: test_multiexp_main POINTS_FILE N_POINTS [LOG_N_SCALAR]
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include <cuda/device_field.h>
#include <cuda/multi_exp.h>

// reference test/multiexp.h:3-13
template <typename FieldT, typename FieldMul>
FieldT multi_exp(std::vector<FieldT> &a, std::vector<FieldMul> &b) {
    FieldT result = FieldT::zero();
    for (size_t i = 0; i < a.size(); i++) result = result + (a[i] * b[i]);
    return result;
}

static fields::Scalar random_scalar(std::mt19937_64 &rng) {
    uint32_t limbs[SIZE];
    for (int k = 0; k < SIZE; ++k) limbs[k] = (uint32_t)rng();
    limbs[SIZE - 1] &= 0xFFFF;  // < 2^752
    return fields::Scalar(limbs);
}

int main(int argc, char **argv) {
    if (argc < 3) { printf("usage: %s POINTS_FILE N_POINTS [LOG_N_SCALAR]\n", argv[0]); return 2; }
    const size_t npts = (size_t)atoi(argv[2]);
    const size_t log_n = argc > 3 ? (size_t)atoi(argv[3]) : 12;
    std::mt19937_64 rng(5);
    try {
        {
            printf("\nTEST MULTI_EXP\n");
            const size_t _size = (size_t)1 << log_n;
            std::vector<fields::Scalar> v1, v2;
            for (size_t i = 0; i < _size; i++) { v1.push_back(random_scalar(rng)); v2.push_back(random_scalar(rng)); }
            std::vector<fields::Scalar> v3 = v1, v4 = v2;
            printf("Field size: %lu, Field count: %lu\n", sizeof(fields::Scalar), v1.size());
            fields::Scalar gpuResult = multiexp<fields::Scalar, fields::Scalar>(v1, v2);
            fields::Scalar cpuResult = multi_exp<fields::Scalar, fields::Scalar>(v3, v4);
            fields::Scalar::testEquality(cpuResult, gpuResult);
            if (!(cpuResult == gpuResult)) return 1;
            printf("\nDONE\n");
        }
        {
            printf("\nTEST MULTI_EXP_MNT4753\n");
            std::vector<fields::mnt4753_G1> v1(npts);
            FILE *f = fopen(argv[1], "rb");
            if (!f || fread(v1.data(), sizeof(fields::mnt4753_G1), npts, f) != npts) { printf("cannot read %zu points from %s\n", npts, argv[1]); return 2; }
            fclose(f);
            std::vector<fields::Scalar> v2;
            for (size_t i = 0; i < npts; i++) v2.push_back(random_scalar(rng));
            v2[0] = fields::Scalar(0u);
            if (npts > 2) { v1[2] = v1[1]; v2[2] = v2[1]; }   // equal summands: a doubling inside the sum
            std::vector<fields::mnt4753_G1> v3 = v1;
            std::vector<fields::Scalar> v4 = v2;
            printf("Field size: %lu, Field count: %lu\n", sizeof(fields::mnt4753_G1), v1.size());
            fields::mnt4753_G1 gpuResult = multiexp<fields::mnt4753_G1, fields::Scalar>(v1, v2);
            fields::mnt4753_G1 cpuResult = multi_exp<fields::mnt4753_G1, fields::Scalar>(v3, v4);
            fields::mnt4753_G1::testEquality(cpuResult, gpuResult);
            if (!cpuResult.same_point(gpuResult) || fields::mnt4753_G1::is_zero(gpuResult)) return 1;
            printf("\nDONE\n");
        }
    } catch (const std::exception &e) {
        printf("ERROR: %s\n", e.what());
        return 2;
    }
    return 0;
}
