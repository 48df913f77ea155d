// tests/cpp/device_field_operator_test.cpp -- the reference's arithmetic unit test
// (reference cuda/device_field_operator_test.cpp) without GMP: the judge is this repo's host field type
// (include/fields/field.h, itself checked against the oracle and Python big-ints on the CPU), the subject
// is the DEVICE arithmetic reached through the C ABI (gsn_fp768_binop_host, gsn_fp768_inner_product_host).
//   testAdd / test_subtract / testMultiply / testPow KATs   <- reference :222-299
//   testEncodeDecode, testConstructor                       <- reference :183-205, :301-320
//   fuzzTest: add / sub / mul on a grid of operands         <- reference :442-483, with full-width random
//             operands instead of the reference's 32-bit ones and the mismatch assert switched ON (:430-437)
// usage: device_field_operator_test [n_fuzz]      exit code 0 = all good
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include <cuda/device_field.h>
#include <gpusnarks_b200.h>

using fields::Scalar;
static int fails = 0;
#define CHECK(c) do { if (!(c)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); ++fails; } } while (0)

static gsn_ctx *ctx;
static Scalar dev(int op, const Scalar &a, const Scalar &b) {
    Scalar r;
    if (gsn_fp768_binop_host(ctx, op, r.im_rep, a.im_rep, b.im_rep, 1) != GSN_OK) { printf("C ABI error: %s\n", gsn_last_error()); exit(2); }
    return r;
}
static Scalar to_monty(const Scalar &a) { return dev(0, a, Scalar(cpu_fields::modulus().r2)); }
static Scalar from_monty(const Scalar &a) { return dev(0, a, Scalar(1u)); }

static void testAdd() { CHECK(dev(1, Scalar(1234), Scalar(1234)) == Scalar(2468)); }
static void test_subtract() {
    CHECK(dev(2, Scalar(1234), Scalar(1234)) == Scalar::zero());
    CHECK(dev(2, Scalar(1235), Scalar(1234)) == Scalar(1u));
    CHECK(dev(2, Scalar::zero(), Scalar(1u)) == -Scalar(1u));  // wraps to p - 1
}
static void testMultiply() {
    const Scalar m = to_monty(Scalar(1234));
    CHECK(from_monty(dev(0, m, m)) == Scalar(1522756));
    CHECK(dev(0, m, Scalar::one()) == m);  // one() is the identity of the Montgomery product
}
static void testPow() {  // 2^0, 2^2, 4^10, 2^20, 2^35 by repeated device multiplication
    const struct { uint32_t base; unsigned e; unsigned long long expect; } kats[] = {{2, 0, 1ull}, {2, 2, 4ull}, {4, 10, 1048576ull}, {2, 20, 1048576ull}, {2, 35, 34359738368ull}};
    for (auto &k : kats) {
        Scalar acc = Scalar::one(), b = to_monty(Scalar(k.base));
        for (unsigned i = 0; i < k.e; ++i) acc = dev(0, acc, b);
        Scalar r = from_monty(acc);
        CHECK(r.im_rep[0] == (uint32_t)k.expect && r.im_rep[1] == (uint32_t)(k.expect >> 32));
        for (int i = 2; i < SIZE; ++i) CHECK(r.im_rep[i] == 0);
    }
}
static void testConstructorAndEncodeDecode() {
    uint32_t limbs[SIZE];
    for (int i = 0; i < SIZE; ++i) limbs[i] = 0x01010101u * (i + 1);
    limbs[SIZE - 1] &= 0xFFFF;
    Scalar s(limbs);
    CHECK(memcmp(s.im_rep, limbs, sizeof(limbs)) == 0);
    CHECK(from_monty(to_monty(s)) == s);  // limb pattern survives the round trip through the device
    CHECK(Scalar(7u).im_rep[0] == 7 && Scalar(7u).im_rep[1] == 0);
}
static void fuzzTest(size_t n) {
    std::mt19937_64 rng(20240917);
    std::vector<Scalar> a(n), b(n), r(n);
    for (size_t i = 0; i < n; ++i) {
        for (int k = 0; k < SIZE; ++k) { a[i].im_rep[k] = (uint32_t)rng(); b[i].im_rep[k] = (uint32_t)rng(); }
        a[i].im_rep[SIZE - 1] &= 0xFFFF;  // < 2^752 < p
        b[i].im_rep[SIZE - 1] &= 0xFFFF;
        if (i % 97 == 0) b[i] = a[i];               // a - a, a + a, a * a
        if (i % 101 == 0) a[i] = -Scalar(1u);       // p - 1
    }
    for (int op = 0; op < 3; ++op) {
        if (gsn_fp768_binop_host(ctx, op, r[0].im_rep, a[0].im_rep, b[0].im_rep, n) != GSN_OK) { printf("C ABI error: %s\n", gsn_last_error()); exit(2); }
        for (size_t i = 0; i < n; ++i) {
            const Scalar expect = op == 0 ? a[i] * b[i] : op == 1 ? a[i] + b[i] : a[i] - b[i];
            if (!(r[i] == expect)) { if (fails < 5) { printf("op %d mismatch at %zu\n", op, i); Scalar::print(r[i]); Scalar::print(expect); } ++fails; }
        }
    }
    // multiexp<Scalar, Scalar>: sum a[i] * b[i]  (reference test/multiexp.h:3-13)
    Scalar acc = Scalar::zero(), got;
    for (size_t i = 0; i < n; ++i) acc = acc + a[i] * b[i];
    if (gsn_fp768_inner_product_host(ctx, got.im_rep, a[0].im_rep, b[0].im_rep, n) != GSN_OK) { printf("C ABI error: %s\n", gsn_last_error()); exit(2); }
    CHECK(got == acc);
}

int main(int argc, char **argv) {
    const size_t n = argc > 1 ? (size_t)atol(argv[1]) : 20000;
    if (gsn_ctx_create(&ctx, 0) != GSN_OK) { printf("gsn_ctx_create: %s\n", gsn_last_error()); return 2; }
    testAdd();
    test_subtract();
    testMultiply();
    testPow();
    testConstructorAndEncodeDecode();
    fuzzTest(n);
    gsn_ctx_destroy(ctx);
    printf(fails ? "%d FAILURES\n" : "device field operators ok\n", fails);
    return fails ? 1 : 0;
}
