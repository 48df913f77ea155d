// cuda/device_field.h -- `fields::Scalar`, same header path and surface as the reference's
// working 768-bit type (reference cuda/device_field.h:67-215): im_rep[SIZE], zero(), one(),
// constructors from uint32_t / const uint32_t*, is_zero, operators * + - unary- ^ ==,
// square(), print(), testEquality().  Host-only here: the reference compiles these methods
// for host and device from one source (device_field.h:26); in this library the device side is
// gpusnarks_b200/csrc/fp768.cuh behind the C ABI, and this header is what host callers such
// as test/main.cpp include.  Differences, all corrections:
//   * arithmetic is a real Montgomery field (correct n', real final reduction);
//   * one() is the identity of operator* (R mod p);   * operator^ is a true power;
//   * unary minus negates (the reference returns x - x = 0, device_field.h:142-149);
//   * the modulus is MNT4-753 Fr; fields::_mod holds ITS limbs (the reference's array holds
//     MNT4-753 Fq, whose 2-adicity 15 admits no 2^16-point transform).
#ifndef GSN_CUDA_DEVICE_FIELD_H
#define GSN_CUDA_DEVICE_FIELD_H
#include <cassert>
#include <cstdint>
#include <cstdio>

#include "../fields/field.h"

namespace fields {

using size_t = decltype(sizeof 1ll);

const uint32_t _mod[SIZE] = GSN_FR_MOD;

struct Scalar {
    uint32_t im_rep[SIZE] = {0};

    static Scalar zero() { return Scalar(); }
    static Scalar one() { return Scalar(cpu_fields::modulus().r1); }
    Scalar() = default;
    Scalar(const uint32_t value) { im_rep[0] = value; }
    Scalar(const uint32_t *value) { for (size_t i = 0; i < SIZE; i++) im_rep[i] = value[i]; }
    static Scalar root_of_unity(size_t n) { return Scalar(cpu_fields::Field::root_of_unity(n).im_rep); }

    bool is_zero() const { for (size_t i = 0; i < SIZE; i++) if (im_rep[i] != 0) return false; return true; }

    Scalar operator*(const Scalar &rhs) const { cpu_fields::Field a(im_rep); cpu_fields::mul(a, cpu_fields::Field(rhs.im_rep)); return Scalar(a.im_rep); }
    Scalar operator+(const Scalar &rhs) const { cpu_fields::Field a(im_rep); cpu_fields::add(a, cpu_fields::Field(rhs.im_rep)); return Scalar(a.im_rep); }
    Scalar operator-(const Scalar &rhs) const { cpu_fields::Field a(im_rep); cpu_fields::subtract(a, cpu_fields::Field(rhs.im_rep)); return Scalar(a.im_rep); }
    Scalar operator-() const { cpu_fields::Field a(im_rep); cpu_fields::negate(a); return Scalar(a.im_rep); }
    Scalar operator^(const size_t &rhs) const { cpu_fields::Field a(im_rep); cpu_fields::pow(a, rhs); return Scalar(a.im_rep); }
    bool operator==(const Scalar &rhs) const { for (size_t i = 0; i < SIZE; i++) if (rhs.im_rep[i] != im_rep[i]) return false; return true; }
    bool operator!=(const Scalar &rhs) const { return !(*this == rhs); }
    Scalar square() const { return *this * *this; }
    Scalar inverse() const { cpu_fields::Field a(im_rep); cpu_fields::mul_inv(a); return Scalar(a.im_rep); }

    static void print(Scalar f) { for (size_t i = 0; i < SIZE; i++) printf("%u, ", f.im_rep[i]); printf("\n"); }
    static void testEquality(Scalar f1, Scalar f2) {
        for (size_t i = 0; i < SIZE; i++)
            if (f1.im_rep[i] != f2.im_rep[i]) {
                printf("Missmatch: \n");
                print(f1);
                print(f2);
                assert(!"missmatch");
            }
    }
};

static_assert(sizeof(Scalar) == 96, "fields::Scalar must be 24 raw limbs (the C ABI memcpy's it, like reference fft_kernel.cu:131)");

}  // namespace fields
#endif
