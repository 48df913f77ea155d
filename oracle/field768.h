// oracle/field768.h -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// CPU restatement of the reference's 768-bit field type `fields::Scalar`
// (reference cuda/device_field.h:67-215) and of its arithmetic
// (reference cuda/device_field_operators.h:92-258), with the defects that SURVEY.md
// section 0 documents put right:
//   * F2  the Montgomery constant is the real  n' = -p^-1 mod 2^32  (the reference
//         hard-codes 0xFFFFFFFF, device_field_operators.h:47);
//   * the final normalisation compares all 24 words (the reference copies 24 *bytes*,
//         device_field_operators.h:142,147, so its conditional subtract is dead);
//   * add() reduces a sum that equals p to 0 (reference uses a strict compare, :193);
//   * operator^ is a true exponentiation whose result lives in the same (Montgomery)
//         domain for every exponent (reference: linear-time, domain depends on e, :216-258);
//   * one() is the identity of operator* (R mod p), not the raw integer 1 (:86-91).
// Fully reduced modular results are unique, so any correct implementation produces the
// same bits: this file is the bit-exact judge for the CUDA path.
//
// The modulus is selectable at run time (MNT4-753 Fr, the product's field, or MNT4-753
// Fq, the reference's literal `_mod`) so that the same code can be pinned against the
// reference compiled from /root/reference (oracle/_ref, see oracle/Makefile).
#pragma once
#include <cstdint>
#include <cstring>
#include <cstddef>
#include "../include/gsn_constants.h"

namespace oracle {

constexpr int NL = 24;  // reference: #define SIZE (768 / 32), device_field.h:35

struct Modulus768 {
    uint32_t p[NL];
    uint32_t r1[NL];   // R mod p
    uint32_t r2[NL];   // R^2 mod p
    uint32_t root[NL]; // primitive 2^s-th root of unity, Montgomery form
    uint32_t np0;      // -p^-1 mod 2^32
    int two_adicity;
};

inline const Modulus768 &modulus_fr() {
    static const Modulus768 m = {GSN_FR_MOD, GSN_FR_R1, GSN_FR_R2, GSN_FR_ROOT_MONT, GSN_FR_NP0, GSN_FR_TWO_ADICITY};
    return m;
}
inline const Modulus768 &modulus_fq() {
    static const Modulus768 m = {GSN_FQ_MOD, GSN_FQ_R1, GSN_FQ_R2, GSN_FQ_ROOT_MONT, GSN_FQ_NP0, GSN_FQ_TWO_ADICITY};
    return m;
}

// The reference keeps its modulus in a namespace-scope array (`fields::_mod`,
// device_field.h:62-65); the oracle keeps a pointer so tests can switch fields.
inline const Modulus768 *&current_modulus() {
    static const Modulus768 *cur = &modulus_fr();
    return cur;
}

// reference `less` (device_field_operators.h:92-102): true iff a < b
inline bool less(const uint32_t *a, const uint32_t *b) {
    for (int i = NL - 1; i >= 0; --i) {
        if (a[i] != b[i]) return a[i] < b[i];
    }
    return false;
}

// reference `_add` (device_field_operators.h:105-118): a += b, returns carry out
inline uint32_t raw_add(uint32_t *a, const uint32_t *b) {
    uint64_t carry = 0;
    for (int i = 0; i < NL; ++i) {
        uint64_t t = (uint64_t)a[i] + b[i] + carry;
        a[i] = (uint32_t)t;
        carry = t >> 32;
    }
    return (uint32_t)carry;
}

// reference `_subtract` (device_field_operators.h:121-137): a -= b, returns borrow out
inline uint32_t raw_sub(uint32_t *a, const uint32_t *b) {
    uint64_t borrow = 0;
    for (int i = 0; i < NL; ++i) {
        uint64_t t = (uint64_t)a[i] - b[i] - borrow;
        a[i] = (uint32_t)t;
        borrow = (t >> 32) & 1;
    }
    return (uint32_t)borrow;
}

// reference `ciosMontgomeryMultiply` + `montyNormalize` (device_field_operators.h:139-187):
// result = a*b*2^-768 mod p, canonical.  Same loop nest (outer over b's words, inner-1
// accumulates a*b[i], inner-2 adds m*p and shifts one word), correct constant, real normalise.
inline void cios_mul(uint32_t *result, const uint32_t *a, const uint32_t *b, const Modulus768 &M) {
    uint32_t t[NL + 2];
    memset(t, 0, sizeof(t));
    for (int i = 0; i < NL; ++i) {
        uint64_t carry = 0;
        for (int j = 0; j < NL; ++j) {
            uint64_t cur = (uint64_t)t[j] + (uint64_t)a[j] * b[i] + carry;
            t[j] = (uint32_t)cur;
            carry = cur >> 32;
        }
        uint64_t cur = (uint64_t)t[NL] + carry;
        t[NL] = (uint32_t)cur;
        t[NL + 1] = (uint32_t)(cur >> 32);

        uint32_t m = t[0] * M.np0;
        cur = (uint64_t)t[0] + (uint64_t)m * M.p[0];
        carry = cur >> 32;
        for (int j = 1; j < NL; ++j) {
            cur = (uint64_t)t[j] + (uint64_t)m * M.p[j] + carry;
            t[j - 1] = (uint32_t)cur;
            carry = cur >> 32;
        }
        cur = (uint64_t)t[NL] + carry;
        t[NL - 1] = (uint32_t)cur;
        t[NL] = t[NL + 1] + (uint32_t)(cur >> 32);
    }
    // t < 2p here; subtract p once if t >= p (t[NL] can only be set when t >= 2^768 > p)
    uint32_t u[NL];
    memcpy(u, t, sizeof(u));
    uint32_t borrow = raw_sub(u, M.p);
    if (t[NL] != 0 || !borrow) memcpy(result, u, sizeof(u));
    else memcpy(result, t, sizeof(u));
}

struct Fp768 {
    uint32_t im_rep[NL];  // little-endian limbs, reference device_field.h:75

    static Fp768 zero() { Fp768 r; memset(r.im_rep, 0, sizeof(r.im_rep)); return r; }
    static Fp768 one() { Fp768 r; memcpy(r.im_rep, current_modulus()->r1, sizeof(r.im_rep)); return r; }
    Fp768() { memset(im_rep, 0, sizeof(im_rep)); }
    // raw limb constructors, exactly as the reference (device_field.h:95-104): no domain conversion
    explicit Fp768(uint32_t v) { memset(im_rep, 0, sizeof(im_rep)); im_rep[0] = v; }
    explicit Fp768(const uint32_t *v) { memcpy(im_rep, v, sizeof(im_rep)); }

    bool is_zero() const { for (int i = 0; i < NL; ++i) if (im_rep[i]) return false; return true; }
    bool operator==(const Fp768 &o) const { return memcmp(im_rep, o.im_rep, sizeof(im_rep)) == 0; }
    bool operator!=(const Fp768 &o) const { return !(*this == o); }

    // reference Scalar::mul (device_field_operators.h:207-214)
    Fp768 operator*(const Fp768 &o) const { Fp768 r; cios_mul(r.im_rep, im_rep, o.im_rep, *current_modulus()); return r; }
    Fp768 square() const { return *this * *this; }
    // reference Scalar::add (device_field_operators.h:190-195)
    Fp768 operator+(const Fp768 &o) const {
        Fp768 r = *this;
        uint32_t carry = raw_add(r.im_rep, o.im_rep);
        if (carry || !less(r.im_rep, current_modulus()->p)) raw_sub(r.im_rep, current_modulus()->p);
        return r;
    }
    // reference Scalar::subtract (device_field_operators.h:198-204)
    Fp768 operator-(const Fp768 &o) const {
        Fp768 r = *this;
        if (less(r.im_rep, o.im_rep)) raw_add(r.im_rep, current_modulus()->p);
        raw_sub(r.im_rep, o.im_rep);
        return r;
    }
    Fp768 operator-() const { return zero() - *this; }
    // reference Scalar::pow / operator^ (device_field_operators.h:229-258): here a real
    // square-and-multiply so that x^e stays in the Montgomery domain for every e (e = 0 -> one()).
    Fp768 operator^(uint64_t e) const {
        Fp768 acc = one(), base = *this;
        while (e) {
            if (e & 1) acc = acc * base;
            base = base * base;
            e >>= 1;
        }
        return acc;
    }
    Fp768 to_monty() const { Fp768 r2(current_modulus()->r2); return *this * r2; }
    Fp768 from_monty() const { return *this * Fp768(1u); }
    // x^(p-2)
    Fp768 inverse() const {
        uint32_t e[NL];
        memcpy(e, current_modulus()->p, sizeof(e));
        uint32_t two[NL] = {2};
        raw_sub(e, two);
        Fp768 acc = one();
        for (int bit = 32 * NL - 1; bit >= 0; --bit) {
            acc = acc * acc;
            if ((e[bit / 32] >> (bit % 32)) & 1) acc = acc * *this;
        }
        return acc;
    }
};

}  // namespace oracle
