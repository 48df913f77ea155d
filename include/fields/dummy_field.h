// fields/dummy_field.h -- host field type for the 32-bit field, in the reference's vocabulary.
//
// Same class shape as the reference's `dummy_fields::Field` (reference
// fields/dummy_field.h:24-62): a uint32_t `im_rep`, static `mod` and `n`, static methods
// zero one is_zero square negate add subtract mul mul_inv pow.  The reference leaves mod = 0
// and implements C wrap-around arithmetic with a no-op pow (fields/dummy_field.cpp:25,50-107),
// so it is not a field; here `mod` is an NTT-friendly prime (default 2013265921 = 15*2^27+1,
// settable) and every method reduces.  Operators * + - ^ == are added so that the host FFT
// templates (test/fft_host.h) instantiate, which the reference's class cannot.
#ifndef GSN_FIELDS_DUMMY_FIELD_H
#define GSN_FIELDS_DUMMY_FIELD_H
#include <cstdint>

#include "../gsn_constants.h"

namespace dummy_fields {

using size_t = decltype(sizeof 1ll);

class Field {
public:
    // Intermediate representation: plain residue in [0, mod)
    uint32_t im_rep;
    // Modulo (odd prime < 2^31) and its multiplicative generator
    static inline uint32_t mod = GSN_P32_DEFAULT;
    static inline uint32_t gen = GSN_P32_DEFAULT_GEN;
    // N: transform length the field was last asked a root for (kept for API parity; unused)
    static inline uint32_t n = 0;

    Field() = default;
    Field(uint32_t value) { im_rep = value; }

    static Field zero() { return Field(0); }
    static Field one() { return Field(1); }
    static bool is_zero(const Field &fld) { return fld.im_rep == 0; }
    static void square(Field &fld) { mul(fld, fld); }
    static void negate(Field &fld) { fld.im_rep = fld.im_rep ? mod - fld.im_rep : 0; }
    static void add(Field &fld1, const Field &fld2) {
        uint32_t s = fld1.im_rep + fld2.im_rep;  // < 2^32 because mod < 2^31
        fld1.im_rep = s >= mod ? s - mod : s;
    }
    static void subtract(Field &fld1, const Field &fld2) {
        fld1.im_rep = fld1.im_rep >= fld2.im_rep ? fld1.im_rep - fld2.im_rep : fld1.im_rep + mod - fld2.im_rep;
    }
    static void mul(Field &fld1, const Field &fld2) { fld1.im_rep = (uint32_t)((uint64_t)fld1.im_rep * fld2.im_rep % mod); }
    static void pow(Field &fld1, const size_t e) {
        Field acc = one(), base = fld1;
        for (size_t k = e; k; k >>= 1) { if (k & 1) mul(acc, base); mul(base, base); }
        fld1 = acc;
    }
    static void mul_inv(Field &fld1) { pow(fld1, (size_t)mod - 2); }
    // primitive len-th root of unity, len a power of two dividing mod - 1
    static Field root_of_unity(size_t len) { n = (uint32_t)len; Field g(gen); pow(g, (size_t)(mod - 1) / len); return g; }

    Field operator*(const Field &o) const { Field r = *this; mul(r, o); return r; }
    Field operator+(const Field &o) const { Field r = *this; add(r, o); return r; }
    Field operator-(const Field &o) const { Field r = *this; subtract(r, o); return r; }
    Field operator-() const { Field r = *this; negate(r); return r; }
    Field operator^(const size_t e) const { Field r = *this; pow(r, e); return r; }
    bool operator==(const Field &o) const { return im_rep == o.im_rep; }
    bool operator!=(const Field &o) const { return im_rep != o.im_rep; }
};

}  // namespace dummy_fields
#endif
