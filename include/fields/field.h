// fields/field.h -- host field type for the 768-bit field, in the reference's vocabulary.
//
// The reference's fields/field.h (reference fields/field.h:25-223) sketches a multi-limb
// `cpu_fields::Field` with free functions  zero() one() Field(uint32_t) == is_zero add
// subtract mul square negate mul_inv pow;  its arithmetic is unfinished (TODOs at :93,169,209,
// an out-of-bounds multiply at :124-139, no modulus) and it lacks the operators that
// test/fft_host.h needs.  This header keeps that vocabulary -- same names, same
// in-place `void f(Field&, const Field&)` shape -- over the field the north-star names,
// MNT4-753 Fr (24 little-endian 32-bit limbs, Montgomery form, the layout of the reference's
// working type fields::Scalar, cuda/device_field.h:75), with complete arithmetic, and adds
// the operator surface (* + - ^ ==) so that the host FFT templates instantiate.
//
// Host-only, header-only, no CUDA: the device arithmetic lives in
// gpusnarks_b200/csrc/fp768.cuh and is reached through the C ABI (gpusnarks_b200.h).
#ifndef GSN_FIELDS_FIELD_H
#define GSN_FIELDS_FIELD_H
#include <cstdint>
#include <cstring>
#include <cstdio>

#include "../gsn_constants.h"
#include "fp768_host.h"

#ifndef SIZE
#define SIZE (768 / 32)  // reference cuda/device_field.h:35
#endif

namespace cpu_fields {

using size_t = decltype(sizeof 1ll);

// modulus selection: 0 = MNT4-753 Fr (default), 1 = MNT4-753 Fq (the reference's literal _mod)
struct ModulusState {
    gsn::host::Field768 f;
    uint32_t p[24], r1[24], r2[24], root[24];
    int two_adicity;
    int which;
    ModulusState() { select(0); }
    void select(int w) {
        static const uint32_t fr_p[24] = GSN_FR_MOD, fr_r1[24] = GSN_FR_R1, fr_r2[24] = GSN_FR_R2, fr_root[24] = GSN_FR_ROOT_MONT;
        static const uint32_t fq_p[24] = GSN_FQ_MOD, fq_r1[24] = GSN_FQ_R1, fq_r2[24] = GSN_FQ_R2, fq_root[24] = GSN_FQ_ROOT_MONT;
        which = w;
        memcpy(p, w ? fq_p : fr_p, 96);
        memcpy(r1, w ? fq_r1 : fr_r1, 96);
        memcpy(r2, w ? fq_r2 : fr_r2, 96);
        memcpy(root, w ? fq_root : fr_root, 96);
        two_adicity = w ? GSN_FQ_TWO_ADICITY : GSN_FR_TWO_ADICITY;
        f.init(p, r1);
    }
};
inline ModulusState &modulus() { static ModulusState m; return m; }

struct Field {
    // Intermediate representation: little-endian limbs, Montgomery form
    uint32_t im_rep[SIZE];

    static Field zero() { Field r; memset(r.im_rep, 0, sizeof(r.im_rep)); return r; }
    // multiplicative identity of mul()/operator*: R mod p
    static Field one() { Field r; memcpy(r.im_rep, modulus().r1, sizeof(r.im_rep)); return r; }
    Field() { memset(im_rep, 0, sizeof(im_rep)); }
    // raw limb constructors, as the reference's (no domain conversion)
    Field(uint32_t value) { memset(im_rep, 0, sizeof(im_rep)); im_rep[0] = value; }
    Field(const uint32_t *value) { memcpy(im_rep, value, sizeof(im_rep)); }
    // the field element with integer value v (converted to Montgomery form)
    static Field from_uint(uint32_t v);
    // primitive n-th root of unity, n a power of two <= 2^two_adicity
    static Field root_of_unity(size_t n);
};

inline bool operator==(const Field &lhs, const Field &rhs) { return memcmp(lhs.im_rep, rhs.im_rep, sizeof(lhs.im_rep)) == 0; }
inline bool operator!=(const Field &lhs, const Field &rhs) { return !(lhs == rhs); }
inline bool is_zero(const Field &fld) { return fld == Field::zero(); }

namespace detail {
inline void to64(uint64_t *d, const Field &f) { memcpy(d, f.im_rep, 96); }
inline void from64(Field &f, const uint64_t *s) { memcpy(f.im_rep, s, 96); }
}  // namespace detail

// Adds two elements: fld1 += fld2 (mod p)
inline void add(Field &fld1, const Field &fld2) {
    uint64_t a[13], b[12];
    detail::to64(a, fld1); detail::to64(b, fld2);
    unsigned __int128 c = 0;
    for (int i = 0; i < 12; ++i) { c += (unsigned __int128)a[i] + b[i]; a[i] = (uint64_t)c; c >>= 64; }
    const uint64_t *p = modulus().f.p;
    if (c || gsn::host::Field768::geq(a, p)) gsn::host::Field768::sub_n(a, a, p);
    detail::from64(fld1, a);
}
// Subtract element two from element one: fld1 -= fld2 (mod p)
inline void subtract(Field &fld1, const Field &fld2) {
    uint64_t a[12], b[12];
    detail::to64(a, fld1); detail::to64(b, fld2);
    if (gsn::host::Field768::sub_n(a, a, b)) {
        const uint64_t *p = modulus().f.p;
        unsigned __int128 c = 0;
        for (int i = 0; i < 12; ++i) { c += (unsigned __int128)a[i] + p[i]; a[i] = (uint64_t)c; c >>= 64; }
    }
    detail::from64(fld1, a);
}
// Multiply two elements: fld1 = fld1 * fld2 * R^-1 (Montgomery product)
inline void mul(Field &fld1, const Field &fld2) {
    uint64_t a[12], b[12];
    detail::to64(a, fld1); detail::to64(b, fld2);
    modulus().f.mul(a, a, b);
    detail::from64(fld1, a);
}
inline void square(Field &fld) { Field t = fld; mul(fld, t); }
inline void negate(Field &fld) { Field z = Field::zero(); subtract(z, fld); fld = z; }
// Exponentiates this element (square and multiply; result stays in Montgomery form, pow 0 = one())
inline void pow(Field &fld1, const size_t e) {
    uint64_t a[12];
    detail::to64(a, fld1);
    modulus().f.pow(a, a, (uint64_t)e);
    detail::from64(fld1, a);
}
// Multiplicative inverse (x^(p-2)); zero stays zero
inline void mul_inv(Field &fld1) {
    const ModulusState &M = modulus();
    uint32_t e[24];
    memcpy(e, M.p, 96);
    e[0] -= 2;  // p is odd and p[0] >= 3 for both moduli
    Field acc = Field::one();
    for (int bit = 32 * 24 - 1; bit >= 0; --bit) {
        square(acc);
        if ((e[bit / 32] >> (bit % 32)) & 1) mul(acc, fld1);
    }
    fld1 = acc;
}

inline Field operator*(const Field &a, const Field &b) { Field r = a; mul(r, b); return r; }
inline Field operator+(const Field &a, const Field &b) { Field r = a; add(r, b); return r; }
inline Field operator-(const Field &a, const Field &b) { Field r = a; subtract(r, b); return r; }
inline Field operator-(const Field &a) { Field r = a; negate(r); return r; }
inline Field operator^(const Field &a, const size_t e) { Field r = a; pow(r, e); return r; }

inline Field Field::from_uint(uint32_t v) { Field r(v); mul(r, Field(modulus().r2)); return r; }
inline Field Field::root_of_unity(size_t n) {
    Field w(modulus().root);
    size_t order = (size_t)1 << modulus().two_adicity;
    while (order > n) { square(w); order >>= 1; }
    return w;
}

}  // namespace cpu_fields
#endif
