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
#include "../fields/g1_host.h"

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

// bit `index` of the raw limbs (reference hasBitAt, cuda/device_field_operators.h; scalars of operator* are raw integers)
inline bool hasBitAt(const Scalar &fld, long index) { return index >= 0 && index < 32 * SIZE && ((fld.im_rep[index / 32] >> (index % 32)) & 1); }

// `fp2` = Fq[u]/(u^2 - 13) over MNT4-753 Fq, the reference's type (cuda/device_field.h:220-294): same members and
// operators.  Coordinates are Fq elements in Montgomery form carried in `Scalar` limbs; the arithmetic below is Fq
// arithmetic (fields/g1_host.h) whatever modulus fields::Scalar's own operators use.  non_residue is the field element 13.
struct fp2 {
    Scalar x, y;
    fp2() = default;
    fp2(Scalar _x, Scalar _y) : x(_x), y(_y) {}
    static fp2 zero() { return fp2(); }
    static Scalar fq(const uint64_t *v) { Scalar s; memcpy(s.im_rep, v, 96); return s; }
    static void lim(uint64_t *d, const Scalar &s) { memcpy(d, s.im_rep, 96); }
    fp2 operator+(const fp2 &r) const { const auto &f = gsn::host::fq_field(); uint64_t a[12], b[12], c[12], d[12]; lim(a, x); lim(b, r.x); lim(c, y); lim(d, r.y); f.add(a, a, b); f.add(c, c, d); return fp2(fq(a), fq(c)); }
    fp2 operator-(const fp2 &r) const { const auto &f = gsn::host::fq_field(); uint64_t a[12], b[12], c[12], d[12]; lim(a, x); lim(b, r.x); lim(c, y); lim(d, r.y); f.sub(a, a, b); f.sub(c, c, d); return fp2(fq(a), fq(c)); }
    fp2 operator-() const { return zero() - *this; }
    fp2 operator*(const fp2 &r) const {   // Karatsuba, as reference device_field.h:253-262
        const auto &f = gsn::host::fq_field();
        uint64_t a[12], b[12], A[12], B[12], aA[12], bB[12], s1[12], s2[12], t[12], t4[12], ox[12], oy[12];
        lim(a, x); lim(b, y); lim(A, r.x); lim(B, r.y);
        f.mul(aA, a, A); f.mul(bB, b, B);
        f.add(s1, a, b); f.add(s2, A, B); f.mul(oy, s1, s2); f.sub(oy, oy, aA); f.sub(oy, oy, bB);
        f.add(t, bB, bB); f.add(t4, t, t); f.add(t, t4, t4); f.add(t, t, t4); f.add(t, t, bB);   // 13 bB
        f.add(ox, aA, t);
        return fp2(fq(ox), fq(oy));
    }
    bool operator==(const fp2 &r) const { return x == r.x && y == r.y; }
    static void print(fp2 f) { printf("FP2: "); Scalar::print(f.x); Scalar::print(f.y); printf("\n"); }
};

// `mnt4753_G1`, the reference's host/device point type (cuda/device_field.h:296-437): homogeneous projective (x : y : z)
// over Fq, a = 2, the same formulas (add-1998-cmo-2, dbl-2007-bl, MSB-first double-and-add for operator*).  Corrections:
// the identity is z == 0 and is handled (the reference's operator+ has no identity / doubling / inverse cases, so
// its zero() + P is (0,0,0)); zero() is (0 : 1 : 0); testEquality compares the POINTS (cross-multiplied), since two
// computations of one point may end in different projective representatives.
struct mnt4753_G1 {
    Scalar x, y, z;
    mnt4753_G1() = default;
    mnt4753_G1(Scalar _x, Scalar _y, Scalar _z) : x(_x), y(_y), z(_z) {}
    gsn::host::G1Host raw() const { gsn::host::G1Host p; memcpy(p.x, x.im_rep, 96); memcpy(p.y, y.im_rep, 96); memcpy(p.z, z.im_rep, 96); return p; }
    static mnt4753_G1 wrap(const gsn::host::G1Host &p) { mnt4753_G1 r; memcpy(r.x.im_rep, p.x, 96); memcpy(r.y.im_rep, p.y, 96); memcpy(r.z.im_rep, p.z, 96); return r; }
    static mnt4753_G1 zero() { gsn::host::G1Host p; gsn::host::G1Ops(gsn::host::fq_field()).identity(p); return wrap(p); }
    static bool is_zero(const mnt4753_G1 &g) { return g.z.is_zero(); }
    mnt4753_G1 operator+(const mnt4753_G1 &o) const { gsn::host::G1Host r; gsn::host::G1Ops(gsn::host::fq_field()).add(r, raw(), o.raw()); return wrap(r); }
    mnt4753_G1 dbl() const { gsn::host::G1Host r; gsn::host::G1Ops(gsn::host::fq_field()).dbl(r, raw()); return wrap(r); }
    mnt4753_G1 operator-() const { const auto &f = gsn::host::fq_field(); uint64_t zero12[12] = {0}, t[12]; memcpy(t, y.im_rep, 96); f.sub(t, zero12, t); mnt4753_G1 r = *this; memcpy(r.y.im_rep, t, 96); return r; }
    mnt4753_G1 operator-(const mnt4753_G1 &o) const { return *this + (-o); }
    void operator+=(const mnt4753_G1 &o) { *this = *this + o; }
    mnt4753_G1 operator*(const Scalar &k) const {   // scalar = the raw 768-bit integer in k.im_rep (reference :394-411)
        mnt4753_G1 result = zero();
        bool one = false;
        for (long i = SIZE * 32 - 1; i >= 0; --i) {
            if (one) result = result.dbl();
            if (hasBitAt(k, i)) { one = true; result = result + *this; }
        }
        return result;
    }
    bool same_point(const mnt4753_G1 &o) const {
        if (is_zero(*this) || is_zero(o)) return is_zero(*this) && is_zero(o);
        const auto &f = gsn::host::fq_field();
        gsn::host::G1Host a = raw(), b = o.raw();
        uint64_t l[12], r[12];
        f.mul(l, a.x, b.z); f.mul(r, b.x, a.z);
        if (memcmp(l, r, 96)) return false;
        f.mul(l, a.y, b.z); f.mul(r, b.y, a.z);
        return memcmp(l, r, 96) == 0;
    }
    static void print(mnt4753_G1 f) { printf("\nmnt4753_G1: \n"); Scalar::print(f.x); Scalar::print(f.y); Scalar::print(f.z); printf("----\n"); }
    static void testEquality(mnt4753_G1 f1, mnt4753_G1 f2) {
        if (!f1.same_point(f2)) { printf("Missmatch: \n"); print(f1); print(f2); assert(!"missmatch"); }
    }
};
static_assert(sizeof(mnt4753_G1) == 288 && sizeof(fp2) == 192, "raw limb layouts (the C ABI memcpy's them)");

}  // namespace fields
#endif
