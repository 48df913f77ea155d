// oracle/field32.h -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the reference's 32-bit field sketch `dummy_fields::Field`
// (reference fields/dummy_field.h:24-62, fields/dummy_field.cpp:25-107): one uint32_t
// `im_rep`, a static `mod`.  The reference leaves `mod = 0` and does C wrap-around
// arithmetic (dummy_field.cpp:25,75-90), so it has no prime and cannot host an NTT
// (SURVEY.md F6); here `mod` is a real NTT-friendly prime (default 2013265921 =
// 15*2^27+1) and every operation reduces into [0, mod).  The operator surface
// (* + - ^ ==, zero(), one()) is what reference test/fft_host.h needs.
#pragma once
#include <cstdint>
#include "../include/gsn_constants.h"

namespace oracle {

struct Fp32 {
    uint32_t im_rep;
    static uint32_t &mod() { static uint32_t m = GSN_P32_DEFAULT; return m; }

    Fp32() : im_rep(0) {}
    Fp32(uint32_t v) : im_rep(v) {}
    static Fp32 zero() { return Fp32(0); }   // dummy_field.cpp:28-33
    static Fp32 one() { return Fp32(1); }    // dummy_field.cpp:36-41
    bool is_zero() const { return im_rep == 0; }
    bool operator==(const Fp32 &o) const { return im_rep == o.im_rep; }
    bool operator!=(const Fp32 &o) const { return im_rep != o.im_rep; }
    Fp32 operator+(const Fp32 &o) const { uint64_t s = (uint64_t)im_rep + o.im_rep; return Fp32((uint32_t)(s % mod())); }
    Fp32 operator-(const Fp32 &o) const { uint64_t s = (uint64_t)im_rep + mod() - o.im_rep; return Fp32((uint32_t)(s % mod())); }
    Fp32 operator*(const Fp32 &o) const { return Fp32((uint32_t)((uint64_t)im_rep * o.im_rep % mod())); }
    Fp32 operator^(uint64_t e) const {
        Fp32 acc = one(), base = *this;
        while (e) { if (e & 1) acc = acc * base; base = base * base; e >>= 1; }
        return acc;
    }
    Fp32 inverse() const { return *this ^ (uint64_t)(mod() - 2); }
};

}  // namespace oracle
