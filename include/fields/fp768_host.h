// fields/fp768_host.h -- host-side 768-bit Montgomery arithmetic: used by the library planner
// (omega validation, omega^-1, n^-1).  Product code: independent of oracle/ (which is test
// infrastructure) -- 12 x 64-bit limbs with unsigned __int128, where the oracle uses the
// reference's 24 x 32-bit CIOS loop nest.
#pragma once
#include <cstdint>
#include <cstring>

namespace gsn {
namespace host {

struct Field768 {
    uint64_t p[12];
    uint64_t r1[12];
    uint64_t np0;  // -p^-1 mod 2^64

    void init(const uint32_t *p32, const uint32_t *r1_32) {
        memcpy(p, p32, 96);
        memcpy(r1, r1_32, 96);
        uint64_t inv = 1;  // Newton: inv = p^-1 mod 2^64
        for (int i = 0; i < 6; ++i) inv *= 2 - p[0] * inv;
        np0 = (uint64_t)0 - inv;
    }
    static bool geq(const uint64_t *a, const uint64_t *b) {
        for (int i = 11; i >= 0; --i)
            if (a[i] != b[i]) return a[i] > b[i];
        return true;
    }
    static uint64_t sub_n(uint64_t *r, const uint64_t *a, const uint64_t *b) {
        unsigned __int128 borrow = 0;
        for (int i = 0; i < 12; ++i) {
            unsigned __int128 t = (unsigned __int128)a[i] - b[i] - borrow;
            r[i] = (uint64_t)t;
            borrow = (t >> 64) & 1;
        }
        return (uint64_t)borrow;
    }
    // r = a * b * 2^-768 mod p (canonical)
    void mul(uint64_t *r, const uint64_t *a, const uint64_t *b) const {
        uint64_t t[14] = {0};
        for (int i = 0; i < 12; ++i) {
            unsigned __int128 c = 0;
            for (int j = 0; j < 12; ++j) {
                c += (unsigned __int128)a[j] * b[i] + t[j];
                t[j] = (uint64_t)c;
                c >>= 64;
            }
            c += t[12];
            t[12] = (uint64_t)c;
            t[13] = (uint64_t)(c >> 64);
            const uint64_t m = t[0] * np0;
            c = ((unsigned __int128)m * p[0] + t[0]) >> 64;
            for (int j = 1; j < 12; ++j) {
                c += (unsigned __int128)m * p[j] + t[j];
                t[j - 1] = (uint64_t)c;
                c >>= 64;
            }
            c += t[12];
            t[11] = (uint64_t)c;
            t[12] = t[13] + (uint64_t)(c >> 64);
        }
        if (t[12] || geq(t, p)) sub_n(r, t, p);
        else memcpy(r, t, 96);
    }
    void pow(uint64_t *r, const uint64_t *a, uint64_t e) const {
        uint64_t acc[12], base[12];
        memcpy(acc, r1, 96);
        memcpy(base, a, 96);
        while (e) {
            if (e & 1) mul(acc, acc, base);
            mul(base, base, base);
            e >>= 1;
        }
        memcpy(r, acc, 96);
    }
    // r = a^-1 (Fermat: a^(p-2)), Montgomery form in and out; a != 0
    void inv(uint64_t *r, const uint64_t *a) const {
        uint64_t e[12], two[12] = {2}, acc[12], base[12];
        sub_n(e, p, two);
        memcpy(acc, r1, 96);
        memcpy(base, a, 96);
        for (int i = 0; i < 768; ++i) {
            if ((e[i >> 6] >> (i & 63)) & 1) mul(acc, acc, base);
            mul(base, base, base);
        }
        memcpy(r, acc, 96);
    }
    // r = a / 2 mod p
    void halve(uint64_t *r, const uint64_t *a) const {
        uint64_t t[13];
        memcpy(t, a, 96);
        t[12] = 0;
        if (t[0] & 1) {
            unsigned __int128 c = 0;
            for (int i = 0; i < 12; ++i) {
                c += (unsigned __int128)t[i] + p[i];
                t[i] = (uint64_t)c;
                c >>= 64;
            }
            t[12] = (uint64_t)c;
        }
        for (int i = 0; i < 12; ++i) r[i] = (t[i] >> 1) | (t[i + 1] << 63);
    }
    bool is_one(const uint64_t *a) const { return memcmp(a, r1, 96) == 0; }
    bool is_minus_one(const uint64_t *a) const {
        uint64_t t[12];
        sub_n(t, p, r1);
        return memcmp(a, t, 96) == 0;
    }
};

}  // namespace host
}  // namespace gsn
