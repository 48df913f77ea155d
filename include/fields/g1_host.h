// fields/g1_host.h -- MNT4-753 G1 arithmetic over Fq on the CPU.  Used (i) by the library for the one strictly
// sequential step of the bucket-method multi-exponentiation (Horner over the per-window sums: ~750 doublings that a
// single GPU thread would take ~25 ms for, a CPU core ~2 ms) and (ii) by the host type fields::mnt4753_G1 of
// cuda/device_field.h.  Same coordinates and formulas as the device code (gpusnarks_b200/csrc/g1.cuh): homogeneous
// projective, a = 2, add-1998-cmo-2 / dbl-2007-bl, identity = Z == 0 (reference cuda/device_field.h:327-372).
// Product code: independent of oracle/.
#pragma once
#include <cstdint>
#include <cstring>

#include "../gsn_constants.h"
#include "fp768_host.h"

namespace gsn {
namespace host {

struct Fq768 : Field768 {
    void add(uint64_t *r, const uint64_t *a, const uint64_t *b) const {
        uint64_t t[12];
        unsigned __int128 c = 0;
        for (int i = 0; i < 12; ++i) { c += (unsigned __int128)a[i] + b[i]; t[i] = (uint64_t)c; c >>= 64; }
        if (c || geq(t, p)) sub_n(r, t, p);
        else memcpy(r, t, 96);
    }
    void sub(uint64_t *r, const uint64_t *a, const uint64_t *b) const {
        uint64_t t[12];
        if (sub_n(t, a, b)) {
            unsigned __int128 c = 0;
            for (int i = 0; i < 12; ++i) { c += (unsigned __int128)t[i] + p[i]; t[i] = (uint64_t)c; c >>= 64; }
        }
        memcpy(r, t, 96);
    }
    static bool is_zero(const uint64_t *a) { for (int i = 0; i < 12; ++i) if (a[i]) return false; return true; }
};

struct G1Host {
    uint64_t x[12], y[12], z[12];
};

// the curve's base field, MNT4-753 Fq (the reference's literal `_mod`), whatever modulus the caller's FFT field uses
inline const Fq768 &fq_field() {
    static const Fq768 f = [] {
        static const uint32_t p[24] = GSN_FQ_MOD, r1[24] = GSN_FQ_R1;
        Fq768 x;
        x.init(p, r1);
        return x;
    }();
    return f;
}

struct G1Ops {
    const Fq768 &f;
    explicit G1Ops(const Fq768 &field) : f(field) {}
    void identity(G1Host &r) const { memset(&r, 0, sizeof(r)); memcpy(r.y, f.r1, 96); }
    bool is_identity(const G1Host &p) const { return Fq768::is_zero(p.z); }

    void dbl(G1Host &r, const G1Host &p) const {
        uint64_t XX[12], ZZ[12], w[12], s[12], ss[12], sss[12], R[12], RR[12], T[12], B[12], h[12], t[12];
        f.mul(XX, p.x, p.x);
        f.mul(ZZ, p.z, p.z);
        f.add(w, ZZ, ZZ);
        f.add(t, XX, XX);
        f.add(t, t, XX);
        f.add(w, w, t);              // w = a*ZZ + 3*XX, a = 2
        f.mul(s, p.y, p.z);
        f.add(s, s, s);
        f.mul(ss, s, s);
        f.mul(sss, s, ss);
        f.mul(R, p.y, s);
        f.mul(RR, R, R);
        f.add(T, p.x, R);
        f.mul(T, T, T);
        f.sub(B, T, XX);
        f.sub(B, B, RR);
        f.mul(h, w, w);
        f.sub(h, h, B);
        f.sub(h, h, B);
        G1Host o;
        f.mul(o.x, h, s);
        f.sub(t, B, h);
        f.mul(t, w, t);
        f.sub(t, t, RR);
        f.sub(o.y, t, RR);
        memcpy(o.z, sss, 96);
        r = o;
    }

    void add(G1Host &r, const G1Host &p, const G1Host &q) const {
        if (is_identity(p)) { r = q; return; }
        if (is_identity(q)) { r = p; return; }
        uint64_t X1Z2[12], Y1Z2[12], Z1Z2[12], u[12], v[12], uu[12], vv[12], vvv[12], R[12], A[12], t[12];
        f.mul(X1Z2, p.x, q.z);
        f.mul(Y1Z2, p.y, q.z);
        f.mul(Z1Z2, p.z, q.z);
        f.mul(u, q.y, p.z);
        f.sub(u, u, Y1Z2);
        f.mul(v, q.x, p.z);
        f.sub(v, v, X1Z2);
        if (Fq768::is_zero(v)) {
            if (Fq768::is_zero(u)) { G1Host c = p; dbl(r, c); }
            else identity(r);
            return;
        }
        f.mul(uu, u, u);
        f.mul(vv, v, v);
        f.mul(vvv, vv, v);
        f.mul(R, vv, X1Z2);
        f.mul(A, uu, Z1Z2);
        f.add(t, R, R);
        f.add(t, t, vvv);
        f.sub(A, A, t);
        G1Host o;
        f.mul(o.x, v, A);
        f.sub(t, R, A);
        f.mul(t, u, t);
        f.mul(Y1Z2, vvv, Y1Z2);
        f.sub(o.y, t, Y1Z2);
        f.mul(o.z, vvv, Z1Z2);
        r = o;
    }
};

}  // namespace host
}  // namespace gsn
