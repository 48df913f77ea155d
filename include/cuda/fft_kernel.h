// cuda/fft_kernel.h -- the reference's host FFT entry point, same header path, same template
//     template <typename FieldT> void best_fft(std::vector<FieldT>& a, const FieldT& omg);
// (reference cuda/fft_kernel.h:24-25).  The reference explicitly instantiates it inside its
// CUDA static library for fields::Scalar only (cuda/fft_kernel.cu:150); here it is header-only
// and forwards, per field type, to the C ABI of libgpusnarks_b200.so (gpusnarks_b200.h):
//     fields::Scalar / cpu_fields::Field  -> gsn_ntt768_host     (768-bit, MNT4-753 Fr)
//     dummy_fields::Field                 -> gsn_ntt32_host      (32-bit prime field)
// Unlike the reference (which ignores `omg`, asserts n == 2^16 and exit(-1)s on a launch
// failure, cuda/fft_kernel.cu:118-142) any power-of-two length is accepted, omega is used
// and validated, and failures are reported by a std::runtime_error carrying gsn_last_error().
// best_ifft is the inverse (omega^-1, n^-1), which the reference does not have.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../gpusnarks_b200.h"
#include "../fields/dummy_field.h"
#include "../fields/field.h"
#include "device_field.h"

#ifdef __cplusplus

namespace gsn {

inline void check(int rc, const char *what) {
    if (rc != GSN_OK) throw std::runtime_error(std::string(what) + ": " + gsn_last_error());
}
// one process-wide context per (device, 768-bit field) for the template entry points (created once, thread safe;
// the field is a property of the context, so the Fr and the Fq context of a device never disturb each other)
inline gsn_ctx *default_ctx(int device = 0, int field = GSN_FIELD_MNT4753_FR) {
    static gsn_ctx *ctx[16][2] = {{nullptr}};
    static std::mutex mu;
    if (device < 0 || device >= 16 || field < 0 || field > 1) throw std::runtime_error("gsn::default_ctx: device / field out of range");
    std::lock_guard<std::mutex> lk(mu);
    if (!ctx[device][field]) {
        check(gsn_ctx_create(&ctx[device][field], device), "gsn_ctx_create");
        if (field != GSN_FIELD_MNT4753_FR) check(gsn_set_field768(ctx[device][field], field), "gsn_set_field768");
    }
    return ctx[device][field];
}
// Transforms of 2^24 elements and more are sharded over every visible GPU (1, 2, 4 or 8 of them; the four-step plan of
// gsn_multi_*, peer access between the devices of this one process).  GSN_MULTI_MIN_LOG_N overrides the threshold,
// GSN_MULTI_DEVICES caps the device count (1 disables sharding).  Plans are cached per (n, omega).
inline gsn_multi *default_multi(size_t n, const uint32_t *omega) {
    struct Entry { size_t n; uint32_t omega[GSN_FP768_LIMBS]; gsn_multi *m; };
    static std::vector<Entry> cache;
    static std::mutex mu;
    static int devices = -1;
    std::lock_guard<std::mutex> lk(mu);
    if (devices < 0) {
        int cnt = 0;
        gsn_device_count(&cnt);
        if (const char *e = std::getenv("GSN_MULTI_DEVICES")) cnt = std::min(cnt, std::atoi(e));
        devices = 1;
        while (devices * 2 <= cnt && devices < 8) devices *= 2;
    }
    const char *e = std::getenv("GSN_MULTI_MIN_LOG_N");
    const unsigned min_log = e ? (unsigned)std::atoi(e) : 24u;
    if (devices < 2 || n < ((size_t)1 << min_log) || (n & (n - 1))) return nullptr;
    for (auto &c : cache)
        if (c.n == n && std::memcmp(c.omega, omega, sizeof(c.omega)) == 0) return c.m;
    int ids[8] = {0, 1, 2, 3, 4, 5, 6, 7};
    Entry ent;
    ent.n = n;
    std::memcpy(ent.omega, omega, sizeof(ent.omega));
    if (gsn_multi_create(&ent.m, ids, (unsigned)devices, n, omega, 3) != GSN_OK) {
        devices = 1;   // e.g. no peer access between the devices: keep to the single-GPU path from now on
        return nullptr;
    }
    if (cache.size() >= 2) { gsn_multi_destroy(cache.front().m); cache.erase(cache.begin()); }
    cache.push_back(ent);
    return ent.m;
}

template <typename FieldT> struct fft_dispatch;  // no generic definition: unknown field types do not link silently

template <> struct fft_dispatch<fields::Scalar> {
    static void run(std::vector<fields::Scalar> &a, const fields::Scalar &omg, int inverse) {
        if (gsn_multi *m = default_multi(a.size(), omg.im_rep)) {
            static std::mutex call_mu;   // a gsn_multi owns its devices for the duration of a call: one sharded transform at a time
            std::lock_guard<std::mutex> lk(call_mu);
            check(gsn_multi_ntt768_host(m, reinterpret_cast<uint32_t *>(a.data()), inverse), "gsn_multi_ntt768_host");
            return;
        }
        check(gsn_ntt768_host(default_ctx(), reinterpret_cast<uint32_t *>(a.data()), a.size(), omg.im_rep, inverse), "gsn_ntt768_host");
    }
};
template <> struct fft_dispatch<cpu_fields::Field> {
    static void run(std::vector<cpu_fields::Field> &a, const cpu_fields::Field &omg, int inverse) {
        static_assert(sizeof(cpu_fields::Field) == 96, "raw limbs");
        // cpu_fields::Field follows the host-side modulus selection (cpu_fields::modulus().select(0 = Fr | 1 = Fq))
        const int field = cpu_fields::modulus().which ? GSN_FIELD_MNT4753_FQ : GSN_FIELD_MNT4753_FR;
        check(gsn_ntt768_host(default_ctx(0, field), reinterpret_cast<uint32_t *>(a.data()), a.size(), omg.im_rep, inverse), "gsn_ntt768_host");
    }
};
template <> struct fft_dispatch<dummy_fields::Field> {
    static void run(std::vector<dummy_fields::Field> &a, const dummy_fields::Field &omg, int inverse) {
        static_assert(sizeof(dummy_fields::Field) == 4, "raw residues");
        check(gsn_ntt32_host(default_ctx(), reinterpret_cast<uint32_t *>(a.data()), a.size(), omg.im_rep, dummy_fields::Field::mod, inverse),
              "gsn_ntt32_host");
    }
};

}  // namespace gsn

template <typename FieldT>
void best_fft(std::vector<FieldT> &a, const FieldT &omg) {
    gsn::fft_dispatch<FieldT>::run(a, omg, 0);
}

template <typename FieldT>
void best_ifft(std::vector<FieldT> &a, const FieldT &omg) {
    gsn::fft_dispatch<FieldT>::run(a, omg, 1);
}

// Several equally long vectors over the 768-bit field in one call: same results as calling best_fft on each, but the
// host<->device copies of consecutive vectors overlap with each other and with the transforms (gsn_ntt768_host_batch).
inline void best_fft_batch(std::vector<std::vector<fields::Scalar>> &vs, const fields::Scalar &omg, bool inverse = false) {
    if (vs.empty()) return;
    std::vector<uint32_t *> ptrs;
    for (auto &v : vs) {
        if (v.size() != vs[0].size()) throw std::runtime_error("best_fft_batch: vectors must have equal length");
        ptrs.push_back(reinterpret_cast<uint32_t *>(v.data()));
    }
    gsn::check(gsn_ntt768_host_batch(gsn::default_ctx(), ptrs.data(), ptrs.size(), vs[0].size(), omg.im_rep, inverse ? 1 : 0),
               "gsn_ntt768_host_batch");
}

#endif
