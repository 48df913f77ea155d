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
#include <stdexcept>
#include <string>
#include <vector>

#include "../gpusnarks_b200.h"
#include "../fields/dummy_field.h"
#include "../fields/field.h"
#include "device_field.h"

#ifdef __cplusplus

namespace gsn {

// one process-wide context per device for the template entry points
inline gsn_ctx *default_ctx(int device = 0) {
    static gsn_ctx *ctx[16] = {nullptr};
    if (device < 0 || device >= 16) throw std::runtime_error("gsn::default_ctx: device out of range");
    if (!ctx[device] && gsn_ctx_create(&ctx[device], device) != GSN_OK)
        throw std::runtime_error(std::string("gsn_ctx_create: ") + gsn_last_error());
    return ctx[device];
}
inline void check(int rc, const char *what) {
    if (rc != GSN_OK) throw std::runtime_error(std::string(what) + ": " + gsn_last_error());
}

template <typename FieldT> struct fft_dispatch;  // no generic definition: unknown field types do not link silently

template <> struct fft_dispatch<fields::Scalar> {
    static void run(std::vector<fields::Scalar> &a, const fields::Scalar &omg, int inverse) {
        check(gsn_ntt768_host(default_ctx(), reinterpret_cast<uint32_t *>(a.data()), a.size(), omg.im_rep, inverse), "gsn_ntt768_host");
    }
};
template <> struct fft_dispatch<cpu_fields::Field> {
    static void run(std::vector<cpu_fields::Field> &a, const cpu_fields::Field &omg, int inverse) {
        static_assert(sizeof(cpu_fields::Field) == 96, "raw limbs");
        check(gsn_ntt768_host(default_ctx(), reinterpret_cast<uint32_t *>(a.data()), a.size(), omg.im_rep, inverse), "gsn_ntt768_host");
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
