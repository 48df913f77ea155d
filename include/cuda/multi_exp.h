// cuda/multi_exp.h -- the reference's multi-exponentiation entry point, same header path, same template
//     template <typename FieldT, typename FieldMul> FieldT multiexp(std::vector<FieldT>& a, std::vector<FieldMul>& mul);
// (reference cuda/multi_exp.h:24-25; defined and explicitly instantiated in cuda/multi_exp.cu:104-142 for
// <Scalar, Scalar> and <mnt4753_G1, Scalar>; called from test/main.cpp:117,168).  Header-only here, forwarding to the C
// ABI of libgpusnarks_b200.so:
//     multiexp<fields::Scalar, fields::Scalar>      -> gsn_fp768_inner_product_host   (sum_i a[i] * mul[i] in the field)
//     multiexp<fields::mnt4753_G1, fields::Scalar>  -> gsn_g1_multiexp_host           (sum_i mul[i] * a[i] on the curve, bucket method)
//                                                      gsn_g1_multiexp_multi_host     (2^18 points and more, several GPUs visible)
// Failures are reported by a std::runtime_error carrying gsn_last_error() (the reference prints and continues).
#pragma once
#include <algorithm>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "../gpusnarks_b200.h"
#include "device_field.h"
#include "fft_kernel.h"

#ifdef __cplusplus

namespace gsn {
template <typename FieldT, typename FieldMul> struct multiexp_dispatch;  // unknown type pairs do not link silently

template <> struct multiexp_dispatch<fields::Scalar, fields::Scalar> {
    static fields::Scalar run(std::vector<fields::Scalar> &a, std::vector<fields::Scalar> &mul) {
        if (a.size() != mul.size()) throw std::runtime_error("multiexp: vectors must have equal length");
        fields::Scalar out;
        check(gsn_fp768_inner_product_host(default_ctx(), out.im_rep, reinterpret_cast<const uint32_t *>(a.data()),
                                           reinterpret_cast<const uint32_t *>(mul.data()), a.size()), "gsn_fp768_inner_product_host");
        return out;
    }
};
template <> struct multiexp_dispatch<fields::mnt4753_G1, fields::Scalar> {
    static fields::mnt4753_G1 run(std::vector<fields::mnt4753_G1> &a, std::vector<fields::Scalar> &mul) {
        if (a.size() != mul.size()) throw std::runtime_error("multiexp: vectors must have equal length");
        fields::mnt4753_G1 out;
        // 2^18 points and more are cut into one slice per visible GPU (independent work; GSN_MULTI_DEVICES=1 switches it off)
        int devices = 1;
        if (a.size() >= ((size_t)1 << 18)) {
            gsn_device_count(&devices);
            if (const char *e = std::getenv("GSN_MULTI_DEVICES")) devices = std::min(devices, std::atoi(e));
        }
        if (devices > 1) {
            std::vector<int> ids(devices);
            for (int d = 0; d < devices; ++d) ids[d] = d;
            check(gsn_g1_multiexp_multi_host(ids.data(), (unsigned)devices, reinterpret_cast<uint32_t *>(&out), reinterpret_cast<const uint32_t *>(a.data()),
                                             reinterpret_cast<const uint32_t *>(mul.data()), a.size()), "gsn_g1_multiexp_multi_host");
            return out;
        }
        check(gsn_g1_multiexp_host(default_ctx(), reinterpret_cast<uint32_t *>(&out), reinterpret_cast<const uint32_t *>(a.data()),
                                   reinterpret_cast<const uint32_t *>(mul.data()), a.size()), "gsn_g1_multiexp_host");
        return out;
    }
};
}  // namespace gsn

template <typename FieldT, typename FieldMul>
FieldT multiexp(std::vector<FieldT> &a, std::vector<FieldMul> &mul) {
    return gsn::multiexp_dispatch<FieldT, FieldMul>::run(a, mul);
}

#endif
