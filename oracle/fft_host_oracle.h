// oracle/fft_host_oracle.h -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the reference's host FFT templates (reference test/fft_host.h):
//   bitreverse_host                   <- test/fft_host.h:7-16
//   _basic_serial_radix2_FFT          <- test/fft_host.h:18-54
//   _basic_parallel_radix2_FFT_inner  <- test/fft_host.h:56-117
// Control flow, loop order, twiddle recurrences and the output layout are the
// reference's.  ONE statement differs, on purpose: the reference butterfly
//     t = w;  w = w * a[k+j+m];                     (test/fft_host.h:44-45)
// adds/subtracts the twiddle instead of twiddle*a[k+j+m] and folds data into the
// twiddle chain, so the reference routine is not a DFT (it disagrees with the naive
// O(n^2) DFT at every index).  The statement libff has there, and the one used below, is
//     t = w * a[k+j+m];
// With it serial == parallel == naive DFT  A[i] = sum_j a[j] * omega^(i*j), natural
// order in, natural order out.  The inverse transform (absent from the reference, F8)
// follows the libff convention: forward transform with omega^-1, then scale by n^-1.
//
// Pinning (tests/test_oracle_pins.py).  The reference holds no golden vectors for the FFT (its only check is
// GPU == host on a constant input, test/main.cpp:80-84), so this restatement is pinned against the reference itself,
// compiled here from /root/reference into oracle/_ref: the reference's own fft_host.h templates with the one butterfly
// statement corrected by sed, instantiated over the oracle's field types, agree with these functions bit for bit
// (serial and parallel entry points, both fields); the UNPATCHED templates are asserted NOT to be a DFT.  Plus the
// DFT definition in Python big-ints and the naive O(n^2) DFT below.
#pragma once
#include <cstddef>
#include <cstdint>
#include <utility>
#include <vector>

namespace oracle {

inline size_t log2_exact(size_t n) { size_t l = 0; while (((size_t)1 << l) < n) ++l; return l; }

// test/fft_host.h:7-16
inline size_t bitreverse_host(size_t n, const size_t l) {
    size_t r = 0;
    for (size_t k = 0; k < l; ++k) {
        r = (r << 1) | (n & 1);
        n >>= 1;
    }
    return r;
}

// test/fft_host.h:18-54
template <typename FieldT>
void _basic_serial_radix2_FFT(std::vector<FieldT> &a, const FieldT omega, const FieldT one) {
    const size_t n = a.size(), logn = log2_exact(n);

    for (size_t k = 0; k < n; ++k) {  // :24-29 swapping in place
        const size_t rk = bitreverse_host(k, logn);
        if (k < rk) std::swap(a[k], a[rk]);
    }

    size_t m = 1;  // invariant: m = 2^{s-1}
    for (size_t s = 1; s <= logn; ++s) {
        FieldT w_m = omega ^ (uint64_t)(n / (2 * m));  // :35-36, 2^s-th root of unity
        for (size_t k = 0; k < n; k += 2 * m) {
            FieldT w = one;
            for (size_t j = 0; j < m; ++j) {
                const FieldT t = w * a[k + j + m];  // corrected statement (reference :44-45)
                a[k + j + m] = a[k + j] - t;        // :46
                a[k + j] = a[k + j] + t;            // :47
                w = w * w_m;                        // :48
            }
        }
        m *= 2;
    }
}

// test/fft_host.h:56-117
template <typename FieldT>
void _basic_parallel_radix2_FFT_inner(std::vector<FieldT> &a, const FieldT omega, const size_t log_cpus, const FieldT one) {
    const size_t num_cpus = (size_t)1 << log_cpus;
    const size_t m = a.size();
    const size_t log_m = log2_exact(m);

    if (log_m < log_cpus) {  // :64-68
        _basic_serial_radix2_FFT(a, omega, one);
        return;
    }

    std::vector<std::vector<FieldT>> tmp(num_cpus);  // :70-74
    for (size_t j = 0; j < num_cpus; ++j) tmp[j].resize((size_t)1 << (log_m - log_cpus), FieldT::zero());

#pragma omp parallel for
    for (size_t j = 0; j < num_cpus; ++j) {  // :76-102
        const FieldT omega_j = omega ^ (uint64_t)j;
        const FieldT omega_step = omega ^ (uint64_t)(j << (log_m - log_cpus));

        FieldT elt = FieldT::one();
        for (size_t i = 0; i < (size_t)1 << (log_m - log_cpus); ++i) {
            for (size_t s = 0; s < num_cpus; ++s) {
                // invariant: elt is omega^(j*idx)
                const size_t idx = (i + (s << (log_m - log_cpus))) % ((size_t)1 << log_m);
                tmp[j][i] = tmp[j][i] + a[idx] * elt;
                elt = elt * omega_step;
            }
            elt = elt * omega_j;
        }
    }

    const FieldT omega_num_cpus = omega ^ (uint64_t)num_cpus;  // :103
#pragma omp parallel for
    for (size_t j = 0; j < num_cpus; ++j) _basic_serial_radix2_FFT(tmp[j], omega_num_cpus, one);  // :104-108

#pragma omp parallel for
    for (size_t i = 0; i < num_cpus; ++i)  // :110-116
        for (size_t j = 0; j < (size_t)1 << (log_m - log_cpus); ++j) a[(j << log_cpus) + i] = tmp[i][j];
}

// Definition check: A[i] = sum_j a[j] * omega^(i*j).  O(n^2); small n only.
template <typename FieldT>
std::vector<FieldT> naive_dft(const std::vector<FieldT> &a, const FieldT omega) {
    const size_t n = a.size();
    std::vector<FieldT> out(n, FieldT::zero());
    FieldT wi = FieldT::one();  // omega^i
    for (size_t i = 0; i < n; ++i) {
        FieldT w = FieldT::one();  // omega^(i*j)
        FieldT acc = FieldT::zero();
        for (size_t j = 0; j < n; ++j) {
            acc = acc + a[j] * w;
            w = w * wi;
        }
        out[i] = acc;
        wi = wi * omega;
    }
    return out;
}

// One output coefficient by Horner: A[k] = sum_j a[j] * (omega^k)^j.  n multiplies.
template <typename FieldT>
FieldT dft_point(const FieldT *a, size_t n, const FieldT omega, uint64_t k) {
    const FieldT x = omega ^ k;
    FieldT acc = FieldT::zero();
    for (size_t j = n; j-- > 0;) acc = acc * x + a[j];
    return acc;
}

// Inverse by the libff convention: forward with omega^-1, then multiply by n^-1.
template <typename FieldT>
void inverse_FFT(std::vector<FieldT> &a, const FieldT omega, const size_t log_cpus) {
    const size_t n = a.size();
    const FieldT omega_inv = omega ^ (uint64_t)(n - 1);  // omega^n = 1
    _basic_parallel_radix2_FFT_inner(a, omega_inv, log_cpus, FieldT::one());
    FieldT nf = FieldT::zero();
    {   // n as a field element: one() added n times via doubling
        FieldT acc = FieldT::one();
        for (size_t bit = 0; ((size_t)1 << bit) <= n; ++bit) {
            if (n & ((size_t)1 << bit)) nf = nf + acc;
            acc = acc + acc;
        }
    }
    const FieldT n_inv = nf.inverse();
#pragma omp parallel for
    for (size_t i = 0; i < n; ++i) a[i] = a[i] * n_inv;
}

}  // namespace oracle
