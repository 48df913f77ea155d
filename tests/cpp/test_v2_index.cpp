// Host model of the large-tile 768-bit pass kernel (gpusnarks_b200/csrc/ntt768.cuh, ntt768_pass2) over a small
// prime field, built from the kernel's own index maps (csrc/v2_index.h): ownership of the 1024 tile positions per
// warp in both phases, butterfly enumeration, unit-twiddle iterations, bank-conflict freedom of the shared-memory
// slots.  The eight warps are replayed one after the other in a random order between the synchronisation points the
// kernel has (none inside phase A, one __syncthreads before phase B), so a butterfly that touched an element of
// another warp's block would read a stale value and the comparison with the DFT definition would fail.
//
//   g++ -O2 -std=c++17 -I gpusnarks_b200/csrc tests/cpp/test_v2_index.cpp -o tests/cpp/test_v2_index && tests/cpp/test_v2_index
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>
#include <vector>

#include "v2_index.h"

static const uint64_t P = 2013265921ull;  // 15 * 2^27 + 1
static uint64_t mulm(uint64_t a, uint64_t b) { return a * b % P; }
static uint64_t powm(uint64_t a, uint64_t e) { uint64_t r = 1; while (e) { if (e & 1) r = mulm(r, a); a = mulm(a, a); e >>= 1; } return r; }
static uint32_t brev(uint32_t x, uint32_t bits) { uint32_t r = 0; for (uint32_t i = 0; i < bits; ++i) { r = (r << 1) | (x & 1); x >>= 1; } return r; }

#define CHECK(c) do { if (!(c)) { printf("FAILED %s:%d: %s (lq=%u)\n", __FILE__, __LINE__, #c, lq); return 1; } } while (0)

static int run(uint32_t lq, std::mt19937 &rng, bool hybrid) {
    const uint32_t T = 1024, L = 1u << lq;
    const uint64_t wL = powm(31, (P - 1) >> lq);  // primitive 2^lq-th root
    // tile input: T/L sub-transforms, natural order x[slot][j]; the kernel places element j at position brev(j)
    std::vector<uint64_t> in(T), smem(T);
    for (auto &v : in) v = rng() % P;
    // ---- load: warp W fills the positions it owns in phase A
    std::vector<int> loaded(T, 0);
    for (uint32_t W = 0; W < 8; ++W)
        for (uint32_t i = 0; i < 128; ++i) {
            const uint32_t pos = gsn::own_a(W, i), slot = pos >> lq, j = brev(pos & (L - 1), lq);
            smem[pos] = in[slot * L + j];
            loaded[pos]++;
        }
    for (uint32_t e = 0; e < T; ++e) CHECK(loaded[e] == 1);
    // ---- stages
    auto run_stage_warp = [&](uint32_t W, uint32_t ph, const std::set<uint32_t> &owned, std::vector<int> &touched) -> int {
        for (uint32_t it = 0; it < 2; ++it) {
            // bank-conflict check per quarter warp: the 8 lanes of an LDS.128 wavefront must hit 8 distinct 16-byte groups
            for (uint32_t half = 0; half < 2; ++half)
                for (uint32_t qw = 0; qw < 4; ++qw) {
                    std::set<uint32_t> groups;
                    for (uint32_t l8 = 0; l8 < 8; ++l8) {
                        uint32_t lo, jj;
                        gsn::v2_butterfly(W, ph, (qw * 8 + l8) + 32 * it, lo, jj);
                        const uint32_t e = half ? lo + (1u << (ph - 1)) : lo;
                        groups.insert((gsn::slot_of(e) * 7) & 7);
                    }
                    CHECK(groups.size() == 8);
                }
            for (uint32_t lane = 0; lane < 32; ++lane) {
                uint32_t lo, jj;
                gsn::v2_butterfly(W, ph, lane + 32 * it, lo, jj);
                const uint32_t m = 1u << (ph - 1), hi = lo + m;
                CHECK(owned.count(lo) && owned.count(hi));
                CHECK(jj < m && (lo & m) == 0 && (lo & (m - 1)) == jj);
                CHECK(ph > lq || (lo >> lq) == (hi >> lq));
                if (gsn::v2_unit(ph, it)) CHECK(jj == 0);
                touched[lo]++; touched[hi]++;
                const uint64_t w = powm(wL, (uint64_t)jj << (lq - ph));   // w_{2^ph}^jj = w_L^(jj * L / 2^ph)
                const uint64_t u = smem[lo], t = gsn::v2_unit(ph, it) ? smem[hi] : mulm(smem[hi], w);
                smem[lo] = (u + t) % P;
                smem[hi] = (u + P - t) % P;
            }
        }
        return 0;
    };
    // hybrid variant: stage ph (3 or 4) enumerated CTA-wide and twiddle-major; the 8 warps are replayed in a random order
    auto run_stage_cta = [&](uint32_t ph, std::vector<int> &touched, std::vector<uint32_t> worder) -> int {
        for (uint32_t W : worder)
            for (uint32_t it = 0; it < 2; ++it) {
                uint32_t jj_warp = ~0u;
                for (uint32_t half = 0; half < 2; ++half)
                    for (uint32_t qw = 0; qw < 4; ++qw) {
                        std::set<uint32_t> groups;
                        for (uint32_t l8 = 0; l8 < 8; ++l8) {
                            uint32_t lo, jj;
                            gsn::cta_butterfly(ph, (W * 32 + qw * 8 + l8) + 256 * it, lo, jj);
                            groups.insert((gsn::slot_of(half ? lo + (1u << (ph - 1)) : lo) * 7) & 7);
                        }
                        CHECK(groups.size() == 8);
                    }
                for (uint32_t lane = 0; lane < 32; ++lane) {
                    uint32_t lo, jj;
                    gsn::cta_butterfly(ph, (W * 32 + lane) + 256 * it, lo, jj);
                    const uint32_t m = 1u << (ph - 1), hi = lo + m;
                    if (jj_warp == ~0u) jj_warp = jj;
                    CHECK(jj == jj_warp);   // warp uniform: a jj == 0 warp skips the product as a whole
                    CHECK(hi < T && jj < m && (lo & m) == 0 && (lo & (m - 1)) == jj);
                    touched[lo]++; touched[hi]++;
                    const uint64_t w = powm(wL, (uint64_t)jj << (lq - ph));
                    const uint64_t u = smem[lo], t = jj == 0 ? smem[hi] : mulm(smem[hi], w);
                    smem[lo] = (u + t) % P;
                    smem[hi] = (u + P - t) % P;
                }
            }
        return 0;
    };
    std::vector<uint32_t> order = {0, 1, 2, 3, 4, 5, 6, 7};
    // phase A: each warp runs ALL its phase-A stages before the next warp starts (no cross-warp sync exists there)
    std::shuffle(order.begin(), order.end(), rng);
    const uint32_t endA = std::min(lq, 7u);
    std::vector<std::vector<int>> touched(lq + 1, std::vector<int>(T, 0));
    auto phase_a = [&](uint32_t from, uint32_t to) -> int {
        std::shuffle(order.begin(), order.end(), rng);
        for (uint32_t W : order) {
            std::set<uint32_t> owned;
            for (uint32_t i = 0; i < 128; ++i) owned.insert(gsn::own_a(W, i));
            CHECK(owned.size() == 128);
            for (uint32_t ph = from; ph <= to; ++ph) if (run_stage_warp(W, ph, owned, touched[ph])) return 1;
        }
        return 0;
    };
    if (!hybrid) {
        if (phase_a(1, endA)) return 1;
    } else {   // stages 1-2 per warp | barrier | stage 3 CTA-wide | barrier | stage 4 CTA-wide | barrier | stages 5-7 per warp
        if (phase_a(1, std::min(endA, 2u))) return 1;
        for (uint32_t ph = 3; ph <= std::min(lq, 4u); ++ph) {
            std::shuffle(order.begin(), order.end(), rng);
            if (run_stage_cta(ph, touched[ph], order)) return 1;
        }
        if (endA >= 5 && phase_a(5, endA)) return 1;
    }
    // __syncthreads, then phase B the same way
    std::shuffle(order.begin(), order.end(), rng);
    std::vector<int> ownedB(T, 0);
    for (uint32_t W : order) {
        std::set<uint32_t> owned;
        for (uint32_t i = 0; i < 128; ++i) { owned.insert(gsn::own_b(W, i)); ownedB[gsn::own_b(W, i)]++; }
        CHECK(owned.size() == 128);
        for (uint32_t ph = 8; ph <= lq; ++ph) if (run_stage_warp(W, ph, owned, touched[ph])) return 1;
    }
    for (uint32_t e = 0; e < T; ++e) CHECK(ownedB[e] == 1);
    for (uint32_t ph = 1; ph <= lq; ++ph)
        for (uint32_t e = 0; e < T; ++e) CHECK(touched[ph][e] == 1);
    // ---- compare every sub-transform with the DFT definition: out[slot][k] = sum_j x[slot][j] w_L^(jk)
    for (uint32_t slot = 0; slot < T / L; ++slot)
        for (uint32_t k = 0; k < L; k += (L > 64 ? 37 : 1)) {
            uint64_t acc = 0;
            for (uint32_t j = 0; j < L; ++j) acc = (acc + mulm(in[slot * L + j], powm(wL, (uint64_t)j * k))) % P;
            CHECK(acc == smem[slot * L + k]);
        }
    return 0;
}

// bank groups of the CTA-wide kernel's enumeration (ntt768_pass, 1024-element tile, 256 threads): twiddle-major in stages
// 2..5, twiddle-minor elsewhere; every quarter warp of an LDS.128 / STS.128 must touch 8 distinct 16-byte groups
static int check_cta_wide_kernel_banks() {
    const uint32_t lq = 10;
    for (uint32_t ph = 1; ph <= 10; ++ph)
        for (uint32_t b0 = 0; b0 < 512; b0 += 8)
            for (uint32_t half = 0; half < 2; ++half) {
                std::set<uint32_t> groups;
                for (uint32_t l8 = 0; l8 < 8; ++l8) {
                    const uint32_t b = b0 + l8, m = 1u << (ph - 1);
                    uint32_t jj, grp;
                    if (ph >= 2 && ph <= 5) { jj = b >> (10 - ph); grp = b & ((1u << (10 - ph)) - 1); }
                    else { jj = b & (m - 1); grp = b >> (ph - 1); }
                    const uint32_t lo = (grp << ph) | jj;
                    groups.insert((gsn::slot_of(half ? lo + m : lo) * 7) & 7);
                }
                CHECK(groups.size() == 8);
            }
    return 0;
}

int main() {
    std::mt19937 rng(12345);
    if (check_cta_wide_kernel_banks()) return 1;
    for (int rep = 0; rep < 3; ++rep)
        for (uint32_t lq = 1; lq <= 10; ++lq)
            if (run(lq, rng, false) || run(lq, rng, true)) return 1;
    printf("test_v2_index: ok (ownership, enumeration, unit iterations, bank groups, DFT parity for lq = 1..10, warp-owned and hybrid schedules)\n");
    return 0;
}
