// v2_index.h -- index maps of the large-tile 768-bit pass kernel (ntt768_pass2), shared between the CUDA kernel and
// the host-side model test (tests/cpp/test_v2_index.cpp), which replays a whole tile over a small prime field with
// exactly these functions and compares every sub-transform with the DFT definition.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define GSN_HD __host__ __device__ __forceinline__
#else
#define GSN_HD inline
#endif

namespace gsn {

// Shared-memory slot of tile element e.  XOR-ing the low three bits with bits 3..5 and bits 6..8 keeps eight consecutive
// elements on eight distinct 16-byte bank groups (the common case) and also makes the stride-2, -4, -8 and -16 element
// patterns of the twiddle-major early stages conflict free (round 1 mixed in bits 3..5 only: stride 16 was 2-way).
GSN_HD uint32_t slot_of(uint32_t e) { return e ^ (((e >> 3) ^ (e >> 6)) & 7u); }

// Tile = 1024 element positions, 8 warps.  Element i (< 128) owned by warp W:
//   phase A  (stages 1..7):   128 consecutive positions
//   phase B  (stages 8..10):  positions (h << 7) | (W << 4) | l, h < 8, l < 16 -- closed under the butterflies of
//                             bits 7, 8 and 9
GSN_HD uint32_t own_a(uint32_t W, uint32_t i) { return (W << 7) | i; }
GSN_HD uint32_t own_b(uint32_t W, uint32_t i) { return ((i >> 4) << 7) | (W << 4) | (i & 15); }

// butterfly `bl` (< 64) of warp W in stage ph (1..10): low element position and twiddle exponent jj; the high
// element is lo + 2^(ph-1), the twiddle w_{2^ph}^jj
GSN_HD void v2_butterfly(uint32_t W, uint32_t ph, uint32_t bl, uint32_t &lo, uint32_t &jj) {
    const uint32_t m = 1u << (ph - 1);
    if (ph <= 7) {
        uint32_t grp;
        // stages 1..3 twiddle-major (consecutive lanes = consecutive groups, stride 2^ph elements: conflict free under
        // slot_of, and iteration 0 of stage 2 is all unit twiddles); later stages twiddle-minor (consecutive elements)
        if (ph <= 3) { jj = bl >> (7 - ph); grp = bl & ((1u << (7 - ph)) - 1); }
        else { jj = bl & (m - 1); grp = bl >> (ph - 1); }
        lo = (W << 7) | (grp << ph) | jj;
    } else {
        const uint32_t b = ph - 8, hp = bl >> 4, l = bl & 15;
        const uint32_t h = ((hp >> b) << (b + 1)) | (hp & ((1u << b) - 1));
        lo = (h << 7) | (W << 4) | l;
        jj = lo & (m - 1);
    }
}
// CTA-wide twiddle-major enumeration of stage ph (hybrid variant, stages 3 and 4): butterfly b (< 512) of the tile.
// jj = b >> (10 - ph) is constant over 128 / 64 consecutive b, i.e. over whole warps: the jj == 0 warps skip the product.
GSN_HD void cta_butterfly(uint32_t ph, uint32_t b, uint32_t &lo, uint32_t &jj) {
    jj = b >> (10 - ph);
    const uint32_t grp = b & ((1u << (10 - ph)) - 1);
    lo = (grp << ph) | jj;
}
// iteration `it` of stage ph skips the product (all its twiddles are 1)
GSN_HD bool v2_unit(uint32_t ph, uint32_t it) { return ph == 1 || (ph == 2 && it == 0); }

}  // namespace gsn
