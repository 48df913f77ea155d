// HostPool (gpusnarks_b200/csrc/host_pool.h): every thread index runs exactly once per run(), run() returns only after
// all of them have finished, back-to-back runs do not lose or repeat work, and a strided copy split over the pool
// equals a plain memcpy.
//   g++ -O2 -std=c++17 -pthread -I gpusnarks_b200/csrc tests/cpp/test_host_pool.cpp -o /tmp/test_host_pool && /tmp/test_host_pool
#include <atomic>
#include <cstdio>
#include <cstring>

#include "host_pool.h"

#define CHECK(c) do { if (!(c)) { printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main() {
    for (unsigned nt : {1u, 2u, 3u, 8u}) {
        HostPool pool(nt);
        std::vector<std::atomic<int>> hits(nt);
        for (auto &h : hits) h = 0;
        for (int rep = 0; rep < 2000; ++rep) {
            std::atomic<unsigned> done{0};
            pool.run([&](unsigned t, unsigned n) {
                if (n == nt && t < nt) hits[t]++;
                done++;
            });
            CHECK(done.load() == nt);   // run() is a barrier: nothing is still in flight when it returns
        }
        for (unsigned t = 0; t < nt; ++t) CHECK(hits[t].load() == 2000);
        // pitched copy: rows split over the pool
        const size_t rows = 257, width = 1000, spitch = 1536, dpitch = 1024;
        std::vector<unsigned char> src(rows * spitch), dst(rows * dpitch, 0), want(rows * dpitch, 0);
        for (size_t i = 0; i < src.size(); ++i) src[i] = (unsigned char)(i * 131 + 7);
        for (size_t r = 0; r < rows; ++r) memcpy(&want[r * dpitch], &src[r * spitch], width);
        pool.run([&](unsigned t, unsigned n) {
            for (size_t r = rows * t / n; r < rows * (t + 1) / n; ++r) memcpy(&dst[r * dpitch], &src[r * spitch], width);
        });
        CHECK(dst == want);
    }
    // stream_copy: aligned large pieces (non-temporal stores), odd sizes and unaligned pointers (plain memcpy) all copy exactly
    {
        std::vector<unsigned char> raw_src(1 << 20), raw_dst(1 << 20);
        for (size_t i = 0; i < raw_src.size(); ++i) raw_src[i] = (unsigned char)(i * 29 + 3);
        unsigned char *s0 = raw_src.data() + ((16 - (uintptr_t)raw_src.data() % 16) % 16), *d0 = raw_dst.data() + ((16 - (uintptr_t)raw_dst.data() % 16) % 16);
        for (size_t bytes : {(size_t)96, (size_t)4096, (size_t)12288, (size_t)12288 + 32, (size_t)100000, (size_t)100001})
            for (size_t so : {(size_t)0, (size_t)16, (size_t)5})
                for (size_t dof : {(size_t)0, (size_t)32, (size_t)7}) {
                    memset(raw_dst.data(), 0, raw_dst.size());
                    stream_copy(d0 + dof, s0 + so, bytes);
                    CHECK(memcmp(d0 + dof, s0 + so, bytes) == 0);
                    CHECK(d0[dof + bytes] == 0);
                }
    }
    printf("test_host_pool: ok\n");
    return 0;
}
