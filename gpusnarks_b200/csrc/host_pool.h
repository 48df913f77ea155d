// host_pool.h -- a few persistent host threads for the staging copies of the pageable host-pointer paths
// (gsn_lib.cu: column blocks of a std::vector <-> pinned bounce buffers).  Plain C++, no CUDA: tests/cpp/test_host_pool.cpp
// exercises it on the CPU.
#pragma once
#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include <cstring>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

// memcpy whose stores bypass the caches (non-temporal) when both pointers are 16-byte aligned and the piece is large:
// the staging copies move hundreds of megabytes that the CPU does not read again soon, and a plain store first READS
// the destination line from DRAM (read for ownership) -- a third of the memory traffic of the copy.
inline void stream_copy(void *dst, const void *src, size_t bytes) {
#if defined(__SSE2__)
    if (bytes >= 4096 && (((uintptr_t)dst | (uintptr_t)src) & 15) == 0) {
        const __m128i *s = (const __m128i *)src;
        __m128i *d = (__m128i *)dst;
        const size_t n = bytes / 64;
        for (size_t i = 0; i < n; ++i) {
            const __m128i a = _mm_load_si128(s + 4 * i), b = _mm_load_si128(s + 4 * i + 1), c = _mm_load_si128(s + 4 * i + 2), e = _mm_load_si128(s + 4 * i + 3);
            _mm_stream_si128(d + 4 * i, a);
            _mm_stream_si128(d + 4 * i + 1, b);
            _mm_stream_si128(d + 4 * i + 2, c);
            _mm_stream_si128(d + 4 * i + 3, e);
        }
        if (bytes % 64) memcpy((char *)dst + n * 64, (const char *)src + n * 64, bytes % 64);
        _mm_sfence();
        return;
    }
#endif
    memcpy(dst, src, bytes);
}

// A few persistent host threads that move data between a caller's pageable memory and the pinned bounce buffers
// (run() blocks; the calling thread works as thread 0).
struct HostPool {
    std::vector<std::thread> threads;
    std::mutex m;
    std::condition_variable cv, done_cv;
    const std::function<void(unsigned, unsigned)> *job = nullptr;
    uint64_t epoch = 0;
    unsigned pending = 0, nt = 1;
    bool stop = false;
    explicit HostPool(unsigned n) : nt(std::max(1u, n)) {
        for (unsigned t = 1; t < nt; ++t)
            threads.emplace_back([this, t] {
                uint64_t seen = 0;
                for (;;) {
                    const std::function<void(unsigned, unsigned)> *fn;
                    {
                        std::unique_lock<std::mutex> lk(m);
                        cv.wait(lk, [&] { return stop || epoch != seen; });
                        if (stop) return;
                        seen = epoch;
                        fn = job;
                    }
                    (*fn)(t, nt);
                    {
                        std::lock_guard<std::mutex> lk(m);
                        if (--pending == 0) done_cv.notify_one();
                    }
                }
            });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(m);
            stop = true;
        }
        cv.notify_all();
        for (auto &t : threads) t.join();
    }
    void run(const std::function<void(unsigned, unsigned)> &fn) {
        if (nt > 1) {
            std::lock_guard<std::mutex> lk(m);
            job = &fn;
            pending = nt - 1;
            ++epoch;
        }
        cv.notify_all();
        fn(0, nt);
        if (nt > 1) {
            std::unique_lock<std::mutex> lk(m);
            done_cv.wait(lk, [&] { return pending == 0; });
        }
    }
};
