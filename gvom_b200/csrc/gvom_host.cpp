// gvom_host.cpp -- see gvom_host.h.  Compiled by the host compiler (intrinsics), linked into libgvom_b200.so.
#include "gvom_host.h"

#include <emmintrin.h>

#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

// Streaming (non-temporal) copy.  Ordinary stores leave the data dirty in the caches of whichever cores
// copied it, and the GPU's PCIe reads of such lines are ~2.5x slower (measured on the B200 box: the zero-copy
// ray-cast took 390 us after a 4-thread memcpy, 157 us after a single-thread one).  NT stores put the lines
// straight into DRAM.
static void stream_copy(char* dst, const char* src, size_t n) {
    size_t head = (16 - ((uintptr_t)dst & 15)) & 15;
    if (head > n) head = n;
    memcpy(dst, src, head);
    dst += head; src += head; n -= head;
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {
        const __m128i a = _mm_loadu_si128((const __m128i*)(src + i)), b = _mm_loadu_si128((const __m128i*)(src + i + 16));
        const __m128i c = _mm_loadu_si128((const __m128i*)(src + i + 32)), d = _mm_loadu_si128((const __m128i*)(src + i + 48));
        _mm_stream_si128((__m128i*)(dst + i), a); _mm_stream_si128((__m128i*)(dst + i + 16), b);
        _mm_stream_si128((__m128i*)(dst + i + 32), c); _mm_stream_si128((__m128i*)(dst + i + 48), d);
    }
    memcpy(dst + i, src + i, n - i);
    _mm_sfence();
}

struct CopyPool::Impl {
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable cv, done;
    char* dst = nullptr; const char* src = nullptr; size_t bytes = 0;
    int gen = 0, pending = 0;
    bool stop = false;

    void slice(int i, int parts) {
        const size_t per = ((bytes / parts) + 4095) & ~size_t(4095);
        const size_t a = std::min(bytes, per * i), b = std::min(bytes, per * (i + 1));
        if (b > a) stream_copy(dst + a, src + a, b - a);
    }
    void work(int i) {
        int seen = 0;
        for (;;) {
            { std::unique_lock<std::mutex> l(m); cv.wait(l, [&] { return gen != seen; }); seen = gen; if (stop) return; }
            slice(i, (int)th.size() + 1);
            { std::lock_guard<std::mutex> l(m); if (--pending == 0) done.notify_one(); }
        }
    }
};

CopyPool::CopyPool(int helpers) : p_(new Impl) {
    for (int i = 0; i < helpers; ++i) p_->th.emplace_back([this, i] { p_->work(i + 1); });
}

CopyPool::~CopyPool() {
    { std::lock_guard<std::mutex> l(p_->m); p_->stop = true; ++p_->gen; }
    p_->cv.notify_all();
    for (auto& t : p_->th) t.join();
    delete p_;
}

void CopyPool::copy(char* dst, const char* src, size_t bytes) {
    const int parts = (int)p_->th.size() + 1;
    if (bytes < (1u << 18) || parts == 1) { stream_copy(dst, src, bytes); return; }
    { std::lock_guard<std::mutex> l(p_->m); p_->dst = dst; p_->src = src; p_->bytes = bytes; p_->pending = parts - 1; ++p_->gen; }
    p_->cv.notify_all();
    p_->slice(0, parts);
    std::unique_lock<std::mutex> l(p_->m);
    p_->done.wait(l, [this] { return p_->pending == 0; });
}
