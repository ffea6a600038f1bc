// gvom_host.cpp -- see gvom_host.h.  Compiled by the host compiler (intrinsics), linked into libgvom_b200.so.
#include "gvom_host.h"

#include <emmintrin.h>
#include <xmmintrin.h>

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

// PointCloud2 records -> packed xyz (see gvom_host.h); dst is 16-byte aligned (pinned staging block)
static void extract_range(char* dst, const char* src, int64_t first, int64_t last, int step, int ox, int oy, int oz,
                          bool as_double) {
    if (!as_double && oy == ox + 4 && oz == ox + 8 && ox + 16 <= step) {
        // x, y, z adjacent and followed by at least 4 more bytes of the record: one unaligned 16-byte load per point
        const __m128 keep = _mm_castsi128_ps(_mm_set_epi32(0, -1, -1, -1));
        for (int64_t i = first; i < last; ++i) {
            const __m128 v = _mm_loadu_ps(reinterpret_cast<const float*>(src + (size_t)i * step + ox));
            _mm_stream_ps(reinterpret_cast<float*>(dst + (size_t)i * 16), _mm_and_ps(v, keep));
        }
    } else if (!as_double) {
        for (int64_t i = first; i < last; ++i) {
            const char* q = src + (size_t)i * step;
            float x, y, z;
            memcpy(&x, q + ox, 4); memcpy(&y, q + oy, 4); memcpy(&z, q + oz, 4);
            _mm_stream_ps(reinterpret_cast<float*>(dst + (size_t)i * 16), _mm_set_ps(0.f, z, y, x));
        }
    } else {
        for (int64_t i = first; i < last; ++i) {
            const char* q = src + (size_t)i * step;
            float x, y, z;
            memcpy(&x, q + ox, 4); memcpy(&y, q + oy, 4); memcpy(&z, q + oz, 4);
            double* d = reinterpret_cast<double*>(dst + (size_t)i * 24);
            _mm_stream_si64(reinterpret_cast<long long*>(d), _mm_cvtsi128_si64(_mm_castpd_si128(_mm_set_sd((double)x))));
            _mm_stream_si64(reinterpret_cast<long long*>(d + 1), _mm_cvtsi128_si64(_mm_castpd_si128(_mm_set_sd((double)y))));
            _mm_stream_si64(reinterpret_cast<long long*>(d + 2), _mm_cvtsi128_si64(_mm_castpd_si128(_mm_set_sd((double)z))));
        }
    }
    _mm_sfence();
}

struct CopyPool::Impl {
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable cv, done;
    char* dst = nullptr; const char* src = nullptr; size_t bytes = 0;
    // extraction job (mode 1) instead of a plain copy (mode 0)
    int mode = 0;
    int64_t n = 0; int step = 0, ox = 0, oy = 0, oz = 0; bool as_double = false;
    int gen = 0, pending = 0, active = 1;   // active: threads (incl. the caller) that take a slice of this job
    bool stop = false;

    void slice(int i, int parts) {
        if (i >= parts) return;
        if (mode == 1) {
            const int64_t per = (((n + parts - 1) / parts) + 255) & ~int64_t(255);   // ceil: per * parts >= n
            const int64_t a = std::min<int64_t>(n, per * i), b = std::min<int64_t>(n, per * (i + 1));
            if (b > a) extract_range(dst, src, a, b, step, ox, oy, oz, as_double);
            return;
        }
        const size_t per = (((bytes + parts - 1) / parts) + 4095) & ~size_t(4095);
        const size_t a = std::min(bytes, per * i), b = std::min(bytes, per * (i + 1));
        if (b > a) stream_copy(dst + a, src + a, b - a);
    }
    void work(int i) {
        int seen = 0;
        for (;;) {
            { std::unique_lock<std::mutex> l(m); cv.wait(l, [&] { return gen != seen; }); seen = gen; if (stop) return; }
            slice(i, active);
            { std::lock_guard<std::mutex> l(m); if (--pending == 0) done.notify_one(); }
        }
    }
    void run(int parts) {       // caller thread takes slice 0
        { std::lock_guard<std::mutex> l(m); active = parts; pending = (int)th.size(); ++gen; }
        cv.notify_all();
        slice(0, parts);
        std::unique_lock<std::mutex> l(m);
        done.wait(l, [this] { return pending == 0; });
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
    // a plain copy saturates the memory system with 4 threads (measured); the extraction below scales further
    const int parts = std::min(4, (int)p_->th.size() + 1);
    if (bytes < (1u << 18) || parts == 1) { stream_copy(dst, src, bytes); return; }
    p_->mode = 0; p_->dst = dst; p_->src = src; p_->bytes = bytes;
    p_->run(parts);
}

void CopyPool::extract_xyz(char* dst, const char* src, int64_t n, int point_step, int ox, int oy, int oz, bool as_double) {
    const int parts = (int)p_->th.size() + 1;
    if (n < 16384 || parts == 1) { extract_range(dst, src, 0, n, point_step, ox, oy, oz, as_double); return; }
    p_->mode = 1; p_->dst = dst; p_->src = src; p_->n = n; p_->step = point_step; p_->ox = ox; p_->oy = oy; p_->oz = oz;
    p_->as_double = as_double;
    p_->run(parts);
}
