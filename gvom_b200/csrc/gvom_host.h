// gvom_host.h -- host-only helpers of the C-ABI library (no CUDA): threaded streaming copy
// into the pinned staging block used for pageable input clouds.
#pragma once
#include <cstddef>

class CopyPool {
public:
    explicit CopyPool(int helpers);
    ~CopyPool();
    // copy `bytes` from pageable `src` to pinned `dst` with non-temporal stores, split over the pool
    void copy(char* dst, const char* src, size_t bytes);
private:
    struct Impl;
    Impl* p_;
};
