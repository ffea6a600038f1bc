// gvom_host.h -- host-only helpers of the C-ABI library (no CUDA): threaded streaming copies
// into the pinned staging block used for pageable input clouds.
#pragma once
#include <cstddef>
#include <cstdint>

class CopyPool {
public:
    explicit CopyPool(int helpers);
    ~CopyPool();
    // copy `bytes` from pageable `src` to pinned `dst` with non-temporal stores, split over the pool
    void copy(char* dst, const char* src, size_t bytes);
    // PointCloud2 field extraction: n records of point_step bytes with float32 x / y / z at the given byte
    // offsets -> packed records in pinned `dst`: 16 bytes (x, y, z, 0 as float32) or, with as_double,
    // 24 bytes (x, y, z widened to float64).  Non-temporal stores, split over the pool.
    void extract_xyz(char* dst, const char* src, int64_t n, int point_step, int ox, int oy, int oz, bool as_double);
private:
    struct Impl;
    Impl* p_;
};
