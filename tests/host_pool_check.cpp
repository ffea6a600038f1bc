// CPU check of the staging helpers (gvom_b200/csrc/gvom_host.cpp): threaded copy and PointCloud2 field extraction
// against plain memcpy semantics.  Built and run by tests/test_host_helpers.py.

#include "gvom_host.h"
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>
int main() {
    const int steps[4] = {12, 16, 48, 32}; int bad = 0;
    for (int si = 0; si < 4; ++si) for (int ox = 0; ox <= 8; ox += 4) {
        const int step = steps[si]; if (ox + 12 > step) continue;
        const int gap = (si == 3) ? 8 : 4;                 // step 32: x / y / z fields 8 bytes apart (not adjacent)
        const int64_t n = 100000;
        std::vector<char> in((size_t)n * step);
        for (size_t i = 0; i < in.size(); ++i) in[i] = (char)(i * 7 + 3);
        char* out = (char*)aligned_alloc(256, (size_t)n * 24); char* out2 = (char*)aligned_alloc(256, (size_t)n * 24);
        CopyPool pool(5);
        pool.extract_xyz(out, in.data(), n, step, ox, ox + gap, ox + 2 * gap, false);
        pool.extract_xyz(out2, in.data(), n, step, ox, ox + gap, ox + 2 * gap, true);
        for (int64_t i = 0; i < n; ++i) {
            float f[4]; memcpy(f, out + i * 16, 16); double d[3]; memcpy(d, out2 + i * 24, 24);
            for (int k = 0; k < 3; ++k) { float w; memcpy(&w, in.data() + i * step + ox + gap * k, 4);
                if (memcmp(&w, &f[k], 4) != 0) ++bad;
                double wd = (double)w; if (memcmp(&wd, &d[k], 8) != 0 && w == w) ++bad; }
            uint32_t z; memcpy(&z, &f[3], 4); if (z != 0) ++bad;
        }
        std::vector<char> c((size_t)n * step + 64); pool.copy(c.data() + ((64 - ((uintptr_t)c.data() & 63)) & 63), in.data(), in.size());
        if (memcmp(c.data() + ((64 - ((uintptr_t)c.data() & 63)) & 63), in.data(), in.size()) != 0) ++bad;
        free(out); free(out2);
    }

    // slicing regression: n / parts already a multiple of 256 with a remainder -- the last n % parts
    // records used to be skipped (n = 16387, 4 slices covered 16384); same for byte copies with 3 slices
    for (int helpers = 2; helpers <= 7; ++helpers) {
        const int64_t ns[4] = {16387, 16384 * 5 + 3, 65536 + 7, 262144 + 5};
        for (int t = 0; t < 4; ++t) {
            const int64_t n = ns[t]; const int step = 48;
            std::vector<char> in((size_t)n * step);
            for (size_t i = 0; i < in.size(); ++i) in[i] = (char)(i * 13 + 1);
            char* out = (char*)aligned_alloc(256, (size_t)n * 16 + 256);
            memset(out, 0x5a, (size_t)n * 16 + 256);
            CopyPool pool(helpers);
            pool.extract_xyz(out, in.data(), n, step, 0, 4, 8, false);
            for (int64_t i = 0; i < n; ++i)
                if (memcmp(out + i * 16, in.data() + i * step, 12) != 0) ++bad;
            free(out);
            const size_t bytes = (size_t)4096 * 3 * 40 + 4096 * 2 + 17;       // floor(bytes / 3) a multiple of 4096, remainder != 0
            std::vector<char> src(bytes), dst(bytes + 64, (char)0x5a);
            for (size_t i = 0; i < bytes; ++i) src[i] = (char)(i * 31 + 7);
            char* d = dst.data() + ((64 - ((uintptr_t)dst.data() & 63)) & 63);
            pool.copy(d, src.data(), bytes);
            if (memcmp(d, src.data(), bytes) != 0) ++bad;
        }
    }
    printf("bad=%d\n", bad); return bad != 0;
}