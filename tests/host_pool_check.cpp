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
    printf("bad=%d\n", bad); return bad != 0;
}