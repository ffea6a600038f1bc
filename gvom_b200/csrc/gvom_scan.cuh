// gvom_scan.cuh -- Process_pointcloud on sm_100a, second build: TWO kernels per scan.
//
//   S1 k_scan_points   per point: transform, voxelise, claim a cell for the voxel on its first hit, accumulate
//                      the point's raw moments in that cell, ray-cast.  Replaces K1 + K3 of the first build and
//                      the compaction half of K2.
//   S2 k_scan_cells    per cell: finalise counts, neighbourhood gather, normalise; beside it (other warps of the
//                      same launch) one streaming pass that derives the slot's group mask and wipes the spare
//                      slot for the next scan.  Replaces K4 and the dense pass of K2.
//
// Results are those of __transform_pointcloud, __point_2_map, __assign_indices, __move_data, __calculate_mean /
// __normalize_mean / __calculate_covariance / __normalize_covariance and __calculate_min_height
// (gvom.py:1121-1421) -- the arithmetic contract of gvom_kernels.cuh applies unchanged.
//
// What is different from the first build:
//   * no dense hit / pass grids and no pass over them: rays decrement the slot's OWN index map, which starts as
//     "all unknown" (-1): after S1 a free voxel already holds its final code -1 - passes (gvom.py:1244-1247)
//   * cell ids are claimed on the fly in an extended, epoch-tagged voxel -> cell grid ((S+2rx)^2 (Z+2rz) words,
//     never cleared between scans: an entry is live only if its tag is this scan's), so the moments of a point go
//     straight into its cell while the point is still in registers -- the cloud is read once
//   * points whose own voxel lies just outside the grid still reach in-grid neighbours in the reference
//     (gvom.py:1262-1279): they claim "ghost" cells in the margin of the extended grid and the gather picks them
//     up like any other neighbour -- no special apron accumulators, no bounds checks in the gather
//   * the ring holds B + 1 physical slots: the one that just dropped out of the ring is wiped (under its own group
//     mask) by S2 of the scan that replaced it, off the critical path
#pragma once
#include "gvom_kernels.cuh"
#include "gvom_merge.cuh"      // mbarrier / bulk-copy helpers

namespace gvom {

constexpr unsigned CELL_PAY = (1u << 23) - 1u;        // payload bits of a cell-grid entry
constexpr unsigned CELL_GHOST = 1u << 23;             // the cell belongs to a margin voxel (no map entry)
constexpr unsigned CELL_PENDING = CELL_PAY;           // claimed, id not published yet
constexpr unsigned CELL_OVERFLOW = CELL_PAY - 1u;     // claimed, but the compact arrays are full
constexpr int CELL_TAG_SHIFT = 24;                    // bits 24..31: scan tag 1..255 (0: never used)
constexpr int MOM = 10;                               // raw moments per cell: S(3), Q(6), n

struct ScanOut {
    int* map;                 // the slot's index map, -1 everywhere on entry
    unsigned* cellid;         // extended voxel -> cell grid
    unsigned tag;             // this scan's tag
    int* counters;            // [0] cells claimed, [1] ghost cells claimed
    double* acc;              // [cap + gcap][MOM] raw moments about the voxel centre
    float* minh;              // [cap]
    int* cell_voxel;          // [cap]
    int cap, gcap;
    int ES;                   // S + 2 rx (extended row length)
};

__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned* p) { return *reinterpret_cast<const volatile unsigned*>(p); }
// Fire-and-forget reductions (RED): nvcc emitted the returning form (ATOMG ..., RZ) for every atomic of this kernel;
// the explicit PTX keeps the return path out of the atomic-bound regime (dense stress configuration).
__device__ __forceinline__ void red_add(int* p, int v) {
    asm volatile("red.global.add.s32 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add(double* p, double v) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "d"(v) : "memory");
}
__device__ __forceinline__ void red_min(int* p, int v) {
    asm volatile("red.global.min.s32 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "r"(v) : "memory");
}

// One world-frame point per lane: hit, cell claim, own-voxel moments, ray (gvom.py:1140-1231, 1249-1421).
template <bool FASTFLOOR>
__device__ __forceinline__ void scan_point(bool ok, double wx, double wy, double wz, const Frame& fr,
                                           const DevParams& P, const ScanOut& O) {
    const int lane = threadIdx.x & 31;
    const double ox = fr.origin[0], oy = fr.origin[1], oz = fr.origin[2];
    // ---- voxel of the point (gvom.py:1153-1171); cls 1: inside the grid, 2: in the margin of the extended grid
    double ex = 0, ey = 0, ez = 0, qx = 0, qy = 0, qz = 0;
    float lzf = 1.0f;
    int cls = 0, v = 0, vext = 0;
    if (ok) {
        ex = __ddiv_rn(wx, P.xy_res); ey = __ddiv_rn(wy, P.xy_res); ez = __ddiv_rn(wz, P.z_res);
        const double fx = __dsub_rn(ex, ox), fy = __dsub_rn(ey, oy), fz = __dsub_rn(ez, oz);
        const double bx = floor(fx), by = floor(fy), bz = floor(fz);
        const double dS = (double)P.S, dZ = (double)P.Z, rx = (double)P.rx, rz = (double)P.rz;
        if (bx >= -rx && bx < dS + rx && by >= -rx && by < dS + rx && bz >= -rz && bz < dZ + rz) {
            const int xi = (int)bx, yi = (int)by, zi = (int)bz;
            const bool inb = ((unsigned)xi < (unsigned)P.S) && ((unsigned)yi < (unsigned)P.S) && ((unsigned)zi < (unsigned)P.Z);
            cls = inb ? 1 : 2;
            v = xi + (yi + zi * P.S) * P.S;
            vext = (xi + P.rx) + ((yi + P.rx) + (zi + P.rz) * O.ES) * O.ES;
            const double lz = __dsub_rn(fz, bz);
            qx = (fx - bx) - 0.5; qy = (fy - by) - 0.5; qz = lz - 0.5;     // about the voxel centre
            lzf = (float)lz;                                              // min height: float32 of the in-voxel z fraction
        }
    }
    // ---- runs of consecutive lanes in the same voxel (neighbouring azimuths): the first lane acts for the run
    const int key = cls ? vext : ~lane;
    const int prev_key = __shfl_up_sync(FULL, key, 1);
    const bool head = (lane == 0) || (prev_key != key);
    const unsigned heads = __ballot_sync(FULL, head);
    const unsigned above = heads & ~((2u << lane) - 1u);                  // heads after my lane
    const int run_end = above ? (__ffs(above) - 2) : 31;                  // last lane of my run
    const bool act = head && cls != 0;
    if (head && cls == 1) red_add(O.map + v, -(run_end - lane + 1));    // the hit counts as a pass too (gvom.py:1169)

    // ---- claim the voxel's cell (first hit of this scan) or find the id somebody else claimed.
    // Phase A: compare-and-swap the grid entry to "pending".  Phase B: the block's winners take their ids from ONE
    // atomic per block and counter (30 k same-address atomics would serialise for ~20 us at the L2 atomic unit:
    // measured 1.5 G/s) and publish them.  Nobody waits on another thread before
    // its own ids are published, so the waits of phase C cannot form a cycle.
    unsigned cur = 0;
    unsigned* e = O.cellid + vext;
    bool won = false;
    if (act) {
        cur = ld_volatile_u32(e);
        while ((cur >> CELL_TAG_SHIFT) != O.tag) {                        // stale entry of an earlier scan: try to take it
            const unsigned old = atomicCAS(e, cur, (O.tag << CELL_TAG_SHIFT) | CELL_PENDING);
            if (old == cur) { won = true; break; }
            cur = old;                                                    // lost the race: `old` carries this scan's tag
        }
    }
    {
        __shared__ int s_cnt[8][2];
        __shared__ int s_base[2];
        const int warp = threadIdx.x >> 5;
        const unsigned wr = __ballot_sync(FULL, won && cls == 1), wg = __ballot_sync(FULL, won && cls == 2);
        if (lane == 0) { s_cnt[warp][0] = __popc(wr); s_cnt[warp][1] = __popc(wg); }
        __syncthreads();
        if (threadIdx.x < 2) {
            int tot = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) tot += s_cnt[w][threadIdx.x];
            s_base[threadIdx.x] = tot ? atomicAdd(O.counters + threadIdx.x, tot) : 0;
        }
        __syncthreads();
        if (won) {
            const int t = cls == 2 ? 1 : 0;
            int id = s_base[t] + __popc((t ? wg : wr) & ((1u << lane) - 1u));
            for (int w = 0; w < warp; ++w) id += s_cnt[w][t];
            unsigned pay = CELL_OVERFLOW;
            if (id < (t ? O.gcap : O.cap)) {
                // the accumulator row and the min-height entry are already clear: S2 of the previous scan reset what
                // the scan before used (double-buffered), so publishing needs no memory fence -- a fence anywhere in
                // this kernel makes ptxas emit every later atomic in its returning form (ATOMG) instead of RED
                if (!t) O.cell_voxel[id] = v;
                pay = (unsigned)id;
            }
            cur = (O.tag << CELL_TAG_SHIFT) | (t ? CELL_GHOST : 0u) | pay;
            *reinterpret_cast<volatile unsigned*>(e) = cur;
        }
    }
    __syncwarp();                // phase C: every claim of this warp is published; the waits below are on other warps only
    // ---- own-voxel raw moments, reduced over the run, one component at a time (register light)
    double* row = nullptr;
    bool real = false;
    int id = 0;
    if (act) {
        // the claimer publishes right after its counter atomic returns and never waits on anybody, so this is short;
        // the bound only keeps a protocol bug from hanging the GPU (the cell would then be dropped like an overflow)
        for (int spin = 0; (cur & CELL_PAY) == CELL_PENDING && spin < (1 << 20); ++spin) cur = ld_volatile_u32(e);
        const unsigned pay = cur & CELL_PAY;
        if (pay < CELL_OVERFLOW) {
            real = (cur & CELL_GHOST) == 0u;
            id = (int)pay;
            row = O.acc + (long long)(real ? id : O.cap + id) * MOM;
        }
    }
    const int max_d = __reduce_max_sync(FULL, run_end - lane);            // longest run of this warp (runs are short)
#pragma unroll
    for (int k = 0; k < MOM; ++k) {
        double val = k == 0 ? qx : k == 1 ? qy : k == 2 ? qz : k == 3 ? qx * qx : k == 4 ? qx * qy : k == 5 ? qx * qz
                   : k == 6 ? qy * qy : k == 7 ? qy * qz : k == 8 ? qz * qz : 1.0;
        if (!cls) val = 0.0;
        for (int off = 1; off <= max_d; off <<= 1) {
            const double t = __shfl_down_sync(FULL, val, off);
            if (lane + off <= run_end) val += t;
        }
        if (row) red_add(row + k, val);
    }
    {
        for (int off = 1; off <= max_d; off <<= 1) {
            const float t = __shfl_down_sync(FULL, lzf, off);
            if (lane + off <= run_end) lzf = fminf(lzf, t);
        }
        if (row && real) red_min(reinterpret_cast<int*>(O.minh) + id, __float_as_int(lzf));   // values in [0,1]: ordered as int bits
    }

    // ---- ray set-up (gvom.py:1174-1207): float32 state, float64 length
    float px = fr.start[0], py = fr.start[1], pz = fr.start[2];
    float ix = 0.f, iy = 0.f, iz = 0.f;
    double dlen = 0.0, lim = 0.0, length = 0.0;
    bool active = false;
    if (ok) {
        float sx = __fsub_rn((float)ex, px), sy = __fsub_rn((float)ey, py), sz = __fsub_rn((float)ez, pz);
        float l2 = __fmul_rn(sx, sx);
        l2 = __fmaf_rn(sy, sy, l2);
        l2 = __fmaf_rn(sz, sz, l2);
        const float L = __fsqrt_rn(l2);
        sx = __fdiv_rn(sx, L); sy = __fdiv_rn(sy, L); sz = __fdiv_rn(sz, L);
        const float a0 = fabsf(sx), a1 = fabsf(sy), a2 = fabsf(sz);
        const float m = fmaxf(a0, fmaxf(a1, a2));
        float sk = sx;                                   // dominant axis; later axis wins ties
        if (m == a1) sk = sy;
        if (m == a2) sk = sz;
        lim = __dadd_rn((double)L, -1.0);
        if (lim > 0.0) {
            active = true;
            const float ak = fabsf(sk);
            ix = __fdiv_rn(sx, ak); iy = __fdiv_rn(sy, ak); iz = __fdiv_rn(sz, ak);
            dlen = fabs(__drcp_rn((double)sk));
        }
    }
    // ---- DDA (gvom.py:1208-1231), warp-synchronous and branch-free; exactness argument in gvom_kernels.cuh
    // (raycast_point).  FASTFLOOR: floorf(p) for |p| < 2^22 is the low mantissa of p + 1.5 * 2^23 rounded DOWN
    // (one FADD.RM on the FP32 pipe instead of a quarter-rate F2I.FLOOR), the host checks the range.
    const int iox = fr.io[0], ioy = fr.io[1], ioz = fr.io[2];
    constexpr int MAGIC_BITS = 0x4B400000;               // bit pattern of 12582912.0f
    const int cx = FASTFLOOR ? MAGIC_BITS + iox : iox, cy = FASTFLOOR ? MAGIC_BITS + ioy : ioy, cz = FASTFLOOR ? MAGIC_BITS + ioz : ioz;
    if (!active) { ix = 0.f; iy = 0.f; iz = 0.f; dlen = 0.0; }
    if (FASTFLOOR && __all_sync(FULL, active)) {
        // Bulk phase: the first n_safe steps of EVERY ray of the warp stay strictly inside the grid and below the
        // length limit, so they need no bounds / length checks and no per-lane activity handling.  n_safe is a
        // conservative closed-form bound (positions accumulate < 0.02 voxel of float32 rounding over 1000 steps; the
        // bound keeps a 2-step margin); the arithmetic per step is unchanged, and the checked loop below finishes every
        // ray exactly as before.
        float kmin = 1.0e6f;
        if (ix > 0.f) kmin = fminf(kmin, ((float)(iox + P.S - 1) - px) / ix); else if (ix < 0.f) kmin = fminf(kmin, (px - (float)(iox + 1)) / -ix);
        if (iy > 0.f) kmin = fminf(kmin, ((float)(ioy + P.S - 1) - py) / iy); else if (iy < 0.f) kmin = fminf(kmin, (py - (float)(ioy + 1)) / -iy);
        if (iz > 0.f) kmin = fminf(kmin, ((float)(ioz + P.Z - 1) - pz) / iz); else if (iz < 0.f) kmin = fminf(kmin, (pz - (float)(ioz + 1)) / -iz);
        kmin = fminf(kmin, (float)(lim / dlen));
        int it = __reduce_min_sync(FULL, (int)fmaxf(kmin - 2.0f, 0.f));
        const unsigned above_me = ~((2u << lane) - 1u), from_me = 0xffffffffu << lane;
        const int l0 = lane == 0 ? (int)0x80000000 : 0;       // lane 0 always starts a run
#pragma unroll 1
        for (; it > 0; --it) {
            px = __fadd_rn(px, ix); py = __fadd_rn(py, iy); pz = __fadd_rn(pz, iz);
            const int bx = __float_as_int(__fadd_rd(px, 12582912.0f)), by = __float_as_int(__fadd_rd(py, 12582912.0f)),
                      bz = __float_as_int(__fadd_rd(pz, 12582912.0f));
            const int vv = (bx - fr.cc) + by * P.S + bz * fr.S2;
            const int p2 = __shfl_up_sync(FULL, vv, 1);
            const bool h2 = p2 != (vv ^ l0);
            const unsigned hs = __ballot_sync(FULL, h2);
            if (h2) {
                const unsigned ab = hs & above_me;
                const unsigned nx = ab & (0u - ab);
                red_add(O.map + vv, -__popc((nx - 1u) & from_me));
            }
            length = __dadd_rn(length, dlen);
        }
    }
    unsigned any = __ballot_sync(FULL, active);
    while (any) {
        px = __fadd_rn(px, ix); py = __fadd_rn(py, iy); pz = __fadd_rn(pz, iz);
        int x, y, z;
        if (FASTFLOOR) {
            x = __float_as_int(__fadd_rd(px, 12582912.0f)) - cx;
            y = __float_as_int(__fadd_rd(py, 12582912.0f)) - cy;
            z = __float_as_int(__fadd_rd(pz, 12582912.0f)) - cz;
        } else {
            x = __float2int_rd(px) - cx; y = __float2int_rd(py) - cy; z = __float2int_rd(pz) - cz;
        }
        const bool inside = active && ((unsigned)x < (unsigned)P.S) && ((unsigned)y < (unsigned)P.S) &&
                            ((unsigned)z < (unsigned)P.Z);
        const int vv = x + (y + z * P.S) * P.S;
        const int k2 = inside ? vv : ~lane;                       // finished lanes: unique negative keys
        const int p2 = __shfl_up_sync(FULL, k2, 1);
        const bool h2 = (lane == 0) || (p2 != k2);
        const unsigned hs = __ballot_sync(FULL, h2);
        if (inside && h2) {
            const unsigned ab = hs & ~((2u << lane) - 1u);        // heads after my lane
            const unsigned nx = ab & (0u - ab);                   // lowest of them (0: my run ends the warp)
            red_add(O.map + vv, -__popc((nx - 1u) & (0xffffffffu << lane)));
        }
        length = __dadd_rn(length, dlen);
        active = inside && (length < lim);
        any = __ballot_sync(FULL, active);
    }
}

// Stage a block's contiguous chunk of a pinned HOST cloud in shared memory (every byte crosses PCIe once, as full
// 128-bit coalesced reads) -- zero-copy input.
template <typename T>
__device__ __forceinline__ const T* stage_chunk(const T* __restrict__ pts, int stride, int n, uint4* chunk) {
    const long long first = (long long)blockIdx.x * blockDim.x;
    const int cnt = (int)min((long long)blockDim.x, (long long)n - first);
    const size_t bytes = (size_t)cnt * stride * sizeof(T);
    const char* g = reinterpret_cast<const char*>(pts + first * stride);
    const int n16 = (int)(bytes >> 4);
    for (int k = threadIdx.x; k < n16; k += blockDim.x) chunk[k] = __ldg(reinterpret_cast<const uint4*>(g) + k);
    if (threadIdx.x < (int)(bytes & 15))                  // tail bytes (none when the chunk is full)
        reinterpret_cast<char*>(chunk)[(n16 << 4) + threadIdx.x] = g[(n16 << 4) + threadIdx.x];
    __syncthreads();
    return reinterpret_cast<const T*>(chunk) + (long long)threadIdx.x * stride;
}

template <typename T, bool FASTFLOOR>
__global__ void __launch_bounds__(256, 8)
k_scan_points(const T* __restrict__ pts, int stride, int n, int from_host, Xform tf, Frame fr, DevParams P, ScanOut O) {
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double wx = 0, wy = 0, wz = 0;
    bool ok = false;
    if (from_host) {
        __shared__ __align__(128) uint4 chunk[256 * 4 * sizeof(double) / 16];
        const T* q;
        const long long first = (long long)blockIdx.x * blockDim.x;
        const int cnt = (int)min((long long)blockDim.x, (long long)n - first);
        const unsigned bytes = (unsigned)((size_t)cnt * stride * sizeof(T));
        if (from_host == 2 && (bytes & 15u) == 0u) {
            // the block's chunk is fetched from pinned host memory by ONE bulk copy of the TMA engine (258 vs 263 us per
            // end-to-end tick against 128-bit loads of every thread; GVOM_VARIANT bit 256 selects the loads)
            __shared__ unsigned long long bar;
            if (threadIdx.x == 0) {
                mbar_init(smem_u32(&bar), 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                mbar_expect_tx(smem_u32(&bar), bytes);
                bulk_g2s(smem_u32(chunk), pts + first * stride, bytes, smem_u32(&bar));
            }
            __syncthreads();
            mbar_wait(smem_u32(&bar), 0u);
            q = reinterpret_cast<const T*>(chunk) + (long long)threadIdx.x * stride;
        } else {
            q = stage_chunk<T>(pts, stride, n, chunk);
        }
        if (i < n) ok = world_from_raw<T>(q[0], q[1], q[2], tf, P.min_d2, wx, wy, wz);
    } else if (i < n) {
        ok = load_world<T>(pts, stride, i, tf, P.min_d2, wx, wy, wz);
    }
    scan_point<FASTFLOOR>(ok, wx, wy, wz, fr, P, O);
}

// S1 for PointCloud2 wire records (see k_voxelize_raycast_pc2): float32 x / y / z at byte offsets, widened to
// float64 (what ros_numpy hands the reference), NaN / Inf dropped; packed 16-byte records are one 128-bit load.
template <bool FASTFLOOR>
__global__ void __launch_bounds__(256, 8)
k_scan_points_pc2(const char* __restrict__ data, int point_step, int offx, int offy, int offz, int n, Xform tf,
                  Frame fr, DevParams P, ScanOut O) {
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double wx = 0, wy = 0, wz = 0;
    bool ok = false;
    if (i < n) {
        const char* q = data + (size_t)i * point_step;
        float x, y, z;
        if (point_step == 16 && offx == 0 && offy == 4 && offz == 8) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(q));
            x = t.x; y = t.y; z = t.z;
        } else {
            x = __ldg(reinterpret_cast<const float*>(q + offx));
            y = __ldg(reinterpret_cast<const float*>(q + offy));
            z = __ldg(reinterpret_cast<const float*>(q + offz));
        }
        ok = world_from_raw<double>((double)x, (double)y, (double)z, tf, P.min_d2, wx, wy, wz);
    }
    scan_point<FASTFLOOR>(ok, wx, wy, wz, fr, P, O);
}

// ---------------------------------------------------------------------------
// S2  cells + housekeeping stream.
//   warps 0..GW-1 of every block (gather): 8 lanes per cell.  counts: hit = n of the own moments (a sum of 1.0s,
//     exact), passes = -1 - map[v]; the cell id replaces the pass code in the map.  metrics = {mean xyz, cov, n} of
//     all points within (rx, rx, rz) voxels in coordinates relative to the cell's voxel corner: a neighbour's raw
//     moments are about ITS centre, shifting by the integer voxel offset d gives moments about this cell's centre
//     (S' = S + n d, Q'_ab = Q_ab + d_a S_b + S_a d_b + n d_a d_b).  Neighbours are looked up in the extended cell
//     grid, so there are no bounds checks and margin ("ghost") cells are ordinary neighbours.
//   warps GW..7 (stream): per 256-voxel segment, read the slot's map -> group-mask word ("anything known in these
//     8 voxels"), and wipe the segment of the spare slot where ITS mask says it holds anything.
// ---------------------------------------------------------------------------
struct CellArgs {
    int* map;                    // this scan's slot
    unsigned* gmask;             // its group mask (written here) or NULL (xy_size % 8 != 0)
    const unsigned* cellid;
    unsigned tag;
    const int* counters;         // [0] cells, [1] ghosts of this scan
    const int* counters_prev;    // the previous scan's pair: the rows of acc_other it used are cleared here
    int* counters_next;          // the next scan's pair, zeroed here
    double* acc_other;           // the accumulator buffer the NEXT scan will use (this scan's is `acc`)
    float* spare_minh;           // min heights of the spare slot (next scan's target), reset to 1.0 here
    const int* spare_count;      // its cell count
    int gcap;
    const double* acc;
    const int* cell_voxel;
    int* hit; int* total; double* metrics; int* slot_count;
    int* old_map;                // spare slot to wipe (NULL: nothing to do)
    unsigned* old_gmask;         // its group mask (NULL: wipe everything)
    int cap, ES;
};

constexpr int S2_GATHER_WARPS = 6;      // of 8 warps per block

template <int RX, int RZ>               // compile-time neighbourhood radius (RX < 0: runtime P.rx / P.rz)
__global__ void __launch_bounds__(256, 3)
k_scan_cells(CellArgs A, DevParams P) {
    pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int raw_count = A.counters[0];
    const int count = min(raw_count, A.cap);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *A.slot_count = raw_count;
        A.counters_next[0] = 0; A.counters_next[1] = 0;
    }
    if (warp >= S2_GATHER_WARPS) {
        // ---- stream: group mask of this slot, wipe of the spare slot
        const int sw = blockIdx.x * (8 - S2_GATHER_WARPS) + (warp - S2_GATHER_WARPS);
        const int nsw = gridDim.x * (8 - S2_GATHER_WARPS);
        const long long V = P.V;
        {   // housekeeping for the next scan: clear the accumulator rows the previous scan used in the other buffer
            // (cells and margin cells) and the min heights of the spare slot, so that S1 can publish cell ids unfenced
            const long long nc = min(A.counters_prev[0], A.cap), ng = min(A.counters_prev[1], A.gcap);
            double2* a0 = reinterpret_cast<double2*>(A.acc_other);
            double2* a1 = reinterpret_cast<double2*>(A.acc_other + (long long)A.cap * MOM);
            const double2 z2 = make_double2(0.0, 0.0);
            for (long long i = (long long)sw * 32 + lane; i < nc * (MOM / 2); i += (long long)nsw * 32) a0[i] = z2;
            for (long long i = (long long)sw * 32 + lane; i < ng * (MOM / 2); i += (long long)nsw * 32) a1[i] = z2;
            if (A.spare_minh) {
                const int ns = min(*A.spare_count, A.cap);
                for (int i = sw * 32 + lane; i < ns; i += nsw * 32) A.spare_minh[i] = 1.0f;
            }
        }
        if (A.gmask) {
            const int nseg = (int)((V + 255) >> 8);
            constexpr int U = 4;                                   // segments in flight per warp
            for (int s0 = sw * U; s0 < nseg; s0 += nsw * U) {
                unsigned wold[U];
                int4 a[U], b[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int seg = s0 + u;
                    wold[u] = 0u;
                    a[u] = make_int4(-1, -1, -1, -1); b[u] = a[u];
                    if (seg < nseg) {
                        if (A.old_map) wold[u] = A.old_gmask ? A.old_gmask[seg] : 0xffffffffu;
                        const long long q = (long long)seg * 256 + lane * 8;
                        if (q < V) {                               // V % 8 == 0: whole groups
                            const int4* src = reinterpret_cast<const int4*>(A.map + q);
                            a[u] = __ldcg(src); b[u] = __ldcg(src + 1);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int seg = s0 + u;
                    if (seg >= nseg) break;                        // uniform
                    const long long q = (long long)seg * 256 + lane * 8;
                    if (wold[u] != 0u) {                           // uniform
                        if (q < V) {
                            int4* dst = reinterpret_cast<int4*>(A.old_map + q);
                            dst[0] = make_int4(-1, -1, -1, -1); dst[1] = make_int4(-1, -1, -1, -1);
                        }
                        if (lane == 0 && A.old_gmask) A.old_gmask[seg] = 0u;
                    }
                    const bool known = (a[u].x & a[u].y & a[u].z & a[u].w & b[u].x & b[u].y & b[u].z & b[u].w) != -1;
                    const unsigned w = __ballot_sync(FULL, known);
                    if (lane == 0) A.gmask[seg] = w;
                }
            }
        } else if (A.old_map) {
            for (long long q = (long long)sw * 32 + lane; q < V; q += (long long)nsw * 32) A.old_map[q] = -1;
        }
        return;
    }
    // ---- gather
    constexpr int L = 8;                                           // lanes per cell
    const int sub = lane & (L - 1);
    const int gid = (blockIdx.x * S2_GATHER_WARPS + warp) * (32 / L) + (lane >> 3);
    const int ngroups = gridDim.x * S2_GATHER_WARPS * (32 / L);
    const int rx = RX >= 0 ? RX : P.rx, rz = RX >= 0 ? RZ : P.rz;
    const int wx = 2 * rx + 1, wz = 2 * rz + 1;
    const int nn = wx * wx * wz;
    const int S = P.S, ES = A.ES;
    constexpr int CH = 4;                                          // look-ups in flight per lane
    const int count_pad = (count + (32 / L) - 1) / (32 / L) * (32 / L);   // whole warps iterate together
    for (int id = gid; id < count_pad; id += ngroups) {
        double r[MOM];
#pragma unroll
        for (int k = 0; k < MOM; ++k) r[k] = 0.0;
        int v = 0;
        if (id < count) {
            v = A.cell_voxel[id];
            int x, y, z;
            if (P.lgS >= 0) { x = v & (S - 1); y = (v >> P.lgS) & (S - 1); z = v >> (2 * P.lgS); }
            else { x = v % S; y = (v / S) % S; z = v / (S * S); }
            const int e0 = x + (y + z * ES) * ES;                  // extended index of the neighbour (-rx, -rx, -rz)
            for (int j0 = sub; j0 < nn; j0 += L * CH) {
                unsigned ce[CH];
#pragma unroll
                for (int u = 0; u < CH; ++u) {
                    const int j = j0 + u * L;
                    const int dx = j % wx, dy = (j / wx) % wx, dz = j / (wx * wx);
                    ce[u] = j < nn ? __ldg(A.cellid + e0 + dx + (dy + dz * ES) * ES) : 0u;
                }
                int rowi[CH];
#pragma unroll
                for (int u = 0; u < CH; ++u) {
                    const unsigned pay = ce[u] & CELL_PAY;
                    rowi[u] = ((ce[u] >> CELL_TAG_SHIFT) == A.tag && pay < CELL_OVERFLOW)
                                  ? (int)pay + ((ce[u] & CELL_GHOST) ? A.cap : 0) : -1;
                }
                // two-deep pipeline over the occupied neighbours: record u+1 is in flight while record u is folded in
                double2 rb[2][MOM / 2];
#pragma unroll
                for (int g = 0; g < MOM / 2; ++g) { rb[0][g] = make_double2(0.0, 0.0); rb[1][g] = make_double2(0.0, 0.0); }
                if (rowi[0] >= 0) {
                    const double2* a2 = reinterpret_cast<const double2*>(A.acc + (long long)rowi[0] * MOM);
#pragma unroll
                    for (int g = 0; g < MOM / 2; ++g) rb[0][g] = __ldcg(a2 + g);
                }
#pragma unroll
                for (int u = 0; u < CH; ++u) {
                    if (u + 1 < CH && rowi[u + 1 < CH ? u + 1 : u] >= 0) {
                        const double2* a2 = reinterpret_cast<const double2*>(A.acc + (long long)rowi[u + 1 < CH ? u + 1 : u] * MOM);
#pragma unroll
                        for (int g = 0; g < MOM / 2; ++g) rb[(u + 1) & 1][g] = __ldcg(a2 + g);
                    }
                    if (rowi[u] < 0) continue;
                    const double2 a01 = rb[u & 1][0], a23 = rb[u & 1][1], a45 = rb[u & 1][2], a67 = rb[u & 1][3], a89 = rb[u & 1][4];
                    const double a0 = a01.x, a1 = a01.y, a2v = a23.x, an = a89.y;
                    const int j = j0 + u * L;
                    const double ddx = (double)(j % wx - rx), ddy = (double)((j / wx) % wx - rx), ddz = (double)(j / (wx * wx) - rz);
                    r[0] += a0 + an * ddx; r[1] += a1 + an * ddy; r[2] += a2v + an * ddz;
                    r[3] += a23.y + 2.0 * ddx * a0 + an * ddx * ddx;
                    r[4] += a45.x + ddx * a1 + ddy * a0 + an * ddx * ddy;
                    r[5] += a45.y + ddx * a2v + ddz * a0 + an * ddx * ddz;
                    r[6] += a67.x + 2.0 * ddy * a1 + an * ddy * ddy;
                    r[7] += a67.y + ddy * a2v + ddz * a1 + an * ddy * ddz;
                    r[8] += a89.x + 2.0 * ddz * a2v + an * ddz * ddz;
                    r[9] += an;
                }
            }
        }
#pragma unroll
        for (int off = L / 2; off > 0; off >>= 1)
#pragma unroll
            for (int k = 0; k < MOM; ++k) r[k] += __shfl_xor_sync(FULL, r[k], off);
        if (sub == 0 && id < count) {
            const double n = r[9];
            double* mo = A.metrics + (long long)id * 10;
            const double m0 = r[0] / n, m1 = r[1] / n, m2 = r[2] / n;
            mo[0] = m0 + 0.5; mo[1] = m1 + 0.5; mo[2] = m2 + 0.5;
            mo[3] = r[3] / n - m0 * m0; mo[4] = r[4] / n - m0 * m1; mo[5] = r[5] / n - m0 * m2;
            mo[6] = r[6] / n - m1 * m1; mo[7] = r[7] / n - m1 * m2; mo[8] = r[8] / n - m2 * m2;
            mo[9] = n;
            A.hit[id] = (int)__ldcg(A.acc + (long long)id * MOM + 9);      // own points: a sum of 1.0s
            A.total[id] = -1 - A.map[v];                                   // hits + passes (gvom.py:1169,1229)
            A.map[v] = id;
        }
    }
}

// start-up / restore: every slot's min heights read 1.0 (S2 keeps the spare slot's that way afterwards)
__global__ void k_fill_f32(float* __restrict__ p, long long n, float v) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace gvom
