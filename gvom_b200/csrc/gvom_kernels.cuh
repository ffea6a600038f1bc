// gvom_kernels.cuh -- sm_100a device code of the G-VOM voxel-mapping path.
//
// Every kernel cites the reference kernel(s) in scripts/gvom.py whose RESULT it
// reproduces; none of them is a translation -- the decomposition is different
// (one fused per-point kernel, compaction without a host sync, own-voxel moment
// accumulation + neighbourhood gather instead of a 27-way float64 atomic scatter,
// a single pass over the ring buffer instead of one pass per slot, ...).
//
// Arithmetic contract.  Voxel indices, ray trip counts and every classification
// are compared BIT-EXACTLY with the reference, so wherever a float decides an
// integer the operation sequence of the reference's compiled kernel (Numba PTX ->
// ptxas SASS on sm_100, see DESIGN.md "arithmetic spec") is reproduced with explicit
// round-to-nearest intrinsics (__dmul_rn, __fma_rn, __fdiv_rn, ...), which neither
// nvcc nor ptxas may contract or reassociate.  The file is compiled with -fmad=false.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace gvom {

constexpr unsigned FULL = 0xffffffffu;
constexpr int MAX_SLOTS = 64;   // ring-buffer slots a single merge pass can take
constexpr int ACC = 20;         // per-cell accumulators: [0..9] own voxel, [10..19] apron

struct DevParams {
    double xy_res, z_res;
    double min_d2;            // min_distance * min_distance (float64 product, gvom.py:1148)
    double pos_thr, neg_thr, slope_thr, robot_height, r2, ground_to_lidar;
    int S, Z;                 // xy_size, z_size
    int rx, rz;               // xy_eigen_dist, z_eigen_dist
    long long V;              // S*S*Z
};

struct Xform {                // rows 0..2 of the 4x4 sensor->world matrix
    double m[12];
    int enabled;
};

struct Frame {
    double ego[3];            // world position of the sensor (float64, as passed)
    double origin[3];         // grid origin in voxel units, integral (gvom.py:138-141)
    float start[3];           // f32(ego / res): DDA start point (gvom.py:1178-1180)
};

// ---------------------------------------------------------------------------
// point load + transform (gvom.py:1121-1138) + world-frame min-distance test
// (gvom.py:1145-1149).  T = element type of the caller's cloud: the reference
// keeps it on the device, so float32 clouds are squared in float32 and the
// transformed point is rounded back to float32.
// ---------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void load3(const T* __restrict__ pts, int stride, long long i, T& a, T& b, T& c) {
    const T* q = pts + i * stride;
    a = q[0]; b = q[1]; c = q[2];
}
template <>
__device__ __forceinline__ void load3<float>(const float* __restrict__ pts, int stride, long long i, float& a, float& b, float& c) {
    if (stride == 4) {                       // xyz + intensity/pad: one 128-bit load
        const float4 q = __ldg(reinterpret_cast<const float4*>(pts) + i);
        a = q.x; b = q.y; c = q.z;
    } else {
        const float* q = pts + i * stride;
        a = __ldg(q); b = __ldg(q + 1); c = __ldg(q + 2);
    }
}
template <>
__device__ __forceinline__ void load3<double>(const double* __restrict__ pts, int stride, long long i, double& a, double& b, double& c) {
    if (stride == 4) {
        const double2 q0 = __ldg(reinterpret_cast<const double2*>(pts) + 2 * i);
        a = q0.x; b = q0.y; c = __ldg(pts + 4 * i + 2);
    } else {
        const double* q = pts + i * stride;
        a = __ldg(q); b = __ldg(q + 1); c = __ldg(q + 2);
    }
}

template <typename T>
__device__ __forceinline__ bool load_world(const T* __restrict__ pts, int stride, long long i, const Xform& tf,
                                           double min_d2, double& wx, double& wy, double& wz) {
    T p0, p1, p2;
    load3<T>(pts, stride, i, p0, p1, p2);
    if (tf.enabled) {
        const double a0 = (double)p0, a1 = (double)p1, a2 = (double)p2;
        double o[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            double t = __dmul_rn(a1, tf.m[4 * r + 1]);
            t = __fma_rn(a0, tf.m[4 * r + 0], t);
            t = __fma_rn(a2, tf.m[4 * r + 2], t);
            o[r] = __dadd_rn(tf.m[4 * r + 3], t);
        }
        p0 = (T)o[0]; p1 = (T)o[1]; p2 = (T)o[2];     // stored back in the cloud's dtype
    }
    double d2;
    if (sizeof(T) == 4) {
        float t = __fmul_rn((float)p1, (float)p1);
        t = __fmaf_rn((float)p0, (float)p0, t);
        t = __fmaf_rn((float)p2, (float)p2, t);
        d2 = (double)t;
    } else {
        double t = __dmul_rn((double)p1, (double)p1);
        t = __fma_rn((double)p0, (double)p0, t);
        t = __fma_rn((double)p2, (double)p2, t);
        d2 = t;
    }
    wx = (double)p0; wy = (double)p1; wz = (double)p2;
    // NaN / Inf coordinates are dropped (the reference has undefined behaviour there;
    // its ROS caller filters NaNs first, gvom_ros.py:108).
    return (d2 >= min_d2) && (d2 < CUDART_INF);
}

// ---------------------------------------------------------------------------
// K1  voxelise + ray-cast.  Result of __point_2_map (gvom.py:1140-1231) on dense
// hit / pass grids that are zero on entry.
//   * one thread per point; the whole warp walks its 32 rays in lock step
//   * every increment is warp-aggregated: lanes that land in the same voxel are
//     found with __match_any_sync and one lane issues a single RED of the group's
//     size.  Azimuth-adjacent rays of a spinning lidar share most voxels, so this
//     removes the bulk of the same-address traffic at the L2 atomic units.
//   * aggregation changes who issues the atomic, never the per-ray arithmetic.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_voxelize_raycast(const T* __restrict__ pts, int stride, int n, Xform tf, Frame fr, DevParams P,
                   int* __restrict__ hit, int* __restrict__ total) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const double ox = fr.origin[0], oy = fr.origin[1], oz = fr.origin[2];
    const double dS = (double)P.S, dZ = (double)P.Z;

    double wx = 0, wy = 0, wz = 0;
    bool ok = false;
    if (i < n) ok = load_world<T>(pts, stride, i, tf, P.min_d2, wx, wy, wz);

    // ---- hit (gvom.py:1153-1171)
    double ex = 0, ey = 0, ez = 0;
    bool inb = false;
    int v = 0;
    if (ok) {
        ex = __ddiv_rn(wx, P.xy_res); ey = __ddiv_rn(wy, P.xy_res); ez = __ddiv_rn(wz, P.z_res);
        const double xi = floor(__dsub_rn(ex, ox)), yi = floor(__dsub_rn(ey, oy)), zi = floor(__dsub_rn(ez, oz));
        inb = (xi >= 0.0) && (xi < dS) && (yi >= 0.0) && (yi < dS) && (zi >= 0.0) && (zi < dZ);
        if (inb) v = (int)xi + ((int)yi + (int)zi * P.S) * P.S;
    }
    const unsigned mh = __ballot_sync(FULL, inb);
    if (inb) {
        const unsigned peers = __match_any_sync(mh, v);
        if (lane == __ffs(peers) - 1) {
            const int c = __popc(peers);
            atomicAdd(hit + v, c);
            atomicAdd(total + v, c);
        }
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

    // ---- DDA (gvom.py:1208-1231), warp-synchronous.
    // The reference evaluates floor(float64(pt) - origin) per axis.  origin is integral and
    // float64(pt) - origin is exact (24-bit pt, |origin| < 2^31), so floor(pt - origin) ==
    // floorf(pt) - origin exactly: the loop runs on float32/int32 only, plus the float64
    // length accumulation whose sequential rounding decides the trip count.
    const int iox = (int)ox, ioy = (int)oy, ioz = (int)oz;
    while (__any_sync(FULL, active)) {
        bool inside = false;
        int vv = 0;
        if (active) {
            px = __fadd_rn(px, ix); py = __fadd_rn(py, iy); pz = __fadd_rn(pz, iz);
            const float bound = 1.0e9f;
            if (fabsf(px) < bound && fabsf(py) < bound && fabsf(pz) < bound) {
                const int x = (int)floorf(px) - iox, y = (int)floorf(py) - ioy, z = (int)floorf(pz) - ioz;
                inside = ((unsigned)x < (unsigned)P.S) && ((unsigned)y < (unsigned)P.S) && ((unsigned)z < (unsigned)P.Z);
                vv = x + (y + z * P.S) * P.S;
            }
            if (!inside) active = false;                  // left the grid: ray ends
        }
        const unsigned ms = __ballot_sync(FULL, inside);
        if (inside) {
            const unsigned peers = __match_any_sync(ms, vv);
            if (lane == __ffs(peers) - 1) atomicAdd(total + vv, __popc(peers));
            length = __dadd_rn(length, dlen);
            active = length < lim;
        }
    }
}

// ---------------------------------------------------------------------------
// K2  index map + compaction.  Result of __assign_indices + __move_data x2
// (gvom.py:1233-1247) without the host round trip of gvom.py:172: the cell count
// stays on the device and compact arrays are sized for the worst case.  Also
// re-zeroes the dense grids for the next scan (replaces the three fill launches
// of gvom.py:125-131) and initialises the per-cell accumulators
// (gvom.py:1079-1085).  Compact ids are allotted per warp (ballot + one atomic).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_build_index(int* __restrict__ hit, int* __restrict__ total, int* __restrict__ index_map,
              int* __restrict__ counter, int* __restrict__ hit_c, int* __restrict__ total_c,
              int* __restrict__ cell_voxel, double* __restrict__ acc, float* __restrict__ minh,
              long long V, int cap) {
    const int lane = threadIdx.x & 31;
    const long long step = (long long)gridDim.x * blockDim.x;
    const long long Vp = (V + 31) & ~31LL;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < Vp; v += step) {
        int h = 0, t = 0;
        if (v < V) { h = hit[v]; t = total[v]; }
        const bool occ = h > 0;
        const unsigned m = __ballot_sync(FULL, occ);
        int base = 0;
        if (m) {
            if (lane == 0) base = atomicAdd(counter, __popc(m));
            base = __shfl_sync(FULL, base, 0);
        }
        if (v < V) {
            int code = -t - 1;
            if (occ) {
                const int id = base + __popc(m & ((1u << lane) - 1u));
                if (id < cap) {
                    code = id;
                    hit_c[id] = h; total_c[id] = t; cell_voxel[id] = (int)v; minh[id] = 1.0f;
                    double* a = acc + (long long)id * ACC;
#pragma unroll
                    for (int k = 0; k < ACC; ++k) a[k] = 0.0;
                }
            }
            index_map[v] = code;
            if (t != 0) total[v] = 0;
            if (h != 0) hit[v] = 0;
        }
    }
}

// ---------------------------------------------------------------------------
// K3  per-point moment accumulation.  Together with K4 it produces the result of
// __calculate_mean/__normalize_mean/__calculate_covariance/__normalize_covariance
// and __calculate_min_height (gvom.py:1249-1421).
// The reference scatters every point into every occupied voxel of its
// (2rx+1)^2(2rz+1) neighbourhood, twice (16.6 M float64 atomics per OS1-128 scan).
// Here a point only adds its raw first/second moments (about its own voxel's
// centre) to its OWN cell -- 10 atomics -- and K4 gathers the neighbourhood.
// Points whose own voxel lies outside the grid still reach in-grid neighbours in
// the reference (gvom.py:1262-1279); they are rare and are scattered directly into
// the neighbour's "apron" accumulators.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_moments(const T* __restrict__ pts, int stride, int n, Xform tf, Frame fr, DevParams P,
          const int* __restrict__ index_map, double* __restrict__ acc, float* __restrict__ minh) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double wx, wy, wz;
    if (!load_world<T>(pts, stride, i, tf, P.min_d2, wx, wy, wz)) return;
    const double fx = __dsub_rn(__ddiv_rn(wx, P.xy_res), fr.origin[0]);
    const double fy = __dsub_rn(__ddiv_rn(wy, P.xy_res), fr.origin[1]);
    const double fz = __dsub_rn(__ddiv_rn(wz, P.z_res), fr.origin[2]);
    const double bx = floor(fx), by = floor(fy), bz = floor(fz);
    const double dS = (double)P.S, dZ = (double)P.Z;
    const bool inb = (bx >= 0.0) && (bx < dS) && (by >= 0.0) && (by < dS) && (bz >= 0.0) && (bz < dZ);
    if (inb) {
        const int v = (int)bx + ((int)by + (int)bz * P.S) * P.S;
        const int id = index_map[v];
        if (id < 0) return;                               // only on compact-capacity overflow
        const double lz = __dsub_rn(fz, bz);
        const double qx = (fx - bx) - 0.5, qy = (fy - by) - 0.5, qz = lz - 0.5;
        double* a = acc + (long long)id * ACC;
        atomicAdd(a + 0, qx); atomicAdd(a + 1, qy); atomicAdd(a + 2, qz);
        atomicAdd(a + 3, qx * qx); atomicAdd(a + 4, qx * qy); atomicAdd(a + 5, qx * qz);
        atomicAdd(a + 6, qy * qy); atomicAdd(a + 7, qy * qz); atomicAdd(a + 8, qz * qz);
        atomicAdd(a + 9, 1.0);
        // min height: float32 of the in-voxel z fraction, in [0,1] -> ordered as int bits
        atomicMin(reinterpret_cast<int*>(minh) + id, __float_as_int((float)lz));
    } else {
        // apron: walk the neighbourhood like the reference does
        const double rx = (double)P.rx, rz = (double)P.rz;
        if (bx < -rx - 1.0 || bx > dS + rx || by < -rx - 1.0 || by > dS + rx || bz < -rz - 1.0 || bz > dZ + rz) return;
        const int x0 = (int)bx - P.rx, y0 = (int)by - P.rx, z0 = (int)bz - P.rz;
        for (int z = z0; z <= z0 + 2 * P.rz; ++z) {
            if (z < 0 || z >= P.Z) continue;
            for (int y = y0; y <= y0 + 2 * P.rx; ++y) {
                if (y < 0 || y >= P.S) continue;
                for (int x = x0; x <= x0 + 2 * P.rx; ++x) {
                    if (x < 0 || x >= P.S) continue;
                    const int id = index_map[x + (y + z * P.S) * P.S];
                    if (id < 0) continue;
                    const double qx = (fx - (double)x) - 0.5, qy = (fy - (double)y) - 0.5, qz = (fz - (double)z) - 0.5;
                    double* a = acc + (long long)id * ACC + 10;
                    atomicAdd(a + 0, qx); atomicAdd(a + 1, qy); atomicAdd(a + 2, qz);
                    atomicAdd(a + 3, qx * qx); atomicAdd(a + 4, qx * qy); atomicAdd(a + 5, qx * qz);
                    atomicAdd(a + 6, qy * qy); atomicAdd(a + 7, qy * qz); atomicAdd(a + 8, qz * qz);
                    atomicAdd(a + 9, 1.0);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// K4  neighbourhood gather: metrics[id] = {mean xyz, cov xx xy xz yy yz zz, n}
// of all points within (rx, rx, rz) voxels, in coordinates relative to the cell's
// own voxel corner (the reference's local_point, gvom.py:1283-1285).
// A neighbour's raw moments are about ITS centre; shifting by the integer voxel
// offset d gives moments about this cell's centre:
//   S' = S + n d,  Q'_ab = Q_ab + d_a S_b + S_a d_b + n d_a d_b.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_gather_metrics(const int* __restrict__ index_map, const int* __restrict__ cell_voxel,
                 const int* __restrict__ counter, const double* __restrict__ acc,
                 double* __restrict__ metrics, DevParams P, int cap) {
    const int count = min(*counter, cap);
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < count; id += gridDim.x * blockDim.x) {
        const int v = cell_voxel[id];
        const int x = v % P.S, y = (v / P.S) % P.S, z = v / (P.S * P.S);
        double s0 = 0, s1 = 0, s2 = 0, q0 = 0, q1 = 0, q2 = 0, q3 = 0, q4 = 0, q5 = 0, n = 0;
        for (int dz = -P.rz; dz <= P.rz; ++dz) {
            const int zz = z + dz;
            if (zz < 0 || zz >= P.Z) continue;
            for (int dy = -P.rx; dy <= P.rx; ++dy) {
                const int yy = y + dy;
                if (yy < 0 || yy >= P.S) continue;
                for (int dx = -P.rx; dx <= P.rx; ++dx) {
                    const int xx = x + dx;
                    if (xx < 0 || xx >= P.S) continue;
                    const int nid = index_map[xx + (yy + zz * P.S) * P.S];
                    if (nid < 0) continue;
                    const double* a = acc + (long long)nid * ACC;
                    const double an = a[9], a0 = a[0], a1 = a[1], a2 = a[2];
                    const double ddx = (double)dx, ddy = (double)dy, ddz = (double)dz;
                    s0 += a0 + an * ddx; s1 += a1 + an * ddy; s2 += a2 + an * ddz;
                    q0 += a[3] + 2.0 * ddx * a0 + an * ddx * ddx;
                    q1 += a[4] + ddx * a1 + ddy * a0 + an * ddx * ddy;
                    q2 += a[5] + ddx * a2 + ddz * a0 + an * ddx * ddz;
                    q3 += a[6] + 2.0 * ddy * a1 + an * ddy * ddy;
                    q4 += a[7] + ddy * a2 + ddz * a1 + an * ddy * ddz;
                    q5 += a[8] + 2.0 * ddz * a2 + an * ddz * ddz;
                    n += an;
                }
            }
        }
        const double* e = acc + (long long)id * ACC + 10;
        s0 += e[0]; s1 += e[1]; s2 += e[2];
        q0 += e[3]; q1 += e[4]; q2 += e[5]; q3 += e[6]; q4 += e[7]; q5 += e[8];
        n += e[9];
        double* mo = metrics + (long long)id * 10;
        const double m0 = s0 / n, m1 = s1 / n, m2 = s2 / n;
        mo[0] = m0 + 0.5; mo[1] = m1 + 0.5; mo[2] = m2 + 0.5;
        mo[3] = q0 / n - m0 * m0; mo[4] = q1 / n - m0 * m1; mo[5] = q2 / n - m0 * m2;
        mo[6] = q3 / n - m1 * m1; mo[7] = q4 / n - m1 * m2; mo[8] = q5 / n - m2 * m2;
        mo[9] = n;
    }
}

// ===========================================================================
// combine_maps
// ===========================================================================
struct SlotRef {
    const int* map;          // dense index map of the source (codes: >=0 id, -1 unknown, <-1 free)
    const void* metrics;     // [cells,10] float64 (ring slot) or float32 (previous combined map)
    const int* hit;
    const int* total;
    const float* minh;
    int dx, dy, dz;          // combined_origin - source_origin, voxels
    int is_prev;             // 1: previous combined map (float32 metrics, [-11,-1] rule)
};
struct MergeArgs {
    SlotRef s[MAX_SLOTS + 1];
    int n;
};

// ---------------------------------------------------------------------------
// C1  merged code per voxel.  Result of __combine_indices run once per slot in
// slot order followed by __combine_old_indices (gvom.py:1009-1063, 242-257), in
// ONE pass: each thread folds all sources of its voxel in the reference's order.
// Once a voxel is occupied nothing later changes it, so the fold stops there.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_merge_codes(MergeArgs A, int* __restrict__ cmap, int* __restrict__ counter,
              int* __restrict__ cell_voxel, DevParams P, int cap) {
    const int lane = threadIdx.x & 31;
    const long long step = (long long)gridDim.x * blockDim.x;
    const long long Vp = (P.V + 31) & ~31LL;
    const int S = P.S, Z = P.Z;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < Vp; v += step) {
        bool occ = false;
        int c = -1;
        if (v < P.V) {
            const int x = (int)(v % S), y = (int)((v / S) % S), z = (int)(v / ((long long)S * S));
            for (int k = 0; k < A.n; ++k) {
                const SlotRef& s = A.s[k];
                const int xs = x + s.dx, ys = y + s.dy, zs = z + s.dz;
                if (xs < 0 || xs >= S || ys < 0 || ys >= S || zs < 0 || zs >= Z) continue;
                const int o = __ldg(s.map + (xs + (ys + (long long)zs * S) * S));
                if (o >= 0) {
                    if (!s.is_prev || c >= -11) { occ = true; break; }
                } else if (o < -1) {
                    c += o + 1;
                }
            }
        }
        const unsigned m = __ballot_sync(FULL, occ);
        int base = 0;
        if (m) {
            if (lane == 0) base = atomicAdd(counter, __popc(m));
            base = __shfl_sync(FULL, base, 0);
        }
        if (v < P.V) {
            if (occ) {
                const int id = base + __popc(m & ((1u << lane) - 1u));
                if (id < cap) { c = id; cell_voxel[id] = (int)v; }
                else c = -1;
            }
            cmap[v] = c;
        }
    }
}

// closed-form eigenvalues of the float32 covariance (gvom.py:1423-1487)
__device__ __forceinline__ void eigen3(const float* m, float* e) {
    const float xx = m[3], xy = m[4], xz = m[5], yy = m[6], yz = m[7], zz = m[8];
    float p1 = __fmul_rn(xz, xz);
    p1 = __fmaf_rn(xy, xy, p1);
    p1 = __fmaf_rn(yz, yz, p1);
    const double q = (double)__fadd_rn(__fadd_rn(xx, yy), zz) / 3.0;
    if (p1 == 0.0f) {
        e[0] = fmaxf(xx, fmaxf(yy, zz));
        e[2] = fminf(xx, fminf(yy, zz));
        e[1] = (float)((q * 3.0 - (double)e[0]) - (double)e[2]);
    } else {
        const double ax = (double)xx - q, ay = (double)yy - q, az = (double)zz - q;
        double p2 = ay * ay;
        p2 = __fma_rn(ax, ax, p2);
        p2 = __fma_rn(az, az, p2);
        p2 = __fma_rn((double)p1, 2.0, p2);
        const double p = sqrt(p2 / 6.0);
        const double B0 = ax / p, B1 = (double)xy / p, B2 = (double)xz / p, B3 = ay / p, B4 = (double)yz / p, B5 = az / p;
        double r = B0 * (B3 * B5 - B4 * B4) - B1 * (B1 * B5 - B4 * B2);
        r = __fma_rn(B2, B1 * B4 - B3 * B2, r);
        r = r * 0.5;
        double phi;
        if (r <= -1.0) phi = CUDART_PI / 3.0;
        else if (r >= 1.0) phi = 0.0;
        else phi = acos(r) / 3.0;
        e[0] = (float)(q + 2.0 * p * cos(phi));
        e[2] = (float)(q + 2.0 * p * cos(phi + (2.0 * CUDART_PI / 3.0)));
        e[1] = (float)((3.0 * q - (double)e[0]) - (double)e[2]);
    }
}

// one pairwise (Chan) merge step of __combine_metrics (gvom.py:926-980): float64
// math on float32-stored running values, stores round to float32.
__device__ __forceinline__ void merge_step(float* c, const double* o) {
    const double n1 = c[9], n2 = o[9], nt = n1 + n2;
    const double c0 = c[0], c1 = c[1], c2 = c[2];
    const double mx = (c0 * n1 + o[0] * n2) / nt;
    const double my = (c1 * n1 + o[1] * n2) / nt;
    const double mz = (c2 * n1 + o[2] * n2) / nt;
    const double cd[3] = {c0 - mx, c1 - my, c2 - mz};
    const double od[3] = {o[0] - mx, o[1] - my, o[2] - mz};
    const int A[6] = {0, 0, 0, 1, 1, 2}, B[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
    for (int e = 0; e < 6; ++e) {
        const double v = (n1 * (double)c[3 + e] + n2 * o[3 + e] + n1 * cd[A[e]] * cd[B[e]] + n2 * od[A[e]] * od[B[e]]) / nt;
        c[3 + e] = (float)v;
    }
    c[0] = (float)mx; c[1] = (float)my; c[2] = (float)mz;
    c[9] = (float)nt;
}

// ---------------------------------------------------------------------------
// C2  per-cell record merge + eigenvalues.  Result of __combine_metrics run per
// slot and for the previous map (gvom.py:280-298, 888-980) and of
// __calculate_eigenvalues (gvom.py:318), one thread per combined cell, sources
// folded in the reference's order so the float32 rounding sequence is the same.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_merge_cells(MergeArgs A, const int* __restrict__ counter, const int* __restrict__ cell_voxel,
              int* __restrict__ chit, int* __restrict__ ctot, float* __restrict__ cminh,
              float* __restrict__ cmet, float* __restrict__ ceig, DevParams P, int cap) {
    const int count = min(*counter, cap);
    const int S = P.S, Z = P.Z;
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < count; id += gridDim.x * blockDim.x) {
        const int v = cell_voxel[id];
        const int x = v % S, y = (v / S) % S, z = v / (S * S);
        float c[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) c[k] = 0.f;
        int hit = 0, tot = 0;
        float mh = 1.0f;
        for (int k = 0; k < A.n; ++k) {
            const SlotRef& s = A.s[k];
            const int xs = x + s.dx, ys = y + s.dy, zs = z + s.dz;
            if (xs < 0 || xs >= S || ys < 0 || ys >= S || zs < 0 || zs >= Z) continue;
            const int io = __ldg(s.map + (xs + (ys + (long long)zs * S) * S));
            if (io < 0) continue;
            double o[10];
            if (s.is_prev) {
                const float* om = reinterpret_cast<const float*>(s.metrics) + (long long)io * 10;
#pragma unroll
                for (int a = 0; a < 10; ++a) o[a] = (double)om[a];
            } else {
                const double* om = reinterpret_cast<const double*>(s.metrics) + (long long)io * 10;
#pragma unroll
                for (int a = 0; a < 10; ++a) o[a] = om[a];
            }
            merge_step(c, o);
            hit += s.hit[io];
            tot += s.total[io];
            mh = fminf(mh, s.minh[io]);
        }
        float* mo = cmet + (long long)id * 10;
#pragma unroll
        for (int k = 0; k < 10; ++k) mo[k] = c[k];
        chit[id] = hit; ctot[id] = tot; cminh[id] = mh;
        float e[3];
        eigen3(c, e);
        ceig[id * 3 + 0] = e[0]; ceig[id * 3 + 1] = e[1]; ceig[id * 3 + 2] = e[2];
    }
}

// 2-D maps are indexed [x,y] row-major like the reference's numpy outputs
#define GVOM_HM(a, x, y) (a)[(long long)(x) * S + (y)]

// ---------------------------------------------------------------------------
// C3  column reduction 3-D -> height maps.  Result of __make_height_map and
// __make_inferred_height_map (gvom.py:560-590) incl. their -1000 fills.
// Thread <-> x fastest so that a warp reads 32 consecutive voxels of each z layer.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_column_maps(const int* __restrict__ cmap, const float* __restrict__ cminh, double o0, double o1, double o2,
              double e0, double e1, double e2, DevParams P, double* __restrict__ height,
              double* __restrict__ inferred) {
    const int S = P.S;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S * S) return;
    const int x = t % S, y = t / S;
    double h = -1000.0, inf = -1000.0;
    // ego disc: xp = (o+x)*res - ego with the product and difference fused (ptxas contracts
    // the reference's mul+sub), then fma(xp,xp, yp*yp) <= r*r
    const double xp = __fma_rn(__dadd_rn(o0, (double)x), P.xy_res, -e0);
    const double yp = __fma_rn(__dadd_rn(o1, (double)y), P.xy_res, -e1);
    if (__fma_rn(xp, xp, __dmul_rn(yp, yp)) <= P.r2) h = __dsub_rn(e2, P.ground_to_lidar);
    bool got_h = false, got_i = false;
    const int* col = cmap + x + (long long)y * S;
    const long long zs = (long long)S * S;
    for (int z = 0; z < P.Z && !(got_h && got_i); ++z) {
        const int idx = __ldg(col + z * zs);
        if (idx >= 0 && !got_h) {
            h = __dmul_rn(__dadd_rn(__dadd_rn((double)z, (double)cminh[idx]), o2), P.z_res);
            got_h = true;
        } else if (idx < -1 && !got_i) {
            inf = __dmul_rn(__dadd_rn(o2, (double)z), P.z_res);
            got_i = true;
        }
    }
    GVOM_HM(height, x, y) = h;
    GVOM_HM(inferred, x, y) = inf;
}

// ---------------------------------------------------------------------------
// C4  surface maps.  Result of __calculate_slope, __guess_height,
// __make_positive_obstacle_map, __make_negative_obstacle_map and
// __make_visability_map (gvom.py:444-452, 505-555, 592-805) and their fills, one
// thread per map cell (slope of the own cell is all the positive-obstacle test needs).
// Thread <-> y fastest: height-map reads of a warp are contiguous.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_surface_maps(const int* __restrict__ cmap, const int* __restrict__ chit, const int* __restrict__ ctot,
               const double* __restrict__ height, const double* __restrict__ inferred, double o2,
               DevParams P, double* __restrict__ rough, double* __restrict__ xs, double* __restrict__ ys,
               double* __restrict__ guessed, int* __restrict__ pos, int* __restrict__ neg, int* __restrict__ vis) {
    const int S = P.S, Z = P.Z;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S * S) return;
    const int y0 = t % S, x0 = t / S;
    const double h0 = GVOM_HM(height, x0, y0);

    // ---- slope + roughness (gvom.py:717-805); contraction pattern = SASS of the reference
    double sxv = 0.0, syv = 0.0, rg = -1.0;
    {
        double px[9], py[9], pz[9];
        int n = 0;
        double sx = 0, sy = 0, sz = 0;
        const int xa = max(0, x0 - 1), xb = min(S, x0 + 2), ya = max(0, y0 - 1), yb = min(S, y0 + 2);
        for (int x = xa; x < xb; ++x)
            for (int y = ya; y < yb; ++y) {
                const double h = GVOM_HM(height, x, y);
                if (h > -1000.0) {
                    px[n] = __dmul_rn((double)x, P.xy_res); py[n] = __dmul_rn((double)y, P.xy_res); pz[n] = h;
                    sx = __dadd_rn(sx, px[n]); sy = __dadd_rn(sy, py[n]); sz = __dadd_rn(sz, pz[n]);
                    ++n;
                }
            }
        if (n >= 3) {
            const double mx = __ddiv_rn(sx, (double)n), my = __ddiv_rn(sy, (double)n), mz = __ddiv_rn(sz, (double)n);
            double xx = 0, xy = 0, xz = 0, yy = 0, yz = 0;
            for (int i = 0; i < n; ++i) {
                const double dx = __dsub_rn(px[i], mx), dy = __dsub_rn(py[i], my), dz = __dsub_rn(pz[i], mz);
                xx = __fma_rn(dx, dx, xx); xy = __fma_rn(dx, dy, xy); xz = __fma_rn(dx, dz, xz);
                yy = __fma_rn(dy, dy, yy); yz = __fma_rn(dy, dz, yz);
            }
            const double det = __fma_rn(xx, yy, -__dmul_rn(xy, xy));
            if (det != 0.0) {
                double a0 = __ddiv_rn(__fma_rn(xz, yy, -__dmul_rn(xy, yz)), det);
                double a1 = __ddiv_rn(__fma_rn(xx, yz, -__dmul_rn(xy, xz)), det);
                const double m = __dsqrt_rn(__dadd_rn(__fma_rn(a0, a0, __dmul_rn(a1, a1)), 1.0));
                a0 = __ddiv_rn(a0, m); a1 = __ddiv_rn(a1, m);
                double err = 0.0;
                for (int i = 0; i < n; ++i) {
                    const double e = __dsub_rn(__dsub_rn(pz[i], mz),
                                               __fma_rn(a0, __dsub_rn(px[i], mx), __dmul_rn(a1, __dsub_rn(py[i], my))));
                    err = __fma_rn(e, e, err);
                }
                err = __ddiv_rn(err, (double)n);
                if (err > 0.0) err = log(err);
                rg = err;
                const double im = __drcp_rn(m);
                sxv = atan2(a0, im);
                syv = atan2(a1, im);
            }
        }
    }
    GVOM_HM(rough, x0, y0) = rg;
    GVOM_HM(xs, x0, y0) = sxv;
    GVOM_HM(ys, x0, y0) = syv;

    // ---- guessed height delta (gvom.py:592-713), quirks kept (see oracle/gvom_oracle.c)
    double dh_out = 0.0;
    if (!(h0 > -1000.0) && GVOM_HM(inferred, x0, y0) != -1000.0) {
        bool xpd = false, xnd = false, ypd = false, ynd = false;
        int x_p = x0, x_n = x0, y_p = y0, y_n = y0;
        double x_ph = -1000.0, x_nh = -1000.0, y_ph = -1000.0, y_nh = -1000.0;
        int i = 0;
        while (i < 15 && !(xnd && ypd && ynd)) {          // x_p_done is NOT part of the condition (gvom.py:619)
            x_p += 1; x_n -= 1; y_p += 1; y_n -= 1; i += 1;
            if (!xpd) {
                if (x_p < S) {
                    for (int d = -i; d < i; ++d) {
                        const int yy = y0 + d;
                        if (yy >= S || yy < 0) continue;
                        const double h = GVOM_HM(height, x_p, yy);
                        if (h > -1000.0) { x_ph = h; xpd = true; break; }
                    }
                } else xpd = true;
            }
            if (!xnd) {
                if (x_n >= 0) {
                    for (int d = -i + 1; d < i + 1; ++d) {
                        const int yy = y0 + d;
                        if (yy >= S || yy < 0) continue;
                        const double h = GVOM_HM(height, x_n, yy);
                        if (h > -1000.0) { x_nh = h; xnd = true; break; }
                    }
                } else xnd = true;
            }
            if (!ypd) {
                if (y_p < S) {
                    for (int d = -i + 1; d < i + 1; ++d) {
                        const int xx = x0 + d;
                        if (xx >= S || xx < 0) continue;
                        const double h = GVOM_HM(height, xx, y_p);
                        if (h > -1000.0) { y_ph = h; ypd = true; break; }
                    }
                } else ypd = true;
            }
            if (!ynd) {
                if (y_n >= 0) {
                    for (int d = -i; d < i; ++d) {
                        const int xx = x0 + d;
                        if (xx >= S || xx < 0) continue;
                        const double h = GVOM_HM(height, xx, y_n);
                        if (h > -1000.0) { y_nh = h; ynd = true; break; }
                    }
                } else ynd = true;
            }
        }
        double mn = 1000.0, mx = GVOM_HM(inferred, x0, y0);
        if (x_ph > -1000.0) { mn = fmin(x_ph, mn); mx = fmax(x_ph, mx); }
        if (x_nh > -1000.0) { mn = fmin(x_nh, mn); mx = fmax(x_nh, mx); }
        if (y_ph > -1000.0) { mn = fmin(y_ph, mn); mx = fmax(y_ph, mx); }
        if (x_nh > -1000.0) { mn = fmin(y_nh, mn); mx = fmax(y_nh, mx); }   // sic (gvom.py:704-706)
        const double dh = __dsub_rn(mx, mn);
        if (dh > 0.0) dh_out = dh;
    }
    GVOM_HM(guessed, x0, y0) = dh_out;
    GVOM_HM(neg, x0, y0) = dh_out > P.neg_thr ? 100 : 0;
    GVOM_HM(vis, x0, y0) = h0 > -1000.0 ? 1 : 0;

    // ---- positive obstacles (gvom.py:515-555)
    int pv = 0;
    const double sl = __dsqrt_rn(__fma_rn(sxv, sxv, __dmul_rn(syv, syv)));
    if (!(sl < P.slope_thr)) {
        pv = 100;
    } else {
        const double lo = floor(__dsub_rn(__ddiv_rn(__dadd_rn(h0, P.pos_thr), P.z_res), o2));
        const double hi = floor(__dsub_rn(__ddiv_rn(__dadd_rn(h0, P.robot_height), P.z_res), o2));
        // |h0| <= ~1e3 here, so the conversions cannot overflow
        const long long zlo = (long long)lo + 1, zhi = (long long)hi;
        if (zlo >= 0 && zlo < Z && zhi >= 0 && zhi < Z) {
            double density = 0.0, n = 0.0;
            for (long long z = zlo; z <= zhi; ++z) {
                const int idx = __ldg(cmap + (x0 + (y0 + z * S) * S));
                if (idx >= 0) {
                    const int hc = chit[idx];
                    if (hc > 10) { n = __dadd_rn(n, (double)ctot[idx]); density = __dadd_rn(density, (double)hc); }
                }
            }
            if (n > 0.0) density = __ddiv_rn(density, n);
            pv = (int)__dmul_rn(density, 100.0);
        }
    }
    GVOM_HM(pos, x0, y0) = pv;
}

// ---------------------------------------------------------------------------
// debug exports (gvom.py:455-503)
// ---------------------------------------------------------------------------
__global__ void k_debug_voxels(const int* __restrict__ counter, const int* __restrict__ cell_voxel,
                               const int* __restrict__ chit, const int* __restrict__ ctot,
                               const float* __restrict__ eig, double o0, double o1, double o2, DevParams P,
                               int cap, float* __restrict__ out) {
    const int count = min(*counter, cap);
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < count; id += gridDim.x * blockDim.x) {
        const int v = cell_voxel[id];
        const int x = v % P.S, y = (v / P.S) % P.S, z = v / (P.S * P.S);
        float* o = out + (long long)id * 8;
        o[0] = (float)__dmul_rn(__dadd_rn((double)x, o0), P.xy_res);
        o[1] = (float)__dmul_rn(__dadd_rn((double)y, o1), P.xy_res);
        o[2] = (float)__dmul_rn(__dadd_rn((double)z, o2), P.z_res);
        o[3] = (float)__ddiv_rn((double)chit[id], (double)ctot[id]);
        o[4] = (float)chit[id];
        const float e0 = eig[id * 3], e1 = eig[id * 3 + 1], e2 = eig[id * 3 + 2];
        o[5] = __fsub_rn(e0, e1);
        o[6] = __fsub_rn(e1, e2);
        o[7] = e2;
    }
}

__global__ void k_debug_height(const double* __restrict__ height, const double* __restrict__ rough,
                               const double* __restrict__ xs, const double* __restrict__ ys,
                               const double* __restrict__ guessed, double o0, double o1, DevParams P,
                               float* __restrict__ out7, float* __restrict__ out3) {
    const int S = P.S;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S * S) return;
    const int x = t % S, y = t / S;                       // row index = x + y*S
    const float wxf = (float)__dmul_rn(__dadd_rn((double)x, o0), P.xy_res);
    const float wyf = (float)__dmul_rn(__dadd_rn((double)y, o1), P.xy_res);
    if (out7) {
        float* o = out7 + (long long)t * 7;
        const double sx = GVOM_HM(xs, x, y), sy = GVOM_HM(ys, x, y);
        o[0] = wxf; o[1] = wyf;
        o[2] = (float)__dsub_rn(GVOM_HM(height, x, y), P.z_res);
        o[3] = (float)GVOM_HM(rough, x, y);
        o[4] = (float)sx; o[5] = (float)sy;
        o[6] = (float)__dsqrt_rn(__fma_rn(sx, sx, __dmul_rn(sy, sy)));
    }
    if (out3) {
        float* o = out3 + (long long)t * 3;
        o[0] = wxf; o[1] = wyf;
        o[2] = (float)__dsub_rn(GVOM_HM(guessed, x, y), P.z_res);
    }
}


// ===========================================================================
// multi-GPU combine: independent sensor streams per GPU, merged at combine time
// (SURVEY.md 8e).  Each rank folds its OWN ring slots into the common frame:
//   code grid  [V] int32 : OCC_FLAG if any own slot is occupied there, else the
//                          summed pass count (>= 0).  Summed across ranks by one
//                          all-reduce; flag and passes stay separable.
//   records    [n][16] f32: per occupied voxel {voxel id, hit, total, min_h,
//                          10 merged metrics, 2 pad}, all-gathered across ranks.
// Occupancy (OR), pass sums, hit/total sums and min-height are order independent
// in the reference's merge rules (gvom.py:1030-1035), so the result equals a
// single Gvom holding all ranks' slots; moments are reduced commutatively (raw
// sums in float64 atomics) and agree to float32 rounding.
// ===========================================================================
constexpr int OCC_FLAG = 1 << 26;
constexpr int REC = 16;

__global__ void __launch_bounds__(256)
k_partial_codes(MergeArgs A, int* __restrict__ grid, int* __restrict__ counter, float* __restrict__ records,
                DevParams P, int cap) {
    const int lane = threadIdx.x & 31;
    const long long step = (long long)gridDim.x * blockDim.x;
    const long long Vp = (P.V + 31) & ~31LL;
    const int S = P.S, Z = P.Z;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < Vp; v += step) {
        bool occ = false;
        int passes = 0;
        if (v < P.V) {
            const int x = (int)(v % S), y = (int)((v / S) % S), z = (int)(v / ((long long)S * S));
            for (int k = 0; k < A.n; ++k) {
                const SlotRef& s = A.s[k];
                const int xs = x + s.dx, ys = y + s.dy, zs = z + s.dz;
                if (xs < 0 || xs >= S || ys < 0 || ys >= S || zs < 0 || zs >= Z) continue;
                const int o = __ldg(s.map + (xs + (ys + (long long)zs * S) * S));
                if (o >= 0) { occ = true; break; }
                if (o < -1) passes += -o - 1;
            }
        }
        const unsigned m = __ballot_sync(FULL, occ);
        int base = 0;
        if (m) {
            if (lane == 0) base = atomicAdd(counter, __popc(m));
            base = __shfl_sync(FULL, base, 0);
        }
        if (v < P.V) {
            if (occ) {
                const int id = base + __popc(m & ((1u << lane) - 1u));
                if (id < cap) records[(long long)id * REC] = __int_as_float((int)v);
            }
            grid[v] = occ ? OCC_FLAG : min(passes, OCC_FLAG - 1);
        }
    }
}

// per record: fold this rank's slots (reference order, float32 rounding as in C2)
__global__ void __launch_bounds__(128)
k_partial_cells(MergeArgs A, const int* __restrict__ counter, float* __restrict__ records, DevParams P, int cap) {
    const int count = min(*counter, cap);
    const int S = P.S, Z = P.Z;
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < count; id += gridDim.x * blockDim.x) {
        float* r = records + (long long)id * REC;
        const int v = __float_as_int(r[0]);
        const int x = v % S, y = (v / S) % S, z = v / (S * S);
        float c[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) c[k] = 0.f;
        int hit = 0, tot = 0;
        float mh = 1.0f;
        for (int k = 0; k < A.n; ++k) {
            const SlotRef& s = A.s[k];
            const int xs = x + s.dx, ys = y + s.dy, zs = z + s.dz;
            if (xs < 0 || xs >= S || ys < 0 || ys >= S || zs < 0 || zs >= Z) continue;
            const int io = __ldg(s.map + (xs + (ys + (long long)zs * S) * S));
            if (io < 0) continue;
            double o[10];
            const double* om = reinterpret_cast<const double*>(s.metrics) + (long long)io * 10;
#pragma unroll
            for (int a = 0; a < 10; ++a) o[a] = om[a];
            merge_step(c, o);
            hit += s.hit[io]; tot += s.total[io]; mh = fminf(mh, s.minh[io]);
        }
        r[1] = __int_as_float(hit); r[2] = __int_as_float(tot); r[3] = mh;
#pragma unroll
        for (int k = 0; k < 10; ++k) r[4 + k] = c[k];
        r[14] = 0.f; r[15] = 0.f;
    }
}

// final code per voxel from the all-reduced grid + this rank's copy of the previous
// combined map (gvom.py:1037-1063); allots compact ids and clears the cell accumulators.
__global__ void __launch_bounds__(256)
k_finish_codes(const int* __restrict__ grid, SlotRef prev, int has_prev, int* __restrict__ cmap,
               int* __restrict__ counter, int* __restrict__ cell_voxel, double* __restrict__ cacc,
               int* __restrict__ chit, int* __restrict__ ctot, float* __restrict__ cminh, DevParams P, int cap) {
    const int lane = threadIdx.x & 31;
    const long long step = (long long)gridDim.x * blockDim.x;
    const long long Vp = (P.V + 31) & ~31LL;
    const int S = P.S, Z = P.Z;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < Vp; v += step) {
        bool occ = false;
        int c = -1;
        if (v < P.V) {
            const int g = grid[v];
            if (g >= OCC_FLAG) occ = true;
            else {
                c = -1 - g;
                if (has_prev) {
                    const int x = (int)(v % S), y = (int)((v / S) % S), z = (int)(v / ((long long)S * S));
                    const int xs = x + prev.dx, ys = y + prev.dy, zs = z + prev.dz;
                    if (!(xs < 0 || xs >= S || ys < 0 || ys >= S || zs < 0 || zs >= Z)) {
                        const int o = __ldg(prev.map + (xs + (ys + (long long)zs * S) * S));
                        if (o >= 0) { if (c >= -11) occ = true; }
                        else if (o < -1) c += o + 1;
                    }
                }
            }
        }
        const unsigned m = __ballot_sync(FULL, occ);
        int base = 0;
        if (m) {
            if (lane == 0) base = atomicAdd(counter, __popc(m));
            base = __shfl_sync(FULL, base, 0);
        }
        if (v < P.V) {
            if (occ) {
                const int id = base + __popc(m & ((1u << lane) - 1u));
                if (id < cap) {
                    c = id; cell_voxel[id] = (int)v; chit[id] = 0; ctot[id] = 0; cminh[id] = 1.0f;
                    double* a = cacc + (long long)id * 10;
#pragma unroll
                    for (int k = 0; k < 10; ++k) a[k] = 0.0;
                } else c = -1;
            }
            cmap[v] = c;
        }
    }
}

// scatter every rank's records into the cells: raw moments n, n*mu, n*(C + mu mu^T) in float64
__global__ void __launch_bounds__(256)
k_scatter_records(const float* __restrict__ records, const int* __restrict__ counts, int nranks, long long capacity,
                  const int* __restrict__ cmap, double* __restrict__ cacc, int* __restrict__ chit,
                  int* __restrict__ ctot, float* __restrict__ cminh) {
    for (int rk = 0; rk < nranks; ++rk) {
        const int count = (int)min((long long)counts[rk], capacity);
        const float* base = records + (long long)rk * capacity * REC;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
            const float* r = base + (long long)i * REC;
            const int id = cmap[__float_as_int(r[0])];
            if (id < 0) continue;
            atomicAdd(chit + id, __float_as_int(r[1]));
            atomicAdd(ctot + id, __float_as_int(r[2]));
            atomicMin(reinterpret_cast<int*>(cminh) + id, __float_as_int(r[3]));   // values in [0,1]
            const double n = r[13], m0 = r[4], m1 = r[5], m2 = r[6];
            double* a = cacc + (long long)id * 10;
            atomicAdd(a + 0, n * m0); atomicAdd(a + 1, n * m1); atomicAdd(a + 2, n * m2);
            atomicAdd(a + 3, n * ((double)r[7] + m0 * m0)); atomicAdd(a + 4, n * ((double)r[8] + m0 * m1));
            atomicAdd(a + 5, n * ((double)r[9] + m0 * m2)); atomicAdd(a + 6, n * ((double)r[10] + m1 * m1));
            atomicAdd(a + 7, n * ((double)r[11] + m1 * m2)); atomicAdd(a + 8, n * ((double)r[12] + m2 * m2));
            atomicAdd(a + 9, n);
        }
    }
}

// per cell: raw sums -> float32 record, then the previous map (reference order: last), eigenvalues
__global__ void __launch_bounds__(128)
k_finish_cells(SlotRef prev, int has_prev, const int* __restrict__ counter, const int* __restrict__ cell_voxel,
               const double* __restrict__ cacc, int* __restrict__ chit, int* __restrict__ ctot,
               float* __restrict__ cminh, float* __restrict__ cmet, float* __restrict__ ceig, DevParams P, int cap) {
    const int count = min(*counter, cap);
    const int S = P.S, Z = P.Z;
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < count; id += gridDim.x * blockDim.x) {
        const double* a = cacc + (long long)id * 10;
        float c[10];
        const double n = a[9];
        if (n > 0.0) {
            const double m0 = a[0] / n, m1 = a[1] / n, m2 = a[2] / n;
            c[0] = (float)m0; c[1] = (float)m1; c[2] = (float)m2;
            c[3] = (float)(a[3] / n - m0 * m0); c[4] = (float)(a[4] / n - m0 * m1); c[5] = (float)(a[5] / n - m0 * m2);
            c[6] = (float)(a[6] / n - m1 * m1); c[7] = (float)(a[7] / n - m1 * m2); c[8] = (float)(a[8] / n - m2 * m2);
            c[9] = (float)n;
        } else {
#pragma unroll
            for (int k = 0; k < 10; ++k) c[k] = 0.f;
        }
        int hit = chit[id], tot = ctot[id];
        float mh = cminh[id];
        if (has_prev) {
            const int v = cell_voxel[id];
            const int x = v % S, y = (v / S) % S, z = v / (S * S);
            const int xs = x + prev.dx, ys = y + prev.dy, zs = z + prev.dz;
            if (!(xs < 0 || xs >= S || ys < 0 || ys >= S || zs < 0 || zs >= Z)) {
                const int io = __ldg(prev.map + (xs + (ys + (long long)zs * S) * S));
                if (io >= 0) {
                    double o[10];
                    const float* om = reinterpret_cast<const float*>(prev.metrics) + (long long)io * 10;
#pragma unroll
                    for (int k = 0; k < 10; ++k) o[k] = (double)om[k];
                    merge_step(c, o);
                    hit += prev.hit[io]; tot += prev.total[io]; mh = fminf(mh, prev.minh[io]);
                }
            }
        }
        float* mo = cmet + (long long)id * 10;
#pragma unroll
        for (int k = 0; k < 10; ++k) mo[k] = c[k];
        chit[id] = hit; ctot[id] = tot; cminh[id] = mh;
        float e[3];
        eigen3(c, e);
        ceig[id * 3 + 0] = e[0]; ceig[id * 3 + 1] = e[1]; ceig[id * 3 + 2] = e[2];
    }
}


// ---------------------------------------------------------------------------
// tooling: L2 atomic-throughput microbenchmark (roofline denominator of the
// ray-cast kernel).  Every thread issues `per_thread` RED.ADD.U32 to
// pseudo-random words of an L2-resident table (mode 0) or to ONE word (mode 1).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_atomic_bench(int* __restrict__ table, unsigned mask, int per_thread, int mode) {
    unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    for (int k = 0; k < per_thread; ++k) {
        s = s * 1664525u + 1013904223u;
        const unsigned a = mode ? 0u : ((s >> 8) & mask);
        atomicAdd(table + a, 1);
    }
}

}  // namespace gvom
