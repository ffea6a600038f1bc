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

// Programmatic dependent launch: every pipeline kernel is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so its blocks are scheduled while the previous
// kernel of the stream drains; nothing may touch that kernel's output before this call returns.
__device__ __forceinline__ void pdl_wait() {
#if __CUDA_ARCH__ >= 900
    cudaGridDependencySynchronize();
#endif
}
// device-side barrier of the peer-to-peer exchanges, folded into the consumer kernel: every block waits until all
// ranks' flag slots (written by their signal kernels after a system-scope fence) have reached `epoch`
// A rank that died or never calls the collective must not hang the others' GPUs: every spin gives up after ~10 s
// (the combine then produces garbage, which the callers' error paths / parity checks report -- but the kernel ends).
constexpr unsigned long long SPIN_LIMIT_NS = 10000000000ULL;
__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ bool spin_until(const volatile int* f, int epoch) {
    if (*f >= epoch) return true;
    const unsigned long long t0 = now_ns();
    while (*f < epoch) {
        __nanosleep(100);
        if (now_ns() - t0 > SPIN_LIMIT_NS) return false;
    }
    return true;
}
__device__ __forceinline__ void wait_flags_block(const int* flags, int n, int epoch) {
    if (flags) {
        if (threadIdx.x == 0) {
            for (int k = 0; k < n; ++k) {
                const volatile int* f = flags + k;
                if (!spin_until(f, epoch)) break;
            }
            __threadfence_system();
        }
        __syncthreads();
    }
}


constexpr int MAX_RANKS = 16;
struct SignalSet { int* slot[MAX_RANKS]; int n; };   // one flag slot per rank (this rank's slot in every rank's memory)

// "The whole grid is done" signal folded into the producing kernel (saves a launch per barrier): every block makes its
// writes visible system-wide and counts itself in; the last one publishes the epoch into every rank's slot
// and resets the counter for the next launch.  Every thread of the block must call it.
struct GridSignal { SignalSet S; int* counter; int epoch; };
__device__ __forceinline__ void signal_when_grid_done(const GridSignal& G) {
    if (G.S.n == 0) return;
    __syncthreads();                                      // the block's writes happen before thread 0's fence (cumulativity)
    if (threadIdx.x == 0) {
        __threadfence_system();
        const int nblocks = gridDim.x * gridDim.y * gridDim.z;
        if (atomicAdd(G.counter, 1) == nblocks - 1) {
            *G.counter = 0;
            __threadfence();
            for (int k = 0; k < G.S.n; ++k) { volatile int* f = G.S.slot[k]; f[0] = G.epoch; }
            __threadfence_system();
        }
    }
}

// The cheaper form of the same signal, used where the NEXT kernel of the stream starts with a wait anyway: that kernel's
// first block publishes the flag before it waits.  Everything the previous kernels of the stream wrote -- remote stores
// included -- is complete and visible when a kernel passes pdl_wait(), so no per-block system fence (ERRBAR: 21 % of the
// stall samples of the cell kernel at 8 ranks) and no counter are needed in the producer.  Call before the wait.
__device__ __forceinline__ void publish_flag_first_block(const SignalSet& S, int epoch) {
    if (S.n == 0 || blockIdx.x != 0 || blockIdx.y != 0 || blockIdx.z != 0) return;
    if (threadIdx.x < S.n) {
        __threadfence_system();
        *reinterpret_cast<volatile int*>(S.slot[threadIdx.x]) = epoch;
    }
}

constexpr int MAX_SLOTS = 64;   // ring-buffer slots a single merge pass can take
constexpr int ACC = 20;         // per-cell accumulators: [0..9] own voxel, [10..19] apron

struct DevParams {
    double xy_res, z_res;
    double min_d2;            // min_distance * min_distance (float64 product, gvom.py:1148)
    double pos_thr, neg_thr, slope_thr, robot_height, r2, ground_to_lidar;
    int S, Z;                 // xy_size, z_size
    int rx, rz;               // xy_eigen_dist, z_eigen_dist
    int lgS;                  // log2(S) if S is a power of two, else -1
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
    int io[3];                // origin as int32
    int cc;                   // (M + io[0]) + (M + io[1]) S + (M + io[2]) S^2 mod 2^32, M = bits of 12582912.0f (bulk DDA phase)
    int S2;                   // S * S
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
__device__ __forceinline__ bool world_from_raw(T p0, T p1, T p2, const Xform& tf, double min_d2,
                                               double& wx, double& wy, double& wz);

template <typename T>
__device__ __forceinline__ bool load_world(const T* __restrict__ pts, int stride, long long i, const Xform& tf,
                                           double min_d2, double& wx, double& wy, double& wz) {
    T p0, p1, p2;
    load3<T>(pts, stride, i, p0, p1, p2);
    return world_from_raw<T>(p0, p1, p2, tf, min_d2, wx, wy, wz);
}

template <typename T>
__device__ __forceinline__ bool world_from_raw(T p0, T p1, T p2, const Xform& tf, double min_d2,
                                               double& wx, double& wy, double& wz) {
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
    const unsigned* gmask;   // one bit per 8-voxel group: something known in the group (NULL: no mask)
};
struct MergeArgs {
    SlotRef s[MAX_SLOTS + 1];
    int n;
    int use_masks;           // every source carries a group mask
};

// ---------------------------------------------------------------------------
// C1  merged code per voxel.  Result of __combine_indices run once per slot in
// slot order followed by __combine_old_indices (gvom.py:1009-1063, 242-257), in
// ONE pass over the combined grid: each thread folds all sources of its voxels in
// the reference's order (once a voxel is occupied nothing later changes it).
//   * VEC = 4: four x-consecutive voxels per thread (S % 4 == 0), sources read with
//     128-bit loads when the slot's x-shift keeps them aligned, SLOT_BATCH sources
//     in flight before the fold -- the kernel is a pure streaming pass.
//   * the column reduction of __make_height_map / __make_inferred_height_map
//     (gvom.py:560-590: lowest occupied / lowest free voxel of every column) is fused
//     in: an atomicMin per candidate on two S*S int maps, pre-filtered by a plain load.
// ---------------------------------------------------------------------------
constexpr int SLOT_BATCH = 3;          // sources in flight per thread (B = 4: 4 slots + new map + previous = 2 batches)
constexpr int OCC_FLAG = 1 << 26;     // multi-GPU partial grids: "occupied" flag; below it: summed pass count
constexpr int REC = 16;               // floats per multi-GPU cell record

// What the merge pass does with the folded codes:
//   MERGE_FULL    single-GPU combine: ring slots + previous map -> combined index map
//   MERGE_PARTIAL multi-GPU, step 1: this rank's ring slots -> encoded grid (OCC_FLAG | passes) + record ids
//   MERGE_FINISH  multi-GPU, step 2: every rank's encoded grid (own HBM or a peer's over NVLink, or one
//                 all-reduced grid) + previous map -> combined index map, cell accumulators cleared
//   MERGE_ROWS    multi-GPU, mirrored ring slots: like MERGE_FULL, restricted to the world rows this rank owns
//                 (O.row_y0 / O.row_n); the sources are local mirrors of every rank's ring slots + the previous map
enum { MERGE_FULL = 0, MERGE_PARTIAL = 1, MERGE_FINISH = 2, MERGE_ROWS = 3 };

struct MergeOut {
    int* cmap;               // combined index map (FULL / FINISH) or encoded grid (PARTIAL)
    int* counter;            // running cell / record counter
    int* cell_voxel;         // FULL / FINISH: compact id -> voxel
    float* records;          // PARTIAL: record rows, voxel id in column 0
    int* col_occ;            // FULL / FINISH: per-column lowest occupied / free z
    int* col_free;
    unsigned* gmask;         // group mask of the output (VEC == 8) or NULL
    double* cacc;            // FINISH: per-cell raw-moment accumulators (cleared here)
    int* chit; int* ctot; float* cminh;
    int cap;
    const int* wait_flags;   // FINISH, peer-to-peer: flags[k] >= wait_epoch once rank k's partial results are visible
    int wait_n, wait_epoch;
    int row_y0, row_n;       // ROWS: this rank merges the rows y = row_y0, row_y0 + row_n, ... of every plane (row_n <= 1: all)
    unsigned* srcmask;       // ROWS: per combined cell, bit k set <=> source k is occupied there (NULL / more than 32 sources: off)
};

__device__ __forceinline__ void column_min(int* __restrict__ col, int z) {
    if (z < *col) atomicMin(col, z);
}

// VEC x-consecutive codes of one source row, starting at xs (may be out of range / unaligned);
// anything outside the source grid reads as -1 (unknown), which folds to nothing.
// does the source hold anything but "unknown" in the <= 2 eight-voxel groups that the run
// [lin, lin+VEC) overlaps?  (one bit per group; the run is clipped to the source row first)
__device__ __forceinline__ bool group_known(const unsigned* __restrict__ gmask, long long row0, int xs, int S, int VEC_) {
    const int a = max(xs, 0), b = min(xs + VEC_ - 1, S - 1);
    if (a > b) return false;
    const long long g0 = (row0 + a) >> 3, g1 = (row0 + b) >> 3;
    unsigned hitb = (__ldg(gmask + (g0 >> 5)) >> (g0 & 31)) & 1u;
    if (g1 != g0) hitb |= (__ldg(gmask + (g1 >> 5)) >> (g1 & 31)) & 1u;
    return hitb != 0;
}

template <int VEC>
__device__ __forceinline__ void load_codes(const int* __restrict__ row, int xs, int S, int (&o)[VEC]) {
    if (VEC >= 4 && xs >= 0 && xs + VEC <= S && (xs & 3) == 0) {
#pragma unroll
        for (int g = 0; g < VEC / 4; ++g) {
            const int4 t = __ldg(reinterpret_cast<const int4*>(row + xs) + g);
            o[4 * g] = t.x; o[4 * g + (VEC > 1 ? 1 : 0)] = t.y; o[4 * g + (VEC > 2 ? 2 : 0)] = t.z; o[4 * g + (VEC > 3 ? 3 : 0)] = t.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) o[j] = (xs + j >= 0 && xs + j < S) ? __ldg(row + xs + j) : -1;
    }
}

// The fold over the ring slots is order independent: a voxel is occupied iff any slot is
// (code >= 0 <=> sign bit clear, so AND the codes), otherwise its code is -1 - sum(passes) with
// passes = -code-1 = ~code for free voxels and 0 for unknown (-1): sum += max(~code, 0).  Only
// the previous combined map has an order dependent rule (the [-11,-1] window) and it comes last.
template <int VEC, int MODE>
__global__ void __launch_bounds__(256, 3)
k_merge_codes(MergeArgs A, MergeOut O, DevParams P) {
    pdl_wait();
    if (MODE == MERGE_FINISH && O.wait_flags) {
        // device-side barrier of the peer-to-peer exchange, folded into the consumer: every block waits
        // until all ranks have published this combine's partial results (their signal kernel wrote the epoch
        // into OUR flag slots after a system-scope fence), then reads the peers' grids over NVLink.
        if (threadIdx.x == 0) {
            for (int k = 0; k < O.wait_n; ++k)
                if (!spin_until(O.wait_flags + k, O.wait_epoch)) break;
            __threadfence_system();
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const long long step = (long long)gridDim.x * blockDim.x;
    const int S = P.S, Z = P.Z;
    const long long NQ = P.V / VEC;
    const long long NQp = (NQ + 31) & ~31LL;
    const int has_prev = (A.n > 0 && A.s[A.n - 1].is_prev) ? 1 : 0;
    for (long long ql = (long long)blockIdx.x * blockDim.x + threadIdx.x; ql < NQp; ql += step) {
        const long long q = ql;
        const bool live = ql < NQ;
        int acc_and[VEC], sum[VEC], enc_or[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) { acc_and[j] = -1; sum[j] = 0; enc_or[j] = 0; }
        int x = 0, y = 0, z = 0;
        int op[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) op[j] = -1;
        if (live) {
            const unsigned v0 = (unsigned)(q * VEC);          // V < 2^31
            if (P.lgS >= 0) {
                x = (int)(v0 & (unsigned)(S - 1));
                const unsigned yz = v0 >> P.lgS;
                y = (int)(yz & (unsigned)(S - 1)); z = (int)(yz >> P.lgS);
            } else {
                x = (int)(v0 % (unsigned)S);
                const unsigned yz = v0 / (unsigned)S;
                y = (int)(yz % (unsigned)S); z = (int)(yz / (unsigned)S);
            }
            // Sources in batches of SLOT_BATCH, the previous combined map riding in the last batch.
            // Phase A: which sources hold anything but "unknown" here (group-mask words, independent
            // loads).  Phase B: the code loads of those sources, all in flight together.  Phase C: fold.
            for (int k0 = 0; k0 < A.n; k0 += SLOT_BATCH) {
                int row0[SLOT_BATCH];                              // V < 2^31: 32-bit linear indices
                int xs[SLOT_BATCH];
                unsigned need = 0;
                unsigned w0[SLOT_BATCH], w1[SLOT_BATCH];
                int b0[SLOT_BATCH], b1[SLOT_BATCH];
#pragma unroll
                for (int u = 0; u < SLOT_BATCH; ++u) {            // branch-free: clamp, load, mask afterwards
                    const int k = min(k0 + u, A.n - 1);
                    const SlotRef& s = A.s[k];
                    const int ys = y + s.dy, zs = z + s.dz;
                    xs[u] = x + s.dx;
                    const int a = max(xs[u], 0), b = min(xs[u] + VEC - 1, S - 1);
                    const bool in = (k0 + u < A.n) && ((unsigned)ys < (unsigned)S) && ((unsigned)zs < (unsigned)Z) && a <= b;
                    row0[u] = in ? (zs * S + ys) * S : 0;
                    if (in) need |= 1u << u;
                    const int g0 = in ? (row0[u] + a) >> 3 : 0, g1 = in ? (row0[u] + b) >> 3 : 0;
                    b0[u] = g0 & 31; b1[u] = g1 & 31;
                    w0[u] = 0xffffffffu; w1[u] = 0xffffffffu;
                    if (A.use_masks) {
                        w0[u] = __ldg(s.gmask + (g0 >> 5));
                        w1[u] = ((g1 >> 5) == (g0 >> 5)) ? w0[u] : __ldg(s.gmask + (g1 >> 5));
                    }
                }
#pragma unroll
                for (int u = 0; u < SLOT_BATCH; ++u)
                    if ((((w0[u] >> b0[u]) | (w1[u] >> b1[u])) & 1u) == 0) need &= ~(1u << u);
                int o[SLOT_BATCH][VEC];
#pragma unroll
                for (int u = 0; u < SLOT_BATCH; ++u) {
                    const bool enc = MODE == MERGE_FINISH && !(has_prev && k0 + u == A.n - 1);
#pragma unroll
                    for (int j = 0; j < VEC; ++j) o[u][j] = enc ? 0 : -1;   // "unknown": folds to nothing
                    if (need & (1u << u)) {
                        if (enc) {                                          // encoded grids: unknown reads as 0
                            int t[VEC];
                            load_codes<VEC>(A.s[k0 + u].map + row0[u], xs[u], S, t);
#pragma unroll
                            for (int j = 0; j < VEC; ++j) o[u][j] = (xs[u] + j >= 0 && xs[u] + j < S) ? t[j] : 0;
                        } else {
                            load_codes<VEC>(A.s[k0 + u].map + row0[u], xs[u], S, o[u]);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < SLOT_BATCH; ++u) {
                    if (!(need & (1u << u))) continue;                      // nothing but "unknown" there: no fold either
                    if (has_prev && k0 + u == A.n - 1) {
#pragma unroll
                        for (int j = 0; j < VEC; ++j) op[j] = o[u][j];
                    } else if (MODE == MERGE_FINISH) {
#pragma unroll
                        for (int j = 0; j < VEC; ++j) { enc_or[j] |= o[u][j]; sum[j] += o[u][j] < OCC_FLAG ? o[u][j] : 0; }
                    } else {
#pragma unroll
                        for (int j = 0; j < VEC; ++j) { acc_and[j] &= o[u][j]; sum[j] += max(~o[u][j], 0); }
                    }
                }
            }
        }
        bool occ[VEC];
        int c[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            occ[j] = (acc_and[j] >= 0) || (enc_or[j] >= OCC_FLAG);
            c[j] = -1 - sum[j];
            if (!occ[j]) {                                    // previous combined map (gvom.py:1058-1063)
                if (op[j] >= 0) { if (c[j] >= -11) occ[j] = true; }
                else if (op[j] < -1) c[j] += op[j] + 1;
            }
        }
        unsigned m[VEC];
        int nocc = 0;
        bool any_free = false, my_occ = false;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            m[j] = __ballot_sync(FULL, occ[j]); nocc += __popc(m[j]);
            any_free |= (!occ[j] && c[j] < -1); my_occ |= occ[j];
        }
        int base = 0;
        if (nocc) {
            if (lane == 0) base = atomicAdd(O.counter, nocc);
            base = __shfl_sync(FULL, base, 0);
        }
        bool known_any = false;
        if (live) {
            if (MODE == MERGE_PARTIAL) {
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    int rid = OCC_FLAG - 1;                    // "occupied, record dropped" (capacity overflow)
                    if (occ[j]) {
                        const int id = base + __popc(m[j] & lt);
                        if (id < O.cap) { O.records[(long long)id * REC] = __int_as_float((int)(q * VEC + j)); rid = id; }
                    }
                    base += __popc(m[j]);
                    // occupied: flag | record id (lets a finishing rank fetch the record without a search)
                    c[j] = occ[j] ? (OCC_FLAG | rid) : min(sum[j], OCC_FLAG - 1);
                    known_any |= c[j] != 0;
                }
            } else if (!my_occ && !any_free) {                    // nothing known here (the common case)
#pragma unroll
                for (int j = 0; j < VEC; ++j) c[j] = -1;
            } else {
                const bool cols = O.col_occ != nullptr;            // (the sharded finish leaves them to the gather pass)
                int* colo = O.col_occ + y * S + x;
                int* colf = O.col_free + y * S + x;
                // current column minima, pre-filter of the atomicMin (vector loads, issued together)
                int cur_occ[VEC], cur_free[VEC];
#pragma unroll
                for (int j = 0; j < VEC; ++j) { cur_occ[j] = cols ? 0x7fffffff : -1; cur_free[j] = cols ? 0x7fffffff : -1; }
                if (!cols) {
                } else if (VEC >= 4) {
#pragma unroll
                    for (int g = 0; g < VEC / 4; ++g) {
                        if (my_occ) {
                            const int4 a = *(reinterpret_cast<const int4*>(colo) + g);
                            cur_occ[4 * g] = a.x; cur_occ[4 * g + (VEC > 1 ? 1 : 0)] = a.y; cur_occ[4 * g + (VEC > 2 ? 2 : 0)] = a.z; cur_occ[4 * g + (VEC > 3 ? 3 : 0)] = a.w;
                        }
                        if (any_free) {
                            const int4 b = *(reinterpret_cast<const int4*>(colf) + g);
                            cur_free[4 * g] = b.x; cur_free[4 * g + (VEC > 1 ? 1 : 0)] = b.y; cur_free[4 * g + (VEC > 2 ? 2 : 0)] = b.z; cur_free[4 * g + (VEC > 3 ? 3 : 0)] = b.w;
                        }
                    }
                } else {
                    cur_occ[0] = colo[0]; cur_free[0] = colf[0];
                }
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    if (occ[j]) {
                        const int id = base + __popc(m[j] & lt);
                        if (id < O.cap) {
                            c[j] = id; O.cell_voxel[id] = (int)(q * VEC + j);
                            if (z < cur_occ[j]) atomicMin(colo + j, z);
                            if (MODE == MERGE_FINISH && O.cacc) {
                                O.chit[id] = 0; O.ctot[id] = 0; O.cminh[id] = 1.0f;
                                double2* a = reinterpret_cast<double2*>(O.cacc + (long long)id * 10);
#pragma unroll
                                for (int k = 0; k < 5; ++k) a[k] = make_double2(0.0, 0.0);
                            }
                        } else c[j] = -1;
                    } else if (c[j] < -1) {
                        if (z < cur_free[j]) atomicMin(colf + j, z);
                    }
                    base += __popc(m[j]);
                    known_any |= c[j] != -1;
                }
            }
            if (VEC >= 4) {
#pragma unroll
                for (int g = 0; g < VEC / 4; ++g)
                    reinterpret_cast<int4*>(O.cmap)[q * (VEC / 4) + g] =
                        make_int4(c[4 * g], c[4 * g + (VEC > 1 ? 1 : 0)], c[4 * g + (VEC > 2 ? 2 : 0)], c[4 * g + (VEC > 3 ? 3 : 0)]);
            } else {
                O.cmap[q] = c[0];
            }
        }
        if (VEC == 8 && O.gmask) {                            // group mask of the output grid
            const unsigned w = __ballot_sync(FULL, known_any);
            if (lane == 0 && live) O.gmask[q >> 5] = w;
        }
    }
}

// ---------------------------------------------------------------------------
// C1 (second build, single-GPU combine, xy_size % 256 == 0)  same results as k_merge_codes<8, MERGE_FULL>,
// organised by ROW SEGMENTS instead of independent 8-voxel items: a warp owns 256 x-consecutive voxels of one
// (y, z) row, lane l the eight voxels [8l, 8l+8).
//   * everything that depends on (y, z) -- source row, range checks -- is warp-uniform
//   * the group-mask bits of a source for the whole segment are two words: lanes 0-15 / 16-31 fetch word 0 / 1
//     of sources 0-15 with ONE load instruction and the warp shares them by shuffle, so the mask phase costs one
//     memory round trip per segment instead of one per source batch, and sources that hold nothing in the segment
//     are skipped by a uniform branch
//   * NB sources' code loads are in flight before the fold.  (Measured and dropped: three aligned 128-bit loads +
//     a uniform register shift for x-shifted sources -- 36 us against 32 us with per-element loads, the 32-byte lane
//     stride wastes half of every 128-bit request; prefetching the next segment's mask words -- no change.)
//   * segments in which no source knows anything (most of the grid) are not even written when the destination
//     buffer's own group mask says it already holds "unknown" there (both combined-map buffers start as all
//     unknown with an empty mask, and every writer keeps map and mask consistent).
// ---------------------------------------------------------------------------
// MODE as in k_merge_codes: MERGE_FULL (single-GPU combine), MERGE_PARTIAL (a rank's own slots -> encoded grid + record
// ids; "unknown" is 0 there), MERGE_FINISH (every rank's encoded grid + previous map -> combined map; waits for the peers'
// partial results itself) and MERGE_ROWS (mirrored multi-GPU combine: MERGE_FULL on the rows this rank owns).
template <int NB, int MODE, bool MASKS = false>      // MASKS (ROWS only): also record which sources are occupied at every combined cell
__device__ __forceinline__ void merge_rows_body(const MergeArgs& A, const MergeOut& O, const DevParams& P) {
    if (MODE == MERGE_FINISH && O.wait_flags) {           // device-side barrier of the peer-to-peer exchange (see k_merge_codes)
        if (threadIdx.x == 0) {
            for (int k = 0; k < O.wait_n; ++k)
                if (!spin_until(O.wait_flags + k, O.wait_epoch)) break;
            __threadfence_system();
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int S = P.S, Z = P.Z;
    const int spr = S >> 8;                               // segments per row
    const bool rslab = MODE == MERGE_ROWS && O.row_n > 1;
    const int my_rows = rslab ? (O.row_y0 < S ? (S - O.row_y0 + O.row_n - 1) / O.row_n : 0) : S;
    const int nseg = rslab ? Z * my_rows * spr : (int)(P.V >> 8);
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int has_prev = (MODE != MERGE_PARTIAL && A.n > 0 && A.s[A.n - 1].is_prev) ? 1 : 0;
    constexpr int UNK = MODE == MERGE_PARTIAL ? 0 : -1;   // what "unknown" looks like in the destination
    // group-mask words of sources kb..kb+15 for segment sg: lanes 0-15 word 0, lanes 16-31 word 1
    auto mask_words = [&](int sg, int kb) -> unsigned {
        const int row = sg / spr;
        const int x0s = (sg - row * spr) << 8;
        const int z = row / S, y = row - z * S;
        const int k = kb + (lane & 15), which = lane >> 4;
        unsigned word = 0;
        if (k < A.n) {
            const SlotRef& s = A.s[k];
            const int dx = s.dx, dy = s.dy, dz = s.dz;
            const int ys = y + dy, zs = z + dz;
            const int wi = ((x0s + dx) >> 8) + which;                    // floor: source word of the segment start, +1
            if ((unsigned)ys < (unsigned)S && (unsigned)zs < (unsigned)Z && wi >= 0 && wi < spr)
                word = __ldg(s.gmask + (zs * S + ys) * spr + wi);
        }
        return word;
    };
    for (int sl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; sl < nseg; sl += warps) {
        int seg = sl;
        if (rslab) {                                      // local segment counter -> (z, own row, x segment)
            const int jr = sl / spr, xsg = sl - jr * spr;
            const int zz = jr / my_rows, yi = jr - zz * my_rows;
            seg = (zz * S + (O.row_y0 + O.row_n * yi)) * spr + xsg;
        }
        const int row = seg / spr;                        // y + z*S
        const int x0s = (seg - row * spr) << 8;
        const int z = row / S, y = row - z * S;
        const unsigned old_word = O.gmask[seg];           // what the destination buffer holds here now
        int acc_and[8], sum[8], op[8], enc_or[8];
        unsigned srcm[8];                                 // ROWS: which sources are occupied at my voxels (for the cell kernel)
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc_and[j] = -1; sum[j] = 0; op[j] = -1; enc_or[j] = 0; srcm[j] = 0u; }
        unsigned seen = 0;                                // any source knows anything in this segment (uniform)
        for (int kb = 0; kb < A.n; kb += 16) {
            const unsigned word = mask_words(seg, kb);
            seen |= __ballot_sync(FULL, word != 0);
            const int kend = min(kb + 16, A.n);
            for (int k0 = kb; k0 < kend; k0 += NB) {
                unsigned need = 0;
                int xs[NB];
                const int* rowp[NB];
#pragma unroll
                for (int u = 0; u < NB; ++u) {
                    xs[u] = 0; rowp[u] = nullptr;
                    const int k = k0 + u;
                    if (k < kend) {                                       // uniform
                        const unsigned w0 = __shfl_sync(FULL, word, k - kb), w1 = __shfl_sync(FULL, word, 16 + k - kb);
                        if ((w0 | w1) != 0) {                             // uniform: the source holds something in this segment
                            const SlotRef& s = A.s[k];
                            const int dx = s.dx, dy = s.dy, dz = s.dz;
                            const int sx0 = x0s + dx;
                            const unsigned long long Wd = ((unsigned long long)w1 << 32) | w0;
                            const int b = ((sx0 >> 3) & 31) + lane;       // my first voxel's group, relative to word 0
                            const unsigned m = (sx0 & 7) ? 3u : 1u;       // an unaligned shift straddles two groups
                            if ((unsigned)(Wd >> b) & m) need |= 1u << u;
                            xs[u] = sx0 + 8 * lane;
                            rowp[u] = s.map + ((z + dz) * S + (y + dy)) * S;
                        }
                    }
                }
                int o[NB][8];
#pragma unroll
                for (int u = 0; u < NB; ++u) {
                    const bool enc = MODE == MERGE_FINISH && !(has_prev && k0 + u == A.n - 1);
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[u][j] = enc ? 0 : -1;   // "unknown": folds to nothing
                    if (need & (1u << u)) load_codes<8>(rowp[u], xs[u], S, o[u]);   // (encoded grids are unshifted: never out of range)
                }
#pragma unroll
                for (int u = 0; u < NB; ++u) {
                    if (!(need & (1u << u))) continue;
                    if (has_prev && k0 + u == A.n - 1) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) op[j] = o[u][j];
                    } else if (MODE == MERGE_FINISH) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) { enc_or[j] |= o[u][j]; sum[j] += o[u][j] < OCC_FLAG ? o[u][j] : 0; }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) { acc_and[j] &= o[u][j]; sum[j] += max(~o[u][j], 0); }
                        if (MASKS) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) srcm[j] |= (o[u][j] >= 0 ? 1u : 0u) << ((k0 + u) & 31);
                        }
                    }
                }
            }
        }
        if (seen == 0) {                                  // uniform: nothing known anywhere in the segment
            if (old_word != 0) {
                int4* dst = reinterpret_cast<int4*>(O.cmap) + (long long)seg * 64 + lane * 2;
                dst[0] = make_int4(UNK, UNK, UNK, UNK); dst[1] = make_int4(UNK, UNK, UNK, UNK);
                if (lane == 0) O.gmask[seg] = 0u;
            }
            continue;
        }
        bool occ[8];
        int c[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            occ[j] = (acc_and[j] >= 0) || (enc_or[j] >= OCC_FLAG);
            c[j] = -1 - sum[j];
            if (!occ[j]) {                                // previous combined map (gvom.py:1058-1063)
                if (op[j] >= 0) { if (c[j] >= -11) occ[j] = true; }
                else if (op[j] < -1) c[j] += op[j] + 1;
            }
        }
        unsigned m[8];
        int nocc = 0;
        bool any_free = false, my_occ = false;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            m[j] = __ballot_sync(FULL, occ[j]); nocc += __popc(m[j]);
            any_free |= (!occ[j] && c[j] < -1); my_occ |= occ[j];
        }
        int base = 0;
        if (nocc) {
            if (lane == 0) base = atomicAdd(O.counter, nocc);
            base = __shfl_sync(FULL, base, 0);
        }
        bool known_any = false;
        const int x = x0s + 8 * lane;
        if (MODE == MERGE_PARTIAL) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int rid = OCC_FLAG - 1;                    // "occupied, record dropped" (capacity overflow)
                if (occ[j]) {
                    const int id = base + __popc(m[j] & lt);
                    if (id < O.cap) { O.records[(long long)id * REC] = __int_as_float(seg * 256 + lane * 8 + j); rid = id; }
                }
                base += __popc(m[j]);
                // occupied: flag | record id (lets a finishing rank fetch the record without a search)
                c[j] = occ[j] ? (OCC_FLAG | rid) : min(sum[j], OCC_FLAG - 1);
                known_any |= c[j] != 0;
            }
        } else if (!my_occ && !any_free) {
#pragma unroll
            for (int j = 0; j < 8; ++j) c[j] = -1;
        } else {
            const bool cols = O.col_occ != nullptr;
            int* colo = O.col_occ + y * S + x;
            int* colf = O.col_free + y * S + x;
            int cur_occ[8], cur_free[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { cur_occ[j] = cols ? 0x7fffffff : -1; cur_free[j] = cols ? 0x7fffffff : -1; }
            if (cols) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    if (my_occ) {
                        const int4 a = *(reinterpret_cast<const int4*>(colo) + g);
                        cur_occ[4 * g] = a.x; cur_occ[4 * g + 1] = a.y; cur_occ[4 * g + 2] = a.z; cur_occ[4 * g + 3] = a.w;
                    }
                    if (any_free) {
                        const int4 b = *(reinterpret_cast<const int4*>(colf) + g);
                        cur_free[4 * g] = b.x; cur_free[4 * g + 1] = b.y; cur_free[4 * g + 2] = b.z; cur_free[4 * g + 3] = b.w;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (occ[j]) {
                    const int id = base + __popc(m[j] & lt);
                    if (id < O.cap) {
                        c[j] = id; O.cell_voxel[id] = seg * 256 + lane * 8 + j;
                        if (MASKS) O.srcmask[id] = srcm[j];
                        if (z < cur_occ[j]) atomicMin(colo + j, z);
                    } else c[j] = -1;
                } else if (c[j] < -1) {
                    if (z < cur_free[j]) atomicMin(colf + j, z);
                }
                base += __popc(m[j]);
                known_any |= c[j] != -1;
            }
        }
        const unsigned w = __ballot_sync(FULL, known_any);
        if (w != 0 || old_word != 0) {                    // uniform
            int4* dst = reinterpret_cast<int4*>(O.cmap) + (long long)seg * 64 + lane * 2;
            dst[0] = make_int4(c[0], c[1], c[2], c[3]); dst[1] = make_int4(c[4], c[5], c[6], c[7]);
            if (lane == 0) O.gmask[seg] = w;
        }
    }
}

template <int NB, int MODE>
__global__ void __launch_bounds__(256, (NB > 3) ? 2 : 3)
k_merge_rows(MergeArgs A, MergeOut O, DevParams P) {
    pdl_wait();
    merge_rows_body<NB, MODE>(A, O, P);
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
    // results are stored as float32: one float64 reciprocal instead of nine divisions changes them
    // by < 1e-15 relative (the parity bar for moments is 1e-4)
    const double n1 = c[9], n2 = o[9], nt = n1 + n2, inv = 1.0 / nt;
    const double c0 = c[0], c1 = c[1], c2 = c[2];
    const double mx = (c0 * n1 + o[0] * n2) * inv;
    const double my = (c1 * n1 + o[1] * n2) * inv;
    const double mz = (c2 * n1 + o[2] * n2) * inv;
    const double cd[3] = {c0 - mx, c1 - my, c2 - mz};
    const double od[3] = {o[0] - mx, o[1] - my, o[2] - mz};
    const int A[6] = {0, 0, 0, 1, 1, 2}, B[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
    for (int e = 0; e < 6; ++e) {
        const double v = (n1 * (double)c[3 + e] + n2 * o[3 + e] + n1 * cd[A[e]] * cd[B[e]] + n2 * od[A[e]] * od[B[e]]) * inv;
        c[3 + e] = (float)v;
    }
    c[0] = (float)mx; c[1] = (float)my; c[2] = (float)mz;
    c[9] = (float)nt;
}

// ---------------------------------------------------------------------------
// C2 (second build)  same results and the same fold order as k_merge_cells; the look-ups of ALL sources of a
// cell are issued together (up to 8 per round) instead of in batches of SLOT_BATCH.  (A build that also fetched the
// next record while merging the current one needed 96 registers and lost more to occupancy than it gained.)
// ---------------------------------------------------------------------------
struct CellRec { double o[10]; int hit, tot; float mh; };

__device__ __forceinline__ void load_cell_rec(const SlotRef& s, int io, CellRec& r) {
    if (s.is_prev) {
        const float2* om = reinterpret_cast<const float2*>(reinterpret_cast<const float*>(s.metrics) + (long long)io * 10);
#pragma unroll
        for (int a = 0; a < 5; ++a) { const float2 t = om[a]; r.o[2 * a] = (double)t.x; r.o[2 * a + 1] = (double)t.y; }
    } else {
        const double2* om = reinterpret_cast<const double2*>(reinterpret_cast<const double*>(s.metrics) + (long long)io * 10);
#pragma unroll
        for (int a = 0; a < 5; ++a) { const double2 t = om[a]; r.o[2 * a] = t.x; r.o[2 * a + 1] = t.y; }
    }
    r.hit = s.hit[io]; r.tot = s.total[io]; r.mh = s.minh[io];
}

struct NoCellHook { __device__ __forceinline__ void operator()(int, int, int, float) const {} };
// hook(x, y, z, min height) runs for every finished cell (the mirrored multi-GPU combine derives the column heights there)
template <typename Hook>
__device__ __forceinline__ void merge_cells2_body(const MergeArgs& A, const int* __restrict__ counter, const int* __restrict__ cell_voxel,
               int* __restrict__ chit, int* __restrict__ ctot, float* __restrict__ cminh,
               float* __restrict__ cmet, float* __restrict__ ceig, const DevParams& P, int cap, const Hook& hook,
               const unsigned* __restrict__ srcmask = nullptr) {
    const int count = min(*counter, cap);
    const int S = P.S, Z = P.Z;
    constexpr int RB = 8;                                   // sources looked up per round
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < count; id += gridDim.x * blockDim.x) {
        const int v = cell_voxel[id];
        const int x = v % S, y = (v / S) % S, z = v / (S * S);
        float c[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) c[k] = 0.f;
        int hit = 0, tot = 0;
        float mh = 1.0f;
        // (mirrored multi-GPU combine: the row merge left a mask of the sources that are occupied here -- with 2N + 1
        // sources most look-ups would find nothing)
        const unsigned want = srcmask ? srcmask[id] : 0xffffffffu;
        for (int k0 = 0; k0 < A.n; k0 += RB) {
            int io[RB];
#pragma unroll
            for (int u = 0; u < RB; ++u) {                  // independent index loads first
                io[u] = -1;
                if (k0 + u < A.n && (!srcmask || A.s[k0 + u].is_prev || ((want >> ((k0 + u) & 31)) & 1u))) {
                    const SlotRef& s = A.s[k0 + u];
                    const int xs = x + s.dx, ys = y + s.dy, zs = z + s.dz;
                    if ((unsigned)xs < (unsigned)S && (unsigned)ys < (unsigned)S && (unsigned)zs < (unsigned)Z)
                        io[u] = __ldg(s.map + (xs + (ys + zs * S) * S));
                }
            }
#pragma unroll
            for (int u = 0; u < RB; ++u) {
                if (io[u] < 0) continue;
                CellRec rc;
                load_cell_rec(A.s[k0 + u], io[u], rc);
                merge_step(c, rc.o);
                hit += rc.hit; tot += rc.tot; mh = fminf(mh, rc.mh);
            }
        }
        float* mo = cmet + (long long)id * 10;
#pragma unroll
        for (int k = 0; k < 10; ++k) mo[k] = c[k];
        chit[id] = hit; ctot[id] = tot; cminh[id] = mh;
        hook(x, y, z, mh);
        float e[3];
        eigen3(c, e);
        ceig[id * 3 + 0] = e[0]; ceig[id * 3 + 1] = e[1]; ceig[id * 3 + 2] = e[2];
    }
}

__global__ void __launch_bounds__(128, 8)
k_merge_cells2(MergeArgs A, const int* __restrict__ counter, const int* __restrict__ cell_voxel,
               int* __restrict__ chit, int* __restrict__ ctot, float* __restrict__ cminh,
               float* __restrict__ cmet, float* __restrict__ ceig, DevParams P, int cap) {
    pdl_wait();
    merge_cells2_body(A, counter, cell_voxel, chit, ctot, cminh, cmet, ceig, P, cap, NoCellHook{});
}

// 2-D maps are indexed [x,y] row-major like the reference's numpy outputs
#define GVOM_HM(a, x, y) (a)[(long long)(x) * S + (y)]

// ---------------------------------------------------------------------------
// C3  height maps from the column minima of C1.  Result of __make_height_map and
// __make_inferred_height_map (gvom.py:560-590) incl. their -1000 fills.  Also
// builds two bit maps of "height known" cells for the ring search of C4:
//   known [x][y/32] (bits over y) and knownT[y][x/32] (bits over x).
// Block = 32x32 cell tile, 1024 threads (x fastest), one column per thread.
// Also publishes the combined cell count (thread 0): C1's running counter -> the map's counter and, through a
// mapped pinned word, the host.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_column_maps(const int* __restrict__ cmap, const float* __restrict__ cminh, const int* __restrict__ col_occ,
              const int* __restrict__ col_free, double o0, double o1, double o2, double e0, double e1, double e2,
              DevParams P, double* __restrict__ height, double* __restrict__ inferred,
              unsigned* __restrict__ known, unsigned* __restrict__ knownT,
              const int* __restrict__ scratch_count, int* __restrict__ map_count, int* __restrict__ host_count) {
    pdl_wait();
    __shared__ unsigned char flag[32][33];
    const int S = P.S;
    const int W = (S + 31) >> 5;                          // words per bit row
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int x = blockIdx.x * 32 + tx, y = blockIdx.y * 32 + ty;
    const long long zs = (long long)S * S;
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
        const int n = *scratch_count;
        *map_count = n;
        if (host_count) *host_count = n;                   // mapped pinned word: the host reads the count without a DMA op
    }
    bool kn = false;
    if (x < S && y < S) {
        double h = -1000.0, inf = -1000.0;
        // ego disc: xp = (o+x)*res - ego with product and difference fused (ptxas contracts the
        // reference's mul+sub), then fma(xp,xp, yp*yp) <= r*r
        const double xp = __fma_rn(__dadd_rn(o0, (double)x), P.xy_res, -e0);
        const double yp = __fma_rn(__dadd_rn(o1, (double)y), P.xy_res, -e1);
        if (__fma_rn(xp, xp, __dmul_rn(yp, yp)) <= P.r2) h = __dsub_rn(e2, P.ground_to_lidar);
        const int zo = col_occ[(long long)y * S + x], zf = col_free[(long long)y * S + x];
        if (zo < P.Z) {
            const int idx = cmap[x + (long long)y * S + zo * zs];
            h = __dmul_rn(__dadd_rn(__dadd_rn((double)zo, (double)cminh[idx]), o2), P.z_res);
        }
        if (zf < P.Z) inf = __dmul_rn(__dadd_rn(o2, (double)zf), P.z_res);
        GVOM_HM(height, x, y) = h;
        GVOM_HM(inferred, x, y) = inf;
        kn = h > -1000.0;
    }
    flag[ty][tx] = kn ? 1 : 0;
    const unsigned wT = __ballot_sync(FULL, kn);          // bits over x for row y
    if (tx == 0 && y < S) knownT[(long long)y * W + blockIdx.x] = wT;
    __syncthreads();
    if (threadIdx.x < 32) {                               // bits over y for column x = tile x0 + threadIdx.x
        unsigned w = 0;
#pragma unroll
        for (int b = 0; b < 32; ++b) w |= (unsigned)flag[b][threadIdx.x] << b;
        const int xx = blockIdx.x * 32 + threadIdx.x;
        if (xx < S) known[(long long)xx * W + blockIdx.y] = w;
    }
}

// Row-sharded multi-GPU combine: rank r owns the grid rows y with (y + origin_y) mod n == r -- whole columns, so the
// column reductions and the 2-D stage of its rows are local -- i.e. the rows y0, y0 + n, ... (nrows of them).
struct RowShard { int y0, n, nrows; };
// The 2-D maps of a row-sharded combine are replicated by PUSHING: the owner of a row stores its values into every
// rank's 2-D block (symmetric memory, same layout everywhere; remote stores are posted).  n == 0: single GPU.
struct PushSet {
    char* base[16];          // every rank's 2-D block as mapped here
    int n, self;
    long long off_maps;      // six float64 maps [height, inferred, -, x slope, y slope, guessed]
    long long off_pos, off_neg, off_vis, off_rough;
};
template <typename T>
__device__ __forceinline__ void push_value(const PushSet& D, long long off, long long idx, T v) {
    for (int d = 0; d < D.n; ++d)
        if (d != D.self) reinterpret_cast<T*>(D.base[d] + off)[idx] = v;
}

// 32 bits of a bit row starting at bit position `pos` (may be negative / run past the row: those bits read 0)
__device__ __forceinline__ unsigned row_window(const unsigned* __restrict__ row, int W, int pos) {
    const int wi = pos >> 5;                                  // floor
    const unsigned lo = (wi >= 0 && wi < W) ? row[wi] : 0u;
    const unsigned hi = (wi + 1 >= 0 && wi + 1 < W) ? row[wi + 1] : 0u;
    return __funnelshift_r(lo, hi, pos & 31);
}

// ---------------------------------------------------------------------------
// C4 (second build)  same results as k_surface_maps, restructured for latency:
//   * the ring search of __guess_height is branch-free per ring: every row / column segment it inspects lies
//     within +-16 cells of the cell, so it is ONE 32-bit window of the "height known" bit rows (two shared-memory
//     words + a funnel shift) ANDed with the wedge mask of the ring, and find-first-set gives the reference's
//     first hit (it scans ascending).  Heights of the (at most four) cells found are loaded after the search,
//     together.
//   * only the ring-search warps stage the bit maps and wait for them (named barrier); the plane-fit warps
//     start on their height loads at once.
//   * optional second output set (pos2 ...): device-resident caller buffers are written by the kernel itself.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void
surface_maps_body(const int* __restrict__ cmap, const int* __restrict__ chit, const int* __restrict__ ctot,
                const double* __restrict__ height, const double* __restrict__ inferred,
                const unsigned* __restrict__ known_g, const unsigned* __restrict__ knownT_g, double o2,
                DevParams P, double* __restrict__ rough, double* __restrict__ xs, double* __restrict__ ys,
                double* __restrict__ guessed, int* __restrict__ pos, int* __restrict__ neg, int* __restrict__ vis,
                int masks_in_smem, int* __restrict__ col_minz, int* __restrict__ scratch_count,
                int* __restrict__ pos2, int* __restrict__ neg2, int* __restrict__ vis2, double* __restrict__ rough2,
                RowShard R, PushSet D) {
    extern __shared__ unsigned smask[];
    const int S = P.S, Z = P.Z;
    const int W = (S + 31) >> 5;
    const int role = threadIdx.x >> 7;
    const int t = blockIdx.x * 128 + (threadIdx.x & 127);
    const int ncell = S * R.nrows;                        // cells this launch owns (all of them on a single GPU)
    const long long S2 = (long long)S * S;
    // 2-D maps are [x][y] like the reference's outputs; in the row-sharded multi-GPU combine (D.n > 0) the exchange block
    // is [y][x] and threads run along x, so that the rows a rank owns are contiguous and the pushes coalesce
    const bool TR = D.n > 0;
#define HMX(a, x, y) (a)[TR ? (long long)(y) * S + (x) : (long long)(x) * S + (y)]
    if (role == 1) {
        const unsigned* known = known_g;
        const unsigned* knownT = knownT_g;
        if (masks_in_smem) {
            const int nw4 = (2 * S * W) >> 2;              // known and knownT are contiguous
            const uint4* g4 = reinterpret_cast<const uint4*>(known_g);
            uint4* s4 = reinterpret_cast<uint4*>(smask);
#pragma unroll 4
            for (int k = threadIdx.x & 127; k < nw4; k += 128) s4[k] = __ldg(g4 + k);
            asm volatile("bar.sync 1, 128;" ::: "memory");   // the four ring-search warps only
            known = smask;
            knownT = smask + S * W;
        }
        if (t >= ncell) return;
        const int y0 = TR ? R.y0 + R.n * (t / S) : t % S, x0 = TR ? t % S : t / S;
        // housekeeping for the next combine: C1's column minima and running counter start clean
        col_minz[y0 * S + x0] = 0x7f7f7f7f;
        col_minz[S2 + y0 * S + x0] = 0x7f7f7f7f;
        if (t == 0) *scratch_count = 0;
        const double h0 = HMX(height, x0, y0);
        const double inf0 = HMX(inferred, x0, y0);
        double dh_out = 0.0;
        if (!(h0 > -1000.0) && inf0 != -1000.0) {
            // ---- guessed height delta (gvom.py:592-713), quirks kept (see oracle/gvom_oracle.c)
            bool xpd = false, xnd = false, ypd = false, ynd = false;
            int fxp = -1, fxn = -1, fyp = -1, fyn = -1;      // found coordinate along the scanned row, -1: none
            int ixp = 0, ixn = 0, iyp = 0, iyn = 0;          // ring index at which it was found
            const int yb = y0 - 16, xb = x0 - 16;
            int i = 0;
            while (i < 15 && !(xnd && ypd && ynd)) {          // x_p_done is NOT part of the condition (gvom.py:619)
                i += 1;
                const unsigned span = (1u << (2 * i)) - 1u;
                const unsigned mP = span << (16 - i);         // offsets [-i, i-1]
                const unsigned mN = span << (17 - i);         // offsets [-i+1, i]
                const int x_p = x0 + i, x_n = x0 - i, y_p = y0 + i, y_n = y0 - i;
                if (!xpd) {
                    if (x_p < S) {
                        const unsigned m = row_window(known + x_p * W, W, yb) & mP;
                        if (m) { fxp = yb + __ffs(m) - 1; ixp = i; xpd = true; }
                    } else xpd = true;
                }
                if (!xnd) {
                    if (x_n >= 0) {
                        const unsigned m = row_window(known + x_n * W, W, yb) & mN;
                        if (m) { fxn = yb + __ffs(m) - 1; ixn = i; xnd = true; }
                    } else xnd = true;
                }
                if (!ypd) {
                    if (y_p < S) {
                        const unsigned m = row_window(knownT + y_p * W, W, xb) & mN;
                        if (m) { fyp = xb + __ffs(m) - 1; iyp = i; ypd = true; }
                    } else ypd = true;
                }
                if (!ynd) {
                    if (y_n >= 0) {
                        const unsigned m = row_window(knownT + y_n * W, W, xb) & mP;
                        if (m) { fyn = xb + __ffs(m) - 1; iyn = i; ynd = true; }
                    } else ynd = true;
                }
            }
            double x_ph = -1000.0, x_nh = -1000.0, y_ph = -1000.0, y_nh = -1000.0;
            if (fxp >= 0) x_ph = HMX(height, x0 + ixp, fxp);
            if (fxn >= 0) x_nh = HMX(height, x0 - ixn, fxn);
            if (fyp >= 0) y_ph = HMX(height, fyp, y0 + iyp);
            if (fyn >= 0) y_nh = HMX(height, fyn, y0 - iyn);
            double mn = 1000.0, mx = inf0;
            if (x_ph > -1000.0) { mn = fmin(x_ph, mn); mx = fmax(x_ph, mx); }
            if (x_nh > -1000.0) { mn = fmin(x_nh, mn); mx = fmax(x_nh, mx); }
            if (y_ph > -1000.0) { mn = fmin(y_ph, mn); mx = fmax(y_ph, mx); }
            if (x_nh > -1000.0) { mn = fmin(y_nh, mn); mx = fmax(y_nh, mx); }   // sic (gvom.py:704-706)
            const double dh = __dsub_rn(mx, mn);
            if (dh > 0.0) dh_out = dh;
        }
        const int nv = dh_out > P.neg_thr ? 100 : 0, vv = h0 > -1000.0 ? 1 : 0;
        HMX(guessed, x0, y0) = dh_out;
        HMX(neg, x0, y0) = nv;
        HMX(vis, x0, y0) = vv;
        if (neg2) { HMX(neg2, x0, y0) = nv; HMX(vis2, x0, y0) = vv; }
        if (D.n) {
            const long long ci = (long long)y0 * S + x0;
            push_value<double>(D, D.off_maps + 5 * S2 * 8, ci, dh_out);
            push_value<int>(D, D.off_neg, ci, nv);
            push_value<int>(D, D.off_vis, ci, vv);
        }
        return;
    }
    if (t >= ncell) return;
    const int y0 = TR ? R.y0 + R.n * (t / S) : t % S, x0 = TR ? t % S : t / S;

    // ---- slope + roughness (gvom.py:717-805); contraction pattern = SASS of the reference.
    double hz[9];
    unsigned okm = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const int x = x0 - 1 + a, y = y0 - 1 + b;
            double h = -1000.0;
            if (x >= 0 && x < S && y >= 0 && y < S) h = HMX(height, x, y);
            hz[a * 3 + b] = h;
            if (h > -1000.0) okm |= 1u << (a * 3 + b);
        }
    const double h0 = hz[4];
    double sxv = 0.0, syv = 0.0, rg = -1.0;
    const int n = __popc(okm);
    if (n >= 3) {
        double sx = 0, sy = 0, sz = 0;
#pragma unroll
        for (int j = 0; j < 9; ++j)
            if (okm & (1u << j)) {
                sx = __dadd_rn(sx, __dmul_rn((double)(x0 - 1 + j / 3), P.xy_res));
                sy = __dadd_rn(sy, __dmul_rn((double)(y0 - 1 + j % 3), P.xy_res));
                sz = __dadd_rn(sz, hz[j]);
            }
        const double dn = (double)n;
        const double mx = __ddiv_rn(sx, dn), my = __ddiv_rn(sy, dn), mz = __ddiv_rn(sz, dn);
        double xx = 0, xy = 0, xz = 0, yy = 0, yz = 0;
#pragma unroll
        for (int j = 0; j < 9; ++j)
            if (okm & (1u << j)) {
                const double dx = __dsub_rn(__dmul_rn((double)(x0 - 1 + j / 3), P.xy_res), mx);
                const double dy = __dsub_rn(__dmul_rn((double)(y0 - 1 + j % 3), P.xy_res), my);
                const double dz = __dsub_rn(hz[j], mz);
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
#pragma unroll
            for (int j = 0; j < 9; ++j)
                if (okm & (1u << j)) {
                    const double dx = __dsub_rn(__dmul_rn((double)(x0 - 1 + j / 3), P.xy_res), mx);
                    const double dy = __dsub_rn(__dmul_rn((double)(y0 - 1 + j % 3), P.xy_res), my);
                    const double e = __dsub_rn(__dsub_rn(hz[j], mz), __fma_rn(a0, dx, __dmul_rn(a1, dy)));
                    err = __fma_rn(e, e, err);
                }
            err = __ddiv_rn(err, dn);
            if (err > 0.0) err = log(err);
            rg = err;
            const double im = __drcp_rn(m);
            sxv = atan2(a0, im);
            syv = atan2(a1, im);
        }
    }
    HMX(rough, x0, y0) = rg;
    if (rough2) HMX(rough2, x0, y0) = rg;
    HMX(xs, x0, y0) = sxv;
    HMX(ys, x0, y0) = syv;

    // ---- positive obstacles (gvom.py:515-555)
    int pv = 0;
    const double sl = __dsqrt_rn(__fma_rn(sxv, sxv, __dmul_rn(syv, syv)));
    if (!(sl < P.slope_thr)) {
        pv = 100;
    } else {
        const double lo = floor(__dsub_rn(__ddiv_rn(__dadd_rn(h0, P.pos_thr), P.z_res), o2));
        const double hi = floor(__dsub_rn(__ddiv_rn(__dadd_rn(h0, P.robot_height), P.z_res), o2));
        if (lo > -2.0e9 && lo < 2.0e9 && hi > -2.0e9 && hi < 2.0e9) {
            const long long zlo = (long long)lo + 1, zhi = (long long)hi;
            if (zlo >= 0 && zlo < Z && zhi >= 0 && zhi < Z) {
                double density = 0.0, nn = 0.0;
                for (long long zb = zlo; zb <= zhi; zb += 8) {        // 8 levels per batch: loads first
                    int idx[8], hc[8], tc[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        idx[u] = (zb + u <= zhi) ? __ldg(cmap + (x0 + (y0 + (zb + u) * S) * S)) : -1;
#pragma unroll
                    for (int u = 0; u < 8; ++u) { hc[u] = idx[u] >= 0 ? __ldg(chit + idx[u]) : 0; tc[u] = idx[u] >= 0 ? __ldg(ctot + idx[u]) : 0; }
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (hc[u] > 10) { nn = __dadd_rn(nn, (double)tc[u]); density = __dadd_rn(density, (double)hc[u]); }
                }
                if (nn > 0.0) density = __ddiv_rn(density, nn);
                pv = (int)__dmul_rn(density, 100.0);
            }
        }
    }
    HMX(pos, x0, y0) = pv;
    if (pos2) HMX(pos2, x0, y0) = pv;
    if (D.n) {
        const long long ci = (long long)y0 * S + x0;
        push_value<double>(D, D.off_rough, ci, rg);
        push_value<double>(D, D.off_maps + 3 * S2 * 8, ci, sxv);
        push_value<double>(D, D.off_maps + 4 * S2 * 8, ci, syv);
        push_value<int>(D, D.off_pos, ci, pv);
    }
}
#undef HMX

__global__ void __launch_bounds__(256, 4)
k_surface_maps2(const int* __restrict__ cmap, const int* __restrict__ chit, const int* __restrict__ ctot,
                const double* __restrict__ height, const double* __restrict__ inferred,
                const unsigned* __restrict__ known_g, const unsigned* __restrict__ knownT_g, double o2,
                DevParams P, double* __restrict__ rough, double* __restrict__ xs, double* __restrict__ ys,
                double* __restrict__ guessed, int* __restrict__ pos, int* __restrict__ neg, int* __restrict__ vis,
                int masks_in_smem, int* __restrict__ col_minz, int* __restrict__ scratch_count,
                int* __restrict__ pos2, int* __restrict__ neg2, int* __restrict__ vis2, double* __restrict__ rough2,
                RowShard R, PushSet D, GridSignal G) {
    pdl_wait();
    surface_maps_body(cmap, chit, ctot, height, inferred, known_g, knownT_g, o2, P, rough, xs, ys, guessed, pos, neg, vis,
                      masks_in_smem, col_minz, scratch_count, pos2, neg2, vis2, rough2, R, D);
    signal_when_grid_done(G);                             // row-sharded combine: this rank's maps are in every rank's block
}

// ---------------------------------------------------------------------------
// Row-sharded multi-GPU combine, 2-D stage helpers.
//   k_rows_known     after the height barrier: the "height known" bit maps of the WHOLE map (every rank builds its own)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_rows_known(const double* __restrict__ height, DevParams P, unsigned* __restrict__ known, unsigned* __restrict__ knownT,
             const int* __restrict__ wait_flags, int wait_n, int wait_epoch, SignalSet pub) {
    pdl_wait();
    publish_flag_first_block(pub, wait_epoch);             // (mirrored combine) this rank's heights are pushed: tell every rank
    wait_flags_block(wait_flags, wait_n, wait_epoch);      // every rank has pushed the heights of its rows
    __shared__ unsigned char flag[32][33];
    const int S = P.S;
    const int W = (S + 31) >> 5;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int x = blockIdx.x * 32 + tx, y = blockIdx.y * 32 + ty;
    bool kn = false;
    if (x < S && y < S) kn = __ldcg(height + (long long)y * S + x) > -1000.0;      // exchange block: [y][x]
    flag[ty][tx] = kn ? 1 : 0;
    const unsigned wT = __ballot_sync(FULL, kn);          // bits over x for row y
    if (tx == 0 && y < S) knownT[(long long)y * W + blockIdx.x] = wT;
    __syncthreads();
    if (threadIdx.x < 32) {                               // bits over y for column x = tile x0 + threadIdx.x
        unsigned w = 0;
#pragma unroll
        for (int b = 0; b < 32; ++b) w |= (unsigned)flag[b][threadIdx.x] << b;
        const int xx = blockIdx.x * 32 + threadIdx.x;
        if (xx < S) known[(long long)xx * W + blockIdx.y] = w;
    }
}

// Delivery of a row-sharded combine: waits (device side) until every rank has pushed the maps of its rows, then
// transposes the exchange block ([y][x]) into the library's own 2-D block and, if given, the caller's buffers
// (device memory or mapped pinned host memory), which are [x][y] like the reference's outputs.
struct MapSet { double* maps6; int* pos; int* neg; int* vis; double* rough; };
__global__ void __launch_bounds__(256)
k_rows_deliver(const char* __restrict__ blk, PushSet D, int S, MapSet own, MapSet user,
               const int* __restrict__ wait_flags, int wait_n, int wait_epoch, int m0, SignalSet pub) {
    pdl_wait();
    publish_flag_first_block(pub, wait_epoch);             // (mirrored combine) this rank's maps are pushed: tell every rank
    wait_flags_block(wait_flags, wait_n, wait_epoch);
    __shared__ double td[32][33];
    __shared__ int ti[32][33];
    const long long S2 = (long long)S * S;
    const int tx = threadIdx.x & 31, ty0 = threadIdx.x >> 5;
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    const int m = blockIdx.z + m0;                         // 0..5: float64 maps of maps6, 6: roughness, 7..9: pos / neg / vis
    const bool is_d = m < 7;
    const double* sd = m < 6 ? reinterpret_cast<const double*>(blk + D.off_maps) + m * S2 : reinterpret_cast<const double*>(blk + D.off_rough);
    const int* si = reinterpret_cast<const int*>(blk + (m == 7 ? D.off_pos : m == 8 ? D.off_neg : D.off_vis));
#pragma unroll
    for (int r = 0; r < 4; ++r) {                          // read along x (the block's fast axis)
        const int yl = ty0 + 8 * r, y = by + yl, x = bx + tx;
        if (x < S && y < S) {
            if (is_d) td[yl][tx] = __ldcg(sd + (long long)y * S + x); else ti[yl][tx] = __ldcg(si + (long long)y * S + x);
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {                          // write along y (the outputs' fast axis)
        const int xl = ty0 + 8 * r, x = bx + xl, y = by + tx;
        if (x < S && y < S) {
            const long long o = (long long)x * S + y;
            if (is_d) {
                const double v = td[tx][xl];
                if (m < 6) own.maps6[m * S2 + o] = v;
                else { own.rough[o] = v; if (user.rough) user.rough[o] = v; }
            } else {
                const int v = ti[tx][xl];
                int* a = m == 7 ? own.pos : m == 8 ? own.neg : own.vis;
                int* u = m == 7 ? user.pos : m == 8 ? user.neg : user.vis;
                a[o] = v;
                if (u) u[o] = v;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// OccupancyGrid post-processing of the reference's ROS node (gvom_ros.py:142-164), on the device: five int8
// grids in Fortran order (cell [x,y] -> y*S + x) from the result block of the last combine.  numpy semantics
// are kept: int32 comparisons against a Python float threshold are float64 comparisons, np.minimum / np.maximum
// propagate NaN, and astype(int8) truncates toward zero and keeps the low byte (x86 cvttsd2si: out of range -> 0).
// ---------------------------------------------------------------------------
__device__ __forceinline__ signed char as_int8(double v) {
    const int t = (v >= -2147483648.0 && v < 2147483648.0) ? __double2int_rz(v) : (int)0x80000000;
    return (signed char)(t & 0xff);
}

__global__ void __launch_bounds__(256)
k_occupancy_grids(const int* __restrict__ pos, const int* __restrict__ neg, const int* __restrict__ vis,
                  const double* __restrict__ rough, int S, double thr, double minr, double maxr,
                  signed char* __restrict__ out) {
    pdl_wait();
    // 32x32 tile transpose through shared memory: reads run along y (the maps' fast axis), writes along x
    __shared__ int tp[32][33], tn[32][33], tv[32][33];
    __shared__ double tr[32][33];
    const int tx = threadIdx.x & 31, ty0 = threadIdx.x >> 5;      // 8 rows per pass
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int xl = ty0 + 8 * r, x = bx + xl, y = by + tx;
        if (x < S && y < S) {
            const long long a = (long long)x * S + y;
            tp[xl][tx] = pos[a]; tn[xl][tx] = neg[a]; tv[xl][tx] = vis[a]; tr[xl][tx] = rough[a];
        }
    }
    __syncthreads();
    const long long S2 = (long long)S * S;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int yl = ty0 + 8 * r, y = by + yl, x = bx + tx;
        if (x < S && y < S) {
            const int p = tp[tx][yl], n = tn[tx][yl], v = tv[tx][yl];
            const double rg = tr[tx][yl];
            const long long o = (long long)y * S + x;
            const int hard = max(((double)p > thr) ? 100 : 0, n);
            const int soft = (((double)p <= thr) && p > 0) ? 100 : 0;
            double c = (rg != rg) ? rg : (rg < maxr ? rg : maxr);          // np.minimum(rough, max)
            c = (c != c) ? c : (c > minr ? c : minr);                      // np.maximum(., min)
            const double q = __dmul_rn(__ddiv_rn(__dadd_rn(c, minr), __dsub_rn(maxr, minr)), 100.0);
            out[0 * S2 + o] = (signed char)(hard & 0xff);
            out[1 * S2 + o] = (signed char)(soft & 0xff);
            out[2 * S2 + o] = (signed char)((v * 100) & 0xff);
            out[3 * S2 + o] = (signed char)(n & 0xff);
            out[4 * S2 + o] = as_int8(q);
        }
    }
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

// Per-rank exchange buffers as the finishing rank sees them.  With the NCCL exchange these
// point into local (all-reduced / all-gathered) memory; with the peer-to-peer exchange they are
// the OTHER GPUs' buffers mapped over NVLink, read directly by the kernels below.
struct RecordSet { const float* r[MAX_RANKS]; const int* count[MAX_RANKS]; int n; };

// header variant: {epoch, ox, oy, oz} -- the origin first, then (after a fence) the epoch the waiters poll
// publish "my partial results of combine `epoch` are complete" into every rank's flag slot for this rank
__global__ void k_signal(SignalSet S, int epoch) {
    pdl_wait();                                   // after this rank's partial kernels
    __threadfence_system();
    if (threadIdx.x < S.n) {
        volatile int* f = S.slot[threadIdx.x];
        *f = epoch;
    }
    __threadfence_system();
}

// per record: fold this rank's slots (reference order, float32 rounding as in C2)
__global__ void __launch_bounds__(128)
k_partial_cells(MergeArgs A, const int* __restrict__ counter, float* __restrict__ records, DevParams P, int cap, GridSignal G) {
    pdl_wait();
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
    signal_when_grid_done(G);                             // partial results complete: tell every rank
}

// scatter every rank's records into the cells: raw moments n, n*mu, n*(C + mu mu^T) in float64
__global__ void __launch_bounds__(256)
k_scatter_records(RecordSet R, long long capacity,
                  const int* __restrict__ cmap, double* __restrict__ cacc, int* __restrict__ chit,
                  int* __restrict__ ctot, float* __restrict__ cminh) {
    pdl_wait();
    for (int rk = 0; rk < R.n; ++rk) {
        const int count = (int)min((long long)*R.count[rk], capacity);
        const float* base = R.r[rk];
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
    pdl_wait();
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
