// gvom_mirror.cuh -- multi-GPU combine_maps with MIRRORED ring slots (sm_100a, NVLink / NVSwitch peer memory).
//
// The reference keeps one ring buffer of per-scan maps and merges all of it on every combine_maps()
// (gvom.py:242-257).  With one sensor stream per GPU the ring is spread over the ranks.  The combined map is sharded
// by WORLD rows (rank r owns the grid rows y with (y + origin_y) mod n == r, whole columns), and instead of
// exchanging pre-merged grids at combine time, every scan is delivered to the owners of its rows when it is made:
//
//   k_push_scan   (end of Process_pointcloud)  the new slot's index-map rows, group-mask words and cell records are
//                 stored into the mirror of (this rank, ring slot) in the memory of the rank that owns the row --
//                 posted NVLink writes, nothing waits for them.  Every rank therefore holds, for the rows it owns,
//                 an exact local copy of EVERY rank's ring slots.
//   k_mirror_args (start of combine_maps)      one warp: publishes "everything I scanned is pushed" (epoch flag in
//                 every rank's block), waits for all ranks' flags, and builds the source list of the merge from the
//                 slot table the pushes carried (slot origins are only known to the rank that made the scan) --
//                 no host round trip, no collective.
//   then the single-GPU merge kernels run over the own rows with the mirrors as sources (k_merge_rows_ind,
//   k_merge_cells2_rows): local memory only.  Beside the cell kernel, on a second stream, k_rows_heights pushes the
//   heights of the own columns and k_rows_known exchanges the "heights" flags and builds the bit maps; after the join
//   the surface stage of the own rows pushes the finished maps.  Only 2-D maps cross the links at combine time.
//
// The rows a rank owns in a mirror are exact: the pusher remembers per (destination rank, 256-voxel segment) which
// group-mask word it last delivered ("held"), overwrites whole segments, and wipes a segment at the owner when the
// new scan knows nothing there but the owner still holds codes of an earlier scan.  Rows a rank does not own under
// the slot's current origin may hold leftovers (the ego moved and the row changed owner); nobody reads them.
#pragma once
#include "gvom_kernels.cuh"
#include "gvom_merge.cuh"      // mbarrier / bulk-copy helpers

namespace gvom {

constexpr int MIRROR_ENTRY = 16;     // ints per slot-table entry: {seq (0: empty), ox, oy, oz, cells, -, -, -, ego xyz (3 float64), -, -}

struct MirrorPush {
    char* base[MAX_RANKS];           // every rank's mirror block as mapped here (own block included)
    int n, self;
    long long o_map, o_gmask, o_hit, o_tot, o_minh, o_met;   // byte offsets of mirror (self, ring slot) inside a block
    long long o_entry;               // byte offset of its slot-table entry
    unsigned* held;                  // local [n][nsegp]: group-mask word last delivered to rank r for a segment
    int nsegp;
    int oy;                          // origin y of the scan (voxels): decides the owner of a row
    int entry[MIRROR_ENTRY];
};

__device__ __forceinline__ int owner_of_row(int y, int oy, int n) {
    int m = (y + oy) % n;
    return m < 0 ? m + n : m;
}

__global__ void __launch_bounds__(256, 4)
k_push_scan(const int* __restrict__ map, const unsigned* __restrict__ gmask, const int* __restrict__ count_ptr,
            const int* __restrict__ cell_voxel, const int* __restrict__ hit, const int* __restrict__ total,
            const float* __restrict__ minh, const double* __restrict__ metrics, MirrorPush M, DevParams P, int cap) {
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int S = P.S;
    const int spr = S >> 8;
    const int nseg = (int)(P.V >> 8);
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    if (blockIdx.x == 0 && threadIdx.x < M.n) {           // slot table: {seq, origin, cells} of this slot, to every rank
        volatile int* e = reinterpret_cast<volatile int*>(M.base[threadIdx.x] + M.o_entry);
#pragma unroll
        for (int k = 1; k < MIRROR_ENTRY; ++k) e[k] = k == 4 ? min(*count_ptr, cap) : M.entry[k];
        e[0] = M.entry[0];
    }
    // ---- index-map rows: whole 256-voxel segments to the owner of the row.  A rank only ever reads the rows it owns
    // under the slot's CURRENT origin, so what an earlier scan of this ring slot left in other ranks' mirrors is never
    // looked at; `held` remembers it, and when such a row comes back to that rank it is overwritten or wiped then.
    constexpr int U = 4;                                   // segments in flight per warp: one pass of the grid covers 256x256x64
    for (int s0 = warp * U; s0 < nseg; s0 += nwarps * U) {
        unsigned nw[U], hw[U];
        int own[U];
        int4 a[U], b[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int seg = s0 + u;
            nw[u] = 0; hw[u] = 0; own[u] = 0;
            a[u] = make_int4(-1, -1, -1, -1); b[u] = a[u];
            if (seg < nseg) {
                own[u] = owner_of_row((seg / spr) % S, M.oy, M.n);
                nw[u] = __ldcg(gmask + seg);
                if (nw[u]) {                              // lane l moves int4 l and l + 32 of the segment: full 512-byte requests
                    const int4* src = reinterpret_cast<const int4*>(map) + (long long)seg * 64 + lane;
                    a[u] = __ldcg(src); b[u] = __ldcg(src + 32);
                } else {
                    hw[u] = M.held[(long long)own[u] * M.nsegp + seg];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int seg = s0 + u;
            if (seg >= nseg) break;
            if (nw[u] == 0u && hw[u] == 0u) continue;      // uniform: nothing known now, nothing stale at the owner
            const int r = own[u];
            int4* dst = reinterpret_cast<int4*>(M.base[r] + M.o_map) + (long long)seg * 64 + lane;
            dst[0] = a[u]; dst[32] = b[u];                 // the codes, or "unknown" over what the owner still holds
            if (lane == 0) {
                reinterpret_cast<unsigned*>(M.base[r] + M.o_gmask)[seg] = nw[u];
                M.held[(long long)r * M.nsegp + seg] = nw[u];
            }
        }
    }
    // ---- cell records: to the owner of the cell's row, at the cell's own id (the codes in the map rows refer to it)
    const int count = min(*count_ptr, cap);
    const int sub = lane & 7;
    const int gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, ngroups = (gridDim.x * blockDim.x) >> 3;
    for (int id = gid; id < count; id += ngroups) {
        const int v = cell_voxel[id];
        const int y = P.lgS >= 0 ? (v >> P.lgS) & (S - 1) : (v / S) % S;
        char* base = M.base[owner_of_row(y, M.oy, M.n)];
        if (sub < 5) reinterpret_cast<double2*>(base + M.o_met)[(long long)id * 5 + sub] = __ldcg(reinterpret_cast<const double2*>(metrics) + (long long)id * 5 + sub);
        else if (sub == 5) reinterpret_cast<int*>(base + M.o_hit)[id] = __ldcg(hit + id);
        else if (sub == 6) reinterpret_cast<int*>(base + M.o_tot)[id] = __ldcg(total + id);
        else reinterpret_cast<float*>(base + M.o_minh)[id] = __ldcg(minh + id);
    }
}

// ---------------------------------------------------------------------------
// k_push_scan with the index-map rows moved by the bulk-copy (TMA) engine: a warp's lane 0 asks for the 1 KB segments
// to be copied HBM -> shared memory (`cp.async.bulk.shared.global`, completion counted in bytes on an mbarrier) and,
// as each lands, from shared memory straight into the owner's mirror (`cp.async.bulk.global.shared::cta` to the peer
// address: 1 KB per request on the NVLink side instead of 64 sixteen-byte lane stores).  The (rare) wipes and the
// cell records use plain stores as in k_push_scan.  Same results; selected by GVOM_VARIANT bit 32.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void bulk_s2g(void* dst, unsigned src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(256, 4)
k_push_scan_bulk(const int* __restrict__ map, const unsigned* __restrict__ gmask, const int* __restrict__ count_ptr,
                 const int* __restrict__ cell_voxel, const int* __restrict__ hit, const int* __restrict__ total,
                 const float* __restrict__ minh, const double* __restrict__ metrics, const __grid_constant__ MirrorPush M,
                 DevParams P, int cap) {
    constexpr int U = 4;
    __shared__ __align__(128) int buf[8][U][256];           // 4 KB per warp
    __shared__ unsigned long long bars[8][U];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int S = P.S;
    const int spr = S >> 8;
    const int nseg = (int)(P.V >> 8);
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    if (lane == 0) {
#pragma unroll
        for (int u = 0; u < U; ++u) mbar_init(smem_u32(&bars[wib][u]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    pdl_wait();
    if (blockIdx.x == 0 && threadIdx.x < M.n) {           // slot table: {seq, origin, cells, ego} of this slot, to every rank
        volatile int* e = reinterpret_cast<volatile int*>(M.base[threadIdx.x] + M.o_entry);
#pragma unroll
        for (int k = 1; k < MIRROR_ENTRY; ++k) e[k] = k == 4 ? min(*count_ptr, cap) : M.entry[k];
        e[0] = M.entry[0];
    }
    unsigned parity = 0;
    for (int s0 = warp * U; s0 < nseg; s0 += nwarps * U) {
        unsigned nw[U], hw[U];
        int own[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int seg = s0 + u;
            nw[u] = 0; hw[u] = 0; own[u] = 0;
            if (seg < nseg) {
                own[u] = owner_of_row((seg / spr) % S, M.oy, M.n);
                nw[u] = __ldcg(gmask + seg);
                if (nw[u]) {
                    if (lane == 0) {
                        const unsigned bar = smem_u32(&bars[wib][u]);
                        mbar_expect_tx(bar, 1024u);
                        bulk_g2s(smem_u32(&buf[wib][u][0]), map + (long long)seg * 256, 1024u, bar);
                    }
                } else {
                    hw[u] = M.held[(long long)own[u] * M.nsegp + seg];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int seg = s0 + u;
            if (seg >= nseg) break;
            if (nw[u] == 0u && hw[u] == 0u) continue;      // uniform
            const int r = own[u];
            if (nw[u]) {
                if (lane == 0) {
                    mbar_wait(smem_u32(&bars[wib][u]), parity);
                    bulk_s2g(M.base[r] + M.o_map + (long long)seg * 1024, smem_u32(&buf[wib][u][0]), 1024u);
                }
            } else {                                       // the owner still holds codes of an earlier scan: "unknown" over them
                int4* dst = reinterpret_cast<int4*>(M.base[r] + M.o_map) + (long long)seg * 64 + lane;
                const int4 unk = make_int4(-1, -1, -1, -1);
                dst[0] = unk; dst[32] = unk;
            }
            if (lane == 0) {
                reinterpret_cast<unsigned*>(M.base[r] + M.o_gmask)[seg] = nw[u];
                M.held[(long long)r * M.nsegp + seg] = nw[u];
            }
        }
        if (lane == 0) {                                   // the staging buffers are reused by the next pass
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncwarp();
        parity ^= 1u;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    // ---- cell records: to the owner of the cell's row, at the cell's own id
    const int count = min(*count_ptr, cap);
    const int sub = lane & 7;
    const int gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, ngroups = (gridDim.x * blockDim.x) >> 3;
    for (int id = gid; id < count; id += ngroups) {
        const int v = cell_voxel[id];
        const int y = P.lgS >= 0 ? (v >> P.lgS) & (S - 1) : (v / S) % S;
        char* base = M.base[owner_of_row(y, M.oy, M.n)];
        if (sub < 5) reinterpret_cast<double2*>(base + M.o_met)[(long long)id * 5 + sub] = __ldcg(reinterpret_cast<const double2*>(metrics) + (long long)id * 5 + sub);
        else if (sub == 5) reinterpret_cast<int*>(base + M.o_hit)[id] = __ldcg(hit + id);
        else if (sub == 6) reinterpret_cast<int*>(base + M.o_tot)[id] = __ldcg(total + id);
        else reinterpret_cast<float*>(base + M.o_minh)[id] = __ldcg(minh + id);
    }
}

// Signal + wait + source list of a mirrored combine (one warp).
struct MirrorSync {
    int* flag_slot[MAX_RANKS];       // this rank's epoch flag in every rank's block
    const int* flags;                // local: n flags, written by the ranks
    const int* table;                // local slot table [n][B][MIRROR_ENTRY]
    char* mirrors;                   // local: mirror (g, k) starts at mirrors + (g * B + k) * mirror_bytes
    long long mirror_bytes, f_gmask, f_hit, f_tot, f_minh, f_met;   // field offsets inside a mirror (the map comes first)
    int n, B, epoch;
    int mode;                        // bit 0: publish my flag; bit 1: wait for everybody's and build the list
    int org[3];                      // origin the map is combined in (this rank's newest scan)
    SlotRef prev; int has_prev;      // this rank's rows of the previous combined map
    MergeArgs* out;
    int* err_flag;                   // raised when the ranks' newest scans disagree on the origin
};

// one warp: publish this rank's epoch flag (mode bit 0; the caller decides which warp of the grid does it)
__device__ __forceinline__ void mirror_signal(const MirrorSync& Y, int lane) {
    // the pushes of every earlier scan were made by earlier kernels of this stream; the fence orders them (and
    // anything else this rank wrote) before the flag for every observer
    __threadfence_system();
    if (lane < Y.n) *reinterpret_cast<volatile int*>(Y.flag_slot[lane]) = Y.epoch;
}

// one warp: wait for every rank's flag, then build the source list from the slot table.  Every lane takes the entries
// lane, lane + 32 (n * B <= 64); valid ones are compacted in table order, the previous map comes last.
__device__ __forceinline__ void mirror_wait_build(const MirrorSync& Y, MergeArgs& A, int lane) {
    if (lane < Y.n) {
        const volatile int* f = Y.flags + lane;
        if (!spin_until(f, Y.epoch) && Y.err_flag) *Y.err_flag = 2;     // a rank never arrived: reported by the host
        __threadfence_system();                            // (only the polling lanes: a system-scope fence is not cheap)
    }
    __syncwarp();
    const int total = Y.n * Y.B;
    int n = 0;
    bool bad = false;
    for (int base = 0; base < total; base += 32) {
        const int idx = base + lane;
        int seq = 0, ox = 0, oy = 0, oz = 0;
        if (idx < total) {
            const volatile int* e = Y.table + idx * MIRROR_ENTRY;
            seq = e[0]; ox = e[1]; oy = e[2]; oz = e[3];
        }
        const unsigned valid = __ballot_sync(FULL, seq != 0);
        if (seq != 0) {
            const int g = idx / Y.B;
            bool newest = true;                           // the newest scan of every rank defines the frame: all must agree
            for (int k = 0; k < Y.B; ++k) newest &= *reinterpret_cast<const volatile int*>(Y.table + (g * Y.B + k) * MIRROR_ENTRY) <= seq;
            if (newest && (ox != Y.org[0] || oy != Y.org[1] || oz != Y.org[2])) bad = true;
            char* m = Y.mirrors + (long long)idx * Y.mirror_bytes;
            SlotRef r;
            r.map = reinterpret_cast<const int*>(m);
            r.gmask = reinterpret_cast<const unsigned*>(m + Y.f_gmask);
            r.hit = reinterpret_cast<const int*>(m + Y.f_hit);
            r.total = reinterpret_cast<const int*>(m + Y.f_tot);
            r.minh = reinterpret_cast<const float*>(m + Y.f_minh);
            r.metrics = m + Y.f_met;
            r.dx = Y.org[0] - ox; r.dy = Y.org[1] - oy; r.dz = Y.org[2] - oz;
            r.is_prev = 0;
            A.s[n + __popc(valid & ((1u << lane) - 1u))] = r;
        }
        n += __popc(valid);
    }
    if (bad && Y.err_flag) *Y.err_flag = 1;
    if (lane == 0) {
        if (Y.has_prev) A.s[n++] = Y.prev;
        A.n = n;
        A.use_masks = 1;
    }
    __syncwarp();
}

// stand-alone form (tests, start-up path: a rank without scans publishes its flag and reads the table on the host)
__global__ void __launch_bounds__(32)
k_mirror_args(const __grid_constant__ MirrorSync Y) {
    pdl_wait();
    const int lane = threadIdx.x;
    if (Y.mode & 1) mirror_signal(Y, lane);
    if (Y.mode & 2) mirror_wait_build(Y, *Y.out, lane);
}

// C1 of the mirrored combine: the row merge of the own rows with its source list in device memory (built by
// k_mirror_args just before).  (Measured and dropped: the flag wait + list build folded into this kernel, every block
// building the list in shared memory -- one launch less, but 183 us against 177 us per step at 2 ranks.)
template <int NB, bool MASKS>
__global__ void __launch_bounds__(256, (NB > 3) ? 2 : 3)
k_merge_rows_ind(const MergeArgs* __restrict__ A, MergeOut O, DevParams P) {
    pdl_wait();
    merge_rows_body<NB, MERGE_ROWS, MASKS>(*A, O, P);
}

// ---------------------------------------------------------------------------
// C2 and C3 of the mirrored combine run SIDE BY SIDE on two streams after the row merge:
//   k_merge_cells2_rows (main stream)  the cells of this rank's rows (source list in device memory, see k_merge_rows_ind);
//                                      publishes the rank's cell count
//   k_rows_heights      (side stream)  the heights of this rank's columns, pushed into every rank's 2-D block ([y][x]),
//                                      followed on the same stream by k_rows_known (flag exchange + bit maps).
// The height of a column needs only the MIN HEIGHT of its lowest occupied cell (__make_height_map, gvom.py:560-580):
// the minimum over the sources that are occupied at that one voxel -- a few look-ups per column instead of waiting for
// the whole cell merge (20-36 us), so the first 2-D exchange overlaps the cell kernel.  Inferred heights from the lowest
// free voxel (gvom.py:582-590); columns without an occupied voxel get the ego-disc height or -1000.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 8)
k_merge_cells2_rows(const MergeArgs* __restrict__ A, const int* __restrict__ counter, const int* __restrict__ cell_voxel,
                    int* __restrict__ chit, int* __restrict__ ctot, float* __restrict__ cminh,
                    float* __restrict__ cmet, float* __restrict__ ceig, DevParams P, int cap,
                    int* __restrict__ map_count, int* __restrict__ host_count, const unsigned* __restrict__ srcmask) {
    pdl_wait();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const int n = *counter;
        *map_count = n;
        if (host_count) *host_count = n;
    }
    merge_cells2_body(*A, counter, cell_voxel, chit, ctot, cminh, cmet, ceig, P, cap, NoCellHook{}, srcmask);
}

__global__ void __launch_bounds__(256)
k_rows_heights(const MergeArgs* __restrict__ Ap, const int* __restrict__ cmap, const int* __restrict__ col_occ,
               const int* __restrict__ col_free, double o0, double o1, double o2, double e0, double e1, double e2,
               DevParams P, RowShard R, const __grid_constant__ PushSet D, const unsigned* __restrict__ srcmask) {
    pdl_wait();
    const MergeArgs& A = *Ap;
    const int S = P.S, Z = P.Z;
    const long long S2 = (long long)S * S;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S * R.nrows) return;
    const int x = t % S, y = R.y0 + R.n * (t / S);        // x fastest: the column minima and the exchange block are [y][x]
    const long long ci = (long long)y * S + x;
    const int zo = col_occ[ci], zf = col_free[ci];
    const double inf = zf < Z ? __dmul_rn(__dadd_rn(o2, (double)zf), P.z_res) : -1000.0;
    double h = -1000.0;
    if (zo < Z) {
        // min height of the column's lowest occupied cell: the minimum over the sources occupied at that voxel -- the
        // same min the cell kernel takes for this cell (order independent, exact)
        unsigned want = 0xffffffffu;
        if (srcmask) {
            const int id = cmap[x + (long long)y * S + zo * S2];
            want = id >= 0 ? srcmask[id] : 0u;
        }
        float mh = 1.0f;
        constexpr int RB = 8;
        for (int k0 = 0; k0 < A.n; k0 += RB) {
            int io[RB];
#pragma unroll
            for (int u = 0; u < RB; ++u) {
                io[u] = -1;
                if (k0 + u < A.n && (!srcmask || A.s[k0 + u].is_prev || ((want >> ((k0 + u) & 31)) & 1u))) {
                    const SlotRef& s = A.s[k0 + u];
                    const int xs = x + s.dx, ys = y + s.dy, zs = zo + s.dz;
                    if ((unsigned)xs < (unsigned)S && (unsigned)ys < (unsigned)S && (unsigned)zs < (unsigned)Z)
                        io[u] = __ldg(s.map + (xs + (ys + zs * S) * S));
                }
            }
#pragma unroll
            for (int u = 0; u < RB; ++u)
                if (io[u] >= 0) mh = fminf(mh, A.s[k0 + u].minh[io[u]]);
        }
        h = __dmul_rn(__dadd_rn(__dadd_rn((double)zo, (double)mh), o2), P.z_res);
    } else {
        const double xp = __fma_rn(__dadd_rn(o0, (double)x), P.xy_res, -e0);
        const double yp = __fma_rn(__dadd_rn(o1, (double)y), P.xy_res, -e1);
        if (__fma_rn(xp, xp, __dmul_rn(yp, yp)) <= P.r2) h = __dsub_rn(e2, P.ground_to_lidar);
    }
    for (int d = 0; d < D.n; ++d) {                        // own copy included
        double* m = reinterpret_cast<double*>(D.base[d] + D.off_maps);
        m[ci] = h;
        m[S2 + ci] = inf;
    }
}

}  // namespace gvom
