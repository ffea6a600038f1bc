// gvom_merge.cuh -- C1 of combine_maps, third build: the row merge as a bulk-copy (TMA engine) pipeline.
//
// Same results as k_merge_codes<8, MERGE_FULL> (gvom_kernels.cuh): the merged code of every voxel from the ring
// slots and the previous combined map (__combine_indices per slot + __combine_old_indices, gvom.py:1009-1063,
// 242-257), compact ids for the occupied voxels, the column minima of __make_height_map /
// __make_inferred_height_map (gvom.py:560-590) and the group mask of the new map.
//
// The second build (k_merge_rows, round 1) kept every source row segment in REGISTERS (8 codes x up to 3 sources per
// lane, 80 registers, 24 warps per SM) and issued 8 bounds-checked 32-bit loads per source and lane: 816 warp
// instructions per 256-voxel segment at 51 % issue utilisation -- an instruction-issue and latency kernel (ncu,
// profiles/r01_v2_ncu_summary.md).  Here the loads leave the register file and the per-code work shrinks:
//   * one warp per block, each running its own 3-stage pipeline in 20 KB of shared memory (10 warps per SM).  The
//     warp's lane 0 asks the copy engine for the 1 KB row segments of the sources that hold anything in a segment
//     (`cp.async.bulk.shared.global`, completion counted in bytes on an mbarrier) up to three segments ahead.
//     The source's x-shift only moves the 16-byte aligned window that is copied; the shift itself is applied when
//     the lanes read shared memory, and the (few) codes that fall outside the source row are pre-set to "unknown"
//     by the warp itself, so the fold has no bounds checks
//   * lanes own voxels l, l+32, ... of the segment (stride 32): shared-memory reads are conflict free whatever
//     the shift, global stores are full 128-byte lines
//   * the fold costs 3 instructions per code: LDS, AND (a voxel is occupied iff any code has its sign bit clear)
//     and ADD (where no source is occupied every code is -1 - passes, so sum(passes) = -sum(codes) - #sources)
//   * a warp walks one (y, x-segment) strip through MR_ZT interleaved z levels: the group-mask words of all its
//     segments are fetched with one round trip (one item ahead), and the column minima live in registers and cost
//     one atomicMin per column and strip instead of a load + compare per voxel
//   * ids of occupied voxels come from a shuffle scan over per-lane counts (occupied voxels are ~2 % of the known
//     ones: most segments skip this entirely), the output group mask from an 8-lane OR + bit spread
#pragma once
#include "gvom_kernels.cuh"

namespace gvom {

constexpr int MR_KS = 6;        // source rows per pipeline stage
constexpr int MR_D = 3;         // stages per warp
constexpr int MR_ROW = 264;     // ints per staged row: 256 + alignment slack (multiple of 4)
constexpr int MR_ZT = 4;        // z levels a warp walks per (y, x-segment) strip
constexpr int MR_META = 4 + 2 * MR_KS + 4;   // ints of per-stage bookkeeping
constexpr int MR_MAX_SRC = 32;  // sources this kernel takes (ring slots + previous map); more: generic kernel

template <int NW>
struct __align__(16) MrWarp {
    int row[MR_D][MR_KS][MR_ROW];
    unsigned wq[2][MR_ZT][NW][32];   // group-mask words of the current / next item, as the lanes fetched them
    unsigned oldq[2][32];            // destination's own mask words (lane t: segment t of the item)
    int meta[MR_D][MR_META];         // [0] rows, [1] flags, [2] segment, [3] old mask word, [4+2r] source, [5+2r] smem offset, then y, z, x0s
    unsigned long long bar[MR_D];
};
enum { MR_FIRST = 1, MR_LAST = 2, MR_ITEM_LAST = 4 };

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    const long long t0 = clock64();
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        // a copy that never completes would be a bug in the byte accounting: give up after ~1 s instead of hanging
        // the GPU (the parity tests then fail on the result)
    } while (!ok && clock64() - t0 < 2000000000LL);
}

template <int NW>                                        // mask-word loads per lane and segment: sources <= 16 NW
__global__ void __launch_bounds__(32, 10)
k_merge_rows_async(MergeArgs A, MergeOut O, DevParams P) {
    extern __shared__ __align__(16) unsigned char mr_smem[];
    MrWarp<NW>& W = *reinterpret_cast<MrWarp<NW>*>(mr_smem);
    const int lane = threadIdx.x;
    const int S = P.S, Z = P.Z;
    const int spr = S >> 8;                               // segments per row
    const int ZC = (Z + MR_ZT - 1) / MR_ZT;               // z chunks: a strip walks z = zc, zc + ZC, zc + 2 ZC, ...
    const int per_zc = S * spr;
    const int nitems = per_zc * ZC;
    const int nwarps = gridDim.x;
    const int has_prev = (A.n > 0 && A.s[A.n - 1].is_prev) ? 1 : 0;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < MR_D; ++s) mbar_init(smem_u32(&W.bar[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    pdl_wait();

    // ---------------- mask words, one item ahead ----------------
    // lanes 0-15 hold word 0, lanes 16-31 word 1 of source 16 w + (lane & 15); lane t also holds the destination's word
    unsigned pre[MR_ZT][NW], pre_old = 0;
    auto fetch_item = [&](int item) {                     // issue the loads (results are stored to shared memory later)
        pre_old = 0;
#pragma unroll
        for (int t = 0; t < MR_ZT; ++t)
#pragma unroll
            for (int w = 0; w < NW; ++w) pre[t][w] = 0;
        if (item >= nitems) return;
        const int zc = item / per_zc, rem = item - zc * per_zc;
        const int y = rem / spr, xsg = rem - y * spr;
        const int which = lane >> 4;
#pragma unroll
        for (int t = 0; t < MR_ZT; ++t) {
            const int z = zc + t * ZC;
            if (z >= Z) break;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const int k = 16 * w + (lane & 15);
                if (k < A.n) {
                    const SlotRef& s = A.s[k];
                    const int ys = y + s.dy, zs = z + s.dz;
                    const int wi = (((xsg << 8) + s.dx) >> 8) + which;        // floor: source word of the segment start, +1
                    if ((unsigned)ys < (unsigned)S && (unsigned)zs < (unsigned)Z && wi >= 0 && wi < spr)
                        pre[t][w] = __ldg(s.gmask + (zs * S + ys) * spr + wi);
                }
            }
            if (lane == t) pre_old = __ldcg(O.gmask + (z * S + y) * spr + xsg);
        }
    };
    auto stash_item = [&](int buf) {                      // registers -> shared memory (the loads have landed long ago)
#pragma unroll
        for (int t = 0; t < MR_ZT; ++t)
#pragma unroll
            for (int w = 0; w < NW; ++w) W.wq[buf][t][w][lane] = pre[t][w];
        W.oldq[buf][lane] = pre_old;
    };

    // ---------------- producer side (state is warp-uniform; lane 0 talks to the copy engine) ----------------
    int p_item = blockIdx.x, p_t = 0, p_buf = 0;          // segment (p_item, p_t) is the next to be started
    bool seg_loaded = false, p_first = true;
    unsigned p_act = 0;                                   // sources of the current segment not yet put into a unit
    int c_seg = 0, c_y = 0, c_z = 0, c_x0 = 0, c_flags = 0;
    unsigned c_old = 0;

    auto load_segment = [&]() -> bool {                   // start segment (p_item, p_t): which sources hold anything in it?
        if (p_item >= nitems) return false;
        const int zc = p_item / per_zc, rem = p_item - zc * per_zc;
        const int y = rem / spr, xsg = rem - y * spr, z = zc + p_t * ZC;
        c_y = y; c_z = z; c_x0 = xsg << 8;
        c_seg = (z * S + y) * spr + xsg;
        const bool item_last = (p_t + 1 >= MR_ZT) || (z + ZC >= Z);
        c_flags = item_last ? MR_ITEM_LAST : 0;
        c_old = W.oldq[p_buf][p_t];
        unsigned act = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const unsigned word = W.wq[p_buf][p_t][w][lane];
            const unsigned other = __shfl_xor_sync(FULL, word, 16);
            const int k = 16 * w + (lane & 15);
            bool a = false;
            if (lane < 16 && k < A.n) {
                const int sx0 = c_x0 + A.s[k].dx;
                const unsigned long long Wd = ((unsigned long long)other << 32) | word;
                const unsigned long long win = (sx0 & 7) ? 0x1ffffffffULL : 0xffffffffULL;   // an unaligned shift straddles one more group
                a = ((Wd >> ((sx0 >> 3) & 31)) & win) != 0ULL;
            }
            act |= (__ballot_sync(FULL, a) & 0xffffu) << (16 * w);
        }
        p_act = act; p_first = true; seg_loaded = true;
        if (item_last) {                                  // next item: its words are in registers since the last switch
            p_item += nwarps; p_t = 0; p_buf ^= 1;
            __syncwarp();
            stash_item(p_buf);
            fetch_item(p_item + nwarps);
            __syncwarp();
        } else {
            ++p_t;
        }
        return true;
    };
    auto produce = [&](int s) -> bool {                   // put one unit (<= MR_KS source rows of one segment) into stage s
        if (!seg_loaded && !load_segment()) return false;
        const int nact = __popc(p_act);
        const int nrows = min(nact, MR_KS);
        int flags = p_first ? MR_FIRST : 0;
        if (nact <= MR_KS) flags |= MR_LAST | (c_flags & MR_ITEM_LAST);
        int* m = W.meta[s];
        const unsigned bar = smem_u32(&W.bar[s]);
        // pass 1: bytes of the whole unit (the barrier is armed before the first copy is issued)
        unsigned total = 0;
        {
            unsigned rem = p_act;
#pragma unroll 1
            for (int r = 0; r < nrows; ++r) {
                const int k = __ffs(rem) - 1;
                rem &= rem - 1;
                const int sx0 = c_x0 + A.s[k].dx;
                total += (unsigned)(((min(sx0 + 256, S) + 3) & ~3) - (max(sx0, 0) & ~3)) * 4u;
            }
        }
        if (lane == 0 && nrows) mbar_expect_tx(bar, total);
#pragma unroll 1
        for (int r = 0; r < nrows; ++r) {
            const int k = __ffs(p_act) - 1;
            p_act &= p_act - 1;
            const SlotRef& sr = A.s[k];
            const int sx0 = c_x0 + sr.dx;
            const int lo0 = sx0 & ~3;                     // staged index i <-> source x = lo0 + i: destination voxel v sits at c0 + v
            const int c0 = sx0 - lo0;
            const int lo = max(lo0, 0), hi = (min(sx0 + 256, S) + 3) & ~3;      // 16-byte aligned window inside the source row
            int* rowp = &W.row[s][r][0];
            // destination voxels whose source x falls outside the row read as "unknown"
            const int lead = min(max(-sx0, 0), 256), trail = min(max(S - sx0, 0), 256);
            for (int i = lane; i < lead; i += 32) rowp[c0 + i] = -1;
            for (int i = trail + lane; i < 256; i += 32) rowp[c0 + i] = -1;
            if (lane == 0) {
                m[4 + 2 * r] = k; m[5 + 2 * r] = c0;
                bulk_g2s(smem_u32(rowp + (lo - lo0)), sr.map + ((c_z + sr.dz) * S + (c_y + sr.dy)) * S + lo,
                         (unsigned)(hi - lo) * 4u, bar);
            }
        }
        if (lane == 0) {
            m[0] = nrows; m[1] = flags; m[2] = c_seg; m[3] = (int)c_old;
            m[4 + 2 * MR_KS] = c_y; m[5 + 2 * MR_KS] = c_z; m[6 + 2 * MR_KS] = c_x0;
        }
        p_first = false;
        if (p_act == 0) seg_loaded = false;
        return true;
    };

    // ---------------- consumer side ----------------
    int acc_and[8], sum[8], op[8], cmo[8], cmf[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc_and[j] = -1; sum[j] = 0; op[j] = -1; cmo[j] = 0x7fffffff; cmf[j] = 0x7fffffff; }
    int nf = 0;
    unsigned phase = 0;                                   // bit s: parity of stage s's next completion

    fetch_item(p_item);
    stash_item(0);
    fetch_item(p_item + nwarps);
    __syncwarp();
    int produced = 0, consumed = 0, s = 0;
#pragma unroll 1
    for (int q = 0; q < MR_D; ++q) { if (produce(q)) ++produced; else break; }
    __syncwarp();
#pragma unroll 1
    while (consumed < produced) {
        const int* m = W.meta[s];
        const int nrows = m[0], flags = m[1], seg = m[2];
        const unsigned old_word = (unsigned)m[3];
        if (nrows) {
            mbar_wait(smem_u32(&W.bar[s]), (phase >> s) & 1u);
            phase ^= 1u << s;
        }
        if (flags & MR_FIRST) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { acc_and[j] = -1; sum[j] = 0; op[j] = -1; }
            nf = 0;
        }
#pragma unroll 1
        for (int r = 0; r < nrows; ++r) {
            const int k = m[4 + 2 * r];
            const int* src = &W.row[s][r][0] + m[5 + 2 * r] + lane;
            if (has_prev && k == A.n - 1) {
#pragma unroll
                for (int j = 0; j < 8; ++j) op[j] = src[32 * j];
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) { const int v = src[32 * j]; acc_and[j] &= v; sum[j] += v; }
                ++nf;
            }
        }
        if (flags & MR_LAST) {
            const int y = m[4 + 2 * MR_KS], z = m[5 + 2 * MR_KS], x0s = m[6 + 2 * MR_KS];
            if (nrows == 0 && (flags & MR_FIRST)) {
                // nothing known anywhere in the segment (half of the grid): write only if the buffer holds something
                if (old_word != 0u) {
                    int* dst = O.cmap + (long long)seg * 256 + lane;
#pragma unroll
                    for (int j = 0; j < 8; ++j) dst[32 * j] = -1;
                    if (lane == 0) O.gmask[seg] = 0u;
                }
            } else {
                int c[8];
                unsigned occm = 0, freem = 0;
                const int nfm1 = nf - 1;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    bool occ = acc_and[j] >= 0;
                    c[j] = sum[j] + nfm1;                 // -1 - sum of passes (meaningful where no source is occupied)
                    if (!occ) {                           // previous combined map (gvom.py:1058-1063)
                        if (op[j] >= 0) { if (c[j] >= -11) occ = true; }
                        else if (op[j] < -1) c[j] += op[j] + 1;
                    }
                    if (occ) occm |= 1u << j; else if (c[j] < -1) freem |= 1u << j;
                }
                if (__any_sync(FULL, occm != 0u)) {       // compact ids: shuffle scan over per-lane counts, one atomic per segment
                    const int cnt = __popc(occm);
                    int incl = cnt;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const int t = __shfl_up_sync(FULL, incl, off);
                        if (lane >= off) incl += t;
                    }
                    const int total = __shfl_sync(FULL, incl, 31);
                    int base = 0;
                    if (lane == 0) base = atomicAdd(O.counter, total);
                    base = __shfl_sync(FULL, base, 0);
                    int id = base + incl - cnt;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (occm & (1u << j)) {
                            if (id < O.cap) { c[j] = id; O.cell_voxel[id] = seg * 256 + lane + 32 * j; }
                            else { c[j] = -1; occm &= ~(1u << j); }
                            ++id;
                        } else if (!(freem & (1u << j))) {
                            c[j] = -1;
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) if (!(freem & (1u << j))) c[j] = -1;
                }
                unsigned knownm = occm | freem;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (occm & (1u << j)) cmo[j] = min(cmo[j], z);
                    if (freem & (1u << j)) cmf[j] = min(cmf[j], z);
                }
                // group mask word: group g = voxels 8g .. 8g+7 = lanes 8 (g % 4) .. +7 at step j = g / 4
                unsigned km = knownm;
                km |= __shfl_xor_sync(FULL, km, 1); km |= __shfl_xor_sync(FULL, km, 2); km |= __shfl_xor_sync(FULL, km, 4);
                unsigned sp = km;                         // bit j -> bit 4 j
                sp = (sp | (sp << 12)) & 0x000f000fu;
                sp = (sp | (sp << 6)) & 0x03030303u;
                sp = (sp | (sp << 3)) & 0x11111111u;
                unsigned w = sp << (lane >> 3);
                w |= __shfl_xor_sync(FULL, w, 8); w |= __shfl_xor_sync(FULL, w, 16);
                if (w != 0u || old_word != 0u) {          // uniform
                    int* dst = O.cmap + (long long)seg * 256 + lane;
#pragma unroll
                    for (int j = 0; j < 8; ++j) dst[32 * j] = c[j];
                    if (lane == 0) O.gmask[seg] = w;
                }
            }
            if (flags & MR_ITEM_LAST) {                   // the strip is done: publish its column minima
                int* colo = O.col_occ + y * S + x0s + lane;
                int* colf = O.col_free + y * S + x0s + lane;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (cmo[j] != 0x7fffffff) atomicMin(colo + 32 * j, cmo[j]);
                    if (cmf[j] != 0x7fffffff) atomicMin(colf + 32 * j, cmf[j]);
                    cmo[j] = 0x7fffffff; cmf[j] = 0x7fffffff;
                }
            }
        }
        ++consumed;
        __syncwarp();                                     // every lane is done with stage s before it is refilled
        if (produce(s)) ++produced;
        __syncwarp();                                     // lane 0's bookkeeping and the "unknown" pre-sets are visible to the warp
        s = (s + 1 == MR_D) ? 0 : s + 1;
    }
}

}  // namespace gvom
