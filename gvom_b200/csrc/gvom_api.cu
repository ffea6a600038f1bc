// gvom_api.cu -- C-ABI (include/gvom_b200.h) and host orchestration of the
// G-VOM voxel-mapping path on B200.  Replaces the host side of the reference
// class (scripts/gvom.py:21-442, 1069-1119): ring buffer of per-scan maps,
// launch sequencing, state carried between combine_maps() calls.
//
// Differences from the reference's host code that matter for speed, not results:
//   * no allocation after gvom_create(): every slot / grid lives in one caller
//     provided workspace (the reference cudaMallocs >= 7 arrays per scan)
//   * no host round trip inside Process_pointcloud (gvom.py:172 blocks on the
//     cell count); counts stay on the device
//   * 4 launches per scan and 4 per combine instead of ~15 and ~3B+30
//   * one stream per handle orders slot overwrites against combine reads, which
//     the reference gets from allocating fresh arrays under Python semaphores.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/gvom_b200.h"
#include "gvom_kernels.cuh"
#include "gvom_scan.cuh"
#include "gvom_merge.cuh"
#include "gvom_mirror.cuh"

using namespace gvom;

static thread_local std::string g_err;
const char* gvom_last_error(void) { return g_err.c_str(); }

static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return fail(GVOM_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));        \
    } while (0)

namespace {

struct Slot {
    int* index_map = nullptr;     // [V]
    int* hit = nullptr;           // [cap]
    int* total = nullptr;         // [cap]
    double* metrics = nullptr;    // [cap,10]
    float* minh = nullptr;        // [cap]
    int* cell_voxel = nullptr;    // [cap]
    int* counter = nullptr;       // device cell count
    unsigned* gmask = nullptr;    // [V/256] one bit per 8-voxel group: something known (valid when has_gmask)
    bool has_gmask = false;
    bool dirty = false;           // the map holds something else than "all unknown" (a physical slot is wiped before reuse)
    double origin[3] = {0, 0, 0};
    bool valid = false;
};

struct Combined {
    int* index_map = nullptr;     // [V]
    int* hit = nullptr;           // [ccap]
    int* total = nullptr;
    float* minh = nullptr;
    float* metrics = nullptr;     // [ccap,10]
    float* eig = nullptr;         // [ccap,3]
    int* cell_voxel = nullptr;
    int* counter = nullptr;
    unsigned* gmask = nullptr;
    bool has_gmask = false;
    double origin[3] = {0, 0, 0};
    bool valid = false;
    int64_t cells = 0;
};

// Pageable host clouds have to pass through pinned memory; a single-threaded memcpy of a 6.3 MB scan costs
// more than everything the GPU does with it, so the staging copy is split over a few persistent threads.
}  // namespace
#include "gvom_host.h"
namespace {

struct Carver {                    // sub-allocates a workspace block, 256-byte aligned
    char* base;
    size_t off = 0;
    explicit Carver(void* b) : base(static_cast<char*>(b)) {}
    template <typename T>
    T* take(size_t count) {
        off = (off + 255) & ~size_t(255);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

enum { EV_START = 0, EV_H2D, EV_POINTS, EV_SCELLS, EV_CSTART, EV_CODES, EV_CELLS, EV_MAPS, EV_D2H,
       EV_X0, EV_X1, EV_X2, EV_P0, EV_P1, EV_COUNT };   // EV_X*: extra marks inside the multi-GPU combine

// GVOM_VARIANT bits (environment / gvom_set_variant): A/B switches for measurements; every setting gives the same results
enum { VAR_GENERIC_MERGE = 2, VAR_ASYNC_ROWS = 4, VAR_NO_SRCMASK = 16, VAR_BULK_PUSH = 32, VAR_DMA_OUT = 64, VAR_NO_FASTFLOOR = 128, VAR_NO_BULK_H2D = 256 };

}  // namespace

struct GvomHandle {
    GvomParams p;
    DevParams dp;
    int device = 0;
    int64_t max_points = 0, cap = 0, gcap = 0, ccap = 0, V = 0, EV = 0;
    int S2 = 0, ES = 0, EZ = 0;
    // device
    unsigned* cellid = nullptr;           // [EV] extended, epoch-tagged voxel -> cell grid of the scan kernels
    unsigned scan_tag = 0;                // tag of the last scan (1..255; wraps through a clear of cellid)
    double* acc[2] = {nullptr, nullptr};  // [cap + gcap][MOM] raw moments of the scan in flight, double-buffered: S2 clears behind S1
    char* stage_dev = nullptr;            // input cloud staging [max_points * 32 B]
    std::vector<Slot> slots;              // B + 1 PHYSICAL slots: B ring entries + one spare that is kept wiped
    std::vector<int> phys;                // ring index (the reference's buffer index) -> physical slot
    int spare = 0;                        // physical slot the next scan is written to
    Combined comb[2];
    int cur = 0;                          // comb[cur] = last combined map (if valid)
    double* maps = nullptr;               // height, inferred, rough_work, xs, ys, guessed  [6][S*S]
    int* imaps = nullptr;                 // result block: pos, neg, vis int32 [3][S*S] then roughness f64 [S*S]
    double* rough_out = nullptr;          // = (double*)(imaps + 3*S*S)
    // where the 2-D maps of the last combine live: the library's own block above, or (row-sharded multi-GPU combine) the
    // exchange block the ranks pushed into -- debug exports, OccupancyGrid post-processing and state save read these
    double* v_maps = nullptr; int* v_imaps = nullptr; double* v_rough = nullptr;
    // row-sharded combine: the six float64 work maps (heights, slopes, guessed heights) stay in the exchange block
    // ([y][x]) and are only transposed into `maps` when a debug export / state save asks for them
    struct Maps6Pending { bool active = false; const char* blk = nullptr; PushSet D{}; } maps6;
    int* col_minz = nullptr;              // [2][S*S] lowest occupied / lowest free z per column (C1 -> C3)
    unsigned* known = nullptr;            // [2][S*ceil(S/32)] "height known" bit maps (rows over y, rows over x)
    float* debug_dev = nullptr;           // [max(ccap*8, S*S*10)]
    int* flags = nullptr;                 // [0..5] cell + ghost counters of scans k % 3, [8] C1's running cell counter, [9..11] grid-done counters
    // multi-GPU scratch
    double* cacc = nullptr;               // [ccap,10] raw-moment scratch of the multi-GPU combine
    // pinned host
    char* stage_host = nullptr;           // [max_points * 32 B]
    int* out_i_host = nullptr;            // [3*S*S]
    int* counters_host = nullptr;         // [8]
    // state
    int buffer_index = 0, last_buffer_index = 0;
    double ego[3] = {0, 0, 0};
    bool have_maps = false;
    cudaStream_t stream = nullptr;        // the handle's own stream
    cudaStream_t active = nullptr;        // stream of the last process / combine call (tooling syncs it)
    cudaEvent_t ev_stage = nullptr;       // completion of the last H2D that read stage_host
    bool stage_busy = false;
    cudaStream_t copy_stream = nullptr;   // H2D of the cloud, chunked so that ray casting overlaps the transfer
    cudaEvent_t ev_chunk[8];              // chunk c has landed in stage_dev
    cudaEvent_t ev_proc_done = nullptr;   // last reader of stage_dev (the previous scan) is done
    cudaEvent_t ev_input = nullptr;       // the scan kernels have consumed the caller's device cloud
    cudaEvent_t ev[EV_COUNT];
    bool profiling = false;
    bool zero_copy = true;                // host clouds: S1 reads pinned memory directly (else chunked DMA)
    bool prof_process = false, prof_combine = false, prof_partial = false, prof_rows = false;
    int sm_count = 148;
    int grid_codes = 0, grid_cells = 0, grid_cells2 = 0, grid_rows3 = 0, grid_scan_cells = 0, grid_rows_mirror = 0, grid_rows_async[2] = {0, 0};   // resident grids (set at create)
    GvomStats stats{};
    float last_stage_copy_ms = 0.f;       // host time of the last pageable->pinned staging copy
    CopyPool* pool = nullptr;             // staging threads for pageable input (created on first use)
    signed char* grids_dev = nullptr;     // [GVOM_GRID_COUNT][S*S] int8 OccupancyGrid payloads
    signed char* grids_host = nullptr;    // pinned mirror
    int host_chunks = 4;                  // pieces a large pageable / PointCloud2 host cloud is staged in (GVOM_CHUNKS), pipelined with S1
    unsigned variant = 0;                 // GVOM_VARIANT bit mask (A/B switches, see VAR_*)
    // outputs of the last combine that still have to be completed on the host (gvom_combine_maps_async)
    struct Pending {
        bool active = false;
        Combined* c = nullptr;
        cudaStream_t st = nullptr;
        bool from_mirror = false;         // pageable outputs: copy out of the pinned mirror after the stream drained
        int32_t *positive = nullptr, *negative = nullptr, *visibility = nullptr;
        double* roughness = nullptr;
    } pend;
    // mirrored multi-GPU combine (gvom_mirror_attach): every scan is pushed to the owners of its rows
    struct Mirror {
        int n = 0, self = 0;
        char* base[MAX_RANKS] = {};
        size_t o_flags = 0, o_table = 0, o_args = 0, o_held = 0, o_mirrors = 0, mirror_bytes = 0;
        size_t f_gmask = 0, f_hit = 0, f_tot = 0, f_minh = 0, f_met = 0, nsegp = 0, total = 0;
    } mir;
    std::mutex mu;
    Slot& ring(int i) { return slots[phys[i]]; }
};

namespace {

size_t carve(GvomHandle* h, void* dev, void* host, size_t* host_bytes) {
    const GvomParams& p = h->p;
    const size_t V = (size_t)h->V, cap = (size_t)h->cap, ccap = (size_t)h->ccap, S2 = (size_t)h->S2;
    Carver d(dev);
    h->cellid = d.take<unsigned>((size_t)h->EV);
    for (int q = 0; q < 2; ++q) h->acc[q] = d.take<double>((cap + (size_t)h->gcap) * MOM);
    h->stage_dev = d.take<char>((size_t)h->max_points * 32);
    h->slots.resize(p.buffer_size + 1);
    h->phys.resize(p.buffer_size);
    for (int i = 0; i < p.buffer_size; ++i) h->phys[i] = i;
    h->spare = p.buffer_size;
    for (auto& s : h->slots) {
        s.index_map = d.take<int>(V);
        s.hit = d.take<int>(cap);
        s.total = d.take<int>(cap);
        s.metrics = d.take<double>(cap * 10);
        s.minh = d.take<float>(cap);
        s.cell_voxel = d.take<int>(cap);
        s.counter = d.take<int>(4);
        s.gmask = d.take<unsigned>(V / 256 + 2);
    }
    for (auto& c : h->comb) {
        c.index_map = d.take<int>(V);
        c.hit = d.take<int>(ccap);
        c.total = d.take<int>(ccap);
        c.minh = d.take<float>(ccap);
        c.metrics = d.take<float>(ccap * 10);
        c.eig = d.take<float>(ccap * 3);
        c.cell_voxel = d.take<int>(ccap);
        c.counter = d.take<int>(4);
        c.gmask = d.take<unsigned>(V / 256 + 2);
    }
    h->maps = d.take<double>(6 * S2);
    // result block: three int32 maps, padding to 8 bytes (odd xy_size), then the float64 roughness map
    h->imaps = d.take<int>(3 * S2 + 2 * S2 + 2);
    h->rough_out = reinterpret_cast<double*>(h->imaps ? h->imaps + ((3 * S2 + 1) & ~size_t(1)) : nullptr);
    h->v_maps = h->maps; h->v_imaps = h->imaps; h->v_rough = h->rough_out;
    h->col_minz = d.take<int>(2 * S2);
    h->known = d.take<unsigned>(2 * (size_t)p.xy_size * ((p.xy_size + 31) / 32));
    h->debug_dev = d.take<float>(std::max(ccap * 8, S2 * 10));
    h->flags = d.take<int>(16);
    h->cacc = d.take<double>(ccap * 10);
    h->grids_dev = d.take<signed char>(GVOM_GRID_COUNT * S2);
    Carver c(host);
    h->stage_host = c.take<char>((size_t)h->max_points * 32);
    h->out_i_host = c.take<int>(3 * S2 + 2 * S2 + 2);   // mirrors the device result block
    h->counters_host = c.take<int>(8);
    h->grids_host = c.take<signed char>(GVOM_GRID_COUNT * S2);
    *host_bytes = c.off + 256;
    return d.off + 256;
}

int check_params(const GvomParams* p, int64_t max_points) {
    if (!p) return fail(GVOM_EINVAL, "params is NULL");
    if (!(p->xy_resolution > 0) || !(p->z_resolution > 0)) return fail(GVOM_EINVAL, "resolutions must be > 0");
    if (p->xy_size < 1 || p->z_size < 1) return fail(GVOM_EINVAL, "grid sizes must be >= 1");
    if (p->buffer_size < 1 || p->buffer_size > MAX_SLOTS) return fail(GVOM_EINVAL, "buffer_size must be in [1,64]");
    if (p->xy_eigen_dist < 0 || p->z_eigen_dist < 0) return fail(GVOM_EINVAL, "eigen distances must be >= 0");
    const double V = (double)p->xy_size * p->xy_size * p->z_size;
    const double EVd = ((double)p->xy_size + 2.0 * p->xy_eigen_dist) * ((double)p->xy_size + 2.0 * p->xy_eigen_dist) *
                       ((double)p->z_size + 2.0 * p->z_eigen_dist);
    if (V >= 2147483647.0 || EVd >= 2147483647.0) return fail(GVOM_EINVAL, "grid (with its eigen-distance margin) has >= 2^31 voxels");
    if (max_points < 1 || max_points > (int64_t)CELL_OVERFLOW - 1) return fail(GVOM_EINVAL, "max_points out of range (1 .. 2^23 - 3)");
    return GVOM_OK;
}

void fill_sizes(GvomHandle* h, const GvomParams* p, int64_t max_points, int64_t max_cells) {
    h->p = *p;
    h->V = (int64_t)p->xy_size * p->xy_size * p->z_size;
    h->S2 = p->xy_size * p->xy_size;
    h->max_points = max_points;
    h->cap = std::min<int64_t>(max_points, h->V);
    h->ES = p->xy_size + 2 * p->xy_eigen_dist;
    h->EZ = p->z_size + 2 * p->z_eigen_dist;
    h->EV = (int64_t)h->ES * h->ES * h->EZ;
    h->gcap = std::min<int64_t>(max_points, h->EV - h->V);      // margin ("ghost") cells
    int64_t dflt = std::min<int64_t>(h->V, 4 * max_points * ((int64_t)p->buffer_size + 1));
    h->ccap = max_cells > 0 ? std::min<int64_t>(max_cells, h->V) : dflt;
    DevParams& d = h->dp;
    d.xy_res = p->xy_resolution; d.z_res = p->z_resolution;
    d.min_d2 = p->min_distance * p->min_distance;
    d.pos_thr = p->positive_obstacle_threshold; d.neg_thr = p->negative_obstacle_threshold;
    d.slope_thr = p->slope_obsacle_threshold; d.robot_height = p->robot_height;
    d.r2 = p->robot_radius * p->robot_radius; d.ground_to_lidar = p->ground_to_lidar_height;
    d.S = p->xy_size; d.Z = p->z_size; d.rx = p->xy_eigen_dist; d.rz = p->z_eigen_dist;
    d.V = h->V;
    d.lgS = -1;
    for (int b = 0; b < 31; ++b) if ((1 << b) == p->xy_size) d.lgS = b;
}

inline int blocks_for(int64_t n, int threads) { return (int)((n + threads - 1) / threads); }

// Launch with programmatic stream serialization (PDL): the kernel may start while its predecessor in
// the stream drains; it calls pdl_wait() before touching the predecessor's results.
template <typename... KArgs, typename... Args>
void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

// grid of a grid-stride kernel: exactly the blocks that are resident at once (one wave, no tail)
template <typename K>
int resident_grid(K kernel, int threads, int sm_count, size_t smem = 0) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 2;
    }
    return per_sm * sm_count;
}

bool is_pinned_or_device(const void* p, bool* is_device) {
    cudaPointerAttributes a;
    *is_device = false;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) { *is_device = true; return true; }
    return a.type == cudaMemoryTypeHost;
}

void rec(GvomHandle* h, int e, cudaStream_t st) {
    if (h->profiling) cudaEventRecord(h->ev[e], st);
}

// Build the ordered source list of a combine: valid ring slots 0..B-1, then the
// previous combined map (gvom.py:242-257).  `org` = combined origin (voxels).
void build_sources(GvomHandle* h, const double org[3], bool with_prev, MergeArgs* A) {
    A->n = 0;
    A->use_masks = 1;
    for (int i = 0; i < h->p.buffer_size; ++i) {         // ring order = the reference's slot order
        Slot& s = h->ring(i);
        if (!s.valid) continue;
        SlotRef& r = A->s[A->n++];
        r.map = s.index_map; r.metrics = s.metrics; r.hit = s.hit; r.total = s.total; r.minh = s.minh;
        r.dx = (int)(org[0] - s.origin[0]); r.dy = (int)(org[1] - s.origin[1]); r.dz = (int)(org[2] - s.origin[2]);
        r.is_prev = 0;
        r.gmask = s.has_gmask ? s.gmask : nullptr;
        if (!r.gmask) A->use_masks = 0;
    }
    Combined& pc = h->comb[h->cur];
    if (with_prev && pc.valid) {
        SlotRef& r = A->s[A->n++];
        r.map = pc.index_map; r.metrics = pc.metrics; r.hit = pc.hit; r.total = pc.total; r.minh = pc.minh;
        r.dx = (int)(org[0] - pc.origin[0]); r.dy = (int)(org[1] - pc.origin[1]); r.dz = (int)(org[2] - pc.origin[2]);
        r.is_prev = 1;
        r.gmask = pc.has_gmask ? pc.gmask : nullptr;
        if (!r.gmask) A->use_masks = 0;
    }
}

template <int MODE>
void launch_merge(GvomHandle* h, const MergeArgs& A, const MergeOut& O, cudaStream_t st, int grid_div = 1) {
    // row-segment kernels: xy_size % 256 == 0, every source carries a group mask, the destination has one
    if (h->p.xy_size % 256 == 0 && A.use_masks && O.gmask && !(h->variant & VAR_GENERIC_MERGE) &&
        (MODE != MERGE_FINISH || O.cacc == nullptr)) {
        // the bulk-copy pipeline build of the row merge is opt-in: parity green, but measured slower (gvom_merge.cuh)
        if (MODE == MERGE_FULL && A.n <= 16 && h->grid_rows_async[0] > 0 && (h->variant & VAR_ASYNC_ROWS))
            launch(k_merge_rows_async<1>, dim3(h->grid_rows_async[0]), dim3(32), sizeof(MrWarp<1>), st, A, O, h->dp);
        else if (MODE == MERGE_FULL && A.n <= MR_MAX_SRC && h->grid_rows_async[1] > 0 && (h->variant & VAR_ASYNC_ROWS))
            launch(k_merge_rows_async<2>, dim3(h->grid_rows_async[1]), dim3(32), sizeof(MrWarp<2>), st, A, O, h->dp);
        else
            launch(k_merge_rows<3, MODE>, dim3(std::max(1, h->grid_rows3 / grid_div)), dim3(256), 0, st, A, O, h->dp);
        h->stats.kernel_launches++;
        return;
    }
    const dim3 g(std::max(1, h->grid_codes / grid_div));
    if (h->p.xy_size % 8 == 0)
        launch(k_merge_codes<8, MODE>, g, dim3(256), 0, st, A, O, h->dp);
    else if (h->p.xy_size % 4 == 0)
        launch(k_merge_codes<4, MODE>, g, dim3(256), 0, st, A, O, h->dp);
    else
        launch(k_merge_codes<1, MODE>, g, dim3(256), 0, st, A, O, h->dp);
    h->stats.kernel_launches++;
}

// C2: per-cell record merge + eigenvalues
void launch_cells(GvomHandle* h, const MergeArgs& A, Combined& c, cudaStream_t st) {
    launch(k_merge_cells2, dim3(h->grid_cells2), dim3(128), 0, st, A, h->flags + 8, c.cell_voxel, c.hit, c.total, c.minh, c.metrics, c.eig,
                                                h->dp, (int)h->ccap);
}

// Row-sharded combine: bring the six work maps of the last combine from the exchange block into the library's own
// [x][y] block (debug exports, state save).  No-op otherwise.
void ensure_maps6(GvomHandle* h) {
    if (!h->maps6.active) return;
    h->maps6.active = false;
    const int S = h->p.xy_size, W = (S + 31) / 32;
    const size_t S2 = (size_t)h->S2;
    MapSet own{h->maps, h->imaps, h->imaps + S2, h->imaps + 2 * S2, h->rough_out};
    MapSet user{nullptr, nullptr, nullptr, nullptr, nullptr};
    k_rows_deliver<<<dim3(W, W, 6), dim3(256), 0, h->active>>>(h->maps6.blk, h->maps6.D, S, own, user, (const int*)nullptr, 0, 0, 0, SignalSet{});
}

// Completes the outputs of the last combine on the host: waits for the stream, copies pageable outputs out of
// the pinned mirror, publishes the cell count.  No-op when nothing is pending.
int finish_outputs(GvomHandle* h) {
    GvomHandle::Pending& pd = h->pend;
    if (!pd.active) return GVOM_OK;
    pd.active = false;
    CUDA_TRY(cudaStreamSynchronize(pd.st));
    const size_t S2 = (size_t)h->S2;
    if (pd.from_mirror) {
        const size_t bi = S2 * sizeof(int), bd = S2 * sizeof(double);
        const size_t rough_off = ((3 * S2 + 1) & ~size_t(1)) * sizeof(int);
        if (pd.positive) memcpy(pd.positive, h->out_i_host, bi);
        if (pd.negative) memcpy(pd.negative, h->out_i_host + S2, bi);
        if (pd.visibility) memcpy(pd.visibility, h->out_i_host + 2 * S2, bi);
        if (pd.roughness) memcpy(pd.roughness, reinterpret_cast<char*>(h->out_i_host) + rough_off, bd);
    }
    if (h->counters_host[1]) {
        const int why = h->counters_host[1];
        h->counters_host[1] = 0;
        if (why == 2) return fail(GVOM_ECUDA, "multi-GPU combine: timed out waiting for a rank (every rank must call combine_maps)");
        return fail(GVOM_EINVAL, "multi-GPU combine: ranks disagree on the map origin (sensors must share the ego position)");
    }
    Combined& c = *pd.c;
    c.cells = std::min<int64_t>(h->counters_host[0], h->ccap);
    h->stats.combined_cells = c.cells;
    if (h->counters_host[0] > h->ccap)
        return fail(GVOM_ECAPACITY, "combined map has more occupied cells than max_combined_cells");
    return GVOM_OK;
}

// 2-D stage + outputs, shared by the single- and multi-GPU combine.  Everything is enqueued on `st`;
// finish_outputs() completes it (the synchronous entry points call it right away).
int enqueue_maps_and_output(GvomHandle* h, Combined& c, double origin[3], int32_t* positive, int32_t* negative,
                            double* roughness, int32_t* visibility, int32_t out_mem, cudaStream_t st) {
    const int S2 = h->S2;
    h->v_maps = h->maps; h->v_imaps = h->imaps; h->v_rough = h->rough_out;
    h->maps6.active = false;
    double* height = h->maps; double* inferred = h->maps + S2; double* rough = h->rough_out;
    double* xs = h->maps + 3 * (size_t)S2; double* ys = h->maps + 4 * (size_t)S2; double* guessed = h->maps + 5 * (size_t)S2;
    int* pos = h->imaps; int* neg = h->imaps + S2; int* vis = h->imaps + 2 * (size_t)S2;
    // device-resident outputs: the surface kernel writes them itself, next to the library's own result block
    // (which the debug exports and the OccupancyGrid post-processing read).  Pinned host outputs are written the
    // same way through their device mapping (posted PCIe writes, coalesced): no DMA operation on the critical path.
    bool direct_dev = out_mem == GVOM_DEVICE && positive && negative && roughness && visibility;
    int32_t *kpos = positive, *kneg = negative, *kvis = visibility;
    double* krough = roughness;
    if (!direct_dev && out_mem == GVOM_HOST && positive && negative && roughness && visibility && h->zero_copy &&
        !(h->variant & VAR_DMA_OUT)) {
        void* m[4] = {nullptr, nullptr, nullptr, nullptr};
        void* hp[4] = {positive, negative, visibility, roughness};
        bool ok = true, dev = false;
        for (int k = 0; k < 4 && ok; ++k) {
            ok = is_pinned_or_device(hp[k], &dev) && !dev && cudaHostGetDevicePointer(&m[k], hp[k], 0) == cudaSuccess && m[k];
            if (!ok) cudaGetLastError();
        }
        if (ok) {
            direct_dev = true;
            kpos = (int32_t*)m[0]; kneg = (int32_t*)m[1]; kvis = (int32_t*)m[2]; krough = (double*)m[3];
        }
    }
    // combined cell count: C3 stores it into a mapped pinned word (else a 4-byte DMA after the kernels)
    int* host_count = nullptr;
    if (h->zero_copy && !(h->variant & VAR_DMA_OUT)) {
        void* m = nullptr;
        if (cudaHostGetDevicePointer(&m, h->counters_host, 0) == cudaSuccess) host_count = (int*)m; else cudaGetLastError();
    }
    const int W = (h->p.xy_size + 31) / 32;
    unsigned* known = h->known; unsigned* knownT = h->known + (size_t)h->p.xy_size * W;
    launch(k_column_maps, dim3(W, W), dim3(1024), 0, st, c.index_map, c.minh, h->col_minz, h->col_minz + S2, c.origin[0], c.origin[1],
                                               c.origin[2], h->ego[0], h->ego[1], h->ego[2], h->dp, height, inferred, known, knownT,
                                               h->flags + 8, c.counter, host_count);
    const size_t mask_bytes = 2 * (size_t)h->p.xy_size * W * sizeof(unsigned);
    const int in_smem = (mask_bytes <= 40 * 1024 && (mask_bytes % 16) == 0) ? 1 : 0;
    {
        launch(k_surface_maps2, dim3(blocks_for(S2, 128)), dim3(256), in_smem ? mask_bytes : 0, st, c.index_map, c.hit, c.total, height, inferred, known, knownT,
                                                                                 c.origin[2], h->dp, rough, xs, ys, guessed, pos, neg, vis,
                                                                                 in_smem, h->col_minz, h->flags + 8,
                                                                                 direct_dev ? kpos : nullptr, direct_dev ? kneg : nullptr,
                                                                                 direct_dev ? kvis : nullptr, direct_dev ? krough : nullptr,
                                                                                 RowShard{0, 1, h->p.xy_size}, PushSet{}, GridSignal{});
    }
    h->stats.kernel_launches += 2;
    rec(h, EV_MAPS, st);
    CUDA_TRY(cudaGetLastError());
    if (!host_count) CUDA_TRY(cudaMemcpyAsync(h->counters_host, c.counter, sizeof(int), cudaMemcpyDeviceToHost, st));
    const size_t bi = (size_t)S2 * sizeof(int), bd = (size_t)S2 * sizeof(double);
    GvomHandle::Pending& pd = h->pend;
    pd = GvomHandle::Pending{};
    pd.active = true; pd.c = &c; pd.st = st;
    if (direct_dev || out_mem == GVOM_NONE) {
        // nothing to move
    } else if (out_mem == GVOM_DEVICE) {
        if (positive) CUDA_TRY(cudaMemcpyAsync(positive, pos, bi, cudaMemcpyDeviceToDevice, st));
        if (negative) CUDA_TRY(cudaMemcpyAsync(negative, neg, bi, cudaMemcpyDeviceToDevice, st));
        if (visibility) CUDA_TRY(cudaMemcpyAsync(visibility, vis, bi, cudaMemcpyDeviceToDevice, st));
        if (roughness) CUDA_TRY(cudaMemcpyAsync(roughness, rough, bd, cudaMemcpyDeviceToDevice, st));
    } else {
        bool dev = false;
        const bool direct = positive && negative && visibility && roughness && is_pinned_or_device(positive, &dev) &&
                            is_pinned_or_device(negative, &dev) && is_pinned_or_device(visibility, &dev) &&
                            is_pinned_or_device(roughness, &dev);
        const size_t rough_off = ((3 * (size_t)S2 + 1) & ~size_t(1)) * sizeof(int);
        if (direct) {                      // caller's buffers are pinned: DMA straight into them
            if (negative == positive + S2 && visibility == negative + S2 &&
                reinterpret_cast<char*>(roughness) == reinterpret_cast<char*>(positive) + rough_off) {
                // laid out like the device result block: one transfer for all four maps
                CUDA_TRY(cudaMemcpyAsync(positive, pos, rough_off + bd, cudaMemcpyDeviceToHost, st));
            } else {
                CUDA_TRY(cudaMemcpyAsync(positive, pos, bi, cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaMemcpyAsync(negative, neg, bi, cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaMemcpyAsync(visibility, vis, bi, cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaMemcpyAsync(roughness, rough, bd, cudaMemcpyDeviceToHost, st));
            }
        } else {                           // pageable: one DMA into the pinned mirror, memcpy when it has landed
            CUDA_TRY(cudaMemcpyAsync(h->out_i_host, h->imaps, rough_off + bd, cudaMemcpyDeviceToHost, st));
            pd.from_mirror = true;
            pd.positive = positive; pd.negative = negative; pd.visibility = visibility; pd.roughness = roughness;
        }
    }
    rec(h, EV_D2H, st);
    h->have_maps = true;
    if (origin) {                          // gvom.py:385-388
        origin[0] = c.origin[0] * h->p.xy_resolution;
        origin[1] = c.origin[1] * h->p.xy_resolution;
        origin[2] = c.origin[2] * h->p.z_resolution;
    }
    return GVOM_OK;
}

int run_maps_and_output(GvomHandle* h, Combined& c, double origin[3], int32_t* positive, int32_t* negative,
                        double* roughness, int32_t* visibility, int32_t out_mem, cudaStream_t st) {
    if (int e = enqueue_maps_and_output(h, c, origin, positive, negative, roughness, visibility, out_mem, st)) return e;
    return finish_outputs(h);
}

}  // namespace

extern "C" {

int gvom_workspace_size(const GvomParams* p, int64_t max_points, int64_t max_combined_cells,
                        size_t* device_bytes, size_t* host_bytes) {
    if (int e = check_params(p, max_points)) return e;
    GvomHandle tmp;
    fill_sizes(&tmp, p, max_points, max_combined_cells);
    size_t hb = 0;
    const size_t db = carve(&tmp, nullptr, nullptr, &hb);
    if (device_bytes) *device_bytes = db;
    if (host_bytes) *host_bytes = hb;
    return GVOM_OK;
}

int gvom_create(const GvomParams* p, int64_t max_points, int64_t max_combined_cells, int device,
                void* device_ws, size_t device_bytes, void* host_ws, size_t host_bytes, GvomHandle** out) {
    if (int e = check_params(p, max_points)) return e;
    if (!out || !device_ws || !host_ws) return fail(GVOM_EINVAL, "NULL workspace or handle pointer");
    CUDA_TRY(cudaSetDevice(device));
    GvomHandle* h = new GvomHandle();
    h->device = device;
    fill_sizes(h, p, max_points, max_combined_cells);
    size_t hb = 0;
    const size_t db = carve(h, device_ws, host_ws, &hb);
    if (db > device_bytes || hb > host_bytes) {
        delete h;
        return fail(GVOM_EINVAL, "workspace smaller than gvom_workspace_size() asked for");
    }
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_stage, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_proc_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_input, cudaEventDisableTiming);
    for (int i = 0; i < 8 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&h->ev_chunk[i], cudaEventDisableTiming);
    for (int i = 0; i < EV_COUNT && e == cudaSuccess; ++i) e = cudaEventCreate(&h->ev[i]);
    int sms = 0;
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (sms > 0) h->sm_count = sms;
    {
        const bool v8 = p->xy_size % 8 == 0, s4 = p->xy_size % 4 == 0;
        h->grid_codes = v8 ? resident_grid(k_merge_codes<8, MERGE_FINISH>, 256, h->sm_count)
                           : s4 ? resident_grid(k_merge_codes<4, MERGE_FINISH>, 256, h->sm_count)
                                : resident_grid(k_merge_codes<1, MERGE_FINISH>, 256, h->sm_count);
        h->grid_cells = resident_grid(k_finish_cells, 128, h->sm_count);
        h->grid_cells2 = resident_grid(k_merge_cells2, 128, h->sm_count);
        h->grid_scan_cells = (p->xy_eigen_dist == 1 && p->z_eigen_dist == 1) ? resident_grid(k_scan_cells<1, 1>, 256, h->sm_count)
                                                                             : resident_grid(k_scan_cells<-1, -1>, 256, h->sm_count);
        if (const char* m = getenv("GVOM_VARIANT")) h->variant = (unsigned)strtoul(m, nullptr, 0);
        if (const char* m = getenv("GVOM_CHUNKS")) h->host_chunks = std::max(1, std::min(16, atoi(m)));
        h->grid_rows3 = std::min(resident_grid(k_merge_rows<3, MERGE_FULL>, 256, h->sm_count),
                                 std::min(resident_grid(k_merge_rows<3, MERGE_PARTIAL>, 256, h->sm_count),
                                          resident_grid(k_merge_rows<3, MERGE_FINISH>, 256, h->sm_count)));
        h->grid_rows_mirror = std::min(resident_grid(k_merge_rows_ind<3, false>, 256, h->sm_count), resident_grid(k_merge_rows_ind<3, true>, 256, h->sm_count));
        h->grid_rows_async[0] = resident_grid(k_merge_rows_async<1>, 32, h->sm_count, sizeof(MrWarp<1>));
        h->grid_rows_async[1] = resident_grid(k_merge_rows_async<2>, 32, h->sm_count, sizeof(MrWarp<2>));
    }
    // the cell grid of the scan kernels starts with tag 0 ("never used") everywhere; every slot map starts as "all
    // unknown" with an empty group mask (S1 ray-casts into a wiped map; the row merge relies on map and mask of its
    // destination being consistent)
    if (e == cudaSuccess) e = cudaMemsetAsync(h->cellid, 0, sizeof(unsigned) * (size_t)h->EV, h->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(h->flags, 0, sizeof(int) * 16, h->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(h->col_minz, 0x7f, sizeof(int) * 2 * (size_t)h->S2, h->stream);
    for (auto& sl : h->slots) {
        if (e == cudaSuccess) e = cudaMemsetAsync(sl.index_map, 0xff, sizeof(int) * (size_t)h->V, h->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(sl.gmask, 0, sizeof(unsigned) * ((size_t)h->V / 256 + 2), h->stream);
        if (e == cudaSuccess) k_fill_f32<<<h->sm_count * 4, 256, 0, h->stream>>>(sl.minh, (long long)h->cap, 1.0f);
    }
    // the scan kernels find their accumulator rows and min heights already clear (S2 resets behind S1)
    for (int q = 0; q < 2; ++q)
        if (e == cudaSuccess) e = cudaMemsetAsync(h->acc[q], 0, sizeof(double) * MOM * (size_t)(h->cap + h->gcap), h->stream);
    for (auto& c : h->comb) {
        if (e == cudaSuccess) e = cudaMemsetAsync(c.index_map, 0xff, sizeof(int) * (size_t)h->V, h->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(c.gmask, 0, sizeof(unsigned) * ((size_t)h->V / 256 + 2), h->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) {
        delete h;
        return fail(GVOM_ECUDA, std::string("gvom_create: ") + cudaGetErrorString(e));
    }
    memset(h->counters_host, 0, sizeof(int) * 8);            // (the caller's pinned block is not zeroed)
    if (const char* m = getenv("GVOM_H2D")) h->zero_copy = std::string(m) != "dma";
    h->active = h->stream;
    *out = h;
    return GVOM_OK;
}

int gvom_destroy(GvomHandle* h) {
    if (!h) return GVOM_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (int i = 0; i < EV_COUNT; ++i) cudaEventDestroy(h->ev[i]);
    cudaEventDestroy(h->ev_stage);
    cudaEventDestroy(h->ev_proc_done);
    cudaEventDestroy(h->ev_input);
    for (int i = 0; i < 8; ++i) cudaEventDestroy(h->ev_chunk[i]);
    cudaStreamDestroy(h->copy_stream);
    cudaStreamDestroy(h->stream);
    delete h->pool;
    delete h;
    return GVOM_OK;
}

// One scan, handle locked.  pc2 = PointCloud2 wire records (point_step / offsets) instead of an array of
// `stride` elements of `dtype`.
struct CloudDesc {
    const void* points; int64_t n; int32_t stride, dtype, mem;
    bool pc2; int32_t point_step, ox, oy, oz;
};

static int process_locked(GvomHandle* h, const CloudDesc& cd, const double ego[3], const double* T, cudaStream_t st) {
    const int64_t n = cd.n;
    const int32_t stride = cd.pc2 ? 3 : cd.stride, dtype = cd.pc2 ? GVOM_F64 : cd.dtype, mem = cd.mem;
    const void* points = cd.points;
    h->active = st;
    const GvomParams& p = h->p;

    // gvom.py:110-112, 138-141
    Frame fr;
    for (int k = 0; k < 3; ++k) { h->ego[k] = ego[k]; fr.ego[k] = ego[k]; }
    fr.origin[0] = std::floor(ego[0] / p.xy_resolution - p.xy_size / 2.0);
    fr.origin[1] = std::floor(ego[1] / p.xy_resolution - p.xy_size / 2.0);
    fr.origin[2] = std::floor(ego[2] / p.z_resolution - p.z_size / 2.0);
    for (int k = 0; k < 3; ++k)
        if (!(std::fabs(fr.origin[k]) < 1.0e9)) return fail(GVOM_EINVAL, "ego position out of range (|ego/res| >= 1e9)");
    fr.start[0] = (float)(ego[0] / p.xy_resolution);
    fr.start[1] = (float)(ego[1] / p.xy_resolution);
    fr.start[2] = (float)(ego[2] / p.z_resolution);
    for (int k = 0; k < 3; ++k) fr.io[k] = (int)fr.origin[k];
    {
        const uint32_t M = 0x4B400000u, S = (uint32_t)p.xy_size;
        fr.S2 = p.xy_size * p.xy_size;
        fr.cc = (int)((M + (uint32_t)fr.io[0]) + (M + (uint32_t)fr.io[1]) * S + (M + (uint32_t)fr.io[2]) * S * S);
    }
    Xform tf;
    tf.enabled = T ? 1 : 0;
    for (int k = 0; k < 12; ++k) tf.m[k] = T ? T[k] : 0.0;
    // the DDA's floor() shortcut holds while every coordinate the ray can take stays below 2^22 in magnitude
    bool fast = !(h->variant & VAR_NO_FASTFLOOR);
    for (int k = 0; k < 3; ++k)
        if (!(std::fabs((double)fr.start[k]) + std::max(p.xy_size, p.z_size) + 8.0 < 4194304.0)) fast = false;

    // ---- the scan is written into the (wiped) spare slot; the slot it replaces in the ring becomes the new spare
    const int target = h->spare;
    Slot& s = h->slots[target];
    if (++h->scan_tag > 255u) {                          // tags wrapped: forget every entry of the cell grid
        CUDA_TRY(cudaMemsetAsync(h->cellid, 0, sizeof(unsigned) * (size_t)h->EV, st));
        h->scan_tag = 1u;
    }
    const int par = (int)(h->stats.process_calls % 3), buf = (int)(h->stats.process_calls & 1);
    ScanOut O{};
    O.map = s.index_map; O.cellid = h->cellid; O.tag = h->scan_tag; O.counters = h->flags + 2 * par;
    O.acc = h->acc[buf]; O.minh = s.minh; O.cell_voxel = s.cell_voxel; O.cap = (int)h->cap; O.gcap = (int)h->gcap; O.ES = h->ES;

    rec(h, EV_START, st);
    // ---- input staging + S1.  Host clouds: zero-copy (the kernel streams pinned memory over PCIe while it ray-casts),
    // pageable ones through the pinned staging block chunk by chunk; or chunked DMA on a copy stream (GVOM_H2D=dma).
    const size_t esz = dtype == GVOM_F32 ? 4 : 8;
    const size_t row = (size_t)stride * esz;
    auto launch_s1 = [&](const void* base, int64_t first, int64_t count, int from_host) {
        if (count <= 0) return;
        const char* p0 = static_cast<const char*>(base) + (size_t)first * row;
        const dim3 g(blocks_for(count, 256)), b(256);
        if (dtype == GVOM_F32) {
            if (fast) launch(k_scan_points<float, true>, g, b, 0, st, (const float*)p0, stride, (int)count, from_host, tf, fr, h->dp, O);
            else launch(k_scan_points<float, false>, g, b, 0, st, (const float*)p0, stride, (int)count, from_host, tf, fr, h->dp, O);
        } else {
            if (fast) launch(k_scan_points<double, true>, g, b, 0, st, (const double*)p0, stride, (int)count, from_host, tf, fr, h->dp, O);
            else launch(k_scan_points<double, false>, g, b, 0, st, (const double*)p0, stride, (int)count, from_host, tf, fr, h->dp, O);
        }
        h->stats.kernel_launches++;
    };
    auto launch_s1_pc2 = [&](const void* base, int step, int ox, int oy, int oz, int64_t first, int64_t count) {
        if (count <= 0) return;
        const char* p0 = static_cast<const char*>(base) + (size_t)first * step;
        const dim3 g(blocks_for(count, 256)), b(256);
        if (fast) launch(k_scan_points_pc2<true>, g, b, 0, st, p0, step, ox, oy, oz, (int)count, tf, fr, h->dp, O);
        else launch(k_scan_points_pc2<false>, g, b, 0, st, p0, step, ox, oy, oz, (int)count, tf, fr, h->dp, O);
        h->stats.kernel_launches++;
    };
    auto ensure_pool = [&]() {
        if (!h->pool) {
            int helpers = std::max(3, std::min(7, (int)std::thread::hardware_concurrency() / 2 - 1));
            if (const char* e = getenv("GVOM_COPY_THREADS")) helpers = std::max(0, atoi(e) - 1);
            h->pool = new CopyPool(helpers);
        }
    };
    bool wait_for_input = false;
    bool pc2_done = false;
    if (cd.pc2 && n > 0) {
        if (mem == GVOM_DEVICE) {
            rec(h, EV_H2D, st);
            launch_s1_pc2(points, cd.point_step, cd.ox, cd.oy, cd.oz, 0, n);
            pc2_done = true;
        } else {
            if (h->stage_busy) CUDA_TRY(cudaEventSynchronize(h->ev_stage));   // stage_host still being read
            void* mapped = nullptr;
            if (h->zero_copy && cudaHostGetDevicePointer(&mapped, h->stage_host, 0) != cudaSuccess) { cudaGetLastError(); mapped = nullptr; }
            if (mapped && ((uintptr_t)mapped & 15)) mapped = nullptr;
            ensure_pool();
            const auto t0 = std::chrono::steady_clock::now();
            if (mapped) {
                // field extraction (threaded, non-temporal) into packed 16-byte records, pipelined chunk by chunk
                // with the zero-copy kernel that streams them over PCIe
                rec(h, EV_H2D, st);
                const int nchunks = n >= 131072 ? h->host_chunks : 1;
                const int64_t per = ((n + nchunks - 1) / nchunks + 255) & ~int64_t(255);
                for (int c = 0; c < nchunks; ++c) {
                    const int64_t first = c * per, count = std::min<int64_t>(per, n - first);
                    if (count <= 0) break;
                    h->pool->extract_xyz(h->stage_host + (size_t)first * 16, static_cast<const char*>(points) + (size_t)first * cd.point_step,
                                         count, cd.point_step, cd.ox, cd.oy, cd.oz, false);
                    launch_s1_pc2(mapped, 16, 0, 4, 8, first, count);
                }
                CUDA_TRY(cudaEventRecord(h->ev_stage, st));
                h->stage_busy = true;
                pc2_done = true;
            } else {
                // no mapped access: widen to float64 x 3 in the staging block and take the array path below
                h->pool->extract_xyz(h->stage_host, static_cast<const char*>(points), n, cd.point_step, cd.ox, cd.oy, cd.oz, true);
                points = h->stage_host;
            }
            h->last_stage_copy_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
        }
    }
    if (pc2_done) {
        // S1 launched above
    } else if (n > 0 && mem == GVOM_DEVICE) {
        const void* src = points;
        if (stride == 4 && ((uintptr_t)points & 15)) {   // vector loads need 16-byte alignment
            CUDA_TRY(cudaMemcpyAsync(h->stage_dev, points, (size_t)n * row, cudaMemcpyDeviceToDevice, st));
            src = h->stage_dev;
        }
        rec(h, EV_H2D, st);
        launch_s1(src, 0, n, 0);
        CUDA_TRY(cudaEventRecord(h->ev_input, st));      // gvom_wait_input(): the caller's buffer has been consumed
    } else if (n > 0) {
        bool dev = false;
        const bool staged = points == h->stage_host;     // PointCloud2 fallback: already in the pinned staging block
        const bool pinned = staged || is_pinned_or_device(points, &dev);
        if (!pinned && h->stage_busy) CUDA_TRY(cudaEventSynchronize(h->ev_stage));   // stage_host still being read
        void* mapped = nullptr;
        if (h->zero_copy) {
            const void* hostp = pinned ? points : h->stage_host;
            if (cudaHostGetDevicePointer(&mapped, const_cast<void*>(hostp), 0) != cudaSuccess) { cudaGetLastError(); mapped = nullptr; }
            if (mapped && ((uintptr_t)mapped & 15)) mapped = nullptr;   // the staged loads are 128-bit
        }
        if (mapped) {
            rec(h, EV_H2D, st);
            const int zc = (h->variant & VAR_NO_BULK_H2D) ? 1 : 2;         // 2: a block's chunk is fetched by the TMA engine
            if (pinned) {
                launch_s1(mapped, 0, n, zc);
            } else {
                // pageable: the staging copy (threaded, non-temporal) is pipelined with the kernel chunk by chunk
                ensure_pool();
                const auto t0 = std::chrono::steady_clock::now();
                const int nchunks = n >= 131072 ? h->host_chunks : 1;
                const int64_t per = ((n + nchunks - 1) / nchunks + 255) & ~int64_t(255);
                for (int c = 0; c < nchunks; ++c) {
                    const int64_t first = c * per, count = std::min<int64_t>(per, n - first);
                    if (count <= 0) break;
                    h->pool->copy(h->stage_host + (size_t)first * row, static_cast<const char*>(points) + (size_t)first * row,
                                  (size_t)count * row);
                    launch_s1(mapped, first, count, zc);
                }
                h->last_stage_copy_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
            }
            CUDA_TRY(cudaEventRecord(h->ev_stage, st));
        } else {
            // chunked DMA on a copy stream; every chunk is processed as soon as it has landed
            const int nchunks = n >= 65536 ? 4 : 1;
            const int64_t per = ((n + nchunks - 1) / nchunks + 255) & ~int64_t(255);
            CUDA_TRY(cudaEventRecord(h->ev_proc_done, st));   // stage_dev: previous scan's readers first
            CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->ev_proc_done, 0));
            for (int c = 0; c < nchunks; ++c) {
                const int64_t first = c * per, count = std::min<int64_t>(per, n - first);
                if (count <= 0) break;
                const size_t off = (size_t)first * row, bytes = (size_t)count * row;
                const char* from = static_cast<const char*>(points) + off;
                if (!pinned) { memcpy(h->stage_host + off, from, bytes); from = h->stage_host + off; }
                CUDA_TRY(cudaMemcpyAsync(h->stage_dev + off, from, bytes, cudaMemcpyHostToDevice, h->copy_stream));
                CUDA_TRY(cudaEventRecord(h->ev_chunk[c], h->copy_stream));
                CUDA_TRY(cudaStreamWaitEvent(st, h->ev_chunk[c], 0));
                launch_s1(h->stage_dev, first, count, 0);
            }
            CUDA_TRY(cudaEventRecord(h->ev_stage, h->copy_stream));
            rec(h, EV_H2D, st);
        }
        h->stage_busy = !pinned || staged;
        wait_for_input = pinned && !staged;              // the caller may reuse its buffer when we return
    } else {
        rec(h, EV_H2D, st);
    }
    rec(h, EV_POINTS, st);

    // ---- S2: cells of this scan; group mask of its slot; wipe of the slot that leaves the ring
    const int ring_i = h->buffer_index;
    const int leaving = h->phys[ring_i];
    Slot& old = h->slots[leaving];
    {
        CellArgs A{};
        A.map = s.index_map;
        A.gmask = (p.xy_size % 8 == 0) ? s.gmask : nullptr;
        A.cellid = h->cellid; A.tag = h->scan_tag;
        A.counters = h->flags + 2 * par; A.counters_prev = h->flags + 2 * ((par + 2) % 3); A.counters_next = h->flags + 2 * ((par + 1) % 3);
        A.acc = h->acc[buf]; A.acc_other = h->acc[1 - buf]; A.cell_voxel = s.cell_voxel;
        A.spare_minh = old.dirty ? old.minh : nullptr; A.spare_count = old.counter; A.gcap = (int)h->gcap;
        A.hit = s.hit; A.total = s.total; A.metrics = s.metrics; A.slot_count = s.counter;
        A.old_map = old.dirty ? old.index_map : nullptr;
        A.old_gmask = (old.dirty && old.has_gmask) ? old.gmask : nullptr;
        A.cap = (int)h->cap; A.ES = h->ES;
        if (p.xy_eigen_dist == 1 && p.z_eigen_dist == 1)
            launch(k_scan_cells<1, 1>, dim3(h->grid_scan_cells), dim3(256), 0, st, A, h->dp);
        else
            launch(k_scan_cells<-1, -1>, dim3(h->grid_scan_cells), dim3(256), 0, st, A, h->dp);
        h->stats.kernel_launches++;
    }
    rec(h, EV_SCELLS, st);
    if (h->mir.n > 0) {
        // mirrored multi-GPU combine: deliver the scan to the owners of its rows (posted NVLink stores)
        const GvomHandle::Mirror& m = h->mir;
        MirrorPush M{};
        M.n = m.n; M.self = m.self;
        for (int k = 0; k < m.n; ++k) M.base[k] = m.base[k];
        const size_t mo = m.o_mirrors + ((size_t)m.self * p.buffer_size + ring_i) * m.mirror_bytes;
        M.o_map = (long long)mo; M.o_gmask = (long long)(mo + m.f_gmask); M.o_hit = (long long)(mo + m.f_hit);
        M.o_tot = (long long)(mo + m.f_tot); M.o_minh = (long long)(mo + m.f_minh); M.o_met = (long long)(mo + m.f_met);
        M.o_entry = (long long)(m.o_table + ((size_t)m.self * p.buffer_size + ring_i) * MIRROR_ENTRY * sizeof(int));
        M.held = reinterpret_cast<unsigned*>(m.base[m.self] + m.o_held) + (size_t)ring_i * m.n * m.nsegp;
        M.nsegp = (int)m.nsegp;
        M.oy = fr.io[1];
        M.entry[0] = (int)((h->stats.process_calls + 1) & 0x7fffffff);
        M.entry[1] = fr.io[0]; M.entry[2] = fr.io[1]; M.entry[3] = fr.io[2];
        memcpy(&M.entry[8], ego, 3 * sizeof(double));       // a rank that has not scanned yet adopts origin and ego from here
        rec(h, EV_P0, st);
        if (h->variant & VAR_BULK_PUSH)
            launch(k_push_scan_bulk, dim3(h->sm_count * 4), dim3(256), 0, st, (const int*)s.index_map, (const unsigned*)s.gmask, (const int*)s.counter,
                   (const int*)s.cell_voxel, (const int*)s.hit, (const int*)s.total, (const float*)s.minh, (const double*)s.metrics, M, h->dp, (int)h->cap);
        else
        launch(k_push_scan, dim3(h->sm_count * 4), dim3(256), 0, st, (const int*)s.index_map, (const unsigned*)s.gmask, (const int*)s.counter,
               (const int*)s.cell_voxel, (const int*)s.hit, (const int*)s.total, (const float*)s.minh, (const double*)s.metrics, M, h->dp, (int)h->cap);
        h->stats.kernel_launches++;
        rec(h, EV_P1, st);
        h->prof_partial = h->profiling;
    }
    CUDA_TRY(cudaGetLastError());
    h->prof_process = h->profiling;
    // a caller-owned host buffer may be reused as soon as we return: wait for the transfer (not the kernels)
    if (wait_for_input) CUDA_TRY(cudaEventSynchronize(h->ev_stage));

    // gvom.py:198-216
    s.has_gmask = p.xy_size % 8 == 0;
    s.dirty = true;
    for (int k = 0; k < 3; ++k) s.origin[k] = fr.origin[k];
    s.valid = true;
    old.valid = false; old.dirty = false; old.has_gmask = false;      // wiped by S2 above (stream order)
    h->phys[ring_i] = target;
    h->spare = leaving;
    h->last_buffer_index = h->buffer_index;
    h->buffer_index = (h->buffer_index + 1) % p.buffer_size;
    h->stats.process_calls++;
    return GVOM_OK;
}

int gvom_process_pointcloud(GvomHandle* h, const void* points, int64_t n, int32_t stride, int32_t dtype,
                            int32_t mem, const double ego[3], const double* T, void* stream) {
    if (!h || !ego) return fail(GVOM_EINVAL, "NULL handle or ego");
    if (n < 0 || n > h->max_points) return fail(GVOM_ECAPACITY, "point count exceeds max_points");
    if (stride < 3 || stride > 4) return fail(GVOM_EINVAL, "stride must be 3 or 4 elements");
    if (dtype != GVOM_F32 && dtype != GVOM_F64) return fail(GVOM_EINVAL, "dtype must be GVOM_F32 or GVOM_F64");
    if (mem != GVOM_HOST && mem != GVOM_DEVICE) return fail(GVOM_EINVAL, "bad mem");
    if (n > 0 && !points) return fail(GVOM_EINVAL, "NULL points");
    std::lock_guard<std::mutex> lock(h->mu);
    CUDA_TRY(cudaSetDevice(h->device));
    const CloudDesc cd{points, n, stride, dtype, mem, false, 0, 0, 0, 0};
    return process_locked(h, cd, ego, T, stream ? (cudaStream_t)stream : h->stream);
}

int gvom_get_stream(GvomHandle* h, void** stream) {
    if (!h || !stream) return fail(GVOM_EINVAL, "NULL argument");
    *stream = (void*)h->stream;
    return GVOM_OK;
}

int gvom_wait_input(GvomHandle* h) {
    if (!h) return fail(GVOM_EINVAL, "NULL handle");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaEventSynchronize(h->ev_input));
    return GVOM_OK;
}

int gvom_process_pointcloud2(GvomHandle* h, const void* data, int64_t n, int32_t point_step, int32_t off_x,
                             int32_t off_y, int32_t off_z, int32_t mem, const double ego[3], const double* T,
                             void* stream) {
    if (!h || !ego) return fail(GVOM_EINVAL, "NULL handle or ego");
    if (n < 0 || n > h->max_points) return fail(GVOM_ECAPACITY, "point count exceeds max_points");
    if (point_step < 12 || point_step > 65536 || (point_step & 3)) return fail(GVOM_EINVAL, "point_step must be a multiple of 4, >= 12");
    const int32_t offs[3] = {off_x, off_y, off_z};
    for (int k = 0; k < 3; ++k)
        if (offs[k] < 0 || offs[k] + 4 > point_step || (offs[k] & 3)) return fail(GVOM_EINVAL, "field offsets must be multiples of 4 inside the record");
    if (mem != GVOM_HOST && mem != GVOM_DEVICE) return fail(GVOM_EINVAL, "bad mem");
    if (n > 0 && !data) return fail(GVOM_EINVAL, "NULL data");
    if (mem == GVOM_DEVICE && ((uintptr_t)data & 3)) return fail(GVOM_EINVAL, "device payload must be 4-byte aligned");
    std::lock_guard<std::mutex> lock(h->mu);
    CUDA_TRY(cudaSetDevice(h->device));
    const CloudDesc cd{data, n, 3, GVOM_F64, mem, true, point_step, off_x, off_y, off_z};
    return process_locked(h, cd, ego, T, stream ? (cudaStream_t)stream : h->stream);
}

// combine_maps with the handle locked: merge, 2-D stage, outputs enqueued; completed unless `async`
static int combine_locked(GvomHandle* h, double origin[3], int32_t* positive, int32_t* negative, double* roughness,
                          int32_t* visibility, int32_t out_mem, cudaStream_t st, bool async) {
    if (int e = finish_outputs(h)) return e;                // a pending asynchronous combine owns the pinned mirrors
    Slot& newest = h->ring(h->last_buffer_index);
    if (!newest.valid) return GVOM_NO_DATA;                 // gvom.py:225-227
    h->active = st;
    rec(h, EV_CSTART, st);
    MergeArgs A;
    build_sources(h, newest.origin, true, &A);
    Combined& c = h->comb[1 - h->cur];
    for (int k = 0; k < 3; ++k) c.origin[k] = newest.origin[k];   // gvom.py:229
    // flags[1] (running cell counter) and the column minima are left clean by the previous combine's C4
    {
        MergeOut O{};
        O.cmap = c.index_map; O.counter = h->flags + 8; O.cell_voxel = c.cell_voxel;
        O.col_occ = h->col_minz; O.col_free = h->col_minz + h->S2;
        O.gmask = (h->p.xy_size % 8 == 0) ? c.gmask : nullptr; O.cap = (int)h->ccap;
        launch_merge<MERGE_FULL>(h, A, O, st);
    }
    rec(h, EV_CODES, st);
    launch_cells(h, A, c, st);
    rec(h, EV_CELLS, st);
    h->stats.kernel_launches += 1;
    h->prof_combine = h->profiling;
    c.has_gmask = h->p.xy_size % 8 == 0;
    if (int e = enqueue_maps_and_output(h, c, origin, positive, negative, roughness, visibility, out_mem, st)) return e;
    if (!async)
        if (int e = finish_outputs(h)) return e;
    c.valid = true;                                          // gvom.py:302-308
    h->cur = 1 - h->cur;
    h->stats.combine_calls++;
    return GVOM_OK;
}

int gvom_combine_maps(GvomHandle* h, double origin[3], int32_t* positive, int32_t* negative, double* roughness,
                      int32_t* visibility, int32_t out_mem, void* stream) {
    if (!h) return fail(GVOM_EINVAL, "NULL handle");
    if (out_mem != GVOM_HOST && out_mem != GVOM_DEVICE && out_mem != GVOM_NONE) return fail(GVOM_EINVAL, "bad out_mem");
    std::lock_guard<std::mutex> lock(h->mu);
    CUDA_TRY(cudaSetDevice(h->device));
    return combine_locked(h, origin, positive, negative, roughness, visibility, out_mem,
                          stream ? (cudaStream_t)stream : h->stream, false);
}

int gvom_combine_maps_async(GvomHandle* h, double origin[3], int32_t* positive, int32_t* negative, double* roughness,
                            int32_t* visibility, int32_t out_mem, void* stream) {
    if (!h) return fail(GVOM_EINVAL, "NULL handle");
    if (out_mem != GVOM_HOST && out_mem != GVOM_DEVICE && out_mem != GVOM_NONE) return fail(GVOM_EINVAL, "bad out_mem");
    std::lock_guard<std::mutex> lock(h->mu);
    CUDA_TRY(cudaSetDevice(h->device));
    return combine_locked(h, origin, positive, negative, roughness, visibility, out_mem,
                          stream ? (cudaStream_t)stream : h->stream, true);
}

int gvom_combine_wait(GvomHandle* h) {
    if (!h) return fail(GVOM_EINVAL, "NULL handle");
    std::lock_guard<std::mutex> lock(h->mu);
    CUDA_TRY(cudaSetDevice(h->device));
    return finish_outputs(h);
}

// OccupancyGrid payloads from the library's result block (handle locked)
static int grids_locked(GvomHandle* h, double thr, double minr, double maxr, int8_t* out, int32_t out_mem, cudaStream_t st) {
    if (!h->have_maps || !h->comb[h->cur].valid) return GVOM_NO_DATA;
    const int S = h->p.xy_size;
    const size_t S2 = (size_t)h->S2, bytes = GVOM_GRID_COUNT * S2;
    const int* pos = h->v_imaps;
    signed char* dst = (out_mem == GVOM_DEVICE) ? reinterpret_cast<signed char*>(out) : h->grids_dev;
    bool mapped_out = false;
    if (out_mem == GVOM_HOST && h->zero_copy && !(h->variant & VAR_DMA_OUT)) {   // pinned: the kernel writes through the mapping
        bool dev = false;
        void* m = nullptr;
        if (is_pinned_or_device(out, &dev) && !dev && cudaHostGetDevicePointer(&m, out, 0) == cudaSuccess && m) {
            dst = static_cast<signed char*>(m);
            mapped_out = true;
        } else cudaGetLastError();
    }
    const int T = (S + 31) / 32;
    launch(k_occupancy_grids, dim3(T, T), dim3(256), 0, st, pos, pos + S2, pos + 2 * S2, (const double*)h->v_rough, S, thr, minr, maxr, dst);
    h->stats.kernel_launches++;
    CUDA_TRY(cudaGetLastError());
    if (out_mem == GVOM_DEVICE || mapped_out) {
        CUDA_TRY(cudaStreamSynchronize(st));
        return GVOM_OK;
    }
    bool dev = false;
    if (is_pinned_or_device(out, &dev) && !dev) {           // pinned: straight into the caller's block
        CUDA_TRY(cudaMemcpyAsync(out, h->grids_dev, bytes, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    } else {
        CUDA_TRY(cudaMemcpyAsync(h->grids_host, h->grids_dev, bytes, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        memcpy(out, h->grids_host, bytes);
    }
    return GVOM_OK;
}

int gvom_occupancy_grids(GvomHandle* h, double density_threshold, double min_roughness, double max_roughness,
                         int8_t* out, int32_t out_mem, void* stream) {
    if (!h || !out) return fail(GVOM_EINVAL, "NULL argument");
    if (out_mem != GVOM_HOST && out_mem != GVOM_DEVICE) return fail(GVOM_EINVAL, "bad out_mem");
    std::lock_guard<std::mutex> lock(h->mu);
    CUDA_TRY(cudaSetDevice(h->device));
    if (int e = finish_outputs(h)) return e;
    cudaStream_t st = stream ? (cudaStream_t)stream : h->active;
    return grids_locked(h, density_threshold, min_roughness, max_roughness, out, out_mem, st);
}

int gvom_combine_maps_grids(GvomHandle* h, double origin[3], double density_threshold, double min_roughness,
                            double max_roughness, int8_t* out, int32_t out_mem, void* stream) {
    if (!h || !out) return fail(GVOM_EINVAL, "NULL argument");
    if (out_mem != GVOM_HOST && out_mem != GVOM_DEVICE) return fail(GVOM_EINVAL, "bad out_mem");
    std::lock_guard<std::mutex> lock(h->mu);
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    // the maps stay on the device (GVOM_NONE); the int8 grids are the only D2H traffic of the tick
    const int r = combine_locked(h, origin, nullptr, nullptr, nullptr, nullptr, GVOM_NONE, st, true);
    if (r != GVOM_OK) return r;
    const int g = grids_locked(h, density_threshold, min_roughness, max_roughness, out, out_mem, st);
    const int f = finish_outputs(h);                        // stream already drained: publishes the cell count
    return g != GVOM_OK ? g : f;
}

int gvom_combined_cell_count(GvomHandle* h, int64_t* cells) {
    if (!h || !cells) return fail(GVOM_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lock(h->mu);
    cudaSetDevice(h->device);
    if (int e = finish_outputs(h)) return e;
    if (!h->comb[h->cur].valid) return GVOM_NO_DATA;
    *cells = h->comb[h->cur].cells;
    return GVOM_OK;
}

int gvom_debug_voxel_map(GvomHandle* h, float* out, int64_t capacity_rows, int64_t* rows) {
    if (!h || !out) return fail(GVOM_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lock(h->mu);
    CUDA_TRY(cudaSetDevice(h->device));
    if (int e = finish_outputs(h)) return e;
    Combined& c = h->comb[h->cur];
    if (!c.valid) return GVOM_NO_DATA;
    if (capacity_rows < c.cells) return fail(GVOM_ECAPACITY, "output has fewer rows than combined cells");
    if (c.cells > 0) {
        k_debug_voxels<<<h->sm_count * 4, 256, 0, h->active>>>(c.counter, c.cell_voxel, c.hit, c.total, c.eig, c.origin[0],
                                                              c.origin[1], c.origin[2], h->dp, (int)h->ccap, h->debug_dev);
        h->stats.kernel_launches++;
        CUDA_TRY(cudaMemcpyAsync(out, h->debug_dev, sizeof(float) * 8 * (size_t)c.cells, cudaMemcpyDeviceToHost, h->active));
        CUDA_TRY(cudaStreamSynchronize(h->active));
    }
    if (rows) *rows = c.cells;
    return GVOM_OK;
}

static int debug_height_common(GvomHandle* h, float* out7, float* out3) {
    std::lock_guard<std::mutex> lock(h->mu);
    CUDA_TRY(cudaSetDevice(h->device));
    if (int e = finish_outputs(h)) return e;
    Combined& c = h->comb[h->cur];
    if (!c.valid || !h->have_maps) return GVOM_NO_DATA;
    ensure_maps6(h);
    const size_t S2 = (size_t)h->S2;
    float* d7 = h->debug_dev; float* d3 = h->debug_dev + 7 * S2;
    k_debug_height<<<blocks_for(h->S2, 256), 256, 0, h->active>>>(h->v_maps, h->v_rough, h->v_maps + 3 * S2, h->v_maps + 4 * S2,
                                                                h->v_maps + 5 * S2, c.origin[0], c.origin[1], h->dp,
                                                                out7 ? d7 : nullptr, out3 ? d3 : nullptr);
    h->stats.kernel_launches++;
    if (out7) CUDA_TRY(cudaMemcpyAsync(out7, d7, sizeof(float) * 7 * S2, cudaMemcpyDeviceToHost, h->active));
    if (out3) CUDA_TRY(cudaMemcpyAsync(out3, d3, sizeof(float) * 3 * S2, cudaMemcpyDeviceToHost, h->active));
    CUDA_TRY(cudaStreamSynchronize(h->active));
    return GVOM_OK;
}

int gvom_debug_height_map(GvomHandle* h, float* out) {
    if (!h || !out) return fail(GVOM_EINVAL, "NULL argument");
    return debug_height_common(h, out, nullptr);
}
int gvom_debug_inferred_height_map(GvomHandle* h, float* out) {
    if (!h || !out) return fail(GVOM_EINVAL, "NULL argument");
    return debug_height_common(h, nullptr, out);
}

// ---------------------------------------------------------------- test hooks
int gvom_last_slot(GvomHandle* h, int32_t* slot) {
    if (!h || !slot) return fail(GVOM_EINVAL, "NULL argument");
    *slot = h->last_buffer_index;
    return GVOM_OK;
}

int gvom_slot_info(GvomHandle* h, int32_t slot, int32_t* valid, int64_t* cells, double origin[3]) {
    if (!h || slot < 0 || slot >= h->p.buffer_size) return fail(GVOM_EINVAL, "bad slot");
    std::lock_guard<std::mutex> lock(h->mu);
    CUDA_TRY(cudaSetDevice(h->device));
    Slot& s = h->ring(slot);
    if (valid) *valid = s.valid ? 1 : 0;
    if (s.valid) {
        int cnt = 0;
        CUDA_TRY(cudaStreamSynchronize(h->active));
        CUDA_TRY(cudaMemcpy(&cnt, s.counter, sizeof(int), cudaMemcpyDeviceToHost));
        if (cells) *cells = std::min<int64_t>(cnt, h->cap);
        h->stats.scan_cells = cnt;
        if (origin) for (int k = 0; k < 3; ++k) origin[k] = s.origin[k];
        if (cnt > h->cap) return fail(GVOM_ECAPACITY, "scan has more occupied voxels than max_points");
    } else if (cells) *cells = 0;
    return GVOM_OK;
}

int gvom_export_slot(GvomHandle* h, int32_t slot, int32_t* index_map, int32_t* hit, int32_t* total, double* metrics,
                     float* min_height) {
    if (!h || slot < 0 || slot >= h->p.buffer_size) return fail(GVOM_EINVAL, "bad slot");
    int64_t cells = 0; int32_t valid = 0;
    if (int e = gvom_slot_info(h, slot, &valid, &cells, nullptr)) return e;
    if (!valid) return GVOM_NO_DATA;
    std::lock_guard<std::mutex> lock(h->mu);
    Slot& s = h->ring(slot);
    if (index_map) CUDA_TRY(cudaMemcpy(index_map, s.index_map, sizeof(int) * (size_t)h->V, cudaMemcpyDeviceToHost));
    if (hit) CUDA_TRY(cudaMemcpy(hit, s.hit, sizeof(int) * (size_t)cells, cudaMemcpyDeviceToHost));
    if (total) CUDA_TRY(cudaMemcpy(total, s.total, sizeof(int) * (size_t)cells, cudaMemcpyDeviceToHost));
    if (metrics) CUDA_TRY(cudaMemcpy(metrics, s.metrics, sizeof(double) * 10 * (size_t)cells, cudaMemcpyDeviceToHost));
    if (min_height) CUDA_TRY(cudaMemcpy(min_height, s.minh, sizeof(float) * (size_t)cells, cudaMemcpyDeviceToHost));
    return GVOM_OK;
}

int gvom_export_combined(GvomHandle* h, int32_t* index_map, int32_t* hit, int32_t* total, float* min_height,
                         float* metrics, float* eig, double* maps6) {
    if (!h) return fail(GVOM_EINVAL, "NULL handle");
    std::lock_guard<std::mutex> lock(h->mu);
    CUDA_TRY(cudaSetDevice(h->device));
    if (int e = finish_outputs(h)) return e;
    Combined& c = h->comb[h->cur];
    if (!c.valid) return GVOM_NO_DATA;
    if (maps6) ensure_maps6(h);
    CUDA_TRY(cudaStreamSynchronize(h->active));
    const size_t n = (size_t)c.cells;
    if (index_map) CUDA_TRY(cudaMemcpy(index_map, c.index_map, sizeof(int) * (size_t)h->V, cudaMemcpyDeviceToHost));
    if (hit) CUDA_TRY(cudaMemcpy(hit, c.hit, sizeof(int) * n, cudaMemcpyDeviceToHost));
    if (total) CUDA_TRY(cudaMemcpy(total, c.total, sizeof(int) * n, cudaMemcpyDeviceToHost));
    if (min_height) CUDA_TRY(cudaMemcpy(min_height, c.minh, sizeof(float) * n, cudaMemcpyDeviceToHost));
    if (metrics) CUDA_TRY(cudaMemcpy(metrics, c.metrics, sizeof(float) * 10 * n, cudaMemcpyDeviceToHost));
    if (eig) CUDA_TRY(cudaMemcpy(eig, c.eig, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost));
    if (maps6) {
        const size_t mb = sizeof(double) * (size_t)h->S2;
        CUDA_TRY(cudaMemcpy(maps6, h->v_maps, 6 * mb, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(maps6 + 2 * (size_t)h->S2, h->v_rough, mb, cudaMemcpyDeviceToHost));
    }
    return GVOM_OK;
}

int gvom_get_stats(GvomHandle* h, GvomStats* out) {
    if (!h || !out) return fail(GVOM_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lock(h->mu);
    *out = h->stats;
    return GVOM_OK;
}

int gvom_set_profiling(GvomHandle* h, int32_t on) {
    if (!h) return fail(GVOM_EINVAL, "NULL handle");
    std::lock_guard<std::mutex> lock(h->mu);
    h->profiling = on != 0;
    if (!on) h->prof_process = h->prof_combine = h->prof_partial = h->prof_rows = false;
    return GVOM_OK;
}

int gvom_stage_times(GvomHandle* h, float ms[16]) {
    if (!h || !ms) return fail(GVOM_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lock(h->mu);
    CUDA_TRY(cudaSetDevice(h->device));
    for (int i = 0; i < 16; ++i) ms[i] = 0.f;
    CUDA_TRY(cudaStreamSynchronize(h->active));
    if (h->prof_process) {          // [0] input staging, [1] S1 points (voxelise + claim + moments + ray-cast), [2] S2 cells
        CUDA_TRY(cudaEventElapsedTime(&ms[0], h->ev[EV_START], h->ev[EV_H2D]));
        CUDA_TRY(cudaEventElapsedTime(&ms[1], h->ev[EV_H2D], h->ev[EV_POINTS]));
        CUDA_TRY(cudaEventElapsedTime(&ms[2], h->ev[EV_POINTS], h->ev[EV_SCELLS]));
    }
    ms[9] = h->last_stage_copy_ms;
    if (h->prof_partial) CUDA_TRY(cudaEventElapsedTime(&ms[13], h->ev[EV_P0], h->ev[EV_P1]));   // multi-GPU: push of the scan (mirrored) or partial merge + cells (generic)
    if (h->prof_rows) {    // mirrored combine: [5] flag exchange + own rows, [6] own cells + heights, [15] wait + bit maps + surface, [8] wait + deliver
        CUDA_TRY(cudaEventElapsedTime(&ms[5], h->ev[EV_CSTART], h->ev[EV_CODES]));
        CUDA_TRY(cudaEventElapsedTime(&ms[6], h->ev[EV_CODES], h->ev[EV_CELLS]));
        CUDA_TRY(cudaEventElapsedTime(&ms[15], h->ev[EV_CELLS], h->ev[EV_X1]));
        CUDA_TRY(cudaEventElapsedTime(&ms[8], h->ev[EV_X1], h->ev[EV_D2H]));
        ms[7] = ms[15];
    }
    if (h->prof_combine) {
        const int a[4] = {EV_CSTART, EV_CODES, EV_CELLS, EV_MAPS};
        const int b[4] = {EV_CODES, EV_CELLS, EV_MAPS, EV_D2H};
        for (int i = 0; i < 4; ++i) CUDA_TRY(cudaEventElapsedTime(&ms[5 + i], h->ev[a[i]], h->ev[b[i]]));
    }
    return GVOM_OK;
}

int gvom_set_variant(GvomHandle* h, uint32_t mask) {
    if (!h) return fail(GVOM_EINVAL, "NULL handle");
    std::lock_guard<std::mutex> lock(h->mu);
    h->variant = mask;
    return GVOM_OK;
}

// ------------------------------------------------------------- state save / restore
}  // extern "C"
namespace {

struct StateHeader {
    char magic[8];                 // "GVOMST01"
    GvomParams p;
    int64_t max_points, cap, ccap, V;
    int32_t buffer_index, last_buffer_index, have_maps, comb_valid;
    double ego[3];
    double comb_origin[3];
    int64_t comb_cells;
    int64_t slot_cells[MAX_SLOTS];
    int32_t slot_valid[MAX_SLOTS];
    double slot_origin[MAX_SLOTS][3];
};

// sections of the blob after the header, in order; `visit(device_ptr, bytes)` is called for each
template <typename F>
void state_sections(GvomHandle* h, const StateHeader& H, F&& visit) {
    const size_t V = (size_t)h->V, S2 = (size_t)h->S2, gm = V / 256 + 2;
    for (int i = 0; i < h->p.buffer_size; ++i) {
        if (!H.slot_valid[i]) continue;
        Slot& s = h->ring(i);
        const size_t n = (size_t)H.slot_cells[i];
        visit(s.index_map, V * sizeof(int)); visit(s.gmask, gm * sizeof(unsigned));
        visit(s.hit, n * sizeof(int)); visit(s.total, n * sizeof(int)); visit(s.metrics, n * 10 * sizeof(double));
        visit(s.minh, n * sizeof(float)); visit(s.cell_voxel, n * sizeof(int)); visit(s.counter, sizeof(int));
    }
    if (H.comb_valid) {
        Combined& c = h->comb[h->cur];
        const size_t n = (size_t)H.comb_cells;
        visit(c.index_map, V * sizeof(int)); visit(c.gmask, gm * sizeof(unsigned));
        visit(c.hit, n * sizeof(int)); visit(c.total, n * sizeof(int)); visit(c.minh, n * sizeof(float));
        visit(c.metrics, n * 10 * sizeof(float)); visit(c.eig, n * 3 * sizeof(float)); visit(c.cell_voxel, n * sizeof(int));
        visit(c.counter, sizeof(int));
    }
    if (H.have_maps) {
        visit(h->v_maps, 6 * S2 * sizeof(double));
        visit(h->v_imaps, (3 * S2 + 2 * S2 + 2) * sizeof(int));
    }
}

int fill_state_header(GvomHandle* h, StateHeader* H) {
    if (int e = finish_outputs(h)) return e;
    ensure_maps6(h);
    CUDA_TRY(cudaStreamSynchronize(h->active));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    memset(H, 0, sizeof(*H));
    memcpy(H->magic, "GVOMST01", 8);
    H->p = h->p; H->max_points = h->max_points; H->cap = h->cap; H->ccap = h->ccap; H->V = h->V;
    H->buffer_index = h->buffer_index; H->last_buffer_index = h->last_buffer_index;
    H->have_maps = h->have_maps ? 1 : 0;
    for (int k = 0; k < 3; ++k) H->ego[k] = h->ego[k];
    for (int i = 0; i < h->p.buffer_size; ++i) {
        Slot& s = h->ring(i);
        H->slot_valid[i] = s.valid ? 1 : 0;
        if (!s.valid) continue;
        int cnt = 0;
        CUDA_TRY(cudaMemcpy(&cnt, s.counter, sizeof(int), cudaMemcpyDeviceToHost));
        H->slot_cells[i] = std::min<int64_t>(cnt, h->cap);
        for (int k = 0; k < 3; ++k) H->slot_origin[i][k] = s.origin[k];
    }
    Combined& c = h->comb[h->cur];
    H->comb_valid = c.valid ? 1 : 0;
    if (c.valid) {
        H->comb_cells = c.cells;
        for (int k = 0; k < 3; ++k) H->comb_origin[k] = c.origin[k];
    }
    return GVOM_OK;
}

}  // namespace
extern "C" {

int gvom_state_size(GvomHandle* h, size_t* bytes) {
    if (!h || !bytes) return fail(GVOM_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lock(h->mu);
    CUDA_TRY(cudaSetDevice(h->device));
    StateHeader H;
    if (int e = fill_state_header(h, &H)) return e;
    size_t total = sizeof(StateHeader);
    state_sections(h, H, [&](void*, size_t b) { total += (b + 15) & ~size_t(15); });
    *bytes = total;
    return GVOM_OK;
}

int gvom_save_state(GvomHandle* h, void* blob, size_t capacity, size_t* written) {
    if (!h || !blob) return fail(GVOM_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lock(h->mu);
    CUDA_TRY(cudaSetDevice(h->device));
    StateHeader H;
    if (int e = fill_state_header(h, &H)) return e;
    size_t total = sizeof(StateHeader);
    state_sections(h, H, [&](void*, size_t b) { total += (b + 15) & ~size_t(15); });
    if (total > capacity) return fail(GVOM_ECAPACITY, "state blob larger than the buffer (ask gvom_state_size)");
    char* out = static_cast<char*>(blob);
    memcpy(out, &H, sizeof(H));
    size_t off = sizeof(StateHeader);
    cudaError_t err = cudaSuccess;
    state_sections(h, H, [&](void* d, size_t b) {
        if (b && err == cudaSuccess) err = cudaMemcpy(out + off, d, b, cudaMemcpyDeviceToHost);
        off += (b + 15) & ~size_t(15);
    });
    if (err != cudaSuccess) return fail(GVOM_ECUDA, std::string("gvom_save_state: ") + cudaGetErrorString(err));
    if (written) *written = total;
    return GVOM_OK;
}

int gvom_load_state(GvomHandle* h, const void* blob, size_t bytes) {
    if (!h || !blob) return fail(GVOM_EINVAL, "NULL argument");
    if (bytes < sizeof(StateHeader)) return fail(GVOM_EINVAL, "state blob too small");
    std::lock_guard<std::mutex> lock(h->mu);
    CUDA_TRY(cudaSetDevice(h->device));
    StateHeader H;
    memcpy(&H, blob, sizeof(H));
    // ---- validate everything before the handle is touched (a rejected blob leaves it as it was)
    if (memcmp(H.magic, "GVOMST01", 8) != 0) return fail(GVOM_EINVAL, "not a gvom_b200 state blob");
    if (memcmp(&H.p, &h->p, sizeof(GvomParams)) != 0 || H.V != h->V)
        return fail(GVOM_EINVAL, "state blob was saved with different parameters");
    // capacities may have GROWN since the blob was saved (the Python side re-creates the handle with a larger
    // max_points when a bigger cloud arrives and carries the state over); shrinking is refused
    if (H.max_points > h->max_points || H.cap > h->cap || H.ccap > h->ccap)
        return fail(GVOM_EINVAL, "state blob was saved with larger capacities than this handle has");
    for (int i = 0; i < h->p.buffer_size; ++i)
        if (H.slot_valid[i] && (H.slot_cells[i] < 0 || H.slot_cells[i] > H.cap)) return fail(GVOM_EINVAL, "corrupt state blob (slot cells)");
    if (H.comb_valid && (H.comb_cells < 0 || H.comb_cells > H.ccap)) return fail(GVOM_EINVAL, "corrupt state blob (combined cells)");
    if (H.buffer_index < 0 || H.buffer_index >= h->p.buffer_size || H.last_buffer_index < 0 || H.last_buffer_index >= h->p.buffer_size)
        return fail(GVOM_EINVAL, "corrupt state blob (ring position)");
    {
        const size_t V = (size_t)h->V, S2 = (size_t)h->S2, gm = V / 256 + 2;
        auto pad = [](size_t b) { return (b + 15) & ~size_t(15); };
        size_t total = sizeof(StateHeader);
        for (int i = 0; i < h->p.buffer_size; ++i) {
            if (!H.slot_valid[i]) continue;
            const size_t n = (size_t)H.slot_cells[i];
            total += pad(V * 4) + pad(gm * 4) + 2 * pad(n * 4) + pad(n * 80) + 2 * pad(n * 4) + pad(4);
        }
        if (H.comb_valid) {
            const size_t n = (size_t)H.comb_cells;
            total += pad(V * 4) + pad(gm * 4) + 3 * pad(n * 4) + pad(n * 40) + pad(n * 12) + pad(n * 4) + pad(4);
        }
        if (H.have_maps) total += pad(6 * S2 * 8) + pad((3 * S2 + 2 * S2 + 2) * 4);
        if (total > bytes) return fail(GVOM_EINVAL, "state blob truncated");
    }
    if (int e = finish_outputs(h)) return e;
    CUDA_TRY(cudaStreamSynchronize(h->active));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    // ---- host state: ring entry i lives in physical slot i again, the extra slot is the (wiped) spare
    h->buffer_index = H.buffer_index; h->last_buffer_index = H.last_buffer_index;
    h->have_maps = H.have_maps != 0;
    for (int k = 0; k < 3; ++k) h->ego[k] = H.ego[k];
    for (int i = 0; i < h->p.buffer_size; ++i) {
        h->phys[i] = i;
        Slot& s = h->slots[i];
        s.valid = H.slot_valid[i] != 0;
        s.dirty = s.valid;
        s.has_gmask = s.valid && (h->p.xy_size % 8 == 0);
        for (int k = 0; k < 3; ++k) s.origin[k] = H.slot_origin[i][k];
    }
    h->spare = h->p.buffer_size;
    { Slot& sp = h->slots[h->spare]; sp.valid = false; sp.dirty = false; sp.has_gmask = false; }
    h->cur = 0;
    Combined& c = h->comb[0];
    c.valid = H.comb_valid != 0; c.cells = H.comb_cells; c.has_gmask = c.valid && (h->p.xy_size % 8 == 0);
    for (int k = 0; k < 3; ++k) c.origin[k] = H.comb_origin[k];
    h->comb[1].valid = false; h->comb[1].cells = 0;
    h->stats.combined_cells = c.cells;
    const char* in = static_cast<const char*>(blob);
    size_t off = sizeof(StateHeader);
    cudaError_t err = cudaSuccess;
    for (auto& sl : h->slots) k_fill_f32<<<h->sm_count * 4, 256>>>(sl.minh, (long long)h->cap, 1.0f);
    for (int q = 0; q < 2 && err == cudaSuccess; ++q) err = cudaMemset(h->acc[q], 0, sizeof(double) * MOM * (size_t)(h->cap + h->gcap));
    if (err == cudaSuccess) err = cudaDeviceSynchronize();
    state_sections(h, H, [&](void* d, size_t b) {
        if (b && err == cudaSuccess) err = cudaMemcpy(d, in + off, b, cudaMemcpyHostToDevice);
        off += (b + 15) & ~size_t(15);
    });
    // buffers the blob does not describe go back to their start-up state
    const size_t gmb = sizeof(unsigned) * ((size_t)h->V / 256 + 2);
    for (auto& sl : h->slots) {
        if (sl.valid || err != cudaSuccess) continue;
        err = cudaMemset(sl.index_map, 0xff, sizeof(int) * (size_t)h->V);
        if (err == cudaSuccess) err = cudaMemset(sl.gmask, 0, gmb);
    }
    for (int q = 0; q < 2 && err == cudaSuccess; ++q) {
        if (q == 0 && c.valid) continue;
        err = cudaMemset(h->comb[q].index_map, 0xff, sizeof(int) * (size_t)h->V);
        if (err == cudaSuccess) err = cudaMemset(h->comb[q].gmask, 0, gmb);
    }
    if (err == cudaSuccess) err = cudaMemset(h->cellid, 0, sizeof(unsigned) * (size_t)h->EV);
    if (err == cudaSuccess) err = cudaMemset(h->flags, 0, sizeof(int) * 16);
    if (err == cudaSuccess) err = cudaMemset(h->col_minz, 0x7f, sizeof(int) * 2 * (size_t)h->S2);
    if (err != cudaSuccess) return fail(GVOM_ECUDA, std::string("gvom_load_state: ") + cudaGetErrorString(err));
    h->scan_tag = 0;
    h->stage_busy = false;
    h->v_maps = h->maps; h->v_imaps = h->imaps; h->v_rough = h->rough_out;
    h->maps6.active = false;
    return GVOM_OK;
}

// ------------------------------------------------------------- tooling: CUDA-graph replay of a tick
// Measures what capturing the tick in a CUDA graph would buy: the launches of ONE
// Process_pointcloud (device cloud) + combine_maps (outputs left in the library's block) are captured -- PDL edges
// included -- instantiated, and replayed `iters` times; then the same tick is issued `iters` times as plain
// PDL-chained launches from this thread.  Both are timed with CUDA events around the whole batch, so host launch
// cost shows up only where the GPU starves.  The replay re-runs the same kernels on the same buffers (the maps it
// produces are not meaningful) and the handle must not be used for mapping afterwards.
int gvom_graph_probe(GvomHandle* h, const void* points_dev, int64_t n, int32_t stride, int32_t dtype, const double ego[3],
                     const double* T, int32_t iters, float* graph_ms_per_tick, float* launch_ms_per_tick, int32_t* graph_nodes) {
    if (!h || !points_dev || !ego || !graph_ms_per_tick || !launch_ms_per_tick || iters < 1) return fail(GVOM_EINVAL, "bad argument");
    if (n < 1 || n > h->max_points) return fail(GVOM_ECAPACITY, "point count exceeds max_points");
    std::lock_guard<std::mutex> lock(h->mu);
    CUDA_TRY(cudaSetDevice(h->device));
    if (int e = finish_outputs(h)) return e;
    cudaStream_t st = h->stream;
    CUDA_TRY(cudaStreamSynchronize(st));
    const bool prof = h->profiling;
    h->profiling = false;
    const CloudDesc cd{points_dev, n, stride, dtype, GVOM_DEVICE, false, 0, 0, 0, 0};
    double org[3];
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    int rc = process_locked(h, cd, ego, T, st);
    if (rc == GVOM_OK) rc = combine_locked(h, org, nullptr, nullptr, nullptr, nullptr, GVOM_NONE, st, true);
    cudaError_t ce = cudaStreamEndCapture(st, &graph);
    h->pend.active = false;                                  // (nothing was executed: no outputs to complete)
    if (rc != GVOM_OK || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        h->profiling = prof;
        return rc != GVOM_OK ? rc : fail(GVOM_ECUDA, std::string("graph capture: ") + cudaGetErrorString(ce));
    }
    size_t nodes = 0;
    cudaGraphGetNodes(graph, nullptr, &nodes);
    if (graph_nodes) *graph_nodes = (int32_t)nodes;
    ce = cudaGraphInstantiate(&exec, graph, 0);
    if (ce != cudaSuccess) {
        cudaGraphDestroy(graph);
        h->profiling = prof;
        return fail(GVOM_ECUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce));
    }
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int w = 0; w < 3; ++w) cudaGraphLaunch(exec, st);
    cudaStreamSynchronize(st);
    cudaEventRecord(a, st);
    for (int i = 0; i < iters; ++i) cudaGraphLaunch(exec, st);
    cudaEventRecord(b, st);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    *graph_ms_per_tick = ms / iters;
    for (int w = 0; w < 3; ++w) { process_locked(h, cd, ego, T, st); combine_locked(h, org, nullptr, nullptr, nullptr, nullptr, GVOM_NONE, st, true); h->pend.active = false; }
    cudaStreamSynchronize(st);
    cudaEventRecord(a, st);
    for (int i = 0; i < iters; ++i) {
        process_locked(h, cd, ego, T, st);
        combine_locked(h, org, nullptr, nullptr, nullptr, nullptr, GVOM_NONE, st, true);
        h->pend.active = false;
    }
    cudaEventRecord(b, st);
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&ms, a, b);
    *launch_ms_per_tick = ms / iters;
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaGraphExecDestroy(exec);
    cudaGraphDestroy(graph);
    h->profiling = prof;
    CUDA_TRY(cudaGetLastError());
    return GVOM_OK;
}

// ------------------------------------------------------------- tooling: atomic roofline
int gvom_bench_atomics(int device, void* table_dev, int64_t table_words, int32_t per_thread, int32_t mode,
                       int32_t repeats, float* best_ms, int64_t* atomics_per_launch) {
    if (!table_dev || !best_ms || table_words < 1 || (table_words & (table_words - 1)))
        return fail(GVOM_EINVAL, "table_words must be a power of two");
    CUDA_TRY(cudaSetDevice(device));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int blocks = sms * 8;
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreate(&a));
    CUDA_TRY(cudaEventCreate(&b));
    CUDA_TRY(cudaMemset(table_dev, 0, sizeof(int) * (size_t)table_words));
    float best = 1e30f;
    for (int r = 0; r < repeats + 2; ++r) {
        CUDA_TRY(cudaEventRecord(a, 0));
        k_atomic_bench<<<blocks, 256>>>((int*)table_dev, (unsigned)(table_words - 1), per_thread, mode);
        CUDA_TRY(cudaEventRecord(b, 0));
        CUDA_TRY(cudaEventSynchronize(b));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
        if (r >= 2 && ms < best) best = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *best_ms = best;
    if (atomics_per_launch) *atomics_per_launch = (int64_t)blocks * 256 * per_thread;
    return GVOM_OK;
}

// ------------------------------------------------------------- multi-GPU
int gvom_newest_origin(GvomHandle* h, double origin[3]) {
    if (!h || !origin) return fail(GVOM_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lock(h->mu);
    Slot& newest = h->ring(h->last_buffer_index);
    if (!newest.valid) return GVOM_NO_DATA;
    for (int k = 0; k < 3; ++k) origin[k] = newest.origin[k];
    return GVOM_OK;
}

static int combine_partial_impl(GvomHandle* h, const double origin[3], int32_t* code_grid_dev, uint32_t* group_mask_dev,
                                float* records_dev, int64_t record_capacity, int32_t* record_count_dev,
                                int32_t* const* signal_slots, int32_t n_signal, int32_t epoch, void* stream) {
    if (!h || !origin || !code_grid_dev || !records_dev || !record_count_dev) return fail(GVOM_EINVAL, "NULL argument");
    if (record_capacity < 1 || record_capacity > 2147483647LL) return fail(GVOM_EINVAL, "bad record capacity");
    std::lock_guard<std::mutex> lock(h->mu);
    CUDA_TRY(cudaSetDevice(h->device));
    if (int e = finish_outputs(h)) return e;
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    h->active = st;
    rec(h, EV_P0, st);
    MergeArgs A;
    build_sources(h, origin, false, &A);       // a rank without data contributes an empty grid
    CUDA_TRY(cudaMemsetAsync(record_count_dev, 0, sizeof(int), st));
    MergeOut O{};
    O.cmap = code_grid_dev; O.counter = record_count_dev; O.records = records_dev;
    O.gmask = (h->p.xy_size % 8 == 0) ? group_mask_dev : nullptr; O.cap = (int)record_capacity;
    launch_merge<MERGE_PARTIAL>(h, A, O, st);
    GridSignal G{};
    if (signal_slots && n_signal > 0) {            // peer-to-peer exchange: the last block of the cell kernel tells every rank
        if (n_signal > MAX_RANKS) return fail(GVOM_EINVAL, "too many ranks");
        G.S.n = n_signal;
        for (int k = 0; k < n_signal; ++k) G.S.slot[k] = signal_slots[k];
        G.counter = h->flags + 9; G.epoch = epoch;
    }
    launch(k_partial_cells, dim3(h->grid_cells), dim3(128), 0, st, A, record_count_dev, records_dev, h->dp, (int)record_capacity, G);
    h->stats.kernel_launches += 1;
    rec(h, EV_P1, st);
    h->prof_partial = h->profiling;
    CUDA_TRY(cudaGetLastError());
    return GVOM_OK;
}

int gvom_combine_partial(GvomHandle* h, const double origin[3], int32_t* code_grid_dev, uint32_t* group_mask_dev,
                         float* records_dev, int64_t record_capacity, int32_t* record_count_dev,
                         int32_t* const* signal_slots, int32_t n_signal, int32_t epoch, void* stream) {
    return combine_partial_impl(h, origin, code_grid_dev, group_mask_dev, records_dev, record_capacity, record_count_dev,
                                signal_slots, n_signal, epoch, stream);
}

int gvom_combine_finish(GvomHandle* h, const double origin[3], const int32_t* const* code_grids,
                        const uint32_t* const* group_masks, int32_t n_grids,
                        const float* const* records, const int32_t* const* record_counts, int32_t nranks,
                        int64_t record_capacity, const int32_t* wait_flags, int32_t wait_epoch, double origin_out[3],
                        int32_t* positive, int32_t* negative, double* roughness, int32_t* visibility, int32_t out_mem,
                        void* stream) {
    if (!h || !origin || !code_grids || !records || !record_counts) return fail(GVOM_EINVAL, "NULL argument");
    if (nranks < 1 || nranks > MAX_RANKS || n_grids < 1 || n_grids > MAX_RANKS) return fail(GVOM_EINVAL, "1..16 ranks / grids");
    RecordSet R;
    R.n = nranks;
    for (int k = 0; k < n_grids; ++k) if (!code_grids[k]) return fail(GVOM_EINVAL, "NULL grid");
    for (int k = 0; k < nranks; ++k) {
        R.r[k] = records[k]; R.count[k] = record_counts[k];
        if (!R.r[k] || !R.count[k]) return fail(GVOM_EINVAL, "NULL record buffer");
    }
    std::lock_guard<std::mutex> lock(h->mu);
    CUDA_TRY(cudaSetDevice(h->device));
    if (int e = finish_outputs(h)) return e;
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    h->active = st;
    rec(h, EV_CSTART, st);
    Combined& pc = h->comb[h->cur];
    Combined& c = h->comb[1 - h->cur];
    for (int k = 0; k < 3; ++k) c.origin[k] = origin[k];
    // sources of the merge pass: every rank's encoded grid (unshifted: all are in the common frame), then
    // this rank's copy of the previous combined map
    MergeArgs A;
    A.n = 0; A.use_masks = (h->p.xy_size % 8 == 0) ? 1 : 0;
    for (int k = 0; k < n_grids; ++k) {
        SlotRef& r = A.s[A.n++];
        r = SlotRef{};
        r.map = code_grids[k];
        r.gmask = group_masks ? group_masks[k] : nullptr;
        if (!r.gmask) A.use_masks = 0;
    }
    SlotRef prev{};
    const int has_prev = pc.valid ? 1 : 0;
    if (has_prev) {
        prev.map = pc.index_map; prev.metrics = pc.metrics; prev.hit = pc.hit; prev.total = pc.total; prev.minh = pc.minh;
        prev.dx = (int)(origin[0] - pc.origin[0]); prev.dy = (int)(origin[1] - pc.origin[1]); prev.dz = (int)(origin[2] - pc.origin[2]);
        prev.is_prev = 1;
        prev.gmask = pc.has_gmask ? pc.gmask : nullptr;
        if (!prev.gmask) A.use_masks = 0;
        A.s[A.n++] = prev;
    }
    double* cacc = h->cacc;
    {
        MergeOut O{};
        O.cmap = c.index_map; O.counter = h->flags + 8; O.cell_voxel = c.cell_voxel;
        O.col_occ = h->col_minz; O.col_free = h->col_minz + h->S2;
        O.gmask = (h->p.xy_size % 8 == 0) ? c.gmask : nullptr; O.cap = (int)h->ccap;
        O.cacc = cacc; O.chit = c.hit; O.ctot = c.total; O.cminh = c.minh;
        O.wait_flags = wait_flags; O.wait_n = nranks; O.wait_epoch = wait_epoch;
        launch_merge<MERGE_FINISH>(h, A, O, st);
    }
    rec(h, EV_CODES, st);
    launch(k_scatter_records, dim3(h->sm_count * 8), dim3(256), 0, st, R, record_capacity, c.index_map,
                                                      cacc, c.hit, c.total, c.minh);
    launch(k_finish_cells, dim3(h->grid_cells), dim3(128), 0, st, prev, has_prev, h->flags + 8, c.cell_voxel, cacc, c.hit, c.total, c.minh,
                                                   c.metrics, c.eig, h->dp, (int)h->ccap);
    rec(h, EV_CELLS, st);
    h->stats.kernel_launches += 2;
    h->prof_combine = h->profiling;
    c.has_gmask = h->p.xy_size % 8 == 0;
    const int r = run_maps_and_output(h, c, origin_out, positive, negative, roughness, visibility, out_mem, st);
    if (r != GVOM_OK) return r;
    c.valid = true;
    h->cur = 1 - h->cur;
    h->stats.combine_calls++;
    return GVOM_OK;
}


// ------------------------------------------------------------- multi-GPU, row-sharded finish
// Rank r owns the grid rows y with (y + origin_y) mod nranks == r -- WORLD rows, so an origin shift never moves a row
// (and with it the previous combined map of that row) to another rank.  Owning whole columns makes everything below
// the exchange local: the merge of the rank's rows from all ranks' encoded grids + its own previous rows, the cells on
// them, the column reductions and the 2-D stage of its columns.  Only 2-D data is replicated, by pushing (posted remote
// stores into every rank's 2-D block): heights after the column stage, the finished maps after the surface stage.
//   phase 1: [wait: partial results]  merge own rows -> cells -> heights of own columns pushed -> signal "heights"
//   phase 2: [wait: heights]          "height known" bit maps (whole map) -> surface stage of own rows pushed -> signal "results"
//   phase 4: [wait: results]          deliver the maps to the caller, refresh the library's own 2-D block
// (separate phases let one process play several ranks in the tests).
static size_t rows_block_bytes(const GvomHandle* h) {
    const size_t S2 = (size_t)h->S2;
    return 6 * S2 * sizeof(double) + (3 * S2 + 2 * S2 + 2) * sizeof(int) + 256;
}

int gvom_rows_block_size(GvomHandle* h, uint64_t* bytes) {
    if (!h || !bytes) return fail(GVOM_EINVAL, "NULL argument");
    *bytes = rows_block_bytes(h);
    return GVOM_OK;
}

// ---- mirrored ring slots (gvom_mirror.cuh).  One block per rank, same layout everywhere:
//   flags[64] | slot table [n][B][8] | source list (MergeArgs, local use) | held masks [B][n][nsegp] (local use) |
//   mirrors [n][B] x {index map [V] | group mask [nsegp] | hit [cap] | total [cap] | min height [cap] | metrics [cap][10] f64}
static void mirror_layout(const GvomHandle* h, int n, GvomHandle::Mirror* m) {
    auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
    const size_t V = (size_t)h->V, cap = (size_t)h->cap, B = (size_t)h->p.buffer_size;
    m->nsegp = (V / 256 + 2 + 63) & ~size_t(63);
    m->o_flags = 0;
    m->o_table = 256;
    m->o_args = up(m->o_table + (size_t)n * B * MIRROR_ENTRY * sizeof(int));
    m->o_held = up(m->o_args + sizeof(MergeArgs));
    m->o_mirrors = up(m->o_held + B * (size_t)n * m->nsegp * sizeof(unsigned));
    m->f_gmask = up(V * sizeof(int));
    m->f_hit = up(m->f_gmask + m->nsegp * sizeof(unsigned));
    m->f_tot = up(m->f_hit + cap * sizeof(int));
    m->f_minh = up(m->f_tot + cap * sizeof(int));
    m->f_met = up(m->f_minh + cap * sizeof(float));
    m->mirror_bytes = up(m->f_met + cap * 10 * sizeof(double));
    m->total = m->o_mirrors + (size_t)n * B * m->mirror_bytes + 256;
}

int gvom_mirror_block_size(GvomHandle* h, int32_t nranks, uint64_t* bytes) {
    if (!h || !bytes) return fail(GVOM_EINVAL, "NULL argument");
    if (nranks < 1 || nranks > MAX_RANKS) return fail(GVOM_EINVAL, "1..16 ranks");
    if ((int64_t)nranks * h->p.buffer_size > MAX_SLOTS) return fail(GVOM_EINVAL, "more than 64 ring slots over all ranks");
    if (h->p.xy_size % 256 != 0) return fail(GVOM_EINVAL, "mirrored ring slots need xy_size % 256 == 0");
    GvomHandle::Mirror m;
    mirror_layout(h, nranks, &m);
    *bytes = m.total;
    return GVOM_OK;
}

int gvom_mirror_attach(GvomHandle* h, int32_t rank, int32_t nranks, void* const* blocks) {
    if (!h || !blocks) return fail(GVOM_EINVAL, "NULL argument");
    uint64_t need = 0;
    if (int e = gvom_mirror_block_size(h, nranks, &need)) return e;
    if (rank < 0 || rank >= nranks) return fail(GVOM_EINVAL, "bad rank");
    for (int k = 0; k < nranks; ++k) if (!blocks[k]) return fail(GVOM_EINVAL, "NULL mirror block");
    std::lock_guard<std::mutex> lock(h->mu);
    CUDA_TRY(cudaSetDevice(h->device));
    if (h->stats.process_calls != 0) return fail(GVOM_EINVAL, "gvom_mirror_attach: the handle already holds scans");
    GvomHandle::Mirror m;
    mirror_layout(h, nranks, &m);
    m.n = nranks; m.self = rank;
    for (int k = 0; k < nranks; ++k) m.base[k] = static_cast<char*>(blocks[k]);
    // own block: flags, table, held masks and group masks clear; every mirror map "all unknown".  The caller may have
    // just filled the block on another stream (torch.zeros on the default stream does not order against this handle's
    // non-blocking stream: found by a racecheck run, where the fill landed AFTER the memsets below): wait for the device
    CUDA_TRY(cudaDeviceSynchronize());
    char* me = m.base[rank];
    CUDA_TRY(cudaMemsetAsync(me, 0, m.o_mirrors, h->stream));
    for (size_t q = 0; q < (size_t)nranks * h->p.buffer_size; ++q) {
        char* mq = me + m.o_mirrors + q * m.mirror_bytes;
        CUDA_TRY(cudaMemsetAsync(mq, 0xff, (size_t)h->V * sizeof(int), h->stream));
        CUDA_TRY(cudaMemsetAsync(mq + m.f_gmask, 0, m.nsegp * sizeof(unsigned), h->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->mir = m;
    return GVOM_OK;
}

int gvom_adopt_ego(GvomHandle* h, const double ego[3]) {
    if (!h || !ego) return fail(GVOM_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lock(h->mu);
    for (int k = 0; k < 3; ++k) h->ego[k] = ego[k];
    return GVOM_OK;
}

int gvom_combine_finish_rows(GvomHandle* h, const double origin[3], const GvomRowsLinks* K, int32_t epoch, int32_t phases,
                             double origin_out[3], int32_t* positive, int32_t* negative, double* roughness,
                             int32_t* visibility, int32_t out_mem, void* stream) {
    if (!h || !origin || !K) return fail(GVOM_EINVAL, "NULL argument");
    const int N = K->nranks, me = K->rank;
    if (N < 1 || N > MAX_RANKS || me < 0 || me >= N) return fail(GVOM_EINVAL, "bad rank / nranks");
    if (h->p.xy_size % 256 != 0) return fail(GVOM_EINVAL, "row-sharded finish needs xy_size % 256 == 0");
    if (out_mem != GVOM_HOST && out_mem != GVOM_DEVICE && out_mem != GVOM_NONE) return fail(GVOM_EINVAL, "bad out_mem");
    const bool mirrored = h->mir.n > 0;
    if (!mirrored) return fail(GVOM_EINVAL, "gvom_combine_finish_rows needs gvom_mirror_attach() first");
    if (h->mir.n != N || h->mir.self != me) return fail(GVOM_EINVAL, "links disagree with gvom_mirror_attach");
    for (int k = 0; k < N; ++k)
        if (!K->blocks2d[k] || !K->heights_slots[k] || !K->results_slots[k]) return fail(GVOM_EINVAL, "NULL rank buffer");
    if (!K->heights_flags || !K->results_flags) return fail(GVOM_EINVAL, "NULL flag table");
    std::lock_guard<std::mutex> lock(h->mu);
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    h->active = st;
    const int S = h->p.xy_size;
    const size_t S2 = (size_t)h->S2;
    Combined& pc = h->comb[h->cur];
    Combined& c = h->comb[1 - h->cur];
    // the rows this rank owns in the frame of this combine
    const long long oy = (long long)origin[1];
    RowShard R;
    R.n = N;
    R.y0 = (int)((((long long)me - oy) % N + N) % N);
    R.nrows = R.y0 < S ? (S - R.y0 + N - 1) / N : 0;
    PushSet D{};
    D.n = N; D.self = me;
    for (int k = 0; k < N; ++k) D.base[k] = static_cast<char*>(K->blocks2d[k]);
    D.off_maps = 0;
    D.off_pos = (long long)(6 * S2 * sizeof(double));
    D.off_neg = D.off_pos + (long long)(S2 * sizeof(int));
    D.off_vis = D.off_neg + (long long)(S2 * sizeof(int));
    D.off_rough = D.off_pos + (long long)(((3 * S2 + 1) & ~size_t(1)) * sizeof(int));
    char* mine = D.base[me];
    double* maps = reinterpret_cast<double*>(mine);
    int* imaps = reinterpret_cast<int*>(mine + D.off_pos);
    double* rough = reinterpret_cast<double*>(mine + D.off_rough);
    const int W = (S + 31) / 32;
    unsigned* known = h->known; unsigned* knownT = h->known + (size_t)S * W;

    auto mirror_sync = [&](int mode, const SlotRef& prev, int has_prev) -> MirrorSync {
        const GvomHandle::Mirror& m = h->mir;
        MirrorSync Y{};
        char* mine_m = m.base[me];
        for (int k = 0; k < N; ++k) Y.flag_slot[k] = reinterpret_cast<int*>(m.base[k] + m.o_flags) + me;
        Y.flags = reinterpret_cast<const int*>(mine_m + m.o_flags);
        Y.table = reinterpret_cast<const int*>(mine_m + m.o_table);
        Y.mirrors = mine_m + m.o_mirrors;
        Y.mirror_bytes = (long long)m.mirror_bytes; Y.f_gmask = (long long)m.f_gmask; Y.f_hit = (long long)m.f_hit;
        Y.f_tot = (long long)m.f_tot; Y.f_minh = (long long)m.f_minh; Y.f_met = (long long)m.f_met;
        Y.n = N; Y.B = h->p.buffer_size; Y.epoch = epoch; Y.mode = mode;
        Y.org[0] = (int)origin[0]; Y.org[1] = (int)origin[1]; Y.org[2] = (int)origin[2];
        Y.prev = prev; Y.has_prev = has_prev;
        Y.out = reinterpret_cast<MergeArgs*>(mine_m + m.o_args);
        void* mp = nullptr;
        if (cudaHostGetDevicePointer(&mp, h->counters_host, 0) == cudaSuccess) Y.err_flag = (int*)mp + 1; else cudaGetLastError();
        return Y;
    };
    if (mirrored && (phases & 16) && !(phases & 1)) {       // publish "my scans are pushed" only (tests: one process plays all ranks)
        launch(k_mirror_args, dim3(1), dim3(32), 0, st, mirror_sync(1, SlotRef{}, 0));
        h->stats.kernel_launches++;
        CUDA_TRY(cudaGetLastError());
    }
    if ((phases & 1) && mirrored) {
        // own world rows from the local mirrors of every rank's ring slots + own previous rows: the single-GPU kernels
        if (int e = finish_outputs(h)) return e;
        for (int k = 0; k < 3; ++k) c.origin[k] = origin[k];
        SlotRef prev{};
        const int has_prev = pc.valid ? 1 : 0;
        if (has_prev) {
            prev.map = pc.index_map; prev.metrics = pc.metrics; prev.hit = pc.hit; prev.total = pc.total; prev.minh = pc.minh;
            prev.dx = (int)(origin[0] - pc.origin[0]); prev.dy = (int)(origin[1] - pc.origin[1]); prev.dz = (int)(origin[2] - pc.origin[2]);
            prev.is_prev = 1;
            prev.gmask = pc.gmask;
        }
        rec(h, EV_CSTART, st);
        h->counters_host[1] = 0;
        const MergeArgs* Ad = reinterpret_cast<const MergeArgs*>(h->mir.base[me] + h->mir.o_args);
        const unsigned* srcmask = nullptr;
        {
            // flag exchange + source list (from the slot table the pushes carried), then the row merge of the own rows
            MergeOut O{};
            O.cmap = c.index_map; O.counter = h->flags + 8; O.cell_voxel = c.cell_voxel;
            O.col_occ = h->col_minz; O.col_free = h->col_minz + S2;
            O.gmask = c.gmask; O.cap = (int)h->ccap;
            O.row_y0 = R.y0; O.row_n = N;
            // which sources are occupied at each combined cell (<= 32 sources): saves the cell kernel 2N - 2 look-ups per cell
            // (from 9 sources up: below that the extra registers cost the row merge more than the cell kernel saves;
            // GVOM_VARIANT bit 16 switches it off for A/B runs)
            const int nsrc = N * h->p.buffer_size;
            if (nsrc >= 8 && nsrc <= 31 && !(h->variant & VAR_NO_SRCMASK)) O.srcmask = reinterpret_cast<unsigned*>(h->cacc);
            srcmask = O.srcmask;
            launch(k_mirror_args, dim3(1), dim3(32), 0, st, mirror_sync((phases & 32) ? 2 : 3, prev, has_prev));
            if (O.srcmask) launch(k_merge_rows_ind<3, true>, dim3(std::max(1, h->grid_rows_mirror)), dim3(256), 0, st, Ad, O, h->dp);
            else launch(k_merge_rows_ind<3, false>, dim3(std::max(1, h->grid_rows_mirror)), dim3(256), 0, st, Ad, O, h->dp);
        }
        rec(h, EV_CODES, st);
        {
            // fork: the heights of the own columns (k_rows_heights, then the flag exchange + bit maps of phase 2) run on the
            // side stream while the cell kernel runs here; joined before the surface stage (or at the end of this phase)
            CUDA_TRY(cudaEventRecord(h->ev_chunk[0], st));
            CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->ev_chunk[0], 0));
            launch(k_rows_heights, dim3(blocks_for((int64_t)S * R.nrows + 1, 256)), dim3(256), 0, h->copy_stream, Ad, (const int*)c.index_map,
                   (const int*)h->col_minz, (const int*)(h->col_minz + S2), c.origin[0], c.origin[1], c.origin[2], h->ego[0], h->ego[1],
                   h->ego[2], h->dp, R, D, srcmask);
            int* host_count = nullptr;
            void* m = nullptr;
            if (cudaHostGetDevicePointer(&m, h->counters_host, 0) == cudaSuccess) host_count = (int*)m; else cudaGetLastError();
            launch(k_merge_cells2_rows, dim3(h->grid_cells2), dim3(128), 0, st, Ad, (const int*)(h->flags + 8), (const int*)c.cell_voxel, c.hit,
                   c.total, c.minh, c.metrics, c.eig, h->dp, (int)h->ccap, c.counter, host_count, srcmask);
            if (!(phases & 2)) {                           // phases run one by one (tests): join here
                CUDA_TRY(cudaEventRecord(h->ev_chunk[1], h->copy_stream));
                CUDA_TRY(cudaStreamWaitEvent(st, h->ev_chunk[1], 0));
            }
        }
        rec(h, EV_CELLS, st);
        h->stats.kernel_launches += 4;
        h->prof_combine = false;
        CUDA_TRY(cudaGetLastError());
    }
    auto flag_set = [&](int32_t* const* slots) { SignalSet P{}; P.n = N; for (int k = 0; k < N; ++k) P.slot[k] = slots[k]; return P; };
    // mirrored combine: the "heights" / "results" flags are published by the first block of the kernel that then waits for
    // everybody's (phases & 256: they do not publish; 64 / 128: publish only -- one process playing several ranks)
    if (mirrored && (phases & 64)) { launch(k_signal, dim3(1), dim3(32), 0, st, flag_set(K->heights_slots), (int)epoch); h->stats.kernel_launches++; }
    if (mirrored && (phases & 128)) { launch(k_signal, dim3(1), dim3(32), 0, st, flag_set(K->results_slots), (int)epoch); h->stats.kernel_launches++; }
    const bool publish = mirrored && !(phases & 256);
    if (phases & 2) {
        // (all phases in one call: the bit maps follow the heights on the side stream, beside the cell kernel)
        const bool beside = mirrored && (phases & 1);
        cudaStream_t ks = beside ? h->copy_stream : st;
        launch(k_rows_known, dim3(W, W), dim3(1024), 0, ks, (const double*)maps, h->dp, known, knownT, K->heights_flags, N, (int)epoch,
               publish ? flag_set(K->heights_slots) : SignalSet{});
        if (beside) {
            CUDA_TRY(cudaEventRecord(h->ev_chunk[1], h->copy_stream));
            CUDA_TRY(cudaStreamWaitEvent(st, h->ev_chunk[1], 0));
        }
        const size_t mask_bytes = 2 * (size_t)S * W * sizeof(unsigned);
        const int in_smem = (mask_bytes <= 40 * 1024 && (mask_bytes % 16) == 0) ? 1 : 0;
        GridSignal G{};
        G.S.n = N;
        for (int k = 0; k < N; ++k) G.S.slot[k] = K->results_slots[k];
        G.counter = h->flags + 11; G.epoch = epoch;
        if (mirrored) G = GridSignal{};                     // published by k_rows_deliver instead
        launch(k_surface_maps2, dim3(std::max(1, blocks_for((int64_t)S * R.nrows, 128))), dim3(256), in_smem ? mask_bytes : 0, st, c.index_map, c.hit,
               c.total, (const double*)maps, (const double*)(maps + S2), known, knownT, c.origin[2], h->dp, rough, maps + 3 * S2,
               maps + 4 * S2, maps + 5 * S2, imaps, imaps + S2, imaps + 2 * S2, in_smem, h->col_minz, h->flags + 8,
               (int*)nullptr, (int*)nullptr, (int*)nullptr, (double*)nullptr, R, D, G);
        h->stats.kernel_launches += 2;
        rec(h, EV_X1, st);
        CUDA_TRY(cudaGetLastError());
    }
    if (phases & 4) {
        // wait for every rank's pushes, then transpose the exchange block ([y][x]) into the library's own 2-D block and the
        // caller's buffers ([x][y]); device and pinned host buffers are written by the kernel itself
        MapSet own{h->maps, h->imaps, h->imaps + S2, h->imaps + 2 * S2, h->rough_out};
        MapSet user{nullptr, nullptr, nullptr, nullptr, nullptr};
        bool direct = false;
        if (out_mem == GVOM_DEVICE && positive && negative && roughness && visibility) {
            user = MapSet{nullptr, positive, negative, visibility, roughness};
            direct = true;
        } else if (out_mem == GVOM_HOST && positive && negative && roughness && visibility && h->zero_copy) {
            void* mp[4] = {nullptr, nullptr, nullptr, nullptr};
            void* hp[4] = {positive, negative, visibility, roughness};
            bool ok = true, dev = false;
            for (int k = 0; k < 4 && ok; ++k) {
                ok = is_pinned_or_device(hp[k], &dev) && !dev && cudaHostGetDevicePointer(&mp[k], hp[k], 0) == cudaSuccess && mp[k];
                if (!ok) cudaGetLastError();
            }
            if (ok) { user = MapSet{nullptr, (int*)mp[0], (int*)mp[1], (int*)mp[2], (double*)mp[3]}; direct = true; }
        }
        // the four output maps now; the six work maps on demand (ensure_maps6)
        launch(k_rows_deliver, dim3(W, W, 4), dim3(256), 0, st, (const char*)mine, D, S, own, user, K->results_flags, N, (int)epoch, 6,
               publish ? flag_set(K->results_slots) : SignalSet{});
        h->maps6.active = true; h->maps6.blk = mine; h->maps6.D = D;
        h->stats.kernel_launches += 1;
        rec(h, EV_MAPS, st);
        CUDA_TRY(cudaGetLastError());
        const size_t bi = S2 * sizeof(int), bd = S2 * sizeof(double);
        const size_t rough_off = ((3 * S2 + 1) & ~size_t(1)) * sizeof(int);
        GvomHandle::Pending& pd = h->pend;
        pd = GvomHandle::Pending{};
        pd.active = true; pd.c = &c; pd.st = st;
        if (!direct && out_mem == GVOM_DEVICE) {
            if (positive) CUDA_TRY(cudaMemcpyAsync(positive, h->imaps, bi, cudaMemcpyDeviceToDevice, st));
            if (negative) CUDA_TRY(cudaMemcpyAsync(negative, h->imaps + S2, bi, cudaMemcpyDeviceToDevice, st));
            if (visibility) CUDA_TRY(cudaMemcpyAsync(visibility, h->imaps + 2 * S2, bi, cudaMemcpyDeviceToDevice, st));
            if (roughness) CUDA_TRY(cudaMemcpyAsync(roughness, h->rough_out, bd, cudaMemcpyDeviceToDevice, st));
        } else if (!direct && out_mem == GVOM_HOST) {       // pageable: one DMA into the pinned mirror, memcpy when it has landed
            CUDA_TRY(cudaMemcpyAsync(h->out_i_host, h->imaps, rough_off + bd, cudaMemcpyDeviceToHost, st));
            pd.from_mirror = true;
            pd.positive = positive; pd.negative = negative; pd.visibility = visibility; pd.roughness = roughness;
        }
        h->v_maps = h->maps; h->v_imaps = h->imaps; h->v_rough = h->rough_out;
        rec(h, EV_D2H, st);
        h->prof_rows = h->profiling;
        h->have_maps = true;
        if (origin_out) {
            origin_out[0] = c.origin[0] * h->p.xy_resolution;
            origin_out[1] = c.origin[1] * h->p.xy_resolution;
            origin_out[2] = c.origin[2] * h->p.z_resolution;
        }
        if (!(phases & 8))                                  // phases & 8: asynchronous, completed by the next call on the handle
            if (int e = finish_outputs(h)) return e;
        c.has_gmask = true;
        c.valid = true;
        h->cur = 1 - h->cur;
        h->stats.combine_calls++;
    }
    return GVOM_OK;
}


}  // extern "C"
