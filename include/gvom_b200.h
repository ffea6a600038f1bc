/*
 * gvom_b200.h -- C-ABI of the B200-native G-VOM voxel-mapping path.
 *
 * The reference has no FFI seam: its boundary is the Python class `Gvom`
 * (reference scripts/gvom.py:8-442, used by scripts/gvom_ros.py:44-59,109,115,
 * 171-185).  This header is the C boundary a replacement binds instead; each
 * entry point names the reference interface it replaces.  Plain pointers and
 * sizes only -- no torch / numpy / CUDA types in the signatures (`stream` is a
 * cudaStream_t passed as void*; NULL = the handle's own stream).
 *
 * Conventions: every function returns 0 on success, a positive GVOM_E* code on
 * error (gvom_last_error() holds the text); gvom_combine_maps* return
 * GVOM_NO_DATA (-1) when the newest ring-buffer slot is empty (the reference
 * prints "ERROR: No data in buffer" and returns None, gvom.py:225-227).
 *
 * Memory: the library never allocates device or pinned memory itself.  The
 * caller asks gvom_workspace_size() how much is needed, allocates both blocks
 * (the Python host side uses torch tensors for ownership) and hands them to
 * gvom_create(); they must outlive the handle.
 */
#ifndef GVOM_B200_H
#define GVOM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GVOM_OK 0
#define GVOM_NO_DATA (-1)
#define GVOM_EINVAL 1   /* bad argument */
#define GVOM_ECUDA 2    /* CUDA runtime error */
#define GVOM_ECAPACITY 3 /* more points / cells than the workspace was sized for */

/* point cloud element types (the reference keeps the caller's dtype on the device,
 * gvom.py:115, and its arithmetic depends on it) */
#define GVOM_F32 0
#define GVOM_F64 1
/* where a caller buffer lives */
#define GVOM_HOST 0
#define GVOM_DEVICE 1

/* Constructor parameters: replaces Gvom.__init__'s 14 positional arguments
 * (gvom.py:21-22), same order and meaning. */
typedef struct GvomParams {
    double xy_resolution;
    double z_resolution;
    int32_t xy_size;
    int32_t z_size;
    int32_t buffer_size;
    int32_t _pad0;
    double min_distance;
    double positive_obstacle_threshold;
    double negative_obstacle_threshold;
    double slope_obsacle_threshold;      /* sic: the reference's spelling */
    double robot_height;
    double robot_radius;
    double ground_to_lidar_height;
    int32_t xy_eigen_dist;
    int32_t z_eigen_dist;
} GvomParams;

/* Work counters of the last calls (device-side counts are read back lazily). */
typedef struct GvomStats {
    int64_t scan_cells;          /* occupied voxels of the last processed scan */
    int64_t combined_cells;      /* cells of the last combined map */
    int64_t kernel_launches;     /* kernels launched by this handle since creation */
    int64_t process_calls;
    int64_t combine_calls;
} GvomStats;

typedef struct GvomHandle GvomHandle;

const char* gvom_last_error(void);

/* Bytes of device and pinned-host workspace a handle needs. */
int gvom_workspace_size(const GvomParams* p, int64_t max_points, int64_t max_combined_cells,
                        size_t* device_bytes, size_t* host_bytes);

/* Replaces Gvom.__init__ (gvom.py:21-101).  max_combined_cells <= 0 selects the
 * default min(V, 4*max_points*(buffer_size+1)). */
int gvom_create(const GvomParams* p, int64_t max_points, int64_t max_combined_cells, int device,
                void* device_ws, size_t device_bytes, void* host_ws, size_t host_bytes,
                GvomHandle** out);
int gvom_destroy(GvomHandle* h);

/* Replaces Gvom.Process_pointcloud (gvom.py:105-220): transform (optional 4x4
 * row-major float64, NULL = none), voxelise, ray-cast, per-voxel moments, and
 * store the per-scan map in the next ring-buffer slot.
 * points: n rows of `stride` elements (>= 3; x,y,z first) of `dtype`, in host
 * (pageable or pinned) or device memory.  Asynchronous on `stream` except for
 * the staging copy of pageable host input.
 * Buffer contract: a HOST buffer may be reused as soon as the call returns.  A DEVICE buffer is read in place,
 * asynchronously, on `stream` (or the handle's own stream, gvom_get_stream()): the caller orders the producer of
 * the buffer before this call on that stream and keeps the buffer alive and unmodified until gvom_wait_input()
 * returns (or the stream has been synchronised). */
int gvom_process_pointcloud(GvomHandle* h, const void* points, int64_t n, int32_t stride,
                            int32_t dtype, int32_t mem, const double ego[3],
                            const double* transform16, void* stream);
/* The handle's own stream (cudaStream_t as void*), the one used when `stream` arguments are NULL; lets a caller
 * order its own work against the library's (cudaStreamWaitEvent). */
int gvom_get_stream(GvomHandle* h, void** stream);
/* Blocks until the device cloud of the last gvom_process_pointcloud* call has been consumed by the kernels. */
int gvom_wait_input(GvomHandle* h);

/* Replaces Gvom.combine_maps (gvom.py:222-393).  Outputs are [x,y]-indexed
 * (row-major, x major) xy_size*xy_size arrays: positive / negative obstacle and
 * visibility int32, roughness float64; origin is the combined map's world origin.
 * Output pointers may be host (pageable or pinned) or device memory (out_mem).
 * Returns after the outputs are complete (synchronises `stream`). */
int gvom_combine_maps(GvomHandle* h, double origin[3], int32_t* positive, int32_t* negative,
                      double* roughness, int32_t* visibility, int32_t out_mem, void* stream);

/* Replaces make_debug_voxel_map / make_debug_height_map /
 * make_debug_inferred_height_map (gvom.py:395-442).  GVOM_NO_DATA before the
 * first combine.  voxel rows: [x,y,z, hit/total, hit, l0-l1, l1-l2, l2] float32. */
int gvom_combined_cell_count(GvomHandle* h, int64_t* cells);
int gvom_debug_voxel_map(GvomHandle* h, float* out_rows8, int64_t capacity_rows, int64_t* rows);
int gvom_debug_height_map(GvomHandle* h, float* out_rows7);
int gvom_debug_inferred_height_map(GvomHandle* h, float* out_rows3);

/* ---- the callers and data formats either side of the path ---- */

/* where a caller buffer lives, continued: GVOM_NONE = "do not deliver this output" (combine only) */
#define GVOM_NONE 2

/* Replaces ros_numpy.point_cloud2.pointcloud2_to_xyz_array(data) followed by Process_pointcloud
 * (gvom_ros.py:108-109): `data` is the byte payload of a sensor_msgs/PointCloud2 -- n_points records of
 * point_step bytes holding little-endian float32 x / y / z at byte offsets off_x / off_y / off_z (all
 * multiples of 4).  The fields are widened to float64 exactly as ros_numpy does, records with a NaN / Inf
 * coordinate are dropped (remove_nans=True), and the result equals Process_pointcloud on that float64
 * array.  Host payloads (pageable or pinned) are reduced to packed 16-byte records by the library's staging
 * threads, so 16 instead of 24 (or point_step) bytes per point cross PCIe; device payloads are read in place. */
int gvom_process_pointcloud2(GvomHandle* h, const void* data, int64_t n_points, int32_t point_step,
                             int32_t off_x, int32_t off_y, int32_t off_z, int32_t mem,
                             const double ego[3], const double* transform16, void* stream);

/* Asynchronous combine (double-buffer the outputs on the caller's side): same arguments and results as
 * gvom_combine_maps, but returns as soon as everything is enqueued.  The outputs (pinned host or device
 * memory; pageable host memory is completed by a memcpy inside gvom_combine_wait) are valid after
 * gvom_combine_wait(), which also reports GVOM_ECAPACITY.  `origin` is written before the call returns.
 * Scans may be processed between the two calls; any other call on the handle waits implicitly. */
int gvom_combine_maps_async(GvomHandle* h, double origin[3], int32_t* positive, int32_t* negative,
                            double* roughness, int32_t* visibility, int32_t out_mem, void* stream);
int gvom_combine_wait(GvomHandle* h);

/* Replaces the OccupancyGrid post-processing of gvom_ros.py:142-164 (five numpy passes per map on the
 * host): from the maps of the last combine, GVOM_GRID_COUNT int8 grids of xy_size*xy_size cells, each
 * flattened in Fortran order (cell [x,y] at y*xy_size + x, the `order='F'` reshape of the node):
 *   HARD      max(100*(positive > density_threshold), negative)        (:143)
 *   SOFT      100*(positive <= density_threshold)*(positive > 0)       (:148)
 *   CERTAINTY visibility*100                                           (:153, published twice)
 *   NEGATIVE  negative                                                 (:159)
 *   ROUGHNESS ((clip(roughness, min, max) + min) / (max - min)) * 100  (:164, the reference's formula, sic)
 * converted like numpy's astype(int8) (truncate toward zero, keep the low byte).
 * out: GVOM_GRID_COUNT * xy_size * xy_size bytes in host (out_mem GVOM_HOST) or device memory. */
#define GVOM_GRID_HARD 0
#define GVOM_GRID_SOFT 1
#define GVOM_GRID_CERTAINTY 2
#define GVOM_GRID_NEGATIVE 3
#define GVOM_GRID_ROUGHNESS 4
#define GVOM_GRID_COUNT 5
int gvom_occupancy_grids(GvomHandle* h, double density_threshold, double min_roughness, double max_roughness,
                         int8_t* out, int32_t out_mem, void* stream);
/* combine_maps + the post-processing above in one call: only the int8 grids (5 bytes per cell instead
 * of 20) leave the device.  Returns GVOM_NO_DATA like gvom_combine_maps. */
int gvom_combine_maps_grids(GvomHandle* h, double origin[3], double density_threshold, double min_roughness,
                            double max_roughness, int8_t* out, int32_t out_mem, void* stream);

/* State save / restore (deterministic replay, regression corpora): ring slots, the last combined map
 * (gvom.py:302-308 `last_combined_*`), ego position and ring position, as one opaque host blob that can
 * be loaded into any handle created with the same parameters and the same or LARGER capacities (that is how the
 * Python host side grows max_points on demand). */
/* (The size changes with every processed scan: a caller that shares the handle with a scan thread retries
 * gvom_state_size + gvom_save_state when the latter answers GVOM_ECAPACITY.) */
int gvom_state_size(GvomHandle* h, size_t* bytes);
int gvom_save_state(GvomHandle* h, void* blob, size_t capacity, size_t* written);
int gvom_load_state(GvomHandle* h, const void* blob, size_t bytes);

/* ---- multi-GPU, generic exchange (any grid size; also the NCCL fallback): partial merge + replicated finish ----
 * Each rank pre-merges its own ring buffer into the common frame `origin`
 * (integral voxel units): a dense int32 code grid (V entries: occupied flag
 * 1<<26, else the pass count) and compact cell records (GVOM_RECORD_FLOATS
 * float32 each).  The caller either sums the grids and gathers the records across ranks
 * (NCCL via torch.distributed) or simply maps every rank's buffers into every other rank
 * (peer-to-peer over NVLink); gvom_combine_finish() then completes the combine on every
 * rank.  See DESIGN.md. */
#define GVOM_RECORD_FLOATS 16   /* voxel id, hit, total, min_h, 10 metrics, 2 pad */
int gvom_newest_origin(GvomHandle* h, double origin[3]);
/* group_mask_dev (V/256 + 2 words, may be NULL): one bit per 8-voxel group of the code grid,
 * set where the group holds anything; lets the finishing pass skip the (mostly empty) rest. */
/* signal_slots (n_signal device pointers, may be NULL): peer-to-peer exchange only -- one int32 slot in
 * every rank's (mapped) memory into which this rank writes `epoch` once its partial results are
 * visible system-wide. */
int gvom_combine_partial(GvomHandle* h, const double origin[3], int32_t* code_grid_dev,
                         uint32_t* group_mask_dev, float* records_dev, int64_t record_capacity,
                         int32_t* record_count_dev, int32_t* const* signal_slots, int32_t n_signal,
                         int32_t epoch, void* stream);
/* code_grids: n_grids device pointers -- ONE all-reduced grid (NCCL exchange) or one grid
 * per rank (peer-to-peer exchange: the other GPUs' buffers mapped over NVLink are read directly by
 * the finishing kernels).  records / record_counts: nranks device pointers each (a rank's records
 * and its record count).
 * wait_flags (nranks local int32 slots, may be NULL): peer-to-peer exchange -- the finishing kernel
 * itself waits until every slot is >= wait_epoch before it reads the peers' buffers (the barrier is
 * part of the consumer kernel; no collective launch, no host synchronisation). */
int gvom_combine_finish(GvomHandle* h, const double origin[3], const int32_t* const* code_grids,
                        const uint32_t* const* group_masks /* n_grids entries, or NULL */,
                        int32_t n_grids, const float* const* records,
                        const int32_t* const* record_counts, int32_t nranks, int64_t record_capacity,
                        const int32_t* wait_flags, int32_t wait_epoch,
                        double origin_out[3], int32_t* positive, int32_t* negative, double* roughness,
                        int32_t* visibility, int32_t out_mem, void* stream);

/* ---- multi-GPU, default (xy_size % 256 == 0): mirrored ring slots + row-sharded combine (gvom_mirror.cuh) ----
 * Replaces the roles of gvom.py:242-257 (merge over the whole ring) when the ring is spread over GPUs.  Rank r owns the
 * WORLD rows y with (y + origin_y) mod nranks == r (whole columns), so an origin shift never moves a row -- and the
 * previous combined map of that row -- to another rank; the 3-D combined state is SHARDED by rows.
 *  * Every rank owns a block of gvom_mirror_block_size() bytes that all ranks can write (symmetric memory).  After
 *    gvom_mirror_attach() every Process_pointcloud ends with one push kernel that stores the new slot's index-map rows,
 *    group-mask words and cell records into the memory of the rank that owns the row -- posted NVLink stores --
 *    together with a slot-table entry; every rank thus holds exact local copies of all ranks' ring slots for the rows
 *    it owns.  blocks[k] = rank k's block as mapped into this process; attach before the first scan and synchronise all
 *    ranks (barrier) before any of them scans.
 *  * gvom_combine_finish_rows() (collective; no NCCL call, no host synchronisation): one warp publishes this rank's epoch
 *    flag, waits for the others' and builds the source list on the device from the slot table; the single-GPU merge
 *    kernels run over the own rows from local memory; the column heights of the own rows are pushed into every rank's
 *    2-D block (symmetric memory, gvom_rows_block_size() bytes); after the "heights" flags the surface stage of the own
 *    rows runs and its maps are pushed; after the "results" flags the four output maps are delivered.  Every flag table
 *    has one entry per rank, written by that rank with the combine's epoch (1, 2, ...); a rank that never arrives makes
 *    the others give up after 10 s (GVOM_ECUDA); ranks whose newest scans disagree on the origin: GVOM_EINVAL.
 *    phases: 1 = own rows + cells + heights, 2 = surface stage, 4 = deliver; 7 = all; + 8 = return without waiting for
 *    the stream (outputs in device memory are valid in stream order; completed, and errors reported, by the next call
 *    on the handle).  One process playing several ranks (tests) publishes every flag for all ranks before any rank waits
 *    for it: 16 / 64 / 128 = publish the epoch / heights / results flag only; 32 (with 1) = the epoch flag is already
 *    out (also the start-up path of a rank without scans); 256 (with 2, 4) = the waiting kernels do not publish --
 *    i.e. 16, 1 | 32, 64, 2 | 256, 128, 4 | 256, each for all ranks in turn. */
#define GVOM_MAX_RANKS 16
typedef struct GvomRowsLinks {
    int32_t rank, nranks;
    void* blocks2d[GVOM_MAX_RANKS];                /* every rank's 2-D block as mapped here */
    int32_t* heights_slots[GVOM_MAX_RANKS];        /* this rank's "heights pushed" flag in every rank's memory */
    const int32_t* heights_flags;                  /* local: nranks flags */
    int32_t* results_slots[GVOM_MAX_RANKS];        /* this rank's "maps pushed" flag in every rank's memory */
    const int32_t* results_flags;                  /* local: nranks flags */
} GvomRowsLinks;
int gvom_mirror_block_size(GvomHandle* h, int32_t nranks, uint64_t* bytes);
int gvom_mirror_attach(GvomHandle* h, int32_t rank, int32_t nranks, void* const* blocks);
int gvom_rows_block_size(GvomHandle* h, uint64_t* bytes);
int gvom_combine_finish_rows(GvomHandle* h, const double origin[3], const GvomRowsLinks* links, int32_t epoch,
                             int32_t phases, double origin_out[3], int32_t* positive, int32_t* negative,
                             double* roughness, int32_t* visibility, int32_t out_mem, void* stream);
/* A rank that has not scanned yet takes part in a combine with the origin AND the vehicle position of a rank that has
 * (slot-table entry: 16 int32 {sequence (0: empty), ox, oy, oz, cells, -, -, -, ego xyz as 3 float64, -, -} at byte 256 +
 * 64 * (rank * buffer_size + slot) of the block): the ego disc of the height map (gvom.py:560-571) needs it. */
int gvom_adopt_ego(GvomHandle* h, const double ego[3]);

/* ---- test / tooling hooks (canonical parity dumps; not on the hot path) ---- */
int gvom_slot_info(GvomHandle* h, int32_t slot, int32_t* valid, int64_t* cells, double origin[3]);
int gvom_last_slot(GvomHandle* h, int32_t* slot);
/* host outputs: index_map[V] int32, hit/total[cells] int32, metrics[cells*10] float64,
 * min_height[cells] float32 (any may be NULL) */
int gvom_export_slot(GvomHandle* h, int32_t slot, int32_t* index_map, int32_t* hit, int32_t* total,
                     double* metrics, float* min_height);
/* host outputs of the last combined map (any may be NULL): index_map[V], hit/total[cells],
 * min_height[cells], metrics[cells*10] float32, eig[cells*3] float32, and the six float64
 * xy*xy maps height, inferred, roughness, x_slope, y_slope, guessed. */
int gvom_export_combined(GvomHandle* h, int32_t* index_map, int32_t* hit, int32_t* total,
                         float* min_height, float* metrics, float* eig, double* maps6);
int gvom_get_stats(GvomHandle* h, GvomStats* out);
/* CUDA-event time (ms) of the stages of the last process / combine call:
 * [0] H2D+staging, [1] scan points (voxelise, claim, moments, ray-cast), [2] scan cells (gather, group mask,
 * spare-slot wipe), [3], [4] unused,
 * [5] merge codes, [6] merge cells, [7] 2-D maps, [8] D2H, [9] host time of the last
 * pageable->pinned staging copy.  [0..8] are only recorded when
 * gvom_set_profiling(h, 1) is on (adds event records to the stream). */
int gvom_set_profiling(GvomHandle* h, int32_t on);
int gvom_stage_times(GvomHandle* h, float ms[16]);

/* Tooling: A/B switches for measurements (bit mask, see gvom_api.cu VAR_*; 0 = defaults); every setting gives the
 * same results.  Also read from the environment variable GVOM_VARIANT at gvom_create(). */
int gvom_set_variant(GvomHandle* h, uint32_t mask);

/* Tooling: what would capturing the tick in a CUDA graph buy?  Captures the launches of one Process_pointcloud
 * (device cloud) + combine_maps (outputs left in the library's block) including their programmatic-dependent-launch
 * edges, replays the graph `iters` times and then issues the same tick `iters` times as plain launches; CUDA-event
 * time per tick of both.  The handle must not be used for mapping afterwards (the replay re-runs the same kernels on
 * the same buffers). */
int gvom_graph_probe(GvomHandle* h, const void* points_dev, int64_t n, int32_t stride, int32_t dtype,
                     const double ego[3], const double* transform16, int32_t iters, float* graph_ms_per_tick,
                     float* launch_ms_per_tick, int32_t* graph_nodes);

/* Tooling: L2 atomic-throughput microbenchmark (denominator of the ray-cast roofline).
 * Launches sm_count*8 blocks of 256 threads, each thread issuing per_thread
 * red.global.add.u32 to pseudo-random words of table_dev (mode 0; table_words a power
 * of two, caller-owned device memory) or to one single word (mode 1).  Best-of-repeats
 * CUDA-event time in ms. */
int gvom_bench_atomics(int device, void* table_dev, int64_t table_words, int32_t per_thread,
                       int32_t mode, int32_t repeats, float* best_ms, int64_t* atomics_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* GVOM_B200_H */
