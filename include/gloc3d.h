/*
 * gloc3d.h -- C ABI of libgloc3d.so: the B200 (sm_100a) implementation of
 * GLoc3D's global-localization query path.  Plain pointers and sizes only; no
 * C++ or torch types cross this boundary.  Every entry point returns an int
 * status (GLOC_OK == 0) and never throws; gloc_last_error() gives the message
 * of the last failure on the calling thread.
 *
 * Each entry point cites the reference interface it replaces (paths relative
 * to /root/reference/registration/).  The C++ shims in gloc3d_b200/host/ put
 * the reference's own class interfaces (InvKeyTree, FastCorrelativeScanMatcher2D)
 * on top of these functions; INTEGRATION.md shows the binding a maintainer adds.
 *
 * There is NO CPU fallback: every compute entry point fails with
 * GLOC_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef GLOC3D_H_
#define GLOC3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLOC_OK 0
#define GLOC_ERR_INVALID 1   /* bad argument (null pointer, k == 0, dim mismatch ...) */
#define GLOC_ERR_CUDA 2      /* CUDA runtime / driver failure, or no usable device   */
#define GLOC_ERR_NOT_BUILT 3 /* query on an empty index (nanoflann: runtime_error,    */
                             /* nanoflann.hpp:1454-1457)                              */
#define GLOC_ERR_RANGE 4     /* parameter outside the supported range                 */
#define GLOC_ERR_NOMEM 5

int gloc_version(void);
const char* gloc_last_error(void);
/* Number of CUDA devices of compute capability 10.x visible to the process. */
int gloc_device_count(void);

/* ===================================================================== stage 1
 * Exhaustive exact top-k L2 retrieval.  Replaces
 *   InvKeyTree = KDTreeVectorOfVectorsAdaptor<KeyMat, float>
 *     ctor  (size_t dim, const KeyMat& mat, int leaf_max_size)   KDTreeVectorOfVectorsAdaptor.h:70-84
 *     query (const float* q, size_t k, size_t* idx, float* d2)   KDTreeVectorOfVectorsAdaptor.h:95-102
 * as used by RpyPCLoopDetector::detect (loop_detector.cpp:34-37,42-45,66-79).
 * Results: the k rows with the smallest squared L2 distance, where the distance
 * is computed in float32 in the operation order of L2_Adaptor::evalMetric
 * (nanoflann.hpp:453-487: groups of 4, left-associated, no FMA), ascending by
 * (d2, idx).  Indices and distances are bit-exact with nanoflann's exact search
 * (ties: nanoflann's order is traversal-dependent, this library's is (d2, idx)).
 */
typedef struct gloc_knn_index gloc_knn_index;

/* Search strategy.  All three return identical results. */
#define GLOC_KNN_AUTO 0      /* tensor shortlist when it applies, else exact scan   */
#define GLOC_KNN_EXACT_SCAN 1 /* FP32 exact scan of every row (K3 as a full scan)    */
#define GLOC_KNN_SHORTLIST 2 /* FP16 tcgen05 GEMM shortlist (K1+K2) + FP32 re-rank  */
                             /* (K3); queries whose shortlist overflows are re-run  */
                             /* through the exact scan on the GPU.  Limits: dim % 64 */
                             /* == 0, dim <= 512, k <= 32 (GLOC_ERR_RANGE beyond;    */
                             /* GLOC_KNN_AUTO takes the exact scan there, k <= 128)  */

typedef struct {
  uint64_t queries;            /* queries answered since creation                   */
  uint64_t kernel_launches;    /* CUDA kernels launched by this index               */
  uint64_t shortlist_queries;  /* queries answered through the tensor shortlist     */
  uint64_t fallback_queries;   /* shortlist overflowed -> exact scan on the GPU     */
  uint64_t shortlist_rows;     /* rows re-ranked in FP32 (sum over queries)         */
  uint64_t last_mode;          /* strategy used by the last call (GLOC_KNN_*)       */
} gloc_knn_stats;

int gloc_knn_create(gloc_knn_index** out, size_t dim, int device);
void gloc_knn_destroy(gloc_knn_index* index);

/* Replace the database with n rows (row-major n x dim float32, host memory).
 * Copy semantics: the reference adaptor keeps a const& to the caller's KeyMat
 * (KDTreeVectorOfVectorsAdaptor.h:88) which must stay unmodified, so a copy is
 * observationally identical.  n == 0 empties the index. */
int gloc_knn_set_db(gloc_knn_index* index, const float* rows, size_t n);
/* Same with rows already in device memory of the index's device. */
int gloc_knn_set_db_device(gloc_knn_index* index, const float* d_rows, size_t n);
/* Append rows (SLAM mode grows the DB one keyframe at a time,
 * loop_detector.cpp:9-20); amortised O(1) re-allocation. */
int gloc_knn_append(gloc_knn_index* index, const float* rows, size_t n);
size_t gloc_knn_size(const gloc_knn_index* index);
size_t gloc_knn_dim(const gloc_knn_index* index);

/* Search only rows [0, n_search) -- SLAM mode searches all but the most recent
 * NUM_EXCLUDE_RECENT=30 keyframes (loop_detector.cpp:66-72).  SIZE_MAX = all. */
int gloc_knn_set_search_limit(gloc_knn_index* index, size_t n_search);
/* Returned index = local row + offset (database sharded over GPUs). */
int gloc_knn_set_index_offset(gloc_knn_index* index, uint64_t offset);
int gloc_knn_set_mode(gloc_knn_index* index, int mode);
int gloc_knn_get_stats(const gloc_knn_index* index, gloc_knn_stats* stats);

/* Live kernel timing for bench.py's roofline line: when enabled, the dominant kernel of
 * every query call (the exact scan, or the shortlist GEMM) is bracketed by CUDA events on
 * the launching stream.  get_profile waits for them, returns their summed duration and
 * count since the last call, and resets. */
typedef struct {
  double dominant_ms;
  uint64_t dominant_launches;
} gloc_profile;
int gloc_knn_set_profiling(gloc_knn_index* index, int enabled);
int gloc_knn_get_profile(gloc_knn_index* index, gloc_profile* out);

/* nq queries (row-major nq x dim) -> out_idx / out_d2 (row-major nq x k), all in
 * HOST memory (caller-allocated, as loop_detector.cpp:42-43).  Slots beyond the
 * number of searchable rows get idx = UINT64_MAX, d2 = FLT_MAX.  Host<->device
 * copies happen inside the call. */
int gloc_knn_query(gloc_knn_index* index, const float* queries, size_t nq, size_t k,
                   uint64_t* out_idx, float* out_d2);
/* Experimental: GLOC_KNN_PAIR=1 in the environment selects the CTA-pair (cta_group::2) form of
 * the shortlist GEMM.  Returns the number of resident CTA pairs it would
 * run with on `device`, 0 when the variable is unset or the variant cannot be launched. */
int gloc_knn_pair_workers(int device);

/* Same with queries and outputs in DEVICE memory; work is enqueued on `stream`
 * (a cudaStream_t, NULL = the legacy default stream) and not synchronised. */
int gloc_knn_query_device(gloc_knn_index* index, const float* d_queries, size_t nq,
                          size_t k, uint64_t* d_out_idx, float* d_out_d2, void* stream);

/* K4: merge g per-shard top-k lists (device memory, laid out [g][nq][k], each
 * ascending by (d2, idx) with GLOBAL indices, UINT64_MAX = empty slot) into the
 * global top-k -- run on every rank after the NCCL all-gather. */
int gloc_knn_merge_topk_device(const uint64_t* d_idx, const float* d_d2, size_t g,
                               size_t nq, size_t k, uint64_t* d_out_idx,
                               float* d_out_d2, int device, void* stream);

/* ===================================================================== stage 2
 * Correlative / branch-and-bound scan matching of BEV probability grids.
 * Replaces cartographer::mapping::scan_matching::FastCorrelativeScanMatcher2D
 * (2d/fast_correlative_scan_matcher_2d.h:137-200) and the pieces it is built
 * from (PrecomputationGridStack2D, SearchParameters, GenerateRotatedScans,
 * DiscretizeScans, ScoreCandidates, BranchAndBound).  A store holds many map
 * grids (one per database keyframe); a match scores one scan against one grid.
 * The selected discrete pose is the global maximum over every (scan, x, y) of
 * the shrunk search window -- what BranchAndBound returns -- with ties
 * resolved to the smallest (scan, x, y).
 */
/* What a store keeps per map grid is the width-1 precomputation grid only, bit-packed (one bit
 * per cell) when every cell is 0 or 255 -- 49 KB for a KITTI-sized 781 x 504 BEV grid -- in a
 * pooled device arena.  The reference builds a PrecomputationGridStack2D per matcher instance
 * (fast_..._2d.cpp:192-215); here the coarser levels are rebuilt on the device for the distinct
 * grids of each batch, so that a database of a million frames stays resident. */
typedef struct gloc_csm_store gloc_csm_store;

typedef struct {
  int found;        /* best score > min_score                  fast_..._2d.cpp:311  */
  float score;      /* PrecomputationGrid2D::ToScore of the best candidate          */
  int scan_index;   /* Candidate2D::scan_index                 correlative_..._2d.h:90 */
  int x_offset;     /* Candidate2D::x_index_offset                                  */
  int y_offset;     /* Candidate2D::y_index_offset                                  */
  int reserved;
  double pose_x;    /* initial x + (-y_offset * resolution)    correlative_..._2d.h:81-84 */
  double pose_y;    /* initial y + (-x_offset * resolution)                         */
  double pose_yaw;  /* initial yaw + (scan_index - n_ang)*step fast_..._2d.cpp:313-317 */
} gloc_csm_result;

typedef struct {
  uint64_t matches;          /* (grid, scan) pairs matched                           */
  uint64_t kernel_launches;
  uint64_t coarse_candidates; /* candidates scored on the coarsest grid              */
  uint64_t refined_nodes;     /* branch-and-bound nodes expanded below it            */
} gloc_csm_stats;

int gloc_csm_create(gloc_csm_store** out, int device);
void gloc_csm_destroy(gloc_csm_store* store);

/* Add a map grid given as Grid2D's uint16 correspondence-cost cells
 * (2d/grid_2d.h:100-101, flat index num_x_cells*y + x, 0 = unknown) with its
 * MapLimits (2d/map_limits.h:40-47).  The width-1 precomputation grid is
 * derived on the GPU exactly as PrecomputationGrid2D does
 * (fast_..._2d.cpp:118-119,130-131,184-190). */
int gloc_csm_add_grid_cells(gloc_csm_store* store, const uint16_t* cells, int nx, int ny,
                            double resolution, double max_x, double max_y, int* grid_id);
/* Add a map grid given directly as the uint8 width-1 precomputation grid
 * (0 = min score ... 255 = max score). */
int gloc_csm_add_grid_u8(gloc_csm_store* store, const uint8_t* level1, int nx, int ny,
                         double resolution, double max_x, double max_y, int* grid_id);
int gloc_csm_num_grids(const gloc_csm_store* store);
/* Device memory held by the store: the grids themselves, and its (reused) work buffers. */
int gloc_csm_store_bytes(const gloc_csm_store* store, uint64_t* grid_bytes, uint64_t* workspace_bytes);
/* Copy the width-`width` precomputation grid of grid_id back to the host
 * ((nx+width-1)*(ny+width-1) bytes) -- PrecomputationGridStack2D::Get. */
int gloc_csm_get_precomputation_grid(gloc_csm_store* store, int grid_id, int width,
                                     uint8_t* out);

/* Match n_pairs independent (grid, scan) pairs.
 *   pts          concatenated scan points, float32 (x, y, z) triples (sensor::PointCloud)
 *   scan_offsets n_scans+1 prefix offsets (in points) into pts
 *   grid_ids, scan_ids, init_xyyaw  per pair: map grid, scan, initial pose estimate
 *   n_lin, n_ang, ang_step  SearchParameters "for testing" ctor
 *                 (correlative_scan_matcher_2d.cpp:57-71); use
 *                 gloc_csm_search_params for the production ctor
 *   depth        branch_and_bound_depth (fast_..._2d.h:51); only affects speed
 *   min_score    Match*'s min_score
 * = MatchWithSearchParameters (fast_..._2d.cpp:270-320) per pair. */
int gloc_csm_match_batch(gloc_csm_store* store, const float* pts, const int64_t* scan_offsets,
                         int n_scans, const int* grid_ids, const int* scan_ids,
                         const double* init_xyyaw, int n_pairs, int n_lin, int n_ang,
                         double ang_step, int depth, float min_score,
                         gloc_csm_result* results);

/* K6 on its own: the rotated + discretised scans MatchWithSearchParameters builds
 * (fast_..._2d.cpp:278-289 = TransformPointCloud by the float initial yaw,
 * GenerateRotatedScans correlative_scan_matcher_2d.cpp:93-109, DiscretizeScans
 * :111-127).  out_cells is (2*n_ang+1) x n_pts x 2 int32 (cell x, cell y), host. */
int gloc_csm_discretize(gloc_csm_store* store, const float* pts, int n_pts, double init_x,
                        double init_y, double init_yaw, int n_ang, double ang_step,
                        double resolution, double max_x, double max_y, int32_t* out_cells);

/* SearchParameters production ctor (correlative_scan_matcher_2d.cpp:27-55). */
int gloc_csm_search_params(double linear_window, double angular_window, const float* pts,
                           int n_pts, double resolution, int* n_lin, int* n_ang,
                           double* ang_step);
/* GridToVirtualPointCloud (fast_..._2d.cpp:78-95): cells with cost < 0.11 ->
 * (ox + i*res, oy + j*res, 0).  Returns the count through n_out; pts may be
 * NULL to size the buffer. */
int gloc_csm_grid_to_points(const uint16_t* cells, int nx, int ny, double resolution,
                            double ox, double oy, float* pts, int capacity, int* n_out);
int gloc_csm_get_stats(const gloc_csm_store* store, gloc_csm_stats* stats);
/* Same live timing for stage 2; the dominant kernel is the coarse-level scorer. */
int gloc_csm_set_profiling(gloc_csm_store* store, int enabled);
int gloc_csm_get_profile(gloc_csm_store* store, gloc_profile* out);

/* ============================================================ multi-GPU
 * The database shards by rows over the GPUs of one box (the reference has no
 * distributed code, F1): every GPU holds the descriptors of rows [offset, offset + n)
 * (gloc_knn_set_index_offset) and the map grids of those rows.  The communicator is an NCCL
 * communicator owned by this library (resolved with dlopen at first use), so a C++ host can
 * drive several GPUs without Python:
 *   one process per GPU   rank 0: gloc_comm_unique_id -> hand the 128 bytes to the other ranks
 *                         (any channel) -> every rank: gloc_comm_create
 *   one process, n GPUs   gloc_comm_create_local (one communicator per device; the collective
 *                         entry points are then called from one thread per device)
 */
typedef struct gloc_comm gloc_comm;
#define GLOC_COMM_ID_BYTES 128
int gloc_comm_unique_id(uint8_t* id, size_t capacity);
int gloc_comm_create(gloc_comm** out, const uint8_t* id, int n_ranks, int rank, int device);
int gloc_comm_create_local(gloc_comm** out /* n_devices handles */, int n_devices, const int* devices);
void gloc_comm_destroy(gloc_comm* comm);
int gloc_comm_rank(const gloc_comm* comm);
int gloc_comm_size(const gloc_comm* comm);
int gloc_comm_nccl_version(void);   /* 0 when NCCL cannot be loaded */

/* Row-sharded exact top-k (BASELINE configs[3]); collective, same k on every rank.
 *   replicated == 0   `queries` is THIS RANK'S SLICE of the batch (nq queries; the same count on
 *                     every rank): queries are all-gathered over NVLink, every rank searches the
 *                     whole batch on its shard, the local top-k lists go to the rank that owns the
 *                     query (all-to-all), which merges them.  Output: this rank's slice, nq x k.
 *   replicated != 0   every rank passes the same nq queries (online localisation: one query);
 *                     local search, all-gather of the lists, merge everywhere.  Output: nq x k.
 * Same result as InvKeyTree::query on the whole database (KDTreeVectorOfVectorsAdaptor.h:95-102):
 * ascending (d2, idx) with global row indices. */
int gloc_knn_query_sharded(gloc_knn_index* shard, gloc_comm* comm, const float* queries, size_t nq,
                           size_t k, uint64_t* out_idx, float* out_d2, int replicated);
int gloc_knn_query_sharded_device(gloc_knn_index* shard, gloc_comm* comm, const float* d_queries,
                                  size_t nq, size_t k, uint64_t* d_out_idx, float* d_out_d2,
                                  int replicated, void* stream);

/* ============================================================ whole query path
 * One call from descriptors + scans to located frames and poses.  Replaces the evaluation loop
 * of the reference driver:
 *   GlocEvaluator::detect_all_query    global_localization.cpp:482-509 (RpyPCLoopDetector::detect,
 *                                      loop_detector.cpp:22-46 -> InvKeyTree::query, top_k_)
 *   GlocEvaluator::global_registraion  global_localization.cpp:511-574 (candidates in retrieval
 *                                      order; loop_detector_.match(q_grid, db_idx, ...) :519-524;
 *                                      the first candidate that matches is the located frame)
 * with FastCorrelativeScanMatcher2D::MatchWithSearchParameters as the verifier.  Retrieval, the
 * gather of the candidates' map grids, verification and the per-query choice run on the device
 * without a host round trip between the stages.  The index and the store are borrowed (they must
 * outlive the localizer and live on the same device); database row r is verified against grid
 * grid_of_row[r] (default: grid r -- db_grids_[db_idx], loop_detector.h:36-39).
 */
typedef struct gloc_localizer gloc_localizer;

#define GLOC_LOC_VERIFY_ALL 0   /* every one of the k candidates is verified (BASELINE config:   */
                                /* "25 candidates/query"); per-candidate results available       */
#define GLOC_LOC_FIRST_MATCH 1  /* the reference's evaluation order: candidate c is verified only */
                                /* if candidates 0..c-1 did not match; same located frame and pose */

typedef struct {
  int k;              /* candidates per query: top_k_ (loop_detector.h:98; 20 there, 25 in BASELINE) */
  int n_lin, n_ang;   /* SearchParameters "for testing" ctor, as gloc_csm_match_batch               */
  double ang_step;
  int depth;          /* branch_and_bound_depth; only affects speed                                 */
  float min_score;
  int policy;         /* GLOC_LOC_*                                                                 */
} gloc_loc_params;

typedef struct {
  int located;            /* global_registraion's return value                                     */
  int candidate;          /* position in the top-k of the located frame = the first candidate, in  */
                          /* retrieval order, whose match succeeded; -1 if none                    */
  int best_candidate;     /* the candidate with the highest score among those verified; -1 if none */
  int n_verified;         /* candidates verified for this query                                    */
  uint64_t db_index;      /* located_db_idx (UINT64_MAX if not located)                            */
  gloc_csm_result match;  /* the match of `candidate`: pose of the query in the located frame      */
} gloc_loc_result;

typedef struct {
  uint64_t queries, pairs_verified, waves, kernel_launches;
  uint64_t pairs_migrated;   /* of pairs_verified: verified here on a peer's grid (gloc_loc_share_grids) */
} gloc_loc_stats;

int gloc_loc_create(gloc_localizer** out, gloc_knn_index* knn, gloc_csm_store* csm);
void gloc_loc_destroy(gloc_localizer* loc);
/* Row -> grid table (n_rows >= rows of the index; ids must exist in the store).  NULL: identity. */
int gloc_loc_set_row_grids(gloc_localizer* loc, const int32_t* grid_of_row, size_t n_rows);
/* nq queries (row-major nq x dim) with one scan each (pts: concatenated float32 xyz triples,
 * scan_offsets: nq+1 prefix offsets in points) and an initial pose estimate per query
 * (init_xyyaw: nq x 3, NULL = zeros), all in HOST memory.  Outputs (host, caller-allocated):
 *   out_idx / out_d2  nq x k    the retrieval result, as gloc_knn_query
 *   cand_results      nq x k    per-candidate match (may be NULL); reserved = -1 marks candidates
 *                               GLOC_LOC_FIRST_MATCH did not evaluate
 *   results           nq        located frame and pose per query                                  */
int gloc_loc_localize(gloc_localizer* loc, const float* queries, size_t nq, const float* pts,
                      const int64_t* scan_offsets, const double* init_xyyaw,
                      const gloc_loc_params* params, uint64_t* out_idx, float* out_d2,
                      gloc_csm_result* cand_results, gloc_loc_result* results);
/* Same with the descriptors and the scan points already in DEVICE memory (scan offsets, initial
 * poses and all outputs stay on the host: a few hundred bytes per query). */
int gloc_loc_localize_device(gloc_localizer* loc, const float* d_queries, size_t nq, const float* d_pts,
                             const int64_t* scan_offsets, const double* init_xyyaw,
                             const gloc_loc_params* params, uint64_t* out_idx, float* out_d2,
                             gloc_csm_result* cand_results, gloc_loc_result* results);
/* The same over a row-sharded database: the localizer's index and store hold this rank's rows
 * and their grids (grid_of_row is indexed by LOCAL row).  Collective; every rank passes the SAME
 * batch and gets the same outputs.  Retrieval as gloc_knn_query_sharded (replicated); a (query,
 * candidate) pair is verified on the rank that owns the candidate's row; one all-reduce of the
 * 8-byte pair results per wave.  All grids of the map must share one resolution.
 * Host buffers (buffers_on_device == 0): every rank uploads the queries and scans of ITS 1/N of the
 * queries only, the parts reach the peers over NVLink (one PCIe upload per byte, not one per rank).
 * buffers_on_device != 0: `queries` and `pts` are device pointers (as gloc_loc_localize_device). */
int gloc_loc_localize_sharded(gloc_localizer* loc, gloc_comm* comm, const float* queries, size_t nq,
                              const float* pts, const int64_t* scan_offsets, const double* init_xyyaw,
                              const gloc_loc_params* params, uint64_t* out_idx, float* out_d2,
                              gloc_csm_result* cand_results, gloc_loc_result* results,
                              int buffers_on_device);
/* Collective, optional, between building the shards and localizing: every rank makes its grid store
 * addressable by its peers (NVLink peer memory: direct pointers inside one process, CUDA IPC between
 * processes) and learns the peers' grid tables.  From then on gloc_loc_localize_sharded balances the
 * work: ranks that own more than their share of a wave's (query, candidate) pairs hand the excess to
 * ranks that own less, which build the working set of such a pair straight from the owner's
 * bit-packed grid over NVLink (83 KB per 800 x 800 grid; nothing is copied ahead of time).  Results
 * are unchanged.  Call again after any rank added grids or changed its row table; returns
 * GLOC_ERR_CUDA (and leaves owner-only verification in place) where peers cannot address each other.
 * gloc_loc_unshare_grids (collective, same communicator) unmaps the peers' stores. */
int gloc_loc_share_grids(gloc_localizer* loc, gloc_comm* comm);
/* The assignment rule itself (host only, no device needed; what gloc_loc_localize_sharded applies per
 * wave once grids are shared): owner[i] = the rank holding pair i's row (or -1), pairs in (query,
 * candidate) order; verifier[i] = the rank that verifies it.  A rank keeps the first
 * ceil(pairs / n_ranks) of its own pairs, the surplus goes in order to the ranks with room. */
int gloc_loc_assign_pairs(const int32_t* owner, size_t n_pairs, int n_ranks, int32_t* verifier);
int gloc_loc_unshare_grids(gloc_localizer* loc, gloc_comm* comm);
int gloc_loc_get_stats(const gloc_localizer* loc, gloc_loc_stats* out);
/* Live device-side timing (CUDA events on the stream the work is launched on) of whole calls and
 * of their retrieval stage; the rest of a call is verification.  Summed since the last get. */
typedef struct {
  double total_ms, retrieval_ms;
  uint64_t calls;
} gloc_loc_profile;
int gloc_loc_set_profiling(gloc_localizer* loc, int enabled);
int gloc_loc_get_profile(gloc_localizer* loc, gloc_loc_profile* out);

/* ============================================================ BEV projection
 * (the producer of both stages' inputs.)  One LiDAR scan ->
 * the reference's bird's-eye-view occupancy image.  Replaces
 *   RpyPCLoopDetector::get_projected_grid   loop_detector.cpp:122-135
 *     = point_cloud_to_range_data           loop_detector.cpp:108-120
 *     + Submap3D::InsertRangeData           3d/submap_3d.cpp:162-177 (fresh submap, identity pose)
 *     + ProjectToCvMat                      3d/submap_3d.cpp:238-326
 *   RpyPCLoopDetector::crop_pad_occupancy   loop_detector.cpp:83-106
 *   ProjectToGrid                           3d/submap_3d.cpp:328-429 (gloc_csm_add_grid_from_bev)
 * A pixel is occupied (0; free = 255) iff at least two distinct 0.2 m voxels of
 * its z column were hit by returns within max_range; the image spans the bounding
 * box of all hit voxels; (ox, oy) = world coordinates of pixel (0, 0).
 */
typedef struct gloc_bev_projector gloc_bev_projector;

typedef struct {
  int width, height;          /* cv::Mat cols, rows                                   */
  int min_ix, min_iy;         /* voxel index of pixel (0, 0)                          */
  double ox, oy, resolution;  /* ProjectToCvMat's out-parameters                      */
  uint64_t n_occupied;        /* pixels with value 0                                  */
  uint64_t n_points_in_range; /* returns (range <= max_range)                         */
} gloc_bev_info;

/* resolution = high_resolution_ (0.2), max_range = high_resolution_max_range_ (100)
 * in the reference (loop_detector.h:111-116). */
int gloc_bev_create(gloc_bev_projector** out, int device, float resolution, float max_range);
void gloc_bev_destroy(gloc_bev_projector* bev);
/* pts: n_pts points in HOST memory, `stride` floats apart, x y z first (KITTI
 * .bin scans: stride 4).  The image stays on the device until fetched. */
int gloc_bev_project(gloc_bev_projector* bev, const float* pts, size_t n_pts, int stride,
                     gloc_bev_info* info);
/* The image of the last projection, height*width bytes, row-major (host). */
int gloc_bev_get_image(gloc_bev_projector* bev, uint8_t* img, size_t capacity);
/* crop_pad_occupancy: centre crop / 255-pad to width x height (768 x 768 in the
 * reference, loop_detector.cpp:144), one channel. */
int gloc_bev_get_cnn_input(gloc_bev_projector* bev, int width, int height, uint8_t* out);
/* The same plus roi_dst of crop_pad_occupancy (loop_detector.cpp:99-102): x0, y0, width, height
 * of the copied image inside the plane -- everything outside is the canvas's padding. */
int gloc_bev_get_cnn_input_roi(gloc_bev_projector* bev, int width, int height, uint8_t* out, int32_t roi[4]);
/* Occupied pixels as GridToVirtualPointCloud produces them from ProjectToGrid's
 * grid (fast_..._2d.cpp:78-95): (ox + i*res, oy + j*res, 0), i outer.  pts may be
 * NULL to query the count. */
int gloc_bev_get_occupied_points(gloc_bev_projector* bev, float* pts, size_t capacity,
                                 size_t* n_out);
uint64_t gloc_bev_kernel_launches(const gloc_bev_projector* bev);
/* Adds the last projection to a scan-match store as ProjectToGrid builds it:
 * num_x_cells = width, num_y_cells = height, max = (max_ix, max_iy) * resolution,
 * occupied -> cost 0.1 (level-1 value 255), free -> cost 0.9 (0); device to device. */
int gloc_csm_add_grid_from_bev(gloc_csm_store* store, gloc_bev_projector* bev, int* grid_id);
/* Same occupancy, laid out so that MapLimits::GetCellIndex (2d/map_limits.h:69-76) of a world
 * point lands on the pixel of its own voxel: num_x_cells = height, num_y_cells = width, cell
 * (cx, cy) = pixel (ix, iy) = (width-1-cy, height-1-cx), max = (max_ix + 0.5, max_iy + 0.5) *
 * resolution.  ProjectToGrid stores pixel (ix, iy) at cell (ix, iy) while the matcher looks
 * points up through GetCellIndex, which swaps and flips the axes -- a lookup then hits the
 * mirror image and no rigid transform can align two such grids.  This entry
 * point is the consistent construction the driver uses for the north-star verifier. */
int gloc_csm_add_grid_from_bev_aligned(gloc_csm_store* store, gloc_bev_projector* bev, int* grid_id);

/* ============================================================ descriptor head
 * (last step of descriptor extraction.)  The NetVLAD_fc pooling layer
 * that turns the encoder's feature map into the 512-d place descriptor stage 1 searches:
 *   NetVLAD.forward   model/netvlad_fc.py:73-109 (vladv2 = False, gating off), as traced into
 *   the TorchScript module of RpyPCLoopDetector::get_place_feature (loop_detector.cpp:137-172)
 * Batched, device to device: descriptors can go straight into gloc_knn_query_device /
 * gloc_knn_set_db_device.  FP32; matches the reference module within 1e-5 of the largest
 * output component (summation order differs).  The VGG16 encoder is NOT part of this library.
 *   conv_w    [clusters][dim]            1x1 conv weight (netvlad_fc.py:34)
 *   conv_b    [clusters] or NULL         its bias (NULL for vladv2 = False)
 *   centroids [clusters][dim]            (:35)
 *   hidden_w  [clusters*dim][out_dim]    hidden1_weights (:37-38; out_dim = dim in the reference)
 * Limits: dim % 32 == 0, clusters <= 64, clusters * dim <= 51200. */
typedef struct gloc_vlad_head gloc_vlad_head;
int gloc_vlad_create(gloc_vlad_head** out, int device, int dim, int clusters, int out_dim,
                     const float* conv_w, const float* conv_b, const float* centroids,
                     const float* hidden_w);
void gloc_vlad_destroy(gloc_vlad_head* head);
/* feat: [batch][dim][n_loc] float32 (the encoder's NCHW output with H*W = n_loc), out:
 * [batch][out_dim]; both in DEVICE memory of the head's device. */
int gloc_vlad_forward_device(gloc_vlad_head* head, const float* d_feat, int batch, int n_loc,
                             float* d_out);
/* The same with HOST buffers (copies inside). */
int gloc_vlad_forward(gloc_vlad_head* head, const float* feat, int batch, int n_loc, float* out);
uint64_t gloc_vlad_kernel_launches(const gloc_vlad_head* head);

/* ============================================================ descriptor encoder
 * (first step of descriptor extraction.)  VGG16 features[:-2] as the
 * reference assembles it (main.py:531-536: 13 3x3 convolutions + ReLU, the first four 2x2
 * max-pools, last ReLU and pool dropped) on the BEV occupancy image that
 * RpyPCLoopDetector::get_place_feature feeds it (loop_detector.cpp:137-172: 768 x 768, three
 * identical channels, 1/255).  Batched, device to device; output [batch][512][H/16 * W/16]
 * float32 is what gloc_vlad_forward_device takes.  Tensor-core implicit GEMM (tcgen05, FP16
 * operands, FP32 accumulation -- the significand of the TF32 path the reference's cuDNN uses).
 *   conv_w[l]  [Cout][Cin][3][3] float32, l = 0..12 (torchvision's layout; Cin = 3 for l = 0)
 *   conv_b[l]  [Cout]
 * images: uint8 [batch][height][width] (one plane: the three channels are identical; the plane
 * gloc_bev_get_cnn_input returns).  height % 128 == 0, width % 256 == 0. */
typedef struct gloc_encoder gloc_encoder;
int gloc_enc_create(gloc_encoder** out, int device, int height, int width,
                    const float* const* conv_w, const float* const* conv_b);
void gloc_enc_destroy(gloc_encoder* enc);
int gloc_enc_feature_shape(const gloc_encoder* enc, int* channels, int* n_loc);
int gloc_enc_forward_device(gloc_encoder* enc, const uint8_t* d_images, int batch, float* d_feat);
/* The same with HOST buffers (copies inside). */
int gloc_enc_forward(gloc_encoder* enc, const uint8_t* images, int batch, float* feat);
uint64_t gloc_enc_kernel_launches(const gloc_encoder* enc);

/* The reference pads small BEV images onto its 768 x 768 canvas with cv::Mat::ones(h, w,
 * CV_8UC3) * 255 (loop_detector.cpp:84); Mat::ones sets only channel 0 of a multi-channel matrix,
 * so the padding is (255, 0, 0) while the copied image has three identical channels.  The padded
 * entry points take, per image, the rectangle of the copied image inside the plane (rois:
 * [batch][4] int32 = x0, y0, width, height, HOST memory; gloc_bev_get_cnn_input_roi returns it)
 * and give the padding the reference's channel-0-only treatment in conv1_1.  rois == NULL, or the
 * plain entry points above: every pixel of the plane is image. */
int gloc_enc_forward_padded_device(gloc_encoder* enc, const uint8_t* d_images, const int32_t* rois,
                                   int batch, float* d_feat);
int gloc_enc_forward_padded(gloc_encoder* enc, const uint8_t* images, const int32_t* rois, int batch,
                            float* feat);
int gloc_desc_extract_padded(gloc_encoder* enc, gloc_vlad_head* head, int out_dim, const uint8_t* images,
                             const int32_t* rois, int batch, float* desc);

/* Host convenience over both halves: uint8 planes [batch][H][W] (host) -> descriptors
 * [batch][out_dim] (host); the feature maps stay on the device.  out_dim is the head's. */
int gloc_desc_extract(gloc_encoder* enc, gloc_vlad_head* head, int out_dim, const uint8_t* images,
                      int batch, float* desc);

/* ============================================================ grid store file
 * (a map's BEV grids on disk, so that a database is projected once.)
 * The reference keeps its grids in memory only (db_grids_, loop_detector.h:36-39) and
 * re-projects every scan at start-up (global_localization.cpp:419-449); this is the
 * file that replaces that pass.  Little-endian:
 *   header   "GLOCGRD1" | u32 version = 1 | u32 tag | u64 n_grids     (tag: the writer's fingerprint
 *            of what the grids were made from, 0 = none; a reader compares it with its own)
 *   per grid i32 nx | i32 ny | f64 resolution | f64 max_x | f64 max_y | u32 encoding | u32 0 |
 *            u64 payload_bytes | payload
 *   encoding 1  bit-packed binary grid: bit (i & 7) of byte (i >> 3) is set iff level-1 cell
 *               i = nx*y + x is 255 (occupied), clear iff it is 0; ceil(nx*ny / 8) bytes
 *   encoding 0  the uint8 level-1 grid as it is, nx*ny bytes (grids with other values)
 * The file functions run on the host only (no device needed). */
typedef struct {
  int32_t nx, ny;              /* MapLimits cell counts                                */
  double resolution, max_x, max_y;
} gloc_grid_info;

typedef struct gloc_grid_file gloc_grid_file;

/* MapLimits of a grid of the store (host metadata). */
int gloc_csm_get_grid_info(const gloc_csm_store* store, int grid_id, gloc_grid_info* out);
/* Writes n grids; level1[i] is the nx*ny uint8 width-1 precomputation grid of grid i
 * (what gloc_csm_add_grid_u8 takes / gloc_csm_get_precomputation_grid(width = 1) returns). */
int gloc_grid_file_write(const char* path, const gloc_grid_info* infos,
                         const uint8_t* const* level1, size_t n);
int gloc_grid_file_write_tagged(const char* path, const gloc_grid_info* infos,
                                const uint8_t* const* level1, size_t n, uint32_t tag);
int gloc_grid_file_open(const char* path, gloc_grid_file** out, size_t* n_grids);
uint32_t gloc_grid_file_tag(const gloc_grid_file* f);
/* Reads the next grid: info always; the cells into level1 when capacity >= nx*ny, otherwise
 * the record is NOT consumed (call again with a large enough buffer).  GLOC_ERR_RANGE after
 * the last grid. */
int gloc_grid_file_next(gloc_grid_file* f, gloc_grid_info* info, uint8_t* level1, size_t capacity);
void gloc_grid_file_close(gloc_grid_file* f);
/* Every grid of the store -> file; file -> grids appended to the store (ids first_grid_id ..
 * first_grid_id + n_grids - 1, in file order). */
int gloc_csm_save_grids(gloc_csm_store* store, const char* path);
int gloc_csm_save_grids_tagged(gloc_csm_store* store, const char* path, uint32_t tag);
int gloc_csm_load_grids(gloc_csm_store* store, const char* path, int* first_grid_id, int* n_grids);

/* ============================================================ measured ceilings
 * Chip-wide rate of random 8-byte shared-memory loads (LDS.64 with the bank conflicts random
 * addresses bring): the unit the stage-2 coarse scorer is bound by (SURVEY.md 8d asks for its
 * gather rate against a measured ceiling).  Takes a few milliseconds of GPU time. */
int gloc_bench_smem_gather(int device, double* loads_per_s);

#ifdef __cplusplus
}
#endif
#endif /* GLOC3D_H_ */
