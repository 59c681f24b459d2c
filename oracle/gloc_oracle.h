/*
 * gloc_oracle.h -- CPU restatement of GLoc3D's global-localization query path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library, and there only as the
 * checker / CPU baseline.  The product (gloc3d_b200/) never links or calls it.
 *
 * Parity status:
 *   stage 1 (retrieval): pinned against the reference's own nanoflann +
 *     KDTreeVectorOfVectorsAdaptor compiled from /root/reference
 *     (oracle/_ref/libnanoflann_ref.so, see oracle/Makefile) and against the
 *     golden vectors minted from it (tests/golden/knn_*.npz).
 *   stage 2 (scan matching): pinned against the reference's own
 *     registration/2d/*.cpp (+ 3d/probability_values.cpp, 3d/point_cloud.cpp)
 *     compiled unmodified from /root/reference into oracle/_ref/libcsm_ref.so
 *     (oracle/csm_ref.cpp, against the stand-in headers of oracle/shim/ for the
 *     absent Eigen / glog / OpenCV / boost / ceres) -- tests/test_oracle_csm_ref.py,
 *     tests/golden/csm_*.npz.  Eigen's arithmetic inside the shim is a
 *     restatement of Eigen 3.3/3.4 (Eigen is not under /root/reference).  Also
 *     checked for self-consistency (B&B == exhaustive, level_w == sliding max
 *     of level_1, float path == u8 path).
 *
 * All file:line citations are relative to /root/reference/registration/.
 */
#ifndef GLOC_ORACLE_H_
#define GLOC_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ stage 1 */

/* L2_Adaptor::evalMetric, nanoflann.hpp:453-487, called with 3 arguments
 * (worst_dist = -1) so the early exit at :476 is dead.  float32, groups of 4,
 * left-associated, no FMA; tail of 0-3 components added one by one. */
float gloc_oracle_l2(const float* q, const float* x, size_t dim);

/* Exhaustive exact kNN: the k rows with the smallest gloc_oracle_l2, ascending
 * by (d2, idx).  Replaces KDTreeVectorOfVectorsAdaptor::query
 * (KDTreeVectorOfVectorsAdaptor.h:95-102) + KNNResultSet (nanoflann.hpp:160-236);
 * the tie-break (d2, idx) is the build's deterministic choice (reference order
 * among equal distances is KD-traversal order).  Slots beyond n rows get
 * idx = UINT64_MAX, d2 = FLT_MAX.  db is row-major n x dim. */
void gloc_oracle_knn(const float* db, size_t n, size_t dim, const float* q,
                     size_t nq, size_t k, uint64_t* out_idx, float* out_d2);

/* Same, queries partitioned over nthreads pthreads (CPU baseline, kind "port"). */
void gloc_oracle_knn_mt(const float* db, size_t n, size_t dim, const float* q,
                        size_t nq, size_t k, uint64_t* out_idx, float* out_d2,
                        int nthreads);

/* Merge G per-shard top-k lists (each ascending by (d2, global idx)) into the
 * global top-k -- the CPU statement of the multi-GPU merge (K4). lists are
 * laid out [g][nq][k]. */
void gloc_oracle_topk_merge(const uint64_t* idx, const float* d2, size_t g,
                            size_t nq, size_t k, uint64_t* out_idx,
                            float* out_d2);

/* ------------------------------------------------------------------ stage 2 */

/* 3d/probability_values.h:64-67 (float32 constants). */
float gloc_oracle_min_cost(void); /* kMinCorrespondenceCost = 1f-(1f-0.1f) */
float gloc_oracle_max_cost(void); /* kMaxCorrespondenceCost = 1f-0.1f      */

/* ValueToCorrespondenceCost, 3d/probability_values.cpp:27-36,59-63:
 * 0 -> kMax; v in [1,32767] -> v*kScale + (kMin - kScale); the table repeats
 * for values with the update marker (bit 15) set. */
float gloc_oracle_value_to_cost(uint16_t value);

/* CorrespondenceCostToValue, 3d/probability_values.h:32-44,77-80. */
uint16_t gloc_oracle_cost_to_value(float cost);

/* PrecomputationGrid2D::ComputeCellValue, 2d/fast_correlative_scan_matcher_2d.cpp:184-190
 * applied to 1 - |cost| (:130-131) with min/max score from :118-119. */
uint8_t gloc_oracle_cell_value(float probability);

/* Width-1 precomputation grid (uint8, nx*ny, flat index nx*y + x,
 * 2d/grid_2d.cpp:168-171) from Grid2D's uint16 correspondence-cost cells. */
void gloc_oracle_level1_from_cells(const uint16_t* cells, int nx, int ny,
                                   uint8_t* out);

/* PrecomputationGrid2D ctor, 2d/fast_correlative_scan_matcher_2d.cpp:112-182,
 * FLOAT path exactly as written (two 1-D sliding maxima over float
 * probabilities, then ComputeCellValue).  out has (nx+w-1)*(ny+w-1) cells,
 * stride nx+w-1, offset (-w+1,-w+1). */
void gloc_oracle_precomp_from_cells(const uint16_t* cells, int nx, int ny,
                                    int width, uint8_t* out);

/* Same grid computed as the sliding maximum of the uint8 width-1 grid (valid
 * because ComputeCellValue is monotone); this is the statement the GPU path
 * implements.  tests assert it equals the float path. */
void gloc_oracle_precomp_from_level1(const uint8_t* level1, int nx, int ny,
                                     int width, uint8_t* out);

/* SearchParameters production ctor, 2d/correlative_scan_matcher_2d.cpp:27-55.
 * pts is P x 3 float (x,y,z).  Outputs n_lin, n_ang, angular step. */
void gloc_oracle_search_params(double linear_window, double angular_window,
                               const float* pts, int n_pts, double resolution,
                               int* n_lin, int* n_ang, double* ang_step);

/* GridToVirtualPointCloud, 2d/fast_correlative_scan_matcher_2d.cpp:78-95:
 * every cell (i,j) with cost < 0.11 -> (ox + i*res, oy + j*res, 0) as float.
 * Iteration order i outer, j inner.  Returns the number of points; pts may be
 * NULL to count only (capacity in points). */
int gloc_oracle_grid_to_points(const uint16_t* cells, int nx, int ny,
                               double resolution, double ox, double oy,
                               float* pts, int capacity);

/* Rotated + discretised scans: MatchWithSearchParameters' prologue
 * (2d/fast_correlative_scan_matcher_2d.cpp:278-289) = TransformPointCloud by
 * the float initial yaw, GenerateRotatedScans (correlative_scan_matcher_2d.cpp:93-109),
 * DiscretizeScans (:111-127, MapLimits::GetCellIndex map_limits.h:69-76).
 * out_cells is S x P x 2 int32 (cell x, cell y), S = 2*n_ang+1. */
void gloc_oracle_discretize(const float* pts, int n_pts, double init_x,
                            double init_y, double init_yaw, int n_ang,
                            double ang_step, double resolution, double max_x,
                            double max_y, int32_t* out_cells);

typedef struct {
  int found;      /* best.score > min_score (fast_..._2d.cpp:311) */
  float score;    /* best fine score (min_score sentinel when !found) */
  int scan_index; /* Candidate2D::scan_index */
  int x_offset;   /* Candidate2D::x_index_offset */
  int y_offset;   /* Candidate2D::y_index_offset */
  double pose_x;  /* init.x + (-y_offset*res)      correlative_..._2d.h:81-84 */
  double pose_y;  /* init.y + (-x_offset*res) */
  double pose_yaw; /* init_yaw + (scan-n_ang)*step  fast_..._2d.cpp:313-317 */
  long long n_scored; /* candidates scored (work counter, not part of parity) */
} gloc_oracle_match_result;

/* mode 0: branch and bound exactly as the reference (DFS, fast_..._2d.cpp:393-438)
 *         with std::stable_sort semantics for the unspecified std::sort tie order.
 * mode 1: exhaustive scan of every (scan, x, y) in the shrunk bounds on the
 *         width-1 grid; ties resolved to the smallest (scan, x, y) -- the
 *         build's canonical tie-break, which the GPU path reproduces bit-exactly.
 * level1 is the uint8 width-1 grid (nx*ny).  Search parameters are the "for
 * testing" ctor (correlative_scan_matcher_2d.cpp:57-71): n_lin, n_ang, step. */
int gloc_oracle_csm_match(const uint8_t* level1, int nx, int ny,
                          double resolution, double max_x, double max_y,
                          int depth, const float* pts, int n_pts,
                          double init_x, double init_y, double init_yaw,
                          int n_lin, int n_ang, double ang_step,
                          float min_score, int mode,
                          gloc_oracle_match_result* out);

/* MatchFullSubmap(point cloud), fast_..._2d.cpp:249-268: window 25*res, +-pi,
 * initial pose = max - 0.5*res*(nx, ny), production SearchParameters. */
int gloc_oracle_csm_match_full_submap(const uint8_t* level1, int nx, int ny,
                                      double resolution, double max_x,
                                      double max_y, int depth, const float* pts,
                                      int n_pts, float min_score, int mode,
                                      gloc_oracle_match_result* out);

/* Batch of independent (map grid, scan) pairs over nthreads pthreads: the CPU
 * baseline for the verification stage.  grids[i] is the width-1 grid of pair
 * i (all nx*ny), pts[i] its P_i x 3 scan. */
void gloc_oracle_csm_match_batch_mt(const uint8_t* const* grids, int nx, int ny,
                                    double resolution, double max_x,
                                    double max_y, int depth,
                                    const float* const* pts, const int* n_pts,
                                    const double* init_xyyaw, int n_pairs,
                                    int n_lin, int n_ang, double ang_step,
                                    float min_score, int mode, int nthreads,
                                    gloc_oracle_match_result* out);

/* ------------------------------------------------------------- BEV projection */

/* One scan -> the reference's BEV image (0 = occupied, 255 = free): see bev_oracle.c for
 * the path restated (loop_detector.cpp:108-135, 3d/submap_3d.cpp:238-326,
 * 3d/range_data_inserter_3d.cpp:57-77, 3d/hybrid_grid.h:429-434).  Pinned against those sources compiled unmodified
 * (oracle/_ref/libbev_ref.so, tests/test_oracle_bev_ref.py). */
int gloc_oracle_bev_project(const float* pts, size_t n, int stride, float resolution,
                            float max_range, uint8_t* img, size_t img_capacity, int* w, int* h,
                            int* min_ix, int* min_iy, double* ox, double* oy,
                            size_t* n_hit_voxels, size_t* n_occupied);
/* RpyPCLoopDetector::crop_pad_occupancy (loop_detector.cpp:83-106), one channel. */
void gloc_oracle_crop_pad(const uint8_t* src, int sw, int sh, int width, int height,
                          uint8_t* dst);

#ifdef __cplusplus
}
#endif
#endif /* GLOC_ORACLE_H_ */
