/*
 * bev_oracle.c -- CPU restatement of the reference's BEV projection of one scan
 * (the producer of both stages' inputs).
 *
 * TEST INFRASTRUCTURE ONLY (see gloc_oracle.h).  PINNED (round 2): the reference's own
 * 3d/submap_3d.cpp, 3d/range_data_inserter_3d.cpp, 3d/range_data.cpp and 3d/hybrid_grid.h
 * compile unmodified against oracle/shim/ into oracle/_ref/libbev_ref.so (oracle/bev_ref.cpp);
 * tests/test_oracle_bev_ref.py checks this file against it image for image and origin bit
 * for bit (the reference's KITTI scan, synthetic scans, rounding boundaries, grid growth), and
 * tests/golden/bev_kitti_subsample.npz is minted from it.  What stays a restatement: the
 * dozen lines of loop_detector.cpp around it (that file needs PCL and libtorch) and the
 * arithmetic of the Eigen shim.  Citations relative to /root/reference/registration/.
 *
 * Reference path for ONE scan inserted into a fresh Submap3D with the identity pose
 * (RpyPCLoopDetector::get_projected_grid, loop_detector.cpp:122-135):
 *   point_cloud_to_range_data (loop_detector.cpp:108-120): a point is a "return" unless
 *     sqrt(x*x + y*y + z*z) > 100 (float products and sums, float sqrt, double compare);
 *   FilterRangeDataByMaxRange(.., 100) (3d/submap_3d.cpp:43-52): keeps returns with
 *     (hit - origin).norm() <= 100.f -- the same predicate on the same float value;
 *   RangeDataInserter3D::Insert (3d/range_data_inserter_3d.cpp:63-77): every return marks
 *     voxel HybridGrid::GetCellIndex(hit) = lround(hit / resolution) per axis, float
 *     division (3d/hybrid_grid.h:429-434) with the hit table: an unknown cell becomes
 *     p = 0.55 (3d/range_data_inserter_3d.cpp:57-61, 3d/probability_values.cpp:73-84) and
 *     stays there for the rest of the update (the update marker, hybrid_grid.h:508-519).
 *     The two free-space voxels before each hit only ever touch cells that were NOT hit
 *     in the same update and leave them at p = 0.49 < 0.501, i.e. invisible below.
 *   ProjectToCvMat (3d/submap_3d.cpp:238-326): voxels with p >= 0.501 (= the hit voxels);
 *     pixel index = lround(cell_center * (1.f / resolution)) = the voxel's (ix, iy) for any
 *     |index| < 2^20; bounding box over ALL hit voxels; a pixel's probability sum is
 *     0.55 x (distinct z voxels of its column); the pixel is 0 (occupied) iff the sum
 *     exceeds kMaxProbability = 0.9, i.e. iff the column holds at least 2 hit voxels,
 *     else 255.  Row = iy - min_iy, column = ix - min_ix.  ox = min_ix * resolution,
 *     oy = min_iy * resolution (double, resolution = (double)0.2f).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "gloc_oracle.h"

typedef struct {
  int x, y, z;
} Vox;

static int vox_cmp(const void* a, const void* b) {
  const Vox* p = (const Vox*)a;
  const Vox* q = (const Vox*)b;
  if (p->y != q->y) return p->y < q->y ? -1 : 1;
  if (p->x != q->x) return p->x < q->x ? -1 : 1;
  if (p->z != q->z) return p->z < q->z ? -1 : 1;
  return 0;
}

/* pts: n points, `stride` floats apart (x, y, z first; KITTI scans are x y z i).
 * Returns 0 and the image geometry; *n_hit_voxels = 0 (w = h = 0) when no point is in range.
 * img (may be NULL to size it) receives h*w bytes, row-major. */
int gloc_oracle_bev_project(const float* pts, size_t n, int stride, float resolution,
                            float max_range, uint8_t* img, size_t img_capacity, int* w, int* h,
                            int* min_ix, int* min_iy, double* ox, double* oy,
                            size_t* n_hit_voxels, size_t* n_occupied) {
  Vox* v = (Vox*)malloc((n ? n : 1) * sizeof(Vox));
  if (!v) return 1;
  size_t m = 0;
  for (size_t i = 0; i < n; ++i) {
    const float x = pts[i * stride], y = pts[i * stride + 1], z = pts[i * stride + 2];
    const float r = sqrtf(x * x + y * y + z * z);
    if (!(r <= max_range)) continue; /* > max_range: a miss; NaN never passes the reference's <= */
    v[m].x = (int)lroundf(x / resolution);
    v[m].y = (int)lroundf(y / resolution);
    v[m].z = (int)lroundf(z / resolution);
    ++m;
  }
  *w = *h = 0;
  *min_ix = *min_iy = 0;
  *ox = *oy = 0.0;
  *n_hit_voxels = 0;
  *n_occupied = 0;
  if (m == 0) {
    free(v);
    return 0;
  }
  qsort(v, m, sizeof(Vox), vox_cmp);
  int mnx = v[0].x, mxx = v[0].x, mny = v[0].y, mxy = v[0].y;
  for (size_t i = 0; i < m; ++i) {
    if (v[i].x < mnx) mnx = v[i].x;
    if (v[i].x > mxx) mxx = v[i].x;
    if (v[i].y < mny) mny = v[i].y;
    if (v[i].y > mxy) mxy = v[i].y;
  }
  *w = mxx - mnx + 1;
  *h = mxy - mny + 1;
  *min_ix = mnx;
  *min_iy = mny;
  *ox = mnx * (double)resolution;
  *oy = mny * (double)resolution;
  const size_t cells = (size_t)*w * (size_t)*h;
  const int write = img != NULL && img_capacity >= cells;
  if (write) memset(img, 255, cells);
  size_t hv = 0, occ = 0;
  for (size_t i = 0; i < m;) {
    size_t j = i, distinct = 0;
    while (j < m && v[j].x == v[i].x && v[j].y == v[i].y) {
      if (j == i || v[j].z != v[j - 1].z) ++distinct;
      ++j;
    }
    hv += distinct;
    if (distinct >= 2) { /* 2 x 0.55 > 0.9 >= 1 x 0.55 */
      ++occ;
      if (write) img[(size_t)(v[i].y - mny) * (size_t)*w + (size_t)(v[i].x - mnx)] = 0;
    }
    i = j;
  }
  *n_hit_voxels = hv;
  *n_occupied = occ;
  free(v);
  return 0;
}

/* RpyPCLoopDetector::crop_pad_occupancy (loop_detector.cpp:83-106) for a 1-channel image:
 * the centre (width x height) window of src pasted into the centre of a 255-filled
 * (width x height) image (the 3-channel conversion only replicates the value). */
void gloc_oracle_crop_pad(const uint8_t* src, int sw, int sh, int width, int height,
                          uint8_t* dst) {
  memset(dst, 255, (size_t)width * (size_t)height);
  const int cw = sw >= width ? width : sw, ch = sh >= height ? height : sh;
  const int sx = (int)floor((sw - cw) / 2.), sy = (int)floor((sh - ch) / 2.);
  const int dx = (int)floor((width - cw) / 2.), dy = (int)floor((height - ch) / 2.);
  for (int r = 0; r < ch; ++r)
    memcpy(dst + (size_t)(dy + r) * width + dx, src + (size_t)(sy + r) * sw + sx, (size_t)cw);
}
