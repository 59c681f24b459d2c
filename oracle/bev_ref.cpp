// bev_ref.cpp -- C wrapper around the REFERENCE's own BEV projection, compiled from the sources
// where they lie under /root/reference/registration (never copied into this repo):
//   3d/submap_3d.cpp (Submap3D::InsertRangeData, ProjectToCvMat)  3d/range_data_inserter_3d.cpp
//   3d/hybrid_grid.h  3d/probability_values.cpp  2d/grid_2d.cpp  2d/probability_grid.cpp (linked: submap_3d.cpp
//   also holds ProjectToGrid)
// into oracle/_ref/libbev_ref.so by oracle/Makefile, UNMODIFIED, against oracle/shim/.
//
// TEST INFRASTRUCTURE ONLY: pins oracle/bev_oracle.c (and through it the GPU path gloc_bev_*) against
// the reference's Submap3D + RangeDataInserter3D + HybridGrid + ProjectToCvMat, i.e. everything
// RpyPCLoopDetector::get_projected_grid (loop_detector.cpp:122-135) calls.  The two helpers of
// loop_detector.cpp itself cannot be compiled here (the file needs PCL and libtorch) and are restated
// below, each a handful of lines:
//   point_cloud_to_range_data  loop_detector.cpp:108-120   returns / misses split at a norm of 100 m
//   get_projected_grid         loop_detector.cpp:122-135   the call sequence, constants of loop_detector.h:114-118
#include <cmath>
#include <cstdint>
#include <cstring>

#include "3d/submap_3d.h"                // -I/root/reference/registration
#include "3d/range_data_inserter_3d.h"

namespace carto = cartographer;

extern "C" {

// One scan (n points, `stride` floats each, xyz first) -> the reference's BEV image.  img == nullptr:
// only the shape and origin are returned.  Returns 0, or -1 when img_capacity is too small.
int gloc_ref_bev_project(const float* pts, size_t n, int stride, uint8_t* img, size_t img_capacity, int* w, int* h,
                         double* ox, double* oy, double* resolution) {
  const float high_resolution_max_range = 100.f, high_resolution = 0.2f, low_resolution = 0.5f;   // loop_detector.h:115-117
  const carto::transform::Rigid3d identity = carto::transform::Rigid3d::Identity();
  carto::mapping::RangeDataInserter3D inserter;
  carto::sensor::RangeData rd;                                     // point_cloud_to_range_data
  rd.origin << 0., 0., 0.;
  for (size_t i = 0; i < n; ++i) {
    const float x = pts[i * stride], y = pts[i * stride + 1], z = pts[i * stride + 2];
    if (sqrt(x * x + y * y + z * z) > 100.) {
      rd.misses.emplace_back(Eigen::Vector3f(x, y, z));
    } else {
      rd.returns.emplace_back(Eigen::Vector3f(x, y, z));
    }
  }
  carto::mapping::Submap3D submap(high_resolution, low_resolution, identity);   // get_projected_grid
  submap.InsertRangeData(rd, inserter, high_resolution_max_range);
  double px = 0, py = 0, res = 0;
  cv::Mat m = carto::mapping::ProjectToCvMat(&submap.high_resolution_hybrid_grid(), identity, px, py, res);
  *w = m.cols;
  *h = m.rows;
  *ox = px;
  *oy = py;
  *resolution = res;
  if (!img) return 0;
  if (img_capacity < (size_t)m.rows * (size_t)m.cols) return -1;
  for (int r = 0; r < m.rows; ++r)
    for (int c = 0; c < m.cols; ++c) img[(size_t)r * m.cols + c] = m.at<uchar>(r, c);
  return 0;
}

}  // extern "C"
