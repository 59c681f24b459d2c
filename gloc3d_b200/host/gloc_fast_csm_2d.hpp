// gloc_fast_csm_2d.hpp -- C++ host mirror of the reference's scan matcher over the C ABI.
//
// Same class names, constructor and Match* signatures as
//   cartographer::mapping::scan_matching::FastCorrelativeScanMatcher2D
//   (/root/reference/registration/2d/fast_correlative_scan_matcher_2d.h:43-52,137-200)
// with the Eigen/glog-dependent parameter types replaced by plain views (Eigen is not
// needed to call the GPU path; INTEGRATION.md shows the two-line adapters from
// Grid2D / sensor::PointCloud / transform::Rigid2d).  Header-only; link with -lgloc3d.
#ifndef GLOC_FAST_CSM_2D_HPP_
#define GLOC_FAST_CSM_2D_HPP_

#include <array>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/gloc3d.h"

namespace cartographer {
namespace mapping {

// MapLimits (2d/map_limits.h:40-90) + CellLimits (2d/xy_index.h:34-45)
struct MapLimitsView {
  double resolution;
  double max_x, max_y;          // MapLimits::max()
  int num_x_cells, num_y_cells; // cell_limits()
};

// What the matcher reads of a Grid2D (2d/grid_2d.h:34-111): limits, the uint16
// correspondence-cost cells (flat index num_x_cells*y + x) and the author's explicit origin.
struct Grid2DView {
  MapLimitsView limits;
  const uint16_t* correspondence_cost_cells;
  double ox = 0., oy = 0.;
};

namespace scan_matching {

struct Rigid2d {  // transform::Rigid2d as translation + Rotation2Dd angle
  double x = 0., y = 0., yaw = 0.;
};
using PointCloud = std::vector<std::array<float, 3>>;  // sensor::PointCloud (Vector3f)

class FastCorrelativeScanMatcherOptions2D {  // fast_correlative_scan_matcher_2d.h:43-52
 public:
  double linear_search_window() const { return linear_search_window_; }
  double angular_search_window() const { return angular_search_window_; }
  int branch_and_bound_depth() const { return branch_and_bound_depth_; }
  double linear_search_window_ = 3.;
  double angular_search_window_ = 3.;
  int branch_and_bound_depth_ = 5;
};

// SearchParameters (2d/correlative_scan_matcher_2d.h:35-61): the fields the GPU path needs.
struct SearchParameters {
  SearchParameters(double linear_search_window, double angular_search_window,
                   const PointCloud& point_cloud, double resolution)
      : resolution(resolution) {
    int n_lin = 0;
    if (gloc_csm_search_params(linear_search_window, angular_search_window,
                               point_cloud.empty() ? nullptr : point_cloud[0].data(),
                               static_cast<int>(point_cloud.size()), resolution, &n_lin,
                               &num_angular_perturbations, &angular_perturbation_step_size) != GLOC_OK)
      throw std::runtime_error(std::string("libgloc3d: ") + gloc_last_error());
    num_linear_perturbations = n_lin;
    num_scans = 2 * num_angular_perturbations + 1;
  }
  // "For testing" (correlative_scan_matcher_2d.cpp:57-71)
  SearchParameters(int num_linear_perturbations, int num_angular_perturbations,
                   double angular_perturbation_step_size, double resolution)
      : num_linear_perturbations(num_linear_perturbations),
        num_angular_perturbations(num_angular_perturbations),
        angular_perturbation_step_size(angular_perturbation_step_size),
        resolution(resolution),
        num_scans(2 * num_angular_perturbations + 1) {}
  int num_linear_perturbations;
  int num_angular_perturbations;
  double angular_perturbation_step_size;
  double resolution;
  int num_scans;
};

class FastCorrelativeScanMatcher2D {
 public:
  FastCorrelativeScanMatcher2D(const Grid2DView& grid, const FastCorrelativeScanMatcherOptions2D& options,
                               int device = 0)
      : options_(options), limits_(grid.limits) {
    if (options.branch_and_bound_depth() < 1)  // CHECK_GE aborts in the reference (fast_..._2d.cpp:195)
      throw std::invalid_argument("Check failed: options.branch_and_bound_depth() >= 1");
    check(gloc_csm_create(&store_, device));
    check(gloc_csm_add_grid_cells(store_, grid.correspondence_cost_cells, limits_.num_x_cells,
                                  limits_.num_y_cells, limits_.resolution, limits_.max_x,
                                  limits_.max_y, &grid_id_));
  }
  ~FastCorrelativeScanMatcher2D() { gloc_csm_destroy(store_); }
  FastCorrelativeScanMatcher2D(const FastCorrelativeScanMatcher2D&) = delete;
  FastCorrelativeScanMatcher2D& operator=(const FastCorrelativeScanMatcher2D&) = delete;

  // Aligns 'point_cloud' within the 'grid' given an 'initial_pose_estimate'.  If a score above
  // 'min_score' (excluding equality) is possible, true is returned, and 'score' and
  // 'pose_estimate' are updated with the result (fast_..._2d.cpp:219-229).
  bool Match(const Rigid2d& initial_pose_estimate, const PointCloud& point_cloud, float min_score,
             float* score, Rigid2d* pose_estimate) const {
    const SearchParameters sp(options_.linear_search_window(), options_.angular_search_window(),
                              point_cloud, limits_.resolution);
    return MatchWithSearchParameters(sp, initial_pose_estimate, point_cloud, min_score, score,
                                     pose_estimate);
  }
  bool Match(const Rigid2d& initial_pose_estimate, const Grid2DView& prob_grid, float min_score,
             float* score, Rigid2d* pose_estimate) const {  // :231-238
    return Match(initial_pose_estimate, GridToVirtualPointCloud(prob_grid), min_score, score,
                 pose_estimate);
  }
  // +-25 cells, +-pi around the grid centre (fast_..._2d.cpp:240-268)
  bool MatchFullSubmap(const PointCloud& point_cloud, float min_score, float* score,
                       Rigid2d* pose_estimate) const {
    const SearchParameters sp(25 * limits_.resolution, M_PI, point_cloud, limits_.resolution);
    Rigid2d center;
    center.x = limits_.max_x - 0.5 * limits_.resolution * limits_.num_x_cells;
    center.y = limits_.max_y - 0.5 * limits_.resolution * limits_.num_y_cells;
    return MatchWithSearchParameters(sp, center, point_cloud, min_score, score, pose_estimate);
  }
  bool MatchFullSubmap(const Grid2DView& prob_grid, float min_score, float* score,
                       Rigid2d* pose_estimate) const {
    return MatchFullSubmap(GridToVirtualPointCloud(prob_grid), min_score, score, pose_estimate);
  }

  // fast_..._2d.cpp:270-320; public in the reference because "// private:" is commented out
  bool MatchWithSearchParameters(SearchParameters search_parameters,
                                 const Rigid2d& initial_pose_estimate, const PointCloud& point_cloud,
                                 float min_score, float* score, Rigid2d* pose_estimate) const {
    if (score == nullptr || pose_estimate == nullptr)  // CHECK_NOTNULL (:275-276)
      throw std::invalid_argument("Check failed: score / pose_estimate must not be null");
    if (point_cloud.empty()) return false;
    const int64_t offs[2] = {0, static_cast<int64_t>(point_cloud.size())};
    const int gid = grid_id_, sid = 0;
    const double init[3] = {initial_pose_estimate.x, initial_pose_estimate.y, initial_pose_estimate.yaw};
    gloc_csm_result r;
    check(gloc_csm_match_batch(store_, point_cloud[0].data(), offs, 1, &gid, &sid, init, 1,
                               search_parameters.num_linear_perturbations,
                               search_parameters.num_angular_perturbations,
                               search_parameters.angular_perturbation_step_size,
                               options_.branch_and_bound_depth(), min_score, &r));
    if (!r.found) return false;  // outputs untouched (:311-319)
    *score = r.score;
    pose_estimate->x = r.pose_x;
    pose_estimate->y = r.pose_y;
    pose_estimate->yaw = r.pose_yaw;
    return true;
  }

  // PrecomputationGridStack2D::Get(index) (fast_..._2d.h:123-125): width 2^index grid,
  // (num_x_cells + w - 1) * (num_y_cells + w - 1) cells.
  std::vector<uint8_t> PrecomputationGrid(int index) const {
    const int w = 1 << index;
    std::vector<uint8_t> out(static_cast<size_t>(limits_.num_x_cells + w - 1) * (limits_.num_y_cells + w - 1));
    check(gloc_csm_get_precomputation_grid(store_, grid_id_, w, out.data()));
    return out;
  }

  // fast_..._2d.cpp:78-95
  static PointCloud GridToVirtualPointCloud(const Grid2DView& grid) {
    int n = 0;
    check(gloc_csm_grid_to_points(grid.correspondence_cost_cells, grid.limits.num_x_cells,
                                  grid.limits.num_y_cells, grid.limits.resolution, grid.ox, grid.oy,
                                  nullptr, 0, &n));
    PointCloud pc(static_cast<size_t>(n));
    if (n > 0)
      check(gloc_csm_grid_to_points(grid.correspondence_cost_cells, grid.limits.num_x_cells,
                                    grid.limits.num_y_cells, grid.limits.resolution, grid.ox, grid.oy,
                                    pc[0].data(), n, &n));
    return pc;
  }

 private:
  static void check(int rc) {
    if (rc != GLOC_OK) throw std::runtime_error(std::string("libgloc3d: ") + gloc_last_error());
  }
  const FastCorrelativeScanMatcherOptions2D options_;
  MapLimitsView limits_;
  gloc_csm_store* store_ = nullptr;
  int grid_id_ = 0;
};

}  // namespace scan_matching
}  // namespace mapping
}  // namespace cartographer

#endif  // GLOC_FAST_CSM_2D_HPP_
