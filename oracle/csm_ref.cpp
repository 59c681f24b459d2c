// csm_ref.cpp -- C wrapper around the REFERENCE's own scan matcher, compiled from the sources
// where they lie under /root/reference/registration (never copied into this repo):
//   2d/fast_correlative_scan_matcher_2d.cpp  2d/correlative_scan_matcher_2d.cpp  2d/grid_2d.cpp
//   2d/probability_grid.cpp  3d/probability_values.cpp  3d/point_cloud.cpp
// into oracle/_ref/libcsm_ref.so by oracle/Makefile, UNMODIFIED, against oracle/shim/ (minimal
// stand-ins for Eigen, glog, OpenCV's cv::Mat, boost::iostreams and ceres::atan2, none of which is
// installed in this image).
//
// TEST INFRASTRUCTURE ONLY: pins oracle/csm_oracle.c (and through it the GPU path) against the
// reference's PrecomputationGrid2D / SlidingWindowMaximum, SearchParameters + ShrinkToFit,
// GenerateRotatedScans, DiscretizeScans, candidate generation, ScoreCandidates, BranchAndBound and
// MatchWithSearchParameters.  What stays a restatement is the arithmetic inside the Eigen shim
// (oracle/shim/Eigen/eigen_shim.h says which formulas).
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "2d/fast_correlative_scan_matcher_2d.h"   // -I/root/reference/registration
#include "2d/probability_grid.h"
#include "3d/probability_values.h"

namespace carto = cartographer;
using carto::mapping::CellLimits;
using carto::mapping::MapLimits;
using carto::mapping::ProbabilityGrid;
using carto::mapping::scan_matching::Candidate2D;
using carto::mapping::scan_matching::DiscreteScan2D;
using carto::mapping::scan_matching::FastCorrelativeScanMatcher2D;
using carto::mapping::scan_matching::FastCorrelativeScanMatcherOptions2D;
using carto::mapping::scan_matching::SearchParameters;

namespace {

// Grid2D keeps its cells protected; a derived class may hand them out (no reference source is touched).
struct RawGrid : ProbabilityGrid {
  explicit RawGrid(const MapLimits& l) : ProbabilityGrid(l) {}
  std::vector<carto::uint16>* cells() { return mutable_correspondence_cost_cells(); }
};

std::unique_ptr<RawGrid> make_grid(const uint16_t* cells, int nx, int ny, double res, double max_x, double max_y) {
  std::unique_ptr<RawGrid> g(new RawGrid(MapLimits(res, Eigen::Vector2d(max_x, max_y), CellLimits(nx, ny))));
  std::memcpy(g->cells()->data(), cells, sizeof(uint16_t) * (size_t)nx * ny);
  return g;
}

carto::sensor::PointCloud make_cloud(const float* pts, int n) {
  carto::sensor::PointCloud pc;
  pc.reserve(n);
  for (int i = 0; i < n; ++i) pc.emplace_back(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
  return pc;
}

}  // namespace

extern "C" {

struct gloc_ref_match_result {
  int found;        // MatchWithSearchParameters' return value
  float score;      // *score (untouched = min_score sentinel written by this wrapper when !found)
  double pose_x, pose_y, pose_yaw;   // *pose_estimate: translation, rotation().angle()
  // the winning Candidate2D, obtained by calling the same public pieces MatchWithSearchParameters
  // calls (fast_..._2d.cpp:278-310); std::sort's order among equal scores is unspecified
  int scan_index, x_offset, y_offset;
  float cand_score;
};

// ValueToCorrespondenceCost through the reference's table (3d/probability_values.cpp)
float gloc_ref_value_to_cost(uint16_t v) { return carto::mapping::ValueToCorrespondenceCost(v); }
uint16_t gloc_ref_cost_to_value(float c) { return carto::mapping::CorrespondenceCostToValue(c); }

// PrecomputationGridStack2D::Get(index) read back cell by cell through the public GetValue
// (out: (nx+w-1) x (ny+w-1), stride nx+w-1, offset (-w+1, -w+1))
void gloc_ref_csm_precomp(const uint16_t* cells, int nx, int ny, double res, double max_x, double max_y,
                          int depth, int index, uint8_t* out) {
  auto g = make_grid(cells, nx, ny, res, max_x, max_y);
  FastCorrelativeScanMatcherOptions2D opt;
  opt.branch_and_bound_depth_ = depth;
  carto::mapping::scan_matching::PrecomputationGridStack2D stack(*g, opt);
  const auto& pg = stack.Get(index);
  const int w = 1 << index, wnx = nx + w - 1, wny = ny + w - 1;
  for (int y = 0; y < wny; ++y)
    for (int x = 0; x < wnx; ++x)
      out[(size_t)y * wnx + x] = (uint8_t)pg.GetValue(Eigen::Array2i(x - w + 1, y - w + 1));
}

// SearchParameters production ctor (correlative_scan_matcher_2d.cpp:27-55)
void gloc_ref_csm_search_params(double lin, double ang, const float* pts, int n, double res, int* n_lin,
                                int* n_ang, double* step) {
  const SearchParameters sp(lin, ang, make_cloud(pts, n), res);
  *n_lin = sp.linear_bounds.empty() ? 0 : sp.linear_bounds[0].max_x;
  *n_ang = sp.num_angular_perturbations;
  *step = sp.angular_perturbation_step_size;
}

// the prologue of MatchWithSearchParameters (fast_..._2d.cpp:278-289): out S x P x 2 int32
void gloc_ref_csm_discretize(const float* pts, int n, double init_x, double init_y, double init_yaw, int n_ang,
                             double step, double res, double max_x, double max_y, int32_t* out) {
  const MapLimits limits(res, Eigen::Vector2d(max_x, max_y), CellLimits(1, 1));
  const SearchParameters sp(0, n_ang, step, res);
  const carto::transform::Rigid2d init({init_x, init_y}, init_yaw);
  const Eigen::Rotation2Dd initial_rotation = init.rotation();
  const carto::sensor::PointCloud rotated = carto::sensor::TransformPointCloud(
      make_cloud(pts, n), carto::transform::Rigid3f::Rotation(Eigen::AngleAxisf(
                              initial_rotation.cast<float>().angle(), Eigen::Vector3f::UnitZ())));
  const auto scans = carto::mapping::scan_matching::GenerateRotatedScans(rotated, sp);
  const auto disc = carto::mapping::scan_matching::DiscretizeScans(
      limits, scans, Eigen::Translation2f(init.translation().x(), init.translation().y()));
  for (size_t s = 0; s < disc.size(); ++s)
    for (int p = 0; p < n; ++p) {
      out[2 * (s * n + p)] = disc[s][p].x();
      out[2 * (s * n + p) + 1] = disc[s][p].y();
    }
}

int gloc_ref_csm_match(const uint16_t* cells, int nx, int ny, double res, double max_x, double max_y, int depth,
                       const float* pts, int n_pts, double init_x, double init_y, double init_yaw, int n_lin,
                       int n_ang, double step, float min_score, gloc_ref_match_result* out) {
  auto g = make_grid(cells, nx, ny, res, max_x, max_y);
  FastCorrelativeScanMatcherOptions2D opt;
  opt.branch_and_bound_depth_ = depth;
  const FastCorrelativeScanMatcher2D matcher(*g, opt);
  const carto::sensor::PointCloud cloud = make_cloud(pts, n_pts);
  const carto::transform::Rigid2d init({init_x, init_y}, init_yaw);
  std::memset(out, 0, sizeof(*out));
  float score = min_score;
  carto::transform::Rigid2d pose = init;
  out->found = matcher.MatchWithSearchParameters(SearchParameters(n_lin, n_ang, step, res), init, cloud,
                                                 min_score, &score, &pose) ? 1 : 0;
  out->score = score;
  out->pose_x = pose.translation().x();
  out->pose_y = pose.translation().y();
  out->pose_yaw = pose.rotation().angle();
  // the candidate behind that pose: the same calls, in the same order, on the public members
  SearchParameters sp(n_lin, n_ang, step, res);
  const Eigen::Rotation2Dd initial_rotation = init.rotation();
  const carto::sensor::PointCloud rotated = carto::sensor::TransformPointCloud(
      cloud, carto::transform::Rigid3f::Rotation(Eigen::AngleAxisf(initial_rotation.cast<float>().angle(),
                                                                    Eigen::Vector3f::UnitZ())));
  const auto scans = carto::mapping::scan_matching::GenerateRotatedScans(rotated, sp);
  const std::vector<DiscreteScan2D> disc = carto::mapping::scan_matching::DiscretizeScans(
      g->limits(), scans, Eigen::Translation2f(init.translation().x(), init.translation().y()));
  sp.ShrinkToFit(disc, g->limits().cell_limits());
  const std::vector<Candidate2D> lowest = matcher.ComputeLowestResolutionCandidates(disc, sp);
  const Candidate2D best = matcher.BranchAndBound(disc, sp, lowest, depth - 1, min_score);
  out->scan_index = best.scan_index;
  out->x_offset = best.x_index_offset;
  out->y_offset = best.y_index_offset;
  out->cand_score = best.score;
  return out->found;
}

// Match(initial pose, Grid2D): GridToVirtualPointCloud (fast_..._2d.cpp:78-95) + the production
// SearchParameters with the default options' windows scaled by the caller
int gloc_ref_csm_match_grid(const uint16_t* cells, int nx, int ny, double res, double max_x, double max_y, int depth,
                            double lin_window, double ang_window, const uint16_t* q_cells, int qnx, int qny,
                            double q_max_x, double q_max_y, double q_ox, double q_oy, double init_x, double init_y,
                            double init_yaw, float min_score, gloc_ref_match_result* out) {
  auto g = make_grid(cells, nx, ny, res, max_x, max_y);
  auto q = make_grid(q_cells, qnx, qny, res, q_max_x, q_max_y);
  q->SetOrigin(q_ox, q_oy);
  FastCorrelativeScanMatcherOptions2D opt;
  opt.branch_and_bound_depth_ = depth;
  opt.linear_search_window_ = lin_window;
  opt.angular_search_window_ = ang_window;
  const FastCorrelativeScanMatcher2D matcher(*g, opt);
  const carto::transform::Rigid2d init({init_x, init_y}, init_yaw);
  std::memset(out, 0, sizeof(*out));
  float score = min_score;
  carto::transform::Rigid2d pose = init;
  out->found = matcher.Match(init, *q, min_score, &score, &pose) ? 1 : 0;
  out->score = score;
  out->pose_x = pose.translation().x();
  out->pose_y = pose.translation().y();
  out->pose_yaw = pose.rotation().angle();
  return out->found;
}

int gloc_ref_csm_match_full_submap(const uint16_t* cells, int nx, int ny, double res, double max_x, double max_y,
                                   int depth, const float* pts, int n_pts, float min_score,
                                   gloc_ref_match_result* out) {
  auto g = make_grid(cells, nx, ny, res, max_x, max_y);
  FastCorrelativeScanMatcherOptions2D opt;
  opt.branch_and_bound_depth_ = depth;
  const FastCorrelativeScanMatcher2D matcher(*g, opt);
  std::memset(out, 0, sizeof(*out));
  float score = min_score;
  carto::transform::Rigid2d pose;
  out->found = matcher.MatchFullSubmap(make_cloud(pts, n_pts), min_score, &score, &pose) ? 1 : 0;
  out->score = score;
  out->pose_x = pose.translation().x();
  out->pose_y = pose.translation().y();
  out->pose_yaw = pose.rotation().angle();
  return out->found;
}

// Independent (map grid, scan) pairs over nthreads threads, one matcher per pair -- how the
// reference would verify candidates of a large database (a PrecomputationGridStack2D per map grid
// cannot be kept for every frame): the CPU baseline of the verification stage.
void gloc_ref_csm_match_batch_mt(const uint16_t* const* cells, int nx, int ny, double res, double max_x,
                                 double max_y, int depth, const float* const* pts, const int* n_pts,
                                 const double* init_xyyaw, int n_pairs, int n_lin, int n_ang, double step,
                                 float min_score, int nthreads, gloc_ref_match_result* out) {
  if (nthreads < 1) nthreads = 1;
  auto work = [&](int t) {
    for (int i = t; i < n_pairs; i += nthreads) {
      auto g = make_grid(cells[i], nx, ny, res, max_x, max_y);
      FastCorrelativeScanMatcherOptions2D opt;
      opt.branch_and_bound_depth_ = depth;
      const FastCorrelativeScanMatcher2D matcher(*g, opt);
      const carto::transform::Rigid2d init({init_xyyaw[3 * i], init_xyyaw[3 * i + 1]}, init_xyyaw[3 * i + 2]);
      std::memset(&out[i], 0, sizeof(out[i]));
      float score = min_score;
      carto::transform::Rigid2d pose = init;
      out[i].found = matcher.MatchWithSearchParameters(SearchParameters(n_lin, n_ang, step, res), init,
                                                       make_cloud(pts[i], n_pts[i]), min_score, &score, &pose) ? 1 : 0;
      out[i].score = score;
      out[i].pose_x = pose.translation().x();
      out[i].pose_y = pose.translation().y();
      out[i].pose_yaw = pose.rotation().angle();
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
}

}  // extern "C"
