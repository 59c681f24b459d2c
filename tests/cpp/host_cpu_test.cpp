// The host-only half of the C++ mirrors: everything that must work (or fail loudly) WITHOUT a
// GPU -- the loop detector's guards, SearchParameters, GridToVirtualPointCloud, the options'
// defaults -- checked against the CPU oracle.  Build: see tests/test_host_cpp.py.
#include <cmath>
#include <cstdio>
#include <random>

#include "../../gloc3d_b200/host/gloc_loop_detector.hpp"
#include "../../oracle/gloc_oracle.h"

using namespace cartographer::mapping;
using namespace cartographer::mapping::scan_matching;

#define EXPECT(c)                                                     \
  do {                                                                \
    if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } \
  } while (0)

int main() {
  std::mt19937 rng(3);
  std::uniform_real_distribution<float> u(-40.f, 40.f);

  // FastCorrelativeScanMatcherOptions2D defaults (fast_correlative_scan_matcher_2d.h:43-52)
  FastCorrelativeScanMatcherOptions2D opt;
  EXPECT(opt.linear_search_window() == 3. && opt.angular_search_window() == 3. && opt.branch_and_bound_depth() == 5);

  // SearchParameters, production ctor (correlative_scan_matcher_2d.cpp:27-55) vs the oracle
  PointCloud cloud;
  for (int i = 0; i < 500; ++i) cloud.push_back({u(rng), u(rng), 0.f});
  for (double res : {0.05, 0.2, 0.5}) {
    const SearchParameters sp(7.0, 1.2, cloud, res);
    int nl = 0, na = 0;
    double st = 0;
    gloc_oracle_search_params(7.0, 1.2, cloud[0].data(), (int)cloud.size(), res, &nl, &na, &st);
    EXPECT(sp.num_linear_perturbations == nl && sp.num_angular_perturbations == na);
    EXPECT(sp.angular_perturbation_step_size == st && sp.num_scans == 2 * na + 1 && sp.resolution == res);
  }
  const SearchParameters test_ctor(100, 180, 2. * M_PI / 360., 0.2);   // "For testing", :57-71
  EXPECT(test_ctor.num_scans == 361 && test_ctor.num_linear_perturbations == 100);

  // GridToVirtualPointCloud (fast_correlative_scan_matcher_2d.cpp:78-95) vs the oracle
  const int nx = 37, ny = 23;
  std::vector<uint16_t> cells((size_t)nx * ny, 0);
  std::uniform_int_distribution<int> pick(0, 9);
  for (auto& c : cells) {
    const int r = pick(rng);
    c = r == 0 ? 1 : r == 1 ? 400 : r == 2 ? 500 : r == 3 ? 32767 : 0;
  }
  Grid2DView grid{{0.2, 5.0, 7.0, nx, ny}, cells.data(), -3.1, 2.7};
  const PointCloud pc = FastCorrelativeScanMatcher2D::GridToVirtualPointCloud(grid);
  std::vector<float> ref((size_t)nx * ny * 3);
  const int n_ref = gloc_oracle_grid_to_points(cells.data(), nx, ny, 0.2, -3.1, 2.7, ref.data(), nx * ny);
  EXPECT((int)pc.size() == n_ref && n_ref > 50);
  for (int i = 0; i < n_ref; ++i)
    EXPECT(pc[i][0] == ref[3 * i] && pc[i][1] == ref[3 * i + 1] && pc[i][2] == ref[3 * i + 2]);

  // RpyPCLoopDetector guards (loop_detector.cpp:27-30, :64): not enough keyframes -> outputs
  // untouched / false, and nothing touches the GPU before the guard passes
  RpyPCLoopDetectorGpu det;
  std::vector<float> feat(512, 0.1f);
  for (int i = 0; i < 50; ++i) det.add_keyframe(feat, grid);   // 50 <= 30 + 20
  std::vector<size_t> idx;
  std::vector<float> d2;
  det.detect(feat, idx, d2);
  EXPECT(idx.empty() && d2.empty());
  size_t qi = 7, li = 9;
  EXPECT(!det.detect(qi, li) && qi == 7 && li == 9);
  EXPECT(det.NUM_EXCLUDE_RECENT == 30);
  std::printf("PASS host cpu\n");
  return 0;
}
