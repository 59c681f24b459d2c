// C++ test of the host mirrors (InvKeyTree, FastCorrelativeScanMatcher2D, the loop detector's
// hot path) against the CPU oracle.  Build: see tests/test_host_cpp.py.  Needs a B200 to run.
#include <cmath>
#include <cstdio>
#include <cstdlib>
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
  std::mt19937 rng(7);
  std::normal_distribution<float> nd(0.f, 1.f / std::sqrt(512.f));
  const size_t N = 3000, D = 512, K = 20;
  KeyMat db(N, std::vector<float>(D));
  std::vector<float> flat(N * D);
  for (size_t i = 0; i < N; ++i)
    for (size_t d = 0; d < D; ++d) flat[i * D + d] = db[i][d] = nd(rng);

  // ---- InvKeyTree exactly as loop_detector.cpp:34-45 uses it
  InvKeyTree tree(D, db, 10);
  std::vector<float> feat(D);
  for (auto& v : feat) v = nd(rng);
  std::vector<size_t> ret_indexes(K);
  std::vector<float> out_dists_sqr(K);
  tree.query(&feat[0], K, &ret_indexes[0], &out_dists_sqr[0]);
  std::vector<uint64_t> oi(K);
  std::vector<float> od(K);
  gloc_oracle_knn(flat.data(), N, D, feat.data(), 1, K, oi.data(), od.data());
  for (size_t i = 0; i < K; ++i) EXPECT(ret_indexes[i] == oi[i] && out_dists_sqr[i] == od[i]);
  EXPECT(tree.kdtree_get_point_count() == N);
  bool threw = false;
  try { KeyMat empty; (void)empty; std::vector<float> q(D); size_t i0; float d0; tree.query(q.data(), 0, &i0, &d0); }
  catch (const std::runtime_error&) { threw = true; }
  EXPECT(threw);

  // ---- FastCorrelativeScanMatcher2D: planted offset, MatchFullSubmap + Match(grid)
  const int nx = 120, ny = 100;
  const double res = 0.2, mx = 14.0, my = 17.0;
  std::vector<uint16_t> cells((size_t)nx * ny, 0);
  std::uniform_int_distribution<int> ux(8, nx - 9), uy(8, ny - 9);
  for (int s = 0; s < 14; ++s) {
    int x = ux(rng), y = uy(rng), dx = (s % 3) - 1, dy = ((s / 3) % 3) - 1;
    if (dx == 0 && dy == 0) dx = 1;
    for (int t = 0; t < 25; ++t) {
      const int cx = x + dx * t, cy = y + dy * t;
      if (cx >= 0 && cy >= 0 && cx < nx && cy < ny) cells[(size_t)nx * cy + cx] = 1;  // occupied
    }
  }
  Grid2DView grid{{res, mx, my, nx, ny}, cells.data(), 0., 0.};
  FastCorrelativeScanMatcherOptions2D opt;
  EXPECT(opt.linear_search_window() == 3. && opt.angular_search_window() == 3. && opt.branch_and_bound_depth() == 5);
  FastCorrelativeScanMatcher2D matcher(grid, opt);
  std::vector<uint8_t> l1((size_t)nx * ny);
  gloc_oracle_level1_from_cells(cells.data(), nx, ny, l1.data());
  EXPECT(matcher.PrecomputationGrid(0) == l1);
  // scan = occupied cell centres relative to the grid centre, shifted by (0.4, -0.2)
  PointCloud cloud;
  const double cx0 = mx - 0.5 * res * nx, cy0 = my - 0.5 * res * ny;
  for (int y = 0; y < ny; ++y)
    for (int x = 0; x < nx; ++x)
      if (cells[(size_t)nx * y + x]) {
        const double wy = my - (x + 0.5) * res, wx = mx - (y + 0.5) * res;
        cloud.push_back({(float)(wx - cx0 - 0.4), (float)(wy - cy0 + 0.2), 0.f});
      }
  float score = -1.f;
  Rigid2d pose;
  EXPECT(matcher.MatchFullSubmap(cloud, 0.5f, &score, &pose));
  gloc_oracle_match_result o;
  gloc_oracle_csm_match_full_submap(l1.data(), nx, ny, res, mx, my, 5, cloud[0].data(), (int)cloud.size(), 0.5f, 0, &o);
  EXPECT(o.found && score == o.score && pose.x == o.pose_x && pose.y == o.pose_y && pose.yaw == o.pose_yaw);
  EXPECT(std::fabs(pose.x - (cx0 + 0.4)) < 0.11 && std::fabs(pose.y - (cy0 - 0.2)) < 0.11);
  float s2 = -7.f;
  Rigid2d p2;
  p2.x = 123.;
  EXPECT(!matcher.MatchFullSubmap(cloud, 0.95f, &s2, &p2) && s2 == -7.f && p2.x == 123.);  // outputs untouched

  // ---- loop detector hot path: guard, global detect, SLAM detect
  RpyPCLoopDetectorGpu det;
  std::vector<size_t> idx;
  std::vector<float> d2;
  for (size_t i = 0; i < 40; ++i) det.add_keyframe(db[i], grid);
  det.detect(feat, idx, d2);
  EXPECT(idx.empty());  // N <= 30 + 20: outputs untouched (loop_detector.cpp:27-30)
  for (size_t i = 40; i < 400; ++i) det.add_keyframe(db[i], grid);
  det.detect(feat, idx, d2);
  gloc_oracle_knn(flat.data(), 400, D, feat.data(), 1, K, oi.data(), od.data());
  for (size_t i = 0; i < K; ++i) EXPECT(idx[i] == oi[i] && d2[i] == od[i]);
  std::vector<float> near = db[123];
  near[5] += 1e-3f;
  det.add_keyframe(near, grid);   // the newest keyframe revisits keyframe 123
  size_t qi = 0, li = 0;
  EXPECT(det.detect(qi, li) && qi == 400 && li == 123);

  // ---- match() against one resident store, and the batched detect_all_query + global_registraion
  {
    float xy_yaw[3] = {9.f, 9.f, 9.f};
    double scale = 0.;
    Grid2DView qg{{res, mx, my, nx, ny}, cells.data(), mx - res * ny, my - res * nx};   // the same place, seen again
    // GridToVirtualPointCloud puts cell (i, j) at (ox + i res, oy + j res): match it against keyframe 7
    const bool ok = det.match(qg, 7, xy_yaw, scale, 20, 10, 0.02, 0.3f);
    const PointCloud qc = FastCorrelativeScanMatcher2D::GridToVirtualPointCloud(qg);
    gloc_oracle_match_result om;
    gloc_oracle_csm_match(l1.data(), nx, ny, res, mx, my, 5, qc[0].data(), (int)qc.size(), 0., 0., 0., 20, 10, 0.02,
                          0.3f, 0, &om);
    EXPECT(ok == (om.found != 0));
    if (ok) EXPECT(xy_yaw[0] == (float)om.pose_x && xy_yaw[1] == (float)om.pose_y && xy_yaw[2] == (float)om.pose_yaw &&
                   det.last_score() == om.score && scale == 1.);
    std::vector<std::vector<float>> qf = {db[123], db[17]};
    qf[0][3] += 2e-3f;
    const auto located = det.localize(qf, {qg, qg}, false, 20, 10, 0.02, 0.3f);
    EXPECT(located.size() == 2 && located[0].loop_indices.size() == K);
    for (int qi2 = 0; qi2 < 2; ++qi2) {
      gloc_oracle_knn(flat.data(), 401, D, qf[qi2].data(), 1, K, oi.data(), od.data());
      // (row 400 is the near-copy of 123 appended above; the flat copy holds the original rows only)
      EXPECT(located[qi2].matched == (om.found != 0));       // every keyframe carries the same grid
      if (located[qi2].matched)
        EXPECT(located[qi2].located_db_idx == located[qi2].loop_indices[0] && located[qi2].score == om.score &&
               located[qi2].xy_yaw[0] == (float)om.pose_x && located[qi2].xy_yaw[2] == (float)om.pose_yaw);
    }
    EXPECT(located[1].loop_indices[0] == 17 && located[1].out_dists_sqr[0] == 0.f);
    const auto first = det.localize(qf, {qg, qg}, true, 20, 10, 0.02, 0.3f);
    EXPECT(first[0].matched == located[0].matched && first[0].located_db_idx == located[0].located_db_idx);
  }

  // ---- BEV projection through the loop detector's own interface vs the oracle
  {
    std::vector<float> scan;
    std::uniform_real_distribution<float> u(-60.f, 60.f), uz(-1.5f, 2.f);
    for (int w = 0; w < 30; ++w) {
      const float x0 = u(rng), y0 = u(rng), dx = u(rng) / 60.f, dy = u(rng) / 60.f;
      for (int t = 0; t < 200; ++t) {
        scan.push_back(x0 + dx * t * 0.1f);
        scan.push_back(y0 + dy * t * 0.1f);
        scan.push_back(uz(rng));
        scan.push_back(0.5f);
      }
    }
    float xy_res[3];
    const BevImage img = det.get_projected_grid(scan.data(), scan.size() / 4, 4, xy_res);
    int w = 0, h = 0, mix = 0, miy = 0;
    double ox = 0, oy = 0;
    size_t nv = 0, no = 0;
    gloc_oracle_bev_project(scan.data(), scan.size() / 4, 4, 0.2f, 100.f, nullptr, 0, &w, &h, &mix, &miy, &ox, &oy, &nv, &no);
    std::vector<uint8_t> ref((size_t)w * h);
    gloc_oracle_bev_project(scan.data(), scan.size() / 4, 4, 0.2f, 100.f, ref.data(), ref.size(), &w, &h, &mix, &miy, &ox, &oy, &nv, &no);
    EXPECT(img.rows == h && img.cols == w && img.data == ref && no > 100);
    EXPECT(xy_res[0] == (float)ox && xy_res[1] == (float)oy && xy_res[2] == 0.2f);
    const BevImage cnn = det.crop_pad_occupancy(768, 768);
    std::vector<uint8_t> cref((size_t)768 * 768);
    gloc_oracle_crop_pad(ref.data(), w, h, 768, 768, cref.data());
    EXPECT(cnn.data == cref);
  }
  std::printf("PASS host mirrors\n");
  return 0;
}
