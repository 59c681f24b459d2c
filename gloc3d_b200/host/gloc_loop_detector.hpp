// gloc_loop_detector.hpp -- the hot-path half of RpyPCLoopDetector
// (/root/reference/registration/loop_detector.{h,cpp}) on top of the GPU path: same member
// names, constants and guards for detect()/match(); descriptors are handed in (the network
// runs through gloc_enc_* / gloc_vlad_*, see tools/global_localization.cpp), BEV grids are
// handed in or come from get_projected_grid() (gloc_bev_*).
//
// NOT the reference's class verbatim: it is named RpyPCLoopDetectorGpu and takes plain views
// (descriptor vectors, Grid2DView, float[3]) where the reference takes pcl::PointCloud::Ptr,
// OccupancyGrid (cv::Mat) and Eigen::Vector3f -- PCL, OpenCV and Eigen are not in this image.
// INTEGRATION.md lists the adapter lines a maintainer adds inside RpyPCLoopDetector to keep its
// own signatures.  The database's grids live in ONE gloc_csm_store (49 KB per KITTI-sized frame);
// localize() is the batched form of detect_all_query + global_registraion
// (global_localization.cpp:482-574) over gloc_loc_localize.
#ifndef GLOC_LOOP_DETECTOR_HPP_
#define GLOC_LOOP_DETECTOR_HPP_

#include <iostream>
#include <memory>
#include <stdexcept>
#include <vector>

#include "gloc_fast_csm_2d.hpp"
#include "gloc_inv_key_tree.hpp"

// cv::Mat stand-in for the BEV occupancy image (CV_8UC1: 0 = occupied, 255 = free).
struct BevImage {
  int rows = 0, cols = 0;
  std::vector<uint8_t> data;   // row-major
  uint8_t at(int r, int c) const { return data[(size_t)r * cols + c]; }
};

class RpyPCLoopDetectorGpu {
 public:
  using Grid2DView = cartographer::mapping::Grid2DView;
  using Matcher = cartographer::mapping::scan_matching::FastCorrelativeScanMatcher2D;
  using Options = cartographer::mapping::scan_matching::FastCorrelativeScanMatcherOptions2D;
  using Rigid2d = cartographer::mapping::scan_matching::Rigid2d;
  using SearchParameters = cartographer::mapping::scan_matching::SearchParameters;

  const int NUM_EXCLUDE_RECENT = 30;  // loop_detector.h:77

  explicit RpyPCLoopDetectorGpu(int device = 0) : device_(device) {}
  ~RpyPCLoopDetectorGpu() {
    gloc_loc_destroy(loc_);
    gloc_csm_destroy(store_);
    gloc_bev_destroy(bev_);
  }
  RpyPCLoopDetectorGpu(const RpyPCLoopDetectorGpu&) = delete;
  RpyPCLoopDetectorGpu& operator=(const RpyPCLoopDetectorGpu&) = delete;

  // get_projected_grid (loop_detector.cpp:122-135): one scan (n points, `stride` floats apart,
  // x y z first) -> BEV occupancy image; xy_res receives (ox, oy, resolution).
  BevImage get_projected_grid(const float* points, size_t n, int stride, float xy_res[3]) {
    ensure_bev();
    gloc_bev_info info;
    csm_check(gloc_bev_project(bev_, points, n, stride, &info));
    BevImage img;
    img.rows = info.height;
    img.cols = info.width;
    img.data.resize((size_t)info.width * info.height);
    csm_check(gloc_bev_get_image(bev_, img.data.data(), img.data.size()));
    xy_res[0] = static_cast<float>(info.ox);
    xy_res[1] = static_cast<float>(info.oy);
    xy_res[2] = static_cast<float>(info.resolution);
    return img;
  }
  // crop_pad_occupancy (loop_detector.cpp:83-106) of the last projected grid, one channel.
  BevImage crop_pad_occupancy(size_t width, size_t height) {
    ensure_bev();
    BevImage img;
    img.rows = (int)height;
    img.cols = (int)width;
    img.data.resize(width * height);
    csm_check(gloc_bev_get_cnn_input(bev_, (int)width, (int)height, img.data.data()));
    return img;
  }

  // add_keyframe (loop_detector.cpp:9-20) with the descriptor and the BEV grid already computed.
  // `cells` must stay alive as long as the detector (the reference keeps cv::Mat copies).
  void add_keyframe(const std::vector<float>& feat, const Grid2DView& grid) {
    db_features_.push_back(feat);   // the grid goes to the device with the first match (sync_store)
    db_grids_.push_back(grid);
    if (kdtree_) kdtree_->append(feat);
  }

  // for global localization (loop_detector.cpp:22-46)
  void detect(const std::vector<float>& q_feat, std::vector<size_t>& loop_indices,
              std::vector<float>& out_dists_sqr) {
    if (db_features_.size() <= num_exclude_recent_ + top_k_) {
      std::cout << "Not enough keyframes in database." << std::endl;
      return;
    }
    ensure_tree();
    kdtree_->set_search_limit(db_features_.size());
    loop_indices.resize(top_k_);
    out_dists_sqr.resize(top_k_);
    kdtree_->query(&q_feat[0], top_k_, &loop_indices[0], &out_dists_sqr[0]);
  }

  // for slam, using the last frame as query (loop_detector.cpp:48-81): searches all but the
  // NUM_EXCLUDE_RECENT most recent keyframes; the reference rebuilds its KD-tree every 30
  // calls over that set, the GPU index just moves its search limit.
  bool detect(size_t& q_idx, size_t& loop_idx) {
    if (db_features_.size() <= num_exclude_recent_ + top_k_) return false;
    ensure_tree();
    if (tree_making_period_counter_ % tree_making_period_ == 0)
      search_rows_ = db_features_.size() - num_exclude_recent_;
    tree_making_period_counter_ = tree_making_period_counter_ + 1;
    kdtree_->set_search_limit(search_rows_);
    const size_t cur_idx = db_features_.size() - 1;
    std::vector<size_t> ret_indexes(top_k_);
    std::vector<float> out_dists_sqr(top_k_);
    kdtree_->query(&db_features_[cur_idx][0], top_k_, &ret_indexes[0], &out_dists_sqr[0]);
    if (out_dists_sqr[0] < loop_metric_dist_th_) {
      q_idx = cur_idx;
      loop_idx = ret_indexes[0];
      return true;
    }
    return false;
  }

  // match (loop_detector.h:73-75): verify candidate db_idx against the query grid by
  // correlative / branch-and-bound scan matching of the BEV grids (the north-star verifier
  // that takes the place of the SURF + RANSAC body, loop_detector.cpp:183-288).
  // xy_yaw receives (x, y, yaw); estimated_scale is 1 (rigid).
  bool match(const Grid2DView& q_grid, const size_t db_idx, float xy_yaw[3], double& estimated_scale,
             int n_lin = 100, int n_ang = 180, double ang_step = 2. * M_PI / 360., float min_score = 0.3f) {
    if (db_idx >= db_grids_.size()) return false;
    const auto cloud = Matcher::GridToVirtualPointCloud(q_grid);
    if (cloud.empty()) return false;
    sync_store();
    const int64_t offs[2] = {0, (int64_t)cloud.size()};
    const int gid = (int)db_idx, sid = 0;
    const double init[3] = {0., 0., 0.};
    gloc_csm_result r;
    csm_check(gloc_csm_match_batch(store_, cloud[0].data(), offs, 1, &gid, &sid, init, 1, n_lin, n_ang, ang_step,
                                   options_.branch_and_bound_depth(), min_score, &r));
    if (!r.found) return false;
    xy_yaw[0] = static_cast<float>(r.pose_x);
    xy_yaw[1] = static_cast<float>(r.pose_y);
    xy_yaw[2] = static_cast<float>(r.pose_yaw);
    estimated_scale = 1.;
    last_score_ = r.score;
    return true;
  }

  // detect_all_query + global_registraion (global_localization.cpp:482-574) for a batch of queries
  // in ONE call: top-k retrieval, the k candidates' grids gathered on the device, every candidate
  // (or, first_match_only, the candidates in retrieval order until one matches -- the reference's
  // loop) verified, the located keyframe and the pose of the query in it per query.
  struct Located {
    bool matched = false;      // global_registraion's return value
    size_t located_db_idx = 0;
    float xy_yaw[3] = {0.f, 0.f, 0.f};
    float score = 0.f;
    std::vector<size_t> loop_indices;     // queried_idx_[q]
    std::vector<float> out_dists_sqr;
  };
  std::vector<Located> localize(const std::vector<std::vector<float>>& q_feats,
                                const std::vector<Grid2DView>& q_grids, bool first_match_only = false,
                                int n_lin = 100, int n_ang = 180, double ang_step = 2. * M_PI / 360.,
                                float min_score = 0.3f) {
    std::vector<Located> out(q_feats.size());
    if (q_feats.empty() || db_features_.size() <= num_exclude_recent_ + top_k_) return out;   // loop_detector.cpp:27-30
    if (q_grids.size() != q_feats.size()) throw std::invalid_argument("one BEV grid per query");
    ensure_tree();
    sync_store();
    kdtree_->set_search_limit(db_features_.size());
    if (!loc_) csm_check(gloc_loc_create(&loc_, kdtree_->index, store_));
    std::vector<float> q(q_feats.size() * k_dim_), pts;
    std::vector<int64_t> offs(q_feats.size() + 1, 0);
    for (size_t i = 0; i < q_feats.size(); ++i) {
      std::copy(q_feats[i].begin(), q_feats[i].end(), q.begin() + i * k_dim_);
      const auto cloud = Matcher::GridToVirtualPointCloud(q_grids[i]);
      if (cloud.empty()) throw std::invalid_argument("a query grid without occupied cells");
      for (const auto& p : cloud) pts.insert(pts.end(), p.begin(), p.end());
      offs[i + 1] = (int64_t)(pts.size() / 3);
    }
    gloc_loc_params prm;
    prm.k = (int)top_k_;
    prm.n_lin = n_lin;
    prm.n_ang = n_ang;
    prm.ang_step = ang_step;
    prm.depth = options_.branch_and_bound_depth();
    prm.min_score = min_score;
    prm.policy = first_match_only ? GLOC_LOC_FIRST_MATCH : GLOC_LOC_VERIFY_ALL;
    std::vector<uint64_t> idx(q_feats.size() * top_k_);
    std::vector<float> d2(q_feats.size() * top_k_);
    std::vector<gloc_loc_result> res(q_feats.size());
    csm_check(gloc_loc_localize(loc_, q.data(), q_feats.size(), pts.data(), offs.data(), nullptr, &prm, idx.data(),
                                d2.data(), nullptr, res.data()));
    for (size_t i = 0; i < q_feats.size(); ++i) {
      Located& L = out[i];
      L.loop_indices.assign(idx.begin() + i * top_k_, idx.begin() + (i + 1) * top_k_);
      L.out_dists_sqr.assign(d2.begin() + i * top_k_, d2.begin() + (i + 1) * top_k_);
      L.matched = res[i].located != 0;
      if (L.matched) {
        L.located_db_idx = (size_t)res[i].db_index;
        L.xy_yaw[0] = static_cast<float>(res[i].match.pose_x);
        L.xy_yaw[1] = static_cast<float>(res[i].match.pose_y);
        L.xy_yaw[2] = static_cast<float>(res[i].match.pose_yaw);
        L.score = res[i].match.score;
      }
    }
    return out;
  }
  float last_score() const { return last_score_; }

 private:
  static void csm_check(int rc) {
    if (rc != GLOC_OK) throw std::runtime_error(gloc_last_error());
  }
  void ensure_bev() {   // high_resolution_ = 0.2, high_resolution_max_range_ = 100 (loop_detector.h:111-116)
    if (!bev_) csm_check(gloc_bev_create(&bev_, device_, 0.2f, 100.f));
  }
  void sync_store() {   // db_grids_[i] = grid i of the store (loop_detector.h:36-39)
    if (!store_) csm_check(gloc_csm_create(&store_, device_));
    for (size_t i = (size_t)gloc_csm_num_grids(store_); i < db_grids_.size(); ++i) {
      const Grid2DView& g = db_grids_[i];
      int gid = -1;
      csm_check(gloc_csm_add_grid_cells(store_, g.correspondence_cost_cells, g.limits.num_x_cells, g.limits.num_y_cells,
                                        g.limits.resolution, g.limits.max_x, g.limits.max_y, &gid));
    }
  }
  void ensure_tree() {
    if (!kdtree_) kdtree_ = std::make_unique<InvKeyTree>(k_dim_, db_features_, 10, device_);
  }
  const size_t k_dim_ = 512;                 // loop_detector.h:97-103
  const size_t top_k_ = 20;
  const size_t num_exclude_recent_ = 30;
  const size_t tree_making_period_ = 30;
  size_t tree_making_period_counter_ = 0;
  const float loop_metric_dist_th_ = 0.8f;
  size_t search_rows_ = 0;
  int device_;
  float last_score_ = 0.f;
  KeyMat db_features_;
  std::unique_ptr<InvKeyTree> kdtree_;
  std::vector<Grid2DView> db_grids_;
  Options options_;
  gloc_bev_projector* bev_ = nullptr;
  gloc_csm_store* store_ = nullptr;    // every keyframe's grid, bit-packed, resident
  gloc_localizer* loc_ = nullptr;
};

#endif  // GLOC_LOOP_DETECTOR_HPP_
