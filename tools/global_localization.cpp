// global_localization -- the reference's evaluation driver on top of libgloc3d.so.
//
//   global_localization VALSET GT_POSE MODEL [any 4th argument]
//
// Same command line, same input formats, same log lines and the same two result files as
// /root/reference/registration/global_localization.cpp (main :577-600): VALSET and GT_POSE
// exactly as dataset/kitti_i2i.py:76-122 writes them, scans as raw float32 x y z i.  The hot
// path runs on the GPU through the C ABI (include/gloc3d.h):
//   BEV projection         gloc_bev_*                (get_projected_grid, loop_detector.cpp:122-135)
//   retrieval, k = 20      gloc_knn_*                (InvKeyTree::query, loop_detector.cpp:34-45)
//   verification           gloc_csm_match_batch      (the slot of loop_detector_.match, :519-524)
// What is NOT part of the query path stays outside:
//   * MODEL: the reference loads a TorchScript CNN here (loop_detector.cpp:157-163), which needs
//     libtorch.  Two forms are accepted instead:
//       - the network's weights in the plain container tools/export_weights.py writes from that
//         TorchScript file ("GLOCW001": 13 VGG16 convolutions + NetVLAD_fc): descriptors are then
//         computed like get_place_feature does (loop_detector.cpp:137-172) -- BEV image, crop/pad
//         to 768 x 768, encoder, pooling head -- through gloc_desc_extract_padded;
//       - a raw float32 table with (db_num + q_num) x 512 descriptors in valset order (what that
//         forward would produce), for runs without a network.
//   * the 4th argument switches ground alignment on like the reference's (:419-449, :482-509,
//     :524-569): every scan is levelled on the host by gloc::GroundEstimator before the BEV
//     projection and the located pose is composed from the 2-D match and the two ground
//     transforms (gloc3d_b200/host/gloc_ground.hpp -- the reference does this on the CPU with
//     PCL; the plane fit there is a RANSAC whose samples cannot be reproduced without PCL).
//   * GLOC_GRID_STORE=<file> keeps the database's BEV grids in a grid store file between runs.
// Own code: a small logger that prints glog-style lines, a reader for each format.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../gloc3d_b200/host/gloc_ground.hpp"
#include "../include/gloc3d.h"

namespace {

// ---- glog-style lines: "I1018 12:34:56.789012 global_localization.cpp:123] message"
struct LogLine {
  std::ostringstream os;
  LogLine(char sev, int line) {
    using namespace std::chrono;
    const auto now = system_clock::now();
    const std::time_t t = system_clock::to_time_t(now);
    const long us = (long)(duration_cast<microseconds>(now.time_since_epoch()).count() % 1000000);
    std::tm tm{};
    localtime_r(&t, &tm);
    char buf[64];
    std::snprintf(buf, sizeof buf, "%c%02d%02d %02d:%02d:%02d.%06ld global_localization.cpp:%d] ", sev,
                  tm.tm_mon + 1, tm.tm_mday, tm.tm_hour, tm.tm_min, tm.tm_sec, us, line);
    os << buf;
  }
  ~LogLine() { std::cerr << os.str() << std::endl; }
};
#define LOG_INFO LogLine('I', __LINE__).os
#define LOG_ERROR LogLine('E', __LINE__).os

struct TicToc {  // registration/tic_toc.h: wall-clock milliseconds
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  double toc() const {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
};

void check(int rc, const char* what) {
  if (rc != GLOC_OK) {
    LOG_ERROR << what << ": " << gloc_last_error();
    std::exit(2);
  }
}

std::vector<std::string> split(const std::string& s, const std::string& sep) {  // global_localization.cpp:40-62
  std::vector<std::string> res;
  if (s.empty()) return res;
  size_t start = 0;
  for (;;) {
    const size_t at = s.find(sep, start);
    if (at == std::string::npos) {
      if (start < s.size()) res.push_back(s.substr(start));
      break;
    }
    if (at > start) res.push_back(s.substr(start, at - start));
    start = at + sep.size();
  }
  return res;
}

using Mat4 = gloc::Mat4f;
using gloc::mul;
using gloc::pose_from;
using gloc::rigid_inverse;

// ReadValset, global_localization.cpp:64-122
bool ReadValset(const std::string& filename, std::vector<std::string>& db_files,
                std::vector<std::string>& q_files, std::vector<std::vector<size_t>>& pos_idx) {
  std::ifstream ifs(filename);
  db_files.clear();
  q_files.clear();
  pos_idx.clear();
  if (!ifs.is_open()) {
    std::cout << "failed to open file " << filename << "\n";
    return false;
  }
  std::string line;
  std::getline(ifs, line);
  std::vector<std::string> sub = split(line, " ");
  if (sub.size() < 2) return false;
  const int db_num = std::atoi(sub[0].c_str()), q_num = std::atoi(sub[1].c_str());
  for (int i = 0; i < db_num; ++i) {
    std::getline(ifs, line);
    db_files.push_back(line);
  }
  for (int i = 0; i < q_num; ++i) {
    std::getline(ifs, line);
    q_files.push_back(line);
  }
  for (int i = 0; i < q_num; ++i) {
    if (!std::getline(ifs, line)) break;
    if (line.empty()) break;
    sub = split(line, ":");
    if (sub.size() == 1) {
      pos_idx.push_back({});
      continue;
    }
    const std::string pos_str = sub[1];
    if (pos_str.empty()) {
      pos_idx.push_back({});
      continue;
    }
    sub = split(pos_str, " ");
    std::vector<size_t> tmp;
    for (const auto& t : sub) tmp.push_back((size_t)std::atoi(t.c_str()));
    pos_idx.push_back(tmp);
  }
  LOG_INFO << "db_num and db_files: " << db_num << ", " << db_files.size();
  LOG_INFO << "q_num and q_files: " << q_num << ", " << q_files.size();
  LOG_INFO << "q_num and q_pos_index: " << q_num << ", " << pos_idx.size();
  return true;
}

// ReadValsetPose, global_localization.cpp:124-156: "qx qy qz qw x y z" per line
bool ReadValsetPose(const std::string& filename, std::vector<Mat4>& poses) {
  std::ifstream ifs(filename);
  poses.clear();
  if (!ifs.is_open()) {
    LOG_ERROR << "failed to open file " << filename;
    return false;
  }
  std::string line;
  while (std::getline(ifs, line)) {
    const std::vector<std::string> sub = split(line, " ");
    if (sub.size() != 7) {
      LOG_ERROR << "Check failed: substrs.size()==7";
      std::exit(1);
    }
    poses.push_back(pose_from((float)std::atof(sub[3].c_str()), (float)std::atof(sub[0].c_str()),
                              (float)std::atof(sub[1].c_str()), (float)std::atof(sub[2].c_str()),
                              (float)std::atof(sub[4].c_str()), (float)std::atof(sub[5].c_str()),
                              (float)std::atof(sub[6].c_str())));
  }
  LOG_INFO << "Read poses with size: " << poses.size();
  return true;
}

// read_lidar_data, global_localization.cpp:160-182: raw float32 x y z i
std::vector<float> read_lidar_data(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  std::vector<float> buf;
  if (!f) return buf;
  f.seekg(0, std::ios::end);
  const size_t n = (size_t)f.tellg() / sizeof(float);
  f.seekg(0, std::ios::beg);
  buf.resize(n / 4 * 4);
  f.read(reinterpret_cast<char*>(buf.data()), (std::streamsize)(buf.size() * sizeof(float)));
  return buf;
}

void caculate_mean_std(const std::vector<double>& v, double& mean, double& stddev) {
  mean = 0.;
  stddev = 0.;
  if (v.empty()) return;
  for (double x : v) mean += x;
  mean /= (double)v.size();
  for (double x : v) stddev += (x - mean) * (x - mean);
  stddev = std::sqrt(stddev / (double)v.size());
}

class GlocEvaluator {
 public:
  bool align_ground_ = false;

  bool load_valset(const std::string& q_db_file, const std::string& pose_file) {
    return ReadValset(q_db_file, db_files_, q_files_, gt_q_pos_idx_) && ReadValsetPose(pose_file, poses_db_q_);
  }

  // GLOC_DRIVER_PARSE_ONLY=1: read every input the run would read (valset, poses, descriptor
  // table, every scan file; with the 4th argument also level every scan) and report what was
  // found, without touching the GPU.
  bool check_inputs(const std::string& model_file_path) {
    load_descriptors(model_file_path);
    size_t n_pts = 0, missing = 0;
    std::vector<std::string> all = db_files_;
    all.insert(all.end(), q_files_.begin(), q_files_.end());
    size_t levelled = 0;
    double height = 0.;
    for (const auto& f : all) {
      std::vector<float> pc = read_lidar_data(f);
      if (pc.empty()) ++missing;
      n_pts += pc.size() / 4;
      if (align_ground_ && !pc.empty()) {   // the host half of the aligned run: level every scan
        const Mat4 T = align_to_ground(&pc);
        if (T.m[2][3] > 0.f) {
          ++levelled;
          height += T.m[2][3];
        }
      }
    }
    if (align_ground_)
      LOG_INFO << "ground alignment: " << levelled << " of " << all.size() << " scans levelled, mean sensor height "
               << (levelled ? height / (double)levelled : 0.) << " m";
    size_t n_pos = 0;
    for (const auto& v : gt_q_pos_idx_) n_pos += v.size();
    LOG_INFO << "inputs: " << db_files_.size() << " db scans, " << q_files_.size() << " query scans, "
             << poses_db_q_.size() << " poses, " << n_pos << " positives, "
             << (have_network_ ? std::string("network") : std::to_string(feats_.size() / kDim) + " descriptors") << ", " << n_pts << " points, " << missing << " unreadable scans";
    return missing == 0 && poses_db_q_.size() == all.size();
  }

  // construct_db, global_localization.cpp:419-449
  void construct_db(const std::string& model_file_path) {
    load_descriptors(model_file_path);
    LOG_INFO << "LOAD descriptor table (stand-in for the libtorch model) from: " << model_file_path;
    check(gloc_bev_create(&bev_, device_, 0.2f, 100.f), "gloc_bev_create");
    check(gloc_csm_create(&store_, device_), "gloc_csm_create");
    check(gloc_knn_create(&index_, kDim, device_), "gloc_knn_create");
    if (have_network_) create_network();
    int i = 0;
    double t_align = 0., t_detect = 0.;
    // GLOC_GRID_STORE=<file>: the database's BEV grids on disk (grid store file, include/gloc3d.h):
    // loaded when the file holds one grid per database scan, written after the projection pass
    // otherwise.  Not used with ground alignment (the per-scan ground transforms are not in it).
    // (a grid store holds no descriptors: with a network every database scan is projected anyway)
    const char* grid_store = (align_ground_ || have_network_) ? nullptr : std::getenv("GLOC_GRID_STORE");
    // what the stored grids were made from: the database scan list and the projection parameters
    uint32_t store_tag = 2166136261u;   // FNV-1a
    auto mix = [&](const void* p, size_t n) {
      for (size_t b = 0; b < n; ++b) store_tag = (store_tag ^ ((const unsigned char*)p)[b]) * 16777619u;
    };
    for (const auto& f : db_files_) mix(f.data(), f.size() + 1);
    const float proj_params[2] = {0.2f, 100.f};
    mix(proj_params, sizeof proj_params);
    if (store_tag == 0) store_tag = 1;
    if (grid_store) {
      gloc_grid_file* gf = nullptr;
      size_t n_stored = 0;
      if (gloc_grid_file_open(grid_store, &gf, &n_stored) == GLOC_OK) {
        const uint32_t tag = gloc_grid_file_tag(gf);
        gloc_grid_file_close(gf);
        if (n_stored == db_files_.size() && tag != store_tag)
          LOG_INFO << grid_store << " was built from another database or projection (fingerprint " << tag
                   << ", expected " << store_tag << "): rebuilding";
        if (n_stored == db_files_.size() && tag == store_tag) {
          int first = 0, n = 0;
          check(gloc_csm_load_grids(store_, grid_store, &first, &n), "gloc_csm_load_grids");
          for (int k = 0; k < n; ++k) db_grid_ids_.push_back(first + k);
          check(gloc_knn_set_db(index_, feats_.data(), db_files_.size()), "gloc_knn_set_db");
          LOG_INFO << "loaded " << n << " database grids from " << grid_store;
          return;
        }
        LOG_INFO << grid_store << " holds " << n_stored << " grids for " << db_files_.size() << " scans: rebuilding";
      }
    }
    for (const auto& filename : db_files_) {
      ++i;
      std::vector<float> kf = read_lidar_data(filename);
      if (align_ground_) {
        TicToc ta;
        db_rpz_estimates_.push_back(align_to_ground(&kf));
        t_align += ta.toc();
      }
      TicToc tb;
      gloc_bev_info info;
      check(gloc_bev_project(bev_, kf.data(), kf.size() / 4, 4, &info), "gloc_bev_project");
      int gid = -1;
      check(gloc_csm_add_grid_from_bev_aligned(store_, bev_, &gid), "gloc_csm_add_grid_from_bev_aligned");
      db_grid_ids_.push_back(gid);
      if (have_network_) {
        push_plane();
        if (i % kCnnBatch == 0) flush_planes((size_t)(i - kCnnBatch));
      }
      if (i > 2) t_detect += tb.toc();
    }
    if (have_network_) flush_planes(db_files_.size() - planes_.size() / ((size_t)kCnnSide * kCnnSide));
    check(gloc_knn_set_db(index_, feats_.data(), db_files_.size()), "gloc_knn_set_db");
    if (grid_store) {
      check(gloc_csm_save_grids_tagged(store_, grid_store, store_tag), "gloc_csm_save_grids_tagged");
      LOG_INFO << "wrote " << db_grid_ids_.size() << " database grids to " << grid_store;
    }
    LOG_INFO << "time cost for align to ground: " << t_align / double(i) << "ms.";
    LOG_INFO << "time cost for feature extraction: " << t_detect / double(i - 2) << "ms.";
  }

  void locate_all_query() {
    if (queried_idx_.empty()) detect_all_query();
    located_db_.clear();
    located_pose_.clear();
    global_registraion_all();
  }

  // detect_all_query, global_localization.cpp:482-509 + RpyPCLoopDetector::detect, loop_detector.cpp:22-46
  void detect_all_query() {
    double t_sum = 0.;
    const size_t n_db = db_files_.size();
    for (size_t qi = 0; qi < q_files_.size(); ++qi) {
      std::vector<float> q_pc = read_lidar_data(q_files_[qi]);
      if (align_ground_) q_rpz_estimates_.push_back(align_to_ground(&q_pc));
      TicToc t;
      std::vector<size_t> loop_indices;
      gloc_bev_info info;
      check(gloc_bev_project(bev_, q_pc.data(), q_pc.size() / 4, 4, &info), "gloc_bev_project");
      size_t n_occ = 0;
      check(gloc_bev_get_occupied_points(bev_, nullptr, 0, &n_occ), "gloc_bev_get_occupied_points");
      std::vector<float> pts(n_occ * 3);
      if (n_occ) check(gloc_bev_get_occupied_points(bev_, pts.data(), n_occ, &n_occ), "gloc_bev_get_occupied_points");
      if (have_network_) {                      // the query's own descriptor, one frame like the reference
        push_plane();
        flush_planes(n_db + qi);
      }
      if (n_db <= kNumExcludeRecent + kTopK) {  // loop_detector.cpp:27-30
        std::cout << "Not enough keyframes in database." << std::endl;
      } else {
        std::vector<uint64_t> idx(kTopK);
        std::vector<float> d2(kTopK);
        check(gloc_knn_query(index_, feats_.data() + (n_db + qi) * kDim, 1, kTopK, idx.data(), d2.data()),
              "gloc_knn_query");
        loop_indices.assign(idx.begin(), idx.end());
      }
      t_sum += t.toc();
      q_scans_.push_back(pts);
      queried_idx_.push_back(loop_indices);
    }
    LOG_INFO << "Each query cost: " << t_sum / q_files_.size() << "ms.";
  }

  // recognition_recalls, global_localization.cpp:221-268
  void recognition_recalls() {
    const std::vector<int> k_values = {1, 5, 10, 20};
    std::vector<float> k_recalls = {0.f, 0.f, 0.f, 0.f};
    int valid_query_num = 0;
    for (size_t i = 0; i < q_files_.size(); ++i) {
      if (i >= gt_q_pos_idx_.size() || gt_q_pos_idx_[i].empty()) continue;
      valid_query_num++;
      if (queried_idx_[i].empty()) {
        failed_detect_indices_.push_back((int)i);
        continue;
      }
      const std::vector<size_t>& cand = queried_idx_[i];
      bool detected = false;
      for (size_t k = 0; k < k_values.size(); ++k) {
        for (int j = 0; j < k_values[k] && j < (int)cand.size(); ++j) {
          if (std::find(gt_q_pos_idx_[i].begin(), gt_q_pos_idx_[i].end(), cand[j]) != gt_q_pos_idx_[i].end()) {
            k_recalls[k] += 1;
            detected = true;
            break;
          }
        }
      }
      if (!detected) failed_detect_indices_.push_back((int)i);
    }
    if (valid_query_num > 0) {
      for (size_t i = 0; i < k_recalls.size(); ++i) {
        k_recalls[i] /= valid_query_num;
        LOG_INFO << "Recall @ " << k_values[i] << ": " << k_recalls[i];
      }
    }
    write_indices("failed_detect_indices.txt", failed_detect_indices_);
  }

  // registration_recalls, global_localization.cpp:270-335
  void registration_recalls() {
    const int all_tests = (int)located_db_.size();
    int succeed_tests = 0;
    std::vector<double> rot_err, pos_err;
    const size_t num_db = db_files_.size();
    for (size_t i = 0; i < located_db_.size(); ++i) {
      const size_t db_idx = located_db_[i];
      if (db_idx >= db_files_.size()) {
        failed_registration_indices_.push_back((int)i);
        continue;
      }
      float err_rot, err_pos;
      gloc::RegistrationError(poses_db_q_[db_idx], poses_db_q_[i + num_db], located_pose_[i], &err_rot, &err_pos);
      if (err_pos < 1.0f && err_rot < 5.f) {
        succeed_tests++;
        rot_err.push_back(err_rot);
        pos_err.push_back(err_pos);
      }
    }
    double mean_rot, std_rot, mean_pos, std_pos;
    caculate_mean_std(pos_err, mean_pos, std_pos);
    caculate_mean_std(rot_err, mean_rot, std_rot);
    LOG_INFO << succeed_tests << ", " << all_tests;
    LOG_INFO << "Success rate: " << static_cast<float>(succeed_tests) / static_cast<float>(all_tests);
    LOG_INFO << "Rot error: " << mean_rot << ", " << std_rot;
    LOG_INFO << "Pos error: " << mean_pos << ", " << std_pos;
    write_indices("failed_registration_indices.txt", failed_registration_indices_);
    LOG_INFO << "Average 2D match costs " << time_sum_match_ / times_call_match_ << "ms.";
  }

  ~GlocEvaluator() {
    gloc_knn_destroy(index_);
    gloc_enc_destroy(enc_);
    gloc_vlad_destroy(head_);
    gloc_csm_destroy(store_);
    gloc_bev_destroy(bev_);
  }

 private:
  static constexpr size_t kDim = 512, kTopK = 20, kNumExcludeRecent = 30;  // loop_detector.h:97-100
  static constexpr int kCnnSide = 768, kCnnBatch = 16;                     // loop_detector.cpp:144-145

  // "GLOCW001" container (gloc3d_b200/weights.py): name -> (shape, float32 data)
  struct Array {
    std::vector<uint64_t> shape;
    std::vector<float> data;
  };
  static bool read_weight_file(const std::string& path, std::map<std::string, Array>* out) {
    std::ifstream f(path, std::ios::binary);
    char magic[8];
    if (!f || !f.read(magic, 8) || std::memcmp(magic, "GLOCW001", 8) != 0) return false;
    uint32_t n = 0;
    f.read(reinterpret_cast<char*>(&n), 4);
    for (uint32_t i = 0; i < n && f; ++i) {
      uint16_t ln = 0;
      f.read(reinterpret_cast<char*>(&ln), 2);
      std::string name(ln, '\0');
      f.read(&name[0], ln);
      uint32_t nd = 0;
      f.read(reinterpret_cast<char*>(&nd), 4);
      if (nd > 8) return false;
      Array a;
      a.shape.resize(nd);
      f.read(reinterpret_cast<char*>(a.shape.data()), 8 * nd);
      uint64_t cnt = 1;
      for (uint64_t d : a.shape) cnt *= d;
      if (cnt > (1ull << 32)) return false;
      a.data.resize(cnt);
      f.read(reinterpret_cast<char*>(a.data.data()), (std::streamsize)(4 * cnt));
      if (!f) return false;
      (*out)[name] = std::move(a);
    }
    return (bool)f;
  }

  // MODEL = weight container: keep the arrays; returns false when MODEL is something else
  bool load_network(const std::string& path) {
    std::map<std::string, Array> w;
    if (!read_weight_file(path, &w)) return false;
    static const uint64_t cout[13] = {64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512};
    uint64_t cin = 3;
    for (int l = 0; l < 13; ++l) {
      const auto wi = w.find("enc." + std::to_string(l) + ".weight"), bi = w.find("enc." + std::to_string(l) + ".bias");
      if (wi == w.end() || bi == w.end() || wi->second.shape != std::vector<uint64_t>{cout[l], cin, 3, 3} ||
          bi->second.shape != std::vector<uint64_t>{cout[l]}) {
        LOG_ERROR << "MODEL: layer " << l << " missing or of the wrong shape";
        std::exit(1);
      }
      cin = cout[l];
    }
    const auto cw = w.find("vlad.conv.weight"), ce = w.find("vlad.centroids"), hi = w.find("vlad.hidden");
    if (cw == w.end() || ce == w.end() || hi == w.end() || cw->second.shape.size() != 2 || cw->second.shape[1] != 512 ||
        ce->second.shape != cw->second.shape || hi->second.shape.size() != 2 ||
        hi->second.shape[0] != cw->second.shape[0] * 512 || hi->second.shape[1] != kDim) {
      LOG_ERROR << "MODEL: NetVLAD_fc arrays missing or of the wrong shape";
      std::exit(1);
    }
    net_ = std::move(w);
    have_network_ = true;
    LOG_INFO << "MODEL: 13 convolutions + NetVLAD_fc (" << net_["vlad.conv.weight"].shape[0] << " clusters, "
             << net_["vlad.hidden"].shape[1] << "-d) from " << path;
    return true;
  }

  void create_network() {   // needs the GPU
    const float* cw[13];
    const float* cb[13];
    for (int l = 0; l < 13; ++l) {
      cw[l] = net_["enc." + std::to_string(l) + ".weight"].data.data();
      cb[l] = net_["enc." + std::to_string(l) + ".bias"].data.data();
    }
    check(gloc_enc_create(&enc_, device_, kCnnSide, kCnnSide, cw, cb), "gloc_enc_create");
    const auto bias = net_.find("vlad.conv.bias");
    check(gloc_vlad_create(&head_, device_, 512, (int)net_["vlad.conv.weight"].shape[0], (int)kDim,
                           net_["vlad.conv.weight"].data.data(), bias == net_.end() ? nullptr : bias->second.data.data(),
                           net_["vlad.centroids"].data.data(), net_["vlad.hidden"].data.data()),
          "gloc_vlad_create");
    feats_.assign((db_files_.size() + q_files_.size()) * kDim, 0.f);
  }

  // get_place_feature (loop_detector.cpp:137-172) for the planes collected so far: rows first .. of feats_
  void flush_planes(size_t first_row) {
    const int n = (int)(planes_.size() / ((size_t)kCnnSide * kCnnSide));
    if (n == 0) return;
    // padded planes: the reference's canvas padding is (255, 0, 0), loop_detector.cpp:84
    check(gloc_desc_extract_padded(enc_, head_, (int)kDim, planes_.data(), rois_.data(), n,
                                   feats_.data() + first_row * kDim),
          "gloc_desc_extract_padded");
    planes_.clear();
    rois_.clear();
  }
  // the projector's current image, cropped / padded to the CNN input, appended to the batch
  void push_plane() {
    const size_t at = planes_.size();
    planes_.resize(at + (size_t)kCnnSide * kCnnSide);
    int32_t roi[4];
    check(gloc_bev_get_cnn_input_roi(bev_, kCnnSide, kCnnSide, planes_.data() + at, roi), "gloc_bev_get_cnn_input_roi");
    rois_.insert(rois_.end(), roi, roi + 4);
  }

  void load_descriptors(const std::string& path) {
    if (load_network(path)) return;
    const size_t rows = db_files_.size() + q_files_.size();
    std::ifstream f(path, std::ios::binary);
    feats_.assign(rows * kDim, 0.f);
    if (!f || !f.read(reinterpret_cast<char*>(feats_.data()), (std::streamsize)(feats_.size() * sizeof(float)))) {
      LOG_ERROR << "MODEL must hold (db_num + q_num) x 512 float32 descriptors in valset order: " << path;
      std::exit(1);
    }
  }

  // EsitmateGroundAndTransform on one scan, in place; returns T_l2g.  A scan without a usable
  // ground keeps its points and gets the identity (the reference hands an EMPTY cloud to the
  // projection in that case, which has no defined result).
  Mat4 align_to_ground(std::vector<float>* scan) {
    std::vector<float> levelled;
    const Mat4 T = ground_estimator_.EsitmateGroundAndTransform(scan->data(), scan->size() / 4, 4, &levelled);
    if (levelled.empty()) {
      LogLine('W', __LINE__).os << "No valid ground found!";
      return Mat4::identity();
    }
    scan->swap(levelled);
    return T;
  }

  void write_indices(const std::string& name, const std::vector<int>& v) {
    std::ofstream ofs(name, std::ios::out);
    if (!ofs) LOG_ERROR << "Failed open " << name;
    for (int idx : v) ofs << idx << " ";
    ofs << "\n";
  }

  // global_registraion_all / global_registraion, global_localization.cpp:342-356, :511-574: the
  // candidates of a query are verified in retrieval order and the first match wins -- here
  // all of them go to the GPU in one batch and the first success in that order is taken.
  void global_registraion_all() {
    const char* ms = std::getenv("GLOC_MATCH_MIN_SCORE");
    const float min_score = ms ? (float)std::atof(ms) : 0.35f;
    located_db_.assign(q_files_.size(), db_files_.size() + 1);
    located_pose_.assign(q_files_.size(), Mat4::identity());
    for (size_t qi = 0; qi < q_files_.size(); ++qi) {
      const std::vector<size_t>& cand = queried_idx_[qi];
      const int n = (int)std::min(kTopK, cand.size());
      if (n == 0 || q_scans_[qi].empty()) continue;
      std::vector<int> gids(n), sids(n, 0);
      std::vector<double> init(3 * (size_t)n, 0.);
      for (int i = 0; i < n; ++i) gids[i] = db_grid_ids_[cand[i]];
      const int64_t offs[2] = {0, (int64_t)(q_scans_[qi].size() / 3)};
      std::vector<gloc_csm_result> res(n);
      TicToc t;
      check(gloc_csm_match_batch(store_, q_scans_[qi].data(), offs, 1, gids.data(), sids.data(), init.data(), n,
                                 /*n_lin*/ 100, /*n_ang*/ 180, 2. * M_PI / 360., /*depth*/ 5, min_score, res.data()),
            "gloc_csm_match_batch");
      time_sum_match_ += t.toc();
      times_call_match_ += n;
      for (int i = 0; i < n; ++i) {
        if (!res[i].found) continue;
        // global_localization.cpp:524-569
        const float xy_yaw[3] = {(float)res[i].pose_x, (float)res[i].pose_y, (float)res[i].pose_yaw};
        const Mat4 p = align_ground_ ? gloc::ComposeLocatedPose(true, xy_yaw, q_rpz_estimates_[qi], db_rpz_estimates_[cand[i]])
                                     : gloc::ComposeLocatedPose(false, xy_yaw, Mat4::identity(), Mat4::identity());
        located_db_[qi] = cand[i];
        located_pose_[qi] = p;
        break;
      }
    }
  }

  int device_ = 0;
  std::vector<std::string> db_files_, q_files_;
  std::vector<std::vector<size_t>> gt_q_pos_idx_;
  std::vector<Mat4> poses_db_q_;
  std::vector<float> feats_;                    // (db_num + q_num) x 512
  std::vector<int> db_grid_ids_;
  std::vector<std::vector<float>> q_scans_;     // occupied-pixel points of every query BEV
  std::vector<std::vector<size_t>> queried_idx_;
  std::vector<size_t> located_db_;
  std::vector<Mat4> located_pose_;
  gloc::GroundEstimator ground_estimator_;
  std::vector<Mat4> db_rpz_estimates_, q_rpz_estimates_;
  std::vector<int> failed_detect_indices_, failed_registration_indices_;
  double time_sum_match_ = 0., times_call_match_ = 0.;
  gloc_bev_projector* bev_ = nullptr;
  gloc_csm_store* store_ = nullptr;
  gloc_knn_index* index_ = nullptr;
  bool have_network_ = false;
  std::map<std::string, Array> net_;
  gloc_encoder* enc_ = nullptr;
  gloc_vlad_head* head_ = nullptr;
  std::vector<uint8_t> planes_;                 // CNN inputs waiting for the next batch
  std::vector<int32_t> rois_;                   // their image rectangles (x0, y0, w, h)
};

}  // namespace

int main(int argc, char* argv[]) {
  if (argc < 4) {
    std::cerr << "usage: global_localization VALSET GT_POSE MODEL [align_ground]\n";
    return 1;
  }
  GlocEvaluator gloc;
  gloc.align_ground_ = argc == 5;
  if (!gloc.load_valset(argv[1], argv[2])) return 1;
  if (std::getenv("GLOC_DRIVER_PARSE_ONLY")) return gloc.check_inputs(argv[3]) ? 0 : 1;   // no GPU needed
  gloc.construct_db(argv[3]);
  gloc.locate_all_query();
  gloc.recognition_recalls();
  gloc.registration_recalls();
  return 0;
}
