// gloc_ground.hpp -- host side of the reference's ground alignment and 6-DoF pose composition
// (SURVEY 8f rank 4): /root/reference/registration/ground_estimator.{h,cpp} and
// global_localization.cpp:511-574.  Header-only, no dependencies (the reference's versions sit
// on PCL and Eigen, neither of which exists in this image), no GPU: this is caller logic around
// the hot path, CPU code in the reference as well.
//
// Two kinds of code live here and the difference matters for parity:
//   * restated bit-for-intent (deterministic arithmetic spelled out in the reference or in
//     Eigen's documented behaviour): TransformPointsToGround, the Euler-angle extraction with
//     Eigen's range convention (first angle in [0, pi]), RollPitchYaw, Embed3D, the pose
//     composition of GlocEvaluator::global_registraion and its error metric;
//   * same algorithm, own implementation ("parity unpinned": PCL is a third-party dependency
//     absent from /root/reference and from this image): k = 10 nearest-neighbour PCA normals
//     flipped towards the sensor, the 10-degree inclination histogram, and the 3-point RANSAC
//     plane with PCL's adaptive stopping rule.  The sample sequence of PCL's generator cannot
//     be reproduced, so plane coefficients agree with the reference's only within the RANSAC
//     threshold (0.1 m), not bit for bit.
#ifndef GLOC_GROUND_HPP_
#define GLOC_GROUND_HPP_

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <random>
#include <vector>

namespace gloc {

// ------------------------------------------------------------------ small fixed-size algebra
struct Mat3f {
  float m[3][3];
  static Mat3f identity() {
    Mat3f r{};
    for (int i = 0; i < 3; ++i) r.m[i][i] = 1.f;
    return r;
  }
};
struct Mat4f {
  float m[4][4];
  static Mat4f identity() {
    Mat4f r{};
    for (int i = 0; i < 4; ++i) r.m[i][i] = 1.f;
    return r;
  }
  Mat3f rotation() const {
    Mat3f r{};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) r.m[i][j] = m[i][j];
    return r;
  }
};
inline Mat4f mul(const Mat4f& a, const Mat4f& b) {
  Mat4f r{};
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float s = 0.f;
      for (int k = 0; k < 4; ++k) s += a.m[i][k] * b.m[k][j];
      r.m[i][j] = s;
    }
  return r;
}
// [R t; 0 1]^-1 = [R^T  -R^T t; 0 1].  (The reference calls Matrix4f::inverse(), a general
// inverse; every matrix it is applied to on this path is rigid.)
inline Mat4f rigid_inverse(const Mat4f& a) {
  Mat4f r = Mat4f::identity();
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i];
  for (int i = 0; i < 3; ++i)
    r.m[i][3] = -(r.m[i][0] * a.m[0][3] + r.m[i][1] * a.m[1][3] + r.m[i][2] * a.m[2][3]);
  return r;
}
// Eigen::Quaternion<T>(w, x, y, z).toRotationMatrix()
template <typename T>
inline void quat_to_matrix(T qw, T qx, T qy, T qz, T R[3][3]) {
  const T tx = T(2) * qx, ty = T(2) * qy, tz = T(2) * qz;
  const T twx = tx * qw, twy = ty * qw, twz = tz * qw;
  const T txx = tx * qx, txy = ty * qx, txz = tz * qx;
  const T tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  R[0][0] = T(1) - (tyy + tzz); R[0][1] = txy - twz;          R[0][2] = txz + twy;
  R[1][0] = txy + twz;          R[1][1] = T(1) - (txx + tzz); R[1][2] = tyz - twx;
  R[2][0] = txz - twy;          R[2][1] = tyz + twx;          R[2][2] = T(1) - (txx + tyy);
}
inline Mat4f pose_from(float qw, float qx, float qy, float qz, float x, float y, float z) {
  Mat4f p = Mat4f::identity();
  float R[3][3];
  quat_to_matrix<float>(qw, qx, qy, qz, R);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) p.m[i][j] = R[i][j];
  p.m[0][3] = x; p.m[1][3] = y; p.m[2][3] = z;
  return p;
}

// cartographer::transform::RollPitchYaw (3d/rigid_transform.cpp:29-36): the quaternion of
// AngleAxis(yaw, Z) * AngleAxis(pitch, Y) * AngleAxis(roll, X), in double.  q = (w, x, y, z).
inline void RollPitchYaw(double roll, double pitch, double yaw, double q[4]) {
  const double cr = std::cos(roll * 0.5), sr = std::sin(roll * 0.5);
  const double cp = std::cos(pitch * 0.5), sp = std::sin(pitch * 0.5);
  const double cy = std::cos(yaw * 0.5), sy = std::sin(yaw * 0.5);
  // (cy, 0, 0, sy) * (cp, 0, sp, 0) = (cy cp, -sy sp, cy sp, sy cp)
  const double aw = cy * cp, ax = -sy * sp, ay = cy * sp, az = sy * cp;
  // ... * (cr, sr, 0, 0)
  q[0] = aw * cr - ax * sr;
  q[1] = aw * sr + ax * cr;
  q[2] = ay * cr + az * sr;
  q[3] = az * cr - ay * sr;
}

// Eigen's MatrixBase::eulerAngles(2, 1, 0) on a Matrix3f: angles (a, b, c) with
// R = Rz(a) Ry(b) Rx(c) and -- Eigen's convention, which the reference inherits -- the FIRST
// angle in [0, pi] (the other two in [-pi, pi]): a rotation whose yaw is negative comes back
// as (yaw + pi, pi - pitch, roll +- pi).  float arithmetic like Eigen's Scalar = float.
inline void eulerAngles210(const Mat3f& R, float res[3]) {
  // a0 = 2, a1 = 1, a2 = 0:  odd = 1, i = 2, j = 1, k = 0
  const int i = 2, j = 1, k = 0;
  const float pi = 3.14159265358979323846f;
  res[0] = std::atan2(R.m[j][k], R.m[k][k]);
  const float c2 = std::sqrt(R.m[i][i] * R.m[i][i] + R.m[i][j] * R.m[i][j]);
  if (res[0] < 0.f) {            // odd && res[0] < 0
    res[0] += pi;
    res[1] = std::atan2(-R.m[i][k], -c2);
  } else {
    res[1] = std::atan2(-R.m[i][k], c2);
  }
  const float s1 = std::sin(res[0]), c1 = std::cos(res[0]);
  res[2] = std::atan2(s1 * R.m[k][i] - c1 * R.m[j][i], c1 * R.m[j][j] - s1 * R.m[k][j]);
  // (Eigen negates the result for even permutations only; (2,1,0) is odd.)
}

// ------------------------------------------------------------------ ground estimation
struct GroundPoint {
  float x, y, z;
};

class GroundEstimator {
 public:
  // ground_estimator.cpp:194-227.  `points`: n points, `stride` floats apart, x y z first (the
  // scan layout of read_lidar_data).  `cloud_out` receives the transformed scan in the same
  // layout (extra channels copied).  Returns T_l2g (identity when no ground was found, with
  // cloud_out left EMPTY like the reference's untouched output cloud).
  Mat4f EsitmateGroundAndTransform(const float* points, size_t n, int stride, std::vector<float>* cloud_out) {
    std::vector<GroundPoint> near;
    for (size_t i = 0; i < n; ++i) {
      const float* p = points + i * (size_t)stride;
      if (p[0] * p[0] + p[1] * p[1] + p[2] * p[2] < 400.) near.push_back({p[0], p[1], p[2]});
    }
    std::vector<GroundPoint> ground;
    if (!FilterGroundByNormals(near, &ground)) {
      if (cloud_out) cloud_out->clear();
      return Mat4f::identity();
    }
    float coeff[4];
    EstimateGround(ground, coeff);
    return TransformPointsToGround(coeff, points, n, stride, cloud_out);
  }

  // ground_estimator.cpp:63-165: normals from the 10 nearest neighbours, inclination
  // theta = atan2(nz, |n_xy|) + 90 deg in 10-degree bins, ground = the fullest bin outside
  // 50..130 deg; returns the points of that bin.  false: "No valid ground found!".
  // (Where the reference would index bin 18 -- theta exactly 180 deg, its assert is compiled
  // out -- the point is counted in bin 17.  Equal bin counts: the lower bin wins; the
  // reference's std::sort leaves that order unspecified.)
  bool FilterGroundByNormals(const std::vector<GroundPoint>& cloud, std::vector<GroundPoint>* ground_points) const {
    ground_points->clear();
    const size_t pt_num = cloud.size();
    if (pt_num < (size_t)kSearchK) return false;
    std::vector<float> normals;
    EstimateNormals(cloud, &normals);
    std::vector<int> bin_flags(pt_num, -1);
    int degree_bins[18] = {0};
    const float rad2deg = 180. / M_PI;
    for (size_t i = 0; i < pt_num; ++i) {
      const float nx = normals[3 * i], ny = normals[3 * i + 1], nz = normals[3 * i + 2];
      if (!(nx == nx)) continue;   // degenerate neighbourhood (PCL: NaN normal)
      const float xy = std::sqrt(nx * nx + ny * ny);
      const float theta = (std::atan2(nz, xy) + M_PI_2) * rad2deg;
      int idx = int(std::floor(theta / 10));
      idx = std::min(17, std::max(0, idx));
      bin_flags[i] = idx;
      degree_bins[idx] += 1;
    }
    int order[18];
    for (int i = 0; i < 18; ++i) order[i] = i;
    std::stable_sort(order, order + 18, [&](int a, int b) { return degree_bins[a] > degree_bins[b]; });
    int ground_bin = -1;
    for (int idx : order) {
      if (idx > 4 && idx < 13) continue;
      ground_bin = idx;
      break;
    }
    if (ground_bin == -1 || degree_bins[ground_bin] == 0) return false;
    for (size_t i = 0; i < pt_num; ++i)
      if (bin_flags[i] == ground_bin) ground_points->push_back(cloud[i]);
    return true;
  }

  // ground_estimator.cpp:19-61: RANSAC plane a x + b y + c z + d = 0, inlier distance 0.1 m,
  // pcl::RandomSampleConsensus::computeModel's loop (probability 0.99, adaptive iteration count
  // k = log(1 - p) / log(1 - w^3), the best 3-point model is returned unrefined).
  void EstimateGround(const std::vector<GroundPoint>& pts, float coeff[4]) const {
    coeff[0] = coeff[1] = 0.f;
    coeff[2] = 1.f;
    coeff[3] = 0.f;
    const size_t n = pts.size();
    if (n < 3) return;
    std::mt19937 rng(12345u);
    const double threshold = 0.1, log_probability = std::log(1.0 - 0.99), one_over_n = 1.0 / (double)n;
    const int max_iterations = 10000;
    double k = 1.0;
    long best = -1;
    int iterations = 0;
    size_t skipped = 0;
    const size_t max_skip = (size_t)max_iterations * 10;
    while (iterations < k && skipped < max_skip) {
      size_t s[3];
      s[0] = rng() % n;
      do s[1] = rng() % n; while (s[1] == s[0]);
      do s[2] = rng() % n; while (s[2] == s[0] || s[2] == s[1]);
      float c[4];
      if (!PlaneFromSamples(pts[s[0]], pts[s[1]], pts[s[2]], c)) {
        ++skipped;
        continue;
      }
      long count = 0;
      for (const GroundPoint& p : pts)
        if (std::fabs(c[0] * p.x + c[1] * p.y + c[2] * p.z + c[3]) < threshold) ++count;
      if (count > best) {
        best = count;
        for (int i = 0; i < 4; ++i) coeff[i] = c[i];
        const double w = (double)best * one_over_n;
        double p_no_outliers = 1.0 - w * w * w;
        p_no_outliers = std::max(std::numeric_limits<double>::epsilon(), p_no_outliers);
        p_no_outliers = std::min(1.0 - std::numeric_limits<double>::epsilon(), p_no_outliers);
        k = log_probability / std::log(p_no_outliers);
      }
      ++iterations;
      if (iterations > max_iterations) break;
    }
  }

  // ground_estimator.cpp:167-192.  T_l2g = [R(roll, pitch, 0) | (0, 0, d)] where (yaw, pitch,
  // roll) are Eigen's Euler angles of FromTwoVectors(ground normal, z) and d the sensor height.
  // With Eigen's [0, pi] range for the first angle, a rotation with a (slightly) negative yaw
  // component comes back in the flipped representation, and "roll, pitch with yaw dropped" then
  // contains a half turn about z -- the reference does exactly that, the search over +-180 deg
  // and the pose composition below absorb it, and so does the 180-degree clause of its error
  // metric.
  Mat4f TransformPointsToGround(const float coeff[4], const float* points, size_t n, int stride,
                                std::vector<float>* cloud_out) const {
    float gx = coeff[0], gy = coeff[1], gz = coeff[2];
    const float norm = std::sqrt(gx * gx + gy * gy + gz * gz);
    const float d = std::fabs(coeff[3]) / norm;
    if (coeff[2] < 0) {   // the ground is below the lidar: normal upwards
      gx *= -1.f; gy *= -1.f; gz *= -1.f;
    }
    gx /= norm; gy /= norm; gz /= norm;
    // Quaternionf::FromTwoVectors(gn_l, z_l), c = gn_l . z_l = gz >= 0: regular branch
    const float c = gz;
    const float ax = gy * 1.f - gz * 0.f, ay = gz * 0.f - gx * 1.f, az = 0.f;   // gn_l x z_l
    const float s = std::sqrt((1.f + c) * 2.f), invs = 1.f / s;
    float qw = s * 0.5f, qx = ax * invs, qy = ay * invs, qz = az * invs;
    const float qn = std::sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
    qw /= qn; qx /= qn; qy /= qn; qz /= qn;
    Mat3f R{};
    quat_to_matrix<float>(qw, qx, qy, qz, R.m);
    float ypr[3];
    eulerAngles210(R, ypr);
    double qd[4];
    RollPitchYaw(ypr[2], ypr[1], 0, qd);
    float Rn[3][3];
    quat_to_matrix<float>((float)qd[0], (float)qd[1], (float)qd[2], (float)qd[3], Rn);
    Mat4f T = Mat4f::identity();
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) T.m[i][j] = Rn[i][j];
    T.m[2][3] = d;
    if (cloud_out) {      // pcl::transformPointCloud: xyz transformed, other channels kept
      cloud_out->resize(n * (size_t)stride);
      for (size_t i = 0; i < n; ++i) {
        const float* p = points + i * (size_t)stride;
        float* o = cloud_out->data() + i * (size_t)stride;
        for (int r = 0; r < 3; ++r)
          o[r] = T.m[r][0] * p[0] + T.m[r][1] * p[1] + T.m[r][2] * p[2] + T.m[r][3];
        for (int ch = 3; ch < stride; ++ch) o[ch] = p[ch];
      }
    }
    return T;
  }

  // pcl::NormalEstimation with setKSearch(10) and the default viewpoint (0, 0, 0): per point
  // the covariance of its 10 nearest neighbours (the point itself included), eigenvector of the
  // smallest eigenvalue, flipped so that it points towards the sensor.  normals: 3 floats per
  // point, NaN when the neighbourhood is degenerate.
  void EstimateNormals(const std::vector<GroundPoint>& cloud, std::vector<float>* normals) const {
    const size_t n = cloud.size();
    normals->assign(3 * n, std::nanf(""));
    if (n < (size_t)kSearchK) return;
    // uniform hash grid over the <= 20 m ball
    const float cell = 0.25f;
    std::vector<uint64_t> key(n);
    std::vector<uint32_t> order(n);
    auto cell_of = [&](float v) { return (int64_t)std::floor(v / cell) + (1 << 20); };
    auto pack = [](int64_t cx, int64_t cy, int64_t cz) { return (uint64_t)cx << 42 | (uint64_t)cy << 21 | (uint64_t)cz; };
    for (size_t i = 0; i < n; ++i) {
      key[i] = pack(cell_of(cloud[i].x), cell_of(cloud[i].y), cell_of(cloud[i].z));
      order[i] = (uint32_t)i;
    }
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return key[a] != key[b] ? key[a] < key[b] : a < b; });
    std::vector<uint64_t> skey(n);
    for (size_t i = 0; i < n; ++i) skey[i] = key[order[i]];
    std::vector<std::pair<float, uint32_t>> cand;
    for (size_t i = 0; i < n; ++i) {
      const GroundPoint& p = cloud[i];
      const int64_t cx = cell_of(p.x), cy = cell_of(p.y), cz = cell_of(p.z);
      // every point outside the cube of cells within Chebyshev distance R is >= R * cell away
      for (int R = 1;; ++R) {
        cand.clear();
        for (int64_t x = cx - R; x <= cx + R; ++x)
          for (int64_t y = cy - R; y <= cy + R; ++y) {
            const uint64_t lo = pack(x, y, cz - R), hi = pack(x, y, cz + R);
            size_t a = std::lower_bound(skey.begin(), skey.end(), lo) - skey.begin();
            for (; a < n && skey[a] <= hi; ++a) {
              const GroundPoint& o = cloud[order[a]];
              const float dx = o.x - p.x, dy = o.y - p.y, dz = o.z - p.z;
              cand.emplace_back(dx * dx + dy * dy + dz * dz, order[a]);
            }
          }
        if (cand.size() >= (size_t)kSearchK) {
          std::partial_sort(cand.begin(), cand.begin() + kSearchK, cand.end());
          const float reach = R * cell;
          if (cand[kSearchK - 1].first <= reach * reach) break;
        }
        if (R > 200) break;   // 50 m: the whole ball has been searched
      }
      if (cand.size() < (size_t)kSearchK) continue;
      if (cand.size() > (size_t)kSearchK) std::partial_sort(cand.begin(), cand.begin() + kSearchK, cand.end());
      double mean[3] = {0, 0, 0};
      for (int j = 0; j < kSearchK; ++j) {
        const GroundPoint& o = cloud[cand[j].second];
        mean[0] += o.x; mean[1] += o.y; mean[2] += o.z;
      }
      for (double& v : mean) v /= kSearchK;
      double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      for (int j = 0; j < kSearchK; ++j) {
        const GroundPoint& o = cloud[cand[j].second];
        const double d[3] = {o.x - mean[0], o.y - mean[1], o.z - mean[2]};
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b) C[a][b] += d[a] * d[b];
      }
      double nvec[3];
      if (!SmallestEigenvector(C, nvec)) continue;
      if (-(p.x * nvec[0] + p.y * nvec[1] + p.z * nvec[2]) < 0) {   // flipNormalTowardsViewpoint, vp = 0
        nvec[0] = -nvec[0]; nvec[1] = -nvec[1]; nvec[2] = -nvec[2];
      }
      (*normals)[3 * i] = (float)nvec[0];
      (*normals)[3 * i + 1] = (float)nvec[1];
      (*normals)[3 * i + 2] = (float)nvec[2];
    }
  }

 private:
  static constexpr int kSearchK = 10;

  // pcl::SampleConsensusModelPlane::computeModelCoefficients
  static bool PlaneFromSamples(const GroundPoint& p0, const GroundPoint& p1, const GroundPoint& p2, float c[4]) {
    const float ax = p1.x - p0.x, ay = p1.y - p0.y, az = p1.z - p0.z;
    const float bx = p2.x - p0.x, by = p2.y - p0.y, bz = p2.z - p0.z;
    float nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
    const float nn = std::sqrt(nx * nx + ny * ny + nz * nz);
    if (!(nn > 1e-12f)) return false;   // collinear sample
    nx /= nn; ny /= nn; nz /= nn;
    c[0] = nx; c[1] = ny; c[2] = nz;
    c[3] = -(nx * p0.x + ny * p0.y + nz * p0.z);
    return true;
  }

  // cyclic Jacobi on a symmetric 3x3; false when the matrix is (numerically) zero
  static bool SmallestEigenvector(const double Cin[3][3], double v[3]) {
    double A[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    double scale = 0;
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        A[a][b] = Cin[a][b];
        scale = std::max(scale, std::fabs(A[a][b]));
      }
    if (!(scale > 0)) return false;
    for (int sweep = 0; sweep < 32; ++sweep) {
      const double off = std::fabs(A[0][1]) + std::fabs(A[0][2]) + std::fabs(A[1][2]);
      if (off <= 1e-15 * scale) break;
      for (int p = 0; p < 2; ++p)
        for (int q = p + 1; q < 3; ++q) {
          if (std::fabs(A[p][q]) <= 1e-300) continue;
          const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
          const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
          const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
          for (int k = 0; k < 3; ++k) {   // A <- A J
            const double akp = A[k][p], akq = A[k][q];
            A[k][p] = c * akp - s * akq;
            A[k][q] = s * akp + c * akq;
          }
          for (int k = 0; k < 3; ++k) {   // A <- J^T A
            const double apk = A[p][k], aqk = A[q][k];
            A[p][k] = c * apk - s * aqk;
            A[q][k] = s * apk + c * aqk;
          }
          for (int k = 0; k < 3; ++k) {
            const double vkp = V[k][p], vkq = V[k][q];
            V[k][p] = c * vkp - s * vkq;
            V[k][q] = s * vkp + c * vkq;
          }
        }
    }
    int m = 0;
    if (A[1][1] < A[m][m]) m = 1;
    if (A[2][2] < A[m][m]) m = 2;
    const double nn = std::sqrt(V[0][m] * V[0][m] + V[1][m] * V[1][m] + V[2][m] * V[2][m]);
    if (!(nn > 0)) return false;
    for (int k = 0; k < 3; ++k) v[k] = V[k][m] / nn;
    return true;
  }
};

// ------------------------------------------------------------------ pose composition
// GlocEvaluator::global_registraion, global_localization.cpp:524-569: the 2-D match result
// xy_yaw = (x, y, yaw) of the query in the candidate's frame -> 4x4 pose of the query in the
// candidate's frame.  Without ground alignment: RollPitchYaw(0, 0, yaw), (x, y, 0).  With:
// roll, pitch, dz from T_db^-1 T_q; x, y, yaw from T_db^-1 * Embed3D(xy_yaw) * T_q, every
// Euler triple by Eigen's eulerAngles(2, 1, 0) (first angle in [0, pi]).
inline Mat4f ComposeLocatedPose(bool align_ground, const float xy_yaw[3], const Mat4f& Tq_l2g, const Mat4f& Tdb_l2g) {
  float e_roll, e_pitch, e_yaw, e_dx, e_dy, e_dz;
  if (align_ground) {
    const Mat4f Tdb_inv = rigid_inverse(Tdb_l2g);
    const Mat4f T_q2db_rpz = mul(Tdb_inv, Tq_l2g);
    float ypr_rpz[3];
    eulerAngles210(T_q2db_rpz.rotation(), ypr_rpz);
    // Rigid2f(xy, yaw) -> Embed3D -> AngleAxis(yaw, Z) quaternion -> rotation matrix
    Mat4f T_qg_dbg = Mat4f::identity();
    float Rz[3][3];
    quat_to_matrix<float>(std::cos(xy_yaw[2] * 0.5f), 0.f, 0.f, std::sin(xy_yaw[2] * 0.5f), Rz);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) T_qg_dbg.m[i][j] = Rz[i][j];
    T_qg_dbg.m[0][3] = xy_yaw[0];
    T_qg_dbg.m[1][3] = xy_yaw[1];
    const Mat4f T_q2db_yawxy = mul(mul(Tdb_inv, T_qg_dbg), Tq_l2g);
    float ypr_yawxy[3];
    eulerAngles210(T_q2db_yawxy.rotation(), ypr_yawxy);
    e_dx = T_q2db_yawxy.m[0][3];
    e_dy = T_q2db_yawxy.m[1][3];
    e_dz = T_q2db_rpz.m[2][3];
    e_roll = ypr_rpz[2];
    e_pitch = ypr_rpz[1];
    e_yaw = ypr_yawxy[0];
  } else {
    e_roll = 0.;
    e_pitch = 0.;
    e_dz = 0.;
    e_dx = xy_yaw[0];
    e_dy = xy_yaw[1];
    e_yaw = xy_yaw[2];
  }
  double q[4], R[3][3];
  RollPitchYaw(e_roll, e_pitch, e_yaw, q);
  quat_to_matrix<double>(q[0], q[1], q[2], q[3], R);
  Mat4f pose = Mat4f::identity();
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) pose.m[i][j] = (float)R[i][j];
  pose.m[0][3] = e_dx;
  pose.m[1][3] = e_dy;
  pose.m[2][3] = e_dz;
  return pose;
}

// registration_recalls, global_localization.cpp:282-303: rotation error in degrees (a result
// within 5 deg of a half turn counts as its distance from the half turn) and position error.
inline void RegistrationError(const Mat4f& pose_db, const Mat4f& pose_q, const Mat4f& located, float* err_rot_deg,
                              float* err_pos) {
  const Mat4f q2db = mul(rigid_inverse(pose_db), pose_q);
  float trace = 0.f;   // trace(gt_rot^T * R_restored)
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) trace += q2db.m[b][a] * located.m[b][a];
  float offset_trace = 0.5f * (trace - 1.f);
  offset_trace = offset_trace < -0.999999f ? -0.999999f : offset_trace;
  offset_trace = offset_trace > 0.999999f ? 0.999999f : offset_trace;
  float err_rot = std::fabs(std::acos(offset_trace));
  const float ex = q2db.m[0][3] - located.m[0][3], ey = q2db.m[1][3] - located.m[1][3], ez = q2db.m[2][3] - located.m[2][3];
  *err_pos = std::sqrt(ex * ex + ey * ey + ez * ez);
  const float rad2deg = 180. / M_PI;
  err_rot = err_rot * rad2deg;
  if (std::fabs(err_rot - 180.f) < 5.f) err_rot = std::fabs(err_rot - 180.f);
  *err_rot_deg = err_rot;
}

}  // namespace gloc

#endif  // GLOC_GROUND_HPP_
