// Command-line face of gloc3d_b200/host/gloc_ground.hpp for tests/test_ground_host.py: reads
// whitespace-separated numbers from stdin, prints results with 9 significant digits.  CPU only.
#include <cstdio>
#include <cstring>
#include <iostream>
#include <vector>

#include "../../gloc3d_b200/host/gloc_ground.hpp"

using namespace gloc;

static Mat4f read_mat4() {
  Mat4f t{};
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) std::cin >> t.m[i][j];
  return t;
}
static void print_mat4(const Mat4f& t) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) std::printf("%.9g ", t.m[i][j]);
  std::printf("\n");
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const std::string mode = argv[1];
  size_t n = 0;
  std::cin >> n;
  if (mode == "euler") {
    for (size_t c = 0; c < n; ++c) {
      Mat3f R{};
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) std::cin >> R.m[i][j];
      float e[3];
      eulerAngles210(R, e);
      std::printf("%.9g %.9g %.9g\n", e[0], e[1], e[2]);
    }
  } else if (mode == "rpy") {
    for (size_t c = 0; c < n; ++c) {
      double r, p, y, q[4];
      std::cin >> r >> p >> y;
      RollPitchYaw(r, p, y, q);
      std::printf("%.17g %.17g %.17g %.17g\n", q[0], q[1], q[2], q[3]);
    }
  } else if (mode == "compose") {
    for (size_t c = 0; c < n; ++c) {
      double align;
      float xy_yaw[3];
      std::cin >> align >> xy_yaw[0] >> xy_yaw[1] >> xy_yaw[2];
      const Mat4f Tq = read_mat4(), Tdb = read_mat4();
      print_mat4(ComposeLocatedPose(align != 0, xy_yaw, Tq, Tdb));
    }
  } else if (mode == "error") {
    for (size_t c = 0; c < n; ++c) {
      const Mat4f a = read_mat4(), b = read_mat4(), l = read_mat4();
      float er, ep;
      RegistrationError(a, b, l, &er, &ep);
      std::printf("%.9g %.9g\n", er, ep);
    }
  } else if (mode == "transform") {   // n plane coefficient sets, each followed by one point
    GroundEstimator ge;
    for (size_t c = 0; c < n; ++c) {
      float coeff[4], p[4];
      std::cin >> coeff[0] >> coeff[1] >> coeff[2] >> coeff[3] >> p[0] >> p[1] >> p[2] >> p[3];
      std::vector<float> out;
      print_mat4(ge.TransformPointsToGround(coeff, p, 1, 4, &out));
      std::printf("%.9g %.9g %.9g %.9g\n", out[0], out[1], out[2], out[3]);
    }
  } else if (mode == "ground") {      // one scan of n points x y z i
    std::vector<float> pts(n * 4);
    for (float& v : pts) std::cin >> v;
    GroundEstimator ge;
    std::vector<GroundPoint> near, ground;
    for (size_t i = 0; i < n; ++i)
      if (pts[4 * i] * pts[4 * i] + pts[4 * i + 1] * pts[4 * i + 1] + pts[4 * i + 2] * pts[4 * i + 2] < 400.)
        near.push_back({pts[4 * i], pts[4 * i + 1], pts[4 * i + 2]});
    const bool ok = ge.FilterGroundByNormals(near, &ground);
    float coeff[4] = {0, 0, 0, 0};
    if (ok) ge.EstimateGround(ground, coeff);
    std::printf("%d %zu %zu\n", ok ? 1 : 0, near.size(), ground.size());
    std::printf("%.9g %.9g %.9g %.9g\n", coeff[0], coeff[1], coeff[2], coeff[3]);
    std::vector<float> out;
    print_mat4(ge.EsitmateGroundAndTransform(pts.data(), n, 4, &out));
    std::printf("%zu\n", out.size());
    for (size_t i = 0; i < out.size(); ++i) std::printf("%.9g%c", out[i], (i % 4 == 3) ? '\n' : ' ');
  } else if (mode == "normals") {     // n points x y z -> normals
    std::vector<GroundPoint> pts(n);
    for (auto& p : pts) std::cin >> p.x >> p.y >> p.z;
    std::vector<float> nrm;
    GroundEstimator().EstimateNormals(pts, &nrm);
    for (size_t i = 0; i < n; ++i) std::printf("%.9g %.9g %.9g\n", nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]);
  } else {
    return 2;
  }
  return 0;
}
