// opencv2/opencv.hpp -- TEST INFRASTRUCTURE ONLY (oracle/).  fast_correlative_scan_matcher_2d.h has
// one inline debug helper (PrecomputationGrid2D::ToCvImage) that needs cv::Mat with at<uchar>(), and
// 3d/submap_3d.cpp's ProjectToCvMat returns its occupancy image as a CV_8UC1 cv::Mat filled with a Scalar.
#ifndef GLOC_ORACLE_OPENCV_SHIM_H_
#define GLOC_ORACLE_OPENCV_SHIM_H_
#include <vector>
typedef unsigned char uchar;
#define CV_8UC1 0
namespace cv {
struct Scalar {
  double v;
  Scalar(double x = 0) : v(x) {}
};
class Mat {
 public:
  Mat() : rows(0), cols(0) {}
  Mat(int r, int c, int /*type*/) : rows(r), cols(c), data_((size_t)r * c, 0) {}
  Mat(int r, int c, int /*type*/, const Scalar& s) : rows(r), cols(c), data_((size_t)r * c, (uchar)s.v) {}
  template <typename T>
  T& at(int r, int c) { return reinterpret_cast<T&>(data_[(size_t)r * cols + c]); }
  template <typename T>
  const T& at(int r, int c) const { return reinterpret_cast<const T&>(data_[(size_t)r * cols + c]); }
  int rows, cols;
 private:
  std::vector<uchar> data_;
};
}  // namespace cv
#endif
