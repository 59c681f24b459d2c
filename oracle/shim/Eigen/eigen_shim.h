// eigen_shim.h -- TEST INFRASTRUCTURE ONLY (oracle/).  The handful of Eigen types the reference's
// registration/2d/*.cpp, 3d/point_cloud.cpp, 3d/probability_values.cpp and (for the BEV projection)
// 3d/submap_3d.cpp, 3d/range_data_inserter_3d.cpp, 3d/range_data.cpp, 3d/hybrid_grid.h use, so that those
// files compile UNMODIFIED into oracle/_ref/libcsm_ref.so and oracle/_ref/libbev_ref.so in an image without Eigen (the build uses
// "Eigen/Core" / "Eigen/Geometry" from PCL's dependency; version unpinned).
//
// What is a restatement here and what is not: the control flow of the matcher (precomputation
// grids, sliding-window maxima, ShrinkToFit, candidate generation, ScoreCandidates,
// BranchAndBound, std::sort) is the reference's own code.  The ARITHMETIC of these classes is
// this file's restatement of Eigen 3.3/3.4, kept in Eigen's evaluation order where floating point
// is involved:
//   Quaternion(AngleAxis)      w = cos(angle/2), vec = sin(angle/2) * axis      (Geometry/Quaternion.h)
//   Quaternion * Vector3       uv = vec x v; uv += uv; v + w*uv + vec x uv      (_transformVector)
//   Quaternion * Quaternion    the Hamilton product, term order as quat_product<>
//   Transform(Translation) * v linear * v + translation with linear = identity  (Geometry/Transform.h)
//   Rotation2D * Rotation2D    angles add; Rotation2D * v = [c -s; s c] v
// Everything else is integer or trivially exact.
#ifndef GLOC_ORACLE_EIGEN_SHIM_H_
#define GLOC_ORACLE_EIGEN_SHIM_H_

#include <algorithm>
#include <array>
#include <climits>
#include <cmath>
#include <cstddef>
#include <limits>
#include <ostream>
#include <string>
#include <type_traits>
#include <vector>

namespace Eigen {

template <typename T, int N> struct Array;

template <int N>
struct BoolArray {
  bool v[N];
  bool all() const {
    for (int i = 0; i < N; ++i)
      if (!v[i]) return false;
    return true;
  }
  bool any() const {
    for (int i = 0; i < N; ++i)
      if (v[i]) return true;
    return false;
  }
};
typedef BoolArray<2> BoolArray2;

template <typename T, int N>
struct CommaInit {
  T* p;
  int i;
  CommaInit& operator,(T x) {
    p[i++] = x;
    return *this;
  }
};

// a writable view of the first K coefficients (v.head<K>() = ...)
template <typename T, int K>
struct HeadRef;

template <typename T, int R, int C = 1>
struct Matrix {
  static_assert(C == 1, "the shim only has column vectors");
  T c[R];
  Matrix() {
    for (int i = 0; i < R; ++i) c[i] = T(0);
  }
  Matrix(T x, T y) : c{x, y} { static_assert(R == 2, "size"); }
  Matrix(T x, T y, T z) : c{x, y, z} { static_assert(R == 3, "size"); }
  Matrix(T x, T y, T z, T w) : c{x, y, z, w} { static_assert(R == 4, "size"); }
  template <int K>
  Matrix(const HeadRef<T, K>& h);
  static Matrix Zero() { return Matrix(); }
  static Matrix unit(int k) {
    Matrix m;
    m.c[k] = T(1);
    return m;
  }
  static Matrix UnitX() { return unit(0); }
  static Matrix UnitY() { return unit(1); }
  static Matrix UnitZ() { return unit(2); }
  T& x() { return c[0]; }
  T& y() { return c[1]; }
  T& z() { return c[2]; }
  T& w() { return c[3]; }
  const T& x() const { return c[0]; }
  const T& y() const { return c[1]; }
  const T& z() const { return c[2]; }
  const T& w() const { return c[3]; }
  T& operator[](int i) { return c[i]; }
  const T& operator[](int i) const { return c[i]; }
  T& operator()(int i) { return c[i]; }
  const T& operator()(int i) const { return c[i]; }
  const T* data() const { return c; }
  T* data() { return c; }
  template <int K>
  Matrix<T, K, 1> head() const {
    Matrix<T, K, 1> r;
    for (int i = 0; i < K; ++i) r.c[i] = c[i];
    return r;
  }
  template <int K>
  HeadRef<T, K> head() {
    return HeadRef<T, K>{c};
  }
  T squaredNorm() const {   // Eigen: sum of abs2, in coefficient order
    T s = c[0] * c[0];
    for (int i = 1; i < R; ++i) s += c[i] * c[i];
    return s;
  }
  T norm() const { return std::sqrt(squaredNorm()); }
  template <typename U>
  Matrix<U, R, 1> cast() const {
    Matrix<U, R, 1> r;
    for (int i = 0; i < R; ++i) r.c[i] = static_cast<U>(c[i]);
    return r;
  }
  Matrix cross(const Matrix& b) const {
    static_assert(R == 3, "cross");
    return Matrix(c[1] * b.c[2] - c[2] * b.c[1], c[2] * b.c[0] - c[0] * b.c[2], c[0] * b.c[1] - c[1] * b.c[0]);
  }
  Matrix operator-() const {
    Matrix r;
    for (int i = 0; i < R; ++i) r.c[i] = -c[i];
    return r;
  }
  Matrix& operator+=(const Matrix& o) {
    for (int i = 0; i < R; ++i) c[i] += o.c[i];
    return *this;
  }
  Matrix& operator-=(const Matrix& o) {
    for (int i = 0; i < R; ++i) c[i] -= o.c[i];
    return *this;
  }
  CommaInit<T, R> operator<<(T x) {
    c[0] = x;
    return CommaInit<T, R>{c, 1};
  }
  Array<T, R> array() const;
  const Matrix& matrix() const { return *this; }
};

template <typename T, int K>
struct HeadRef {
  T* p;
  HeadRef& operator=(const Matrix<T, K, 1>& m) {
    for (int i = 0; i < K; ++i) p[i] = m.c[i];
    return *this;
  }
  T norm() const { return Matrix<T, K, 1>(*this).norm(); }
  T& x() { return p[0]; }
  T& y() { return p[1]; }
};

template <typename T, int R, int C>
template <int K>
Matrix<T, R, C>::Matrix(const HeadRef<T, K>& h) {
  static_assert(K == R, "size");
  for (int i = 0; i < R; ++i) c[i] = h.p[i];
}

template <typename T, int R>
Matrix<T, R, 1> operator+(const Matrix<T, R, 1>& a, const Matrix<T, R, 1>& b) {
  Matrix<T, R, 1> r;
  for (int i = 0; i < R; ++i) r.c[i] = a.c[i] + b.c[i];
  return r;
}
template <typename T, int R>
Matrix<T, R, 1> operator-(const Matrix<T, R, 1>& a, const Matrix<T, R, 1>& b) {
  Matrix<T, R, 1> r;
  for (int i = 0; i < R; ++i) r.c[i] = a.c[i] - b.c[i];
  return r;
}
template <typename T, int R, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
Matrix<T, R, 1> operator*(S s, const Matrix<T, R, 1>& a) {
  Matrix<T, R, 1> r;
  for (int i = 0; i < R; ++i) r.c[i] = static_cast<T>(s) * a.c[i];
  return r;
}
template <typename T, int R, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
Matrix<T, R, 1> operator*(const Matrix<T, R, 1>& a, S s) {
  Matrix<T, R, 1> r;
  for (int i = 0; i < R; ++i) r.c[i] = a.c[i] * static_cast<T>(s);
  return r;
}
template <typename T, int R>
std::ostream& operator<<(std::ostream& os, const Matrix<T, R, 1>& m) {
  for (int i = 0; i < R; ++i) os << (i ? " " : "") << m.c[i];
  return os;
}

template <typename T, int N>
struct Array {
  T c[N];
  Array() {
    for (int i = 0; i < N; ++i) c[i] = T(0);
  }
  Array(T x, T y) : c{x, y} { static_assert(N == 2, "size"); }
  Array(const Matrix<T, N, 1>& m) {   // Eigen converts between Matrix and Array of one shape
    for (int i = 0; i < N; ++i) c[i] = m.c[i];
  }
  Array(T x, T y, T z) : c{x, y, z} { static_assert(N == 3, "size"); }
  Array(T x, T y, T z, T w) : c{x, y, z, w} { static_assert(N == 4, "size"); }
  static Array Zero() { return Array(); }
  static Array Constant(T v) {
    Array r;
    for (int i = 0; i < N; ++i) r.c[i] = v;
    return r;
  }
  T& x() { return c[0]; }
  T& y() { return c[1]; }
  T& z() { return c[2]; }
  T& w() { return c[3]; }
  const T& x() const { return c[0]; }
  const T& y() const { return c[1]; }
  const T& z() const { return c[2]; }
  const T& w() const { return c[3]; }
  T& operator[](int i) { return c[i]; }
  const T& operator[](int i) const { return c[i]; }
  T& operator()(int i) { return c[i]; }
  const T& operator()(int i) const { return c[i]; }
  template <int K>
  Array<T, K> head() const {
    Array<T, K> r;
    for (int i = 0; i < K; ++i) r.c[i] = c[i];
    return r;
  }
  Array cwiseMin(const Array& o) const { return min(o); }
  Array cwiseMax(const Array& o) const { return max(o); }
  Array cwiseAbs() const {
    Array r;
    for (int i = 0; i < N; ++i) r.c[i] = c[i] < T(0) ? -c[i] : c[i];
    return r;
  }
  T maxCoeff() const {
    T m = c[0];
    for (int i = 1; i < N; ++i)
      if (m < c[i]) m = c[i];
    return m;
  }
  T minCoeff() const {
    T m = c[0];
    for (int i = 1; i < N; ++i)
      if (c[i] < m) m = c[i];
    return m;
  }
  template <typename U>
  Array<U, N> cast() const {
    Array<U, N> r;
    for (int i = 0; i < N; ++i) r.c[i] = static_cast<U>(c[i]);
    return r;
  }
  Array& operator+=(const Array& o) {
    for (int i = 0; i < N; ++i) c[i] += o.c[i];
    return *this;
  }
  Array operator-() const {
    Array r;
    for (int i = 0; i < N; ++i) r.c[i] = -c[i];
    return r;
  }
  Array min(const Array& o) const {
    Array r;
    for (int i = 0; i < N; ++i) r.c[i] = o.c[i] < c[i] ? o.c[i] : c[i];
    return r;
  }
  Array max(const Array& o) const {
    Array r;
    for (int i = 0; i < N; ++i) r.c[i] = c[i] < o.c[i] ? o.c[i] : c[i];
    return r;
  }
  Matrix<T, N, 1> matrix() const {
    Matrix<T, N, 1> m;
    for (int i = 0; i < N; ++i) m.c[i] = c[i];
    return m;
  }
};
template <typename T, int N>
Array<T, N> operator+(const Array<T, N>& a, const Array<T, N>& b) {
  Array<T, N> r;
  for (int i = 0; i < N; ++i) r.c[i] = a.c[i] + b.c[i];
  return r;
}
template <typename T, int N>
Array<T, N> operator-(const Array<T, N>& a, const Array<T, N>& b) {
  Array<T, N> r;
  for (int i = 0; i < N; ++i) r.c[i] = a.c[i] - b.c[i];
  return r;
}
#define GLOC_SHIM_ARRAY_CMP(OP)                                            \
  template <typename T, int N>                                             \
  BoolArray<N> operator OP(const Array<T, N>& a, const Array<T, N>& b) {   \
    BoolArray<N> r;                                                        \
    for (int i = 0; i < N; ++i) r.v[i] = a.c[i] OP b.c[i];                 \
    return r;                                                              \
  }
GLOC_SHIM_ARRAY_CMP(<=)
GLOC_SHIM_ARRAY_CMP(<)
GLOC_SHIM_ARRAY_CMP(>=)
GLOC_SHIM_ARRAY_CMP(>)
GLOC_SHIM_ARRAY_CMP(==)
#undef GLOC_SHIM_ARRAY_CMP
template <typename T, int N, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
BoolArray<N> operator>=(const Array<T, N>& a, S s) {
  BoolArray<N> r;
  for (int i = 0; i < N; ++i) r.v[i] = a.c[i] >= static_cast<T>(s);
  return r;
}
template <typename T, int N, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
Array<T, N> operator+(const Array<T, N>& a, S s) {
  Array<T, N> r;
  for (int i = 0; i < N; ++i) r.c[i] = a.c[i] + static_cast<T>(s);
  return r;
}
template <typename T, int N, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
Array<T, N> operator-(const Array<T, N>& a, S s) {
  Array<T, N> r;
  for (int i = 0; i < N; ++i) r.c[i] = a.c[i] - static_cast<T>(s);
  return r;
}
// integer arrays: Eigen evaluates coefficient-wise with the scalar's own arithmetic (int * int, int / int)
template <typename T, int N, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
Array<T, N> operator*(const Array<T, N>& a, S s) {
  Array<T, N> r;
  for (int i = 0; i < N; ++i) r.c[i] = a.c[i] * static_cast<T>(s);
  return r;
}
template <typename T, int N, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
Array<T, N> operator/(const Array<T, N>& a, S s) {
  Array<T, N> r;
  for (int i = 0; i < N; ++i) r.c[i] = a.c[i] / static_cast<T>(s);
  return r;
}
template <typename T, int N>
std::ostream& operator<<(std::ostream& os, const Array<T, N>& m) {
  for (int i = 0; i < N; ++i) os << (i ? " " : "") << m.c[i];
  return os;
}
template <typename T, int R, int C>
Array<T, R> Matrix<T, R, C>::array() const {
  Array<T, R> a;
  for (int i = 0; i < R; ++i) a.c[i] = c[i];
  return a;
}

template <typename T, int N>
struct AlignedBox {
  Matrix<T, N, 1> lo, hi;
  bool empty_ = true;
  AlignedBox() {}
  AlignedBox(const Matrix<T, N, 1>& a, const Matrix<T, N, 1>& b) : lo(a), hi(b), empty_(false) {}
  bool isEmpty() const { return empty_; }
  const Matrix<T, N, 1>& min() const { return lo; }
  const Matrix<T, N, 1>& max() const { return hi; }
  Matrix<T, N, 1> sizes() const { return hi - lo; }
  AlignedBox& extend(const Matrix<T, N, 1>& p) {
    if (empty_) {
      lo = hi = p;
      empty_ = false;
    } else {
      for (int i = 0; i < N; ++i) {
        if (p.c[i] < lo.c[i]) lo.c[i] = p.c[i];
        if (p.c[i] > hi.c[i]) hi.c[i] = p.c[i];
      }
    }
    return *this;
  }
  AlignedBox& translate(const Matrix<T, N, 1>& t) {
    lo += t;
    hi += t;
    return *this;
  }
};

template <typename T>
struct Rotation2D {
  T a;
  Rotation2D() : a(T(0)) {}
  Rotation2D(T angle) : a(angle) {}   // implicit, as in Eigen (Rigid2(translation, double))
  static Rotation2D Identity() { return Rotation2D(T(0)); }
  T angle() const { return a; }
  T& angle() { return a; }
  template <typename U>
  Rotation2D<U> cast() const {
    return Rotation2D<U>(static_cast<U>(a));
  }
  Rotation2D inverse() const { return Rotation2D(-a); }
  Rotation2D operator*(const Rotation2D& o) const { return Rotation2D(a + o.a); }
  Matrix<T, 2, 1> operator*(const Matrix<T, 2, 1>& v) const {
    const T s = std::sin(a), c = std::cos(a);     // toRotationMatrix() * v
    return Matrix<T, 2, 1>(c * v.c[0] + (-s) * v.c[1], s * v.c[0] + c * v.c[1]);
  }
};

template <typename T>
struct AngleAxis {
  T angle_;
  Matrix<T, 3, 1> axis_;
  template <typename A>
  AngleAxis(A angle, const Matrix<T, 3, 1>& axis) : angle_(static_cast<T>(angle)), axis_(axis) {}
  T angle() const { return angle_; }
  const Matrix<T, 3, 1>& axis() const { return axis_; }
};

template <typename T>
struct Quaternion {
  T w_, x_, y_, z_;
  Quaternion() : w_(T(1)), x_(T(0)), y_(T(0)), z_(T(0)) {}
  Quaternion(T w, T x, T y, T z) : w_(w), x_(x), y_(y), z_(z) {}
  Quaternion(const AngleAxis<T>& aa) {            // Quaternion.h: operator=(const AngleAxisType&)
    const T ha = T(0.5) * aa.angle();
    w_ = std::cos(ha);
    const Matrix<T, 3, 1> v = std::sin(ha) * aa.axis();
    x_ = v.c[0];
    y_ = v.c[1];
    z_ = v.c[2];
  }
  static Quaternion Identity() { return Quaternion(T(1), T(0), T(0), T(0)); }
  T& w() { return w_; }
  T& x() { return x_; }
  T& y() { return y_; }
  T& z() { return z_; }
  const T& w() const { return w_; }
  const T& x() const { return x_; }
  const T& y() const { return y_; }
  const T& z() const { return z_; }
  Matrix<T, 3, 1> vec() const { return Matrix<T, 3, 1>(x_, y_, z_); }
  Quaternion conjugate() const { return Quaternion(w_, -x_, -y_, -z_); }
  T squaredNorm() const { return x_ * x_ + y_ * y_ + z_ * z_ + w_ * w_; }   // coeffs() order: x y z w
  T norm() const { return std::sqrt(squaredNorm()); }
  Quaternion normalized() const {
    const T n = norm();
    return Quaternion(w_ / n, x_ / n, y_ / n, z_ / n);
  }
  template <typename U>
  Quaternion<U> cast() const {
    return Quaternion<U>(static_cast<U>(w_), static_cast<U>(x_), static_cast<U>(y_), static_cast<U>(z_));
  }
  Quaternion operator*(const Quaternion& b) const {   // quat_product<> (generic path)
    const Quaternion& a = *this;
    return Quaternion(a.w_ * b.w_ - a.x_ * b.x_ - a.y_ * b.y_ - a.z_ * b.z_,
                      a.w_ * b.x_ + a.x_ * b.w_ + a.y_ * b.z_ - a.z_ * b.y_,
                      a.w_ * b.y_ + a.y_ * b.w_ + a.z_ * b.x_ - a.x_ * b.z_,
                      a.w_ * b.z_ + a.z_ * b.w_ + a.x_ * b.y_ - a.y_ * b.x_);
  }
  Matrix<T, 3, 1> operator*(const Matrix<T, 3, 1>& v) const {   // _transformVector
    Matrix<T, 3, 1> uv = vec().cross(v);
    uv += uv;
    return v + w_ * uv + vec().cross(uv);
  }
};

template <typename T, int N>
struct Translation {
  Matrix<T, N, 1> t;
  Translation(T x, T y) : t(x, y) { static_assert(N == 2, "size"); }
  const Matrix<T, N, 1>& vector() const { return t; }
};

enum TransformTraits { Isometry = 1, Affine = 2 };

template <typename T, int N, int Mode>
struct Transform {
  T lin[N][N];
  Matrix<T, N, 1> t;
  Transform(const Translation<T, N>& tr) : t(tr.t) {   // linear part = identity
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j) lin[i][j] = i == j ? T(1) : T(0);
  }
  Matrix<T, N, 1> operator*(const Matrix<T, N, 1>& v) const {   // linear() * v + translation()
    Matrix<T, N, 1> r;
    for (int i = 0; i < N; ++i) {
      T s = lin[i][0] * v.c[0];
      for (int j = 1; j < N; ++j) s += lin[i][j] * v.c[j];
      r.c[i] = s + t.c[i];
    }
    return r;
  }
};

template <typename V>
struct Map {
  V v;
  template <typename P>
  Map(P* p) {
    for (std::size_t i = 0; i < sizeof(v.c) / sizeof(v.c[0]); ++i) v.c[i] = p[i];
  }
  operator V() const { return v; }
};

typedef Matrix<float, 2, 1> Vector2f;
typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<int, 2, 1> Vector2i;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<float, 4, 1> Vector4f;
typedef Array<int, 2> Array2i;
typedef Array<int, 3> Array3i;
typedef Array<float, 3> Array3f;
typedef Array<int, 4> Array4i;
typedef AlignedBox<int, 2> AlignedBox2i;
typedef Rotation2D<double> Rotation2Dd;
typedef Rotation2D<float> Rotation2Df;
typedef AngleAxis<float> AngleAxisf;
typedef AngleAxis<double> AngleAxisd;
typedef Quaternion<float> Quaternionf;
typedef Quaternion<double> Quaterniond;
typedef Translation<float, 2> Translation2f;
typedef Transform<float, 2, Affine> Affine2f;

}  // namespace Eigen
#endif  // GLOC_ORACLE_EIGEN_SHIM_H_
