// SPDX-License-Identifier: MIT
// Host-side mirror of the reference's data contract, without Eigen (absent from this image):
//   * fixed-size algebra types under the names the reference surface uses
//     (Vector4d, Matrix4d, Matrix<double,6,6>, Isometry3d ...; SURVEY.md Appendix B, hot-path subset)
//   * traits::Traits<T> free functions              (/root/reference/include/small_gicp/points/traits.hpp:15-78)
//   * PointCloud                                     (.../points/point_cloud.hpp:15-94)
//   * se3_exp / so3_exp / skew                       (.../util/lie.hpp:13-96)
// Storage is column-major like Eigen's default so `.data()` has the layout the C-ABI documents.
// When this backend is added to the reference tree itself the real Eigen types are used instead
// (INTEGRATION.md); nothing here is needed there.
#pragma once
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

namespace small_gicp_b200 {

// ---------------------------------------------------------------------------------------------
// Mat<R,C>: dense column-major fixed-size matrix of doubles
// ---------------------------------------------------------------------------------------------
template <int R, int C>
struct Mat {
  std::array<double, static_cast<size_t>(R) * C> a{};  // zero-initialised

  static Mat Zero() { return Mat(); }
  static Mat Identity() {
    Mat m;
    for (int i = 0; i < (R < C ? R : C); i++) m(i, i) = 1.0;
    return m;
  }
  double& operator()(int r, int c) { return a[static_cast<size_t>(c) * R + r]; }
  double operator()(int r, int c) const { return a[static_cast<size_t>(c) * R + r]; }
  double& operator[](int i) { return a[i]; }
  double operator[](int i) const { return a[i]; }
  double* data() { return a.data(); }
  const double* data() const { return a.data(); }
  static constexpr int rows() { return R; }
  static constexpr int cols() { return C; }

  Mat& operator+=(const Mat& o) {
    for (size_t i = 0; i < a.size(); i++) a[i] += o.a[i];
    return *this;
  }
  Mat& operator-=(const Mat& o) {
    for (size_t i = 0; i < a.size(); i++) a[i] -= o.a[i];
    return *this;
  }
  Mat& operator*=(double s) {
    for (double& v : a) v *= s;
    return *this;
  }
  Mat operator+(const Mat& o) const { return Mat(*this) += o; }
  Mat operator-(const Mat& o) const { return Mat(*this) -= o; }
  Mat operator-() const { return Mat(*this) *= -1.0; }
  Mat operator*(double s) const { return Mat(*this) *= s; }
  Mat operator/(double s) const { return Mat(*this) *= (1.0 / s); }

  Mat<C, R> transpose() const {
    Mat<C, R> t;
    for (int r = 0; r < R; r++)
      for (int c = 0; c < C; c++) t(c, r) = (*this)(r, c);
    return t;
  }
  double squaredNorm() const {
    double s = 0.0;
    for (double v : a) s += v * v;
    return s;
  }
  double norm() const { return std::sqrt(squaredNorm()); }
  double dot(const Mat& o) const {
    double s = 0.0;
    for (size_t i = 0; i < a.size(); i++) s += a[i] * o.a[i];
    return s;
  }
  /// first / last N entries of a column vector
  template <int N>
  Mat<N, 1> head() const {
    static_assert(C == 1 && N <= R, "head<N> of a vector");
    Mat<N, 1> h;
    for (int i = 0; i < N; i++) h[i] = a[i];
    return h;
  }
  template <int N>
  Mat<N, 1> tail() const {
    static_assert(C == 1 && N <= R, "tail<N> of a vector");
    Mat<N, 1> t;
    for (int i = 0; i < N; i++) t[i] = a[R - N + i];
    return t;
  }
  template <int BR, int BC>
  Mat<BR, BC> block(int r0, int c0) const {
    Mat<BR, BC> b;
    for (int r = 0; r < BR; r++)
      for (int c = 0; c < BC; c++) b(r, c) = (*this)(r0 + r, c0 + c);
    return b;
  }
  template <int BR, int BC>
  void set_block(int r0, int c0, const Mat<BR, BC>& b) {
    for (int r = 0; r < BR; r++)
      for (int c = 0; c < BC; c++) (*this)(r0 + r, c0 + c) = b(r, c);
  }
};

template <int R, int K, int C>
Mat<R, C> operator*(const Mat<R, K>& A, const Mat<K, C>& B) {
  Mat<R, C> out;
  for (int c = 0; c < C; c++)
    for (int k = 0; k < K; k++) {
      const double b = B(k, c);
      for (int r = 0; r < R; r++) out(r, c) += A(r, k) * b;
    }
  return out;
}
template <int R, int C>
Mat<R, C> operator*(double s, const Mat<R, C>& A) {
  return A * s;
}

using Vector3d = Mat<3, 1>;
using Vector4d = Mat<4, 1>;
using Matrix3d = Mat<3, 3>;
using Matrix4d = Mat<4, 4>;
using Vector6d = Mat<6, 1>;
using Matrix6d = Mat<6, 6>;

inline Vector3d vec3(double x, double y, double z) {
  Vector3d v;
  v[0] = x;
  v[1] = y;
  v[2] = z;
  return v;
}
inline Vector4d vec4(double x, double y, double z, double w) {
  Vector4d v;
  v[0] = x;
  v[1] = y;
  v[2] = z;
  v[3] = w;
  return v;
}

/// Rigid transform stored as a homogeneous 4x4 (the subset of Eigen::Isometry3d the reference path uses).
struct Isometry3d {
  Matrix4d m = Matrix4d::Identity();
  static Isometry3d Identity() { return Isometry3d(); }
  const Matrix4d& matrix() const { return m; }
  Matrix4d& matrix() { return m; }
  const double* data() const { return m.data(); }
  Matrix3d linear() const { return m.block<3, 3>(0, 0); }
  Vector3d translation() const { return vec3(m(0, 3), m(1, 3), m(2, 3)); }
  void set_linear(const Matrix3d& R) { m.set_block<3, 3>(0, 0, R); }
  void set_translation(const Vector3d& t) {
    for (int i = 0; i < 3; i++) m(i, 3) = t[i];
  }
  Isometry3d operator*(const Isometry3d& o) const {
    Isometry3d r;
    r.m = m * o.m;
    r.m(3, 0) = r.m(3, 1) = r.m(3, 2) = 0.0;
    r.m(3, 3) = 1.0;
    return r;
  }
  Vector4d operator*(const Vector4d& p) const { return m * p; }
  Isometry3d inverse() const {
    Isometry3d r;
    const Matrix3d Rt = linear().transpose();
    r.set_linear(Rt);
    r.set_translation(-(Rt * translation()));
    return r;
  }
};

// ---------------------------------------------------------------------------------------------
// Lie group helpers (twist order [rx ry rz tx ty tz], rotation first; util/lie.hpp)
// ---------------------------------------------------------------------------------------------
inline Matrix3d skew(const Vector3d& x) {
  Matrix3d s;
  s(0, 1) = -x[2];
  s(0, 2) = x[1];
  s(1, 0) = x[2];
  s(1, 2) = -x[0];
  s(2, 0) = -x[1];
  s(2, 1) = x[0];
  return s;
}

/// exp of a rotation vector as a rotation matrix (via the unit quaternion, like lie.hpp:52-69 + toRotationMatrix)
inline Matrix3d so3_exp(const Vector3d& omega) {
  const double th2 = omega.dot(omega);
  double im, re;
  if (th2 < 1e-10) {
    const double th4 = th2 * th2;
    im = 0.5 - th2 / 48.0 + th4 / 3840.0;
    re = 1.0 - th2 / 8.0 + th4 / 384.0;
  } else {
    const double th = std::sqrt(th2);
    im = std::sin(0.5 * th) / th;
    re = std::cos(0.5 * th);
  }
  const double w = re, x = im * omega[0], y = im * omega[1], z = im * omega[2];
  Matrix3d R;
  R(0, 0) = 1.0 - 2.0 * (y * y + z * z);
  R(0, 1) = 2.0 * (x * y - z * w);
  R(0, 2) = 2.0 * (x * z + y * w);
  R(1, 0) = 2.0 * (x * y + z * w);
  R(1, 1) = 1.0 - 2.0 * (x * x + z * z);
  R(1, 2) = 2.0 * (y * z - x * w);
  R(2, 0) = 2.0 * (x * z - y * w);
  R(2, 1) = 2.0 * (y * z + x * w);
  R(2, 2) = 1.0 - 2.0 * (x * x + y * y);
  return R;
}

inline Isometry3d se3_exp(const Vector6d& a) {
  const Vector3d omega = a.head<3>(), v = a.tail<3>();
  const double th2 = omega.dot(omega), th = std::sqrt(th2);
  Isometry3d T;
  const Matrix3d R = so3_exp(omega);
  T.set_linear(R);
  if (th < 1e-10) {
    T.set_translation(R * v);
  } else {
    const Matrix3d W = skew(omega);
    const Matrix3d V = Matrix3d::Identity() + W * ((1.0 - std::cos(th)) / th2) + (W * W) * ((th - std::sin(th)) / (th2 * th));
    T.set_translation(V * v);
  }
  return T;
}

/// Solve (symmetric positive definite) A x = b for 6 unknowns by LDL^T with diagonal pivoting
/// (what `A.ldlt().solve(b)` does at optimizer.hpp:46,109).
inline Vector6d solve_ldlt(Matrix6d A, const Vector6d& b) {
  constexpr int n = 6;
  int perm[n];
  for (int i = 0; i < n; i++) perm[i] = i;
  for (int k = 0; k < n; k++) {
    int piv = k;
    for (int i = k + 1; i < n; i++)
      if (std::abs(A(i, i)) > std::abs(A(piv, piv))) piv = i;
    if (piv != k) {
      for (int j = 0; j < n; j++) std::swap(A(k, j), A(piv, j));
      for (int i = 0; i < n; i++) std::swap(A(i, k), A(i, piv));
      std::swap(perm[k], perm[piv]);
    }
    for (int j = 0; j < k; j++) A(k, k) -= A(k, j) * A(k, j) * A(j, j);
    for (int i = k + 1; i < n; i++) {
      double s = A(i, k);
      for (int j = 0; j < k; j++) s -= A(i, j) * A(k, j) * A(j, j);
      A(i, k) = A(k, k) != 0.0 ? s / A(k, k) : 0.0;
    }
  }
  double y[n];
  for (int i = 0; i < n; i++) y[i] = b[perm[i]];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < i; j++) y[i] -= A(i, j) * y[j];
  for (int i = 0; i < n; i++) y[i] = std::abs(A(i, i)) > 0.0 ? y[i] / A(i, i) : 0.0;
  for (int i = n - 1; i >= 0; i--)
    for (int j = i + 1; j < n; j++) y[i] -= A(j, i) * y[j];
  Vector6d x;
  for (int i = 0; i < n; i++) x[perm[i]] = y[i];
  return x;
}

// ---------------------------------------------------------------------------------------------
// traits (points/traits.hpp:11-78) -- compile-time polymorphism over cloud types, same names
// ---------------------------------------------------------------------------------------------
namespace traits {

template <typename T>
struct Traits;

template <typename T>
size_t size(const T& points) {
  return Traits<T>::size(points);
}
template <typename T>
bool has_points(const T& points) {
  return Traits<T>::has_points(points);
}
template <typename T>
bool has_normals(const T& points) {
  return Traits<T>::has_normals(points);
}
template <typename T>
bool has_covs(const T& points) {
  return Traits<T>::has_covs(points);
}
template <typename T>
auto point(const T& points, size_t i) {
  return Traits<T>::point(points, i);
}
template <typename T>
auto normal(const T& points, size_t i) {
  return Traits<T>::normal(points, i);
}
template <typename T>
auto cov(const T& points, size_t i) {
  return Traits<T>::cov(points, i);
}
template <typename T>
void resize(T& points, size_t n) {
  Traits<T>::resize(points, n);
}
template <typename T>
void set_point(T& points, size_t i, const Vector4d& pt) {
  Traits<T>::set_point(points, i, pt);
}
template <typename T>
void set_normal(T& points, size_t i, const Vector4d& n) {
  Traits<T>::set_normal(points, i, n);
}
template <typename T>
void set_cov(T& points, size_t i, const Matrix4d& cov) {
  Traits<T>::set_cov(points, i, cov);
}

}  // namespace traits

/// Three parallel arrays: (x,y,z,1), (nx,ny,nz,0), 4x4 zero-padded covariance (point_cloud.hpp:69-71).
struct PointCloud {
  using Ptr = std::shared_ptr<PointCloud>;
  using ConstPtr = std::shared_ptr<const PointCloud>;

  PointCloud() = default;
  /// from N x (3|4) doubles or floats, row per point
  template <typename Scalar>
  PointCloud(const Scalar* xyz, size_t n, int stride) {
    resize(n);
    for (size_t i = 0; i < n; i++) points[i] = vec4(xyz[i * stride + 0], xyz[i * stride + 1], xyz[i * stride + 2], 1.0);
  }

  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void resize(size_t n) {
    points.resize(n);
    normals.resize(n);
    covs.resize(n);
  }
  Vector4d& point(size_t i) { return points[i]; }
  Vector4d& normal(size_t i) { return normals[i]; }
  Matrix4d& cov(size_t i) { return covs[i]; }
  const Vector4d& point(size_t i) const { return points[i]; }
  const Vector4d& normal(size_t i) const { return normals[i]; }
  const Matrix4d& cov(size_t i) const { return covs[i]; }

  std::vector<Vector4d> points;
  std::vector<Vector4d> normals;
  std::vector<Matrix4d> covs;
};

namespace traits {
template <>
struct Traits<PointCloud> {
  using Points = PointCloud;
  static size_t size(const Points& p) { return p.size(); }
  static bool has_points(const Points& p) { return !p.points.empty(); }
  static bool has_normals(const Points& p) { return !p.normals.empty(); }
  static bool has_covs(const Points& p) { return !p.covs.empty(); }
  static const Vector4d& point(const Points& p, size_t i) { return p.point(i); }
  static const Vector4d& normal(const Points& p, size_t i) { return p.normal(i); }
  static const Matrix4d& cov(const Points& p, size_t i) { return p.cov(i); }
  static void resize(Points& p, size_t n) { p.resize(n); }
  static void set_point(Points& p, size_t i, const Vector4d& v) { p.point(i) = v; }
  static void set_normal(Points& p, size_t i, const Vector4d& v) { p.normal(i) = v; }
  static void set_cov(Points& p, size_t i, const Matrix4d& c) { p.cov(i) = c; }
};
}  // namespace traits

static_assert(sizeof(Vector4d) == 32 && sizeof(Matrix4d) == 128, "PointCloud arrays must have the reference's memory layout");

}  // namespace small_gicp_b200
