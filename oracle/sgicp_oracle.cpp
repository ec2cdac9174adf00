// SPDX-License-Identifier: MIT
//
// sgicp_oracle.cpp -- CPU ORACLE for the small_gicp hot path.  TEST INFRASTRUCTURE ONLY.
//
// This file is a from-scratch, double-precision restatement of the reference's algorithm
// (koide3/small_gicp @ aea1313) for the per-iteration correspondence + linearisation +
// reduction path and for the input-preparation steps the reference's tests run before it.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load it; the product (small_gicp_b200/) never does.
//
// It deliberately shares NO code with the product: it has its own tiny fixed-size linear algebra
// (plain arrays, plain loops) instead of Eigen, which is absent from this image.
// Eigen 3.4.0 (pinned by the reference's CMakeLists.txt:50 FetchContent URL) is the one third-party
// dependency whose arithmetic sits on the path; the closed-form pieces used there are restated from
// Eigen's published algorithms:
//   * 3x3 inverse by cofactors / determinant           (Eigen/src/LU/InverseImpl.h, compute_inverse<.,.,3>)
//   * SelfAdjointEigenSolver<Matrix3d>::computeDirect   (Eigen/src/Eigenvalues/SelfAdjointEigenSolver.h,
//                                                        direct_selfadjoint_eigenvalues<.,3,false>)
//   * LDLT with diagonal pivoting + solve              (Eigen/src/Cholesky/LDLT.h)
//   * Quaternion::toRotationMatrix                     (Eigen/src/Geometry/Quaternion.h)
// Parity status: pinned end-to-end against the reference's own golden data (tests/golden/*, generated
// from /root/reference/data by tests/golden/make_fixtures.py) at the reference's own test tolerances
// (src/test/registration_test.cpp:139-151, kdtree_test.cpp:81-105, python_test.py:143-166); the individual
// Eigen ops are NOT pinned at the bit level (no test in the reference pins them either).
//
// Every function cites the reference file:line (relative to /root/reference/) it follows.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <numeric>
#include <unordered_map>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#else
static inline int omp_get_thread_num() { return 0; }
static inline int omp_get_max_threads() { return 1; }
#endif

namespace orc {

// ---------------------------------------------------------------------------------------------
// tiny linear algebra (row-major plain arrays)
// ---------------------------------------------------------------------------------------------
struct V4 {
  double v[4];
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
};
struct M4 {
  double m[4][4];
};
struct M3 {
  double m[3][3];
};
struct V6 {
  double v[6];
};
struct M6 {
  double m[6][6];
};
struct M46 {
  double m[4][6];
};  // Jacobian 4x6 (row 3 is zero)

static inline V4 v4_zero() { return V4{{0, 0, 0, 0}}; }
static inline M4 m4_zero() {
  M4 r;
  std::memset(&r, 0, sizeof(r));
  return r;
}
static inline M6 m6_zero() {
  M6 r;
  std::memset(&r, 0, sizeof(r));
  return r;
}
static inline V6 v6_zero() {
  V6 r;
  std::memset(&r, 0, sizeof(r));
  return r;
}
static inline M4 m4_identity() {
  M4 r = m4_zero();
  for (int i = 0; i < 4; i++) r.m[i][i] = 1.0;
  return r;
}
static inline V4 operator+(const V4& a, const V4& b) { return V4{{a[0] + b[0], a[1] + b[1], a[2] + b[2], a[3] + b[3]}}; }
static inline V4 operator-(const V4& a, const V4& b) { return V4{{a[0] - b[0], a[1] - b[1], a[2] - b[2], a[3] - b[3]}}; }
static inline double sq_norm(const V4& a) { return a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3]; }
static inline V4 mul(const M4& A, const V4& x) {
  V4 r;
  for (int i = 0; i < 4; i++) r.v[i] = A.m[i][0] * x[0] + A.m[i][1] * x[1] + A.m[i][2] * x[2] + A.m[i][3] * x[3];
  return r;
}
static inline M4 mul(const M4& A, const M4& B) {
  M4 r;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) r.m[i][j] = A.m[i][0] * B.m[0][j] + A.m[i][1] * B.m[1][j] + A.m[i][2] * B.m[2][j] + A.m[i][3] * B.m[3][j];
  return r;
}
static inline M4 transpose(const M4& A) {
  M4 r;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) r.m[i][j] = A.m[j][i];
  return r;
}
static inline M4 add(const M4& A, const M4& B) {
  M4 r;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) r.m[i][j] = A.m[i][j] + B.m[i][j];
  return r;
}

/// Isometry3d: a 4x4 homogeneous matrix [R t; 0 1].
struct Iso {
  M4 T;
};
static inline Iso iso_identity() { return Iso{m4_identity()}; }
static inline Iso iso_mul(const Iso& a, const Iso& b) {
  Iso r{mul(a.T, b.T)};
  r.T.m[3][0] = r.T.m[3][1] = r.T.m[3][2] = 0.0;
  r.T.m[3][3] = 1.0;
  return r;
}
static inline Iso iso_inverse(const Iso& a) {  // Eigen Isometry inverse: [R^T, -R^T t]
  Iso r = iso_identity();
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.T.m[i][j] = a.T.m[j][i];
  for (int i = 0; i < 3; i++) r.T.m[i][3] = -(r.T.m[i][0] * a.T.m[0][3] + r.T.m[i][1] * a.T.m[1][3] + r.T.m[i][2] * a.T.m[2][3]);
  return r;
}

/// 3x3 inverse by cofactors (Eigen compute_inverse_size3_helper); used at factors/gicp_factor.hpp:60.
static inline M3 inverse3(const M3& A) {
  const double(*a)[3] = A.m;
  M3 c;  // cofactor matrix, transposed (adjugate)
  c.m[0][0] = a[1][1] * a[2][2] - a[1][2] * a[2][1];
  c.m[1][0] = a[1][2] * a[2][0] - a[1][0] * a[2][2];
  c.m[2][0] = a[1][0] * a[2][1] - a[1][1] * a[2][0];
  c.m[0][1] = a[0][2] * a[2][1] - a[0][1] * a[2][2];
  c.m[1][1] = a[0][0] * a[2][2] - a[0][2] * a[2][0];
  c.m[2][1] = a[0][1] * a[2][0] - a[0][0] * a[2][1];
  c.m[0][2] = a[0][1] * a[1][2] - a[0][2] * a[1][1];
  c.m[1][2] = a[0][2] * a[1][0] - a[0][0] * a[1][2];
  c.m[2][2] = a[0][0] * a[1][1] - a[0][1] * a[1][0];
  const double det = a[0][0] * c.m[0][0] + a[0][1] * c.m[1][0] + a[0][2] * c.m[2][0];
  const double invdet = 1.0 / det;
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.m[i][j] = c.m[i][j] * invdet;
  return r;
}

/// util/lie.hpp:13-24
static inline M3 skew(const double x[3]) {
  M3 s;
  std::memset(&s, 0, sizeof(s));
  s.m[0][1] = -x[2];
  s.m[0][2] = x[1];
  s.m[1][0] = x[2];
  s.m[1][2] = -x[0];
  s.m[2][0] = -x[1];
  s.m[2][1] = x[0];
  return s;
}

/// util/lie.hpp:52-69 (so3_exp -> quaternion w,x,y,z) + Eigen Quaternion::toRotationMatrix
static inline M3 so3_exp_matrix(const double omega[3]) {
  const double theta_sq = omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2];
  double imag_factor, real_factor;
  if (theta_sq < 1e-10) {
    const double theta_quad = theta_sq * theta_sq;
    imag_factor = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * theta_quad;
    real_factor = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * theta_quad;
  } else {
    const double theta = std::sqrt(theta_sq);
    const double half_theta = 0.5 * theta;
    imag_factor = std::sin(half_theta) / theta;
    real_factor = std::cos(half_theta);
  }
  const double w = real_factor, x = imag_factor * omega[0], y = imag_factor * omega[1], z = imag_factor * omega[2];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  M3 R;
  R.m[0][0] = 1 - (tyy + tzz);
  R.m[0][1] = txy - twz;
  R.m[0][2] = txz + twy;
  R.m[1][0] = txy + twz;
  R.m[1][1] = 1 - (txx + tzz);
  R.m[1][2] = tyz - twx;
  R.m[2][0] = txz - twy;
  R.m[2][1] = tyz + twx;
  R.m[2][2] = 1 - (txx + tyy);
  return R;
}

/// util/lie.hpp:73-96 (rotation-first twist [rx ry rz tx ty tz])
static inline Iso se3_exp(const V6& a) {
  const double* omega = a.v;
  const double theta_sq = omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2];
  const double theta = std::sqrt(theta_sq);
  Iso se3 = iso_identity();
  const M3 R = so3_exp_matrix(omega);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) se3.T.m[i][j] = R.m[i][j];
  const double* t = a.v + 3;
  if (theta < 1e-10) {
    for (int i = 0; i < 3; i++) se3.T.m[i][3] = R.m[i][0] * t[0] + R.m[i][1] * t[1] + R.m[i][2] * t[2];
  } else {
    const M3 O = skew(omega);
    M3 OO;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) OO.m[i][j] = O.m[i][0] * O.m[0][j] + O.m[i][1] * O.m[1][j] + O.m[i][2] * O.m[2][j];
    const double c1 = (1.0 - std::cos(theta)) / theta_sq;
    const double c2 = (theta - std::sin(theta)) / (theta_sq * theta);
    for (int i = 0; i < 3; i++) {
      double acc = 0.0;
      for (int j = 0; j < 3; j++) {
        const double V = (i == j ? 1.0 : 0.0) + c1 * O.m[i][j] + c2 * OO.m[i][j];
        acc += V * t[j];
      }
      se3.T.m[i][3] = acc;
    }
  }
  return se3;
}

/// Eigen LDLT (robust Cholesky with diagonal pivoting) + solve, as used at registration/optimizer.hpp:46,109.
static inline V6 ldlt_solve6(const M6& A, const V6& rhs) {
  const int n = 6;
  double a[6][6];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) a[i][j] = A.m[i][j];
  int perm[6];
  for (int i = 0; i < n; i++) perm[i] = i;
  // in-place LDLT on the lower triangle with symmetric pivoting (largest |diagonal|)
  for (int k = 0; k < n; k++) {
    int piv = k;
    double best = std::abs(a[k][k]);
    for (int i = k + 1; i < n; i++)
      if (std::abs(a[i][i]) > best) {
        best = std::abs(a[i][i]);
        piv = i;
      }
    if (piv != k) {
      for (int j = 0; j < n; j++) std::swap(a[k][j], a[piv][j]);
      for (int i = 0; i < n; i++) std::swap(a[i][k], a[i][piv]);
      std::swap(perm[k], perm[piv]);
    }
    // a[k][k] -= sum_{j<k} L[k][j]^2 D[j]; column update
    for (int j = 0; j < k; j++) a[k][k] -= a[k][j] * a[k][j] * a[j][j];
    for (int i = k + 1; i < n; i++) {
      double s = a[i][k];
      for (int j = 0; j < k; j++) s -= a[i][j] * a[k][j] * a[j][j];
      a[i][k] = (a[k][k] != 0.0) ? s / a[k][k] : 0.0;
    }
  }
  double y[6];
  for (int i = 0; i < n; i++) y[i] = rhs.v[perm[i]];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < i; j++) y[i] -= a[i][j] * y[j];
  const double tol = std::numeric_limits<double>::min();
  for (int i = 0; i < n; i++) y[i] = (std::abs(a[i][i]) > tol) ? y[i] / a[i][i] : 0.0;
  for (int i = n - 1; i >= 0; i--)
    for (int j = i + 1; j < n; j++) y[i] -= a[j][i] * y[j];
  V6 x;
  for (int i = 0; i < n; i++) x.v[perm[i]] = y[i];
  return x;
}

/// Eigen SelfAdjointEigenSolver<Matrix3d>::computeDirect (closed form, eigenvalues ascending, eigenvectors in columns).
/// Used at util/normal_estimation.hpp:88-89.
static void eigen_sym3_direct(const M3& mat, double evals[3], M3& evecs) {
  const double eps = std::numeric_limits<double>::epsilon();
  const double shift = (mat.m[0][0] + mat.m[1][1] + mat.m[2][2]) / 3.0;
  M3 sm = mat;
  for (int i = 0; i < 3; i++) sm.m[i][i] -= shift;
  // Eigen scales by the max |coeff| of the (shifted) lower triangle
  double scale = 0.0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j <= i; j++) scale = std::max(scale, std::abs(sm.m[i][j]));
  if (scale > 0.0)
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) sm.m[i][j] /= scale;
  // symmetric view of the lower triangle
  const double m00 = sm.m[0][0], m11 = sm.m[1][1], m22 = sm.m[2][2], m10 = sm.m[1][0], m20 = sm.m[2][0], m21 = sm.m[2][1];
  {  // computeRoots
    const double s_inv3 = 1.0 / 3.0, s_sqrt3 = std::sqrt(3.0);
    const double c0 = m00 * m11 * m22 + 2.0 * m10 * m20 * m21 - m00 * m21 * m21 - m11 * m20 * m20 - m22 * m10 * m10;
    const double c1 = m00 * m11 - m10 * m10 + m00 * m22 - m20 * m20 + m11 * m22 - m21 * m21;
    const double c2 = m00 + m11 + m22;
    const double c2_over_3 = c2 * s_inv3;
    double a_over_3 = (c2 * c2_over_3 - c1) * s_inv3;
    a_over_3 = std::max(a_over_3, 0.0);
    const double half_b = 0.5 * (c0 + c2_over_3 * (2.0 * c2_over_3 * c2_over_3 - c1));
    double q = a_over_3 * a_over_3 * a_over_3 - half_b * half_b;
    q = std::max(q, 0.0);
    const double rho = std::sqrt(a_over_3);
    const double theta = std::atan2(std::sqrt(q), half_b) * s_inv3;
    const double cos_theta = std::cos(theta), sin_theta = std::sin(theta);
    evals[0] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
    evals[1] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
    evals[2] = c2_over_3 + 2.0 * rho * cos_theta;
  }
  auto sym = [&](int i, int j) { return i >= j ? sm.m[i][j] : sm.m[j][i]; };
  auto extract_kernel = [&](double tmp[3][3], double res[3], double representative[3]) {
    int i0 = 0;
    double best = std::abs(tmp[0][0]);
    for (int i = 1; i < 3; i++)
      if (std::abs(tmp[i][i]) > best) {
        best = std::abs(tmp[i][i]);
        i0 = i;
      }
    for (int i = 0; i < 3; i++) representative[i] = tmp[i][i0];
    const int i1 = (i0 + 1) % 3, i2 = (i0 + 2) % 3;
    double c0[3], c1[3];
    auto cross = [](const double a[3], const double b[3], double o[3]) {
      o[0] = a[1] * b[2] - a[2] * b[1];
      o[1] = a[2] * b[0] - a[0] * b[2];
      o[2] = a[0] * b[1] - a[1] * b[0];
    };
    double col1[3] = {tmp[0][i1], tmp[1][i1], tmp[2][i1]}, col2[3] = {tmp[0][i2], tmp[1][i2], tmp[2][i2]};
    cross(representative, col1, c0);
    cross(representative, col2, c1);
    const double n0 = c0[0] * c0[0] + c0[1] * c0[1] + c0[2] * c0[2];
    const double n1 = c1[0] * c1[0] + c1[1] * c1[1] + c1[2] * c1[2];
    if (n0 > n1) {
      const double s = std::sqrt(n0);
      for (int i = 0; i < 3; i++) res[i] = c0[i] / s;
    } else {
      const double s = std::sqrt(n1);
      for (int i = 0; i < 3; i++) res[i] = c1[i] / s;
    }
  };
  if ((evals[2] - evals[0]) <= eps) {
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) evecs.m[i][j] = (i == j) ? 1.0 : 0.0;
  } else {
    double tmp[3][3];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) tmp[i][j] = sym(i, j);
    double d0 = evals[2] - evals[1];
    const double d1 = evals[1] - evals[0];
    int k = 0, l = 2;
    if (d0 > d1) {
      std::swap(k, l);
      d0 = d1;
    }
    double vk[3], vl[3];
    for (int i = 0; i < 3; i++) tmp[i][i] -= evals[k];
    extract_kernel(tmp, vk, vl);  // vl <- representative
    if (d0 <= 2.0 * eps * d1) {
      const double dot = vk[0] * vl[0] + vk[1] * vl[1] + vk[2] * vl[2];
      for (int i = 0; i < 3; i++) vl[i] -= dot * vl[i];
      const double nrm = std::sqrt(vl[0] * vl[0] + vl[1] * vl[1] + vl[2] * vl[2]);
      for (int i = 0; i < 3; i++) vl[i] /= nrm;
    } else {
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) tmp[i][j] = sym(i, j);
      for (int i = 0; i < 3; i++) tmp[i][i] -= evals[l];
      double dummy[3];
      extract_kernel(tmp, vl, dummy);
    }
    for (int i = 0; i < 3; i++) {
      evecs.m[i][k] = vk[i];
      evecs.m[i][l] = vl[i];
    }
    // col(1) = col(2).cross(col(0)).normalized()
    const double a[3] = {evecs.m[0][2], evecs.m[1][2], evecs.m[2][2]}, b[3] = {evecs.m[0][0], evecs.m[1][0], evecs.m[2][0]};
    double c[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    const double nrm = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    for (int i = 0; i < 3; i++) evecs.m[i][1] = c[i] / nrm;
  }
  for (int i = 0; i < 3; i++) evals[i] = evals[i] * scale + shift;
}

// ---------------------------------------------------------------------------------------------
// points/point_cloud.hpp:15-72  (x,y,z,1) / (nx,ny,nz,0) / 4x4 zero-padded covariance
// ---------------------------------------------------------------------------------------------
struct Cloud {
  std::vector<V4> points, normals;
  std::vector<M4> covs;
  size_t size() const { return points.size(); }
  void resize(size_t n) {
    points.resize(n);
    normals.resize(n);
    covs.resize(n);
  }
};

// ---------------------------------------------------------------------------------------------
// util/fast_floor.hpp:12-15, util/downsampling.hpp:22-78 (serial voxelgrid_sampling)
// ---------------------------------------------------------------------------------------------
static inline int fast_floor1(double x) {
  const int n = static_cast<int>(x);
  return n - (x < static_cast<double>(n) ? 1 : 0);
}

static std::unique_ptr<Cloud> voxelgrid_sampling(const Cloud& points, double leaf_size) {
  auto downsampled = std::make_unique<Cloud>();
  if (points.size() == 0) return downsampled;
  const double inv_leaf_size = 1.0 / leaf_size;
  constexpr std::uint64_t invalid_coord = std::numeric_limits<std::uint64_t>::max();
  constexpr int coord_bit_size = 21;
  constexpr size_t coord_bit_mask = (1 << 21) - 1;
  constexpr int coord_offset = 1 << (coord_bit_size - 1);

  std::vector<std::pair<std::uint64_t, size_t>> coord_pt(points.size());
  for (size_t i = 0; i < points.size(); i++) {
    int coord[4];
    bool bad = false;
    for (int d = 0; d < 4; d++) {
      coord[d] = fast_floor1(points.points[i][d] * inv_leaf_size) + coord_offset;
      if (coord[d] < 0 || static_cast<size_t>(coord[d]) > coord_bit_mask) bad = true;
    }
    if (bad) {
      coord_pt[i] = {invalid_coord, i};
      continue;
    }
    const std::uint64_t bits = (static_cast<std::uint64_t>(coord[0] & coord_bit_mask) << (coord_bit_size * 0)) |
                               (static_cast<std::uint64_t>(coord[1] & coord_bit_mask) << (coord_bit_size * 1)) |
                               (static_cast<std::uint64_t>(coord[2] & coord_bit_mask) << (coord_bit_size * 2));
    coord_pt[i] = {bits, i};
  }
  const auto compare = [](const auto& lhs, const auto& rhs) { return lhs.first < rhs.first; };
  std::sort(coord_pt.begin(), coord_pt.end(), compare);

  downsampled->resize(points.size());
  size_t num_points = 0;
  V4 sum_pt = points.points[coord_pt.front().second];
  auto emit = [&](const V4& s) {
    V4 p{{s[0] / s[3], s[1] / s[3], s[2] / s[3], s[3] / s[3]}};
    downsampled->points[num_points++] = p;
  };
  for (size_t i = 1; i < points.size(); i++) {
    if (coord_pt[i].first == invalid_coord) continue;
    if (coord_pt[i - 1].first != coord_pt[i].first) {
      emit(sum_pt);
      sum_pt = v4_zero();
    }
    sum_pt = sum_pt + points.points[coord_pt[i].second];
  }
  emit(sum_pt);
  downsampled->resize(num_points);
  return downsampled;
}

// ---------------------------------------------------------------------------------------------
// ann/knn_result.hpp:13-108
// ---------------------------------------------------------------------------------------------
struct KnnSetting {
  double epsilon = 0.0;
};

/// KnnResult<N> with N==1 fast path or dynamic capacity; index_transform adds `index_base`
/// (voxel_id << 32 for voxel maps, ann/incremental_voxelmap.hpp:102-103,151).
struct KnnResult {
  static constexpr size_t INVALID = std::numeric_limits<size_t>::max();
  int capacity;
  int num_found_neighbors = 0;
  size_t* indices;
  double* distances;
  size_t index_base = 0;
  KnnResult(size_t* idx, double* dist, int k) : capacity(k), indices(idx), distances(dist) {
    std::fill(indices, indices + capacity, INVALID);
    std::fill(distances, distances + capacity, std::numeric_limits<double>::max());
  }
  size_t num_found() const { return num_found_neighbors; }
  double worst_distance() const { return distances[capacity - 1]; }
  void push(size_t index, double distance) {
    if (distance >= worst_distance()) return;
    if (capacity == 1) {
      indices[0] = index_base | index;
      distances[0] = distance;
    } else {
      int insert_loc = std::min<int>(num_found_neighbors, capacity - 1);
      for (; insert_loc > 0 && distance < distances[insert_loc - 1]; insert_loc--) {
        indices[insert_loc] = indices[insert_loc - 1];
        distances[insert_loc] = distances[insert_loc - 1];
      }
      indices[insert_loc] = index_base | index;
      distances[insert_loc] = distance;
    }
    num_found_neighbors = std::min<int>(num_found_neighbors + 1, capacity);
  }
  bool fulfilled(const KnnSetting& s) const { return worst_distance() < s.epsilon; }
};

// ---------------------------------------------------------------------------------------------
// ann/kdtree.hpp:56-241 + ann/projection.hpp:18-55 (AxisAlignedProjection)
// ---------------------------------------------------------------------------------------------
using NodeIndexType = std::uint32_t;
static constexpr NodeIndexType INVALID_NODE = std::numeric_limits<NodeIndexType>::max();

/// Same 24-byte layout as KdTreeNode<AxisAlignedProjection> (ann/kdtree.hpp:56-71).
struct KdTreeNode {
  union {
    struct {
      NodeIndexType first, last;
    } lr;
    struct {
      int axis;
      double thresh;
    } sub;
  } node_type;
  NodeIndexType left = INVALID_NODE;
  NodeIndexType right = INVALID_NODE;
};
static_assert(sizeof(KdTreeNode) == 24, "node layout");

struct KdTree {
  const Cloud* points = nullptr;
  std::vector<size_t> indices;
  NodeIndexType root = 0;
  std::vector<KdTreeNode> nodes;
  int max_leaf_size = 20;
  int max_scan_count = 128;

  /// projection.hpp:27-49
  int find_axis(const size_t* first, const size_t* last) const {
    const size_t N = last - first;
    V4 sum_pt = v4_zero(), sum_sq = v4_zero();
    const size_t step = N < static_cast<size_t>(max_scan_count) ? 1 : N / max_scan_count;
    const size_t num_steps = N / step;
    for (size_t i = 0; i < num_steps; i++) {
      const V4& pt = points->points[*(first + step * i)];
      for (int d = 0; d < 4; d++) {
        sum_pt.v[d] += pt[d];
        sum_sq.v[d] += pt[d] * pt[d];
      }
    }
    double var[4];
    for (int d = 0; d < 4; d++) {
      const double mean = sum_pt[d] / sum_pt[3];
      var[d] = sum_sq[d] - mean * sum_pt[d];
    }
    return var[0] > var[1] ? (var[0] > var[2] ? 0 : 2) : (var[1] > var[2] ? 1 : 2);
  }

  /// kdtree.hpp:99-126
  NodeIndexType create_node(size_t& node_count, size_t* global_first, size_t* first, size_t* last) {
    const size_t N = last - first;
    const NodeIndexType node_index = node_count++;
    if (N <= static_cast<size_t>(max_leaf_size)) {
      nodes[node_index].node_type.lr.first = first - global_first;
      nodes[node_index].node_type.lr.last = last - global_first;
      return node_index;
    }
    const int axis = find_axis(first, last);
    size_t* median_itr = first + N / 2;
    const Cloud& pc = *points;
    std::nth_element(first, median_itr, last, [&](size_t i, size_t j) { return pc.points[i][axis] < pc.points[j][axis]; });
    nodes[node_index].node_type.sub.axis = axis;
    nodes[node_index].node_type.sub.thresh = pc.points[*median_itr][axis];
    const NodeIndexType l = create_node(node_count, global_first, first, median_itr);
    nodes[node_index].left = l;
    const NodeIndexType r = create_node(node_count, global_first, median_itr, last);
    nodes[node_index].right = r;
    return node_index;
  }

  /// kdtree.hpp:82-91, 146-153
  void build(const Cloud& cloud) {
    points = &cloud;
    indices.clear();
    nodes.clear();
    if (cloud.size() == 0) return;
    indices.resize(cloud.size());
    std::iota(indices.begin(), indices.end(), 0);
    size_t node_count = 0;
    nodes.resize(cloud.size());
    root = create_node(node_count, indices.data(), indices.data(), indices.data() + indices.size());
    nodes.resize(node_count);
  }

  /// kdtree.hpp:193-233
  bool knn_search(const V4& query, NodeIndexType node_index, KnnResult& result, const KnnSetting& setting) const {
    const KdTreeNode& node = nodes[node_index];
    if (node.left == INVALID_NODE) {
      for (size_t i = node.node_type.lr.first; i < node.node_type.lr.last; i++) {
        const double sq_dist = sq_norm(points->points[indices[i]] - query);
        result.push(indices[i], sq_dist);
      }
      return !result.fulfilled(setting);
    }
    const double val = query[node.node_type.sub.axis];
    const double diff = val - node.node_type.sub.thresh;
    const double cut_sq_dist = diff * diff;
    NodeIndexType best_child, other_child;
    if (diff < 0.0) {
      best_child = node.left;
      other_child = node.right;
    } else {
      best_child = node.right;
      other_child = node.left;
    }
    if (!knn_search(query, best_child, result, setting)) return false;
    if (result.worst_distance() > cut_sq_dist) return knn_search(query, other_child, result, setting);
    return true;
  }

  /// kdtree.hpp:161-189
  size_t knn_search(const V4& query, int k, size_t* k_indices, double* k_sq_dists) const {
    KnnResult result(k_indices, k_sq_dists, k);
    if (nodes.empty()) return 0;
    knn_search(query, root, result, KnnSetting());
    return result.num_found();
  }
  size_t nearest_neighbor_search(const V4& query, size_t* k_index, double* k_sq_dist) const { return knn_search(query, 1, k_index, k_sq_dist); }
};

// ---------------------------------------------------------------------------------------------
// util/normal_estimation.hpp:12-92 (+ _omp.hpp:9-26)
// ---------------------------------------------------------------------------------------------
enum FeatureMode { FEAT_NORMAL = 1, FEAT_COV = 2, FEAT_NORMAL_COV = 3 };

static void set_normal_invalid(Cloud& c, size_t i) { c.normals[i] = v4_zero(); }
static void set_cov_invalid(Cloud& c, size_t i) {
  M4 cov = m4_identity();
  cov.m[3][3] = 0.0;
  c.covs[i] = cov;
}
static void set_normal(Cloud& c, size_t i, const M3& ev) {  // normal_estimation.hpp:17-24
  double n[3] = {ev.m[0][0], ev.m[1][0], ev.m[2][0]};
  const double nrm = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  V4 normal{{n[0] / nrm, n[1] / nrm, n[2] / nrm, 0.0}};
  const V4& p = c.points[i];
  const double dot = p[0] * normal[0] + p[1] * normal[1] + p[2] * normal[2] + p[3] * normal[3];
  if (dot > 0)
    c.normals[i] = V4{{-normal[0], -normal[1], -normal[2], -normal[3]}};
  else
    c.normals[i] = normal;
}
static void set_cov(Cloud& c, size_t i, const M3& ev) {  // normal_estimation.hpp:40-45
  const double values[3] = {1e-3, 1.0, 1.0};
  M4 cov = m4_zero();
  // (V * diag) * V^T
  double VD[3][3];
  for (int r = 0; r < 3; r++)
    for (int k = 0; k < 3; k++) VD[r][k] = ev.m[r][k] * values[k];
  for (int r = 0; r < 3; r++)
    for (int cc = 0; cc < 3; cc++) cov.m[r][cc] = VD[r][0] * ev.m[cc][0] + VD[r][1] * ev.m[cc][1] + VD[r][2] * ev.m[cc][2];
  c.covs[i] = cov;
}

static void estimate_local_features(Cloud& cloud, const KdTree& tree, int num_neighbors, size_t point_index, int mode) {
  std::vector<size_t> k_indices(num_neighbors);
  std::vector<double> k_sq_dists(num_neighbors);
  const size_t n = tree.knn_search(cloud.points[point_index], num_neighbors, k_indices.data(), k_sq_dists.data());
  if (n < 5) {
    if (mode & FEAT_NORMAL) set_normal_invalid(cloud, point_index);
    if (mode & FEAT_COV) set_cov_invalid(cloud, point_index);
    return;
  }
  V4 sum_points = v4_zero();
  M4 sum_cross = m4_zero();
  for (size_t i = 0; i < n; i++) {
    const V4& pt = cloud.points[k_indices[i]];
    sum_points = sum_points + pt;
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) sum_cross.m[r][c] += pt[r] * pt[c];
  }
  V4 mean;
  for (int d = 0; d < 4; d++) mean.v[d] = sum_points[d] / static_cast<double>(n);
  M3 cov;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) cov.m[r][c] = (sum_cross.m[r][c] - mean[r] * sum_points[c]) / static_cast<double>(n);
  double evals[3];
  M3 evecs;
  eigen_sym3_direct(cov, evals, evecs);
  if (mode & FEAT_NORMAL) set_normal(cloud, point_index, evecs);
  if (mode & FEAT_COV) set_cov(cloud, point_index, evecs);
}

static void estimate_features(Cloud& cloud, const KdTree& tree, int num_neighbors, int mode, int num_threads) {
  cloud.resize(cloud.size());
  const std::int64_t N = cloud.size();
  if (num_threads <= 1) {
    for (std::int64_t i = 0; i < N; i++) estimate_local_features(cloud, tree, num_neighbors, i, mode);
  } else {
#pragma omp parallel for num_threads(num_threads)
    for (std::int64_t i = 0; i < N; i++) estimate_local_features(cloud, tree, num_neighbors, i, mode);
  }
}

// ---------------------------------------------------------------------------------------------
// ann/gaussian_voxelmap.hpp:15-60 + ann/incremental_voxelmap.hpp:46-186 (single insert; no LRU eviction
// can happen within one insert: lru_counter becomes 1, clear cycle is 10)
// ---------------------------------------------------------------------------------------------
struct Vec3i {
  int x, y, z;
  bool operator==(const Vec3i& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct XORVector3iHash {  // util/vector3i_hash.hpp:15-20
  size_t operator()(const Vec3i& v) const {
    const size_t p1 = 73856093, p2 = 19349669, p3 = 83492791;
    return static_cast<size_t>((v.x * p1) ^ (v.y * p2) ^ (v.z * p3));
  }
};
struct GaussianVoxel {
  bool finalized = false;
  size_t num_points = 0;
  V4 mean = v4_zero();
  M4 cov = m4_zero();
  Vec3i coord{0, 0, 0};
  void add(const V4& transformed_pt, const M4& transformed_cov) {  // gaussian_voxelmap.hpp:31-41
    if (finalized) {
      finalized = false;
      for (int d = 0; d < 4; d++) mean.v[d] *= num_points;
      for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) cov.m[r][c] *= num_points;
    }
    num_points++;
    mean = mean + transformed_pt;
    cov = add_m(cov, transformed_cov);
  }
  static M4 add_m(const M4& a, const M4& b) { return orc::add(a, b); }
  void finalize() {  // gaussian_voxelmap.hpp:44-52
    if (finalized) return;
    for (int d = 0; d < 4; d++) mean.v[d] /= num_points;
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) cov.m[r][c] /= num_points;
    finalized = true;
  }
};

struct GaussianVoxelMap {
  double inv_leaf_size;
  std::vector<Vec3i> search_offsets;
  std::vector<GaussianVoxel> flat_voxels;
  std::unordered_map<Vec3i, size_t, XORVector3iHash> voxels;
  explicit GaussianVoxelMap(double leaf_size) : inv_leaf_size(1.0 / leaf_size) { set_search_offsets(1); }
  size_t size() const { return flat_voxels.size(); }

  void set_search_offsets(int num_offsets) {  // incremental_voxelmap.hpp:157-186
    search_offsets.clear();
    switch (num_offsets) {
      default:
      case 1:
        search_offsets = {{0, 0, 0}};
        break;
      case 7:
        search_offsets = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {-1, 0, 0}, {0, -1, 0}, {0, 0, -1}};
        break;
      case 27:
        for (int i = -1; i <= 1; i++)
          for (int j = -1; j <= 1; j++)
            for (int k = -1; k <= 1; k++) search_offsets.push_back({i, j, k});
        break;
    }
  }

  void insert(const Cloud& points, const Iso& T) {  // incremental_voxelmap.hpp:55-92
    for (size_t i = 0; i < points.size(); i++) {
      const V4 pt = mul(T.T, points.points[i]);
      const Vec3i coord{fast_floor1(pt[0] * inv_leaf_size), fast_floor1(pt[1] * inv_leaf_size), fast_floor1(pt[2] * inv_leaf_size)};
      auto found = voxels.find(coord);
      if (found == voxels.end()) {
        found = voxels.emplace_hint(found, coord, flat_voxels.size());
        flat_voxels.emplace_back();
        flat_voxels.back().coord = coord;
      }
      const M4 tc = mul(mul(T.T, points.covs[i]), transpose(T.T));
      flat_voxels[found->second].add(pt, tc);
    }
    for (auto& v : flat_voxels) v.finalize();
  }

  size_t nearest_neighbor_search(const V4& pt, size_t* index, double* sq_dist) const {  // incremental_voxelmap.hpp:99-119
    const Vec3i center{fast_floor1(pt[0] * inv_leaf_size), fast_floor1(pt[1] * inv_leaf_size), fast_floor1(pt[2] * inv_leaf_size)};
    KnnResult result(index, sq_dist, 1);
    for (const auto& offset : search_offsets) {
      const Vec3i coord{center.x + offset.x, center.y + offset.y, center.z + offset.z};
      const auto found = voxels.find(coord);
      if (found == voxels.end()) continue;
      result.index_base = found->second << 32;                                 // calc_index(voxel_id, 0), :151
      result.push(0, sq_norm(flat_voxels[found->second].mean - pt));           // gaussian_voxelmap.hpp:83-85
    }
    return result.num_found();
  }
};

// ---------------------------------------------------------------------------------------------
// Targets: a (cloud, kd-tree) pair or a Gaussian voxel map (which is both "cloud" and "tree",
// registration_helper.cpp:136).
// ---------------------------------------------------------------------------------------------
struct CloudTarget {
  const Cloud* cloud;
  const KdTree* tree;
  size_t nn(const V4& q, size_t* k, double* d) const { return tree->nearest_neighbor_search(q, k, d); }
  const V4& point(size_t i) const { return cloud->points[i]; }
  const V4& normal(size_t i) const { return cloud->normals[i]; }
  const M4& cov(size_t i) const { return cloud->covs[i]; }
};
struct VoxelTarget {
  const GaussianVoxelMap* map;
  V4 zero = v4_zero();
  size_t nn(const V4& q, size_t* k, double* d) const { return map->nearest_neighbor_search(q, k, d); }
  const V4& point(size_t i) const { return map->flat_voxels[i >> 32].mean; }  // incremental_voxelmap.hpp:210-215
  const V4& normal(size_t) const { return zero; }
  const M4& cov(size_t i) const { return map->flat_voxels[i >> 32].cov; }
};

// ---------------------------------------------------------------------------------------------
// Factors: factors/icp_factor.hpp:14-70, plane_icp_factor.hpp:14-75, gicp_factor.hpp:14-97,
// robust_kernel.hpp:11-106 ; rejectors: registration/rejector.hpp:11-28
// ---------------------------------------------------------------------------------------------
enum FactorKind { FACTOR_ICP = 0, FACTOR_PLANE = 1, FACTOR_GICP = 2 };
enum RobustKind { ROBUST_NONE = 0, ROBUST_HUBER = 1, ROBUST_CAUCHY = 2 };
enum RejectorKind { REJECT_NONE = 0, REJECT_DISTANCE = 1 };
static constexpr size_t NO_INDEX = std::numeric_limits<size_t>::max();

struct Setting {
  int factor = FACTOR_GICP;
  int robust = ROBUST_NONE;
  double robust_c = 1.0;
  int rejector = REJECT_DISTANCE;
  double max_dist_sq = 1.0;
};

struct Factor {
  size_t target_index = NO_INDEX;
  size_t source_index = NO_INDEX;
  M4 mahalanobis = m4_zero();  // GICP only
  bool inlier() const { return target_index != NO_INDEX; }
};

static inline double robust_weight(const Setting& s, double e) {
  if (s.robust == ROBUST_HUBER) {  // robust_kernel.hpp:24-27
    const double e_abs = std::abs(e);
    return e_abs < s.robust_c ? 1.0 : s.robust_c / e_abs;
  }
  if (s.robust == ROBUST_CAUCHY) return s.robust_c / (s.robust_c + e * e);  // :47
  return 1.0;
}

/// J^T * W * J etc. with the reference's 4x6 Jacobian and 4x4 weights.
static inline void jtwj(const M46& J, const M4* W, const V4& r, M6* H, V6* b) {
  // WJ = W * J (4x6); Wr = W * r
  M46 WJ;
  V4 Wr;
  if (W) {
    for (int i = 0; i < 4; i++) {
      for (int j = 0; j < 6; j++) WJ.m[i][j] = W->m[i][0] * J.m[0][j] + W->m[i][1] * J.m[1][j] + W->m[i][2] * J.m[2][j] + W->m[i][3] * J.m[3][j];
      Wr.v[i] = W->m[i][0] * r[0] + W->m[i][1] * r[1] + W->m[i][2] * r[2] + W->m[i][3] * r[3];
    }
  } else {
    WJ = J;
    Wr = r;
  }
  for (int i = 0; i < 6; i++) {
    for (int j = 0; j < 6; j++) H->m[i][j] = J.m[0][i] * WJ.m[0][j] + J.m[1][i] * WJ.m[1][j] + J.m[2][i] * WJ.m[2][j] + J.m[3][i] * WJ.m[3][j];
    b->v[i] = J.m[0][i] * Wr[0] + J.m[1][i] * Wr[1] + J.m[2][i] * Wr[2] + J.m[3][i] * Wr[3];
  }
}

template <typename Target>
static bool factor_linearize(const Setting& s, Factor& f, const Target& target, const Cloud& source, const Iso& T, size_t source_index, M6* H, V6* b, double* e) {
  f.source_index = source_index;
  f.target_index = NO_INDEX;
  const V4 transed_source_pt = mul(T.T, source.points[source_index]);
  size_t k_index;
  double k_sq_dist;
  if (!target.nn(transed_source_pt, &k_index, &k_sq_dist)) return false;
  if (s.rejector == REJECT_DISTANCE && k_sq_dist > s.max_dist_sq) return false;  // rejector.hpp:23-25
  f.target_index = k_index;

  const V4 residual = target.point(k_index) - transed_source_pt;
  // J = [R * skew(p) | -R], 4x6 with a zero last row
  M46 J;
  std::memset(&J, 0, sizeof(J));
  const V4& sp = source.points[source_index];
  const double p3[3] = {sp[0], sp[1], sp[2]};
  const M3 S = skew(p3);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      J.m[i][j] = T.T.m[i][0] * S.m[0][j] + T.T.m[i][1] * S.m[1][j] + T.T.m[i][2] * S.m[2][j];
      J.m[i][3 + j] = -T.T.m[i][j];
    }

  if (s.factor == FACTOR_ICP) {  // icp_factor.hpp:45-51
    jtwj(J, nullptr, residual, H, b);
    *e = 0.5 * sq_norm(residual);
  } else if (s.factor == FACTOR_PLANE) {  // plane_icp_factor.hpp:44-55 (element-wise n .* r)
    const V4& n = target.normal(k_index);
    V4 err{{n[0] * residual[0], n[1] * residual[1], n[2] * residual[2], n[3] * residual[3]}};
    M46 Jn;
    std::memset(&Jn, 0, sizeof(Jn));
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 6; j++) Jn.m[i][j] = n[i] * J.m[i][j];
    jtwj(Jn, nullptr, err, H, b);
    *e = 0.5 * sq_norm(err);
  } else {  // gicp_factor.hpp:59-70
    const M4 RCR = add(target.cov(k_index), mul(mul(T.T, source.covs[source_index]), transpose(T.T)));
    M3 rcr3;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) rcr3.m[i][j] = RCR.m[i][j];
    const M3 inv = inverse3(rcr3);
    f.mahalanobis = m4_zero();
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) f.mahalanobis.m[i][j] = inv.m[i][j];
    jtwj(J, &f.mahalanobis, residual, H, b);
    const V4 Mr = mul(f.mahalanobis, residual);
    *e = 0.5 * (residual[0] * Mr[0] + residual[1] * Mr[1] + residual[2] * Mr[2] + residual[3] * Mr[3]);
  }

  if (s.robust != ROBUST_NONE) {  // robust_kernel.hpp:70-91
    const double w = robust_weight(s, std::sqrt(*e));
    for (int i = 0; i < 6; i++) {
      for (int j = 0; j < 6; j++) H->m[i][j] *= w;
      b->v[i] *= w;
    }
    *e *= w;
  }
  return true;
}

template <typename Target>
static double factor_error(const Setting& s, const Factor& f, const Target& target, const Cloud& source, const Iso& T) {
  if (f.target_index == NO_INDEX) return 0.0;
  const V4 transed_source_pt = mul(T.T, source.points[f.source_index]);
  const V4 residual = target.point(f.target_index) - transed_source_pt;
  double e;
  if (s.factor == FACTOR_ICP) {
    e = 0.5 * sq_norm(residual);  // icp_factor.hpp:57-64
  } else if (s.factor == FACTOR_PLANE) {
    const V4& n = target.normal(f.target_index);  // plane_icp_factor.hpp:60-69
    V4 err{{n[0] * residual[0], n[1] * residual[1], n[2] * residual[2], n[3] * residual[3]}};
    e = 0.5 * sq_norm(err);
  } else {
    const V4 Mr = mul(f.mahalanobis, residual);  // gicp_factor.hpp:81-89 (mahalanobis frozen at linearisation)
    e = 0.5 * (residual[0] * Mr[0] + residual[1] * Mr[1] + residual[2] * Mr[2] + residual[3] * Mr[3]);
  }
  if (s.robust != ROBUST_NONE) e = robust_weight(s, std::sqrt(e)) * e;  // robust_kernel.hpp:93-98
  return e;
}

// ---------------------------------------------------------------------------------------------
// Reductions: registration/reduction.hpp:20-62 (serial), reduction_omp.hpp:24-70 (OMP guided,8)
// ---------------------------------------------------------------------------------------------
struct Linearized {
  M6 H;
  V6 b;
  double e;
};

template <typename Target>
static Linearized reduce_linearize(const Setting& s, int num_threads, const Target& target, const Cloud& source, const Iso& T, std::vector<Factor>& factors) {
  if (num_threads <= 0) {  // SerialReduction
    Linearized sum{m6_zero(), v6_zero(), 0.0};
    for (size_t i = 0; i < factors.size(); i++) {
      M6 H;
      V6 b;
      double e;
      if (!factor_linearize(s, factors[i], target, source, T, i, &H, &b, &e)) continue;
      for (int r = 0; r < 6; r++) {
        for (int c = 0; c < 6; c++) sum.H.m[r][c] += H.m[r][c];
        sum.b.v[r] += b.v[r];
      }
      sum.e += e;
    }
    return sum;
  }
  // ParallelReductionOMP
  std::vector<M6> Hs(num_threads, m6_zero());
  std::vector<V6> bs(num_threads, v6_zero());
  std::vector<double> es(num_threads, 0.0);
  const std::int64_t N = factors.size();
#pragma omp parallel for num_threads(num_threads) schedule(guided, 8)
  for (std::int64_t i = 0; i < N; i++) {
    M6 H;
    V6 b;
    double e;
    if (!factor_linearize(s, factors[i], target, source, T, i, &H, &b, &e)) continue;
    const int tid = omp_get_thread_num();
    for (int r = 0; r < 6; r++) {
      for (int c = 0; c < 6; c++) Hs[tid].m[r][c] += H.m[r][c];
      bs[tid].v[r] += b.v[r];
    }
    es[tid] += e;
  }
  for (int t = 1; t < num_threads; t++) {
    for (int r = 0; r < 6; r++) {
      for (int c = 0; c < 6; c++) Hs[0].m[r][c] += Hs[t].m[r][c];
      bs[0].v[r] += bs[t].v[r];
    }
    es[0] += es[t];
  }
  return Linearized{Hs[0], bs[0], es[0]};
}

template <typename Target>
static double reduce_error(const Setting& s, int num_threads, const Target& target, const Cloud& source, const Iso& T, std::vector<Factor>& factors) {
  double sum_e = 0.0;
  const std::int64_t N = factors.size();
  if (num_threads <= 0) {
    for (std::int64_t i = 0; i < N; i++) sum_e += factor_error(s, factors[i], target, source, T);
    return sum_e;
  }
#pragma omp parallel for num_threads(num_threads) schedule(guided, 8) reduction(+ : sum_e)
  for (std::int64_t i = 0; i < N; i++) sum_e += factor_error(s, factors[i], target, source, T);
  return sum_e;
}

// ---------------------------------------------------------------------------------------------
// Optimizers: registration/optimizer.hpp:24-63 (GN), :83-149 (LM); termination_criteria.hpp:10-20;
// registration_result.hpp:11-30 ; general factor = NullFactor (general_factor.hpp:11-39)
// ---------------------------------------------------------------------------------------------
struct OptimizerSetting {
  int type = 1;  // 0 = GaussNewton, 1 = LevenbergMarquardt (registration.hpp:22 default)
  int max_iterations = 20;
  double gn_lambda = 1e-6;
  int max_inner_iterations = 10;
  double init_lambda = 1e-3;
  double lambda_factor = 10.0;
  double translation_eps = 1e-3;
  double rotation_eps = 0.1 * M_PI / 180.0;
};

struct Result {
  Iso T_target_source;
  int converged = 0;
  size_t iterations = 0;
  size_t num_inliers = 0;
  M6 H = m6_zero();
  V6 b = v6_zero();
  double error = 0.0;
};

/// Optional per-iteration trace (oracle extension used to freeze golden vectors: pose, H|b|e before the update).
struct Trace {
  std::vector<double> rows;  // per linearize: 16 (T, row-major) + 36 (H) + 6 (b) + 1 (e) = 59 doubles
  void add(const Iso& T, const Linearized& L) {
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) rows.push_back(T.T.m[i][j]);
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 6; j++) rows.push_back(L.H.m[i][j]);
    for (int i = 0; i < 6; i++) rows.push_back(L.b.v[i]);
    rows.push_back(L.e);
  }
};

static inline bool criteria_converged(const OptimizerSetting& o, const V6& d) {  // termination_criteria.hpp:17
  const double nr = std::sqrt(d.v[0] * d.v[0] + d.v[1] * d.v[1] + d.v[2] * d.v[2]);
  const double nt = std::sqrt(d.v[3] * d.v[3] + d.v[4] * d.v[4] + d.v[5] * d.v[5]);
  return nr <= o.rotation_eps && nt <= o.translation_eps;
}

static inline V6 solve_damped(const M6& H, const V6& b, double lambda) {
  M6 A = H;
  for (int i = 0; i < 6; i++) A.m[i][i] += lambda;
  V6 nb;
  for (int i = 0; i < 6; i++) nb.v[i] = -b.v[i];
  return ldlt_solve6(A, nb);
}

template <typename Target>
static Result optimize(const Setting& s, const OptimizerSetting& o, int num_threads, const Target& target, const Cloud& source, const Iso& init_T, std::vector<Factor>& factors, Trace* trace) {
  Result result;
  result.T_target_source = init_T;
  if (o.type == 0) {  // GaussNewtonOptimizer::optimize
    for (int i = 0; i < o.max_iterations && !result.converged; i++) {
      Linearized L = reduce_linearize(s, num_threads, target, source, result.T_target_source, factors);
      if (trace) trace->add(result.T_target_source, L);
      const V6 delta = solve_damped(L.H, L.b, o.gn_lambda);
      result.converged = criteria_converged(o, delta);
      result.T_target_source = iso_mul(result.T_target_source, se3_exp(delta));
      result.iterations = i;
      result.H = L.H;
      result.b = L.b;
      result.error = L.e;
    }
  } else {  // LevenbergMarquardtOptimizer::optimize
    double lambda = o.init_lambda;
    for (int i = 0; i < o.max_iterations && !result.converged; i++) {
      Linearized L = reduce_linearize(s, num_threads, target, source, result.T_target_source, factors);
      if (trace) trace->add(result.T_target_source, L);
      double e = L.e;
      bool success = false;
      for (int j = 0; j < o.max_inner_iterations; j++) {
        const V6 delta = solve_damped(L.H, L.b, lambda);
        const Iso new_T = iso_mul(result.T_target_source, se3_exp(delta));
        const double new_e = reduce_error(s, num_threads, target, source, new_T, factors);
        if (new_e <= e) {
          result.converged = criteria_converged(o, delta);
          result.T_target_source = new_T;
          lambda /= o.lambda_factor;
          success = true;
          e = new_e;
          break;
        } else {
          lambda *= o.lambda_factor;
        }
      }
      result.iterations = i;
      result.H = L.H;
      result.b = L.b;
      result.error = e;
      if (!success) break;
    }
  }
  result.num_inliers = std::count_if(factors.begin(), factors.end(), [](const Factor& f) { return f.inlier(); });
  return result;
}

/// A registration problem instance (factors persist between linearize() and error(), registration.hpp:41).
struct Registration {
  Setting setting;
  OptimizerSetting optimizer;
  int num_threads = 0;  // 0 = SerialReduction, >0 = ParallelReductionOMP with that many threads
  std::vector<Factor> factors;
  Trace trace;
};

}  // namespace orc

// =============================================================================================
// C API (ctypes).  Matrices cross the boundary ROW-MAJOR; clouds as N x 4 doubles.
// =============================================================================================
using namespace orc;

extern "C" {

int orc_max_threads() { return omp_get_max_threads(); }

// ---- clouds ----
Cloud* orc_cloud_create(size_t n, const double* xyz, int stride) {
  auto* c = new Cloud();
  c->resize(n);
  for (size_t i = 0; i < n; i++) {
    c->points[i] = V4{{xyz[i * stride + 0], xyz[i * stride + 1], xyz[i * stride + 2], 1.0}};
    c->normals[i] = v4_zero();
    c->covs[i] = m4_zero();
  }
  return c;
}
void orc_cloud_destroy(Cloud* c) { delete c; }
size_t orc_cloud_size(const Cloud* c) { return c->size(); }
void orc_cloud_get(const Cloud* c, double* points4, double* normals4, double* covs16) {
  const size_t n = c->size();
  if (points4) std::memcpy(points4, c->points.data(), n * sizeof(V4));
  if (normals4) std::memcpy(normals4, c->normals.data(), n * sizeof(V4));
  if (covs16) std::memcpy(covs16, c->covs.data(), n * sizeof(M4));
}
void orc_cloud_set_features(Cloud* c, const double* normals4, const double* covs16) {
  const size_t n = c->size();
  if (normals4) std::memcpy(c->normals.data(), normals4, n * sizeof(V4));
  if (covs16) std::memcpy(c->covs.data(), covs16, n * sizeof(M4));
}
Cloud* orc_cloud_transformed(const Cloud* c, const double* T16) {  // points only (registration_test.cpp:83-85)
  Iso T;
  std::memcpy(&T.T, T16, sizeof(M4));
  auto* o = new Cloud();
  o->resize(c->size());
  for (size_t i = 0; i < c->size(); i++) {
    o->points[i] = mul(T.T, c->points[i]);
    o->normals[i] = v4_zero();
    o->covs[i] = m4_zero();
  }
  return o;
}
Cloud* orc_voxelgrid_sampling(const Cloud* c, double leaf) { return voxelgrid_sampling(*c, leaf).release(); }

// ---- kd-tree ----
KdTree* orc_kdtree_create(const Cloud* c) {
  auto* t = new KdTree();
  t->build(*c);
  return t;
}
void orc_kdtree_destroy(KdTree* t) { delete t; }
size_t orc_kdtree_num_nodes(const KdTree* t) { return t->nodes.size(); }
/// raw 24-byte nodes (reference layout) + size_t indices, to feed sgb_target_set_kdtree
void orc_kdtree_export(const KdTree* t, void* nodes24, uint64_t* indices) {
  std::memcpy(nodes24, t->nodes.data(), t->nodes.size() * sizeof(KdTreeNode));
  for (size_t i = 0; i < t->indices.size(); i++) indices[i] = t->indices[i];
}
/// batch kNN: queries nq x 4, outputs nq x k (unfound slots: index = SIZE_MAX, distance = DBL_MAX); returns found counts
void orc_kdtree_knn(const KdTree* t, size_t nq, const double* q4, int k, uint64_t* idx, double* d2, uint64_t* counts, int num_threads) {
  const std::int64_t N = nq;
#pragma omp parallel for num_threads(num_threads > 0 ? num_threads : 1) schedule(guided, 8)
  for (std::int64_t i = 0; i < N; i++) {
    std::vector<size_t> ki(k);
    const V4 q{{q4[i * 4 + 0], q4[i * 4 + 1], q4[i * 4 + 2], q4[i * 4 + 3]}};
    const size_t n = t->knn_search(q, k, ki.data(), d2 + i * k);
    for (int j = 0; j < k; j++) idx[i * k + j] = ki[j];
    if (counts) counts[i] = n;
  }
}
void orc_estimate_features(Cloud* c, const KdTree* t, int k, int mode, int num_threads) { estimate_features(*c, *t, k, mode, num_threads); }

// ---- Gaussian voxel map ----
GaussianVoxelMap* orc_voxelmap_create(const Cloud* c, double leaf, int search_offsets) {
  auto* m = new GaussianVoxelMap(leaf);
  m->set_search_offsets(search_offsets);
  m->insert(*c, iso_identity());
  return m;
}
void orc_voxelmap_destroy(GaussianVoxelMap* m) { delete m; }
size_t orc_voxelmap_size(const GaussianVoxelMap* m) { return m->size(); }
void orc_voxelmap_export(const GaussianVoxelMap* m, int32_t* coords3, double* means4, double* covs16, uint64_t* counts) {
  for (size_t i = 0; i < m->size(); i++) {
    const auto& v = m->flat_voxels[i];
    coords3[i * 3 + 0] = v.coord.x;
    coords3[i * 3 + 1] = v.coord.y;
    coords3[i * 3 + 2] = v.coord.z;
    std::memcpy(means4 + i * 4, &v.mean, sizeof(V4));
    std::memcpy(covs16 + i * 16, &v.cov, sizeof(M4));
    if (counts) counts[i] = v.num_points;
  }
}
void orc_voxelmap_nn(const GaussianVoxelMap* m, size_t nq, const double* q4, uint64_t* idx, double* d2, uint64_t* counts) {
  for (size_t i = 0; i < nq; i++) {
    const V4 q{{q4[i * 4 + 0], q4[i * 4 + 1], q4[i * 4 + 2], q4[i * 4 + 3]}};
    size_t k = KnnResult::INVALID;
    double d = std::numeric_limits<double>::max();
    counts[i] = m->nearest_neighbor_search(q, &k, &d);
    idx[i] = k;
    d2[i] = d;
  }
}

// ---- registration ----
Registration* orc_reg_create(int factor, int robust, double robust_c, int rejector, double max_dist_sq, int num_threads) {
  auto* r = new Registration();
  r->setting.factor = factor;
  r->setting.robust = robust;
  r->setting.robust_c = robust_c;
  r->setting.rejector = rejector;
  r->setting.max_dist_sq = max_dist_sq;
  r->num_threads = num_threads;
  return r;
}
void orc_reg_destroy(Registration* r) { delete r; }
void orc_reg_set_optimizer(Registration* r, int type, int max_iterations, double gn_lambda, int max_inner, double init_lambda, double lambda_factor, double rot_eps, double trans_eps) {
  r->optimizer.type = type;
  r->optimizer.max_iterations = max_iterations;
  r->optimizer.gn_lambda = gn_lambda;
  r->optimizer.max_inner_iterations = max_inner;
  r->optimizer.init_lambda = init_lambda;
  r->optimizer.lambda_factor = lambda_factor;
  r->optimizer.rotation_eps = rot_eps;
  r->optimizer.translation_eps = trans_eps;
}

static void pack43(const Linearized& L, double* out43) {
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) out43[i * 6 + j] = L.H.m[i][j];
  for (int i = 0; i < 6; i++) out43[36 + i] = L.b.v[i];
  out43[42] = L.e;
}

/// One Reduction::linearize call; (re)allocates the factor vector like Registration::align does (registration.hpp:41).
/// Exactly one of (tree, voxelmap) is non-null.
void orc_reg_linearize(Registration* r, const Cloud* target, const KdTree* tree, const GaussianVoxelMap* vmap, const Cloud* source, const double* T16, double* out43) {
  Iso T;
  std::memcpy(&T.T, T16, sizeof(M4));
  if (r->factors.size() != source->size()) r->factors.assign(source->size(), Factor());
  Linearized L;
  if (vmap)
    L = reduce_linearize(r->setting, r->num_threads, VoxelTarget{vmap}, *source, T, r->factors);
  else
    L = reduce_linearize(r->setting, r->num_threads, CloudTarget{target, tree}, *source, T, r->factors);
  pack43(L, out43);
}
double orc_reg_error(Registration* r, const Cloud* target, const GaussianVoxelMap* vmap, const Cloud* source, const double* T16) {
  Iso T;
  std::memcpy(&T.T, T16, sizeof(M4));
  if (vmap) return reduce_error(r->setting, r->num_threads, VoxelTarget{vmap}, *source, T, r->factors);
  return reduce_error(r->setting, r->num_threads, CloudTarget{target, nullptr}, *source, T, r->factors);
}
void orc_reg_correspondences(const Registration* r, uint64_t* target_index) {
  for (size_t i = 0; i < r->factors.size(); i++) target_index[i] = r->factors[i].target_index;
}

/// Registration::align (registration.hpp:31-43). out: T (16, row-major), then scalars.
/// result_scalars = [converged, iterations, num_inliers, error]; H36, b6 optional.
void orc_reg_align(Registration* r, const Cloud* target, const KdTree* tree, const GaussianVoxelMap* vmap, const Cloud* source, const double* init_T16, int want_trace, double* T_out16, double* result_scalars, double* H36, double* b6) {
  Iso T;
  std::memcpy(&T.T, init_T16, sizeof(M4));
  r->factors.assign(source->size(), Factor());
  r->trace.rows.clear();
  Result res;
  if (vmap)
    res = optimize(r->setting, r->optimizer, r->num_threads, VoxelTarget{vmap}, *source, T, r->factors, want_trace ? &r->trace : nullptr);
  else
    res = optimize(r->setting, r->optimizer, r->num_threads, CloudTarget{target, tree}, *source, T, r->factors, want_trace ? &r->trace : nullptr);
  std::memcpy(T_out16, &res.T_target_source.T, sizeof(M4));
  result_scalars[0] = res.converged;
  result_scalars[1] = static_cast<double>(res.iterations);
  result_scalars[2] = static_cast<double>(res.num_inliers);
  result_scalars[3] = res.error;
  if (H36)
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 6; j++) H36[i * 6 + j] = res.H.m[i][j];
  if (b6)
    for (int i = 0; i < 6; i++) b6[i] = res.b.v[i];
}
size_t orc_reg_trace_rows(const Registration* r) { return r->trace.rows.size() / 59; }
void orc_reg_trace_get(const Registration* r, double* out) { std::memcpy(out, r->trace.rows.data(), r->trace.rows.size() * sizeof(double)); }

// ---- small algebra exports for unit tests ----
void orc_se3_exp(const double* a6, double* T16) {
  V6 a;
  std::memcpy(a.v, a6, sizeof(a.v));
  const Iso T = se3_exp(a);
  std::memcpy(T16, &T.T, sizeof(M4));
}
void orc_ldlt_solve6(const double* A36, const double* b6, double* x6) {
  M6 A;
  V6 b;
  std::memcpy(&A, A36, sizeof(A));
  std::memcpy(&b, b6, sizeof(b));
  const V6 x = ldlt_solve6(A, b);
  std::memcpy(x6, x.v, sizeof(x.v));
}
void orc_eigen_sym3(const double* A9, double* evals3, double* evecs9) {
  M3 A, V;
  std::memcpy(&A, A9, sizeof(A));
  eigen_sym3_direct(A, evals3, V);
  std::memcpy(evecs9, &V, sizeof(V));
}
}  // extern "C"
