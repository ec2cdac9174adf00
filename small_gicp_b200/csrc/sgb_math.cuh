// SPDX-License-Identifier: MIT
// Per-point arithmetic of the hot path: the factor algebra of ICP / point-to-plane / GICP (+ robust kernels), the GICP
// precision matrix, and the mapping of the 28 compact sums to H | b | e.  No memory access pattern, no warp primitive:
// everything here is plain C++ on scalars (device functions under nvcc, plain inline functions under g++), so the SAME source that the kernels inline is also compiled
// by g++ into tests/host_math/libsgb_host_math.so and checked against the oracle and the numpy leg without a GPU
// (tests/test_device_math_on_host.py).  That library is test infrastructure; the product never loads it.
#pragma once
#include <stdint.h>
#include <vector_types.h>  // float4

#if defined(__CUDACC__)
#define SGB_HD __device__ __forceinline__
#else
#include <cmath>
#define SGB_HD inline
#endif

namespace sgb {

constexpr int kAcc = 28;  // 6 (H_rr sym) + 9 (H_rt) + 6 (H_tt sym) + 6 (b) + 1 (e)

/// Symmetric 3x3 stored as (xx, xy, xz, yy, yz, zz).
struct Sym3 {
  double xx, xy, xz, yy, yz, zz;
};

SGB_HD Sym3 sym3_inverse(const Sym3& a) {
  // cofactors / determinant (the closed form Eigen uses for 3x3, gicp_factor.hpp:60)
  const double c00 = a.yy * a.zz - a.yz * a.yz;
  const double c01 = a.xz * a.yz - a.xy * a.zz;
  const double c02 = a.xy * a.yz - a.xz * a.yy;
  const double c11 = a.xx * a.zz - a.xz * a.xz;
  const double c12 = a.xy * a.xz - a.xx * a.yz;
  const double c22 = a.xx * a.yy - a.xy * a.xy;
  const double det = a.xx * c00 + a.xy * c01 + a.xz * c02;
  const double inv = 1.0 / det;
  return Sym3{c00 * inv, c01 * inv, c02 * inv, c11 * inv, c12 * inv, c22 * inv};
}

/// RCR = Ct + R Cs R^T (3x3 blocks of gicp_factor.hpp:59) ; returns its inverse (the fused precision matrix).
SGB_HD Sym3 gicp_precision(const double* R, const float4& sA, const float4& sB, const float4& tA, const float4& tB) {
  const double sxx = sA.x, sxy = sA.y, sxz = sA.z, syy = sA.w, syz = sB.x, szz = sB.y;
  // A = R * Cs
  double A[9];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double r0 = R[i * 3 + 0], r1 = R[i * 3 + 1], r2 = R[i * 3 + 2];
    A[i * 3 + 0] = r0 * sxx + r1 * sxy + r2 * sxz;
    A[i * 3 + 1] = r0 * sxy + r1 * syy + r2 * syz;
    A[i * 3 + 2] = r0 * sxz + r1 * syz + r2 * szz;
  }
  Sym3 rcr;
  rcr.xx = static_cast<double>(tA.x) + (A[0] * R[0] + A[1] * R[1] + A[2] * R[2]);
  rcr.xy = static_cast<double>(tA.y) + (A[0] * R[3] + A[1] * R[4] + A[2] * R[5]);
  rcr.xz = static_cast<double>(tA.z) + (A[0] * R[6] + A[1] * R[7] + A[2] * R[8]);
  rcr.yy = static_cast<double>(tA.w) + (A[3] * R[3] + A[4] * R[4] + A[5] * R[5]);
  rcr.yz = static_cast<double>(tB.x) + (A[3] * R[6] + A[4] * R[7] + A[5] * R[8]);
  rcr.zz = static_cast<double>(tB.y) + (A[6] * R[6] + A[7] * R[7] + A[8] * R[8]);
  return sym3_inverse(rcr);
}

/// Accumulate one point's  J^T M J | J^T M r | 1/2 r^T M r  (scaled by w) into acc[28], where
/// J = [R skew(p) | -R]  (icp_factor.hpp:45-47) and M is the 3x3 weight in the target frame:
/// identity (ICP), diag(n.^2) (point-to-plane, plane_icp_factor.hpp:46-55) or (Ct + R Cs R^T)^-1 (GICP).
/// Returns the unweighted error e.
template <int ROBUST>
SGB_HD void accumulate_factor(const double* R, const Sym3& M, double rx, double ry, double rz, double px, double py, double pz,
                                                  double robust_c, double* acc) {
  // Mr, e
  const double mrx = M.xx * rx + M.xy * ry + M.xz * rz;
  const double mry = M.xy * rx + M.yy * ry + M.yz * rz;
  const double mrz = M.xz * rx + M.yz * ry + M.zz * rz;
  const double e = 0.5 * (rx * mrx + ry * mry + rz * mrz);
  double w = 1.0;
  if (ROBUST == 1) {  // Huber, robust_kernel.hpp:24-27 on sqrt(e) (robust_kernel.hpp:84)
    const double x = sqrt(e);
    w = x < robust_c ? 1.0 : robust_c / x;
  } else if (ROBUST == 2) {  // Cauchy, robust_kernel.hpp:47 : c / (c + x^2) with x = sqrt(e)
    const double x = sqrt(e);
    w = robust_c / (robust_c + x * x);
  }
  // MR = M * R ; D = R^T * MR (symmetric)
  double MR[9];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const double r0 = R[0 + j], r1 = R[3 + j], r2 = R[6 + j];
    MR[0 + j] = M.xx * r0 + M.xy * r1 + M.xz * r2;
    MR[3 + j] = M.xy * r0 + M.yy * r1 + M.yz * r2;
    MR[6 + j] = M.xz * r0 + M.yz * r1 + M.zz * r2;
  }
  const double d00 = w * (R[0] * MR[0] + R[3] * MR[3] + R[6] * MR[6]);
  const double d01 = w * (R[0] * MR[1] + R[3] * MR[4] + R[6] * MR[7]);
  const double d02 = w * (R[0] * MR[2] + R[3] * MR[5] + R[6] * MR[8]);
  const double d11 = w * (R[1] * MR[1] + R[4] * MR[4] + R[7] * MR[7]);
  const double d12 = w * (R[1] * MR[2] + R[4] * MR[5] + R[7] * MR[8]);
  const double d22 = w * (R[2] * MR[2] + R[5] * MR[5] + R[8] * MR[8]);
  // g = w * R^T (M r)
  const double gx = w * (R[0] * mrx + R[3] * mry + R[6] * mrz);
  const double gy = w * (R[1] * mrx + R[4] * mry + R[7] * mrz);
  const double gz = w * (R[2] * mrx + R[5] * mry + R[8] * mrz);
  // U = skew(p) * D : column j = p x D[:, j]      (H_rt = U)
  const double u00 = py * d02 - pz * d01, u01 = py * d12 - pz * d11, u02 = py * d22 - pz * d12;
  const double u10 = pz * d00 - px * d02, u11 = pz * d01 - px * d12, u12 = pz * d02 - px * d22;
  const double u20 = px * d01 - py * d00, u21 = px * d11 - py * d01, u22 = px * d12 - py * d02;
  // H_rr = skew(p)^T D skew(p) : row i = p x U[i, :]
  acc[0] += py * u02 - pz * u01;
  acc[1] += pz * u00 - px * u02;
  acc[2] += px * u01 - py * u00;
  acc[3] += pz * u10 - px * u12;
  acc[4] += px * u11 - py * u10;
  acc[5] += px * u21 - py * u20;
  acc[6] += u00;
  acc[7] += u01;
  acc[8] += u02;
  acc[9] += u10;
  acc[10] += u11;
  acc[11] += u12;
  acc[12] += u20;
  acc[13] += u21;
  acc[14] += u22;
  acc[15] += d00;
  acc[16] += d01;
  acc[17] += d02;
  acc[18] += d11;
  acc[19] += d12;
  acc[20] += d22;
  // b = J^T M r = [ g x p ; -g ]
  acc[21] += gy * pz - gz * py;
  acc[22] += gz * px - gx * pz;
  acc[23] += gx * py - gy * px;
  acc[24] -= gx;
  acc[25] -= gy;
  acc[26] -= gz;
  acc[27] += w * e;
}

/// GICP weight in the SOURCE frame.  With R orthogonal,  R^T (Ct + R Cs R^T)^-1 R = (R^T Ct R + Cs)^-1 =: D.  D is all the
/// Hessian blocks need (H_tt = D, H_rt = skew(p) D, H_rr = skew(p)^T D skew(p)), and  R^T M r = D (R^T r),
/// r^T M r = (R^T r)^T D (R^T r):  the target-frame precision matrix M of gicp_factor.hpp:59-60 never has to be formed
/// (~50 of the ~240 FP64 operations per point of the target-frame formulation above; same value up to rounding).
SGB_HD Sym3 gicp_precision_source(const double* R, const float4& sA, const float4& sB, const float4& tA, const float4& tB) {
  const double txx = tA.x, txy = tA.y, txz = tA.z, tyy = tA.w, tyz = tB.x, tzz = tB.y;
  double B[9];  // B = Ct * R
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const double r0 = R[0 + j], r1 = R[3 + j], r2 = R[6 + j];
    B[0 + j] = txx * r0 + txy * r1 + txz * r2;
    B[3 + j] = txy * r0 + tyy * r1 + tyz * r2;
    B[6 + j] = txz * r0 + tyz * r1 + tzz * r2;
  }
  Sym3 n;  // N = Cs + R^T B
  n.xx = static_cast<double>(sA.x) + (R[0] * B[0] + R[3] * B[3] + R[6] * B[6]);
  n.xy = static_cast<double>(sA.y) + (R[0] * B[1] + R[3] * B[4] + R[6] * B[7]);
  n.xz = static_cast<double>(sA.z) + (R[0] * B[2] + R[3] * B[5] + R[6] * B[8]);
  n.yy = static_cast<double>(sA.w) + (R[1] * B[1] + R[4] * B[4] + R[7] * B[7]);
  n.yz = static_cast<double>(sB.x) + (R[1] * B[2] + R[4] * B[5] + R[7] * B[8]);
  n.zz = static_cast<double>(sB.y) + (R[2] * B[2] + R[5] * B[5] + R[8] * B[8]);
  return sym3_inverse(n);
}

/// Point-to-plane weight diag(n.^2) (plane_icp_factor.hpp:46-55) carried to the source frame: D = R^T diag(n.^2) R.
SGB_HD Sym3 plane_weight_source(const double* R, double nx, double ny, double nz) {
  const double a = nx * nx, b = ny * ny, c = nz * nz;
  Sym3 d;
  d.xx = a * R[0] * R[0] + b * R[3] * R[3] + c * R[6] * R[6];
  d.xy = a * R[0] * R[1] + b * R[3] * R[4] + c * R[6] * R[7];
  d.xz = a * R[0] * R[2] + b * R[3] * R[5] + c * R[6] * R[8];
  d.yy = a * R[1] * R[1] + b * R[4] * R[4] + c * R[7] * R[7];
  d.yz = a * R[1] * R[2] + b * R[4] * R[5] + c * R[7] * R[8];
  d.zz = a * R[2] * R[2] + b * R[5] * R[5] + c * R[8] * R[8];
  return d;
}

/// Same sums as accumulate_factor, from the source-frame weight D = R^T M R and the source-frame residual rs = R^T r.
template <int ROBUST>
SGB_HD void accumulate_factor_source(const Sym3& D, double rsx, double rsy, double rsz, double px, double py, double pz, double robust_c,
                                                         double* acc) {
  const double mrx = D.xx * rsx + D.xy * rsy + D.xz * rsz;
  const double mry = D.xy * rsx + D.yy * rsy + D.yz * rsz;
  const double mrz = D.xz * rsx + D.yz * rsy + D.zz * rsz;
  const double e = 0.5 * (rsx * mrx + rsy * mry + rsz * mrz);
  double w = 1.0;
  if (ROBUST == 1) {  // Huber, robust_kernel.hpp:24-27 on sqrt(e) (robust_kernel.hpp:84)
    const double x = sqrt(e);
    w = x < robust_c ? 1.0 : robust_c / x;
  } else if (ROBUST == 2) {  // Cauchy, robust_kernel.hpp:47 : c / (c + x^2) with x = sqrt(e)
    const double x = sqrt(e);
    w = robust_c / (robust_c + x * x);
  }
  const double d00 = w * D.xx, d01 = w * D.xy, d02 = w * D.xz, d11 = w * D.yy, d12 = w * D.yz, d22 = w * D.zz;
  const double gx = w * mrx, gy = w * mry, gz = w * mrz;  // g = w R^T M r
  const double u00 = py * d02 - pz * d01, u01 = py * d12 - pz * d11, u02 = py * d22 - pz * d12;
  const double u10 = pz * d00 - px * d02, u11 = pz * d01 - px * d12, u12 = pz * d02 - px * d22;
  const double u20 = px * d01 - py * d00, u21 = px * d11 - py * d01, u22 = px * d12 - py * d02;
  acc[0] += py * u02 - pz * u01;
  acc[1] += pz * u00 - px * u02;
  acc[2] += px * u01 - py * u00;
  acc[3] += pz * u10 - px * u12;
  acc[4] += px * u11 - py * u10;
  acc[5] += px * u21 - py * u20;
  acc[6] += u00;
  acc[7] += u01;
  acc[8] += u02;
  acc[9] += u10;
  acc[10] += u11;
  acc[11] += u12;
  acc[12] += u20;
  acc[13] += u21;
  acc[14] += u22;
  acc[15] += d00;
  acc[16] += d01;
  acc[17] += d02;
  acc[18] += d11;
  acc[19] += d12;
  acc[20] += d22;
  acc[21] += gy * pz - gz * py;
  acc[22] += gz * px - gx * pz;
  acc[23] += gx * py - gy * px;
  acc[24] -= gx;
  acc[25] -= gy;
  acc[26] -= gz;
  acc[27] += w * e;
}

/// One accepted-or-rejected correspondence of `linearize`, evaluated exactly as factor_reduce_kernel does it:
///   sp        source point, FP32, relative to the source centre (csx, csy, csz)
///   tq        matched target point, FP32, relative to the target centre
///   (tpx, tpy, tpz) = R c_s + t - c_t : the translation with both centres folded in (FP64)
///   sA, sB    source covariance (GICP);  t1, t2 = target normal, - (point-to-plane) or target covariance (GICP): pointers,
///             read only by the factor that needs them and only for accepted points (unused ones may be null)
/// Transform and residual in FP64, DistanceRejector on the FP64 residual (rejector.hpp:24: d2 > max rejects), weight and
/// residual carried to the SOURCE frame, sums into acc[0..27], inlier count into acc[28].  Returns false if rejected.
template <int FACTOR, int ROBUST>
SGB_HD bool point_linearize_source(const double* R, double tpx, double tpy, double tpz, double csx, double csy, double csz, double max_dist_sq_d,
                                   double robust_c, const float4& sp, const float4* sA, const float4* sB, const float4& tq, const float4* t1,
                                   const float4* t2, double* acc) {
  const double sx = sp.x, sy = sp.y, sz = sp.z;
  const double qx = R[0] * sx + R[1] * sy + R[2] * sz + tpx;
  const double qy = R[3] * sx + R[4] * sy + R[5] * sz + tpy;
  const double qz = R[6] * sx + R[7] * sy + R[8] * sz + tpz;
  const double rx = static_cast<double>(tq.x) - qx, ry = static_cast<double>(tq.y) - qy, rz = static_cast<double>(tq.z) - qz;
  if (rx * rx + ry * ry + rz * rz > max_dist_sq_d) return false;  // DistanceRejector on the FP64 residual (rejector.hpp:24)
  // weight and residual in the SOURCE frame (D = R^T M R, rs = R^T r): see gicp_precision_source
  Sym3 D;
  if (FACTOR == 0) {
    D = Sym3{1.0, 0.0, 0.0, 1.0, 0.0, 1.0};
  } else if (FACTOR == 1) {
    const float4 nrm = *t1;
    D = plane_weight_source(R, nrm.x, nrm.y, nrm.z);
  } else {
    D = gicp_precision_source(R, *sA, *sB, *t1, *t2);
  }
  const double rsx = R[0] * rx + R[3] * ry + R[6] * rz, rsy = R[1] * rx + R[4] * ry + R[7] * rz, rsz = R[2] * rx + R[5] * ry + R[8] * rz;
  accumulate_factor_source<ROBUST>(D, rsx, rsy, rsz, csx + sx, csy + sy, csz + sz, robust_c, acc);
  acc[kAcc] += 1.0;
  return true;
}

/// One cached correspondence of `error` (reduction.hpp:55-62): residual at the trial pose R | t', GICP precision matrix
/// re-derived from the rotation of the LAST LINEARIZE (gicp_factor.hpp:81-89 keeps the cached mahalanobis), robust weight.
template <int FACTOR, int ROBUST>
SGB_HD double point_error(const double* R, double tpx, double tpy, double tpz, const double* Rlin, double robust_c, const float4& sp, const float4& sA,
                          const float4& sB, const float4& tq, const float4& t1, const float4& t2) {
  const double sx = sp.x, sy = sp.y, sz = sp.z;
  const double qx = R[0] * sx + R[1] * sy + R[2] * sz + tpx;
  const double qy = R[3] * sx + R[4] * sy + R[5] * sz + tpy;
  const double qz = R[6] * sx + R[7] * sy + R[8] * sz + tpz;
  const double rx = static_cast<double>(tq.x) - qx, ry = static_cast<double>(tq.y) - qy, rz = static_cast<double>(tq.z) - qz;
  double e;
  if (FACTOR == 0) {
    e = 0.5 * (rx * rx + ry * ry + rz * rz);
  } else if (FACTOR == 1) {
    const double ex = static_cast<double>(t1.x) * rx, ey = static_cast<double>(t1.y) * ry, ez = static_cast<double>(t1.z) * rz;
    e = 0.5 * (ex * ex + ey * ey + ez * ez);
  } else {
    const Sym3 M = gicp_precision(Rlin, sA, sB, t1, t2);
    const double mrx = M.xx * rx + M.xy * ry + M.xz * rz;
    const double mry = M.xy * rx + M.yy * ry + M.yz * rz;
    const double mrz = M.xz * rx + M.yz * ry + M.zz * rz;
    e = 0.5 * (rx * mrx + ry * mry + rz * mrz);
  }
  if (ROBUST == 1) {
    const double x = sqrt(e);
    e *= (x < robust_c ? 1.0 : robust_c / x);
  } else if (ROBUST == 2) {
    const double x = sqrt(e);
    e *= robust_c / (robust_c + x * x);
  }
  return e;
}

/// Where compact sum k (0..27, 28 = inlier count) goes in the expanded output  H (6x6 row-major) | b (6) | e | inliers:
/// p0 always, p1 too when the function returns 2 (the mirrored entry of the symmetric H).
SGB_HD int expand_positions(int k, int& p0, int& p1) {
  p1 = -1;
  if (k < 6) {  // H_rr upper triangle: (0,0)(0,1)(0,2)(1,1)(1,2)(2,2)
    const int r = k < 3 ? 0 : (k < 5 ? 1 : 2);
    const int c = k < 3 ? k : (k < 5 ? k - 2 : 2);
    p0 = r * 6 + c;
    p1 = c * 6 + r;
    return 2;
  }
  if (k < 15) {  // H_rt 3x3 row-major -> rows 0..2, cols 3..5 (+ transpose)
    const int r = (k - 6) / 3, c = (k - 6) % 3;
    p0 = r * 6 + 3 + c;
    p1 = (3 + c) * 6 + r;
    return 2;
  }
  if (k < 21) {  // H_tt upper triangle
    const int kk = k - 15;
    const int r = kk < 3 ? 0 : (kk < 5 ? 1 : 2);
    const int c = kk < 3 ? kk : (kk < 5 ? kk - 2 : 2);
    p0 = (3 + r) * 6 + 3 + c;
    p1 = (3 + c) * 6 + 3 + r;
    return 2;
  }
  if (k < 27) {
    p0 = 36 + (k - 21);  // b
    return 1;
  }
  p0 = k == 27 ? 42 : 43;  // e, inlier count
  return 1;
}

}  // namespace sgb
