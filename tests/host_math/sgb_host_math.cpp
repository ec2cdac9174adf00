// SPDX-License-Identifier: MIT
// TEST INFRASTRUCTURE -- not part of the product, never loaded by small_gicp_b200.
// The per-point arithmetic the CUDA kernels inline (small_gicp_b200/csrc/sgb_math.cuh: factor algebra, GICP precision matrix,
// robust kernels, compact-sum layout) compiled for the HOST from the same source, behind a small C entry point, so that
// tests/test_device_math_on_host.py can hold it against the oracle and the numpy leg on a machine without a GPU.
// Inputs arrive in the reference's host layout (Vector4d points, Matrix4d covariances, already matched pairs) and are
// converted exactly as the device upload does it (convert_kernel in sgb_kernels.cu): coordinates relative to the cloud's
// centre rounded to FP32, the 6 unique covariance entries rounded to FP32.
#include <cstddef>
#include <cstring>

#include "sgb_math.cuh"

namespace {

struct Frame {
  double R[9], tp[3];
};

// R row-major + t' = R c_s + t - c_t (fill_params / the kernels' prologue)
Frame make_frame(const double* T_colmajor16, const double* cs, const double* ct) {
  Frame f;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) f.R[r * 3 + c] = T_colmajor16[c * 4 + r];
  for (int r = 0; r < 3; r++) f.tp[r] = f.R[r * 3 + 0] * cs[0] + f.R[r * 3 + 1] * cs[1] + f.R[r * 3 + 2] * cs[2] + T_colmajor16[12 + r] - ct[r];
  return f;
}

float4 centred(const double* p4, const double* c) {
  return float4{static_cast<float>(p4[0] - c[0]), static_cast<float>(p4[1] - c[1]), static_cast<float>(p4[2] - c[2]), 0.0f};
}
void cov6(const double* m16, float4& a, float4& b) {  // Matrix4d, column-major -> (xx,xy,xz,yy) (yz,zz,0,0): pack_covA / pack_covB of sgb_kernels.cu
  a = float4{static_cast<float>(m16[0]), static_cast<float>(m16[1]), static_cast<float>(m16[2]), static_cast<float>(m16[5])};
  b = float4{static_cast<float>(m16[6]), static_cast<float>(m16[10]), 0.0f, 0.0f};
}

template <int FACTOR, int ROBUST>
void run_linearize(const Frame& f, const double* cs, double max_dist_sq, double robust_c, size_t n, const double* src_pts4, const double* src_covs16,
                   const double* tgt_pts4, const double* tgt_normals4, const double* tgt_covs16, const double* ct, double* out44, unsigned char* accepted) {
  double acc[sgb::kAcc + 1] = {0.0};
  for (size_t i = 0; i < n; i++) {
    const float4 sp = centred(src_pts4 + 4 * i, cs), tq = centred(tgt_pts4 + 4 * i, ct);
    float4 sA{}, sB{}, t1{}, t2{};
    if (FACTOR == 1) t1 = float4{static_cast<float>(tgt_normals4[4 * i]), static_cast<float>(tgt_normals4[4 * i + 1]), static_cast<float>(tgt_normals4[4 * i + 2]), 0.0f};
    if (FACTOR == 2) {
      cov6(src_covs16 + 16 * i, sA, sB);
      cov6(tgt_covs16 + 16 * i, t1, t2);
    }
    const bool ok = sgb::point_linearize_source<FACTOR, ROBUST>(f.R, f.tp[0], f.tp[1], f.tp[2], cs[0], cs[1], cs[2], max_dist_sq, robust_c, sp, &sA, &sB, tq, &t1, &t2, acc);
    if (accepted) accepted[i] = ok ? 1 : 0;
  }
  std::memset(out44, 0, 44 * sizeof(double));
  for (int k = 0; k <= sgb::kAcc; k++) {  // the finishing CTA's expansion (block_reduce_and_finish)
    int p0, p1;
    if (sgb::expand_positions(k, p0, p1) == 2) out44[p1] = acc[k];
    out44[p0] = acc[k];
  }
}

// target-frame formulation (fused / voxel kernels: gicp_precision + accumulate_factor)
template <int FACTOR, int ROBUST>
void run_linearize_target_frame(const Frame& f, const double* cs, double max_dist_sq, double robust_c, size_t n, const double* src_pts4, const double* src_covs16,
                                const double* tgt_pts4, const double* tgt_normals4, const double* tgt_covs16, const double* ct, double* out44) {
  double acc[sgb::kAcc + 1] = {0.0};
  for (size_t i = 0; i < n; i++) {
    const float4 sp = centred(src_pts4 + 4 * i, cs), tq = centred(tgt_pts4 + 4 * i, ct);
    const double sx = sp.x, sy = sp.y, sz = sp.z;
    const double qx = f.R[0] * sx + f.R[1] * sy + f.R[2] * sz + f.tp[0], qy = f.R[3] * sx + f.R[4] * sy + f.R[5] * sz + f.tp[1],
                 qz = f.R[6] * sx + f.R[7] * sy + f.R[8] * sz + f.tp[2];
    const double rx = static_cast<double>(tq.x) - qx, ry = static_cast<double>(tq.y) - qy, rz = static_cast<double>(tq.z) - qz;
    if (rx * rx + ry * ry + rz * rz > max_dist_sq) continue;
    sgb::Sym3 M;
    if (FACTOR == 0) {
      M = sgb::Sym3{1.0, 0.0, 0.0, 1.0, 0.0, 1.0};
    } else if (FACTOR == 1) {
      const float nx = static_cast<float>(tgt_normals4[4 * i]), ny = static_cast<float>(tgt_normals4[4 * i + 1]), nz = static_cast<float>(tgt_normals4[4 * i + 2]);
      M = sgb::Sym3{static_cast<double>(nx) * nx, 0.0, 0.0, static_cast<double>(ny) * ny, 0.0, static_cast<double>(nz) * nz};
    } else {
      float4 sA, sB, tA, tB;
      cov6(src_covs16 + 16 * i, sA, sB);
      cov6(tgt_covs16 + 16 * i, tA, tB);
      M = sgb::gicp_precision(f.R, sA, sB, tA, tB);
    }
    sgb::accumulate_factor<ROBUST>(f.R, M, rx, ry, rz, cs[0] + sx, cs[1] + sy, cs[2] + sz, robust_c, acc);
    acc[sgb::kAcc] += 1.0;
  }
  std::memset(out44, 0, 44 * sizeof(double));
  for (int k = 0; k <= sgb::kAcc; k++) {
    int p0, p1;
    if (sgb::expand_positions(k, p0, p1) == 2) out44[p1] = acc[k];
    out44[p0] = acc[k];
  }
}

template <int FACTOR, int ROBUST>
double run_error(const Frame& f, const Frame& flin, const double* cs, double robust_c, size_t n, const double* src_pts4, const double* src_covs16,
                 const double* tgt_pts4, const double* tgt_normals4, const double* tgt_covs16, const double* ct) {
  double e = 0.0;
  for (size_t i = 0; i < n; i++) {
    const float4 sp = centred(src_pts4 + 4 * i, cs), tq = centred(tgt_pts4 + 4 * i, ct);
    float4 sA{}, sB{}, t1{}, t2{};
    if (FACTOR == 1) t1 = float4{static_cast<float>(tgt_normals4[4 * i]), static_cast<float>(tgt_normals4[4 * i + 1]), static_cast<float>(tgt_normals4[4 * i + 2]), 0.0f};
    if (FACTOR == 2) {
      cov6(src_covs16 + 16 * i, sA, sB);
      cov6(tgt_covs16 + 16 * i, t1, t2);
    }
    e += sgb::point_error<FACTOR, ROBUST>(f.R, f.tp[0], f.tp[1], f.tp[2], flin.R, robust_c, sp, sA, sB, tq, t1, t2);
  }
  return e;
}

}  // namespace

#define SGBM_DISPATCH(FN, ...)                 \
  switch (factor * 3 + robust) {               \
    case 0: FN<0, 0>(__VA_ARGS__); break;      \
    case 1: FN<0, 1>(__VA_ARGS__); break;      \
    case 2: FN<0, 2>(__VA_ARGS__); break;      \
    case 3: FN<1, 0>(__VA_ARGS__); break;      \
    case 4: FN<1, 1>(__VA_ARGS__); break;      \
    case 5: FN<1, 2>(__VA_ARGS__); break;      \
    case 6: FN<2, 0>(__VA_ARGS__); break;      \
    case 7: FN<2, 1>(__VA_ARGS__); break;      \
    case 8: FN<2, 2>(__VA_ARGS__); break;      \
    default: return 1;                         \
  }

extern "C" {

/// Sums of `linearize` over n MATCHED pairs (pair i = source point i with the target point/normal/covariance given at row i).
/// frame: 0 = source-frame formulation (factor_reduce_kernel), 1 = target-frame formulation (fused / voxel kernels).
/// out44 = H (36, row-major) | b (6) | e | accepted count; accepted (n bytes, or NULL; frame 0 only) = 1 where the rejector kept the pair.
int sgbm_linearize_pairs(int factor, int robust, double robust_c, double max_dist_sq, const double* T_colmajor16, const double* src_centre3,
                         const double* tgt_centre3, size_t n, const double* src_pts4, const double* src_covs16, const double* tgt_pts4,
                         const double* tgt_normals4, const double* tgt_covs16, int frame, double* out44, unsigned char* accepted) {
  if (factor < 0 || factor > 2 || robust < 0 || robust > 2) return 1;
  const Frame f = make_frame(T_colmajor16, src_centre3, tgt_centre3);
  if (frame == 0) {
    SGBM_DISPATCH(run_linearize, f, src_centre3, max_dist_sq, robust_c, n, src_pts4, src_covs16, tgt_pts4, tgt_normals4, tgt_covs16, tgt_centre3, out44, accepted)
  } else {
    SGBM_DISPATCH(run_linearize_target_frame, f, src_centre3, max_dist_sq, robust_c, n, src_pts4, src_covs16, tgt_pts4, tgt_normals4, tgt_covs16, tgt_centre3, out44)
  }
  return 0;
}

/// `error` over n matched pairs at the trial pose T, GICP precision frozen at T_lin.
int sgbm_error_pairs(int factor, int robust, double robust_c, const double* T_colmajor16, const double* Tlin_colmajor16, const double* src_centre3,
                     const double* tgt_centre3, size_t n, const double* src_pts4, const double* src_covs16, const double* tgt_pts4, const double* tgt_normals4,
                     const double* tgt_covs16, double* out_e) {
  if (factor < 0 || factor > 2 || robust < 0 || robust > 2) return 1;
  const Frame f = make_frame(T_colmajor16, src_centre3, tgt_centre3), fl = make_frame(Tlin_colmajor16, src_centre3, tgt_centre3);
  double e = 0.0;
  switch (factor * 3 + robust) {
    case 0: e = run_error<0, 0>(f, fl, src_centre3, robust_c, n, src_pts4, src_covs16, tgt_pts4, tgt_normals4, tgt_covs16, tgt_centre3); break;
    case 1: e = run_error<0, 1>(f, fl, src_centre3, robust_c, n, src_pts4, src_covs16, tgt_pts4, tgt_normals4, tgt_covs16, tgt_centre3); break;
    case 2: e = run_error<0, 2>(f, fl, src_centre3, robust_c, n, src_pts4, src_covs16, tgt_pts4, tgt_normals4, tgt_covs16, tgt_centre3); break;
    case 3: e = run_error<1, 0>(f, fl, src_centre3, robust_c, n, src_pts4, src_covs16, tgt_pts4, tgt_normals4, tgt_covs16, tgt_centre3); break;
    case 4: e = run_error<1, 1>(f, fl, src_centre3, robust_c, n, src_pts4, src_covs16, tgt_pts4, tgt_normals4, tgt_covs16, tgt_centre3); break;
    case 5: e = run_error<1, 2>(f, fl, src_centre3, robust_c, n, src_pts4, src_covs16, tgt_pts4, tgt_normals4, tgt_covs16, tgt_centre3); break;
    case 6: e = run_error<2, 0>(f, fl, src_centre3, robust_c, n, src_pts4, src_covs16, tgt_pts4, tgt_normals4, tgt_covs16, tgt_centre3); break;
    case 7: e = run_error<2, 1>(f, fl, src_centre3, robust_c, n, src_pts4, src_covs16, tgt_pts4, tgt_normals4, tgt_covs16, tgt_centre3); break;
    default: e = run_error<2, 2>(f, fl, src_centre3, robust_c, n, src_pts4, src_covs16, tgt_pts4, tgt_normals4, tgt_covs16, tgt_centre3); break;
  }
  *out_e = e;
  return 0;
}

}  // extern "C"
