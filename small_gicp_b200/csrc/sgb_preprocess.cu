// SPDX-License-Identifier: MIT
// Per-cloud preparation on the device:
//   * k-nearest-neighbour search of every cloud point in its own kd-tree + local covariance + closed-form 3x3
//     symmetric eigen-decomposition -> normal and regularised covariance
//       replaces estimate_normals / estimate_covariances / estimate_normals_covariances{,_omp,_tbb}
//       (/root/reference/include/small_gicp/util/normal_estimation.hpp:12-140, normal_estimation_omp.hpp:9-60)
//   * voxel-grid down-sampling: 3 x 21-bit voxel key, radix sort, segmented mean
//       replaces voxelgrid_sampling{,_omp,_tbb} (.../util/downsampling.hpp:22-78)
#include <cfloat>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "sgb_device.cuh"
#include "sgb_kernels.h"

namespace sgb {

// ---------------------------------------------------------------------------------------------
// k-NN with a register-resident sorted candidate list (ann/knn_result.hpp:80-101 semantics: only strictly
// closer candidates enter, equal distances keep their arrival order).
// ---------------------------------------------------------------------------------------------
template <int KMAX>
struct KnnList {
  float d[KMAX];
  uint32_t i[KMAX];
  float worst;
  int k;
  __device__ __forceinline__ void init(int k_) {
    k = k_;
#pragma unroll
    for (int j = 0; j < KMAX; j++) {
      d[j] = FLT_MAX;
      i[j] = kNone;
    }
    worst = FLT_MAX;
  }
  __device__ __forceinline__ void offer(float v, uint32_t vi) {
    if (!(v < worst)) return;
#pragma unroll
    for (int j = 0; j < KMAX; j++) {
      if (j < k && v < d[j]) {
        const float td = d[j];
        const uint32_t ti = i[j];
        d[j] = v;
        i[j] = vi;
        v = td;
        vi = ti;
      }
    }
#pragma unroll
    for (int j = 0; j < KMAX; j++)
      if (j == k - 1) worst = d[j];
  }
};

__device__ __forceinline__ float pre_box_dist2(float qx, float qy, float qz, const float4& lo, const float4& hi) {
  const float dx = fmaxf(fmaxf(lo.x - qx, qx - hi.x), 0.0f);
  const float dy = fmaxf(fmaxf(lo.y - qy, qy - hi.y), 0.0f);
  const float dz = fmaxf(fmaxf(lo.z - qz, qz - hi.z), 0.0f);
  return dx * dx + dy * dy + dz * dz;
}

/// Exact k-NN of one query per thread in the packet (BVH2) records of the cloud's tree (sgb_kernels_packet.cu has the
/// record layout): near child first, the far child is kept pending only while its box is closer than the k-th distance.
/// stack entry s of thread t: desc[s * kLinBlock + t] (child descriptor) and dist[s * kLinBlock + t] (its box distance).
template <int KMAX>
__device__ __forceinline__ void bvh_knn(const float4* __restrict__ pnodes, const float4* __restrict__ pts, float qx, float qy, float qz, KnnList<KMAX>& L,
                                        uint2* desc, float* dist) {
  int sp = 0;
  uint2* my_desc = desc + threadIdx.x;
  float* my_dist = dist + threadIdx.x;
  uint32_t cur = 0;
  bool expand = true;
  uint2 leaf = make_uint2(0u, 0u);
  for (;;) {
    if (expand) {
      const float4 n0 = __ldg(&pnodes[cur * 4 + 0]), n1 = __ldg(&pnodes[cur * 4 + 1]);
      const float4 n2 = __ldg(&pnodes[cur * 4 + 2]), n3 = __ldg(&pnodes[cur * 4 + 3]);
      const float dl = pre_box_dist2(qx, qy, qz, n0, n1), dr = pre_box_dist2(qx, qy, qz, n2, n3);
      const bool left_first = dl <= dr;
      const float dn = left_first ? dl : dr, df = left_first ? dr : dl;
      const uint2 cn = left_first ? make_uint2(__float_as_uint(n0.w), __float_as_uint(n1.w)) : make_uint2(__float_as_uint(n2.w), __float_as_uint(n3.w));
      const uint2 cf = left_first ? make_uint2(__float_as_uint(n2.w), __float_as_uint(n3.w)) : make_uint2(__float_as_uint(n0.w), __float_as_uint(n1.w));
      if (df < L.worst) {
        my_desc[sp * kLinBlock] = cf;
        my_dist[sp * kLinBlock] = df;
        sp++;
      }
      if (!(dn < L.worst)) {
        expand = false;
        continue;
      }
      if (cn.y == 0u) {
        cur = cn.x;
        continue;
      }
      leaf = cn;
    } else {
      bool got = false;
      while (sp > 0) {
        sp--;
        if (my_dist[sp * kLinBlock] < L.worst) {  // worst_distance() > lower bound of the subtree (ann/kdtree.hpp:228 analogue)
          leaf = my_desc[sp * kLinBlock];
          got = true;
          break;
        }
      }
      if (!got) break;
      if (leaf.y == 0u) {
        cur = leaf.x;
        expand = true;
        continue;
      }
    }
    for (uint32_t j = 0; j < leaf.y; j++) {
      const float4 t = __ldg(&pts[leaf.x + j]);
      const float dx = t.x - qx, dy = t.y - qy, dz = t.z - qz;
      L.offer(dx * dx + dy * dy + dz * dz, leaf.x + j);
    }
    expand = false;
  }
}

// Eigenvector of the smallest eigenvalue of a symmetric 3x3 (closed-form roots of the characteristic polynomial,
// eigenvector from the best-conditioned cross product of two rows of A - lambda I): the only part of the
// decomposition the regularised covariance and the normal need (normal_estimation.hpp:17-24,40-45).
__device__ __forceinline__ void smallest_eigenvector(double a00, double a01, double a02, double a11, double a12, double a22, double v[3]) {
  // shift + scale for conditioning
  const double shift = (a00 + a11 + a22) / 3.0;
  double b00 = a00 - shift, b11 = a11 - shift, b22 = a22 - shift, b01 = a01, b02 = a02, b12 = a12;
  double scale = fmax(fmax(fabs(b00), fabs(b11)), fmax(fabs(b22), fmax(fabs(b01), fmax(fabs(b02), fabs(b12)))));
  if (!(scale > 0.0)) {
    v[0] = 1.0;
    v[1] = 0.0;
    v[2] = 0.0;
    return;
  }
  const double is = 1.0 / scale;
  b00 *= is; b11 *= is; b22 *= is; b01 *= is; b02 *= is; b12 *= is;
  // trace(B) = 0: characteristic polynomial  x^3 - c1' x - det = 0
  const double p2 = (b00 * b00 + b11 * b11 + b22 * b22 + 2.0 * (b01 * b01 + b02 * b02 + b12 * b12)) / 6.0;  // = (sum eig^2)/6
  const double det = b00 * (b11 * b22 - b12 * b12) - b01 * (b01 * b22 - b12 * b02) + b02 * (b01 * b12 - b11 * b02);
  const double p = sqrt(p2);
  double lam_min;
  if (!(p > 0.0)) {
    lam_min = 0.0;
  } else {
    double r = det / (2.0 * p * p * p);
    r = fmin(fmax(r, -1.0), 1.0);
    const double phi = acos(r) / 3.0;
    // eigenvalues 2p cos(phi + 2 pi k / 3); the smallest is k = 1
    lam_min = 2.0 * p * cos(phi + 2.0943951023931953);
  }
  // rows of C = B - lam_min I ; kernel = cross product of the two most independent rows
  const double c00 = b00 - lam_min, c11 = b11 - lam_min, c22 = b22 - lam_min;
  const double r0[3] = {c00, b01, b02}, r1[3] = {b01, c11, b12}, r2[3] = {b02, b12, c22};
  double x01[3] = {r0[1] * r1[2] - r0[2] * r1[1], r0[2] * r1[0] - r0[0] * r1[2], r0[0] * r1[1] - r0[1] * r1[0]};
  double x02[3] = {r0[1] * r2[2] - r0[2] * r2[1], r0[2] * r2[0] - r0[0] * r2[2], r0[0] * r2[1] - r0[1] * r2[0]};
  double x12[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
  const double n01 = x01[0] * x01[0] + x01[1] * x01[1] + x01[2] * x01[2];
  const double n02 = x02[0] * x02[0] + x02[1] * x02[1] + x02[2] * x02[2];
  const double n12 = x12[0] * x12[0] + x12[1] * x12[1] + x12[2] * x12[2];
  const double* best = x01;
  double nb = n01;
  if (n02 > nb) {
    best = x02;
    nb = n02;
  }
  if (n12 > nb) {
    best = x12;
    nb = n12;
  }
  if (!(nb > 1e-280)) {  // (numerically) isotropic: any direction is an eigenvector
    v[0] = 1.0;
    v[1] = 0.0;
    v[2] = 0.0;
    return;
  }
  const double inv = rsqrt(nb);
  v[0] = best[0] * inv;
  v[1] = best[1] * inv;
  v[2] = best[2] * inv;
}

/// One thread per point of the cloud in LEAF order (neighbouring threads query neighbouring points).
/// mode bit 0: normals, bit 1: covariances.  Outputs are written in the cloud's ORIGINAL order
/// (index carried in pts[].w): device layout (float4 streams) and / or the reference's double layout.
///
/// Neighbour sets.  The search runs on the centred FP32 coordinates (rounded by up to ~2 um at 50 m), the reference on doubles: at the
/// k-th neighbour a near-tie could swap one point of the set -- rare (0.5 % of the points), but a different covariance where it happens.
/// So the search collects k + kFeatExtra candidates, and the k nearest of them are picked by their distance on the EXACT coordinates
/// (hi + lo, lo = what the FP32 rounding dropped, kept in original order): the set is then the reference's unless two candidates are
/// equidistant to ~1e-12 m^2, and the covariance sums use the exact coordinates as well.  lo == nullptr: FP32 coordinates throughout.
constexpr int kFeatExtra = 4;
template <int KMAX>
__global__ void __launch_bounds__(kLinBlock) features_kernel(const float4* __restrict__ pnodes, const float4* __restrict__ pts, uint32_t n, int k,
                                                             const double* __restrict__ centre, int mode, float4* out_normals, float4* out_covA,
                                                             float4* out_covB, double* out_normals_d, double* out_covs_d, int leaf_order_out, int depth,
                                                             const float4* __restrict__ lo) {
  extern __shared__ uint2 s_stack[];  // [depth][kLinBlock] descriptors, then [depth][kLinBlock] distances
  constexpr int KC = KMAX + kFeatExtra;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 q = __ldg(&pts[i]);
  KnnList<KC> L;
  const int kc = lo ? k + kFeatExtra : k;
  L.init(kc);
  bvh_knn<KC>(pnodes, pts, q.x, q.y, q.z, L, s_stack, reinterpret_cast<float*>(s_stack + static_cast<size_t>(depth) * kLinBlock));
  const uint32_t qo = static_cast<uint32_t>(__float_as_int(q.w));
  double qx = q.x, qy = q.y, qz = q.z;
  if (lo) {
    const float4 l = __ldg(&lo[qo]);
    qx += l.x; qy += l.y; qz += l.z;
  }
  // exact offset of candidate j from the query (hi + lo of both)
  auto offset = [&](int j, double& x, double& y, double& z) {
    const float4 t = __ldg(&pts[L.i[j]]);
    double tx = t.x, ty = t.y, tz = t.z;
    if (lo) {
      const float4 l = __ldg(&lo[static_cast<uint32_t>(__float_as_int(t.w))]);
      tx += l.x; ty += l.y; tz += l.z;
    }
    x = tx - qx; y = ty - qy; z = tz - qz;
  };
  // exact squared distances of the candidates (-1 = no candidate / dropped)
  double d2[KC];
  int found = 0;
#pragma unroll
  for (int j = 0; j < KC; j++) {
    d2[j] = -1.0;
    if (j < kc && L.i[j] != kNone) {
      double x, y, z;
      offset(j, x, y, z);
      d2[j] = x * x + y * y + z * z;
      found++;
    }
  }
  // drop the (found - k) farthest candidates (exact distances; among exact ties the later arrival goes, knn_result.hpp:88-101)
  for (int drop = found - k; drop > 0; drop--) {
    double worst = -1.0;
    int wj = -1;
#pragma unroll
    for (int j = 0; j < KC; j++)
      if (d2[j] >= worst && d2[j] >= 0.0) {
        worst = d2[j];
        wj = j;
      }
#pragma unroll
    for (int j = 0; j < KC; j++)
      if (j == wj) d2[j] = -1.0;
    found--;
  }
  // sums relative to the query point (covariance is translation invariant; keeps the FP64 sums well conditioned); the offsets are
  // re-derived rather than kept: 3 x KC doubles of registers would spill
  double s[3] = {0, 0, 0}, ss[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int j = 0; j < KC; j++) {
    if (d2[j] >= 0.0) {
      double x, y, z;
      offset(j, x, y, z);
      s[0] += x; s[1] += y; s[2] += z;
      ss[0] += x * x; ss[1] += x * y; ss[2] += x * z; ss[3] += y * y; ss[4] += y * z; ss[5] += z * z;
    }
  }
  const uint32_t orig = leaf_order_out ? i : qo;
  double nrm[3] = {0.0, 0.0, 0.0};
  double cov[6] = {1.0, 0.0, 0.0, 1.0, 0.0, 1.0};  // < 5 neighbours: identity covariance, zero normal (normal_estimation.hpp:33-37,71-75)
  if (found >= 5) {
    const double inv_n = 1.0 / found;
    const double mx = s[0] * inv_n, my = s[1] * inv_n, mz = s[2] * inv_n;
    // cov = (sum_cross - mean * sum^T) / n   (normal_estimation.hpp:85-86)
    const double c00 = (ss[0] - mx * s[0]) * inv_n, c01 = (ss[1] - mx * s[1]) * inv_n, c02 = (ss[2] - mx * s[2]) * inv_n;
    const double c11 = (ss[3] - my * s[1]) * inv_n, c12 = (ss[4] - my * s[2]) * inv_n, c22 = (ss[5] - mz * s[2]) * inv_n;
    double v[3];
    smallest_eigenvector(c00, c01, c02, c11, c12, c22, v);
    // V diag(1e-3, 1, 1) V^T = I - (1 - 1e-3) v v^T for orthonormal V
    const double w = 1.0 - 1e-3;
    cov[0] = 1.0 - w * v[0] * v[0]; cov[1] = -w * v[0] * v[1]; cov[2] = -w * v[0] * v[2];
    cov[3] = 1.0 - w * v[1] * v[1]; cov[4] = -w * v[1] * v[2]; cov[5] = 1.0 - w * v[2] * v[2];
    // flip the normal toward the origin of the cloud's own frame (normal_estimation.hpp:19-23)
    const double px = qx + centre[0], py = qy + centre[1], pz = qz + centre[2];
    const double sgn = (px * v[0] + py * v[1] + pz * v[2]) > 0.0 ? -1.0 : 1.0;
    nrm[0] = sgn * v[0]; nrm[1] = sgn * v[1]; nrm[2] = sgn * v[2];
  }
  if (mode & 1) {
    if (out_normals) out_normals[orig] = make_float4((float)nrm[0], (float)nrm[1], (float)nrm[2], 0.f);
    if (out_normals_d) {
      double* o = out_normals_d + static_cast<size_t>(orig) * 4;
      o[0] = nrm[0]; o[1] = nrm[1]; o[2] = nrm[2]; o[3] = 0.0;
    }
  }
  if (mode & 2) {
    if (out_covA) {
      out_covA[orig] = make_float4((float)cov[0], (float)cov[1], (float)cov[2], (float)cov[3]);
      out_covB[orig] = make_float4((float)cov[4], (float)cov[5], 0.f, 0.f);
    }
    if (out_covs_d) {
      double* o = out_covs_d + static_cast<size_t>(orig) * 16;
      o[0] = cov[0]; o[1] = cov[1]; o[2] = cov[2]; o[3] = 0.0;
      o[4] = cov[1]; o[5] = cov[3]; o[6] = cov[4]; o[7] = 0.0;
      o[8] = cov[2]; o[9] = cov[4]; o[10] = cov[5]; o[11] = 0.0;
      o[12] = 0.0; o[13] = 0.0; o[14] = 0.0; o[15] = 0.0;
    }
  }
}

template <int KMAX>
static cudaError_t launch_features_t(const float4* nodes, const float4* pts, uint32_t n, int k, const double* centre, int mode, float4* on, float4* oa, float4* ob,
                                     double* ond, double* ocd, int depth, int leaf_order_out, cudaStream_t st, const float4* lo) {
  if (depth < 1) depth = 1;
  const size_t smem = static_cast<size_t>(depth) * kLinBlock * (sizeof(uint2) + sizeof(float));
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(features_kernel<KMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  const int grid = static_cast<int>((n + kLinBlock - 1) / kLinBlock);
  features_kernel<KMAX><<<grid, kLinBlock, smem, st>>>(nodes, pts, n, k, centre, mode, on, oa, ob, ond, ocd, leaf_order_out, depth, lo);
  return cudaGetLastError();
}

cudaError_t launch_features(const float4* nodes, const float4* pts, uint32_t n, int k, const double* centre, int mode, float4* out_normals, float4* out_covA,
                            float4* out_covB, double* out_normals_d, double* out_covs_d, int depth, int leaf_order_out, cudaStream_t st, const float4* lo_orig) {
  if (n == 0) return cudaSuccess;
  if (k <= 10) return launch_features_t<10>(nodes, pts, n, k, centre, mode, out_normals, out_covA, out_covB, out_normals_d, out_covs_d, depth, leaf_order_out, st, lo_orig);
  if (k <= 20) return launch_features_t<20>(nodes, pts, n, k, centre, mode, out_normals, out_covA, out_covB, out_normals_d, out_covs_d, depth, leaf_order_out, st, lo_orig);
  if (k <= 32) return launch_features_t<32>(nodes, pts, n, k, centre, mode, out_normals, out_covA, out_covB, out_normals_d, out_covs_d, depth, leaf_order_out, st, lo_orig);
  return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------------------------
// k-NN of arbitrary query points in the target's search structure (KdTree::knn_search, ann/kdtree.hpp:165-189;
// batch_knn_search / batch_nearest_neighbor_search of the Python binding): one query per thread, results ascending.
// The search runs on the FP32 centred coordinates; the reported squared distances are recomputed in FP64.
// ---------------------------------------------------------------------------------------------
template <int KMAX>
__global__ void __launch_bounds__(kLinBlock) batch_knn_kernel(const float4* __restrict__ pnodes, const float4* __restrict__ pts, const double4* __restrict__ queries,
                                                              uint32_t n, int k, const double* __restrict__ centre, unsigned long long* out_idx, double* out_d,
                                                              int depth) {
  extern __shared__ uint2 s_stack[];  // [depth][kLinBlock] descriptors, then [depth][kLinBlock] distances
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 q = queries[i];
  const double qx = q.x - centre[0], qy = q.y - centre[1], qz = q.z - centre[2];
  KnnList<KMAX> L;
  L.init(k);
  bvh_knn<KMAX>(pnodes, pts, static_cast<float>(qx), static_cast<float>(qy), static_cast<float>(qz), L, s_stack,
                reinterpret_cast<float*>(s_stack + static_cast<size_t>(depth) * kLinBlock));
#pragma unroll
  for (int j = 0; j < KMAX; j++) {
    if (j < k) {
      unsigned long long idx = ~0ull;  // fewer than k points: size_t(-1) / max() like KnnResult's initial state (knn_result.hpp:60-66)
      double d = DBL_MAX;
      if (L.i[j] != kNone) {
        const float4 t = __ldg(&pts[L.i[j]]);
        const double dx = static_cast<double>(t.x) - qx, dy = static_cast<double>(t.y) - qy, dz = static_cast<double>(t.z) - qz;
        idx = static_cast<uint32_t>(__float_as_int(t.w));
        d = dx * dx + dy * dy + dz * dz;
      }
      out_idx[static_cast<size_t>(i) * k + j] = idx;
      out_d[static_cast<size_t>(i) * k + j] = d;
    }
  }
}

template <int KMAX>
static cudaError_t launch_batch_knn_t(const float4* pnodes, const float4* pts, const double* queries4, uint32_t n, int k, const double* centre,
                                      unsigned long long* out_idx, double* out_d, int depth, cudaStream_t st) {
  if (depth < 1) depth = 1;
  const size_t smem = static_cast<size_t>(depth) * kLinBlock * (sizeof(uint2) + sizeof(float));
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(batch_knn_kernel<KMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  batch_knn_kernel<KMAX><<<(n + kLinBlock - 1) / kLinBlock, kLinBlock, smem, st>>>(pnodes, pts, reinterpret_cast<const double4*>(queries4), n, k, centre, out_idx,
                                                                                  out_d, depth);
  return cudaGetLastError();
}

cudaError_t launch_batch_knn(const float4* pnodes, const float4* pts, const double* queries4, uint32_t n, int k, const double* centre, unsigned long long* out_idx,
                             double* out_d, int depth, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  if (k <= 1) return launch_batch_knn_t<1>(pnodes, pts, queries4, n, k, centre, out_idx, out_d, depth, st);
  if (k <= 10) return launch_batch_knn_t<10>(pnodes, pts, queries4, n, k, centre, out_idx, out_d, depth, st);
  if (k <= 20) return launch_batch_knn_t<20>(pnodes, pts, queries4, n, k, centre, out_idx, out_d, depth, st);
  if (k <= 32) return launch_batch_knn_t<32>(pnodes, pts, queries4, n, k, centre, out_idx, out_d, depth, st);
  return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------------------------
// voxel-grid down-sampling
// ---------------------------------------------------------------------------------------------
__global__ void voxel_keys_kernel(const double4* __restrict__ pts, size_t n, double inv_leaf, uint64_t* keys, uint32_t* vals) {
  constexpr int coord_bit_size = 21;
  constexpr int64_t coord_bit_mask = (1 << 21) - 1;
  constexpr int coord_offset = 1 << (coord_bit_size - 1);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const double4 p = pts[i];
    const int64_t cx = static_cast<int64_t>(floor(p.x * inv_leaf)) + coord_offset;
    const int64_t cy = static_cast<int64_t>(floor(p.y * inv_leaf)) + coord_offset;
    const int64_t cz = static_cast<int64_t>(floor(p.z * inv_leaf)) + coord_offset;
    uint64_t key = ~0ull;  // out of the 21-bit range: dropped (downsampling.hpp:41-45 warns and marks them invalid)
    if (cx >= 0 && cy >= 0 && cz >= 0 && cx <= coord_bit_mask && cy <= coord_bit_mask && cz <= coord_bit_mask)
      key = static_cast<uint64_t>(cx) | (static_cast<uint64_t>(cy) << coord_bit_size) | (static_cast<uint64_t>(cz) << (2 * coord_bit_size));
    keys[i] = key;
    vals[i] = static_cast<uint32_t>(i);
  }
}

__global__ void voxel_heads_kernel(const uint64_t* __restrict__ keys, size_t n, uint32_t* heads) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint64_t k = keys[i];
    heads[i] = (k != ~0ull && (i == 0 || keys[i - 1] != k)) ? 1u : 0u;
  }
}

// one thread per segment head: mean of the segment's points in sorted (= original index) order, FP64
__global__ void voxel_means_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, const uint32_t* __restrict__ heads,
                                   const uint32_t* __restrict__ slots, size_t n, const double4* __restrict__ pts, double4* out) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    if (!heads[i]) continue;
    const uint64_t k = keys[i];
    double sx = 0, sy = 0, sz = 0, cnt = 0;
    for (size_t j = i; j < n && keys[j] == k; j++) {
      const double4 p = pts[vals[j]];
      sx += p.x; sy += p.y; sz += p.z; cnt += 1.0;
    }
    out[slots[i]] = make_double4(sx / cnt, sy / cnt, sz / cnt, 1.0);
  }
}

static int vgrid(size_t n, int cap) {
  size_t g = (n + 255) / 256;
  if (g < 1) g = 1;
  if (g > static_cast<size_t>(cap)) g = cap;
  return static_cast<int>(g);
}

cudaError_t launch_voxel_keys(const double* d_pts4, size_t n, double inv_leaf, uint64_t* keys, uint32_t* vals, int sm_count, cudaStream_t st) {
  if (!n) return cudaSuccess;
  voxel_keys_kernel<<<vgrid(n, sm_count * 8), 256, 0, st>>>(reinterpret_cast<const double4*>(d_pts4), n, inv_leaf, keys, vals);
  return cudaGetLastError();
}
cudaError_t launch_voxel_heads(const uint64_t* keys, size_t n, uint32_t* heads, int sm_count, cudaStream_t st) {
  if (!n) return cudaSuccess;
  voxel_heads_kernel<<<vgrid(n, sm_count * 8), 256, 0, st>>>(keys, n, heads);
  return cudaGetLastError();
}
cudaError_t launch_voxel_means(const uint64_t* keys, const uint32_t* vals, const uint32_t* heads, const uint32_t* slots, size_t n, const double* d_pts4,
                               double* d_out4, int sm_count, cudaStream_t st) {
  if (!n) return cudaSuccess;
  voxel_means_kernel<<<vgrid(n, sm_count * 8), 256, 0, st>>>(keys, vals, heads, slots, n, reinterpret_cast<const double4*>(d_pts4),
                                                             reinterpret_cast<double4*>(d_out4));
  return cudaGetLastError();
}
// ---------------------------------------------------------------------------------------------
// Gaussian voxel map construction (incremental_voxelmap.hpp:55-92 with GaussianVoxel::add / finalize,
// gaussian_voxelmap.hpp:30-62): one thread per voxel adds its points and covariances in original-index order
// (the order the reference's insert loop meets them) and divides by the count.
// ---------------------------------------------------------------------------------------------
__global__ void voxel_stats_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, const uint32_t* __restrict__ heads,
                                   const uint32_t* __restrict__ slots, size_t n, const double4* __restrict__ pts, const double* __restrict__ covs,
                                   double4* out_means, double* out_covs, int4* out_coords) {
  constexpr int coord_bit_size = 21;
  constexpr uint64_t coord_bit_mask = (1ull << 21) - 1;
  constexpr int coord_offset = 1 << (coord_bit_size - 1);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    if (!heads[i]) continue;
    const uint64_t k = keys[i];
    double sx = 0, sy = 0, sz = 0, cnt = 0;
    double c[16];
#pragma unroll
    for (int a = 0; a < 16; a++) c[a] = 0.0;
    for (size_t j = i; j < n && keys[j] == k; j++) {
      const uint32_t v = vals[j];
      const double4 p = pts[v];
      sx += p.x;
      sy += p.y;
      sz += p.z;
      cnt += 1.0;
      if (covs) {
        const double* cv = covs + static_cast<size_t>(v) * 16;
#pragma unroll
        for (int a = 0; a < 16; a++) c[a] += cv[a];
      }
    }
    const uint32_t id = slots[i];
    out_means[id] = make_double4(sx / cnt, sy / cnt, sz / cnt, 1.0);
    if (covs) {
#pragma unroll
      for (int a = 0; a < 16; a++) out_covs[static_cast<size_t>(id) * 16 + a] = c[a] / cnt;
    }
    out_coords[id] = make_int4(static_cast<int>(k & coord_bit_mask) - coord_offset, static_cast<int>((k >> coord_bit_size) & coord_bit_mask) - coord_offset,
                               static_cast<int>((k >> (2 * coord_bit_size)) & coord_bit_mask) - coord_offset, static_cast<int>(id));
  }
}

__global__ void vox_table_clear_kernel(int4* table, uint32_t capacity) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < capacity) table[i] = make_int4(0, 0, 0, -1);
}

// voxel coordinates are distinct, so an insert only has to find a free slot of its probe sequence
__global__ void vox_table_insert_kernel(const int4* __restrict__ coords, uint32_t n_voxels, int4* table, uint32_t mask) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_voxels) return;
  const int4 c = coords[i];
  uint32_t slot = vox_hash(c.x, c.y, c.z) & mask;
  for (;;) {
    if (atomicCAS(&table[slot].w, -1, c.w) == -1) {
      table[slot].x = c.x;
      table[slot].y = c.y;
      table[slot].z = c.z;
      return;
    }
    slot = (slot + 1u) & mask;
  }
}

cudaError_t launch_voxel_stats(const uint64_t* keys, const uint32_t* vals, const uint32_t* heads, const uint32_t* slots, size_t n, const double* d_pts4,
                               const double* d_covs16, double* out_means4, double* out_covs16, int4* out_coords, int sm_count, cudaStream_t st) {
  if (!n) return cudaSuccess;
  voxel_stats_kernel<<<vgrid(n, sm_count * 8), 256, 0, st>>>(keys, vals, heads, slots, n, reinterpret_cast<const double4*>(d_pts4), d_covs16,
                                                             reinterpret_cast<double4*>(out_means4), out_covs16, out_coords);
  return cudaGetLastError();
}
cudaError_t launch_vox_table_build(const int4* coords, uint32_t n_voxels, int4* table, uint32_t capacity, cudaStream_t st) {
  vox_table_clear_kernel<<<(capacity + 255u) / 256u, 256, 0, st>>>(table, capacity);
  if (n_voxels) vox_table_insert_kernel<<<(n_voxels + 255u) / 256u, 256, 0, st>>>(coords, n_voxels, table, capacity - 1u);
  return cudaGetLastError();
}

cudaError_t exclusive_sum_u32(void* d_temp, size_t& temp_bytes, const uint32_t* in, uint32_t* out, size_t n, cudaStream_t st) {
  return cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, in, out, static_cast<int>(n), st);
}

}  // namespace sgb
