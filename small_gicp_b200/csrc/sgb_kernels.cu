// SPDX-License-Identifier: MIT
// Hot-path kernels (sm_100a): fused transform -> kd-tree NN -> reject -> factor -> reduction, the
// error() kernel for the LM inner loop, and the one-off layout kernels (upload conversion, Morton
// ordering of the source, leaf ordering of the target).  See DESIGN.md §4 for the roofline of each.
#include "sgb_device.cuh"
#include "sgb_kernels.h"

#include <cfloat>
#include <cub/device/device_radix_sort.cuh>

namespace sgb {

// =============================================================================================
// linearize: one thread per source point (grid-stride, Morton-ordered so a warp's queries share
// their path through the tree), FP32 search on FP32 coordinates, FP64 factor algebra and sums.
// =============================================================================================
#ifdef SGB_PROFILING  // single fused kernel (SGB_SEARCH=0): the round-1 starting point, kept for A/B runs only
template <int FACTOR, int ROBUST>
__global__ void __launch_bounds__(kLinBlock) linearize_kd_kernel(const __grid_constant__ LinParams P) {
  extern __shared__ uint2 s_stack[];
  double acc[kAcc + 1];
#pragma unroll
  for (int k = 0; k <= kAcc; k++) acc[k] = 0.0;

  // pose in the centred frames:  p' - c_t = R p_f + (R c_s + t - c_t)
  const double* R = P.T;
  const double csx = P.src.centre[0], csy = P.src.centre[1], csz = P.src.centre[2];
  const double tpx = R[0] * csx + R[1] * csy + R[2] * csz + P.T[9] - P.tgt.centre[0];
  const double tpy = R[3] * csx + R[4] * csy + R[5] * csz + P.T[10] - P.tgt.centre[1];
  const double tpz = R[6] * csx + R[7] * csy + R[8] * csz + P.T[11] - P.tgt.centre[2];

  // Work unit = chunk of 32*K source positions per warp; lane l handles positions base + k*32 + l, k = 0..K-1,
  // which are K consecutive points along the Morton curve: the previous result of the lane is a tight search
  // seed for the next query, and so is the previous linearize's correspondence of the same point.
  const uint32_t K = P.src.run, chunk_pts = 32u * K;
  const uint32_t n_chunks = (P.src.n + chunk_pts - 1) / chunk_pts;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t chunk = warp; chunk < n_chunks; chunk += n_warps) {
    uint32_t chain = kNone;
    for (uint32_t k = 0; k < K; k++) {
      const uint32_t i = chunk * chunk_pts + k * 32u + lane;
      if (i >= P.src.n) break;
      const float4 sp = __ldg(&P.src.pts[i]);
      const double sx = sp.x, sy = sp.y, sz = sp.z;
      const double qx = R[0] * sx + R[1] * sy + R[2] * sz + tpx;
      const double qy = R[3] * sx + R[4] * sy + R[5] * sz + tpy;
      const double qz = R[6] * sx + R[7] * sy + R[8] * sz + tpz;
      const float fx = static_cast<float>(qx), fy = static_cast<float>(qy), fz = static_cast<float>(qz);

      float best_d = P.max_dist_sq;
      uint32_t best = kNone;
      {  // seeds
        const uint32_t prev = P.use_prev ? P.corr[i] : kNone;
        if (prev != kNone) {
          const float4 t = __ldg(&P.tgt.pts[prev]);
          const float dx = t.x - fx, dy = t.y - fy, dz = t.z - fz;
          const float d = dx * dx + dy * dy + dz * dz;
          if (d < best_d) {
            best_d = d;
            best = prev;
          }
        }
        if (chain != kNone && chain != prev) {
          const float4 t = __ldg(&P.tgt.pts[chain]);
          const float dx = t.x - fx, dy = t.y - fy, dz = t.z - fz;
          const float d = dx * dx + dy * dy + dz * dz;
          if (d < best_d) {
            best_d = d;
            best = chain;
          }
        }
      }
      best = kd_nearest(P.tgt.nodes, P.tgt.pts, fx, fy, fz, best_d, best, s_stack);
      if (best != kNone) chain = best;

      double rx = 0.0, ry = 0.0, rz = 0.0;
      if (best != kNone) {
        const float4 tq = __ldg(&P.tgt.pts[best]);
        rx = static_cast<double>(tq.x) - qx;
        ry = static_cast<double>(tq.y) - qy;
        rz = static_cast<double>(tq.z) - qz;
        if (rx * rx + ry * ry + rz * rz > P.max_dist_sq_d) best = kNone;  // DistanceRejector on the FP64 residual (rejector.hpp:24)
      }
      P.corr[i] = best;
      if (best == kNone) continue;

      Sym3 M;
      if (FACTOR == 0) {
        M = Sym3{1.0, 0.0, 0.0, 1.0, 0.0, 1.0};
      } else if (FACTOR == 1) {
        const float4 n = __ldg(&P.tgt.normals[best]);
        M = Sym3{static_cast<double>(n.x) * n.x, 0.0, 0.0, static_cast<double>(n.y) * n.y, 0.0, static_cast<double>(n.z) * n.z};
      } else {
        M = gicp_precision(R, __ldg(&P.src.covA[i]), __ldg(&P.src.covB[i]), __ldg(&P.tgt.covA[best]), __ldg(&P.tgt.covB[best]));
      }
      accumulate_factor<ROBUST>(R, M, rx, ry, rz, csx + sx, csy + sy, csz + sz, P.robust_c, acc);
      acc[kAcc] += 1.0;
    }
  }
  block_reduce_and_finish<kAcc + 1, true>(acc, P.partials, P.ticket, P.out, P.comm);
}

#endif  // SGB_PROFILING

// Gaussian-voxel-map target (VGICP): hash probe of 1 / 7 / 27 voxels in the reference's offset order
// (incremental_voxelmap.hpp:157-186), nearest voxel mean wins, first-found wins ties (knn_result.hpp:81-83).
__device__ __forceinline__ void vox_offset(int num, int o, int& dx, int& dy, int& dz) {
  if (num == 27) {  // for i, j, k in -1..1 (lexicographic)
    dx = o / 9 - 1;
    dy = (o / 3) % 3 - 1;
    dz = o % 3 - 1;
  } else {  // 1 or 7: centre, +x, +y, +z, -x, -y, -z
    const int s = o == 0 ? 0 : (o <= 3 ? 1 : -1);
    const int a = o == 0 ? -1 : (o - 1) % 3;
    dx = a == 0 ? s : 0;
    dy = a == 1 ? s : 0;
    dz = a == 2 ? s : 0;
  }
}

template <int FACTOR, int ROBUST>
__global__ void __launch_bounds__(kLinBlock) linearize_vox_kernel(const __grid_constant__ LinParams P) {
  double acc[kAcc + 1];
#pragma unroll
  for (int k = 0; k <= kAcc; k++) acc[k] = 0.0;
  const double* R = P.T;
  const double csx = P.src.centre[0], csy = P.src.centre[1], csz = P.src.centre[2];
  const double ctx = P.tgt.centre[0], cty = P.tgt.centre[1], ctz = P.tgt.centre[2];
  const double tpx = R[0] * csx + R[1] * csy + R[2] * csz + P.T[9] - ctx;
  const double tpy = R[3] * csx + R[4] * csy + R[5] * csz + P.T[10] - cty;
  const double tpz = R[6] * csx + R[7] * csy + R[8] * csz + P.T[11] - ctz;
  const double max_d = P.max_dist_sq_d;

  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P.src.n; i += stride) {
    const float4 sp = __ldg(&P.src.pts[i]);
    const double sx = sp.x, sy = sp.y, sz = sp.z;
    const double qx = R[0] * sx + R[1] * sy + R[2] * sz + tpx;
    const double qy = R[3] * sx + R[4] * sy + R[5] * sz + tpy;
    const double qz = R[6] * sx + R[7] * sy + R[8] * sz + tpz;
    // voxel coordinate of the un-centred point, fast_floor semantics == floor (util/fast_floor.hpp:12-15)
    const int cx = static_cast<int>(floor((qx + ctx) * P.tgt.vox_inv_leaf));
    const int cy = static_cast<int>(floor((qy + cty) * P.tgt.vox_inv_leaf));
    const int cz = static_cast<int>(floor((qz + ctz) * P.tgt.vox_inv_leaf));
    double best_d = DBL_MAX;
    uint32_t best = kNone;
    double brx = 0, bry = 0, brz = 0;
    for (int o = 0; o < P.tgt.vox_num_offsets; o++) {
      int ox, oy, oz;
      vox_offset(P.tgt.vox_num_offsets, o, ox, oy, oz);
      const int vx = cx + ox, vy = cy + oy, vz = cz + oz;
      uint32_t slot = vox_hash(vx, vy, vz) & P.tgt.vox_mask;
      for (;;) {
        const int4 e = __ldg(&P.tgt.vox_table[slot]);
        if (e.w < 0) break;
        if (e.x == vx && e.y == vy && e.z == vz) {
          const float4 m = __ldg(&P.tgt.pts[e.w]);
          const double dx = static_cast<double>(m.x) - qx, dy = static_cast<double>(m.y) - qy, dz = static_cast<double>(m.z) - qz;
          const double d = dx * dx + dy * dy + dz * dz;
          if (d < best_d) {
            best_d = d;
            best = static_cast<uint32_t>(e.w);
            brx = dx;
            bry = dy;
            brz = dz;
          }
          break;
        }
        slot = (slot + 1u) & P.tgt.vox_mask;
      }
    }
    if (best != kNone && best_d > max_d) best = kNone;  // DistanceRejector, rejector.hpp:23-25
    P.corr[i] = best;
    if (best == kNone) continue;
    Sym3 M;
    if (FACTOR == 2) {
      M = gicp_precision(R, __ldg(&P.src.covA[i]), __ldg(&P.src.covB[i]), __ldg(&P.tgt.covA[best]), __ldg(&P.tgt.covB[best]));
    } else {
      M = Sym3{1.0, 0.0, 0.0, 1.0, 0.0, 1.0};
    }
    accumulate_factor<ROBUST>(R, M, brx, bry, brz, csx + sx, csy + sy, csz + sz, P.robust_c, acc);
    acc[kAcc] += 1.0;
  }
  block_reduce_and_finish<kAcc + 1, true>(acc, P.partials, P.ticket, P.out, P.comm);
}

// =============================================================================================
// error(): cached correspondences, trial pose T, GICP precision matrix re-derived from the pose of
// the last linearize (gicp_factor.hpp:81-89).  Pure streaming + gather.
// =============================================================================================
template <int FACTOR, int ROBUST>
__global__ void __launch_bounds__(kLinBlock) error_kernel(const __grid_constant__ LinParams P) {
  double acc[1] = {0.0};
  const double* R = P.T;
  const double csx = P.src.centre[0], csy = P.src.centre[1], csz = P.src.centre[2];
  const double tpx = R[0] * csx + R[1] * csy + R[2] * csz + P.T[9] - P.tgt.centre[0];
  const double tpy = R[3] * csx + R[4] * csy + R[5] * csz + P.T[10] - P.tgt.centre[1];
  const double tpz = R[6] * csx + R[7] * csy + R[8] * csz + P.T[11] - P.tgt.centre[2];
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P.src.n; i += stride) {
    const uint32_t k = __ldg(&P.corr[i]);
    if (k == kNone) continue;
    const float4 sp = __ldg(&P.src.pts[i]);
    const float4 tq = __ldg(&P.tgt.pts[k]);
    float4 sA = make_float4(0.f, 0.f, 0.f, 0.f), sB = sA, t1 = sA, t2 = sA;
    if (FACTOR == 1) t1 = __ldg(&P.tgt.normals[k]);
    if (FACTOR == 2) {
      sA = __ldg(&P.src.covA[i]);
      sB = __ldg(&P.src.covB[i]);
      t1 = __ldg(&P.tgt.covA[k]);
      t2 = __ldg(&P.tgt.covB[k]);
    }
    const double e = point_error<FACTOR, ROBUST>(R, tpx, tpy, tpz, P.Tlin, P.robust_c, sp, sA, sB, tq, t1, t2);
    acc[0] += e;
  }
  block_reduce_and_finish<1, false>(acc, P.partials, P.ticket, P.out, P.comm);
}

// ---------------------------------------------------------------------------------------------
template <int FACTOR, int ROBUST>
static cudaError_t launch_lin(const LinParams& P, bool voxel, int grid, size_t smem, cudaStream_t st) {
  if (voxel) {
    if (FACTOR == 1) return cudaErrorInvalidValue;
    linearize_vox_kernel<(FACTOR == 1 ? 0 : FACTOR), ROBUST><<<grid, kLinBlock, 0, st>>>(P);
  } else {
#ifdef SGB_PROFILING
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(linearize_kd_kernel<FACTOR, ROBUST>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return e;
    }
    linearize_kd_kernel<FACTOR, ROBUST><<<grid, kLinBlock, smem, st>>>(P);
#else
    return cudaErrorNotSupported;  // point targets go through the search + factor kernels (sgb_grid.cu, sgb_kernels_packet.cu, sgb_kernels_split.cu)
#endif
  }
  return cudaGetLastError();
}
template <int FACTOR, int ROBUST>
static cudaError_t launch_err(const LinParams& P, int grid, cudaStream_t st) {
  error_kernel<FACTOR, ROBUST><<<grid, kLinBlock, 0, st>>>(P);
  return cudaGetLastError();
}

#define SGB_DISPATCH(FN, ...)                                      \
  switch (factor * 3 + robust) {                                   \
    case 0: return FN<0, 0>(__VA_ARGS__);                          \
    case 1: return FN<0, 1>(__VA_ARGS__);                          \
    case 2: return FN<0, 2>(__VA_ARGS__);                          \
    case 3: return FN<1, 0>(__VA_ARGS__);                          \
    case 4: return FN<1, 1>(__VA_ARGS__);                          \
    case 5: return FN<1, 2>(__VA_ARGS__);                          \
    case 6: return FN<2, 0>(__VA_ARGS__);                          \
    case 7: return FN<2, 1>(__VA_ARGS__);                          \
    case 8: return FN<2, 2>(__VA_ARGS__);                          \
    default: return cudaErrorInvalidValue;                         \
  }

cudaError_t launch_linearize(const LinParams& P, int factor, int robust, bool voxel, int grid, int stack_depth, cudaStream_t st) {
  const size_t smem = voxel ? 0 : static_cast<size_t>(stack_depth) * kLinBlock * sizeof(uint2);
  SGB_DISPATCH(launch_lin, P, voxel, grid, smem, st);
}
cudaError_t launch_error(const LinParams& P, int factor, int robust, int grid, cudaStream_t st) { SGB_DISPATCH(launch_err, P, grid, st); }

// A rank whose shard is empty still has to take part in the fused exchange: one CTA that reduces nothing.
template <int NACC, bool EXPAND>
__global__ void __launch_bounds__(kLinBlock) reduce_nothing_kernel(const __grid_constant__ LinParams P) {
  double acc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; k++) acc[k] = 0.0;
  block_reduce_and_finish<NACC, EXPAND>(acc, P.partials, P.ticket, P.out, P.comm);
}
cudaError_t launch_reduce_nothing(const LinParams& P, bool linearize, cudaStream_t st) {
  if (linearize)
    reduce_nothing_kernel<kAcc + 1, true><<<1, kLinBlock, 0, st>>>(P);
  else
    reduce_nothing_kernel<1, false><<<1, kLinBlock, 0, st>>>(P);
  return cudaGetLastError();
}

int linearize_occupancy(int stack_depth) {
  int nb = 0;
#ifdef SGB_PROFILING
  if (stack_depth > 0) {
    const size_t smem = static_cast<size_t>(stack_depth) * kLinBlock * sizeof(uint2);
    if (smem > 48 * 1024) cudaFuncSetAttribute(linearize_kd_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, linearize_kd_kernel<2, 0>, kLinBlock, smem) != cudaSuccess) return 1;
    return nb > 0 ? nb : 1;
  }
#endif
  (void)stack_depth;  // resident CTAs of the voxel-map kernel (VGICP): its grid-stride loop is sized to one wave
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, linearize_vox_kernel<2, 0>, kLinBlock, 0) != cudaSuccess) return 1;
  return nb > 0 ? nb : 1;
}

// =============================================================================================
// One-off layout kernels
// =============================================================================================
// min/max of N x 4 doubles (xyz used) -> bounds[6] via ordered-int atomics on doubles
__device__ __forceinline__ void atomic_min_double(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if (__longlong_as_double(assumed) <= v) break;
    old = atomicCAS(a, assumed, static_cast<unsigned long long>(__double_as_longlong(v)));
  } while (assumed != old);
}
__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if (__longlong_as_double(assumed) >= v) break;
    old = atomicCAS(a, assumed, static_cast<unsigned long long>(__double_as_longlong(v)));
  } while (assumed != old);
}

__global__ void bounds_init_kernel(double* bounds) {
  if (threadIdx.x < 3) bounds[threadIdx.x] = DBL_MAX;
  if (threadIdx.x >= 3 && threadIdx.x < 6) bounds[threadIdx.x] = -DBL_MAX;
}

__global__ void bounds_kernel(const double4* __restrict__ pts, size_t n, double* bounds) {
  double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const double4 p = pts[i];
    lo[0] = fmin(lo[0], p.x); lo[1] = fmin(lo[1], p.y); lo[2] = fmin(lo[2], p.z);
    hi[0] = fmax(hi[0], p.x); hi[1] = fmax(hi[1], p.y); hi[2] = fmax(hi[2], p.z);
  }
#pragma unroll
  for (int d = 0; d < 3; d++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = fmin(lo[d], __shfl_down_sync(0xffffffffu, lo[d], o));
      hi[d] = fmax(hi[d], __shfl_down_sync(0xffffffffu, hi[d], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; d++) {
      atomic_min_double(&bounds[d], lo[d]);
      atomic_max_double(&bounds[3 + d], hi[d]);
    }
  }
}

// centre = box centre ; also emit the Morton scale
__global__ void centre_kernel(const double* bounds, double* centre) {
  if (threadIdx.x < 3) {
    const double lo = bounds[threadIdx.x], hi = bounds[3 + threadIdx.x];
    centre[threadIdx.x] = (lo <= hi) ? 0.5 * (lo + hi) : 0.0;
  }
  if (threadIdx.x == 3) {
    double ext = 0.0;
    for (int d = 0; d < 3; d++) ext = fmax(ext, bounds[3 + d] - bounds[d]);
    centre[3] = ext > 0.0 ? ext : 1.0;  // largest extent
  }
}

__device__ __forceinline__ float4 pack_covA(const double* c) { return make_float4((float)c[0], (float)c[1], (float)c[2], (float)c[5]); }
__device__ __forceinline__ float4 pack_covB(const double* c) { return make_float4((float)c[6], (float)c[10], 0.f, 0.f); }

__device__ int d_use_hilbert = 1;  // space-filling curve of the source order: 1 Hilbert, 0 Morton (profiling switch)
cudaError_t set_source_curve(int hilbert) { return cudaMemcpyToSymbol(d_use_hilbert, &hilbert, sizeof(int)); }

__device__ __forceinline__ uint64_t spread21(uint64_t x) {  // 21 bits -> every third bit
  x &= 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}


/// 63-bit space-filling-curve key of a centred point (h = half extent of the cloud's box, inv_ext = (2^21 - 1) / extent).
__device__ __forceinline__ uint64_t curve_key(double x, double y, double z, double h, double inv_ext) {
  const uint64_t kx = static_cast<uint64_t>(fmin(fmax((x + h) * inv_ext, 0.0), 2097151.0));
  const uint64_t ky = static_cast<uint64_t>(fmin(fmax((y + h) * inv_ext, 0.0), 2097151.0));
  const uint64_t kz = static_cast<uint64_t>(fmin(fmax((z + h) * inv_ext, 0.0), 2097151.0));
  if (!d_use_hilbert) return spread21(kx) | (spread21(ky) << 1) | (spread21(kz) << 2);  // Morton
  // Hilbert index (Skilling's transpose algorithm, 21 bits x 3): unlike the Z-curve it has no jumps, so 32 consecutive
  // points form a tighter patch
  uint32_t X[3] = {static_cast<uint32_t>(kx), static_cast<uint32_t>(ky), static_cast<uint32_t>(kz)};
  const uint32_t M = 1u << 20;
  for (uint32_t Q = M; Q > 1; Q >>= 1) {
    const uint32_t Pm = Q - 1;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      if (X[a] & Q) {
        X[0] ^= Pm;
      } else {
        const uint32_t t = (X[0] ^ X[a]) & Pm;
        X[0] ^= t;
        X[a] ^= t;
      }
    }
  }
  X[1] ^= X[0];
  X[2] ^= X[1];
  uint32_t t = 0;
  for (uint32_t Q = M; Q > 1; Q >>= 1)
    if (X[2] & Q) t ^= Q - 1;
  X[0] ^= t;
  X[1] ^= t;
  X[2] ^= t;
  return (spread21(X[0]) << 2) | (spread21(X[1]) << 1) | spread21(X[2]);
}

// raw (double AoS) -> centred float4 in ORIGINAL order (w = index), optional features, optional curve key
__global__ void convert_kernel(const double4* __restrict__ pts, const double4* __restrict__ normals, const double* __restrict__ covs, size_t n,
                               const double* __restrict__ centre, float4* out_pts, float4* out_normals, float4* out_covA, float4* out_covB,
                               uint64_t* keys, uint32_t* vals, float4* out_lo) {
  const double cx = centre[0], cy = centre[1], cz = centre[2];
  const double inv_ext = 2097151.0 / centre[3];
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const double4 p = pts[i];
    const double x = p.x - cx, y = p.y - cy, z = p.z - cz;
    out_pts[i] = make_float4((float)x, (float)y, (float)z, __int_as_float(static_cast<int>(i)));
    // what the rounding to FP32 dropped (hi + lo carries ~48 bits of x - c): the feature estimation ranks its neighbour candidates and
    // sums its covariances on the exact coordinates, so that its neighbour SETS are the reference's, not FP32 near-tie variants of them
    if (out_lo) out_lo[i] = make_float4((float)(x - (double)(float)x), (float)(y - (double)(float)y), (float)(z - (double)(float)z), 0.f);
    if (normals) {
      const double4 nn = normals[i];
      out_normals[i] = make_float4((float)nn.x, (float)nn.y, (float)nn.z, 0.f);
    }
    if (covs) {
      const double* c = covs + i * 16;
      out_covA[i] = pack_covA(c);
      out_covB[i] = pack_covB(c);
    }
    if (keys) {
      keys[i] = curve_key(x, y, z, 0.5 * centre[3], inv_ext);
      vals[i] = static_cast<uint32_t>(i);
    }
  }
}

// rank[perm[j]] = j: where the point with original index i ends up in the search order
__global__ void inverse_perm_kernel(const uint32_t* __restrict__ perm, size_t n, uint32_t* rank) {
  for (size_t j = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; j < n; j += static_cast<size_t>(gridDim.x) * blockDim.x) rank[perm[j]] = static_cast<uint32_t>(j);
}

// one CHUNK of raw 4x4 double covariances (original indices first .. first + n) -> packed floats written straight to their place in
// the search order (and, optionally, to the original-order copy): the chunk is converted while the next one is still on the PCIe bus
__global__ void convert_cov_scatter_kernel(const double* __restrict__ covs, size_t first, size_t n, const uint32_t* __restrict__ rank, float4* outA,
                                           float4* outB, float4* origA, float4* origB) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const double* c = covs + (first + i) * 16;
    const float4 a = pack_covA(c), b = pack_covB(c);
    const uint32_t r = rank[first + i];
    outA[r] = a;
    outB[r] = b;
    if (origA) {
      origA[first + i] = a;
      origB[first + i] = b;
    }
  }
}

// curve keys of points that are already on the device in centred FP32 form (target: device-side tree construction)
__global__ void curve_keys_kernel(const float4* __restrict__ pts, size_t n, const double* __restrict__ centre, uint64_t* keys, uint32_t* vals) {
  const double inv_ext = 2097151.0 / centre[3], h = 0.5 * centre[3];
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 p = pts[i];
    keys[i] = curve_key(p.x, p.y, p.z, h, inv_ext);
    vals[i] = static_cast<uint32_t>(i);
  }
}

// ---------------------------------------------------------------------------------------------
// Device-side construction of the search structure (replaces KdTreeBuilder::build_tree, ann/kdtree.hpp:74-131, and its
// OMP / TBB variants).  Exact nearest-neighbour results do not depend on how the hierarchy is formed, so instead of
// recursive median splits the tree is a linear BVH: points sorted along the Hilbert curve, leaves = runs of kLbvhLeaf
// consecutive points, hierarchy = the implicit perfect binary tree over P = 2^ceil(log2(#leaves)) leaf slots (node i has
// children 2i+1, 2i+2; missing leaves are empty boxes that no query ever enters).  Every step is a data-parallel kernel;
// the output is the same 64-byte packet record array the search kernel consumes.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void lbvh_store_child(float4* pnodes, uint32_t child_heap, const float lo[3], const float hi[3], uint32_t a, uint32_t b) {
  const uint32_t parent = (child_heap - 1u) >> 1, side = (child_heap & 1u) ? 0u : 1u;  // odd heap index = left child
  pnodes[parent * 4 + side * 2 + 0] = make_float4(lo[0], lo[1], lo[2], __uint_as_float(a));
  pnodes[parent * 4 + side * 2 + 1] = make_float4(hi[0], hi[1], hi[2], __uint_as_float(b));
}

/// first point of node j at a level with `count` nodes (balanced implicit tree over n points)
__device__ __forceinline__ uint32_t tree_bound(uint32_t j, uint32_t count, uint32_t n) {
  return static_cast<uint32_t>((static_cast<uint64_t>(j) * n) / count);
}
/// node (at a level with `count` nodes) that owns position i
__device__ __forceinline__ uint32_t tree_node_of(uint32_t i, uint32_t count, uint32_t n) {
  return static_cast<uint32_t>(((static_cast<uint64_t>(i) + 1u) * count + n - 1u) / n) - 1u;
}

// ---- median-split refinement (device-side kd construction): one pass per level ------------------------------------
// boxes: count x 6 floats stored as order-preserving uints (lo xyz, hi xyz)
__device__ __forceinline__ uint32_t float_flip(float f) {
  const uint32_t u = __float_as_uint(f);
  return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float float_unflip(uint32_t u) { return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xFFFFFFFFu)); }

__global__ void kd_boxes_init_kernel(uint32_t* boxes, uint32_t count) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < count * 6u) boxes[t] = (t % 6u) < 3u ? 0xFFFFFFFFu : 0u;
}

__global__ void kd_boxes_kernel(const float4* __restrict__ pts, uint32_t n, uint32_t count, uint32_t* boxes) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < n;
  const uint32_t idx = valid ? i : n - 1u;
  const float4 p = pts[idx];
  const uint32_t node = tree_node_of(idx, count, n);
  uint32_t v[6] = {float_flip(p.x), float_flip(p.y), float_flip(p.z), float_flip(p.x), float_flip(p.y), float_flip(p.z)};
  // pre-reduction: whole CTA in one node (upper levels: otherwise thousands of atomics hit the same six words), else
  // whole warp in one node, else per-thread atomics (deepest levels, little contention)
  __shared__ uint32_t s_box[6];
  __shared__ uint32_t s_first_node, s_last_node;
  if (threadIdx.x == 0) s_first_node = node;
  if (threadIdx.x == blockDim.x - 1) s_last_node = node;
  if (threadIdx.x < 6) s_box[threadIdx.x] = threadIdx.x < 3 ? 0xFFFFFFFFu : 0u;
  __syncthreads();
  const bool cta_uniform = s_first_node == s_last_node;  // positions are contiguous: equal ends => one node
  const uint32_t node0 = __shfl_sync(0xffffffffu, node, 0);
  if (__all_sync(0xffffffffu, node == node0)) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      v[a] = __reduce_min_sync(0xffffffffu, v[a]);
      v[3 + a] = __reduce_max_sync(0xffffffffu, v[3 + a]);
    }
    if ((threadIdx.x & 31u) == 0) {
      uint32_t* dst = cta_uniform ? s_box : &boxes[node * 6];
#pragma unroll
      for (int a = 0; a < 3; a++) {
        atomicMin(&dst[a], v[a]);
        atomicMax(&dst[3 + a], v[3 + a]);
      }
    }
    if (cta_uniform) {
      __syncthreads();
      if (threadIdx.x < 3) atomicMin(&boxes[node * 6 + threadIdx.x], s_box[threadIdx.x]);
      if (threadIdx.x >= 3 && threadIdx.x < 6) atomicMax(&boxes[node * 6 + threadIdx.x], s_box[threadIdx.x]);
    }
  } else if (valid) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      atomicMin(&boxes[node * 6 + a], v[a]);
      atomicMax(&boxes[node * 6 + 3 + a], v[3 + a]);
    }
  }
}

// key = (node << 32) | coordinate along the node's widest axis: sorting by it orders every node's points along its split
// axis, and the balanced position boundaries of the next level then ARE the median splits
__global__ void kd_keys_kernel(const float4* __restrict__ pts, uint32_t n, uint32_t count, const uint32_t* __restrict__ boxes, uint64_t* keys) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t node = tree_node_of(i, count, n);
  const float ex = float_unflip(boxes[node * 6 + 3]) - float_unflip(boxes[node * 6 + 0]);
  const float ey = float_unflip(boxes[node * 6 + 4]) - float_unflip(boxes[node * 6 + 1]);
  const float ez = float_unflip(boxes[node * 6 + 5]) - float_unflip(boxes[node * 6 + 2]);
  const float4 p = pts[i];
  const float c = ex >= ey ? (ex >= ez ? p.x : p.z) : (ey >= ez ? p.y : p.z);
  keys[i] = (static_cast<uint64_t>(node) << 32) | float_flip(c);
}

// The same median-split refinement for a whole SUBTREE of at most kKdSegMax points in ONE launch: CTA b owns node b of level `level0`
// (a contiguous position range), keeps its points in shared memory and runs every remaining level there -- per level: node boxes
// (shared-memory atomics on order-preserving integer images of the floats), key = (node, coordinate along the node's widest axis),
// bitonic sort of (key, slot), permutation of the points.  Above level0 a level is ~12 launches (boxes, keys, radix sort, gather); for a
// 16k-point frame that was 9 levels = ~110 launches = 0.9 ms of launch latency, now 3 levels + this kernel.  Same keys, same balanced
// position boundaries as the per-level path (ties may land differently -- any order is a valid tree).
constexpr uint32_t kKdSegMax = 2048;
constexpr uint32_t kKdSegNodes = 128;
__global__ void __launch_bounds__(1024) kd_refine_smem_kernel(float4* __restrict__ leaf_pts, uint32_t* __restrict__ perm, uint32_t n, uint32_t level0, uint32_t levels) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  float4* s_pts = reinterpret_cast<float4*>(s_raw);                                            // [2][kKdSegMax]
  unsigned long long* s_key = reinterpret_cast<unsigned long long*>(s_raw + 2 * kKdSegMax * sizeof(float4));  // [kKdSegMax]
  uint32_t* s_idx = reinterpret_cast<uint32_t*>(s_key + kKdSegMax);                            // [kKdSegMax]
  uint32_t* s_box = s_idx + kKdSegMax;                                                         // [kKdSegNodes][6]
  const uint32_t b = blockIdx.x, tid = threadIdx.x;
  const uint32_t first = tree_bound(b, 1u << level0, n), last = tree_bound(b + 1u, 1u << level0, n), m = last - first;
  for (uint32_t t = tid; t < m; t += blockDim.x) s_pts[t] = leaf_pts[first + t];
  uint32_t cur = 0;
  __syncthreads();
  for (uint32_t level = level0; level < levels; level++) {
    const uint32_t count = 1u << level, node_first = b << (level - level0), n_nodes = 1u << (level - level0);
    for (uint32_t t = tid; t < n_nodes * 6u; t += blockDim.x) s_box[t] = (t % 6u) < 3u ? 0xFFFFFFFFu : 0u;
    __syncthreads();
    const float4* P0 = s_pts + cur * kKdSegMax;
    for (uint32_t t = tid; t < m; t += blockDim.x) {
      const float4 p = P0[t];
      uint32_t* bx = s_box + (tree_node_of(first + t, count, n) - node_first) * 6u;
      atomicMin(&bx[0], float_flip(p.x));
      atomicMin(&bx[1], float_flip(p.y));
      atomicMin(&bx[2], float_flip(p.z));
      atomicMax(&bx[3], float_flip(p.x));
      atomicMax(&bx[4], float_flip(p.y));
      atomicMax(&bx[5], float_flip(p.z));
    }
    __syncthreads();
    for (uint32_t t = tid; t < kKdSegMax; t += blockDim.x) {
      unsigned long long key = ~0ull;  // padding sorts to the end
      if (t < m) {
        const uint32_t node = tree_node_of(first + t, count, n) - node_first;
        const uint32_t* bx = s_box + node * 6u;
        const float ex = float_unflip(bx[3]) - float_unflip(bx[0]), ey = float_unflip(bx[4]) - float_unflip(bx[1]), ez = float_unflip(bx[5]) - float_unflip(bx[2]);
        const float4 p = P0[t];
        const float c = ex >= ey ? (ex >= ez ? p.x : p.z) : (ey >= ez ? p.y : p.z);
        key = (static_cast<unsigned long long>(node) << 32) | float_flip(c);
      }
      s_key[t] = key;
      s_idx[t] = t;
    }
    __syncthreads();
    // bitonic sort of (key, slot) over kKdSegMax entries: one compare-exchange per thread and step
    for (uint32_t k = 2; k <= kKdSegMax; k <<= 1) {
      for (uint32_t j = k >> 1; j > 0; j >>= 1) {
        for (uint32_t t = tid; t < kKdSegMax / 2u; t += blockDim.x) {
          const uint32_t lo = ((t & ~(j - 1u)) << 1) | (t & (j - 1u)), hi = lo | j;
          const bool up = (lo & k) == 0u;
          const unsigned long long a = s_key[lo], c = s_key[hi];
          if ((a > c) == up) {
            s_key[lo] = c;
            s_key[hi] = a;
            const uint32_t ia = s_idx[lo];
            s_idx[lo] = s_idx[hi];
            s_idx[hi] = ia;
          }
        }
        __syncthreads();
      }
    }
    float4* P1 = s_pts + (cur ^ 1u) * kKdSegMax;
    for (uint32_t t = tid; t < m; t += blockDim.x) P1[t] = P0[s_idx[t]];
    cur ^= 1u;
    __syncthreads();
  }
  const float4* Pf = s_pts + cur * kKdSegMax;
  for (uint32_t t = tid; t < m; t += blockDim.x) {
    const float4 p = Pf[t];
    leaf_pts[first + t] = p;
    perm[first + t] = static_cast<uint32_t>(__float_as_int(p.w));  // w carries the original index through every permutation
  }
}

// one warp per leaf slot: bounding box of its (at most 32) consecutive points -> the parent's child record
__global__ void lbvh_leaf_kernel(const float4* __restrict__ leaf_pts, uint32_t n, uint32_t P, float4* pnodes) {
  const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
  if (slot >= P) return;
  const uint32_t first = tree_bound(slot, P, n), last = tree_bound(slot + 1u, P, n), idx = first + lane;
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  if (idx < last) {
    const float4 p = leaf_pts[idx];
    lo[0] = hi[0] = p.x;
    lo[1] = hi[1] = p.y;
    lo[2] = hi[2] = p.z;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
  }
  if (lane == 0) {
    const uint32_t cnt = last - first;
    lbvh_store_child(pnodes, (P - 1u) + slot, lo, hi, cnt ? first : 0u, cnt);  // cnt == 0: inverted box, never entered
  }
}

// one thread per inner node of a level (children records complete): union box -> the parent's child record
__global__ void lbvh_level_kernel(uint32_t level_first, uint32_t level_count, float4* pnodes) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= level_count) return;
  const uint32_t node = level_first + t;
  const float4 l0 = pnodes[node * 4 + 0], l1 = pnodes[node * 4 + 1], r0 = pnodes[node * 4 + 2], r1 = pnodes[node * 4 + 3];
  const float lo[3] = {fminf(l0.x, r0.x), fminf(l0.y, r0.y), fminf(l0.z, r0.z)};
  const float hi[3] = {fmaxf(l1.x, r1.x), fmaxf(l1.y, r1.y), fmaxf(l1.z, r1.z)};
  lbvh_store_child(pnodes, node, lo, hi, node, 0u);  // inner child: a = node index, b = 0
}

cudaError_t launch_curve_keys(const float4* pts, size_t n, const double* centre4, uint64_t* keys, uint32_t* vals, int sm_count, cudaStream_t st) {
  if (!n) return cudaSuccess;
  size_t g = (n + 255) / 256;
  if (g > static_cast<size_t>(sm_count) * 8) g = static_cast<size_t>(sm_count) * 8;
  curve_keys_kernel<<<static_cast<int>(g), 256, 0, st>>>(pts, n, centre4, keys, vals);
  return cudaGetLastError();
}

/// pnodes: (P - 1) * 4 float4 records, P = number of leaf slots (power of two >= 2).  Returns the number of launches in *launches.
cudaError_t launch_lbvh_build(const float4* leaf_pts, uint32_t n, uint32_t P, float4* pnodes, int* launches, cudaStream_t st) {
  const uint32_t warps_per_block = 8;
  lbvh_leaf_kernel<<<(P + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, st>>>(leaf_pts, n, P, pnodes);
  int count = 1;
  for (uint32_t level_count = P >> 1; level_count >= 2; level_count >>= 1) {  // deepest inner level first; the root (level_count 1) has no parent
    const uint32_t level_first = level_count - 1u;
    lbvh_level_kernel<<<(level_count + 255) / 256, 256, 0, st>>>(level_first, level_count, pnodes);
    count++;
  }
  if (launches) *launches = count;
  return cudaGetLastError();
}

/// One level of the median-split refinement: boxes of the `count` nodes of this level over the points in their current
/// order, then the sort keys.  The caller sorts (keys, perm) and re-gathers the points.
cudaError_t launch_kd_level_keys(const float4* cur_pts, uint32_t n, uint32_t count, uint32_t* boxes, uint64_t* keys, cudaStream_t st) {
  kd_boxes_init_kernel<<<(count * 6u + 255u) / 256u, 256, 0, st>>>(boxes, count);
  kd_boxes_kernel<<<(n + 255u) / 256u, 256, 0, st>>>(cur_pts, n, count, boxes);
  kd_keys_kernel<<<(n + 255u) / 256u, 256, 0, st>>>(cur_pts, n, count, boxes, keys);
  return cudaGetLastError();
}

/// first level whose nodes hold at most kKdSegMax points (from there on kd_refine_smem_kernel refines whole subtrees in shared memory)
uint32_t kd_smem_first_level(uint32_t n, uint32_t levels) {
  uint32_t l = 0;
  while (l < levels && (static_cast<uint64_t>(n) + (1ull << l) - 1ull) / (1ull << l) > kKdSegMax) l++;
  return l;
}
cudaError_t launch_kd_refine_smem(float4* leaf_pts, uint32_t* perm, uint32_t n, uint32_t level0, uint32_t levels, cudaStream_t st) {
  if (level0 >= levels) return cudaSuccess;
  const size_t smem = 2 * kKdSegMax * sizeof(float4) + kKdSegMax * (sizeof(unsigned long long) + sizeof(uint32_t)) + kKdSegNodes * 6 * sizeof(uint32_t);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kd_refine_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    configured = true;
  }
  kd_refine_smem_kernel<<<1u << level0, 1024, smem, st>>>(leaf_pts, perm, n, level0, levels);
  return cudaGetLastError();
}

cudaError_t sort_pairs_u64_u32_bits(void* d_temp, size_t& temp_bytes, const uint64_t* keys_in, uint64_t* keys_out, const uint32_t* vals_in, uint32_t* vals_out,
                                    size_t n, int end_bit, cudaStream_t st) {
  return cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, static_cast<int>(n), 0, end_bit, st);
}

// out[j] = in[perm[j]] for up to four float4 streams (leaf ordering of the target / Morton ordering of the source)
__global__ void gather_kernel(const uint32_t* __restrict__ perm, size_t n, const float4* __restrict__ in0, float4* out0, const float4* __restrict__ in1,
                              float4* out1, const float4* __restrict__ in2, float4* out2, const float4* __restrict__ in3, float4* out3) {
  for (size_t j = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; j < n; j += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint32_t s = perm[j];
    if (in0) out0[j] = in0[s];
    if (in1) out1[j] = in1[s];
    if (in2) out2[j] = in2[s];
    if (in3) out3[j] = in3[s];
  }
}

// out[perm[j]] = in[j]: leaf order -> original order (features estimated on the device in leaf order)
__global__ void scatter_kernel(const uint32_t* __restrict__ perm, size_t n, const float4* __restrict__ in0, float4* out0, const float4* __restrict__ in1,
                               float4* out1, const float4* __restrict__ in2, float4* out2) {
  for (size_t j = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; j < n; j += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint32_t d = perm[j];
    if (in0) out0[d] = in0[j];
    if (in1) out1[d] = in1[j];
    if (in2) out2[d] = in2[j];
  }
}

// Morton-rank order -> chunk-transposed order (see LinParams / DevSource::run): out[p] = in[rank(p)]
__global__ void chunk_transpose_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t n, uint32_t K) {
  const size_t chunk_pts = 32ull * K;
  for (size_t p = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; p < n; p += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t chunk = p / chunk_pts, r = p % chunk_pts;
    size_t m = p;
    if ((chunk + 1) * chunk_pts <= n) m = chunk * chunk_pts + (r % 32) * K + r / 32;  // full chunks only; the ragged tail keeps Morton order
    out[p] = in[m];
  }
}

// correspondences in the caller's order: out[perm[i]] = original target index (or voxel id << 32)
__global__ void correspondences_kernel(const uint32_t* __restrict__ corr, const uint32_t* __restrict__ perm, size_t n, const float4* __restrict__ tgt_pts,
                                       int voxel, uint64_t* out) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint32_t k = corr[i];
    uint64_t v = ~0ull;
    if (k != kNone) v = voxel ? (static_cast<uint64_t>(k) << 32) : static_cast<uint64_t>(static_cast<uint32_t>(__float_as_int(tgt_pts[k].w)));
    out[perm[i]] = v;
  }
}

static int grid_for(size_t n, int block, int cap) {
  size_t g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > static_cast<size_t>(cap)) g = cap;
  return static_cast<int>(g);
}

cudaError_t launch_bounds_centre(const double* d_pts4, size_t n, double* d_bounds6, double* d_centre4, int sm_count, cudaStream_t st) {
  bounds_init_kernel<<<1, 32, 0, st>>>(d_bounds6);
  if (n) bounds_kernel<<<grid_for(n, 256, sm_count * 8), 256, 0, st>>>(reinterpret_cast<const double4*>(d_pts4), n, d_bounds6);
  centre_kernel<<<1, 32, 0, st>>>(d_bounds6, d_centre4);
  return cudaGetLastError();
}

cudaError_t launch_convert(const double* d_pts4, const double* d_normals4, const double* d_covs16, size_t n, const double* d_centre4, float4* out_pts,
                           float4* out_normals, float4* out_covA, float4* out_covB, uint64_t* keys, uint32_t* vals, int sm_count, cudaStream_t st, float4* out_lo) {
  if (!n) return cudaSuccess;
  convert_kernel<<<grid_for(n, 256, sm_count * 8), 256, 0, st>>>(reinterpret_cast<const double4*>(d_pts4), reinterpret_cast<const double4*>(d_normals4), d_covs16,
                                                                 n, d_centre4, out_pts, out_normals, out_covA, out_covB, keys, vals, out_lo);
  return cudaGetLastError();
}

cudaError_t launch_inverse_perm(const uint32_t* perm, size_t n, uint32_t* rank, int sm_count, cudaStream_t st) {
  if (!n) return cudaSuccess;
  inverse_perm_kernel<<<grid_for(n, 256, sm_count * 8), 256, 0, st>>>(perm, n, rank);
  return cudaGetLastError();
}

cudaError_t launch_convert_cov_scatter(const double* d_covs16, size_t first, size_t n, const uint32_t* rank, float4* outA, float4* outB, float4* origA,
                                       float4* origB, int sm_count, cudaStream_t st) {
  if (!n) return cudaSuccess;
  convert_cov_scatter_kernel<<<grid_for(n, 256, sm_count * 8), 256, 0, st>>>(d_covs16, first, n, rank, outA, outB, origA, origB);
  return cudaGetLastError();
}

cudaError_t launch_gather(const uint32_t* perm, size_t n, const float4* in0, float4* out0, const float4* in1, float4* out1, const float4* in2, float4* out2,
                          const float4* in3, float4* out3, int sm_count, cudaStream_t st) {
  if (!n) return cudaSuccess;
  gather_kernel<<<grid_for(n, 256, sm_count * 8), 256, 0, st>>>(perm, n, in0, out0, in1, out1, in2, out2, in3, out3);
  return cudaGetLastError();
}

cudaError_t launch_scatter(const uint32_t* perm, size_t n, const float4* in0, float4* out0, const float4* in1, float4* out1, const float4* in2, float4* out2,
                           int sm_count, cudaStream_t st) {
  if (!n) return cudaSuccess;
  scatter_kernel<<<grid_for(n, 256, sm_count * 8), 256, 0, st>>>(perm, n, in0, out0, in1, out1, in2, out2);
  return cudaGetLastError();
}

cudaError_t launch_correspondences(const uint32_t* corr, const uint32_t* perm, size_t n, const float4* tgt_pts, int voxel, uint64_t* out, int sm_count,
                                   cudaStream_t st) {
  if (!n) return cudaSuccess;
  correspondences_kernel<<<grid_for(n, 256, sm_count * 8), 256, 0, st>>>(corr, perm, n, tgt_pts, voxel, out);
  return cudaGetLastError();
}

cudaError_t launch_chunk_transpose(const uint32_t* in, uint32_t* out, size_t n, uint32_t K, int sm_count, cudaStream_t st) {
  if (!n) return cudaSuccess;
  chunk_transpose_kernel<<<grid_for(n, 256, sm_count * 8), 256, 0, st>>>(in, out, n, K);
  return cudaGetLastError();
}

cudaError_t sort_pairs_u64_u32(void* d_temp, size_t& temp_bytes, const uint64_t* keys_in, uint64_t* keys_out, const uint32_t* vals_in, uint32_t* vals_out,
                               size_t n, cudaStream_t st) {
  return cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, static_cast<int>(n), 0, 63, st);
}

}  // namespace sgb
