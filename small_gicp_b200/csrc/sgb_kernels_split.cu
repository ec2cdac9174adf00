// SPDX-License-Identifier: MIT
// Two-phase variant of linearize (profiles/r01):
//   phase 1  nn_search_kernel      transform + exact kd-tree NN per source point, ~40 registers -> 3-4x the resident
//                                  warps of the fused kernel (the search is latency-bound: dependent node loads)
//   phase 2  factor_reduce_kernel  streams source + correspondences, gathers the matched target point / covariance,
//                                  FP64 rejector + factor algebra + block reduction + last-CTA finish
// Both phases run back to back on the context's stream; the only extra traffic is the 4-byte correspondence
// per point that sgb_error()/sgb_correspondences() need anyway.
#include <cfloat>

#include "sgb_device.cuh"
#include "sgb_kernels.h"

namespace sgb {

// ---------------------------------------------------------------------------------------------------------------
// Phase 1.  Every lane owns a run of K consecutive points of the Morton curve (chunk-transposed layout) and walks
// it INDEPENDENTLY of the other lanes: a lane that finishes a query starts its next one in the same iteration of the
// warp loop instead of idling until the slowest query of the warp is done.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kLinBlock) nn_search_kernel(const __grid_constant__ LinParams P) {
  extern __shared__ uint2 s_stack[];
  const double* R = P.T;
  const double csx = P.src.centre[0], csy = P.src.centre[1], csz = P.src.centre[2];
  const double tpx = R[0] * csx + R[1] * csy + R[2] * csz + P.T[9] - P.tgt.centre[0];
  const double tpy = R[3] * csx + R[4] * csy + R[5] * csz + P.T[10] - P.tgt.centre[1];
  const double tpz = R[6] * csx + R[7] * csy + R[8] * csz + P.T[11] - P.tgt.centre[2];
  // the FP32 query is the FP64 transform rounded once (same value phase 2 derives its residual from)
  const KdNode* __restrict__ nodes = P.tgt.nodes;
  const float4* __restrict__ pts = P.tgt.pts;

  const uint32_t K = P.src.run, chunk_pts = 32u * K;
  const uint32_t n_chunks = (P.src.n + chunk_pts - 1) / chunk_pts;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  uint2* my_stack = s_stack + threadIdx.x;

  for (uint32_t chunk = warp; chunk < n_chunks; chunk += n_warps) {
    uint32_t k = 0, chain = kNone;
    bool active = false;
    uint32_t i = 0, node = 0, best = kNone;
    int sp = 0;
    float fx = 0.f, fy = 0.f, fz = 0.f, best_d = 0.f;
    for (;;) {
      if (!active) {  // start the lane's next query
        i = chunk * chunk_pts + k * 32u + lane;
        if (k >= K || i >= P.src.n) break;
        const float4 s = __ldg(&P.src.pts[i]);
        const double sx = s.x, sy = s.y, sz = s.z;
        fx = static_cast<float>(R[0] * sx + R[1] * sy + R[2] * sz + tpx);
        fy = static_cast<float>(R[3] * sx + R[4] * sy + R[5] * sz + tpy);
        fz = static_cast<float>(R[6] * sx + R[7] * sy + R[8] * sz + tpz);
        best_d = P.max_dist_sq;
        best = kNone;
        const uint32_t prev = P.use_prev ? P.corr[i] : kNone;
        if (prev != kNone) {
          const float4 t = __ldg(&pts[prev]);
          const float dx = t.x - fx, dy = t.y - fy, dz = t.z - fz;
          const float d = dx * dx + dy * dy + dz * dz;
          if (d < best_d) {
            best_d = d;
            best = prev;
          }
        }
        if (chain != kNone && chain != prev) {
          const float4 t = __ldg(&pts[chain]);
          const float dx = t.x - fx, dy = t.y - fy, dz = t.z - fz;
          const float d = dx * dx + dy * dy + dz * dz;
          if (d < best_d) {
            best_d = d;
            best = chain;
          }
        }
        node = 0;
        sp = 0;
        active = true;
      }
      // descend to a leaf
      KdNode nd = __ldg(&nodes[node]);
      uint32_t kind = nd.y & 3u;
      while (kind != 3u) {
        const float qv = kind == 0u ? fx : (kind == 1u ? fy : fz);
        const float diff = qv - __uint_as_float(nd.x);
        const uint32_t right = nd.y >> 2, left = node + 1u;
        const bool go_left = diff < 0.0f;
        const float cut = diff * diff;
        if (cut < best_d) {
          my_stack[sp * kLinBlock] = make_uint2(go_left ? right : left, __float_as_uint(cut));
          sp++;
        }
        node = go_left ? left : right;
        nd = __ldg(&nodes[node]);
        kind = nd.y & 3u;
      }
      {  // scan the leaf
        const uint32_t first = nd.x, cnt = nd.y >> 2;
        const float4* lp = pts + first;
        for (uint32_t j = 0; j < cnt; j++) {
          const float4 t = __ldg(&lp[j]);
          const float dx = t.x - fx, dy = t.y - fy, dz = t.z - fz;
          const float d = dx * dx + dy * dy + dz * dz;
          if (d < best_d) {
            best_d = d;
            best = first + j;
          }
        }
      }
      // next pending subtree that can still hold a closer point, else the query is finished
      bool resumed = false;
      while (sp > 0) {
        sp--;
        const uint2 e = my_stack[sp * kLinBlock];
        if (__uint_as_float(e.y) < best_d) {
          node = e.x;
          resumed = true;
          break;
        }
      }
      if (!resumed) {
        P.corr[i] = best;
        if (best != kNone) chain = best;
        k++;
        active = false;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Phase 2.
// ---------------------------------------------------------------------------------------------------------------
template <int FACTOR, int ROBUST>
__global__ void __launch_bounds__(kLinBlock) factor_reduce_kernel(const __grid_constant__ LinParams P) {
  double acc[kAcc + 1];
#pragma unroll
  for (int k = 0; k <= kAcc; k++) acc[k] = 0.0;
  const double* R = P.T;
  const double csx = P.src.centre[0], csy = P.src.centre[1], csz = P.src.centre[2];
  const double tpx = R[0] * csx + R[1] * csy + R[2] * csz + P.T[9] - P.tgt.centre[0];
  const double tpy = R[3] * csx + R[4] * csy + R[5] * csz + P.T[10] - P.tgt.centre[1];
  const double tpz = R[6] * csx + R[7] * csy + R[8] * csz + P.T[11] - P.tgt.centre[2];
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P.src.n; i += stride) {
    const uint32_t best = P.corr[i];
    if (best == kNone) continue;
    const float4 sp = __ldg(&P.src.pts[i]);
    const double sx = sp.x, sy = sp.y, sz = sp.z;
    const double qx = R[0] * sx + R[1] * sy + R[2] * sz + tpx;
    const double qy = R[3] * sx + R[4] * sy + R[5] * sz + tpy;
    const double qz = R[6] * sx + R[7] * sy + R[8] * sz + tpz;
    const float4 tq = __ldg(&P.tgt.pts[best]);
    const double rx = static_cast<double>(tq.x) - qx, ry = static_cast<double>(tq.y) - qy, rz = static_cast<double>(tq.z) - qz;
    if (rx * rx + ry * ry + rz * rz > P.max_dist_sq_d) {  // DistanceRejector on the FP64 residual (rejector.hpp:24)
      P.corr[i] = kNone;
      continue;
    }
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
  block_reduce_and_finish<kAcc + 1, true>(acc, P.partials, P.ticket, P.out);
}

template <int FACTOR, int ROBUST>
static cudaError_t launch_factor(const LinParams& P, int grid, cudaStream_t st) {
  factor_reduce_kernel<FACTOR, ROBUST><<<grid, kLinBlock, 0, st>>>(P);
  return cudaGetLastError();
}

int search_occupancy(int stack_depth) {
  int nb = 0;
  const size_t smem = static_cast<size_t>(stack_depth) * kLinBlock * sizeof(uint2);
  if (smem > 48 * 1024) cudaFuncSetAttribute(nn_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, nn_search_kernel, kLinBlock, smem) != cudaSuccess) return 1;
  return nb > 0 ? nb : 1;
}

cudaError_t launch_search(const LinParams& P, int grid, int stack_depth, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(stack_depth) * kLinBlock * sizeof(uint2);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(nn_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  nn_search_kernel<<<grid, kLinBlock, smem, st>>>(P);
  return cudaGetLastError();
}

cudaError_t launch_factor_reduce(const LinParams& P, int factor, int robust, int grid, cudaStream_t st) {
  switch (factor * 3 + robust) {
    case 0: return launch_factor<0, 0>(P, grid, st);
    case 1: return launch_factor<0, 1>(P, grid, st);
    case 2: return launch_factor<0, 2>(P, grid, st);
    case 3: return launch_factor<1, 0>(P, grid, st);
    case 4: return launch_factor<1, 1>(P, grid, st);
    case 5: return launch_factor<1, 2>(P, grid, st);
    case 6: return launch_factor<2, 0>(P, grid, st);
    case 7: return launch_factor<2, 1>(P, grid, st);
    case 8: return launch_factor<2, 2>(P, grid, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace sgb
