// SPDX-License-Identifier: MIT
// Device-side data layout and per-point math of the B200 hot path.  See DESIGN.md §3-4.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sgb {

constexpr uint32_t kNone = 0xFFFFFFFFu;  // no correspondence
constexpr int kLinBlock = 128;           // threads per CTA of the linearize / error kernels
constexpr int kAcc = 28;                 // 6 (H_rr sym) + 9 (H_rt) + 6 (H_tt sym) + 6 (b) + 1 (e)
constexpr int kPartialStride = 32;       // doubles per CTA partial (28 sums + inlier count, padded)

/// Flattened kd-tree node, 8 bytes, nodes in PRE-ORDER so that the left child of node i is i+1.
///   inner: x = split threshold (float bits), y = (right_child << 2) | axis        (axis in 0..2)
///   leaf : x = first point (offset into the leaf-ordered point arrays), y = (count << 2) | 3
using KdNode = uint2;

struct DevTarget {
  const float4* pts;      // leaf order; xyz relative to `centre`; w = original index (int bits) / unused for voxels
  const float4* normals;  // leaf order (plane ICP) or null
  const float4* covA;     // (xx, xy, xz, yy)
  const float4* covB;     // (yz, zz, 0, 0)
  const KdNode* nodes;    // kd-tree (null for voxel maps)
  const double* centre;   // 3 doubles, device
  // voxel map (VGICP)
  const int4* vox_table;  // open addressing: (x, y, z, voxel_id), voxel_id < 0 = empty
  uint32_t vox_mask;      // capacity - 1 (power of two)
  int vox_num_offsets;    // 1, 7 or 27
  double vox_inv_leaf;
};

struct DevSource {
  const float4* pts;   // Morton order; xyz relative to `centre`
  const float4* covA;  // Morton order
  const float4* covB;
  const double* centre;  // 3 doubles, device
  uint32_t n;
  uint32_t run;  // K: source positions chunk*32K + k*32 + lane hold Morton ranks chunk*32K + lane*K + k (full chunks),
                 // i.e. one lane walks K spatially consecutive points while every load of the warp stays coalesced
};

/// Multi-GPU exchange fused into the finishing CTA of the reduction (one process per GPU, source sharded).
/// Every rank owns a MAILBOX in its device memory that its peers map through CUDA IPC:
///   doubles [2 parities][kMaxPeers senders][kMailStride]   then   unsigned long long flags [2 parities][kMaxPeers senders]
/// Rank r stores its sums into slot (parity, r) of every peer's mailbox over NVLink, fences, stores the call's sequence
/// number into the matching flag, waits until its OWN mailbox shows the sequence number of every sender and adds the
/// slots in rank order (deterministic, identical on all ranks).  Two parities: a rank can be at most one call ahead.
constexpr int kMaxPeers = 8;
constexpr int kMailStride = 64;
constexpr size_t kMailFlagOffset = 2 * kMaxPeers * kMailStride * sizeof(double);
constexpr size_t kMailBytes = kMailFlagOffset + 2 * kMaxPeers * sizeof(unsigned long long);
struct CommParams {
  unsigned char* mail[kMaxPeers];  // mailbox of every rank (mail[rank] is local memory)
  int world, rank;                 // world <= 1: no exchange
  unsigned long long seq;          // sequence number of this call (same on all ranks, > 0)
  unsigned long long timeout_ns;   // give up waiting for a peer after this long (the result is then NaN)
};

struct LinParams {
  DevTarget tgt;
  DevSource src;
  CommParams comm;
  double T[12];      // row-major R (9) then t (3): T_target_source of this call
  double Tlin[12];   // pose of the last linearize (error kernel, GICP precision matrix)
  float max_dist_sq;     // search bound in FP32, a hair above the rejector's threshold (FLT_MAX for NullRejector)
  double max_dist_sq_d;  // the rejector's threshold itself, applied to the FP64 residual (DBL_MAX for NullRejector)
  int use_prev;          // corr[] holds the previous linearize's correspondences for the same clouds: use them as search seeds
  double robust_c;
  uint32_t* corr;     // per source point (Morton order): leaf-order position / voxel id, or kNone
  double* partials;   // gridDim.x * kPartialStride
  unsigned int* ticket;
  double* out;        // 44 doubles: H(36) | b(6) | e | num_inliers     (error kernel: out[0] = e)
};

/// Programmatic dependent launch (sm_90+).  A kernel launched with launch_dependent() (sgb_kernels.h) may be set up and
/// scheduled while its predecessor on the stream is still draining; it must call this before it touches anything the
/// predecessor wrote.  Without the launch attribute the instruction returns immediately.
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t vox_hash(int x, int y, int z) {
  // any hash works (lookups are exact-match); this is a 32-bit mix of the three coordinates
  uint32_t h = static_cast<uint32_t>(x) * 73856093u ^ static_cast<uint32_t>(y) * 19349669u ^ static_cast<uint32_t>(z) * 83492791u;
  h ^= h >> 15;
  h *= 0x2c1b3c6du;
  h ^= h >> 12;
  return h;
}

/// Symmetric 3x3 stored as (xx, xy, xz, yy, yz, zz).
struct Sym3 {
  double xx, xy, xz, yy, yz, zz;
};

__device__ __forceinline__ Sym3 sym3_inverse(const Sym3& a) {
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
__device__ __forceinline__ Sym3 gicp_precision(const double* R, const float4& sA, const float4& sB, const float4& tA, const float4& tB) {
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
__device__ __forceinline__ void accumulate_factor(const double* R, const Sym3& M, double rx, double ry, double rz, double px, double py, double pz,
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
__device__ __forceinline__ Sym3 gicp_precision_source(const double* R, const float4& sA, const float4& sB, const float4& tA, const float4& tB) {
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
__device__ __forceinline__ Sym3 plane_weight_source(const double* R, double nx, double ny, double nz) {
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
__device__ __forceinline__ void accumulate_factor_source(const Sym3& D, double rsx, double rsy, double rsz, double px, double py, double pz, double robust_c,
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

/// Exact nearest neighbour of q in the flattened kd-tree (restates the visiting order of
/// UnsafeKdTree::knn_search, ann/kdtree.hpp:193-233, with an explicit stack of far children).
/// `best_d` enters as the search bound (only strictly closer points are accepted, knn_result.hpp:81-83).
/// stack: shared memory, entry (s, lane) at stack[s * kLinBlock + threadIdx.x].
__device__ __forceinline__ uint32_t kd_nearest(const KdNode* __restrict__ nodes, const float4* __restrict__ pts, float qx, float qy, float qz,
                                               float& best_d, uint32_t best, uint2* stack) {
  // `best` / `best_d` may enter seeded with a candidate (an upper bound only prunes: the result is still the
  // exact nearest neighbour, the seed wins only if nothing is strictly closer).
  uint32_t node = 0;
  int sp = 0;
  uint2* my_stack = stack + threadIdx.x;
  for (;;) {
    KdNode nd = __ldg(&nodes[node]);
    uint32_t kind = nd.y & 3u;
    while (kind != 3u) {  // descend; remember the far child only if the current bound does not already prune it
      const float qv = kind == 0u ? qx : (kind == 1u ? qy : qz);
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
    {  // leaf: scan its contiguous block of points
      const uint32_t first = nd.x, cnt = nd.y >> 2;
      const float4* lp = pts + first;
#pragma unroll 4
      for (uint32_t j = 0; j < cnt; j++) {
        const float4 t = __ldg(&lp[j]);
        const float dx = t.x - qx, dy = t.y - qy, dz = t.z - qz;
        const float d = dx * dx + dy * dy + dz * dz;
        if (d < best_d) {
          best_d = d;
          best = first + j;
        }
      }
    }
    // backtrack: nearest pending far child whose splitting plane is closer than the best distance
    bool found = false;
    while (sp > 0) {
      sp--;
      const uint2 e = my_stack[sp * kLinBlock];
      if (__uint_as_float(e.y) < best_d) {
        node = e.x;
        found = true;
        break;
      }
    }
    if (!found) break;
  }
  return best;
}

/// All-reduce(sum) of one value per thread (threads < NVALS of ONE CTA per rank) over the peer mailboxes -- see CommParams.
/// Called by ALL threads of the finishing CTA.  A rank whose peers never show up gives up after c.timeout_ns and returns NaN
/// (a hung collective must not take the GPU down with it).
template <int NVALS>
__device__ __forceinline__ double comm_all_reduce(const CommParams& c, double v) {
  static_assert(NVALS <= kMailStride, "mailbox slot too small");
  const int tid = threadIdx.x;
  const size_t par = static_cast<size_t>(c.seq & 1ull);
  if (tid < NVALS) {
    for (int p = 0; p < c.world; p++)  // my sums -> slot (par, rank) of every rank's mailbox (remote stores over NVLink)
      reinterpret_cast<double*>(c.mail[p])[(par * kMaxPeers + c.rank) * kMailStride + tid] = v;
  }
  __threadfence_system();
  __syncthreads();
  __shared__ int s_timeout;
  if (tid == 0) s_timeout = 0;
  __syncthreads();
  if (tid < c.world) {
    __threadfence_system();
    volatile unsigned long long* theirs = reinterpret_cast<volatile unsigned long long*>(c.mail[tid] + kMailFlagOffset) + par * kMaxPeers + c.rank;
    *theirs = c.seq;  // publish: rank `tid` may now read my slot
    volatile unsigned long long* mine = reinterpret_cast<volatile unsigned long long*>(c.mail[c.rank] + kMailFlagOffset) + par * kMaxPeers + tid;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (*mine != c.seq) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > c.timeout_ns) {
        s_timeout = 1;
        break;
      }
    }
  }
  __syncthreads();
  __threadfence_system();
  double s = 0.0;
  if (tid < NVALS) {
    const volatile double* slots = reinterpret_cast<const volatile double*>(c.mail[c.rank]) + par * kMaxPeers * kMailStride;
    for (int p = 0; p < c.world; p++) s += slots[p * kMailStride + tid];  // rank order: the same sum on every rank
    if (s_timeout) s = __longlong_as_double(0x7ff8000000000000ll);
  }
  return s;
}

/// CTAs per group of the two-level final reduction (see block_reduce_and_finish)
constexpr unsigned int kFinishGroup = 16;
/// doubles / tickets the reduction needs for a grid of `grid` CTAs
__host__ __device__ inline size_t reduction_partials_doubles(size_t grid) { return (grid + (grid + kFinishGroup - 1) / kFinishGroup) * kPartialStride; }
__host__ __device__ inline size_t reduction_tickets(size_t grid) { return 1 + (grid + kFinishGroup - 1) / kFinishGroup; }

/// Block-wide sum of per-thread accumulators -> partials[blockIdx.x].  The sums of all CTAs are then combined by a
/// two-level tree of tickets: the last CTA of every group of kFinishGroup consecutive CTAs adds that group's partials,
/// the last group to finish adds the group sums and writes the (expanded 44-double) result.  Each level is one batch of
/// independent L2 loads per lane -- a single CTA adding all gridDim.x partials cost ~15-25 us of an otherwise idle GPU
/// (profiles/r01/p, r01/t).  The order of the additions depends on gridDim.x only: deterministic.
/// partials: reduction_partials_doubles(gridDim.x) doubles; ticket: reduction_tickets(gridDim.x) zeroed counters (left zeroed).
/// comm.world > 1: the finishing CTA then exchanges the sums with the other ranks (comm_all_reduce) before writing `out`.
template <int NACC, bool EXPAND>
__device__ __forceinline__ void block_reduce_and_finish(double* acc, double* partials, unsigned int* ticket, double* out, const CommParams& comm) {
  static_assert(NACC <= 32, "one lane per accumulator");
  __shared__ double s_red[kLinBlock / 32][NACC];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NACC; k++) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < NACC) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kLinBlock / 32; w++) v += s_red[w][threadIdx.x];
    partials[static_cast<size_t>(blockIdx.x) * kPartialStride + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  const unsigned int group = blockIdx.x / kFinishGroup, n_groups = (gridDim.x + kFinishGroup - 1) / kFinishGroup;
  const unsigned int group_first = group * kFinishGroup;
  const unsigned int group_size = min(kFinishGroup, gridDim.x - group_first);
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(&ticket[1 + group], 1u);
    s_last = (t == group_size - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double* group_sums = partials + static_cast<size_t>(gridDim.x) * kPartialStride;
  if (threadIdx.x < NACC) {  // level 1: this group's CTAs, in CTA order
    double v = 0.0;
#pragma unroll
    for (unsigned int j = 0; j < kFinishGroup; j++)
      if (j < group_size) v += __ldcg(&partials[static_cast<size_t>(group_first + j) * kPartialStride + threadIdx.x]);
    group_sums[static_cast<size_t>(group) * kPartialStride + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    ticket[1 + group] = 0u;  // ready for the next launch on this stream
    const unsigned int t = atomicAdd(&ticket[0], 1u);
    s_last = (t == n_groups - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x == 0) ticket[0] = 0u;
  double v = 0.0;
  if (threadIdx.x < NACC) {  // level 2: the group sums, in group order
#pragma unroll 8
    for (unsigned int gI = 0; gI < n_groups; gI++) v += __ldcg(&group_sums[static_cast<size_t>(gI) * kPartialStride + threadIdx.x]);
  }
  if (comm.world > 1) v = comm_all_reduce<NACC>(comm, v);  // level 3: the other GPUs (fused exchange over peer memory)
  if (threadIdx.x < NACC) {
    if (!EXPAND) {
      out[threadIdx.x] = v;
    } else {
      // acc index -> positions in H(6x6 row-major) | b | e | inliers
      const int k = threadIdx.x;
      if (k < 6) {  // H_rr upper triangle: (0,0)(0,1)(0,2)(1,1)(1,2)(2,2)
        const int r = k < 3 ? 0 : (k < 5 ? 1 : 2);
        const int c = k < 3 ? k : (k < 5 ? k - 2 : 2);
        out[r * 6 + c] = v;
        out[c * 6 + r] = v;
      } else if (k < 15) {  // H_rt 3x3 row-major -> rows 0..2, cols 3..5 (+ transpose)
        const int r = (k - 6) / 3, c = (k - 6) % 3;
        out[r * 6 + 3 + c] = v;
        out[(3 + c) * 6 + r] = v;
      } else if (k < 21) {  // H_tt upper triangle
        const int kk = k - 15;
        const int r = kk < 3 ? 0 : (kk < 5 ? 1 : 2);
        const int c = kk < 3 ? kk : (kk < 5 ? kk - 2 : 2);
        out[(3 + r) * 6 + 3 + c] = v;
        out[(3 + c) * 6 + 3 + r] = v;
      } else if (k < 27) {
        out[36 + (k - 21)] = v;
      } else if (k == 27) {
        out[42] = v;
      } else {
        out[43] = v;  // inlier count
      }
    }
  }
}

}  // namespace sgb
