// SPDX-License-Identifier: MIT
// Device-side data layout and per-point math of the B200 hot path.  See DESIGN.md §3-4.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sgb_math.cuh"

namespace sgb {

constexpr uint32_t kNone = 0xFFFFFFFFu;  // no correspondence
constexpr int kLinBlock = 128;           // threads per CTA of the linearize / error kernels
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
constexpr unsigned long long kCommPoison = 1ull << 63;  // flag bit: the sender could not run its reduction (its call failed on the host)
constexpr int kCommStampRing = 64;                      // wait times of the last 64 exchanges (sgb_comm_wait_ns)
struct CommParams {
  unsigned char* mail[kMaxPeers];  // mailbox of every rank (mail[rank] is local memory)
  int world, rank;                 // world <= 1: no exchange
  unsigned long long seq;          // sequence number of this call (same on all ranks, > 0)
  unsigned long long timeout_ns;   // give up waiting for a peer after this long (the result is then NaN and *status becomes non-zero)
  unsigned int* status;            // sticky device word: 1 = gave up on a peer, 2 = a peer reported a failed call (never cleared by the kernels)
  unsigned long long* stamps;      // [kCommStampRing]: ns this rank waited for its peers in exchange seq % kCommStampRing (or null)
  int poison;                      // this rank has nothing valid to contribute: tell the peers (host-side failure after the sequence number was taken)
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

// Sym3, the factor algebra and the per-point helpers live in sgb_math.cuh (also compiled for the host by tests/host_math).

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
/// Called by ALL threads of the finishing CTA.  A rank whose peers never show up gives up after c.timeout_ns, returns NaN and
/// raises the context's sticky status word, which the host-facing calls turn into a non-zero return code (a hung collective
/// must not take the GPU down with it, and must not pass for a result either).
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
  __shared__ int s_fail;
  if (tid == 0) {
    s_fail = 0;
    if (c.stamps) c.stamps[c.seq % kCommStampRing] = 0ull;
  }
  __syncthreads();
  if (tid < c.world) {
    __threadfence_system();
    volatile unsigned long long* theirs = reinterpret_cast<volatile unsigned long long*>(c.mail[tid] + kMailFlagOffset) + par * kMaxPeers + c.rank;
    *theirs = c.poison ? (c.seq | kCommPoison) : c.seq;  // publish: rank `tid` may now read my slot
    volatile unsigned long long* mine = reinterpret_cast<volatile unsigned long long*>(c.mail[c.rank] + kMailFlagOffset) + par * kMaxPeers + tid;
    unsigned long long t0, t1 = 0, f;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (((f = *mine) & ~kCommPoison) != c.seq) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > c.timeout_ns) {
        atomicOr(&s_fail, 1);
        break;
      }
    }
    if (f == (c.seq | kCommPoison)) atomicOr(&s_fail, 2);
    // how long this rank waited for peer `tid` (its own flag is there at once): the slowest peer is the skew + transport,
    // the rank that arrives last sees the bare transport latency (bench.py: comm_wait_ns, min / max over the ranks)
    if (c.stamps) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      atomicMax(&c.stamps[c.seq % kCommStampRing], t1 - t0);
    }
  }
  __syncthreads();
  __threadfence_system();
  double s = 0.0;
  if (tid < NVALS) {
    const volatile double* slots = reinterpret_cast<const volatile double*>(c.mail[c.rank]) + par * kMaxPeers * kMailStride;
    for (int p = 0; p < c.world; p++) s += slots[p * kMailStride + tid];  // rank order: the same sum on every rank
    if (s_fail) s = __longlong_as_double(0x7ff8000000000000ll);
  }
  if (tid == 0 && s_fail && c.status) atomicOr(c.status, static_cast<unsigned int>(s_fail));
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
  if (NACC > 4) {
    // Transposing butterfly: at every step a lane keeps one half of the values it still owns and hands the other half to its partner,
    // so the 32 lanes end up with ONE fully reduced accumulator each (lane l holds sum l): 16 + 8 + 4 + 2 + 1 = 31 exchanges instead of
    // 5 per accumulator (145 for the 29 sums of linearize -- 9 % of the factor kernel's instructions, profiles/r01/an).  The order of
    // the additions is fixed by the lane numbers: deterministic.
    double v[32];
#pragma unroll
    for (int k = 0; k < 32; k++) v[k] = k < NACC ? acc[k] : 0.0;
#pragma unroll
    for (int w = 16; w >= 1; w >>= 1) {
      const bool upper = (lane & w) != 0;
#pragma unroll
      for (int k = 0; k < w; k++) {
        if (k >= NACC) continue;  // both halves are padding
        const double keep = upper ? v[k + w] : v[k];
        const double send = upper ? v[k] : v[k + w];
        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, w);
      }
    }
    if (lane < NACC) s_red[warp][lane] = v[0];
  } else {
#pragma unroll
    for (int k = 0; k < NACC; k++) {
      double v = acc[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0) s_red[warp][k] = v;
    }
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
      // compact sum -> position(s) in H(6x6 row-major) | b | e | inliers (sgb_math.cuh)
      int p0, p1;
      if (expand_positions(threadIdx.x, p0, p1) == 2) out[p1] = v;
      out[p0] = v;
    }
  }
}

}  // namespace sgb
