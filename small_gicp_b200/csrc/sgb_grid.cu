// SPDX-License-Identifier: MIT
// Uniform-grid front end of the exact nearest-neighbour search.
//
// In a registration that is not wildly misaligned almost every query's nearest neighbour is a fraction of the point
// spacing away.  For such a query a hash probe of the 2 x 2 x 2 block of grid cells (cell edge c) that covers the cube
// [q - c/2, q + c/2]^3 sees EVERY target point within c/2 of q, so if the closest point found is within c/2 it is the
// exact nearest neighbour and the query is finished: 8 independent loads + a few dozen distance tests, no tree walk,
// no dependent pointer chasing.  Queries it cannot settle (nothing within c/2: misaligned first iterations, holes,
// outliers) are left pending -- with the best candidate found so far as an upper bound -- for the packet tree search
// (sgb_kernels_packet.cu), which still returns exact results for them.  Results are therefore identical to a pure tree
// search (exact ties aside); ncu evidence and the A/B switch (SGB_GRID=0) are recorded in profiles/.
#include <cfloat>
#include <cub/device/device_radix_sort.cuh>

#include "sgb_device.cuh"
#include "sgb_kernels.h"

namespace sgb {

__device__ __forceinline__ uint64_t grid_key(int ix, int iy, int iz) {
  return (static_cast<uint64_t>(static_cast<uint32_t>(ix + (1 << 20)) & 0x1fffffu)) | (static_cast<uint64_t>(static_cast<uint32_t>(iy + (1 << 20)) & 0x1fffffu) << 21) |
         (static_cast<uint64_t>(static_cast<uint32_t>(iz + (1 << 20)) & 0x1fffffu) << 42);
}
__device__ __forceinline__ uint32_t grid_hash(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return static_cast<uint32_t>(k);
}

// cell key of every target point (leaf-ordered, centred FP32); value = leaf position
__global__ void grid_keys_kernel(const float4* __restrict__ pts, uint32_t n, GridParams g, uint64_t* keys, uint32_t* vals) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  const int ix = __float2int_rd((p.x - g.origin[0]) * g.inv_cell), iy = __float2int_rd((p.y - g.origin[1]) * g.inv_cell),
            iz = __float2int_rd((p.z - g.origin[2]) * g.inv_cell);
  keys[i] = grid_key(ix, iy, iz);
  vals[i] = i;
}

// after the sort: gather the points into cell order (w = leaf position) and insert one table entry per run of equal keys
__global__ void grid_fill_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ leaf_pos, const float4* __restrict__ leaf_pts, uint32_t n,
                                 float4* grid_pts, GridSlot* table, uint32_t mask) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t lp = leaf_pos[i];
  const float4 p = leaf_pts[lp];
  grid_pts[i] = make_float4(p.x, p.y, p.z, __uint_as_float(lp));
  const uint64_t k = keys[i];
  if (i == 0 || keys[i - 1] != k) {  // head of a cell: count its points, claim a slot
    uint32_t cnt = 1;
    while (i + cnt < n && keys[i + cnt] == k) cnt++;
    uint32_t slot = grid_hash(k) & mask;
    for (;;) {
      const unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(&table[slot].key), ~0ull, static_cast<unsigned long long>(k));
      if (prev == ~0ull) break;
      slot = (slot + 1u) & mask;
    }
    table[slot].start = i;
    table[slot].count = cnt;
  }
}

__global__ void grid_table_init_kernel(GridSlot* table, uint32_t capacity) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < capacity) {
    table[i].key = ~0ull;
    table[i].start = 0u;
    table[i].count = 0u;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// probe: one query per thread (Hilbert order -> neighbouring lanes hit neighbouring cells)
// state[i] = 1: settled (corr[i] is the exact nearest neighbour or kNone is impossible here), 0: pending for the tree search
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grid_probe_kernel(const __grid_constant__ LinParams P, const float4* __restrict__ grid_pts,
                                                        const GridSlot* __restrict__ table, uint32_t mask, GridParams g, uint8_t* state,
                                                        uint32_t* pending_count, uint32_t* pending_list) {
  const uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = gi < P.src.n;
  const uint32_t i = in_range ? gi : P.src.n - 1u;  // out-of-range lanes shadow the last query (no divergent exit before the warp vote below)
  const double* R = P.T;
  const double csx = P.src.centre[0], csy = P.src.centre[1], csz = P.src.centre[2];
  const double tpx = R[0] * csx + R[1] * csy + R[2] * csz + P.T[9] - P.tgt.centre[0];
  const double tpy = R[3] * csx + R[4] * csy + R[5] * csz + P.T[10] - P.tgt.centre[1];
  const double tpz = R[6] * csx + R[7] * csy + R[8] * csz + P.T[11] - P.tgt.centre[2];
  const float4 s = __ldg(&P.src.pts[i]);
  const double sx = s.x, sy = s.y, sz = s.z;
  const float fx = static_cast<float>(R[0] * sx + R[1] * sy + R[2] * sz + tpx);
  const float fy = static_cast<float>(R[3] * sx + R[4] * sy + R[5] * sz + tpy);
  const float fz = static_cast<float>(R[6] * sx + R[7] * sy + R[8] * sz + tpz);

  float best_d = P.max_dist_sq;
  uint32_t best = kNone;
  // the 2 x 2 x 2 block covering [q - c/2, q + c/2]^3
  const int bx = __float2int_rd((fx - g.origin[0]) * g.inv_cell - 0.5f), by = __float2int_rd((fy - g.origin[1]) * g.inv_cell - 0.5f),
            bz = __float2int_rd((fz - g.origin[2]) * g.inv_cell - 0.5f);
  uint32_t starts[8], counts[8];
#pragma unroll
  for (int c = 0; c < 8; c++) {
    const uint64_t k = grid_key(bx + (c & 1), by + ((c >> 1) & 1), bz + (c >> 2));
    uint32_t slot = grid_hash(k) & mask;
    uint32_t st = 0, cn = 0;
    for (;;) {
      const GridSlot e = table[slot];
      if (e.key == k) {
        st = e.start;
        cn = e.count;
        break;
      }
      if (e.key == ~0ull) break;
      slot = (slot + 1u) & mask;
    }
    starts[c] = st;
    counts[c] = cn;
  }
#pragma unroll
  for (int c = 0; c < 8; c++) {
    const float4* cp = grid_pts + starts[c];
    for (uint32_t j = 0; j < counts[c]; j++) {
      const float4 t = __ldg(&cp[j]);
      const float dx = t.x - fx, dy = t.y - fy, dz = t.z - fz;
      const float d = dx * dx + dy * dy + dz * dz;
      if (d < best_d) {
        best_d = d;
        best = __float_as_uint(t.w);
      }
    }
  }
  const bool settled = best != kNone && best_d <= g.settle_d2;
  if (!settled && P.use_prev) {  // keep the better of (grid candidate, previous correspondence) as the tree search's seed
    const uint32_t prev = P.corr[i];
    if (prev != kNone && best == kNone) best = prev;
    else if (prev != kNone) {
      const float4 t = __ldg(&P.tgt.pts[prev]);
      const float dx = t.x - fx, dy = t.y - fy, dz = t.z - fz;
      if (dx * dx + dy * dy + dz * dz < best_d) best = prev;
    }
  }
  if (in_range) {
    P.corr[i] = best;
    state[i] = settled ? 1 : 0;
  }
  // pending queries go to a compact work list (warp-aggregated append) for the per-thread tree search
  const bool pend = in_range && !settled;
  const unsigned m = __ballot_sync(0xffffffffu, pend);
  if (m) {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t base = 0;
    if (lane == static_cast<uint32_t>(__ffs(m) - 1)) base = atomicAdd(pending_count, static_cast<uint32_t>(__popc(m)));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (pend) pending_list[base + __popc(m & ((1u << lane) - 1u))] = i;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Pending queries (a few per cent, scattered): exact NN by an individual walk of the packet records, seeded with the
// candidate left in corr[].  Skipped (the warp-cooperative search runs instead) when more than `max_pending` are pending.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float grid_box_dist2(float qx, float qy, float qz, const float4& lo, const float4& hi) {
  const float dx = fmaxf(fmaxf(lo.x - qx, qx - hi.x), 0.0f);
  const float dy = fmaxf(fmaxf(lo.y - qy, qy - hi.y), 0.0f);
  const float dz = fmaxf(fmaxf(lo.z - qz, qz - hi.z), 0.0f);
  return dx * dx + dy * dy + dz * dz;
}

__global__ void __launch_bounds__(kLinBlock) pending_search_kernel(const __grid_constant__ LinParams P, const float4* __restrict__ pnodes, int depth,
                                                                   const uint32_t* __restrict__ pending_count, const uint32_t* __restrict__ pending_list,
                                                                   uint32_t max_pending) {
  extern __shared__ uint2 s_desc[];  // [depth][kLinBlock] descriptors, then [depth][kLinBlock] box distances
  const uint32_t count = *pending_count;
  if (count > max_pending) return;
  float* s_dist = reinterpret_cast<float*>(s_desc + static_cast<size_t>(depth) * kLinBlock);
  uint2* my_desc = s_desc + threadIdx.x;
  float* my_dist = s_dist + threadIdx.x;
  const double* R = P.T;
  const double csx = P.src.centre[0], csy = P.src.centre[1], csz = P.src.centre[2];
  const double tpx = R[0] * csx + R[1] * csy + R[2] * csz + P.T[9] - P.tgt.centre[0];
  const double tpy = R[3] * csx + R[4] * csy + R[5] * csz + P.T[10] - P.tgt.centre[1];
  const double tpz = R[6] * csx + R[7] * csy + R[8] * csz + P.T[11] - P.tgt.centre[2];
  const float4* __restrict__ pts = P.tgt.pts;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
    const uint32_t i = pending_list[k];
    const float4 s = __ldg(&P.src.pts[i]);
    const double sx = s.x, sy = s.y, sz = s.z;
    const float qx = static_cast<float>(R[0] * sx + R[1] * sy + R[2] * sz + tpx);
    const float qy = static_cast<float>(R[3] * sx + R[4] * sy + R[5] * sz + tpy);
    const float qz = static_cast<float>(R[6] * sx + R[7] * sy + R[8] * sz + tpz);
    float best_d = P.max_dist_sq;
    uint32_t best = kNone;
    const uint32_t seed = P.corr[i];
    if (seed != kNone) {
      const float4 t = __ldg(&pts[seed]);
      const float dx = t.x - qx, dy = t.y - qy, dz = t.z - qz;
      const float d = dx * dx + dy * dy + dz * dz;
      if (d < best_d) {
        best_d = d;
        best = seed;
      }
    }
    int sp = 0;
    uint32_t cur = 0;
    bool expand = true;
    uint2 leaf = make_uint2(0u, 0u);
    for (;;) {
      if (expand) {
        const float4 n0 = __ldg(&pnodes[cur * 4 + 0]), n1 = __ldg(&pnodes[cur * 4 + 1]);
        const float4 n2 = __ldg(&pnodes[cur * 4 + 2]), n3 = __ldg(&pnodes[cur * 4 + 3]);
        const float dl = grid_box_dist2(qx, qy, qz, n0, n1), dr = grid_box_dist2(qx, qy, qz, n2, n3);
        const bool left_first = dl <= dr;
        const float dn = left_first ? dl : dr, df = left_first ? dr : dl;
        const uint2 cn = left_first ? make_uint2(__float_as_uint(n0.w), __float_as_uint(n1.w)) : make_uint2(__float_as_uint(n2.w), __float_as_uint(n3.w));
        const uint2 cf = left_first ? make_uint2(__float_as_uint(n2.w), __float_as_uint(n3.w)) : make_uint2(__float_as_uint(n0.w), __float_as_uint(n1.w));
        if (df < best_d) {
          my_desc[sp * kLinBlock] = cf;
          my_dist[sp * kLinBlock] = df;
          sp++;
        }
        if (!(dn < best_d)) {
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
          if (my_dist[sp * kLinBlock] < best_d) {
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
        const float d = dx * dx + dy * dy + dz * dz;
        if (d < best_d) {
          best_d = d;
          best = leaf.x + j;
        }
      }
      expand = false;
    }
    P.corr[i] = best;
  }
}

// per-leaf spacing estimate from the packet records (two largest box extents / count) for the choice of the cell size
__global__ void grid_spacing_kernel(const float4* __restrict__ pnodes, uint32_t n_inner, float* out) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_inner * 2u) return;
  const float4 lo = pnodes[(t >> 1) * 4 + (t & 1u) * 2 + 0], hi = pnodes[(t >> 1) * 4 + (t & 1u) * 2 + 1];
  const uint32_t cnt = __float_as_uint(hi.w);
  float v = -1.0f;
  if (cnt >= 4u && hi.x >= lo.x) {
    float e0 = hi.x - lo.x, e1 = hi.y - lo.y, e2 = hi.z - lo.z;
    const float mn = fminf(e0, fminf(e1, e2));
    const float area = (e0 * e1 * e2 > 0.f && mn > 0.f) ? (e0 * e1 * e2 / mn) : fmaxf(e0 * e1, fmaxf(e0 * e2, e1 * e2));
    v = sqrtf(area / static_cast<float>(cnt));
  }
  out[t] = v;
}

// ---------------------------------------------------------------------------------------------------------------
cudaError_t launch_grid_spacing(const float4* pnodes, uint32_t n_inner, float* out, cudaStream_t st) {
  if (!n_inner) return cudaSuccess;
  grid_spacing_kernel<<<(n_inner * 2u + 255u) / 256u, 256, 0, st>>>(pnodes, n_inner, out);
  return cudaGetLastError();
}

cudaError_t launch_grid_build(const float4* leaf_pts, uint32_t n, const GridParams& g, uint64_t* keys_in, uint64_t* keys_out, uint32_t* vals_in, uint32_t* vals_out,
                              void* sort_temp, size_t sort_temp_bytes, float4* grid_pts, GridSlot* table, uint32_t capacity, cudaStream_t st) {
  grid_keys_kernel<<<(n + 255u) / 256u, 256, 0, st>>>(leaf_pts, n, g, keys_in, vals_in);
  cudaError_t e = cub::DeviceRadixSort::SortPairs(sort_temp, sort_temp_bytes, keys_in, keys_out, vals_in, vals_out, static_cast<int>(n), 0, 63, st);
  if (e != cudaSuccess) return e;
  grid_table_init_kernel<<<(capacity + 255u) / 256u, 256, 0, st>>>(table, capacity);
  grid_fill_kernel<<<(n + 255u) / 256u, 256, 0, st>>>(keys_out, vals_out, leaf_pts, n, grid_pts, table, capacity - 1u);
  return cudaGetLastError();
}

cudaError_t launch_grid_probe(const LinParams& P, const float4* grid_pts, const GridSlot* table, uint32_t capacity, const GridParams& g, uint8_t* state,
                              uint32_t* pending_count, uint32_t* pending_list, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(pending_count, 0, sizeof(uint32_t), st);
  if (e != cudaSuccess) return e;
  grid_probe_kernel<<<(P.src.n + 255u) / 256u, 256, 0, st>>>(P, grid_pts, table, capacity - 1u, g, state, pending_count, pending_list);
  return cudaGetLastError();
}

cudaError_t launch_pending_search(const LinParams& P, const float4* pnodes, int depth, const uint32_t* pending_count, const uint32_t* pending_list,
                                  uint32_t max_pending, int grid, cudaStream_t st) {
  if (depth < 1) depth = 1;
  const size_t smem = static_cast<size_t>(depth) * kLinBlock * (sizeof(uint2) + sizeof(float));
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(pending_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  pending_search_kernel<<<grid, kLinBlock, smem, st>>>(P, pnodes, depth, pending_count, pending_list, max_pending);
  return cudaGetLastError();
}

}  // namespace sgb
