// SPDX-License-Identifier: MIT
// Uniform-grid front end of the exact nearest-neighbour search.
//
// In a registration that is not wildly misaligned almost every query's nearest neighbour is a fraction of the point
// spacing away.  The 2 x 2 x 2 block of grid cells (cell edge c) anchored at floor(u - 1/2) covers the cube
// [q - c/2, q + c/2]^3, i.e. it holds EVERY target point within c/2 of q: if the closest point of the block is within c/2
// it is the exact nearest neighbour and the query is finished -- no tree walk, no dependent pointer chasing.
//   * BLOCK LISTS (default): every target point is stored under the eight blocks that contain its cell, so a block is
//     ONE hash lookup and ONE contiguous run of points (grid_probe_blocks_kernel).  8x the point storage buys the removal
//     of seven lookups and of the divergent per-cell scans.
//   * per-cell lists (SGB_GRID_BLOCKS=0, kept for the A/B in profiles/r01): eight lookups, own cell first, the other
//     cells pruned by their box distance (grid_probe_kernel).
// Queries the probe cannot settle (nothing within c/2: misaligned first iterations, holes, outliers) are appended to a
// compact pending list -- with the best candidate found so far as an upper bound -- and finished exactly:
//   * few of them: a warp per query (pending_search_kernel): ring search over the 27 blocks at stride 2 around the query
//     (everything within 2.5 c), then, only if that radius does not decide it, a walk of the packet tree;
//   * many of them: the packet search (sgb_kernels_packet.cu) over the chunk-ordered queries, settled lanes idle.
// Which of the two runs is decided on the device from the pending counter.  Results are identical to a pure tree search
// (exact ties aside; tests/test_gpu_parity.py::test_search_structures_agree); evidence and A/B runs: profiles/r01.
#include <cfloat>
#include <cstdlib>
#include <cub/device/device_radix_sort.cuh>

#include "sgb_device.cuh"
#include "sgb_grid.cuh"
#include "sgb_kernels.h"

namespace sgb {

// Morton (Z-order) code of the three 21-bit offset coordinates: the ORDER in which the lists are laid out in memory.  The table is keyed
// by the plain packed coordinates (grid_key); only the sort uses the curve, so that the lists a warp of Hilbert-ordered queries needs --
// a compact patch of ~11 neighbouring blocks -- are neighbours in memory too (DRAM pages, L2 sectors, TLB) instead of being strung along
// x only (profiles/r02: probe and the 10M - 100M sizes).
__device__ __forceinline__ uint64_t grid_spread21(uint64_t x) {
  x &= 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}
__device__ __forceinline__ uint64_t grid_compact21(uint64_t x) {
  x &= 0x1249249249249249ull;
  x = (x | x >> 2) & 0x10c30c30c30c30c3ull;
  x = (x | x >> 4) & 0x100f00f00f00f00full;
  x = (x | x >> 8) & 0x1f0000ff0000ffull;
  x = (x | x >> 16) & 0x1f00000000ffffull;
  x = (x | x >> 32) & 0x1fffffull;
  return x;
}
__device__ __forceinline__ uint64_t grid_order_key(int ix, int iy, int iz, bool curve) {
  if (!curve) return grid_key(ix, iy, iz);
  return grid_spread21(static_cast<uint32_t>(ix + (1 << 20))) | (grid_spread21(static_cast<uint32_t>(iy + (1 << 20))) << 1) |
         (grid_spread21(static_cast<uint32_t>(iz + (1 << 20))) << 2);
}
/// sort key -> table key (packed coordinates)
__device__ __forceinline__ uint64_t grid_table_key(uint64_t k, bool curve) {
  if (!curve) return k;
  return grid_compact21(k) | (grid_compact21(k >> 1) << 21) | (grid_compact21(k >> 2) << 42);
}

// cell key of every target point (leaf-ordered, centred FP32); value = leaf position.
// BLOCKS: every point is entered under the eight 2 x 2 x 2 cell blocks that contain its cell (block anchor = lowest cell),
// so that a query finds every point of the block around it in ONE contiguous run (8x the points, one lookup, one loop).
template <bool BLOCKS>
__global__ void grid_keys_kernel(const float4* __restrict__ pts, uint32_t n, GridParams g, uint64_t* keys, uint32_t* vals, bool curve) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t i = BLOCKS ? t >> 3 : t, o = BLOCKS ? t & 7u : 0u;
  if (i >= n) return;
  const float4 p = pts[i];
  const int ix = __float2int_rd((p.x - g.origin[0]) * g.inv_cell), iy = __float2int_rd((p.y - g.origin[1]) * g.inv_cell),
            iz = __float2int_rd((p.z - g.origin[2]) * g.inv_cell);
  keys[t] = grid_order_key(ix - static_cast<int>(o & 1u), iy - static_cast<int>((o >> 1) & 1u), iz - static_cast<int>(o >> 2), curve);
  vals[t] = i;
}

// number of distinct keys of the sorted key array (sizes the hash table)
__global__ void grid_count_heads_kernel(const uint64_t* __restrict__ keys, uint32_t m, uint32_t* count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool head = i < m && (i == 0 || keys[i - 1] != keys[i]);
  const unsigned b = __ballot_sync(0xffffffffu, head);
  if ((threadIdx.x & 31u) == 0 && b) atomicAdd(count, static_cast<uint32_t>(__popc(b)));
}

// after the sort: gather the points into cell order (w = leaf position) and insert one table entry per run of equal keys
__global__ void grid_fill_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ leaf_pos, const float4* __restrict__ leaf_pts, uint32_t n,
                                 float4* grid_pts, GridSlot* table, uint32_t mask, uint32_t* max_count, bool curve) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t lp = leaf_pos[i];
  const float4 p = leaf_pts[lp];
  grid_pts[i] = make_float4(p.x, p.y, p.z, __uint_as_float(lp));
  const uint64_t k = keys[i];
  if (i == 0 || keys[i - 1] != k) {  // head of a cell: count its points, claim a slot
    uint32_t cnt = 1;
    while (i + cnt < n && keys[i + cnt] == k) cnt++;
    const uint64_t tk = grid_table_key(k, curve);
    uint32_t slot = grid_hash_of_key(tk) & mask;
    for (;;) {
      const unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(&table[slot].key), ~0ull, static_cast<unsigned long long>(tk));
      if (prev == ~0ull) break;
      slot = (slot + 1u) & mask;
    }
    table[slot].start = i;
    table[slot].count = cnt;
    atomicMax(max_count, cnt);  // longest list: build_grid drops the front end when one query would have to scan thousands of points
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

__device__ __forceinline__ void grid_scan(const float4* __restrict__ cp, uint32_t cnt, float fx, float fy, float fz, float& best_d, uint32_t& best) {
#pragma unroll 4
  for (uint32_t j = 0; j < cnt; j++) {
    const float4 t = __ldg(&cp[j]);
    const float dx = t.x - fx, dy = t.y - fy, dz = t.z - fz;
    const float d = dx * dx + dy * dy + dz * dz;
    if (d < best_d) {
      best_d = d;
      best = __float_as_uint(t.w);
    }
  }
}

#ifdef SGB_PROFILING  // per-cell lists, eight lookups per query (SGB_GRID_BLOCKS=0): superseded by the block lists, kept for A/B runs only
// ---------------------------------------------------------------------------------------------------------------
// probe: one query per thread (Hilbert order -> neighbouring lanes hit neighbouring cells)
// state[i] = 1: settled (corr[i] is the exact nearest neighbour), 0: pending for the tree search (corr[i] = best candidate)
// ---------------------------------------------------------------------------------------------------------------
template <int MIN_CTAS>
__global__ void __launch_bounds__(256, MIN_CTAS) grid_probe_kernel(const __grid_constant__ LinParams P, const float4* __restrict__ grid_pts,
                                                           const GridSlot* __restrict__ table, uint32_t mask, GridParams g, float cell_sq, uint8_t* state,
                                                           uint32_t* pending_count, uint32_t* pending_list, float4* pending_q, uint32_t* next_count) {
  const uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = gi < P.src.n;
  if (gi == 0u) *next_count = 0u;  // the counter of the NEXT linearize (two counters alternate: no memset between launches)
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
  // own cell, and per axis the neighbour on the side the query leans to: together the 2 x 2 x 2 block around [q - c/2, q + c/2]^3
  const float ux = (fx - g.origin[0]) * g.inv_cell, uy = (fy - g.origin[1]) * g.inv_cell, uz = (fz - g.origin[2]) * g.inv_cell;
  const float flx = floorf(ux), fly = floorf(uy), flz = floorf(uz);
  // queries far outside the target's box cannot be settled anyway; clamping keeps the integer cell coordinates in range
  const int ox = static_cast<int>(fminf(fmaxf(flx, -1e5f), 1e5f)), oy = static_cast<int>(fminf(fmaxf(fly, -1e5f), 1e5f)),
            oz = static_cast<int>(fminf(fmaxf(flz, -1e5f), 1e5f));
  const float frx = ux - flx, fry = uy - fly, frz = uz - flz;
  const int nx = frx >= 0.5f ? 1 : -1, ny = fry >= 0.5f ? 1 : -1, nz = frz >= 0.5f ? 1 : -1;
  // squared distance (cell units, less the rounding slack) to the neighbour cell's slab along each axis
  const float ax = fmaxf((frx >= 0.5f ? 1.0f - frx : frx) - kGridSlack, 0.0f), ay = fmaxf((fry >= 0.5f ? 1.0f - fry : fry) - kGridSlack, 0.0f),
              az = fmaxf((frz >= 0.5f ? 1.0f - frz : frz) - kGridSlack, 0.0f);
  const float ax2 = ax * ax * cell_sq, ay2 = ay * ay * cell_sq, az2 = az * az * cell_sq;
  // all eight table lookups first: eight independent loads in flight, no divergence (the first slot of each probe
  // sequence is fetched unconditionally; collisions are resolved below)
  uint4 ent[8];
#pragma unroll
  for (int c = 0; c < 8; c++) {
    const int ix = ox + ((c & 1) ? nx : 0), iy = oy + ((c & 2) ? ny : 0), iz = oz + ((c & 4) ? nz : 0);
    ent[c] = __ldg(reinterpret_cast<const uint4*>(&table[grid_hash(ix, iy, iz) & mask]));
  }
#pragma unroll
  for (int c = 0; c < 8; c++) {
    // the own cell first, then the neighbours whose box is still closer than the best distance so far
    const float bd = ((c & 1) ? ax2 : 0.f) + ((c & 2) ? ay2 : 0.f) + ((c & 4) ? az2 : 0.f);
    if (!(bd < best_d)) continue;
    const int ix = ox + ((c & 1) ? nx : 0), iy = oy + ((c & 2) ? ny : 0), iz = oz + ((c & 4) ? nz : 0);
    const uint64_t key = grid_key(ix, iy, iz);
    uint4 e = ent[c];
    uint64_t ek = static_cast<uint64_t>(e.x) | (static_cast<uint64_t>(e.y) << 32);
    uint32_t slot = grid_hash(ix, iy, iz) & mask;
    while (ek != key && ek != ~0ull) {  // rare: the table is sparsely filled
      slot = (slot + 1u) & mask;
      e = __ldg(reinterpret_cast<const uint4*>(&table[slot]));
      ek = static_cast<uint64_t>(e.x) | (static_cast<uint64_t>(e.y) << 32);
    }
    if (ek == key) grid_scan(grid_pts + e.z, e.w, fx, fy, fz, best_d, best);
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
  // pending queries go to a compact work list (warp-aggregated append)
  const bool pend = in_range && !settled;
  const unsigned m = __ballot_sync(0xffffffffu, pend);
  if (m) {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t base = 0;
    if (lane == static_cast<uint32_t>(__ffs(m) - 1)) base = atomicAdd(pending_count, static_cast<uint32_t>(__popc(m)));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (pend) {
      const uint32_t slot = base + __popc(m & ((1u << lane) - 1u));
      pending_list[slot] = i;
      pending_q[slot] = make_float4(fx, fy, fz, (best == kNone || !(best_d < P.max_dist_sq)) ? P.max_dist_sq : best_d);
    }
  }
}

#endif  // SGB_PROFILING

// ---------------------------------------------------------------------------------------------------------------
// probe, block lists: the block anchored at floor(u - 1/2) covers [q - c/2, q + c/2]^3; one lookup, one contiguous scan
// ---------------------------------------------------------------------------------------------------------------
// BATCH_TAIL (A/B switch SGB_PROBE_TAIL=1; measured once at the end of round 1: no gain, stays off -- profiles/r01/am_linearize.md): the plain loop below is unrolled by 8, which leaves a
// remainder loop of count % 8 iterations with ONE load in flight each -- 15 % of the kernel's stall samples sit on that
// load's first use (profiles/r01/am).  The batched form always issues eight loads, clamping the index to the last point of
// the list: a repeated point can never be strictly closer than itself, so the result is unchanged.
template <int MIN_CTAS, bool BATCH_TAIL>
__global__ void __launch_bounds__(256, MIN_CTAS) grid_probe_blocks_kernel(const __grid_constant__ LinParams P, const float4* __restrict__ grid_pts,
                                                               const GridSlot* __restrict__ table, uint32_t mask, GridParams g, uint8_t* state,
                                                               uint32_t* pending_count, uint32_t* pending_list, float4* pending_q, uint32_t* next_count,
                                                               ChunkClasses cc) {
  const uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = gi < P.src.n;
  if (gi == 0u) *next_count = 0u;  // the counter of the NEXT linearize (two counters alternate: no memset between launches)
  if (gi < kChunkClasses && cc.count_next) cc.count_next[gi] = 0u;
  const uint32_t i = in_range ? gi : P.src.n - 1u;
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
  const float ax = floorf((fx - g.origin[0]) * g.inv_cell - 0.5f), ay = floorf((fy - g.origin[1]) * g.inv_cell - 0.5f),
              az = floorf((fz - g.origin[2]) * g.inv_cell - 0.5f);
  // queries far outside the target's box cannot be settled anyway; clamping keeps the integer cell coordinates in range
  const int bx = static_cast<int>(fminf(fmaxf(ax, -1e5f), 1e5f)), by = static_cast<int>(fminf(fmaxf(ay, -1e5f), 1e5f)),
            bz = static_cast<int>(fminf(fmaxf(az, -1e5f), 1e5f));
  const uint2 e = grid_lookup(table, mask, bx, by, bz);
  {
    // the list is a contiguous run of e.y points: pull all of its lines into L2 at once (the scan below would otherwise
    // pay one DRAM round trip per batch of loads: profiles/r01/q, 29 warps stalled on the scoreboard per issue)
    const char* lp = reinterpret_cast<const char*>(grid_pts + e.x);
    const uint32_t bytes = e.y * 16u;
    for (uint32_t off = 128u; off < bytes; off += 128u) asm volatile("prefetch.global.L2 [%0];" ::"l"(lp + off));
    if (bytes > 128u) asm volatile("prefetch.global.L2 [%0];" ::"l"(lp + bytes - 16u));  // last line of a run that is not 128 B aligned
    const float4* __restrict__ cp = grid_pts + e.x;
    if (BATCH_TAIL) {
      const uint32_t last = e.y - 1u;  // only used when e.y > 0
      for (uint32_t base = 0; base < e.y; base += 8u) {
        float4 t[8];
#pragma unroll
        for (int u = 0; u < 8; u++) t[u] = __ldg(&cp[min(base + static_cast<uint32_t>(u), last)]);
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const float dx = t[u].x - fx, dy = t[u].y - fy, dz = t[u].z - fz;
          const float d = dx * dx + dy * dy + dz * dz;
          if (d < best_d) {
            best_d = d;
            best = __float_as_uint(t[u].w);
          }
        }
      }
    } else {
#pragma unroll 8
      for (uint32_t j = 0; j < e.y; j++) {
        const float4 t = __ldg(&cp[j]);
        const float dx = t.x - fx, dy = t.y - fy, dz = t.z - fz;
        const float d = dx * dx + dy * dy + dz * dz;
        if (d < best_d) {
          best_d = d;
          best = __float_as_uint(t.w);
        }
      }
    }
  }
  const bool settled = best != kNone && best_d <= g.settle_d2;
  if (!settled && P.use_prev) {  // keep the better of (grid candidate, previous correspondence) as the tree search's seed
    const uint32_t prev = P.corr[i];
    if (prev != kNone) {
      const float4 t = __ldg(&P.tgt.pts[prev]);
      const float dx = t.x - fx, dy = t.y - fy, dz = t.z - fz;
      const float d = dx * dx + dy * dy + dz * dz;
      if (d < best_d) {  // (a seed beyond the rejector's bound is no use: best_d starts there)
        best = prev;
        best_d = d;
      }
    }
  }
  if (in_range) {
    P.corr[i] = best;
    state[i] = settled ? 1 : 0;
  }
  // Pending queries go to a compact list, and -- for the packet search -- the CHUNK (one warp here = one 32-query chunk there) goes into
  // the work list of its cost class = how far its widest search ball reaches (best_d is the squared radius a pending lane starts its
  // walk with).  The packet search takes the wide classes first, so that what is left when the GPU drains are the cheap walks
  // (profiles/r01/ai: a third of that kernel was a tail of late-starting heavy chunks); chunks without a pending lane are in no list.
  // Both appends are aggregated over the CTA's eight warps: at most four atomics per CTA -- one per warp and list made the hot
  // counters the probe's own tail (r02d: +9 us at the identity pose).
  const bool pend = in_range && !settled;
  const unsigned m = __ballot_sync(0xffffffffu, pend);
  const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  __shared__ uint32_t s_cnt[8], s_cls[8], s_base[8];
  {
    const uint32_t r2 = __reduce_max_sync(0xffffffffu, pend ? __float_as_uint(best_d) : 0u);  // non-negative floats order like their bits
    if (lane == 0) {
      const float r = __uint_as_float(r2);
      s_cnt[wib] = static_cast<uint32_t>(__popc(m));
      s_cls[wib] = !m ? kChunkClasses : (r >= cc.wide_r2 ? 0u : (r >= cc.mid_r2 ? 1u : 2u));
    }
  }
  __syncthreads();
  if (wib == 0) {
    const uint32_t nw = blockDim.x >> 5;
    const uint32_t cnt = lane < nw ? s_cnt[lane] : 0u, cls = lane < nw ? s_cls[lane] : kChunkClasses;
    uint32_t incl = cnt;  // inclusive prefix over the warps
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= static_cast<uint32_t>(o)) incl += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 7);
    uint32_t pbase = 0;
    if (lane == 0 && total) pbase = atomicAdd(pending_count, total);
    pbase = __shfl_sync(0xffffffffu, pbase, 0);
    if (lane < nw) s_base[lane] = pbase + incl - cnt;
    if (cc.lists) {
#pragma unroll
      for (uint32_t c = 0; c < kChunkClasses; c++) {
        const unsigned mc = __ballot_sync(0xffffffffu, cls == c);
        if (!mc) continue;
        uint32_t cb = 0;
        if (lane == static_cast<uint32_t>(__ffs(mc) - 1)) cb = atomicAdd(&cc.count[c], static_cast<uint32_t>(__popc(mc)));
        cb = __shfl_sync(0xffffffffu, cb, __ffs(mc) - 1);
        if (cls == c) cc.lists[static_cast<size_t>(c) * cc.n_chunks + cb + __popc(mc & ((1u << lane) - 1u))] = (blockIdx.x * blockDim.x >> 5) + lane;
      }
    }
  }
  __syncthreads();
  if (pend) {  // the pending record: who, where (transformed query) and how far its best candidate is -- the finishing kernel needs nothing else
    const uint32_t slot = s_base[wib] + __popc(m & ((1u << lane) - 1u));
    pending_list[slot] = i;
    pending_q[slot] = make_float4(fx, fy, fz, best_d);
  }
}

#ifdef SGB_PROFILING
// ---------------------------------------------------------------------------------------------------------------
// Many pending queries AND a rejector whose radius the ring covers (DistanceRejector(1 m) with the usual cell: 2.5 c = 1.1 m): one THREAD
// per pending query finishes it exactly from the block lists alone -- the 26 blocks at stride 2 around its own one (which the probe has
// already scanned), each skipped when its box lies beyond the best distance so far.  Everything within sqrt(max_dist_sq) of the query is
// inside those 27 blocks, so whatever is nearest there is the exact nearest neighbour, or provably nothing is in range.  The idea: trade
// the packet walk's ~3,350 warp instructions per 32-query chunk for per-thread scans as wide as each query's own search ball.
// MEASURED AND REJECTED (r02m, SGB_RING_SCAN=1): 0.417 vs 0.241 ms at the identity pose, 0.275 vs 0.181 ms at T1 -- up to 26 dependent
// lookup + scan round trips per thread with neighbouring lanes needing different blocks is far slower than one shared, divergence-free
// walk.  Exact (test_search_structures_agree: device-kd/grid-ring-scan-pending); profiling library only.
// ---------------------------------------------------------------------------------------------------------------
// the 26 neighbours of the centre block, faces first, then edges, then corners; ox | oy << 2 | oz << 4 with o in {0, 1, 2} = {-1, 0, +1}
__constant__ unsigned char kRingOrder[26] = {
  0x14, 0x16, 0x11, 0x19, 0x05, 0x25,                                                  // faces:  x-, x+, y-, y+, z-, z+
  0x10, 0x12, 0x18, 0x1a, 0x04, 0x06, 0x24, 0x26, 0x01, 0x09, 0x21, 0x29,            // edges:  xy (4), xz (4), yz (4)
  0x00, 0x02, 0x08, 0x0a, 0x20, 0x22, 0x28, 0x2a};                                    // corners
__global__ void __launch_bounds__(256, 4) ring_scan_kernel(const __grid_constant__ LinParams P, const uint32_t* __restrict__ pending_count,
                                                           const uint32_t* __restrict__ pending_list, uint32_t min_pending, const float4* __restrict__ grid_pts,
                                                           const GridSlot* __restrict__ table, uint32_t mask, GridParams g, float cell) {
  grid_dependency_wait();  // probe (and the warp-per-query kernel, which exits at once in this regime) wrote corr[] and the pending list
  const uint32_t count = *pending_count;
  if (count <= min_pending) return;
  const double* R = P.T;
  const double csx = P.src.centre[0], csy = P.src.centre[1], csz = P.src.centre[2];
  const double tpx = R[0] * csx + R[1] * csy + R[2] * csz + P.T[9] - P.tgt.centre[0];
  const double tpy = R[3] * csx + R[4] * csy + R[5] * csz + P.T[10] - P.tgt.centre[1];
  const double tpz = R[6] * csx + R[7] * csy + R[8] * csz + P.T[11] - P.tgt.centre[2];
  const float cell_sq = cell * cell;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
    const uint32_t i = pending_list[k];
    const float4 s = __ldg(&P.src.pts[i]);
    const double sx = s.x, sy = s.y, sz = s.z;
    const float qx = static_cast<float>(R[0] * sx + R[1] * sy + R[2] * sz + tpx);
    const float qy = static_cast<float>(R[3] * sx + R[4] * sy + R[5] * sz + tpy);
    const float qz = static_cast<float>(R[6] * sx + R[7] * sy + R[8] * sz + tpz);
    float best_d = P.max_dist_sq;
    uint32_t best = P.corr[i];  // the probe's candidate (own block, or the previous correspondence): an upper bound
    if (best != kNone) {
      const float4 t = __ldg(&P.tgt.pts[best]);
      const float dx = t.x - qx, dy = t.y - qy, dz = t.z - qz;
      const float d = dx * dx + dy * dy + dz * dz;
      if (d < best_d) best_d = d;
      else best = kNone;  // beyond the rejector's bound: no use
    }
    const float ux = (qx - g.origin[0]) * g.inv_cell, uy = (qy - g.origin[1]) * g.inv_cell, uz = (qz - g.origin[2]) * g.inv_cell;
    const int ax = static_cast<int>(fminf(fmaxf(floorf(ux - 0.5f), -1e5f), 1e5f)), ay = static_cast<int>(fminf(fmaxf(floorf(uy - 0.5f), -1e5f), 1e5f)),
              az = static_cast<int>(fminf(fmaxf(floorf(uz - 0.5f), -1e5f), 1e5f));
    // per axis: squared distance (cell units, less the rounding slack) from the query to the lower / own / upper block slab
    // (the query sits inside its own slab, so that distance is zero; scalars, not arrays: dynamic indexing would go to local memory)
    auto slab = [&](float u, int a, int o) {
      const float l = static_cast<float>(a + 2 * (o - 1));
      const float f = fmaxf(fmaxf(l - u, u - (l + 2.0f)) - kGridSlack, 0.0f);
      return f * f * cell_sq;
    };
    const float ex0 = slab(ux, ax, 0), ex2 = slab(ux, ax, 2), ey0 = slab(uy, ay, 0), ey2 = slab(uy, ay, 2), ez0 = slab(uz, az, 0), ez2 = slab(uz, az, 2);
    // nearer blocks first: faces, then edges, then corners of the 3 x 3 x 3 arrangement (the centre is the probe's own block)
    {
#pragma unroll 1
      for (int b = 0; b < 26; b++) {
        const int code = kRingOrder[b], ox = code & 3, oy = (code >> 2) & 3, oz = code >> 4;
        const float bd = (ox == 0 ? ex0 : (ox == 2 ? ex2 : 0.0f)) + (oy == 0 ? ey0 : (oy == 2 ? ey2 : 0.0f)) + (oz == 0 ? ez0 : (oz == 2 ? ez2 : 0.0f));
        if (!(bd < best_d)) continue;
        const uint2 e = grid_lookup(table, mask, ax + 2 * (ox - 1), ay + 2 * (oy - 1), az + 2 * (oz - 1));
        const float4* __restrict__ cp = grid_pts + e.x;
#pragma unroll 4
        for (uint32_t j = 0; j < e.y; j++) {
          const float4 t = __ldg(&cp[j]);
          const float dx = t.x - qx, dy = t.y - qy, dz = t.z - qz;
          const float d = dx * dx + dy * dy + dz * dz;
          if (d < best_d) {
            best_d = d;
            best = __float_as_uint(t.w);
          }
        }
      }
    }
    P.corr[i] = best;
  }
}

#endif  // SGB_PROFILING

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

cudaError_t launch_grid_sort(const float4* leaf_pts, uint32_t n, const GridParams& g, bool blocks, uint64_t* keys_in, uint64_t* keys_out, uint32_t* vals_in,
                             uint32_t* vals_out, void* sort_temp, size_t sort_temp_bytes, uint32_t* d_distinct, bool curve, cudaStream_t st) {
  const uint32_t m = blocks ? n * 8u : n;
  if (blocks)
    grid_keys_kernel<true><<<(m + 255u) / 256u, 256, 0, st>>>(leaf_pts, n, g, keys_in, vals_in, curve);
  else
    grid_keys_kernel<false><<<(m + 255u) / 256u, 256, 0, st>>>(leaf_pts, n, g, keys_in, vals_in, curve);
  cudaError_t e = cub::DeviceRadixSort::SortPairs(sort_temp, sort_temp_bytes, keys_in, keys_out, vals_in, vals_out, static_cast<int>(m), 0, 63, st);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(d_distinct, 0, sizeof(uint32_t), st);
  if (e != cudaSuccess) return e;
  grid_count_heads_kernel<<<(m + 255u) / 256u, 256, 0, st>>>(keys_out, m, d_distinct);
  return cudaGetLastError();
}

cudaError_t launch_grid_fill(const uint64_t* keys_sorted, const uint32_t* vals_sorted, const float4* leaf_pts, uint32_t m, float4* grid_pts, GridSlot* table,
                             uint32_t capacity, uint32_t* d_max_count, bool curve, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(d_max_count, 0, sizeof(uint32_t), st);
  if (e != cudaSuccess) return e;
  grid_table_init_kernel<<<(capacity + 255u) / 256u, 256, 0, st>>>(table, capacity);
  grid_fill_kernel<<<(m + 255u) / 256u, 256, 0, st>>>(keys_sorted, vals_sorted, leaf_pts, m, grid_pts, table, capacity - 1u, d_max_count, curve);
  return cudaGetLastError();
}

cudaError_t launch_grid_probe(const LinParams& P, const float4* grid_pts, const GridSlot* table, uint32_t capacity, const GridParams& g, bool blocks,
                              bool batch_tail, uint8_t* state, uint32_t* pending_count, uint32_t* pending_list, float4* pending_q, uint32_t* next_count,
                              const ChunkClasses& cc, cudaStream_t st) {
  // *pending_count must be zero on entry: the previous probe (or the context) cleared it
  const float cell = 1.0f / g.inv_cell;
  const uint32_t grid = (P.src.n + 255u) / 256u;
#ifdef SGB_PROFILING
  if (!blocks) {
    grid_probe_kernel<4><<<grid, 256, 0, st>>>(P, grid_pts, table, capacity - 1u, g, cell * cell, state, pending_count, pending_list, pending_q, next_count);
    return cudaGetLastError();
  }
  static const int ctas = std::getenv("SGB_PROBE_CTAS") ? std::atoi(std::getenv("SGB_PROBE_CTAS")) : 5;
  if (batch_tail) {
    grid_probe_blocks_kernel<5, true><<<grid, 256, 0, st>>>(P, grid_pts, table, capacity - 1u, g, state, pending_count, pending_list, pending_q, next_count, cc);
    return cudaGetLastError();
  }
  if (ctas == 6) {
    grid_probe_blocks_kernel<6, false><<<grid, 256, 0, st>>>(P, grid_pts, table, capacity - 1u, g, state, pending_count, pending_list, pending_q, next_count, cc);
    return cudaGetLastError();
  }
  if (ctas == 8) {
    grid_probe_blocks_kernel<8, false><<<grid, 256, 0, st>>>(P, grid_pts, table, capacity - 1u, g, state, pending_count, pending_list, pending_q, next_count, cc);
    return cudaGetLastError();
  }
#else
  (void)blocks;
  (void)batch_tail;
  (void)cell;
#endif
  grid_probe_blocks_kernel<5, false><<<grid, 256, 0, st>>>(P, grid_pts, table, capacity - 1u, g, state, pending_count, pending_list, pending_q, next_count, cc);
  return cudaGetLastError();
}

cudaError_t launch_ring_scan(const LinParams& P, const uint32_t* pending_count, const uint32_t* pending_list, uint32_t min_pending, const float4* grid_pts,
                             const GridSlot* block_table, uint32_t capacity, const GridParams& g, int grid, cudaStream_t st) {
#ifndef SGB_PROFILING
  return cudaErrorNotSupported;
#else
  return launch_dependent(ring_scan_kernel, grid, 256, 0, st, P, pending_count, pending_list, min_pending, grid_pts, block_table, capacity - 1u, g, 1.0f / g.inv_cell);
#endif
}

}  // namespace sgb
