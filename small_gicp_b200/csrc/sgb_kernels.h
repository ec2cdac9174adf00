// SPDX-License-Identifier: MIT
// Host-visible launchers of the kernels in sgb_kernels.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>

#include "sgb_device.cuh"

namespace sgb {

/// Launch `kernel` as a programmatic dependent of the previous kernel on `st` (its set-up overlaps the predecessor's tail;
/// the kernel itself waits with grid_dependency_wait()).  Profiling build: SGB_PDL=0 falls back to plain stream order (A/B switch).
inline bool pdl_enabled() {
#ifdef SGB_PROFILING
  static const bool on = !(std::getenv("SGB_PDL") && std::getenv("SGB_PDL")[0] == '0');
  return on;
#else
  return true;
#endif
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_dependent(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(static_cast<unsigned>(block));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

cudaError_t launch_linearize(const LinParams& P, int factor, int robust, bool voxel, int grid, int stack_depth, cudaStream_t st);
cudaError_t launch_error(const LinParams& P, int factor, int robust, int grid, cudaStream_t st);
cudaError_t launch_reduce_nothing(const LinParams& P, bool linearize, cudaStream_t st);
int linearize_occupancy(int stack_depth);
cudaError_t set_source_curve(int hilbert);
// sgb_kernels_split.cu
int search_occupancy(int stack_depth);
cudaError_t launch_search(const LinParams& P, int grid, int stack_depth, cudaStream_t st);
cudaError_t launch_factor_reduce(const LinParams& P, int factor, int robust, int grid, cudaStream_t st);
int factor_reduce_occupancy(int factor, int robust);
/// Reduction::error through the factor kernel's operand pipeline (one wave of CTAs looping over tiles)
cudaError_t launch_error_pipelined(const LinParams& P, int factor, int robust, int grid, cudaStream_t st);
int error_pipelined_occupancy(int factor, int robust);

cudaError_t launch_bounds_centre(const double* d_pts4, size_t n, double* d_bounds6, double* d_centre4, int sm_count, cudaStream_t st);
cudaError_t launch_convert(const double* d_pts4, const double* d_normals4, const double* d_covs16, size_t n, const double* d_centre4, float4* out_pts,
                           float4* out_normals, float4* out_covA, float4* out_covB, uint64_t* keys, uint32_t* vals, int sm_count, cudaStream_t st,
                           float4* out_lo = nullptr);
cudaError_t launch_inverse_perm(const uint32_t* perm, size_t n, uint32_t* rank, int sm_count, cudaStream_t st);
cudaError_t launch_convert_cov_scatter(const double* d_covs16, size_t first, size_t n, const uint32_t* rank, float4* outA, float4* outB, float4* origA,
                                       float4* origB, int sm_count, cudaStream_t st);
cudaError_t launch_gather(const uint32_t* perm, size_t n, const float4* in0, float4* out0, const float4* in1, float4* out1, const float4* in2, float4* out2,
                          const float4* in3, float4* out3, int sm_count, cudaStream_t st);
cudaError_t launch_scatter(const uint32_t* perm, size_t n, const float4* in0, float4* out0, const float4* in1, float4* out1, const float4* in2, float4* out2,
                           int sm_count, cudaStream_t st);
cudaError_t launch_correspondences(const uint32_t* corr, const uint32_t* perm, size_t n, const float4* tgt_pts, int voxel, uint64_t* out, int sm_count,
                                   cudaStream_t st);
cudaError_t launch_chunk_transpose(const uint32_t* in, uint32_t* out, size_t n, uint32_t K, int sm_count, cudaStream_t st);
cudaError_t sort_pairs_u64_u32(void* d_temp, size_t& temp_bytes, const uint64_t* keys_in, uint64_t* keys_out, const uint32_t* vals_in, uint32_t* vals_out,
                               size_t n, cudaStream_t st);

/// Work lists of the packet search, filled by the probe (one warp of the probe = one 32-query chunk): chunk ids by cost class
/// (0 = widest search ball ... kChunkClasses - 1 = narrowest), so that the packet search can take the expensive walks first.
constexpr uint32_t kChunkClasses = 3;
struct ChunkClasses {
  uint32_t* count;       // [kChunkClasses] entries of each list (zero on entry: the previous probe cleared them)
  uint32_t* count_next;  // [kChunkClasses] the counters of the NEXT linearize (this probe clears them)
  uint32_t* lists;       // [kChunkClasses][n_chunks] chunk ids, or null: no lists (the packet search scans the settled flags itself)
  uint32_t n_chunks;
  float wide_r2, mid_r2;  // class boundaries on the squared search radius
  uint32_t fallback_pct;  // more than this share of all chunks listed (i.e. with pending lanes): the packet search ignores the lists (curve order)
};
// sgb_kernels_packet.cu
int packet_occupancy(int max_depth);
// sgb_grid.cu: uniform-grid front end of the search
struct GridParams {
  float origin[3];  // minimum corner of the target's box (centred frame)
  float inv_cell;   // 1 / cell edge
  float settle_d2;  // a candidate within this squared distance is provably the nearest neighbour ((cell/2 - margin)^2)
};
struct __align__(16) GridSlot {
  unsigned long long key;  // packed cell coordinates, ~0 = empty
  uint32_t start, count;   // run of the cell's points in the cell-ordered copy of the target
};
/// What the warp-per-query search of few pending queries needs (it runs inside packet_search_kernel; see sgb_grid.cuh)
struct PendingParams {
  const uint32_t* list;    // pending list (source positions)
  const float4* q;         // parallel: transformed query + squared distance of the probe's best candidate
  const float4* grid_pts;  // block lists (pair records, sgb_grid.cuh)
  const GridSlot* table;   // block table, or null: no ring phase
  uint32_t capacity;
  GridParams g;
};
/// queue / queue_next: two alternating zero-initialised counters of the dynamic chunk queue (this launch clears queue_next), or null = static stride.
/// pending_count <= min_pending: the launch finishes the pending queries warp-per-query (ring phase / tree walk); more: the packet walk.
cudaError_t launch_packet_search(const LinParams& P, const float4* pnodes, int grid, int max_depth, const uint8_t* settled, const uint32_t* pending_count,
                                 uint32_t min_pending, uint32_t* queue, uint32_t* queue_next, const ChunkClasses& cc, const PendingParams& pp, bool tma_leaf, cudaStream_t st);
cudaError_t launch_grid_spacing(const float4* pnodes, uint32_t n_inner, float* out, cudaStream_t st);
/// Block lists as PAIR RECORDS (sgb_grid.cuh: GridPair).  sort: keys of the 8 n entries, radix sort, then records per run at the run heads
/// (into vals_in) and d_counters[0] = runs, [1] = longest run.  fill: exclusive scan (pairs -> pair_start), table, records.
/// grid_pts must hold (m + runs) / 2 records of 32 B, i.e. (m + runs) float4.
cudaError_t launch_grid_sort(const float4* leaf_pts, uint32_t n, const GridParams& g, uint64_t* keys_in, uint64_t* keys_out, uint32_t* vals_in, uint32_t* vals_out,
                             void* sort_temp, size_t sort_temp_bytes, uint32_t* d_counters, bool curve, cudaStream_t st);
cudaError_t launch_grid_fill(const uint64_t* keys_sorted, const uint32_t* vals_sorted, const float4* leaf_pts, uint32_t m, const uint32_t* pairs, uint32_t* pair_start,
                             void* scan_temp, size_t scan_temp_bytes, float4* grid_pts, GridSlot* table, uint32_t capacity, bool curve, cudaStream_t st);
/// longest list (points) the grid front end accepts: beyond it a single query's scan would dominate (a cluster of near-duplicate
/// points, or a few far outliers stretching the box so that the cell-count bound inflates the cell) and the exact tree search alone is used
constexpr uint32_t kGridMaxList = 2048;
cudaError_t launch_grid_probe(const LinParams& P, const float4* grid_pts, const GridSlot* table, uint32_t capacity, const GridParams& g, uint8_t* state,
                              uint32_t* pending_count, uint32_t* pending_list, float4* pending_q, uint32_t* next_count, const ChunkClasses& cc, cudaStream_t st);
/// what the 27-block ring is guaranteed to cover: squared radius ((2.5 - slack) cells)^2
inline float ring_cover_sq(float cell) {
  const float c = (2.5f - 4e-3f) * cell;
  return c * c;
}
// device-side tree construction (sgb_kernels.cu)
constexpr uint32_t kLbvhLeafPoints = 32;
cudaError_t launch_curve_keys(const float4* pts, size_t n, const double* centre4, uint64_t* keys, uint32_t* vals, int sm_count, cudaStream_t st);
cudaError_t launch_lbvh_build(const float4* leaf_pts, uint32_t n, uint32_t P, float4* pnodes, int* launches, cudaStream_t st);
cudaError_t launch_kd_level_keys(const float4* cur_pts, uint32_t n, uint32_t count, uint32_t* boxes, uint64_t* keys, cudaStream_t st);
uint32_t kd_smem_first_level(uint32_t n, uint32_t levels);
cudaError_t launch_kd_refine_smem(float4* leaf_pts, uint32_t* perm, uint32_t n, uint32_t level0, uint32_t levels, cudaStream_t st);
cudaError_t sort_pairs_u64_u32_bits(void* d_temp, size_t& temp_bytes, const uint64_t* keys_in, uint64_t* keys_out, const uint32_t* vals_in, uint32_t* vals_out,
                                    size_t n, int end_bit, cudaStream_t st);
// sgb_preprocess.cu
cudaError_t launch_features(const float4* pnodes, const float4* pts, uint32_t n, int k, const double* centre, int mode, float4* out_normals, float4* out_covA,
                            float4* out_covB, double* out_normals_d, double* out_covs_d, int depth, int leaf_order_out, cudaStream_t st,
                            const float4* lo_orig = nullptr);
cudaError_t launch_batch_knn(const float4* pnodes, const float4* pts, const double* queries4, uint32_t n, int k, const double* centre, unsigned long long* out_idx,
                             double* out_d, int depth, cudaStream_t st);
cudaError_t launch_voxel_keys(const double* d_pts4, size_t n, double inv_leaf, uint64_t* keys, uint32_t* vals, int sm_count, cudaStream_t st);
cudaError_t launch_voxel_heads(const uint64_t* keys, size_t n, uint32_t* heads, int sm_count, cudaStream_t st);
cudaError_t launch_voxel_means(const uint64_t* keys, const uint32_t* vals, const uint32_t* heads, const uint32_t* slots, size_t n, const double* d_pts4,
                               double* d_out4, int sm_count, cudaStream_t st);
cudaError_t launch_voxel_stats(const uint64_t* keys, const uint32_t* vals, const uint32_t* heads, const uint32_t* slots, size_t n, const double* d_pts4,
                               const double* d_covs16, double* out_means4, double* out_covs16, int4* out_coords, int sm_count, cudaStream_t st);
cudaError_t launch_vox_table_build(const int4* coords, uint32_t n_voxels, int4* table, uint32_t capacity, cudaStream_t st);
cudaError_t exclusive_sum_u32(void* d_temp, size_t& temp_bytes, const uint32_t* in, uint32_t* out, size_t n, cudaStream_t st);

}  // namespace sgb
