// SPDX-License-Identifier: MIT
// C-ABI entry points of the per-cloud preparation steps: normal / covariance estimation and
// voxel-grid down-sampling on the device.  Kernels: sgb_preprocess.cu.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include "../../include/sgicp_b200.h"
#include "sgb_context.hpp"
#include "sgb_kdtree_host.hpp"
#include "sgb_kernels.h"

using namespace sgb;

namespace sgb {

/// Device-side construction of the search structure over `n` centred FP32 points (original order, w = index):
/// Hilbert keys -> radix sort -> leaf-ordered points -> leaf boxes -> one kernel per level of the implicit binary tree.
/// No host round trip, no synchronisation.  Outputs: perm (leaf order -> original index), leaf_pts, pnodes, depth.
int build_lbvh(sgb_ctx* ctx, const float4* d_orig_pts, size_t n, const double* d_centre4, DevBuf& perm, DevBuf& leaf_pts, DevBuf& pnodes, int* depth) {
  if (n >= (1ull << 30)) return fail(ctx, 1, "too many points for the device tree");
  // P leaf slots (power of two), leaf j = positions [floor(j n / P), floor((j+1) n / P)): at most 32 points each
  uint32_t P = 2;
  int d = 1;
  while (static_cast<uint64_t>(P) * kLbvhLeafPoints < n) {
    P <<= 1;
    d++;
  }
  CU(ctx->keys_in.reserve(n * sizeof(uint64_t)));
  CU(ctx->keys_out.reserve(n * sizeof(uint64_t)));
  CU(ctx->vals_in.reserve(n * sizeof(uint32_t)));
  CU(perm.reserve(n * sizeof(uint32_t)));
  CU(leaf_pts.reserve(n * sizeof(float4)));
  CU(pnodes.reserve(static_cast<size_t>(P - 1) * 64));
  CU(launch_curve_keys(d_orig_pts, n, d_centre4, ctx->keys_in.as<uint64_t>(), ctx->vals_in.as<uint32_t>(), ctx->sm_count, ctx->stream));
  size_t temp_bytes = 0;
  CU(sort_pairs_u64_u32(nullptr, temp_bytes, ctx->keys_in.as<uint64_t>(), ctx->keys_out.as<uint64_t>(), ctx->vals_in.as<uint32_t>(), perm.as<uint32_t>(), n,
                        ctx->stream));
  CU(ctx->sort_temp.reserve(temp_bytes));
  CU(sort_pairs_u64_u32(ctx->sort_temp.p, temp_bytes, ctx->keys_in.as<uint64_t>(), ctx->keys_out.as<uint64_t>(), ctx->vals_in.as<uint32_t>(),
                        perm.as<uint32_t>(), n, ctx->stream));
  CU(launch_gather(perm.as<uint32_t>(), n, d_orig_pts, leaf_pts.as<float4>(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, ctx->sm_count, ctx->stream));
  ctx->launches += 5;
  // (Skipping this refinement for small clouds -- its d radix sorts are pure launch latency at 15k points -- was measured and rejected:
  // the k = 20 searches of the covariance estimation in the unrefined tree cost more than the sorts save, r02j: 6.8 vs 4.1 ms per frame.)
  if (ctx->tree_quality >= 1) {
    // median-split refinement: level by level, sort every node's points along the widest axis of its box; the balanced
    // position boundaries of the next level are then exactly the median splits of a kd-tree (same quality as the host
    // builder, ~d radix sorts instead of a D2H + recursive host build + H2D)
    CU(ctx->pre_boxes.reserve(static_cast<size_t>(P) * 6 * sizeof(uint32_t)));
    // top levels (nodes larger than a shared-memory segment): one radix sort per level; every subtree below is refined by ONE launch
    const int level0 = ctx->kd_smem_refine ? static_cast<int>(kd_smem_first_level(static_cast<uint32_t>(n), static_cast<uint32_t>(d))) : d;
    for (int level = 0; level < level0; level++) {
      const uint32_t count = 1u << level;
      CU(launch_kd_level_keys(leaf_pts.as<float4>(), static_cast<uint32_t>(n), count, ctx->pre_boxes.as<uint32_t>(), ctx->keys_in.as<uint64_t>(), ctx->stream));
      CU(cudaMemcpyAsync(ctx->vals_in.p, perm.p, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
      size_t tb = 0;
      CU(sort_pairs_u64_u32_bits(nullptr, tb, ctx->keys_in.as<uint64_t>(), ctx->keys_out.as<uint64_t>(), ctx->vals_in.as<uint32_t>(), perm.as<uint32_t>(), n,
                                 32 + level, ctx->stream));
      CU(ctx->sort_temp.reserve(tb));
      CU(sort_pairs_u64_u32_bits(ctx->sort_temp.p, tb, ctx->keys_in.as<uint64_t>(), ctx->keys_out.as<uint64_t>(), ctx->vals_in.as<uint32_t>(),
                                 perm.as<uint32_t>(), n, 32 + level, ctx->stream));
      CU(launch_gather(perm.as<uint32_t>(), n, d_orig_pts, leaf_pts.as<float4>(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, ctx->sm_count,
                       ctx->stream));
      ctx->launches += 8;
    }
    if (level0 < d) {
      CU(launch_kd_refine_smem(leaf_pts.as<float4>(), perm.as<uint32_t>(), static_cast<uint32_t>(n), static_cast<uint32_t>(level0), static_cast<uint32_t>(d), ctx->stream));
      ctx->launches += 1;
    }
  }
  int launches = 0;
  CU(launch_lbvh_build(leaf_pts.as<float4>(), static_cast<uint32_t>(n), P, pnodes.as<float4>(), &launches, ctx->stream));
  ctx->launches += launches;
  *depth = d + 1;
  return 0;
}

/// Uniform grid over the leaf-ordered target points.  Cell edge = 3 x the median point spacing of the tree's leaves, so the
/// settle radius (half a cell) is ~3 mean nearest-neighbour distances of a surface sampled like the target.
int build_grid(sgb_ctx* ctx) {
  ctx->grid_ready = false;
  const size_t n = ctx->n_tgt;
  if (!ctx->use_grid || ctx->search_mode != 2 || ctx->tgt_is_voxel || n < 1024 || ctx->n_pnodes == 0) return 0;
  // spacing estimate from the leaf boxes + the target box (one small D2H, construction is a per-target one-off)
  const size_t m = ctx->n_pnodes * 2;
  CU(ctx->grid_spacing.reserve(m * sizeof(float)));
  CU(launch_grid_spacing(ctx->tgt_pnodes.as<float4>(), static_cast<uint32_t>(ctx->n_pnodes), ctx->grid_spacing.as<float>(), ctx->stream));
  std::vector<float> sp(m);
  double bounds[6], centre[4];
  CU(cudaMemcpyAsync(sp.data(), ctx->grid_spacing.p, m * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(bounds, ctx->tgt_bounds.p, sizeof(bounds), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(centre, ctx->tgt_centre.p, sizeof(centre), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  size_t k = 0;
  for (size_t i = 0; i < m; i++)
    if (sp[i] > 0.f) sp[k++] = sp[i];
  if (k == 0) return 0;
  std::nth_element(sp.begin(), sp.begin() + k / 2, sp.begin() + k);
  double cell = ctx->grid_cell_factor * sp[k / 2];
  double ext = 0.0;
  for (int a = 0; a < 3; a++) ext = std::max(ext, bounds[3 + a] - bounds[a]);
  if (!(cell > 0.0) || !(ext > 0.0)) return 0;
  cell = std::max(cell, ext / 8000.0);  // cell indices < 2^13: FP32 cell coordinates stay accurate to 1e-3 cells (kGridSlack in sgb_grid.cu)
  GridParams g;
  for (int a = 0; a < 3; a++) g.origin[a] = static_cast<float>(bounds[a] - centre[a]);
  g.inv_cell = static_cast<float>(1.0 / cell);
  const double settle = (0.5 - 4e-3) * cell;  // margin for the FP32 cell-coordinate arithmetic (kGridSlack cells, sgb_grid.cu)
  g.settle_d2 = static_cast<float>(settle * settle);
  const size_t n_ent = n * 8;  // block lists hold every point under its eight enclosing 2 x 2 x 2 blocks
  if (n_ent >= (1ull << 31)) return 0;
  CU(ctx->keys_in.reserve(n_ent * sizeof(uint64_t)));
  CU(ctx->keys_out.reserve(n_ent * sizeof(uint64_t)));
  CU(ctx->vals_in.reserve(n_ent * sizeof(uint32_t)));
  CU(ctx->pre_vals_out.reserve(n_ent * sizeof(uint32_t)));
  CU(ctx->grid_pending.reserve(2 * sizeof(uint32_t)));
  size_t tb = 0;
  CU(sort_pairs_u64_u32(nullptr, tb, ctx->keys_in.as<uint64_t>(), ctx->keys_out.as<uint64_t>(), ctx->vals_in.as<uint32_t>(), ctx->pre_vals_out.as<uint32_t>(), n_ent,
                        ctx->stream));
  CU(ctx->sort_temp.reserve(tb));
  uint32_t* d_distinct = ctx->grid_pending.as<uint32_t>();  // the two pending counters double as scratch during construction
  CU(launch_grid_sort(ctx->tgt_pts.as<float4>(), static_cast<uint32_t>(n), g, ctx->keys_in.as<uint64_t>(), ctx->keys_out.as<uint64_t>(),
                      ctx->vals_in.as<uint32_t>(), ctx->pre_vals_out.as<uint32_t>(), ctx->sort_temp.p, tb, d_distinct, ctx->grid_curve_order, ctx->stream));
  ctx->pending_clean = false;  // the two counters of the construction live in the pending counters
  uint32_t counters[2] = {0, 0};  // runs (= lists), longest run
  CU(cudaMemcpyAsync(counters, d_distinct, sizeof(counters), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  const uint32_t distinct = counters[0], max_list = counters[1];
  ctx->launches += 4;
  if (ctx->debug_pending) std::fprintf(stderr, "[sgb] grid front end: cell %.4g, %u distinct lists, longest %u points\n", cell, distinct, max_list);
  if (max_list > kGridMaxList) return 0;  // degenerate density (see kGridMaxList): the tree search alone stays exact and bounded
  uint32_t capacity = 1024;
  while (capacity < 2ull * distinct) capacity <<= 1;  // load factor <= 1/2
  CU(ctx->grid_table.reserve(static_cast<size_t>(capacity) * sizeof(GridSlot)));
  // lists are stored as pair records (two points each, sgb_grid.cuh): a list of odd length carries one pad
  CU(ctx->grid_pts.reserve((n_ent + distinct + 2) * sizeof(float4)));
  CU(launch_grid_fill(ctx->keys_out.as<uint64_t>(), ctx->pre_vals_out.as<uint32_t>(), ctx->tgt_pts.as<float4>(), static_cast<uint32_t>(n_ent), ctx->vals_in.as<uint32_t>(),
                      ctx->keys_in.as<uint32_t>(), ctx->sort_temp.p, tb, ctx->grid_pts.as<float4>(), ctx->grid_table.as<GridSlot>(), capacity, ctx->grid_curve_order,
                      ctx->stream));
  ctx->grid_blocks = true;
  ctx->launches += 5;
  for (int a = 0; a < 3; a++) ctx->grid_origin[a] = g.origin[a];
  ctx->grid_inv_cell = g.inv_cell;
  ctx->grid_settle_d2 = g.settle_d2;
  ctx->grid_cell = static_cast<float>(cell);
  ctx->grid_capacity = capacity;
  ctx->grid_ready = true;
  return 0;
}

}  // namespace sgb

extern "C" {

int sgb_estimate_features(sgb_ctx* ctx, size_t n, const double* points, int num_neighbors, double* out_normals, double* out_covs) {
  if (!ctx) return 1;
  if (n && !points) return fail(ctx, 1, "sgb_estimate_features: null points");
  if (num_neighbors < 1 || num_neighbors > 32) return fail(ctx, 1, "sgb_estimate_features: num_neighbors must be in 1..32");
  if (n >= (1ull << 30)) return fail(ctx, 1, "sgb_estimate_features: too many points");
  if (n == 0 || (!out_normals && !out_covs)) return 0;
  CU(cudaSetDevice(ctx->device));
  CU(ctx->stage_pts.reserve(n * 4 * sizeof(double)));
  CU(cudaMemcpyAsync(ctx->stage_pts.p, points, n * 4 * sizeof(double), cudaMemcpyDefault, ctx->stream));
  CU(ctx->pre_centre.reserve(4 * sizeof(double)));
  CU(ctx->pre_bounds.reserve(6 * sizeof(double)));
  CU(ctx->pre_pts.reserve(n * sizeof(float4)));
  CU(ctx->pre_lo.reserve(n * sizeof(float4)));
  CU(launch_bounds_centre(ctx->stage_pts.as<double>(), n, ctx->pre_bounds.as<double>(), ctx->pre_centre.as<double>(), ctx->sm_count, ctx->stream));
  CU(launch_convert(ctx->stage_pts.as<double>(), nullptr, nullptr, n, ctx->pre_centre.as<double>(), ctx->pre_pts.as<float4>(), nullptr, nullptr, nullptr, nullptr,
                    nullptr, ctx->sm_count, ctx->stream, ctx->pre_lo.as<float4>()));
  ctx->launches += 4;
  int depth = 0;
  ctx->src_tree_valid = false;  // the scratch tree buffers are about to hold THIS cloud's tree
  if (int rc = build_lbvh(ctx, ctx->pre_pts.as<float4>(), n, ctx->pre_centre.as<double>(), ctx->pre_perm, ctx->pre_leaf_pts, ctx->pre_nodes, &depth)) return rc;
  if (out_normals) CU(ctx->pre_out_normals.reserve(n * 4 * sizeof(double)));
  if (out_covs) CU(ctx->pre_out_covs.reserve(n * 16 * sizeof(double)));
  const int mode = (out_normals ? 1 : 0) | (out_covs ? 2 : 0);
  CU(launch_features(ctx->pre_nodes.as<float4>(), ctx->pre_leaf_pts.as<float4>(), static_cast<uint32_t>(n), num_neighbors, ctx->pre_centre.as<double>(), mode,
                     nullptr, nullptr, nullptr, out_normals ? ctx->pre_out_normals.as<double>() : nullptr, out_covs ? ctx->pre_out_covs.as<double>() : nullptr,
                     depth, 0, ctx->stream, ctx->pre_lo.as<float4>()));
  ctx->launches += 1;
  if (out_normals) CU(cudaMemcpyAsync(out_normals, ctx->pre_out_normals.p, n * 4 * sizeof(double), cudaMemcpyDefault, ctx->stream));
  if (out_covs) CU(cudaMemcpyAsync(out_covs, ctx->pre_out_covs.p, n * 16 * sizeof(double), cudaMemcpyDefault, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int sgb_target_estimate_features(sgb_ctx* ctx, int num_neighbors) {
  if (!ctx) return 1;
  if (num_neighbors < 1 || num_neighbors > 32) return fail(ctx, 1, "sgb_target_estimate_features: num_neighbors must be in 1..32");
  if (ctx->tgt_is_voxel || !ctx->tgt_ready) return fail(ctx, 1, "sgb_target_estimate_features: needs a point target with its kd-tree (set points, then set/build the tree)");
  const size_t n = ctx->n_tgt;
  ctx->tgt_has_normals = ctx->tgt_has_covs = true;
  ctx->have_lin = false;
  ctx->corr_seeds = false;
  if (n == 0) return 0;
  CU(cudaSetDevice(ctx->device));
  CU(ctx->tgt_normals.reserve(n * sizeof(float4)));
  CU(ctx->tgt_covA.reserve(n * sizeof(float4)));
  CU(ctx->tgt_covB.reserve(n * sizeof(float4)));
  // the target's own tree and leaf-ordered points are already resident: write the features straight into the leaf-ordered streams
  CU(launch_features(ctx->tgt_pnodes.as<float4>(), ctx->tgt_pts.as<float4>(), static_cast<uint32_t>(n), num_neighbors, ctx->tgt_centre.as<double>(), 3,
                     ctx->tgt_normals.as<float4>(), ctx->tgt_covA.as<float4>(), ctx->tgt_covB.as<float4>(), nullptr, nullptr, ctx->tree_depth, 1, ctx->stream,
                     ctx->tgt_has_lo ? ctx->tgt_orig_lo.as<float4>() : nullptr));
  ctx->launches += 1;
  // keep the ORIGINAL-order copies in step (a later sgb_target_build_kdtree / _set_kdtree re-gathers the leaf-ordered streams from them)
  CU(ctx->tgt_orig_normals.reserve(n * sizeof(float4)));
  CU(ctx->tgt_orig_covA.reserve(n * sizeof(float4)));
  CU(ctx->tgt_orig_covB.reserve(n * sizeof(float4)));
  CU(launch_scatter(ctx->tgt_perm.as<uint32_t>(), n, ctx->tgt_normals.as<float4>(), ctx->tgt_orig_normals.as<float4>(), ctx->tgt_covA.as<float4>(),
                    ctx->tgt_orig_covA.as<float4>(), ctx->tgt_covB.as<float4>(), ctx->tgt_orig_covB.as<float4>(), ctx->sm_count, ctx->stream));
  ctx->launches += 1;
  return 0;
}

int sgb_source_estimate_features(sgb_ctx* ctx, int num_neighbors) {
  if (!ctx) return 1;
  if (num_neighbors < 1 || num_neighbors > 32) return fail(ctx, 1, "sgb_source_estimate_features: num_neighbors must be in 1..32");
  const size_t n = ctx->n_src;
  ctx->src_has_covs = true;
  ctx->have_lin = false;
  ctx->corr_seeds = false;
  if (n == 0) return 0;
  CU(cudaSetDevice(ctx->device));
  // tmp_pts still holds the source in ORIGINAL order (centred FP32, w = index) from sgb_source_set_points
  if (!ctx->src_orig_valid) return fail(ctx, 1, "sgb_source_estimate_features: no source points (call sgb_source_set_points first)");
  int depth = 0;
  ctx->src_tree_valid = false;
  if (int rc = build_lbvh(ctx, ctx->tmp_pts.as<float4>(), n, ctx->src_centre.as<double>(), ctx->pre_perm, ctx->pre_leaf_pts, ctx->pre_nodes, &depth)) return rc;
  ctx->src_tree_valid = true;  // kept: sgb_target_adopt_source turns this cloud, tree and covariances into the next target
  ctx->src_tree_depth = depth;
  ctx->src_cov_orig_valid = true;
  CU(ctx->tmp_covA.reserve(n * sizeof(float4)));
  CU(ctx->tmp_covB.reserve(n * sizeof(float4)));
  CU(ctx->src_covA.reserve(n * sizeof(float4)));
  CU(ctx->src_covB.reserve(n * sizeof(float4)));
  CU(launch_features(ctx->pre_nodes.as<float4>(), ctx->pre_leaf_pts.as<float4>(), static_cast<uint32_t>(n), num_neighbors, ctx->src_centre.as<double>(), 2, nullptr,
                     ctx->tmp_covA.as<float4>(), ctx->tmp_covB.as<float4>(), nullptr, nullptr, depth, 0, ctx->stream, ctx->src_has_lo ? ctx->tmp_lo.as<float4>() : nullptr));
  // original order -> the search order of the source (chunk-transposed Morton)
  CU(launch_gather(ctx->src_perm.as<uint32_t>(), n, ctx->tmp_covA.as<float4>(), ctx->src_covA.as<float4>(), ctx->tmp_covB.as<float4>(), ctx->src_covB.as<float4>(),
                   nullptr, nullptr, nullptr, nullptr, ctx->sm_count, ctx->stream));
  ctx->launches += 2;
  return 0;
}

int sgb_voxelgrid_sampling(sgb_ctx* ctx, size_t n, const double* points, double leaf_size, double* out_points, size_t* n_out) {
  if (!ctx || !n_out) return 1;
  *n_out = 0;
  if (n && (!points || !out_points)) return fail(ctx, 1, "sgb_voxelgrid_sampling: null buffer");
  if (!(leaf_size > 0.0)) return fail(ctx, 1, "sgb_voxelgrid_sampling: leaf_size must be positive");
  if (n >= (1ull << 31)) return fail(ctx, 1, "sgb_voxelgrid_sampling: too many points");
  if (n == 0) return 0;
  CU(cudaSetDevice(ctx->device));
  CU(ctx->stage_pts.reserve(n * 4 * sizeof(double)));
  CU(cudaMemcpyAsync(ctx->stage_pts.p, points, n * 4 * sizeof(double), cudaMemcpyDefault, ctx->stream));
  CU(ctx->keys_in.reserve(n * sizeof(uint64_t)));
  CU(ctx->keys_out.reserve(n * sizeof(uint64_t)));
  CU(ctx->vals_in.reserve(n * sizeof(uint32_t)));
  CU(ctx->pre_vals_out.reserve(n * sizeof(uint32_t)));
  CU(ctx->pre_heads.reserve((n + 1) * sizeof(uint32_t)));
  CU(ctx->pre_slots.reserve((n + 1) * sizeof(uint32_t)));
  CU(ctx->stage_covs.reserve(n * 4 * sizeof(double)));  // output staging (re-uses a scratch buffer)
  CU(launch_voxel_keys(ctx->stage_pts.as<double>(), n, 1.0 / leaf_size, ctx->keys_in.as<uint64_t>(), ctx->vals_in.as<uint32_t>(), ctx->sm_count, ctx->stream));
  size_t temp_bytes = 0;
  CU(sort_pairs_u64_u32(nullptr, temp_bytes, ctx->keys_in.as<uint64_t>(), ctx->keys_out.as<uint64_t>(), ctx->vals_in.as<uint32_t>(),
                        ctx->pre_vals_out.as<uint32_t>(), n, ctx->stream));
  size_t scan_bytes = 0;
  CU(exclusive_sum_u32(nullptr, scan_bytes, ctx->pre_heads.as<uint32_t>(), ctx->pre_slots.as<uint32_t>(), n + 1, ctx->stream));
  CU(ctx->sort_temp.reserve(temp_bytes > scan_bytes ? temp_bytes : scan_bytes));
  CU(sort_pairs_u64_u32(ctx->sort_temp.p, temp_bytes, ctx->keys_in.as<uint64_t>(), ctx->keys_out.as<uint64_t>(), ctx->vals_in.as<uint32_t>(),
                        ctx->pre_vals_out.as<uint32_t>(), n, ctx->stream));
  CU(cudaMemsetAsync(ctx->pre_heads.p, 0, (n + 1) * sizeof(uint32_t), ctx->stream));
  CU(launch_voxel_heads(ctx->keys_out.as<uint64_t>(), n, ctx->pre_heads.as<uint32_t>(), ctx->sm_count, ctx->stream));
  CU(exclusive_sum_u32(ctx->sort_temp.p, scan_bytes, ctx->pre_heads.as<uint32_t>(), ctx->pre_slots.as<uint32_t>(), n + 1, ctx->stream));
  CU(launch_voxel_means(ctx->keys_out.as<uint64_t>(), ctx->pre_vals_out.as<uint32_t>(), ctx->pre_heads.as<uint32_t>(), ctx->pre_slots.as<uint32_t>(), n,
                        ctx->stage_pts.as<double>(), ctx->stage_covs.as<double>(), ctx->sm_count, ctx->stream));
  ctx->launches += 8;
  uint32_t count = 0;
  CU(cudaMemcpyAsync(&count, ctx->pre_slots.as<uint32_t>() + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (count) {
    CU(cudaMemcpyAsync(out_points, ctx->stage_covs.p, static_cast<size_t>(count) * 4 * sizeof(double), cudaMemcpyDefault, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  *n_out = count;
  return 0;
}

// The frame-stream hand-over (odometry_benchmark_small_gicp_tbb.cpp:41-43: `target_points = points; target_tree = tree;`): the cloud that was
// the source of this align() is the target of the next one.  Everything that already exists on the device is taken over, not rebuilt.
int sgb_target_adopt_source(sgb_ctx* ctx) {
  if (!ctx) return 1;
  if (!ctx->src_orig_valid || ctx->n_src == 0) return fail(ctx, 1, "sgb_target_adopt_source: no source points (call sgb_source_set_points first)");
  CU(cudaSetDevice(ctx->device));
  const size_t n = ctx->n_src;
  const bool covs = ctx->src_has_covs;
  // covariances in ORIGINAL order: either still there from the device-side estimation, or un-permuted from the search-ordered streams
  if (covs && !ctx->src_cov_orig_valid) {
    CU(ctx->tmp_covA.reserve(n * sizeof(float4)));
    CU(ctx->tmp_covB.reserve(n * sizeof(float4)));
    CU(launch_scatter(ctx->src_perm.as<uint32_t>(), n, ctx->src_covA.as<float4>(), ctx->tmp_covA.as<float4>(), ctx->src_covB.as<float4>(), ctx->tmp_covB.as<float4>(),
                      nullptr, nullptr, ctx->sm_count, ctx->stream));
    ctx->launches += 1;
  }
  ctx->tgt_orig_pts.swap(ctx->tmp_pts);
  ctx->tgt_has_lo = ctx->src_has_lo;
  if (ctx->src_has_lo) ctx->tgt_orig_lo.swap(ctx->tmp_lo);
  if (covs) {
    ctx->tgt_orig_covA.swap(ctx->tmp_covA);
    ctx->tgt_orig_covB.swap(ctx->tmp_covB);
  }
  ctx->tgt_centre.swap(ctx->src_centre);
  ctx->tgt_bounds.swap(ctx->src_bounds);
  CU(ctx->src_centre.reserve(4 * sizeof(double)));
  CU(ctx->src_bounds.reserve(6 * sizeof(double)));
  ctx->n_tgt = n;
  ctx->tgt_has_normals = false;
  ctx->tgt_has_covs = covs;
  ctx->tgt_is_voxel = false;
  ctx->tgt_has_kd = false;
  ctx->tgt_feats_leaf_only = false;
  ctx->grid_ready = false;
  ctx->have_lin = false;
  ctx->corr_seeds = false;
  int depth = ctx->src_tree_depth;
  if (ctx->src_tree_valid) {
    ctx->tgt_perm.swap(ctx->pre_perm);
    ctx->tgt_pts.swap(ctx->pre_leaf_pts);
    ctx->tgt_pnodes.swap(ctx->pre_nodes);
  } else if (int rc = build_lbvh(ctx, ctx->tgt_orig_pts.as<float4>(), n, ctx->tgt_centre.as<double>(), ctx->tgt_perm, ctx->tgt_pts, ctx->tgt_pnodes, &depth)) {
    return rc;
  }
  if (depth > 40) return fail(ctx, 1, "sgb_target_adopt_source: tree too deep");
  if (covs) {
    CU(ctx->tgt_covA.reserve(n * sizeof(float4)));
    CU(ctx->tgt_covB.reserve(n * sizeof(float4)));
    CU(launch_gather(ctx->tgt_perm.as<uint32_t>(), n, nullptr, nullptr, nullptr, nullptr, ctx->tgt_orig_covA.as<float4>(), ctx->tgt_covA.as<float4>(),
                     ctx->tgt_orig_covB.as<float4>(), ctx->tgt_covB.as<float4>(), ctx->sm_count, ctx->stream));
    ctx->launches += 1;
  }
  ctx->tree_depth = depth;
  ctx->n_pnodes = (static_cast<size_t>(1) << (depth - 1)) - 1;
  ctx->tgt_ready = true;
  // the context has no source until the next sgb_source_set_points (its buffers now belong to the target)
  ctx->n_src = 0;
  ctx->src_has_covs = ctx->src_orig_valid = ctx->src_tree_valid = ctx->src_cov_orig_valid = ctx->src_has_lo = false;
  return build_grid(ctx);
}

// Gaussian voxel map built on the device: replaces IncrementalVoxelMap<GaussianVoxel>::insert
// (incremental_voxelmap.hpp:55-92) + GaussianVoxel::add / finalize (gaussian_voxelmap.hpp:30-62) for a one-shot map.
int sgb_target_build_voxelmap(sgb_ctx* ctx, size_t n, const double* points, const double* covs, double leaf_size, int search_offsets) {
  if (!ctx) return 1;
  if (!(leaf_size > 0.0)) return fail(ctx, 1, "sgb_target_build_voxelmap: leaf_size must be positive");
  if (n && !points) return fail(ctx, 1, "sgb_target_build_voxelmap: null points");
  if (n >= (1ull << 31)) return fail(ctx, 1, "sgb_target_build_voxelmap: too many points");
  if (search_offsets != 1 && search_offsets != 7 && search_offsets != 27) search_offsets = 1;  // incremental_voxelmap.hpp:159-163
  CU(cudaSetDevice(ctx->device));
  ctx->tgt_has_normals = false;
  ctx->tgt_has_covs = covs != nullptr;
  ctx->tgt_is_voxel = true;
  ctx->tgt_has_kd = false;
  ctx->grid_ready = false;
  ctx->tgt_ready = true;
  ctx->have_lin = false;
  ctx->corr_seeds = false;
  ctx->vox_offsets = search_offsets;
  ctx->vox_inv_leaf = 1.0 / leaf_size;
  ctx->n_tgt = 0;
  CU(ctx->tgt_centre.reserve(4 * sizeof(double)));
  CU(ctx->tgt_bounds.reserve(6 * sizeof(double)));
  if (n == 0) return 0;
  CU(ctx->stage_pts.reserve(n * 4 * sizeof(double)));
  CU(cudaMemcpyAsync(ctx->stage_pts.p, points, n * 4 * sizeof(double), cudaMemcpyDefault, ctx->stream));
  if (covs) {
    CU(ctx->stage_covs.reserve(n * 16 * sizeof(double)));
    CU(cudaMemcpyAsync(ctx->stage_covs.p, covs, n * 16 * sizeof(double), cudaMemcpyDefault, ctx->stream));
  }
  CU(ctx->keys_in.reserve(n * sizeof(uint64_t)));
  CU(ctx->keys_out.reserve(n * sizeof(uint64_t)));
  CU(ctx->vals_in.reserve(n * sizeof(uint32_t)));
  CU(ctx->pre_vals_out.reserve(n * sizeof(uint32_t)));
  CU(ctx->pre_heads.reserve((n + 1) * sizeof(uint32_t)));
  CU(ctx->pre_slots.reserve((n + 1) * sizeof(uint32_t)));
  CU(launch_voxel_keys(ctx->stage_pts.as<double>(), n, 1.0 / leaf_size, ctx->keys_in.as<uint64_t>(), ctx->vals_in.as<uint32_t>(), ctx->sm_count, ctx->stream));
  size_t temp_bytes = 0, scan_bytes = 0;
  CU(sort_pairs_u64_u32(nullptr, temp_bytes, ctx->keys_in.as<uint64_t>(), ctx->keys_out.as<uint64_t>(), ctx->vals_in.as<uint32_t>(),
                        ctx->pre_vals_out.as<uint32_t>(), n, ctx->stream));
  CU(exclusive_sum_u32(nullptr, scan_bytes, ctx->pre_heads.as<uint32_t>(), ctx->pre_slots.as<uint32_t>(), n + 1, ctx->stream));
  CU(ctx->sort_temp.reserve(temp_bytes > scan_bytes ? temp_bytes : scan_bytes));
  CU(sort_pairs_u64_u32(ctx->sort_temp.p, temp_bytes, ctx->keys_in.as<uint64_t>(), ctx->keys_out.as<uint64_t>(), ctx->vals_in.as<uint32_t>(),
                        ctx->pre_vals_out.as<uint32_t>(), n, ctx->stream));
  CU(cudaMemsetAsync(ctx->pre_heads.p, 0, (n + 1) * sizeof(uint32_t), ctx->stream));
  CU(launch_voxel_heads(ctx->keys_out.as<uint64_t>(), n, ctx->pre_heads.as<uint32_t>(), ctx->sm_count, ctx->stream));
  CU(exclusive_sum_u32(ctx->sort_temp.p, scan_bytes, ctx->pre_heads.as<uint32_t>(), ctx->pre_slots.as<uint32_t>(), n + 1, ctx->stream));
  uint32_t n_vox = 0;
  CU(cudaMemcpyAsync(&n_vox, ctx->pre_slots.as<uint32_t>() + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->launches += 7;
  if (n_vox == 0) return 0;  // every point outside the 21-bit voxel range
  // per-voxel mean / covariance / integer coordinates (FP64), then the same device-side conversion as sgb_target_set_voxelmap
  CU(ctx->pre_out_normals.reserve(static_cast<size_t>(n_vox) * 4 * sizeof(double)));
  if (covs) CU(ctx->pre_out_covs.reserve(static_cast<size_t>(n_vox) * 16 * sizeof(double)));
  CU(ctx->pre_vox_coords.reserve(static_cast<size_t>(n_vox) * sizeof(int4)));  // NOT tmp_pts: that buffer holds the source in original order
  CU(launch_voxel_stats(ctx->keys_out.as<uint64_t>(), ctx->pre_vals_out.as<uint32_t>(), ctx->pre_heads.as<uint32_t>(), ctx->pre_slots.as<uint32_t>(), n,
                        ctx->stage_pts.as<double>(), covs ? ctx->stage_covs.as<double>() : nullptr, ctx->pre_out_normals.as<double>(),
                        covs ? ctx->pre_out_covs.as<double>() : nullptr, ctx->pre_vox_coords.as<int4>(), ctx->sm_count, ctx->stream));
  CU(ctx->tgt_pts.reserve(static_cast<size_t>(n_vox) * sizeof(float4)));
  if (covs) {
    CU(ctx->tgt_covA.reserve(static_cast<size_t>(n_vox) * sizeof(float4)));
    CU(ctx->tgt_covB.reserve(static_cast<size_t>(n_vox) * sizeof(float4)));
  }
  CU(launch_bounds_centre(ctx->pre_out_normals.as<double>(), n_vox, ctx->tgt_bounds.as<double>(), ctx->tgt_centre.as<double>(), ctx->sm_count, ctx->stream));
  CU(launch_convert(ctx->pre_out_normals.as<double>(), nullptr, covs ? ctx->pre_out_covs.as<double>() : nullptr, n_vox, ctx->tgt_centre.as<double>(),
                    ctx->tgt_pts.as<float4>(), nullptr, ctx->tgt_covA.as<float4>(), ctx->tgt_covB.as<float4>(), nullptr, nullptr, ctx->sm_count, ctx->stream));
  uint32_t capacity = 16;
  while (capacity < 2ull * n_vox) capacity <<= 1;  // load factor <= 1/2
  CU(ctx->vox_table.reserve(static_cast<size_t>(capacity) * sizeof(int4)));
  CU(launch_vox_table_build(ctx->pre_vox_coords.as<int4>(), n_vox, ctx->vox_table.as<int4>(), capacity, ctx->stream));
  ctx->launches += 7;
  ctx->vox_mask = capacity - 1;
  ctx->n_tgt = n_vox;
  return 0;
}

// k nearest neighbours of arbitrary queries in the target's search structure (the batch NN API)
int sgb_target_batch_knn(sgb_ctx* ctx, size_t n_queries, const double* queries, int k, uint64_t* out_indices, double* out_sq_dists) {
  if (!ctx) return 1;
  if (k < 1 || k > 32) return fail(ctx, 1, "sgb_target_batch_knn: k must be in 1..32");
  if (n_queries && (!queries || !out_indices || !out_sq_dists)) return fail(ctx, 1, "sgb_target_batch_knn: null buffer");
  if (n_queries >= (1ull << 31) / static_cast<size_t>(k)) return fail(ctx, 1, "sgb_target_batch_knn: too many queries");
  if (ctx->tgt_is_voxel || !ctx->tgt_ready) return fail(ctx, 1, "sgb_target_batch_knn: needs a point target with its kd-tree (set points, then set/build the tree)");
  if (n_queries == 0) return 0;
  CU(cudaSetDevice(ctx->device));
  const size_t n = n_queries, nk = n * static_cast<size_t>(k);
  if (ctx->n_tgt == 0) {  // nothing to find: KnnResult's initial state
    for (size_t j = 0; j < nk; j++) {
      out_indices[j] = ~0ull;
      out_sq_dists[j] = std::numeric_limits<double>::max();
    }
    return 0;
  }
  CU(ctx->stage_pts.reserve(n * 4 * sizeof(double)));
  CU(cudaMemcpyAsync(ctx->stage_pts.p, queries, n * 4 * sizeof(double), cudaMemcpyDefault, ctx->stream));
  CU(ctx->pre_out_covs.reserve(nk * sizeof(double)));
  CU(ctx->corr64.reserve(nk * sizeof(uint64_t)));
  CU(launch_batch_knn(ctx->tgt_pnodes.as<float4>(), ctx->tgt_pts.as<float4>(), ctx->stage_pts.as<double>(), static_cast<uint32_t>(n), k, ctx->tgt_centre.as<double>(),
                      ctx->corr64.as<unsigned long long>(), ctx->pre_out_covs.as<double>(), ctx->tree_depth, ctx->stream));
  ctx->launches += 1;
  CU(cudaMemcpyAsync(out_indices, ctx->corr64.p, nk * sizeof(uint64_t), cudaMemcpyDefault, ctx->stream));
  CU(cudaMemcpyAsync(out_sq_dists, ctx->pre_out_covs.p, nk * sizeof(double), cudaMemcpyDefault, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // extern "C"
