// SPDX-License-Identifier: MIT
// The context behind the opaque sgb_ctx handle of include/sgicp_b200.h (shared by sgb_capi.cu / sgb_capi_preprocess.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace sgb {

/// Grow-only device buffer (re-used across calls so that streams of similarly sized frames never reallocate).
/// Owns its allocation: freed with the context (sgb_destroy makes the context's device current first).
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    const size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const {
    return static_cast<T*>(p);
  }
  void swap(DevBuf& o) {
    void* tp = p;
    p = o.p;
    o.p = tp;
    const size_t tc = cap;
    cap = o.cap;
    o.cap = tc;
  }
};

}  // namespace sgb

struct sgb_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  // second stream + events of the pipelined source upload (sgb_source_set_points): the covariance chunks travel on it while the
  // context's stream converts / orders the points and the chunks that have already arrived
  cudaStream_t copy_stream = nullptr;
  static constexpr int kUploadChunks = 8;
  cudaEvent_t ev_upload[kUploadChunks] = {};
  cudaEvent_t ev_entry = nullptr;
  std::string err;
  uint64_t launches = 0;
  // The product library has ONE search path (grid front end + packet / pending search, then the factor kernel).  The fields below that
  // select anything else are only ever changed by the profiling build (-DSGB_PROFILING, libsgicp_b200_prof.so: A/B runs in profiles/).
  int search_mode = 2;  // 2: packet (warp-cooperative) search + factor kernel; [profiling build] 1: per-thread search + factor kernel; 0: single fused kernel

  // ---- target ----
  size_t n_tgt = 0;
  bool tgt_has_normals = false, tgt_has_covs = false;
  bool tgt_is_voxel = false, tgt_ready = false;
  sgb::DevBuf tgt_orig_pts, tgt_orig_normals, tgt_orig_covA, tgt_orig_covB;  // original order
  sgb::DevBuf tgt_orig_lo, tmp_lo, pre_lo;  // what the FP32 rounding of the coordinates dropped (target / source / scratch cloud, original order): feature estimation only
  sgb::DevBuf tgt_pts, tgt_normals, tgt_covA, tgt_covB;                      // leaf order (or voxel order)
  sgb::DevBuf tgt_nodes, tgt_perm, tgt_pnodes;  // kd nodes (8 B), leaf permutation, packet records (64 B / inner node)
  int tree_depth = 0;
  bool tgt_has_kd = false;  // 8-byte kd nodes present (adopted / host-built trees); device-built trees only have packet records
  bool host_tree = false;   // profiling switch: sgb_target_build_kdtree builds a median-split kd-tree on the host
  int tree_quality = 1;     // device construction: 1 = Hilbert order + per-level median-split refinement (kd quality), 0 = Hilbert order only (linear BVH)
  sgb::DevBuf pre_boxes;
  size_t n_pnodes = 0;      // packet records of the current target tree
  // uniform-grid front end of the search (sgb_grid.cu)
  bool use_grid = true, grid_ready = false, grid_blocks = true;  // grid_blocks: the front end holds 2 x 2 x 2 block lists (one lookup per query)
  sgb::DevBuf grid_pts, grid_table, grid_state, grid_spacing, grid_pending;  // grid_pending: [0], [1] = alternating counters, [2..] = pending query positions
  sgb::DevBuf grid_pending_q;  // parallel to the list: transformed query + squared distance of the probe's best candidate
  bool use_ring = true, debug_pending = false;  // profiling switches (sgb_create)
  int pending_div = 16;           // more than n_src / pending_div pending queries: packet search, else a warp per pending query
  double grid_cell_factor = 2.5;  // cell edge in units of the median point spacing (sweep in profiles/r01: 2 / 2.5 / 3 / 4)
  uint32_t grid_capacity = 0;
  bool pending_clean = false;  // both pending counters are zero / maintained by the probe kernels
  int pending_parity = 0;
  sgb::DevBuf packet_queue;      // two alternating chunk-queue counters of the packet search (each launch clears the other)
  sgb::DevBuf chunk_lists;       // [2 parities][3] class counters, then [3][n_chunks] chunk ids: the packet search's work lists, filled by the probe
  bool chunk_lists_clean = false;
  uint32_t class_fallback_pct = 85;  // more than this share of the chunks listed (pending lanes almost everywhere): curve order instead of the lists
  float class_wide_cells = 2.0f;     // search radius (in cells) from which a chunk counts as wide
  bool grid_curve_order = true;  // block lists laid out along a Morton curve (profiling switch SGB_GRID_ORDER=0: raster order of the packed coordinates)
  bool kd_smem_refine = true;    // subtrees of <= 2048 points refined in shared memory by one launch (profiling switch SGB_KD_SMEM=0: one radix sort per level)
  bool error_pipelined = true;   // Reduction::error through the factor kernel's cp.async operand pipeline (profiling switch SGB_ERROR_PIPE=0: plain kernel)
  bool tma_leaf = false;         // profiling switch SGB_TMA_LEAF=1 (A/B of the north-star's TMA leaf staging)
  bool use_chunk_classes = true; // profiling switch SGB_CHUNK_CLASSES=0: the packet search scans all chunks in curve order
  int packet_parity = 0;
  bool use_packet_queue = true;  // profiling switch SGB_PACKET_QUEUE=0: static stride
  float grid_origin[3] = {0, 0, 0}, grid_inv_cell = 1.f, grid_settle_d2 = 0.f, grid_cell = 0.f;
  sgb::DevBuf tgt_centre, tgt_bounds;  // 4 / 6 doubles
  sgb::DevBuf vox_table;
  uint32_t vox_mask = 0;
  int vox_offsets = 1;
  double vox_inv_leaf = 1.0;

  // ---- source ----
  size_t n_src = 0;
  bool src_has_covs = false;
  sgb::DevBuf src_pts, src_covA, src_covB, src_perm, src_rank, src_centre, src_bounds;
  uint32_t src_run = 1;  // K of the chunk-transposed source layout (DevSource::run)

  // ---- scratch ----
  sgb::DevBuf stage_pts, stage_normals, stage_covs;  // raw double uploads
  sgb::DevBuf tmp_pts, tmp_covA, tmp_covB, keys_in, keys_out, vals_in, sort_temp;
  sgb::DevBuf corr, partials, ticket, out44, corr64;
  double* h_out = nullptr;      // page-locked AND mapped, 64 doubles: [0,44) H|b|e|inliers, [48] error(), [56] scratch, [60] comm status
  double* h_out_dev = nullptr;  // device address of the same memory (the reduction's finishing CTA writes into it)

  // ---- preprocessing scratch (sgb_capi_preprocess.cu) ----
  sgb::DevBuf pre_pts, pre_leaf_pts, pre_nodes, pre_perm, pre_centre, pre_bounds, pre_out_normals, pre_out_covs, pre_heads, pre_slots, pre_vals_out;
  sgb::DevBuf pre_vox_coords;  // integer voxel coordinates of sgb_target_build_voxelmap (its own scratch: tmp_pts belongs to the source)
  bool tgt_has_lo = false, src_has_lo = false;  // tgt_orig_lo / tmp_lo hold the residuals of the current target / source
  bool src_tree_valid = false;  // pre_perm / pre_leaf_pts / pre_nodes hold the SOURCE's tree (sgb_source_estimate_features built it), depth src_tree_depth
  bool src_cov_orig_valid = false;  // tmp_covA / tmp_covB hold the source's covariances in original order (estimated on the device)
  int src_tree_depth = 0;
  bool src_orig_valid = false;  // tmp_pts holds the current source in original order (sgb_source_estimate_features needs it)
  bool tgt_feats_leaf_only = false;  // normals / covariances were estimated on the device into the leaf-ordered streams only

  // ---- multi-GPU exchange fused into the reduction's finishing CTA (sgb_comm_*, CommParams in sgb_device.cuh) ----
  sgb::DevBuf comm_mail;                  // this rank's mailbox (zeroed at allocation)
  unsigned char* comm_peers[8] = {};      // mailbox of every rank as mapped into this process (own entry = comm_mail.p)
  bool comm_ipc_opened[8] = {};           // mapped by cudaIpcOpenMemHandle (to be closed)
  int comm_world = 0, comm_rank = 0;      // world <= 1: single GPU
  unsigned long long comm_seq = 0;        // collective calls issued so far
  unsigned long long comm_timeout_ns = 5000000000ull;  // a peer that has not arrived after 5 s (sgb_comm_set_timeout_ms) is given up on
  sgb::DevBuf comm_status;                // one sticky word: != 0 after an exchange gave up on a peer (read back with every host-visible result)
  bool comm_failed = false;               // host copy of that word: every later call of this context fails until sgb_comm_connect* is called again

  // ---- state of the last linearize (cached for error(), gicp_factor.hpp:94-96) ----
  bool have_lin = false;
  bool corr_seeds = false;  // corr[] holds correspondences of the current (target, tree, source) triple
  int lin_factor = 0, lin_robust = 0;
  double lin_c = 1.0;
  double Tlin[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
  const double* last_out = nullptr;  // device pointer holding H|b|e|inliers of the last linearize
  int lin_grid = 0;
};

namespace sgb {

inline int fail(sgb_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}

/// device-side tree construction (sgb_capi_preprocess.cu)
int build_lbvh(sgb_ctx* ctx, const float4* d_orig_pts, size_t n, const double* d_centre4, DevBuf& perm, DevBuf& leaf_pts, DevBuf& pnodes, int* depth);
/// uniform grid over the current (leaf-ordered) target for the search front end; no-op unless enabled
int build_grid(sgb_ctx* ctx);

}  // namespace sgb

#define CU(expr)                                                                                                 \
  do {                                                                                                           \
    cudaError_t _e = (expr);                                                                                     \
    if (_e != cudaSuccess) return sgb::fail(ctx, 2, std::string(#expr) + ": " + cudaGetErrorString(_e));        \
  } while (0)
