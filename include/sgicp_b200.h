/* SPDX-License-Identifier: MIT
 *
 * sgicp_b200.h -- C-ABI of the B200-native small_gicp hot path (libsgicp_b200.so).
 *
 * One context = one GPU = one CUDA stream.  A context is NOT thread-safe; distinct contexts are
 * (the reference runs several align() calls concurrently from TBB flow-graph nodes,
 * src/benchmark/odometry_benchmark_small_gicp_tbb_flow.cpp:81-98 -- give each its own context).
 *
 * Every entry point returns 0 on success and a non-zero code on failure; sgb_last_error() then
 * returns a description.  There is NO CPU fallback: without a CUDA device sgb_create() fails.
 *
 * All host input buffers are borrowed for the duration of the call only; the context owns all
 * device memory.  One exception, by CUDA's own rules: a cloud passed in PAGE-LOCKED host memory
 * (cudaMallocHost / cudaHostRegister / torch pin_memory) is copied asynchronously on the context's
 * stream, so sgb_{target,source}_set_points / sgb_target_set_voxelmap may return before it has been
 * read -- leave such a buffer unchanged until sgb_synchronize() or any call that returns results to
 * the host (sgb_linearize, sgb_error, sgb_correspondences, ...).  Pageable memory (std::vector,
 * numpy) is consumed when the call returns.  Every bulk input / output pointer (points, normals, covariances, queries, feature and
 * correspondence outputs) may also be a DEVICE pointer of the context's GPU (copies use cudaMemcpyDefault): a device-resident producer hands
 * its buffers over without a host round trip.  Host layouts are exactly the reference's in-memory layouts so that the header
 * glue can pass `cloud.points[0].data()` etc. without repacking:
 *   points / normals : N x 4 doubles (x,y,z,1) / (nx,ny,nz,0)   -- std::vector<Eigen::Vector4d>,
 *                      include/small_gicp/points/point_cloud.hpp:69-70
 *   covariances      : N x 16 doubles, 4x4 with zero 4th row/col (symmetric, so row/col-major
 *                      agree)                                    -- point_cloud.hpp:71
 *   poses            : 16 doubles, COLUMN-major 4x4 = Eigen::Isometry3d::data()
 *   H | b | e        : 36 + 6 + 1 doubles (H symmetric)          -- the tuple returned by
 *                      Reduction::linearize, registration/reduction.hpp:20-27
 */
#ifndef SGICP_B200_H_
#define SGICP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sgb_ctx sgb_ctx;

/* Per-point factor evaluated by the kernel.
 * replaces ICPFactor::linearize/error          include/small_gicp/factors/icp_factor.hpp:20-64
 *          PointToPlaneICPFactor::linearize/.. include/small_gicp/factors/plane_icp_factor.hpp:20-69
 *          GICPFactor::linearize/error         include/small_gicp/factors/gicp_factor.hpp:35-89 */
enum sgb_factor_kind { SGB_FACTOR_ICP = 0, SGB_FACTOR_PLANE_ICP = 1, SGB_FACTOR_GICP = 2 };
/* RobustFactor<Huber|Cauchy, F>   include/small_gicp/factors/robust_kernel.hpp:11-106 */
enum sgb_robust_kind { SGB_ROBUST_NONE = 0, SGB_ROBUST_HUBER = 1, SGB_ROBUST_CAUCHY = 2 };
/* NullRejector / DistanceRejector include/small_gicp/registration/rejector.hpp:11-28 */
enum sgb_rejector_kind { SGB_REJECT_NONE = 0, SGB_REJECT_DISTANCE = 1 };

#define SGB_NO_CORRESPONDENCE UINT64_MAX /* std::numeric_limits<size_t>::max(), icp_factor.hpp:66 */

/* ---- context ------------------------------------------------------------------------------ */
int sgb_create(int device_id, sgb_ctx** out_ctx);
void sgb_destroy(sgb_ctx* ctx);
/* ctx may be NULL: returns the error of the last failed sgb_create() on this thread. */
const char* sgb_last_error(const sgb_ctx* ctx);
/* Run all work of this context on an existing cudaStream_t (e.g. the caller's / torch's current
 * stream) instead of the context's own.  Pass NULL to go back to the private stream. */
int sgb_set_stream(sgb_ctx* ctx, void* cuda_stream);
/* Block until everything queued by this context has finished. */
int sgb_synchronize(sgb_ctx* ctx);
/* Number of this library's kernels launched by the context since creation (bench.py gpu_launches). */
uint64_t sgb_kernel_launches(const sgb_ctx* ctx);

/* ---- multi-GPU (one process per GPU, source sharded, target replicated).  The reference has no
 *      counterpart: its reductions (reduction_omp.hpp:24-58, reduction_tbb.hpp:70-92) sum within one process.  After
 *      sgb_comm_connect* every sgb_linearize* / sgb_error* of this context returns the SUM OVER ALL RANKS: the CTA that
 *      finishes the reduction writes its sums straight into the peers' mailboxes over NVLink and adds theirs (one kernel,
 *      no NCCL call, deterministic rank order).  All ranks must issue the same sequence of linearize / error calls. ---- */
#define SGB_COMM_HANDLE_BYTES 64
#define SGB_COMM_MAX_RANKS 8
/* CUDA IPC handle (64 bytes) of this context's mailbox; allocate it on first use. */
int sgb_comm_handle(sgb_ctx* ctx, void* out_handle64);
/* `handles` = world x 64 bytes, rank order, gathered by the caller (e.g. torch.distributed.all_gather). */
int sgb_comm_connect(sgb_ctx* ctx, int rank, int world, const void* handles);
/* Same wiring from raw device pointers (contexts of ONE process: tests, single-process multi-stream setups). */
int sgb_comm_mailbox(sgb_ctx* ctx, void** out_device_ptr);
int sgb_comm_connect_ptrs(sgb_ctx* ctx, int rank, int world, void* const* mailboxes);
int sgb_comm_disconnect(sgb_ctx* ctx);
/* How long a rank waits for a peer before it gives up on an exchange (default 5000 ms). */
int sgb_comm_set_timeout_ms(sgb_ctx* ctx, int milliseconds);
/* Failure handling of the fused exchange.  A rank that gives up on a peer (timeout), or that hears from a peer that its call failed on
 * the host, gets NaN sums AND a sticky status word: the host-returning calls (sgb_linearize, sgb_error) then return a non-zero code
 * (4), every later collective call of the context returns 5 until sgb_comm_connect* is called again.  Users of the *_device variants
 * read the status themselves: 0 = fine, bit 0 = timeout, bit 1 = a peer reported a failed call.  A call that fails on the host of one
 * rank (bad argument, missing covariances) still takes part in the exchange -- it tells the peers instead of leaving them waiting. */
int sgb_comm_status(sgb_ctx* ctx, int* out_status);
/* Diagnostics: nanoseconds this rank's finishing CTA waited for its slowest peer in each of the last 64 exchanges (ring indexed by
 * call number % 64; *out_calls = exchanges so far).  The rank that arrives last sees the bare transport latency, the first one skew
 * + transport: the minimum and maximum over the ranks separate the two (bench.py `comm_wait_us`). */
int sgb_comm_wait_ns(sgb_ctx* ctx, uint64_t* out_ring64, uint64_t* out_calls);

/* ---- target: replaces traits::point/normal/cov(target, k) (points/traits.hpp:38-54) and the
 *      target_tree argument of Reduction::linearize (reduction.hpp:23) -------------------------- */
int sgb_target_set_points(sgb_ctx* ctx, size_t n, const double* points_xyz1, const double* normals_xyz0 /*or NULL*/,
                          const double* covs_4x4 /*or NULL*/);
/* Adopt a tree built by the reference: `nodes24` = KdTree<PointCloud>::kdtree.nodes.data()
 * (KdTreeNode<AxisAlignedProjection>, 24 bytes, ann/kdtree.hpp:56-71), `indices` = ...kdtree.indices.data()
 * (ann/kdtree.hpp:237).  Any builder's node order is accepted (serial / OMP / TBB). Leaf scan order is
 * preserved, so exact-tie behaviour follows the reference's (ann/knn_result.hpp:80-83). */
int sgb_target_set_kdtree(sgb_ctx* ctx, const void* nodes24, size_t n_nodes, uint32_t root, const uint64_t* indices);
/* Or build this library's own kd-tree over the current target points ON THE DEVICE (replaces
 * KdTreeBuilder::build_tree, ann/kdtree.hpp:74-131, and the OMP / TBB builders): Hilbert-order sort, then one radix sort
 * per level along the widest axis of every node's box -- a balanced median-split kd-tree with at most 32 points per leaf
 * (one per lane of the warp-cooperative leaf scan; the reference's builder uses 20), stored implicitly.  Asynchronous on
 * the context's stream, no host round trip.  Exact nearest-neighbour results do not depend on the split choices (only
 * exact ties can).  max_leaf_size is accepted for source compatibility with KdTreeBuilder::max_leaf_size and ignored. */
int sgb_target_build_kdtree(sgb_ctx* ctx, int max_leaf_size);
/* Gaussian voxel map target (VGICP): replaces IncrementalVoxelMap<GaussianVoxel>::nearest_neighbor_search
 * (ann/incremental_voxelmap.hpp:99-119) and its point/cov traits (:207-222).  Voxel i of the arrays is
 * flat_voxels[i]; reported correspondences are (i << 32) like calc_index (:151).
 * search_offsets is 1, 7 or 27 (:157-186). */
int sgb_target_set_voxelmap(sgb_ctx* ctx, double leaf_size, size_t n_voxels, const int32_t* coords_xyz, const double* means_xyz1,
                            const double* covs_4x4, int search_offsets);
/* Build the Gaussian voxel map on the device from raw points (+ 4x4 covariances): replaces the one-shot use of
 * IncrementalVoxelMap<GaussianVoxel>::insert (ann/incremental_voxelmap.hpp:55-92) with GaussianVoxel::add / finalize
 * (ann/gaussian_voxelmap.hpp:30-62): voxel = floor(p / leaf_size), mean = sum p / n, cov = sum C / n.  Voxel ids are in
 * ascending key order (the reference numbers them in first-insertion order; only the ids differ). */
int sgb_target_build_voxelmap(sgb_ctx* ctx, size_t n, const double* points_xyz1, const double* covs_4x4 /*or NULL*/, double leaf_size,
                              int search_offsets);
/* k nearest neighbours (1 <= k <= 32) of arbitrary query points in the target's kd-tree: replaces KdTree::knn_search /
 * nearest_neighbor_search (ann/kdtree.hpp:165-189,254-274) and the batch_knn_search / batch_nearest_neighbor_search of the
 * Python binding (src/python/kdtree.cpp).  out_indices / out_sq_dists: n_queries x k, ascending distance, original target
 * indices; UINT64_MAX / DBL_MAX pad when the target has fewer than k points (knn_result.hpp:60-66). */
int sgb_target_batch_knn(sgb_ctx* ctx, size_t n_queries, const double* queries_xyz1, int k, uint64_t* out_indices, double* out_sq_dists);
size_t sgb_target_size(const sgb_ctx* ctx);

/* ---- source: replaces traits::point/cov(source, i) ------------------------------------------- */
int sgb_source_set_points(sgb_ctx* ctx, size_t n, const double* points_xyz1, const double* covs_4x4 /*or NULL*/);
size_t sgb_source_size(const sgb_ctx* ctx);

/* ---- the hot path ------------------------------------------------------------------------------
 * sgb_linearize replaces {Serial,ParallelReductionOMP,ParallelReductionTBB}::linearize
 * (registration/reduction.hpp:20-47, reduction_omp.hpp:24-59, reduction_tbb.hpp:117-131): for every
 * source point transform -> nearest neighbour -> reject -> factor -> sum.  out_Hbe = H(36) | b(6) | e(1).
 * The correspondences (and, for GICP, the linearisation pose from which the fused precision matrix is
 * re-derived) stay cached in the context for sgb_error(), as the reference caches them in its factor
 * vector (gicp_factor.hpp:94-96). */
int sgb_linearize(sgb_ctx* ctx, int factor_kind, int robust_kind, double robust_c, int rejector_kind, double max_dist_sq,
                  const double* T_colmajor16, double* out_Hbe43);
/* replaces {...}::error (reduction.hpp:55-62, reduction_omp.hpp:61-70, reduction_tbb.hpp:133-138):
 * sum of factor errors at a trial pose with the cached correspondences (no new search). */
int sgb_error(sgb_ctx* ctx, const double* T_colmajor16, double* out_e);
/* Same, but asynchronous on the context's stream with the result left in DEVICE memory
 * (d_out44 = H|b|e|num_inliers as doubles; d_out1 = e) -- for multi-GPU all-reduce without a host hop. */
int sgb_linearize_device(sgb_ctx* ctx, int factor_kind, int robust_kind, double robust_c, int rejector_kind, double max_dist_sq,
                         const double* T_colmajor16, double* d_out44);
int sgb_error_device(sgb_ctx* ctx, const double* T_colmajor16, double* d_out1);

/* Forget the correspondences of the previous linearize as search seeds: the next sgb_linearize searches like the first iteration of a
 * fresh align() (the seeds only ever prune -- results are identical with or without them; this exists so that a benchmark can time an
 * unseeded first iteration without re-uploading the source). */
int sgb_drop_seeds(sgb_ctx* ctx);

/* factors[i].target_index of the last linearize, in the caller's source order
 * (SGB_NO_CORRESPONDENCE = rejected; icp_factor.hpp:66-69). */
int sgb_correspondences(sgb_ctx* ctx, uint64_t* target_index);
/* count_if(factors, inlier()) of the last linearize (registration/optimizer.hpp:60,146). */
int sgb_num_inliers(sgb_ctx* ctx, size_t* n);

/* ---- per-cloud preparation on the device (not per-iteration) --------------------
 * sgb_estimate_features replaces estimate_normals / estimate_covariances / estimate_normals_covariances and
 * their _omp / _tbb variants (include/small_gicp/util/normal_estimation.hpp:12-140, normal_estimation_omp.hpp:9-60):
 * k nearest neighbours of every point in the cloud's own kd-tree (the point itself included), local covariance,
 * smallest-eigenvalue direction -> normal (flipped toward the origin, w = 0) and covariance with eigenvalues replaced
 * by (1e-3, 1, 1); fewer than 5 neighbours -> zero normal / identity covariance.  Either output may be NULL.
 * num_neighbors in 1..32. */
int sgb_estimate_features(sgb_ctx* ctx, size_t n, const double* points_xyz1, int num_neighbors, double* out_normals_xyz0 /*or NULL*/,
                          double* out_covs_4x4 /*or NULL*/);
/* Device-resident variants: fill the normals + covariances of the CURRENT target (after its kd-tree is set/built) or
 * the covariances of the CURRENT source without any host round trip (frame streams, benchmark_odom.hpp:59 +
 * odometry_benchmark_small_gicp_tbb.cpp:26-27). */
int sgb_target_estimate_features(sgb_ctx* ctx, int num_neighbors);
int sgb_source_estimate_features(sgb_ctx* ctx, int num_neighbors);
/* Frame streams (odometry_benchmark_small_gicp_tbb.cpp:41-43: `target_points = points; target_tree = tree;`): the CURRENT SOURCE becomes the
 * target -- its points, the kd-tree sgb_source_estimate_features built over them and its covariances are taken over on the device (a tree
 * is built only if none exists; host-supplied covariances are taken over too), the grid front end is built, and the context has no source
 * until the next sgb_source_set_points.  No normals: call sgb_target_estimate_features for point-to-plane. */
int sgb_target_adopt_source(sgb_ctx* ctx);
/* replaces voxelgrid_sampling{,_omp,_tbb} (include/small_gicp/util/downsampling.hpp:22-78): one output point per occupied
 * voxel = mean of its points, voxels ordered by the reference's 63-bit key (x | y<<21 | z<<42 of floor(p/leaf)+2^20);
 * points whose voxel coordinate leaves the 21-bit range are dropped.  out_points_xyz1 must hold n points. */
int sgb_voxelgrid_sampling(sgb_ctx* ctx, size_t n, const double* points_xyz1, double leaf_size, double* out_points_xyz1, size_t* n_out);

#ifdef __cplusplus
}
#endif
#endif /* SGICP_B200_H_ */
