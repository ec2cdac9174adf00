// SPDX-License-Identifier: MIT
// libsgicp_b200_host.so -- a thin C entry point over the C++ host mirror so that the Python tests can
// drive  Registration<Factor, ParallelReductionCUDA, GeneralFactor, Rejector, Optimizer>::align()
// exactly as a C++ user of the reference would (the template instantiations below are the ones the
// reference's registration_test.cpp:284-292 and registration_helper.cpp:81-137 exercise).
#include <cstring>
#include <memory>
#include <string>

#include <small_gicp_b200/core.hpp>
#include <small_gicp_b200/factors.hpp>
#include <small_gicp_b200/kdtree.hpp>
#include <small_gicp_b200/reduction_cuda.hpp>
#include <small_gicp_b200/registration.hpp>
#include <small_gicp_b200/registration_helper.hpp>
#include <small_gicp_b200/voxelmap.hpp>

using namespace small_gicp_b200;

namespace {
thread_local std::string g_error;

struct Options {
  int factor;          // sgb_factor_kind
  int robust;          // sgb_robust_kind
  double robust_c;
  int rejector;        // sgb_rejector_kind
  double max_dist_sq;
  int optimizer;       // 0 = GaussNewton, 1 = LevenbergMarquardt
  int max_iterations;
  double rotation_eps, translation_eps;
  int tree;            // 0 = host KdTree (reference-layout nodes), 1 = DeviceKdTree, 2 = GaussianVoxelMap target
  double voxel_resolution;
  int voxel_search_offsets;
  int restrict_dof;    // 0 = NullFactor, 1 = RestrictDoFFactor with the mask below
  double dof_mask[6];
  int device;
};

template <typename Reg>
void configure(Reg& reg, const Options& o) {
  reg.criteria.rotation_eps = o.rotation_eps;
  reg.criteria.translation_eps = o.translation_eps;
  reg.optimizer.max_iterations = o.max_iterations;
  reg.reduction.device = o.device;
}
template <typename Reg>
void set_rejector(Reg& reg, const Options& o, std::true_type) {
  reg.rejector.max_dist_sq = o.max_dist_sq;
}
template <typename Reg>
void set_rejector(Reg&, const Options&, std::false_type) {}
template <typename Reg>
void set_general(Reg& reg, const Options& o, std::true_type) {
  reg.general_factor.set_rotation_mask(o.dof_mask[0], o.dof_mask[1], o.dof_mask[2]);
  reg.general_factor.set_translation_mask(o.dof_mask[3], o.dof_mask[4], o.dof_mask[5]);
}
template <typename Reg>
void set_general(Reg&, const Options&, std::false_type) {}
template <typename Setting>
void set_robust(Setting& s, const Options& o, std::true_type) {
  s.robust_kernel.c = o.robust_c;
}
template <typename Setting>
void set_robust(Setting&, const Options&, std::false_type) {}

template <typename Factor, typename General, typename Rejector, typename Optimizer, bool kRobust>
RegistrationResult run(const Options& o, const std::shared_ptr<PointCloud>& target, const std::shared_ptr<PointCloud>& source, const Isometry3d& init) {
  Registration<Factor, ParallelReductionCUDA, General, Rejector, Optimizer> reg;
  configure(reg, o);
  set_rejector(reg, o, std::is_same<Rejector, DistanceRejector>());
  set_general(reg, o, std::is_same<General, RestrictDoFFactor>());
  set_robust(reg.point_factor, o, std::integral_constant<bool, kRobust>());
  if (o.tree == 2) {
    GaussianVoxelMap map(o.voxel_resolution);
    map.set_search_offsets(o.voxel_search_offsets);
    map.insert(*target);
    return reg.align(map, *source, map, init);
  }
  if (o.tree == 1) {
    DeviceKdTree<PointCloud> tree(target);
    return reg.align(*target, *source, tree, init);
  }
  KdTree<PointCloud> tree(target);
  return reg.align(*target, *source, tree, init);
}

template <typename Factor, bool kRobust>
RegistrationResult dispatch_rest(const Options& o, const std::shared_ptr<PointCloud>& t, const std::shared_ptr<PointCloud>& s, const Isometry3d& init) {
  const int key = (o.restrict_dof ? 4 : 0) | (o.rejector == SGB_REJECT_DISTANCE ? 2 : 0) | (o.optimizer ? 1 : 0);
  switch (key) {
    case 0: return run<Factor, NullFactor, NullRejector, GaussNewtonOptimizer, kRobust>(o, t, s, init);
    case 1: return run<Factor, NullFactor, NullRejector, LevenbergMarquardtOptimizer, kRobust>(o, t, s, init);
    case 2: return run<Factor, NullFactor, DistanceRejector, GaussNewtonOptimizer, kRobust>(o, t, s, init);
    case 3: return run<Factor, NullFactor, DistanceRejector, LevenbergMarquardtOptimizer, kRobust>(o, t, s, init);
    case 4: return run<Factor, RestrictDoFFactor, NullRejector, GaussNewtonOptimizer, kRobust>(o, t, s, init);
    case 5: return run<Factor, RestrictDoFFactor, NullRejector, LevenbergMarquardtOptimizer, kRobust>(o, t, s, init);
    case 6: return run<Factor, RestrictDoFFactor, DistanceRejector, GaussNewtonOptimizer, kRobust>(o, t, s, init);
    default: return run<Factor, RestrictDoFFactor, DistanceRejector, LevenbergMarquardtOptimizer, kRobust>(o, t, s, init);
  }
}

template <typename Base>
RegistrationResult dispatch_robust(const Options& o, const std::shared_ptr<PointCloud>& t, const std::shared_ptr<PointCloud>& s, const Isometry3d& init) {
  if (o.robust == SGB_ROBUST_HUBER) return dispatch_rest<RobustFactor<Huber, Base>, true>(o, t, s, init);
  if (o.robust == SGB_ROBUST_CAUCHY) return dispatch_rest<RobustFactor<Cauchy, Base>, true>(o, t, s, init);
  return dispatch_rest<Base, false>(o, t, s, init);
}

std::shared_ptr<PointCloud> make_cloud(size_t n, const double* pts4, const double* normals4, const double* covs16) {
  auto c = std::make_shared<PointCloud>();
  c->resize(n);
  if (n) {
    std::memcpy(c->points[0].data(), pts4, n * sizeof(Vector4d));
    if (normals4) std::memcpy(c->normals[0].data(), normals4, n * sizeof(Vector4d));
    if (covs16) std::memcpy(c->covs[0].data(), covs16, n * sizeof(Matrix4d));
  }
  if (!normals4) c->normals.clear();
  if (!covs16) c->covs.clear();
  return c;
}
}  // namespace

extern "C" {

const char* sgbh_last_error() { return g_error.c_str(); }

/// options: 21 doubles in the order of `Options` (ints passed as doubles).  init_T / T_out: column-major 4x4.
/// scalars_out: [converged, iterations, num_inliers, error].
int sgbh_align(size_t n_target, const double* target_pts4, const double* target_normals4, const double* target_covs16, size_t n_source,
               const double* source_pts4, const double* source_covs16, const double* options21, const double* init_T16, double* T_out16, double* scalars_out4,
               double* H_out36, double* b_out6) {
  try {
    Options o;
    const double* q = options21;
    o.factor = static_cast<int>(q[0]);
    o.robust = static_cast<int>(q[1]);
    o.robust_c = q[2];
    o.rejector = static_cast<int>(q[3]);
    o.max_dist_sq = q[4];
    o.optimizer = static_cast<int>(q[5]);
    o.max_iterations = static_cast<int>(q[6]);
    o.rotation_eps = q[7];
    o.translation_eps = q[8];
    o.tree = static_cast<int>(q[9]);
    o.voxel_resolution = q[10];
    o.voxel_search_offsets = static_cast<int>(q[11]);
    o.restrict_dof = static_cast<int>(q[12]);
    for (int i = 0; i < 6; i++) o.dof_mask[i] = q[13 + i];
    o.device = static_cast<int>(q[19]);
    auto target = make_cloud(n_target, target_pts4, target_normals4, target_covs16);
    auto source = make_cloud(n_source, source_pts4, nullptr, source_covs16);
    Isometry3d init;
    std::memcpy(init.matrix().data(), init_T16, sizeof(double) * 16);
    RegistrationResult r;
    switch (o.factor) {
      case SGB_FACTOR_ICP: r = dispatch_robust<ICPFactor>(o, target, source, init); break;
      case SGB_FACTOR_PLANE_ICP: r = dispatch_robust<PointToPlaneICPFactor>(o, target, source, init); break;
      case SGB_FACTOR_GICP: r = dispatch_robust<GICPFactor>(o, target, source, init); break;
      default: g_error = "sgbh_align: invalid factor kind"; return 1;
    }
    std::memcpy(T_out16, r.T_target_source.matrix().data(), sizeof(double) * 16);
    scalars_out4[0] = r.converged ? 1.0 : 0.0;
    scalars_out4[1] = static_cast<double>(r.iterations);
    scalars_out4[2] = static_cast<double>(r.num_inliers);
    scalars_out4[3] = r.error;
    if (H_out36) std::memcpy(H_out36, r.H.data(), sizeof(double) * 36);
    if (b_out6) std::memcpy(b_out6, r.b.data(), sizeof(double) * 6);
    return 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return 2;
  }
}

/// small_gicp::align(target_points, source_points, init_T, setting) of the helper API (registration_helper.cpp:58-69):
/// raw points in; settings8 = [type, voxel_resolution, downsampling_resolution, max_correspondence_distance, rotation_eps,
/// translation_eps, max_iterations, device].  sizes_out2 = down-sampled target / source sizes (informational).
int sgbh_helper_align(size_t n_target, const double* target_pts4, size_t n_source, const double* source_pts4, const double* settings8, const double* init_T16,
                      double* T_out16, double* scalars_out4, double* sizes_out2) {
  try {
    RegistrationSetting s;
    s.type = static_cast<RegistrationSetting::RegistrationType>(static_cast<int>(settings8[0]));
    s.voxel_resolution = settings8[1];
    s.downsampling_resolution = settings8[2];
    s.max_correspondence_distance = settings8[3];
    s.rotation_eps = settings8[4];
    s.translation_eps = settings8[5];
    s.max_iterations = static_cast<int>(settings8[6]);
    s.device = static_cast<int>(settings8[7]);
    auto target = make_cloud(n_target, target_pts4, nullptr, nullptr);
    auto source = make_cloud(n_source, source_pts4, nullptr, nullptr);
    Isometry3d init;
    std::memcpy(init.matrix().data(), init_T16, sizeof(double) * 16);
    auto [tp, tt] = preprocess_points(*target, s.downsampling_resolution, 10, s.num_threads, s.device);
    auto [sp, st] = preprocess_points(*source, s.downsampling_resolution, 10, s.num_threads, s.device);
    RegistrationResult r;
    if (s.type == RegistrationSetting::VGICP) {
      auto vm = create_gaussian_voxelmap(*tp, s.voxel_resolution);
      r = align(*vm, *sp, init, s);
    } else {
      r = align(*tp, *sp, *tt, init, s);
    }
    std::memcpy(T_out16, r.T_target_source.matrix().data(), sizeof(double) * 16);
    scalars_out4[0] = r.converged ? 1.0 : 0.0;
    scalars_out4[1] = static_cast<double>(r.iterations);
    scalars_out4[2] = static_cast<double>(r.num_inliers);
    scalars_out4[3] = r.error;
    if (sizes_out2) {
      sizes_out2[0] = static_cast<double>(tp->size());
      sizes_out2[1] = static_cast<double>(sp->size());
    }
    return 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return 2;
  }
}

/// Host KdTree<PointCloud>::knn_search over a batch of queries (API-surface check of the mirror, CPU-only).
int sgbh_kdtree_knn(size_t n, const double* pts4, size_t nq, const double* queries4, int k, uint64_t* idx_out, double* d2_out) {
  try {
    auto cloud = make_cloud(n, pts4, nullptr, nullptr);
    KdTree<PointCloud> tree(cloud);
    std::vector<size_t> ki(k);
    for (size_t i = 0; i < nq; i++) {
      Vector4d q;
      std::memcpy(q.data(), queries4 + i * 4, sizeof(double) * 4);
      tree.knn_search(q, k, ki.data(), d2_out + i * k);
      for (int j = 0; j < k; j++) idx_out[i * k + j] = ki[j];
    }
    return 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return 2;
  }
}

/// read_ply / read_points of the host mirror (read_points.hpp): fills at most `capacity` points (x, y, z, 1 as floats), returns the
/// number of points in the file through n_out (API-surface check of the mirror, CPU-only).  kind: 0 = PLY, 1 = KITTI .bin.
int sgbh_read_points(const char* filename, int kind, size_t capacity, float* out_xyz1, size_t* n_out) {
  try {
    const std::vector<Vector4f> pts = kind == 0 ? read_ply(filename) : read_points(filename);
    *n_out = pts.size();
    const size_t m = pts.size() < capacity ? pts.size() : capacity;
    if (m) std::memcpy(out_xyz1, pts.data(), m * sizeof(Vector4f));
    return 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return 2;
  }
}

/// write_points of the host mirror.
int sgbh_write_points(const char* filename, size_t n, const float* xyz1) {
  try {
    std::vector<Vector4f> pts(n);
    if (n) std::memcpy(pts.data(), xyz1, n * sizeof(Vector4f));
    write_points(filename, pts);
    return 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return 2;
  }
}

}  // extern "C"
