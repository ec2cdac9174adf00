// SPDX-License-Identifier: MIT
// oracle/_ref: the REFERENCE'S OWN headers (included in place from /root/reference/include, never copied) compiled
// against the Eigen API shim in ./eigen_shim, behind the same C API as the oracle (oracle/sgicp_oracle.cpp), so that
// tests can pin the oracle's restatement against the reference's code itself.  TEST INFRASTRUCTURE ONLY; built only
// where /root/reference exists (this container), never on the GPU box.
#include <cstring>
#include <memory>
#include <vector>

#include <small_gicp/ann/gaussian_voxelmap.hpp>
#include <small_gicp/ann/kdtree.hpp>
#include <small_gicp/factors/gicp_factor.hpp>
#include <small_gicp/factors/icp_factor.hpp>
#include <small_gicp/factors/plane_icp_factor.hpp>
#include <small_gicp/factors/robust_kernel.hpp>
#include <small_gicp/points/point_cloud.hpp>
#include <small_gicp/registration/reduction.hpp>
#include <small_gicp/registration/reduction_omp.hpp>
#include <small_gicp/registration/registration.hpp>
#include <small_gicp/util/downsampling.hpp>
#include <small_gicp/util/lie.hpp>
#include <small_gicp/util/normal_estimation.hpp>
#include <small_gicp/util/normal_estimation_omp.hpp>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace small_gicp;
using Tree = UnsafeKdTree<PointCloud>;

namespace {

struct Reg {
  int factor = 2, robust = 0, rejector = 1, num_threads = 0;
  double robust_c = 1.0, max_dist_sq = 1.0;
  int opt_type = 1, max_iterations = 20, max_inner = 10;
  double gn_lambda = 1e-6, init_lambda = 1e-3, lambda_factor = 10.0, rot_eps = 0.1 * M_PI / 180.0, trans_eps = 1e-3;
  std::shared_ptr<void> factors;  // std::vector<Factor> of the current factor type
  size_t n_factors = 0;
  int factors_key = -1;
  std::vector<uint64_t> corr;
};

Eigen::Isometry3d iso_from_rowmajor(const double* T16) {
  Eigen::Isometry3d T;
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) T.matrix()(r, c) = T16[r * 4 + c];
  return T;
}
void iso_to_rowmajor(const Eigen::Isometry3d& T, double* T16) {
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) T16[r * 4 + c] = T.matrix()(r, c);
}

template <typename F>
F make_factor(const Reg&, const F*) {
  return F();
}
template <typename K, typename F>
RobustFactor<K, F> make_factor(const Reg& r, const RobustFactor<K, F>*) {
  typename RobustFactor<K, F>::Setting s;
  s.robust_kernel.c = r.robust_c;
  return RobustFactor<K, F>(s);
}

/// call fn(Factor prototype) for the runtime (factor, robust) pair
template <typename Fn>
void with_factor(const Reg& r, Fn&& fn) {
  auto robustify = [&](auto base) {
    using B = decltype(base);
    if (r.robust == 1)
      fn(make_factor(r, static_cast<RobustFactor<Huber, B>*>(nullptr)));
    else if (r.robust == 2)
      fn(make_factor(r, static_cast<RobustFactor<Cauchy, B>*>(nullptr)));
    else
      fn(B());
  };
  if (r.factor == 0)
    robustify(ICPFactor());
  else if (r.factor == 1)
    robustify(PointToPlaneICPFactor());
  else
    robustify(GICPFactor());
}

template <typename Factor>
std::vector<Factor>& factor_vector(Reg& r, size_t n, const Factor& proto, bool reset) {
  const int key = r.factor * 3 + r.robust;
  if (!r.factors || r.factors_key != key || r.n_factors != n || reset) {
    r.factors = std::make_shared<std::vector<Factor>>(n, proto);
    r.factors_key = key;
    r.n_factors = n;
  }
  return *static_cast<std::vector<Factor>*>(r.factors.get());
}

template <typename Factor>
void store_corr(Reg& r, const std::vector<Factor>& f) {
  r.corr.resize(f.size());
  for (size_t i = 0; i < f.size(); i++) {
    size_t k;
    if constexpr (std::is_same_v<Factor, ICPFactor> || std::is_same_v<Factor, PointToPlaneICPFactor> || std::is_same_v<Factor, GICPFactor>)
      k = f[i].target_index;
    else
      k = f[i].factor.target_index;
    r.corr[i] = k;
  }
}

template <typename Target, typename TargetTree, typename Factor>
void do_linearize(Reg& r, const Target& target, const TargetTree& tree, const PointCloud& source, const Eigen::Isometry3d& T, const Factor& proto, double* out43) {
  auto& factors = factor_vector(r, source.size(), proto, false);
  Eigen::Matrix<double, 6, 6> H;
  Eigen::Matrix<double, 6, 1> b;
  double e = 0.0;
  auto run = [&](const auto& rejector) {
    if (r.num_threads <= 0) {
      SerialReduction red;
      std::tie(H, b, e) = red.linearize(target, source, tree, rejector, T, factors);
    } else {
      ParallelReductionOMP red;
      red.num_threads = r.num_threads;
      std::tie(H, b, e) = red.linearize(target, source, tree, rejector, T, factors);
    }
  };
  if (r.rejector == 1) {
    DistanceRejector rej;
    rej.max_dist_sq = r.max_dist_sq;
    run(rej);
  } else {
    run(NullRejector());
  }
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) out43[i * 6 + j] = H(i, j);
  for (int i = 0; i < 6; i++) out43[36 + i] = b[i];
  out43[42] = e;
  store_corr(r, factors);
}

template <typename Target, typename Factor>
double do_error(Reg& r, const Target& target, const PointCloud& source, const Eigen::Isometry3d& T, const Factor& proto) {
  auto& factors = factor_vector(r, source.size(), proto, false);
  if (r.num_threads <= 0) return SerialReduction().error(target, source, T, factors);
  ParallelReductionOMP red;
  red.num_threads = r.num_threads;
  return red.error(target, source, T, factors);
}

template <typename Factor, typename Reduction, typename Rejector, typename Optimizer, typename Target, typename TargetTree>
RegistrationResult do_align(const Reg& r, const Target& target, const TargetTree& tree, const PointCloud& source, const Eigen::Isometry3d& init, const Factor& proto) {
  Registration<Factor, Reduction, NullFactor, Rejector, Optimizer> reg;
  if constexpr (std::is_same_v<Reduction, ParallelReductionOMP>) reg.reduction.num_threads = r.num_threads;
  if constexpr (std::is_same_v<Rejector, DistanceRejector>) reg.rejector.max_dist_sq = r.max_dist_sq;
  reg.criteria.rotation_eps = r.rot_eps;
  reg.criteria.translation_eps = r.trans_eps;
  reg.optimizer.max_iterations = r.max_iterations;
  if constexpr (std::is_same_v<Optimizer, GaussNewtonOptimizer>) {
    reg.optimizer.lambda = r.gn_lambda;
  } else {
    reg.optimizer.max_inner_iterations = r.max_inner;
    reg.optimizer.init_lambda = r.init_lambda;
    reg.optimizer.lambda_factor = r.lambda_factor;
  }
  if constexpr (!std::is_same_v<Factor, ICPFactor> && !std::is_same_v<Factor, PointToPlaneICPFactor> && !std::is_same_v<Factor, GICPFactor>)
    reg.point_factor.robust_kernel.c = r.robust_c;
  return reg.align(target, source, tree, init);
}

template <typename Target, typename TargetTree, typename Factor>
RegistrationResult dispatch_align(const Reg& r, const Target& target, const TargetTree& tree, const PointCloud& source, const Eigen::Isometry3d& init, const Factor& proto) {
  const int key = (r.num_threads > 0 ? 4 : 0) | (r.rejector == 1 ? 2 : 0) | (r.opt_type ? 1 : 0);
  switch (key) {
    case 0: return do_align<Factor, SerialReduction, NullRejector, GaussNewtonOptimizer>(r, target, tree, source, init, proto);
    case 1: return do_align<Factor, SerialReduction, NullRejector, LevenbergMarquardtOptimizer>(r, target, tree, source, init, proto);
    case 2: return do_align<Factor, SerialReduction, DistanceRejector, GaussNewtonOptimizer>(r, target, tree, source, init, proto);
    case 3: return do_align<Factor, SerialReduction, DistanceRejector, LevenbergMarquardtOptimizer>(r, target, tree, source, init, proto);
    case 4: return do_align<Factor, ParallelReductionOMP, NullRejector, GaussNewtonOptimizer>(r, target, tree, source, init, proto);
    case 5: return do_align<Factor, ParallelReductionOMP, NullRejector, LevenbergMarquardtOptimizer>(r, target, tree, source, init, proto);
    case 6: return do_align<Factor, ParallelReductionOMP, DistanceRejector, GaussNewtonOptimizer>(r, target, tree, source, init, proto);
    default: return do_align<Factor, ParallelReductionOMP, DistanceRejector, LevenbergMarquardtOptimizer>(r, target, tree, source, init, proto);
  }
}

}  // namespace

extern "C" {

int orc_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// ---- clouds ----
PointCloud* orc_cloud_create(size_t n, const double* xyz, int stride) {
  auto* c = new PointCloud();
  c->resize(n);
  for (size_t i = 0; i < n; i++) {
    c->point(i) = Eigen::Vector4d(xyz[i * stride + 0], xyz[i * stride + 1], xyz[i * stride + 2], 1.0);
    c->normal(i) = Eigen::Vector4d::Zero();
    c->cov(i) = Eigen::Matrix4d::Zero();
  }
  return c;
}
void orc_cloud_destroy(PointCloud* c) { delete c; }
size_t orc_cloud_size(const PointCloud* c) { return c->size(); }
void orc_cloud_get(const PointCloud* c, double* points4, double* normals4, double* covs16) {
  const size_t n = c->size();
  for (size_t i = 0; i < n; i++) {
    if (points4) std::memcpy(points4 + i * 4, c->point(i).data(), 32);
    if (normals4) std::memcpy(normals4 + i * 4, c->normal(i).data(), 32);
    if (covs16)
      for (int r = 0; r < 4; r++)
        for (int col = 0; col < 4; col++) covs16[i * 16 + r * 4 + col] = c->cov(i)(r, col);
  }
}
void orc_cloud_set_features(PointCloud* c, const double* normals4, const double* covs16) {
  const size_t n = c->size();
  for (size_t i = 0; i < n; i++) {
    if (normals4) std::memcpy(c->normal(i).data(), normals4 + i * 4, 32);
    if (covs16)
      for (int r = 0; r < 4; r++)
        for (int col = 0; col < 4; col++) c->cov(i)(r, col) = covs16[i * 16 + r * 4 + col];
  }
}
PointCloud* orc_cloud_transformed(const PointCloud* c, const double* T16) {
  const Eigen::Isometry3d T = iso_from_rowmajor(T16);
  auto* o = new PointCloud();
  o->resize(c->size());
  for (size_t i = 0; i < c->size(); i++) {
    o->point(i) = T * c->point(i);
    o->normal(i) = Eigen::Vector4d::Zero();
    o->cov(i) = Eigen::Matrix4d::Zero();
  }
  return o;
}
PointCloud* orc_voxelgrid_sampling(const PointCloud* c, double leaf) {
  auto out = voxelgrid_sampling(*c, leaf);
  auto* o = new PointCloud(*out);
  o->normals.assign(o->size(), Eigen::Vector4d::Zero());
  o->covs.assign(o->size(), Eigen::Matrix4d::Zero());
  return o;
}

// ---- kd-tree ----
Tree* orc_kdtree_create(const PointCloud* c) { return new Tree(*c); }
void orc_kdtree_destroy(Tree* t) { delete t; }
size_t orc_kdtree_num_nodes(const Tree* t) { return t->nodes.size(); }
void orc_kdtree_export(const Tree* t, void* nodes24, uint64_t* indices) {
  static_assert(sizeof(Tree::Node) == 24, "reference node layout");
  std::memcpy(nodes24, t->nodes.data(), t->nodes.size() * sizeof(Tree::Node));
  for (size_t i = 0; i < t->indices.size(); i++) indices[i] = t->indices[i];
}
void orc_kdtree_knn(const Tree* t, size_t nq, const double* q4, int k, uint64_t* idx, double* d2, uint64_t* counts, int num_threads) {
  const std::int64_t N = nq;
#pragma omp parallel for num_threads(num_threads > 0 ? num_threads : 1) schedule(guided, 8)
  for (std::int64_t i = 0; i < N; i++) {
    std::vector<size_t> ki(k);
    const Eigen::Vector4d q(q4[i * 4 + 0], q4[i * 4 + 1], q4[i * 4 + 2], q4[i * 4 + 3]);
    size_t n = 0;
    if (!t->nodes.empty()) {
      n = t->knn_search(q, k, ki.data(), d2 + i * k);
    } else {
      for (int j = 0; j < k; j++) {
        ki[j] = std::numeric_limits<size_t>::max();
        d2[i * k + j] = std::numeric_limits<double>::max();
      }
    }
    for (int j = 0; j < k; j++) idx[i * k + j] = ki[j];
    if (counts) counts[i] = n;
  }
}
void orc_estimate_features(PointCloud* c, const Tree* t, int k, int mode, int num_threads) {
  if (num_threads <= 1) {
    if (mode == 1) estimate_normals(*c, *t, k);
    if (mode == 2) estimate_covariances(*c, *t, k);
    if (mode == 3) estimate_normals_covariances(*c, *t, k);
  } else {
    if (mode == 1) estimate_normals_omp(*c, *t, k, num_threads);
    if (mode == 2) estimate_covariances_omp(*c, *t, k, num_threads);
    if (mode == 3) estimate_normals_covariances_omp(*c, *t, k, num_threads);
  }
}

// ---- Gaussian voxel map ----
GaussianVoxelMap* orc_voxelmap_create(const PointCloud* c, double leaf, int search_offsets) {
  auto* m = new GaussianVoxelMap(leaf);
  m->set_search_offsets(search_offsets);
  m->insert(*c);
  return m;
}
void orc_voxelmap_destroy(GaussianVoxelMap* m) { delete m; }
size_t orc_voxelmap_size(const GaussianVoxelMap* m) { return m->size(); }
void orc_voxelmap_export(const GaussianVoxelMap* m, int32_t* coords3, double* means4, double* covs16, uint64_t* counts) {
  for (size_t i = 0; i < m->size(); i++) {
    const auto& v = *m->flat_voxels[i];
    for (int d = 0; d < 3; d++) coords3[i * 3 + d] = v.first.coord[d];
    std::memcpy(means4 + i * 4, v.second.mean.data(), 32);
    for (int r = 0; r < 4; r++)
      for (int col = 0; col < 4; col++) covs16[i * 16 + r * 4 + col] = v.second.cov(r, col);
    if (counts) counts[i] = v.second.num_points;
  }
}
void orc_voxelmap_nn(const GaussianVoxelMap* m, size_t nq, const double* q4, uint64_t* idx, double* d2, uint64_t* counts) {
  for (size_t i = 0; i < nq; i++) {
    const Eigen::Vector4d q(q4[i * 4 + 0], q4[i * 4 + 1], q4[i * 4 + 2], q4[i * 4 + 3]);
    size_t k = std::numeric_limits<size_t>::max();
    double d = std::numeric_limits<double>::max();
    counts[i] = m->nearest_neighbor_search(q, &k, &d);
    idx[i] = k;
    d2[i] = d;
  }
}

// ---- registration ----
Reg* orc_reg_create(int factor, int robust, double robust_c, int rejector, double max_dist_sq, int num_threads) {
  auto* r = new Reg();
  r->factor = factor;
  r->robust = robust;
  r->robust_c = robust_c;
  r->rejector = rejector;
  r->max_dist_sq = max_dist_sq;
  r->num_threads = num_threads;
  return r;
}
void orc_reg_destroy(Reg* r) { delete r; }
void orc_reg_set_optimizer(Reg* r, int type, int max_iterations, double gn_lambda, int max_inner, double init_lambda, double lambda_factor, double rot_eps, double trans_eps) {
  r->opt_type = type;
  r->max_iterations = max_iterations;
  r->gn_lambda = gn_lambda;
  r->max_inner = max_inner;
  r->init_lambda = init_lambda;
  r->lambda_factor = lambda_factor;
  r->rot_eps = rot_eps;
  r->trans_eps = trans_eps;
}
void orc_reg_linearize(Reg* r, const PointCloud* target, const Tree* tree, const GaussianVoxelMap* vmap, const PointCloud* source, const double* T16, double* out43) {
  const Eigen::Isometry3d T = iso_from_rowmajor(T16);
  with_factor(*r, [&](auto proto) {
    if (vmap) {
      if constexpr (!std::is_same_v<decltype(proto), PointToPlaneICPFactor> && !std::is_same_v<decltype(proto), RobustFactor<Huber, PointToPlaneICPFactor>> &&
                    !std::is_same_v<decltype(proto), RobustFactor<Cauchy, PointToPlaneICPFactor>>)
        do_linearize(*r, *vmap, *vmap, *source, T, proto, out43);
    } else {
      do_linearize(*r, *target, *tree, *source, T, proto, out43);
    }
  });
}
double orc_reg_error(Reg* r, const PointCloud* target, const GaussianVoxelMap* vmap, const PointCloud* source, const double* T16) {
  const Eigen::Isometry3d T = iso_from_rowmajor(T16);
  double e = 0.0;
  with_factor(*r, [&](auto proto) {
    if (vmap) {
      if constexpr (!std::is_same_v<decltype(proto), PointToPlaneICPFactor> && !std::is_same_v<decltype(proto), RobustFactor<Huber, PointToPlaneICPFactor>> &&
                    !std::is_same_v<decltype(proto), RobustFactor<Cauchy, PointToPlaneICPFactor>>)
        e = do_error(*r, *vmap, *source, T, proto);
    } else {
      e = do_error(*r, *target, *source, T, proto);
    }
  });
  return e;
}
void orc_reg_correspondences(const Reg* r, uint64_t* target_index) { std::memcpy(target_index, r->corr.data(), r->corr.size() * sizeof(uint64_t)); }

void orc_reg_align(Reg* r, const PointCloud* target, const Tree* tree, const GaussianVoxelMap* vmap, const PointCloud* source, const double* init_T16, int /*want_trace*/,
                   double* T_out16, double* result_scalars, double* H36, double* b6) {
  const Eigen::Isometry3d init = iso_from_rowmajor(init_T16);
  RegistrationResult res;
  with_factor(*r, [&](auto proto) {
    if (vmap) {
      if constexpr (!std::is_same_v<decltype(proto), PointToPlaneICPFactor> && !std::is_same_v<decltype(proto), RobustFactor<Huber, PointToPlaneICPFactor>> &&
                    !std::is_same_v<decltype(proto), RobustFactor<Cauchy, PointToPlaneICPFactor>>)
        res = dispatch_align(*r, *vmap, *vmap, *source, init, proto);
    } else {
      res = dispatch_align(*r, *target, *tree, *source, init, proto);
    }
  });
  iso_to_rowmajor(res.T_target_source, T_out16);
  result_scalars[0] = res.converged;
  result_scalars[1] = static_cast<double>(res.iterations);
  result_scalars[2] = static_cast<double>(res.num_inliers);
  result_scalars[3] = res.error;
  if (H36)
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 6; j++) H36[i * 6 + j] = res.H(i, j);
  if (b6)
    for (int i = 0; i < 6; i++) b6[i] = res.b[i];
}
size_t orc_reg_trace_rows(const Reg*) { return 0; }
void orc_reg_trace_get(const Reg*, double*) {}

// ---- small algebra exports ----
void orc_se3_exp(const double* a6, double* T16) {
  Eigen::Matrix<double, 6, 1> a;
  for (int i = 0; i < 6; i++) a[i] = a6[i];
  iso_to_rowmajor(se3_exp(a), T16);
}
void orc_ldlt_solve6(const double* A36, const double* b6, double* x6) {
  Eigen::Matrix<double, 6, 6> A;
  Eigen::Matrix<double, 6, 1> b;
  for (int i = 0; i < 6; i++) {
    for (int j = 0; j < 6; j++) A(i, j) = A36[i * 6 + j];
    b[i] = b6[i];
  }
  const Eigen::Matrix<double, 6, 1> x = A.ldlt().solve(b);
  for (int i = 0; i < 6; i++) x6[i] = x[i];
}
void orc_eigen_sym3(const double* A9, double* evals3, double* evecs9) {
  Eigen::Matrix3d A;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) A(i, j) = A9[i * 3 + j];
  Eigen::SelfAdjointEigenSolver<Eigen::Matrix3d> eig;
  eig.computeDirect(A);
  for (int i = 0; i < 3; i++) {
    evals3[i] = eig.eigenvalues()[i];
    for (int j = 0; j < 3; j++) evecs9[i * 3 + j] = eig.eigenvectors()(i, j);
  }
}
}  // extern "C"
