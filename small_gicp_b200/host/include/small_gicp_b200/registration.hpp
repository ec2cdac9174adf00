// SPDX-License-Identifier: MIT
// Host-side registration driver with the reference's template surface:
//   Registration<PointFactor, Reduction, GeneralFactor, CorrespondenceRejector, Optimizer>::align()
//       /root/reference/include/small_gicp/registration/registration.hpp:17-54
//   GaussNewtonOptimizer / LevenbergMarquardtOptimizer   .../registration/optimizer.hpp:10-158
//   TerminationCriteria                                   .../registration/termination_criteria.hpp:10-20
//   RegistrationResult                                    .../registration/registration_result.hpp:11-30
// The optimizers only ever touch the 6x6 system; everything per-point happens inside Reduction.
// Two hooks let a device-resident reduction avoid per-iteration host work (both optional, detected at
// compile time): reduction.sync_factors(factors) after the last iteration, reduction.num_inliers().
#pragma once
#include <algorithm>
#include <iostream>
#include <tuple>
#include <type_traits>
#include <vector>

#include "core.hpp"
#include "factors.hpp"

namespace small_gicp_b200 {

struct RegistrationResult {
  explicit RegistrationResult(const Isometry3d& T = Isometry3d::Identity()) : T_target_source(T) {}
  Isometry3d T_target_source;  ///< estimated transformation (source -> target)
  bool converged = false;
  size_t iterations = 0;   ///< index of the last outer iteration that ran
  size_t num_inliers = 0;  ///< source points with an accepted correspondence at the last linearisation
  Matrix6d H;              ///< final information matrix
  Vector6d b;              ///< final information vector
  double error = 0.0;      ///< final error
};

struct TerminationCriteria {
  TerminationCriteria() : translation_eps(1e-3), rotation_eps(0.1 * M_PI / 180.0) {}
  /// both the rotational and the translational part of the update must be below their tolerances
  bool converged(const Vector6d& delta) const { return delta.head<3>().norm() <= rotation_eps && delta.tail<3>().norm() <= translation_eps; }
  double translation_eps;  ///< [m]
  double rotation_eps;     ///< [rad]
};

namespace detail {
template <typename R, typename F, typename = void>
struct has_sync_factors : std::false_type {};
template <typename R, typename F>
struct has_sync_factors<R, F, std::void_t<decltype(std::declval<R&>().sync_factors(std::declval<std::vector<F>&>()))>> : std::true_type {};

/// Bring the per-point state back to the host once, then count the accepted correspondences.
template <typename Reduction, typename Factor>
size_t finish_factors(Reduction& reduction, std::vector<Factor>& factors) {
  if constexpr (has_sync_factors<Reduction, Factor>::value) reduction.sync_factors(factors);
  return static_cast<size_t>(std::count_if(factors.begin(), factors.end(), [](const Factor& f) { return f.inlier(); }));
}

inline Vector6d damped_step(const Matrix6d& H, const Vector6d& b, double lambda) {
  Matrix6d A = H;
  for (int i = 0; i < 6; i++) A(i, i) += lambda;
  return solve_ldlt(A, -b);
}
}  // namespace detail

/// Gauss-Newton with a small constant damping.
struct GaussNewtonOptimizer {
  GaussNewtonOptimizer() : verbose(false), max_iterations(20), lambda(1e-6) {}

  template <typename Target, typename Source, typename Tree, typename Rejector, typename Criteria, typename Reduction, typename Factor, typename GeneralFactor>
  RegistrationResult optimize(const Target& target, const Source& source, const Tree& tree, const Rejector& rejector, const Criteria& criteria, Reduction& reduction,
                              const Isometry3d& init_T, std::vector<Factor>& factors, GeneralFactor& general_factor) const {
    RegistrationResult result(init_T);
    for (int it = 0; it < max_iterations && !result.converged; it++) {
      auto [H, b, e] = reduction.linearize(target, source, tree, rejector, result.T_target_source, factors);
      general_factor.update_linearized_system(target, source, tree, result.T_target_source, &H, &b, &e);
      const Vector6d delta = detail::damped_step(H, b, lambda);
      if (verbose) std::cout << "gn iter=" << it << " e=" << e << " |dr|=" << delta.head<3>().norm() << " |dt|=" << delta.tail<3>().norm() << std::endl;
      result.converged = criteria.converged(delta);
      result.T_target_source = result.T_target_source * se3_exp(delta);
      result.iterations = it;
      result.H = H;
      result.b = b;
      result.error = e;
    }
    result.num_inliers = detail::finish_factors(reduction, factors);
    return result;
  }

  bool verbose;
  int max_iterations;
  double lambda;
};

/// Levenberg-Marquardt: one linearisation per outer iteration, then up to max_inner_iterations damped trials,
/// each scored with Reduction::error at the trial pose (cached correspondences).
struct LevenbergMarquardtOptimizer {
  LevenbergMarquardtOptimizer() : verbose(false), max_iterations(20), max_inner_iterations(10), init_lambda(1e-3), lambda_factor(10.0) {}

  template <typename Target, typename Source, typename Tree, typename Rejector, typename Criteria, typename Reduction, typename Factor, typename GeneralFactor>
  RegistrationResult optimize(const Target& target, const Source& source, const Tree& tree, const Rejector& rejector, const Criteria& criteria, Reduction& reduction,
                              const Isometry3d& init_T, std::vector<Factor>& factors, GeneralFactor& general_factor) const {
    RegistrationResult result(init_T);
    double lambda = init_lambda;
    for (int it = 0; it < max_iterations && !result.converged; it++) {
      auto [H, b, e] = reduction.linearize(target, source, tree, rejector, result.T_target_source, factors);
      general_factor.update_linearized_system(target, source, tree, result.T_target_source, &H, &b, &e);
      bool accepted = false;
      for (int trial = 0; trial < max_inner_iterations; trial++) {
        const Vector6d delta = detail::damped_step(H, b, lambda);
        const Isometry3d candidate = result.T_target_source * se3_exp(delta);
        const double candidate_e = reduction.error(target, source, candidate, factors);
        general_factor.update_error(target, source, candidate, &e);  // sic: the reference passes &e here (optimizer.hpp:114)
        if (verbose) std::cout << "lm iter=" << it << " trial=" << trial << " e=" << e << " new_e=" << candidate_e << " lambda=" << lambda << std::endl;
        if (candidate_e <= e) {
          result.converged = criteria.converged(delta);
          result.T_target_source = candidate;
          lambda /= lambda_factor;
          e = candidate_e;
          accepted = true;
          break;
        }
        lambda *= lambda_factor;
      }
      result.iterations = it;
      result.H = H;
      result.b = b;
      result.error = e;
      if (!accepted) break;
    }
    result.num_inliers = detail::finish_factors(reduction, factors);
    return result;
  }

  bool verbose;
  int max_iterations;
  int max_inner_iterations;
  double init_lambda;
  double lambda_factor;
};

template <typename PointFactor, typename Reduction, typename GeneralFactor = NullFactor, typename CorrespondenceRejector = DistanceRejector,
          typename Optimizer = LevenbergMarquardtOptimizer>
struct Registration {
  using PointFactorSetting = typename PointFactor::Setting;

  template <typename Target, typename Source, typename Tree>
  RegistrationResult align(const Target& target, const Source& source, const Tree& target_tree, const Isometry3d& init_T = Isometry3d::Identity()) const {
    if (traits::size(target) <= 10) std::cerr << "warning: target point cloud is too small. |target|=" << traits::size(target) << std::endl;
    if (traits::size(source) <= 10) std::cerr << "warning: source point cloud is too small. |source|=" << traits::size(source) << std::endl;
    std::vector<PointFactor> factors(traits::size(source), PointFactor(point_factor));
    return optimizer.optimize(target, source, target_tree, rejector, criteria, reduction, init_T, factors, general_factor);
  }

  TerminationCriteria criteria;
  CorrespondenceRejector rejector;
  PointFactorSetting point_factor;
  GeneralFactor general_factor;
  Reduction reduction;
  Optimizer optimizer;
};

}  // namespace small_gicp_b200
