// SPDX-License-Identifier: MIT
// Per-point factor, robust-kernel, rejector and general-factor TYPES with the reference's names and settings.
// On this backend a factor type is a compile-time descriptor: the arithmetic of
//   ICPFactor::linearize/error            /root/reference/include/small_gicp/factors/icp_factor.hpp:20-64
//   PointToPlaneICPFactor::linearize/..   .../factors/plane_icp_factor.hpp:20-69
//   GICPFactor::linearize/error           .../factors/gicp_factor.hpp:35-89
//   RobustFactor<Huber|Cauchy, F>         .../factors/robust_kernel.hpp:11-106
// runs inside the CUDA kernels (small_gicp_b200/csrc/sgb_device.cuh); the objects below only carry the
// per-point state the optimizers read back (target_index / source_index / inlier()).
#pragma once
#include <limits>

#include "../../../../include/sgicp_b200.h"
#include "core.hpp"

namespace small_gicp_b200 {

namespace detail {
struct PointFactorState {
  size_t target_index = std::numeric_limits<size_t>::max();
  size_t source_index = std::numeric_limits<size_t>::max();
  bool inlier() const { return target_index != std::numeric_limits<size_t>::max(); }
};
}  // namespace detail

/// Point-to-point ICP error.
struct ICPFactor : detail::PointFactorState {
  struct Setting {};
  ICPFactor(const Setting& = Setting()) {}
  static constexpr int kind = SGB_FACTOR_ICP;
};
/// Point-to-plane ICP error (element-wise n .* r, as the reference defines it).
struct PointToPlaneICPFactor : detail::PointFactorState {
  struct Setting {};
  PointToPlaneICPFactor(const Setting& = Setting()) {}
  static constexpr int kind = SGB_FACTOR_PLANE_ICP;
};
/// Distribution-to-distribution (GICP) error.
struct GICPFactor : detail::PointFactorState {
  struct Setting {};
  GICPFactor(const Setting& = Setting()) {}
  static constexpr int kind = SGB_FACTOR_GICP;
};

struct Huber {
  struct Setting {
    double c = 1.0;
  };
  Huber() : c(1.0) {}
  Huber(const Setting& s) : c(s.c) {}
  static constexpr int kind = SGB_ROBUST_HUBER;
  double c;
};
struct Cauchy {
  struct Setting {
    double c = 1.0;
  };
  Cauchy() : c(1.0) {}
  Cauchy(const Setting& s) : c(s.c) {}
  static constexpr int kind = SGB_ROBUST_CAUCHY;
  double c;
};

template <typename Kernel, typename Factor>
struct RobustFactor {
  struct Setting {
    typename Kernel::Setting robust_kernel;
    typename Factor::Setting factor;
  };
  RobustFactor(const Setting& s = Setting()) : robust_kernel(s.robust_kernel), factor(s.factor) {}
  bool inlier() const { return factor.inlier(); }
  Kernel robust_kernel;
  Factor factor;
};

/// What the kernel launch needs to know about a factor type.
struct FactorDescriptor {
  int factor_kind;
  int robust_kind;
  double robust_c;
};
template <typename Factor>
struct factor_traits {
  static FactorDescriptor describe(const Factor&) { return {Factor::kind, SGB_ROBUST_NONE, 1.0}; }
  static detail::PointFactorState& state(Factor& f) { return f; }
};
template <typename Kernel, typename Factor>
struct factor_traits<RobustFactor<Kernel, Factor>> {
  static FactorDescriptor describe(const RobustFactor<Kernel, Factor>& f) { return {Factor::kind, Kernel::kind, f.robust_kernel.c}; }
  static detail::PointFactorState& state(RobustFactor<Kernel, Factor>& f) { return f.factor; }
};

// ---- correspondence rejectors (registration/rejector.hpp:11-28) ----
struct NullRejector {
  static constexpr int kind = SGB_REJECT_NONE;
  double threshold() const { return 0.0; }
};
struct DistanceRejector {
  DistanceRejector() : max_dist_sq(1.0) {}
  static constexpr int kind = SGB_REJECT_DISTANCE;
  double threshold() const { return max_dist_sq; }
  double max_dist_sq;  ///< correspondences with a larger squared distance are dropped
};

// ---- general factors: applied to the reduced 6x6 system on the host (factors/general_factor.hpp:11-75) ----
struct NullFactor {
  template <typename Target, typename Source, typename Tree>
  void update_linearized_system(const Target&, const Source&, const Tree&, const Isometry3d&, Matrix6d*, Vector6d*, double*) const {}
  template <typename Target, typename Source>
  void update_error(const Target&, const Source&, const Isometry3d&, double*) const {}
};

/// Soft constraint that freezes selected degrees of freedom (mask entries of 0 are frozen).
struct RestrictDoFFactor {
  RestrictDoFFactor() : lambda(1e9) {
    for (int i = 0; i < 6; i++) mask[i] = 1.0;
  }
  void set_rotation_mask(double rx, double ry, double rz) { mask[0] = rx, mask[1] = ry, mask[2] = rz; }
  void set_translation_mask(double tx, double ty, double tz) { mask[3] = tx, mask[4] = ty, mask[5] = tz; }
  template <typename Target, typename Source, typename Tree>
  void update_linearized_system(const Target&, const Source&, const Tree&, const Isometry3d&, Matrix6d* H, Vector6d*, double*) const {
    for (int i = 0; i < 6; i++) (*H)(i, i) += lambda * std::abs(mask[i] - 1.0);
  }
  template <typename Target, typename Source>
  void update_error(const Target&, const Source&, const Isometry3d&, double*) const {}
  double lambda;
  Vector6d mask;  ///< (rx, ry, rz, tx, ty, tz)
};

}  // namespace small_gicp_b200
