// SPDX-License-Identifier: MIT
// The reference's src/example/01_basic_registration.cpp, line for line in intent, against the B200 backend: the only
// differences are the include paths / namespace of this repository's Eigen-free host mirror (inside the reference tree
// itself one would include <small_gicp/registration/reduction_cuda.hpp> instead) and the optional
// `setting.device`.  Every per-point step -- down-sampling, normals / covariances, nearest neighbours, factors, the
// H | b | e reduction -- runs on the GPU through libsgicp_b200.so; the 6 x 6 solves stay on the host.
//
//   make -C examples && examples/01_basic_registration target.ply source.ply
#include <iostream>

#include <small_gicp_b200/read_points.hpp>
#include <small_gicp_b200/registration_helper.hpp>

using namespace small_gicp_b200;

static void print(const RegistrationResult& result) {
  std::cout << "--- T_target_source ---" << std::endl;
  for (int r = 0; r < 4; r++) {
    for (int c = 0; c < 4; c++) std::cout << result.T_target_source.matrix()(r, c) << (c == 3 ? "\n" : " ");
  }
  std::cout << "converged:" << result.converged << std::endl;
  std::cout << "error:" << result.error << std::endl;
  std::cout << "iterations:" << result.iterations << std::endl;
  std::cout << "num_inliers:" << result.num_inliers << std::endl;
  std::cout << "--- H ---" << std::endl;
  for (int r = 0; r < 6; r++) {
    for (int c = 0; c < 6; c++) std::cout << result.H(r, c) << (c == 5 ? "\n" : " ");
  }
  std::cout << "--- b ---" << std::endl;
  for (int r = 0; r < 6; r++) std::cout << result.b[r] << (r == 5 ? "\n" : " ");
}

/// Most basic registration example (example1 of the reference).
static void example1(const std::vector<Vector4f>& target_points, const std::vector<Vector4f>& source_points) {
  RegistrationSetting setting;
  setting.num_threads = 4;                    // accepted for source compatibility; the device does the work
  setting.downsampling_resolution = 0.25;     // Downsampling resolution
  setting.max_correspondence_distance = 1.0;  // Maximum correspondence distance between points (e.g., triming threshold)

  Isometry3d init_T_target_source = Isometry3d::Identity();
  RegistrationResult result = align(target_points, source_points, init_T_target_source, setting);
  print(result);
}

/// Preprocessing and registration performed separately (example2 of the reference).
static void example2(const std::vector<Vector4f>& target_points, const std::vector<Vector4f>& source_points) {
  int num_threads = 4;
  double downsampling_resolution = 0.25;  // Downsampling resolution
  int num_neighbors = 10;                 // Number of neighbor points used for normal and covariance estimation

  // std::pair<PointCloud::Ptr, KdTree<PointCloud>::Ptr>
  auto [target, target_tree] = preprocess_points(target_points, downsampling_resolution, num_neighbors, num_threads);
  auto [source, source_tree] = preprocess_points(source_points, downsampling_resolution, num_neighbors, num_threads);

  RegistrationSetting setting;
  setting.num_threads = num_threads;
  setting.max_correspondence_distance = 1.0;

  Isometry3d init_T_target_source = Isometry3d::Identity();
  RegistrationResult result = align(*target, *source, *target_tree, init_T_target_source, setting);
  print(result);

  // Preprocessed points and trees can be reused for the next registration for efficiency
  RegistrationResult result2 = align(*source, *target, *source_tree, Isometry3d::Identity(), setting);
  std::cout << "reverse: converged:" << result2.converged << " iterations:" << result2.iterations << std::endl;
}

/// The template surface itself (03_registration_template.cpp of the reference): pick factor, reduction, rejector, optimizer as types.
static void example3(const std::vector<Vector4f>& target_points, const std::vector<Vector4f>& source_points) {
  auto [target, target_tree] = preprocess_points(target_points, 0.25, 10, 4);
  auto [source, source_tree] = preprocess_points(source_points, 0.25, 10, 4);

  Registration<GICPFactor, ParallelReductionCUDA> registration;  // where the reference writes ParallelReductionOMP / ParallelReductionTBB
  registration.reduction.device = 0;
  registration.rejector.max_dist_sq = 1.0;
  registration.optimizer.max_iterations = 20;
  RegistrationResult result = registration.align(*target, *source, *target_tree, Isometry3d::Identity());
  std::cout << "template: converged:" << result.converged << " iterations:" << result.iterations << " num_inliers:" << result.num_inliers << std::endl;
}

int main(int argc, char** argv) {
  const std::string target_path = argc > 1 ? argv[1] : "data/target.ply";
  const std::string source_path = argc > 2 ? argv[2] : "data/source.ply";
  std::vector<Vector4f> target_points = read_ply(target_path);
  std::vector<Vector4f> source_points = read_ply(source_path);
  if (target_points.empty() || source_points.empty()) {
    std::cerr << "error: failed to read points from " << target_path << " / " << source_path << std::endl;
    return 1;
  }
  try {
    example1(target_points, source_points);
    example2(target_points, source_points);
    example3(target_points, source_points);
  } catch (const std::exception& e) {  // e.g. no CUDA device: the backend has no CPU fallback
    std::cerr << "error: " << e.what() << std::endl;
    return 2;
  }
  return 0;
}
