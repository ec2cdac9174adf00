// SPDX-License-Identifier: MIT
// Gaussian voxel map target for VGICP with the reference's surface
// (GaussianVoxelMap = IncrementalVoxelMap<GaussianVoxel>): insert(), size(), nearest_neighbor_search(),
// set_search_offsets(), calc_index / voxel_id / point_id, point/cov traits
//   /root/reference/include/small_gicp/ann/incremental_voxelmap.hpp:30-237, gaussian_voxelmap.hpp:15-89.
// Voxels are stored by value in one flat vector (no shared_ptr per voxel) so the arrays handed to
// sgb_target_set_voxelmap() are produced by a single pass.  Single-insert use (one target cloud per map)
// is what the registration path needs; the LRU eviction of the incremental map is not part of it.
#pragma once
#include <unordered_map>

#include "core.hpp"

namespace small_gicp_b200 {

struct VoxelCoord {
  int32_t x, y, z;
  bool operator==(const VoxelCoord& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct VoxelCoordHash {
  size_t operator()(const VoxelCoord& c) const {
    uint64_t h = static_cast<uint32_t>(c.x) * 0x9E3779B97F4A7C15ull;
    h ^= (static_cast<uint32_t>(c.y) + 0x7F4A7C15ull + (h << 6) + (h >> 2));
    h ^= (static_cast<uint32_t>(c.z) + 0x9E3779B9ull + (h << 6) + (h >> 2));
    return static_cast<size_t>(h);
  }
};

/// Running mean of the points and of their covariances inside one voxel.
struct GaussianVoxel {
  VoxelCoord coord{0, 0, 0};
  size_t num_points = 0;
  bool finalized = false;
  Vector4d mean;
  Matrix4d cov;
  void add(const Vector4d& pt, const Matrix4d& c) {
    if (finalized) {  // re-open: back to sums
      finalized = false;
      mean *= static_cast<double>(num_points);
      cov *= static_cast<double>(num_points);
    }
    num_points++;
    mean += pt;
    cov += c;
  }
  void finalize() {
    if (finalized) return;
    mean *= 1.0 / static_cast<double>(num_points);
    cov *= 1.0 / static_cast<double>(num_points);
    finalized = true;
  }
};

struct GaussianVoxelMap {
  using Ptr = std::shared_ptr<GaussianVoxelMap>;
  using ConstPtr = std::shared_ptr<const GaussianVoxelMap>;
  static constexpr int point_id_bits = 32;

  explicit GaussianVoxelMap(double leaf_size) : inv_leaf_size(1.0 / leaf_size), leaf_size(leaf_size) { set_search_offsets(1); }

  size_t size() const { return flat_voxels.size(); }

  static int32_t floor_coord(double v) {
    const int32_t t = static_cast<int32_t>(v);
    return t - (v < static_cast<double>(t) ? 1 : 0);
  }
  VoxelCoord coord_of(const Vector4d& pt) const { return {floor_coord(pt[0] * inv_leaf_size), floor_coord(pt[1] * inv_leaf_size), floor_coord(pt[2] * inv_leaf_size)}; }

  /// Add every point of `points` (moved by T) to its voxel, then normalise all voxels.
  template <typename PointCloud>
  void insert(const PointCloud& points, const Isometry3d& T = Isometry3d::Identity()) {
    const Matrix4d Tm = T.matrix(), Tt = T.matrix().transpose();
    for (size_t i = 0; i < traits::size(points); i++) {
      const Vector4d pt = T * traits::point(points, i);
      const VoxelCoord c = coord_of(pt);
      auto it = voxels.find(c);
      if (it == voxels.end()) {
        it = voxels.emplace(c, flat_voxels.size()).first;
        flat_voxels.emplace_back();
        flat_voxels.back().coord = c;
      }
      flat_voxels[it->second].add(pt, Tm * traits::cov(points, i) * Tt);
    }
    for (auto& v : flat_voxels) v.finalize();
    generation++;
  }

  /// 1 (centre), 7 (+ face neighbours) or 27 (full 3x3x3) voxels are probed per query.
  void set_search_offsets(int num_offsets) {
    search_offsets.clear();
    if (num_offsets == 7) {
      search_offsets = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {-1, 0, 0}, {0, -1, 0}, {0, 0, -1}};
    } else if (num_offsets == 27) {
      for (int i = -1; i <= 1; i++)
        for (int j = -1; j <= 1; j++)
          for (int k = -1; k <= 1; k++) search_offsets.push_back({i, j, k});
    } else {
      search_offsets = {{0, 0, 0}};
    }
  }

  size_t calc_index(size_t voxel_id, size_t point_id) const { return (voxel_id << point_id_bits) | point_id; }
  size_t voxel_id(size_t i) const { return i >> point_id_bits; }
  size_t point_id(size_t i) const { return i & ((1ull << point_id_bits) - 1); }

  /// Host-side lookup with the reference's semantics (first strictly-closest voxel mean in offset order).
  size_t nearest_neighbor_search(const Vector4d& pt, size_t* index, double* sq_dist) const {
    const VoxelCoord c = coord_of(pt);
    double best = std::numeric_limits<double>::max();
    size_t found = 0;
    for (const auto& o : search_offsets) {
      const auto it = voxels.find(VoxelCoord{c.x + o.x, c.y + o.y, c.z + o.z});
      if (it == voxels.end()) continue;
      const double d = (flat_voxels[it->second].mean - pt).squaredNorm();
      if (d < best) {
        best = d;
        *index = calc_index(it->second, 0);
        *sq_dist = d;
        found = 1;
      }
    }
    return found;
  }

  double inv_leaf_size, leaf_size;
  std::vector<VoxelCoord> search_offsets;
  std::vector<GaussianVoxel> flat_voxels;
  std::unordered_map<VoxelCoord, size_t, VoxelCoordHash> voxels;
  uint64_t generation = 0;  ///< bumped by insert(); lets the CUDA glue notice in-place changes
};

namespace traits {
template <>
struct Traits<GaussianVoxelMap> {
  static size_t size(const GaussianVoxelMap& m) { return m.size(); }
  static bool has_points(const GaussianVoxelMap&) { return true; }
  static bool has_normals(const GaussianVoxelMap&) { return false; }
  static bool has_covs(const GaussianVoxelMap&) { return true; }
  static Vector4d point(const GaussianVoxelMap& m, size_t i) { return m.flat_voxels[m.voxel_id(i)].mean; }
  static Matrix4d cov(const GaussianVoxelMap& m, size_t i) { return m.flat_voxels[m.voxel_id(i)].cov; }
  static size_t nearest_neighbor_search(const GaussianVoxelMap& m, const Vector4d& p, size_t* k, double* d) { return m.nearest_neighbor_search(p, k, d); }
};
}  // namespace traits

}  // namespace small_gicp_b200
