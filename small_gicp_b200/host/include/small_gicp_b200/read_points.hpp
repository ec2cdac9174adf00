// SPDX-License-Identifier: MIT
// File readers for the inputs the reference's examples and benchmarks use, for the Eigen-free host mirror.  Behavioural model:
// /root/reference/include/small_gicp/benchmark/read_points.hpp
//   read_points  (:15-33)   KITTI velodyne .bin -- packed float32 records (x, y, z, intensity); the fourth value is replaced by 1
//   write_points (:38-46)   the same records back to disk
//   read_ply     (:52-109)  binary PLY whose vertex properties are all `float`, x y z first
// Like the reference, bad input never throws: one line on stderr, empty result.  One deliberate difference: read_ply consumes
// (#properties x #vertices) floats, where the reference always reads 4 x #vertices (:100-101) -- identical for its bundled
// x/y/z/intensity files, a short read for any other property count.
#pragma once
#include <array>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "core.hpp"

namespace small_gicp_b200 {

/// (x, y, z, 1) in single precision: what LiDAR drivers, KITTI .bin files and the bundled PLY files deliver
using Vector4f = std::array<float, 4>;
static_assert(sizeof(Vector4f) == 16, "Vector4f must be four packed floats");

namespace detail {

struct FileCloser {
  void operator()(std::FILE* f) const {
    if (f) std::fclose(f);
  }
};
using File = std::unique_ptr<std::FILE, FileCloser>;

inline File open_file(const std::string& path, const char* mode) {
  File f(std::fopen(path.c_str(), mode));
  if (!f) std::fprintf(stderr, "error: failed to open %s\n", path.c_str());
  return f;
}

/// What a PLY header tells us: vertex count, number of float properties per vertex, and whether it is usable here.
struct PlyHeader {
  size_t vertices = 0;
  size_t floats_per_vertex = 0;
  bool ok = false;
};

inline PlyHeader parse_ply_header(std::FILE* f, const std::string& path) {
  PlyHeader h;
  std::vector<std::string> names;
  char buf[512];
  bool finished = false;
  while (std::fgets(buf, sizeof(buf), f)) {
    std::string line(buf);
    while (!line.empty() && (line.back() == '\n' || line.back() == '\r')) line.pop_back();
    if (line == "end_header") {
      finished = true;
      break;
    }
    char key[32] = {0}, a[64] = {0}, b[64] = {0};
    const int got = std::sscanf(line.c_str(), "%31s %63s %63s", key, a, b);
    if (got >= 3 && std::strcmp(key, "element") == 0) {
      if (std::strcmp(a, "vertex") != 0) {  // faces etc.: not a plain point cloud
        std::fprintf(stderr, "error: %s: unsupported ply element (%s)\n", path.c_str(), line.c_str());
        return h;
      }
      char* end = nullptr;
      h.vertices = static_cast<size_t>(std::strtoull(b, &end, 10));
      if (end == b) {
        std::fprintf(stderr, "error: %s: bad vertex count (%s)\n", path.c_str(), line.c_str());
        return h;
      }
    } else if (got >= 3 && std::strcmp(key, "property") == 0) {
      if (std::strcmp(a, "float") != 0 && std::strcmp(a, "float32") != 0) {
        std::fprintf(stderr, "error: %s: only float properties are supported (%s)\n", path.c_str(), line.c_str());
        return h;
      }
      names.emplace_back(b);
    }
  }
  auto axis = [&](size_t k, char lower) { return names[k].size() == 1 && (names[k][0] | 0x20) == lower; };
  if (!finished || names.size() < 3 || !axis(0, 'x') || !axis(1, 'y') || !axis(2, 'z')) {
    std::fprintf(stderr, "error: %s: header incomplete or the first three properties are not x y z\n", path.c_str());
    return h;
  }
  h.floats_per_vertex = names.size();
  h.ok = true;
  return h;
}

}  // namespace detail

inline std::vector<Vector4f> read_ply(const std::string& filename) {
  detail::File f = detail::open_file(filename, "rb");
  if (!f) return {};
  const detail::PlyHeader h = detail::parse_ply_header(f.get(), filename);
  if (!h.ok) return {};
  std::vector<float> raw(h.vertices * h.floats_per_vertex);
  if (std::fread(raw.data(), sizeof(float), raw.size(), f.get()) != raw.size()) {
    std::fprintf(stderr, "error: %s: vertex data shorter than the header promises\n", filename.c_str());
    return {};
  }
  std::vector<Vector4f> cloud(h.vertices);
  const float* src = raw.data();
  for (Vector4f& p : cloud) {
    p = {src[0], src[1], src[2], 1.0f};
    src += h.floats_per_vertex;
  }
  return cloud;
}

inline std::vector<Vector4f> read_points(const std::string& filename) {
  detail::File f = detail::open_file(filename, "rb");
  if (!f) return {};
  std::fseek(f.get(), 0, SEEK_END);
  const long bytes = std::ftell(f.get());
  std::fseek(f.get(), 0, SEEK_SET);
  std::vector<Vector4f> cloud(bytes > 0 ? static_cast<size_t>(bytes) / sizeof(Vector4f) : 0);  // a trailing partial record is ignored
  if (!cloud.empty() && std::fread(cloud.data(), sizeof(Vector4f), cloud.size(), f.get()) != cloud.size()) {
    std::fprintf(stderr, "error: %s: short read\n", filename.c_str());
    return {};
  }
  for (Vector4f& p : cloud) p[3] = 1.0f;  // the intensity slot becomes the homogeneous coordinate
  return cloud;
}

inline void write_points(const std::string& filename, const std::vector<Vector4f>& cloud) {
  detail::File f = detail::open_file(filename, "wb");
  if (!f) return;
  if (std::fwrite(cloud.data(), sizeof(Vector4f), cloud.size(), f.get()) != cloud.size()) std::fprintf(stderr, "error: %s: short write\n", filename.c_str());
}

/// PointCloud from raw single-precision points (what `std::make_shared<PointCloud>(points)` does in the reference, point_cloud.hpp:24-35)
inline PointCloud::Ptr make_point_cloud(const std::vector<Vector4f>& raw) {
  auto cloud = std::make_shared<PointCloud>();
  cloud->resize(raw.size());
  for (size_t i = 0; i < raw.size(); i++) cloud->points[i] = vec4(raw[i][0], raw[i][1], raw[i][2], 1.0);
  return cloud;
}

}  // namespace small_gicp_b200
