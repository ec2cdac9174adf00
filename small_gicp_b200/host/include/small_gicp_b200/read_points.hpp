// SPDX-License-Identifier: MIT
// On-disk readers of the reference's benchmark / example inputs, for the Eigen-free host mirror
// (/root/reference/include/small_gicp/benchmark/read_points.hpp):
//   read_points  : KITTI velodyne .bin = N x (x, y, z, intensity) float32; w is overwritten with 1   (:15-33)
//   write_points : the inverse                                                                       (:38-46)
//   read_ply     : binary PLY whose vertex properties are all `float`, the first three named x y z   (:52-109)
// Same behaviour on bad input as the reference: a message on std::cerr and an empty result, never an exception.
// One deliberate difference: read_ply reads `#properties x #vertices` floats; the reference reads `4 x #vertices`
// (read_points.hpp:100-101), which is the same thing for its bundled files (x, y, z, intensity) and a short read otherwise.
#pragma once
#include <array>
#include <cctype>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "core.hpp"

namespace small_gicp_b200 {

/// (x, y, z, 1) in single precision: what LiDAR drivers, KITTI .bin files and the bundled PLY files deliver
using Vector4f = std::array<float, 4>;
static_assert(sizeof(Vector4f) == 16, "Vector4f must be four packed floats");

inline std::vector<Vector4f> read_points(const std::string& filename) {
  std::ifstream ifs(filename, std::ios::binary | std::ios::ate);
  if (!ifs) {
    std::cerr << "error: failed to open " << filename << std::endl;
    return {};
  }
  const std::streamsize bytes = ifs.tellg();
  const size_t n = bytes > 0 ? static_cast<size_t>(bytes) / sizeof(Vector4f) : 0;
  ifs.seekg(0, std::ios::beg);
  std::vector<Vector4f> points(n);
  ifs.read(reinterpret_cast<char*>(points.data()), static_cast<std::streamsize>(sizeof(Vector4f) * n));
  for (auto& p : points) p[3] = 1.0f;
  return points;
}

inline void write_points(const std::string& filename, const std::vector<Vector4f>& points) {
  std::ofstream ofs(filename, std::ios::binary);
  if (!ofs) {
    std::cerr << "error: failed to open " << filename << std::endl;
    return;
  }
  ofs.write(reinterpret_cast<const char*>(points.data()), static_cast<std::streamsize>(sizeof(Vector4f) * points.size()));
}

inline std::vector<Vector4f> read_ply(const std::string& filename) {
  std::ifstream ifs(filename, std::ios::binary);
  if (!ifs) {
    std::cerr << "error: failed to open " << filename << std::endl;
    return {};
  }
  std::vector<std::string> properties;
  size_t n = 0;
  bool header_done = false;
  std::string line;
  while (std::getline(ifs, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line == "end_header") {
      header_done = true;
      break;
    }
    std::stringstream sst(line);
    std::string token;
    sst >> token;
    if (token == "element") {
      std::string what, count;
      sst >> what >> count;
      if (what != "vertex") {
        std::cerr << "error: invalid ply format (line=" << line << ")" << std::endl;
        return {};
      }
      try {
        n = static_cast<size_t>(std::stoull(count));
      } catch (const std::exception&) {
        std::cerr << "error: invalid ply format (line=" << line << ")" << std::endl;
        return {};
      }
    } else if (token == "property") {
      std::string type, name;
      sst >> type >> name;
      if (type != "float") {
        std::cerr << "error: only float properties are supported!! (line=" << line << ")" << std::endl;
        return {};
      }
      properties.push_back(name);
    }
  }
  auto is_axis = [&](size_t k, char c) { return properties[k].size() == 1 && std::tolower(static_cast<unsigned char>(properties[k][0])) == c; };
  if (!header_done || properties.size() < 3 || !is_axis(0, 'x') || !is_axis(1, 'y') || !is_axis(2, 'z')) {
    std::cerr << "error: invalid ply header or properties (the first three must be float x, y, z)" << std::endl;
    return {};
  }
  const size_t stride = properties.size();
  std::vector<float> buffer(stride * n);
  ifs.read(reinterpret_cast<char*>(buffer.data()), static_cast<std::streamsize>(sizeof(float) * buffer.size()));
  if (static_cast<size_t>(ifs.gcount()) != sizeof(float) * buffer.size()) {
    std::cerr << "error: truncated vertex data in " << filename << std::endl;
    return {};
  }
  std::vector<Vector4f> points(n);
  for (size_t i = 0; i < n; i++) points[i] = Vector4f{buffer[i * stride + 0], buffer[i * stride + 1], buffer[i * stride + 2], 1.0f};
  return points;
}

/// PointCloud from raw single-precision points (the reference's `std::make_shared<PointCloud>(points)`, point_cloud.hpp:24-35)
inline PointCloud::Ptr make_point_cloud(const std::vector<Vector4f>& points) {
  auto cloud = std::make_shared<PointCloud>();
  cloud->resize(points.size());
  for (size_t i = 0; i < points.size(); i++) cloud->points[i] = vec4(points[i][0], points[i][1], points[i][2], 1.0);
  return cloud;
}

}  // namespace small_gicp_b200
