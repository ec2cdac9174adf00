// SPDX-License-Identifier: MIT
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

namespace sgb {

struct FlatNode {
  uint32_t x, y;  // see KdNode in sgb_device.cuh
};

struct FlatTree {
  std::vector<FlatNode> nodes;  // pre-order: left child of i is i+1
  std::vector<uint32_t> perm;   // leaf order -> original point index
  int depth = 0;                // number of inner nodes on the deepest root-to-leaf path (= max stack entries)
};

/// 64-byte inner-node record of the packet search: {L.lo|a, L.hi|b, R.lo|a, R.hi|b} (sgb_kernels_packet.cu)
struct PacketNode {
  float v[16];
};

/// Adopt a reference-built tree (24-byte nodes, size_t indices). Thresholds are re-expressed relative to `centre` in FP32.
bool flatten_reference_tree(const void* nodes24, size_t n_nodes, uint32_t root, const uint64_t* indices, size_t n_points, const double centre[3],
                            FlatTree& out, std::string& err);

/// Own builder over centred FP32 points (float4 stride): widest-extent axis, median split, leaves <= max_leaf_size.
bool build_flat_tree(const float* pts_xyzw, size_t n_points, int max_leaf_size, FlatTree& out, std::string& err);

/// Children bounding boxes for every inner node of `tree` over the points `pts_xyzw` (original order, float4 stride).
bool build_packet_nodes(const FlatTree& tree, const float* pts_xyzw, std::vector<PacketNode>& out, int* max_pending);

}  // namespace sgb
