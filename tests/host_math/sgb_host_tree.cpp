// SPDX-License-Identifier: MIT
// TEST INFRASTRUCTURE -- not part of the product, never loaded by small_gicp_b200.
// The host-side tree code of the product (small_gicp_b200/csrc/sgb_kdtree_host.cpp: adoption of a reference-built kd-tree,
// the library's own median-split builder, the 64-byte packet records) compiled stand-alone, plus two plain host walkers
// that follow the device kernels' traversal rules (kd_nearest in sgb_device.cuh, packet_search_kernel in
// sgb_kernels_packet.cu) -- so that tests/test_host_tree_structures.py can check on a CPU-only machine that the structures
// the library uploads are valid exact-nearest-neighbour structures, and that malformed input is refused.
#include <cfloat>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "sgb_kdtree_host.hpp"

namespace {

std::string g_err;

struct Query {
  float x, y, z;
};

// exact NN over the flattened 8-byte nodes: descend to the query's leaf, then unwind the far children whose splitting plane is
// closer than the best distance (kd_nearest, sgb_device.cuh)
void walk_kd(const sgb::FlatTree& t, const std::vector<float>& leaf_pts, const Query& q, uint32_t& best, float& best_d) {
  struct Far {
    uint32_t node;
    float cut;
  };
  std::vector<Far> stack;
  uint32_t node = 0;
  for (;;) {
    sgb::FlatNode nd = t.nodes[node];
    while ((nd.y & 3u) != 3u) {
      const uint32_t axis = nd.y & 3u;
      float th;
      std::memcpy(&th, &nd.x, 4);
      const float qv = axis == 0 ? q.x : (axis == 1 ? q.y : q.z);
      const float diff = qv - th;
      const uint32_t right = nd.y >> 2, left = node + 1;
      const bool go_left = diff < 0.0f;
      if (diff * diff < best_d) stack.push_back({go_left ? right : left, diff * diff});
      node = go_left ? left : right;
      nd = t.nodes[node];
    }
    const uint32_t first = nd.x, cnt = nd.y >> 2;
    for (uint32_t j = 0; j < cnt; j++) {
      const float* p = &leaf_pts[4 * static_cast<size_t>(first + j)];
      const float dx = p[0] - q.x, dy = p[1] - q.y, dz = p[2] - q.z;
      const float d = dx * dx + dy * dy + dz * dz;
      if (d < best_d) {
        best_d = d;
        best = first + j;
      }
    }
    bool found = false;
    while (!stack.empty()) {
      const Far f = stack.back();
      stack.pop_back();
      if (f.cut < best_d) {
        node = f.node;
        found = true;
        break;
      }
    }
    if (!found) return;
  }
}

float box_dist2(const Query& q, const float* lo, const float* hi) {
  const float dx = std::fmax(std::fmax(lo[0] - q.x, q.x - hi[0]), 0.0f);
  const float dy = std::fmax(std::fmax(lo[1] - q.y, q.y - hi[1]), 0.0f);
  const float dz = std::fmax(std::fmax(lo[2] - q.z, q.z - hi[2]), 0.0f);
  return dx * dx + dy * dy + dz * dz;
}

// exact NN over the packet records: a child is entered only if its tight box is closer than the best distance
void walk_packet(const std::vector<sgb::PacketNode>& pn, const std::vector<float>& leaf_pts, const Query& q, uint32_t& best, float& best_d, int max_pending, int* deepest) {
  struct Item {
    uint32_t a, b;
    float d;
  };
  std::vector<Item> stack;
  stack.push_back({0u, 0u, 0.0f});
  while (!stack.empty()) {
    const Item it = stack.back();
    stack.pop_back();
    if (!(it.d < best_d)) continue;
    if (it.b != 0u) {  // leaf: first = a, count = b
      for (uint32_t j = 0; j < it.b; j++) {
        const float* p = &leaf_pts[4 * static_cast<size_t>(it.a + j)];
        const float dx = p[0] - q.x, dy = p[1] - q.y, dz = p[2] - q.z;
        const float d = dx * dx + dy * dy + dz * dz;
        if (d < best_d) {
          best_d = d;
          best = it.a + j;
        }
      }
      continue;
    }
    const float* v = pn[it.a].v;
    uint32_t la, lb, ra, rb;
    std::memcpy(&la, &v[3], 4);
    std::memcpy(&lb, &v[7], 4);
    std::memcpy(&ra, &v[11], 4);
    std::memcpy(&rb, &v[15], 4);
    const float dl = box_dist2(q, v, v + 4), dr = box_dist2(q, v + 8, v + 12);
    // nearer child on top of the stack
    if (dl <= dr) {
      if (dr < best_d) stack.push_back({ra, rb, dr});
      if (dl < best_d) stack.push_back({la, lb, dl});
    } else {
      if (dl < best_d) stack.push_back({la, lb, dl});
      if (dr < best_d) stack.push_back({ra, rb, dr});
    }
    if (static_cast<int>(stack.size()) > *deepest) *deepest = static_cast<int>(stack.size());
  }
  (void)max_pending;
}

// centred FP32 copies: original order (float4 stride) and leaf order
void centre_points(const double* pts4, size_t n, const double* centre, std::vector<float>& orig) {
  orig.resize(4 * n);
  for (size_t i = 0; i < n; i++) {
    for (int a = 0; a < 3; a++) orig[4 * i + a] = static_cast<float>(pts4[4 * i + a] - centre[a]);
    orig[4 * i + 3] = 0.0f;
  }
}

int search_all(const sgb::FlatTree& tree, const std::vector<float>& orig, const double* centre, size_t nq, const double* queries4, int mode, uint64_t* out_idx,
               float* out_d2, int* out_info) {
  const size_t n = tree.perm.size();
  std::vector<float> leaf(4 * n);
  for (size_t i = 0; i < n; i++) std::memcpy(&leaf[4 * i], &orig[4 * static_cast<size_t>(tree.perm[i])], 4 * sizeof(float));
  std::vector<sgb::PacketNode> pn;
  int max_pending = 1, deepest = 0;
  if (mode == 1 && !sgb::build_packet_nodes(tree, orig.data(), pn, &max_pending)) {
    g_err = "build_packet_nodes failed";
    return 1;
  }
  for (size_t k = 0; k < nq; k++) {
    const Query q{static_cast<float>(queries4[4 * k] - centre[0]), static_cast<float>(queries4[4 * k + 1] - centre[1]), static_cast<float>(queries4[4 * k + 2] - centre[2])};
    uint32_t best = 0xFFFFFFFFu;
    float best_d = FLT_MAX;
    if (n) {
      if (mode == 0)
        walk_kd(tree, leaf, q, best, best_d);
      else
        walk_packet(pn, leaf, q, best, best_d, max_pending, &deepest);
    }
    out_idx[k] = best == 0xFFFFFFFFu ? ~0ull : static_cast<uint64_t>(tree.perm[best]);
    out_d2[k] = best_d;
  }
  if (out_info) {
    out_info[0] = static_cast<int>(tree.nodes.size());
    out_info[1] = tree.depth;
    out_info[2] = static_cast<int>(pn.size());
    out_info[3] = deepest;
  }
  return 0;
}

}  // namespace

extern "C" {

const char* sgbt_last_error() { return g_err.c_str(); }

/// Adopt a reference-layout kd-tree (sgb_target_set_kdtree's host half) and answer `nq` nearest-neighbour queries with it.
/// mode 0: 8-byte kd nodes, 1: packet records.  out_idx = ORIGINAL point indices, out_d2 = FP32 squared distances (centred coordinates).
int sgbt_adopt_and_search(const void* nodes24, size_t n_nodes, uint32_t root, const uint64_t* indices, size_t n_points, const double* pts4, const double* centre3,
                          size_t nq, const double* queries4, int mode, uint64_t* out_idx, float* out_d2, int* out_info4) {
  sgb::FlatTree tree;
  g_err.clear();
  if (!sgb::flatten_reference_tree(nodes24, n_nodes, root, indices, n_points, centre3, tree, g_err)) return 1;
  std::vector<float> orig;
  centre_points(pts4, n_points, centre3, orig);
  return search_all(tree, orig, centre3, nq, queries4, mode, out_idx, out_d2, out_info4);
}

/// The library's own host builder (SGB_TREE=host) over the same points.
int sgbt_build_and_search(size_t n_points, const double* pts4, const double* centre3, int max_leaf, size_t nq, const double* queries4, int mode, uint64_t* out_idx,
                          float* out_d2, int* out_info4) {
  std::vector<float> orig;
  centre_points(pts4, n_points, centre3, orig);
  sgb::FlatTree tree;
  g_err.clear();
  if (!sgb::build_flat_tree(orig.data(), n_points, max_leaf, tree, g_err)) return 1;
  // every point exactly once, leaves within the bound
  std::vector<unsigned char> seen(n_points, 0);
  for (uint32_t p : tree.perm) {
    if (p >= n_points || seen[p]) {
      g_err = "permutation is not a bijection";
      return 2;
    }
    seen[p] = 1;
  }
  const int bound = max_leaf <= 0 ? 32 : (max_leaf > 64 ? 64 : max_leaf);
  for (const sgb::FlatNode& nd : tree.nodes)
    if ((nd.y & 3u) == 3u && static_cast<int>(nd.y >> 2) > bound) {
      g_err = "leaf larger than max_leaf_size";
      return 2;
    }
  return search_all(tree, orig, centre3, nq, queries4, mode, out_idx, out_d2, out_info4);
}

}  // extern "C"
