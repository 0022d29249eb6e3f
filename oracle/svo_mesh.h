// ORACLE — test infrastructure only (see svo_math.h header; parity unpinned end-to-end).
//
// CPU restatement of the collider geometry of the reference (paths relative to
// /root/reference/rust/crates):
//   util/src/aabb.rs:117-167                              Aabb (inclusive overlap / contains)
//   mesh_util/src/bounding_volume_hierarchy.rs:85-217     64-ary integer-lattice BVH build + query
//   mesh_util/src/mesh.rs:30-142                          Topology (opposites, closed-fan vertex lists)
//   mesh_util/src/mesh.rs:228-309                         point-segment / point-triangle distance
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "svo_math.h"

namespace svo {

struct Triangle { uint32_t a, b, c; uint32_t operator[](int i) const { return (&a)[i]; } };
struct Opposites { uint32_t ab, bc, ca; };

struct AabbI {
  Vec3i min{INT32_MAX, INT32_MAX, INT32_MAX};  // splat(f32::MAX as i32) saturates
  Vec3i max{INT32_MIN, INT32_MIN, INT32_MIN};
  static bool leq(const Vec3i& a, const Vec3i& b) { return a.x <= b.x && a.y <= b.y && a.z <= b.z; }
  bool has_overlap(const AabbI& o) const { return leq(min, o.max) && leq(o.min, max); }
  bool contains(const Vec3i& p) const { return leq(min, p) && leq(p, max); }
  void extend(const Vec3i& p) {
    for (int i = 0; i < 3; ++i) { min[i] = std::min(min[i], p[i]); max[i] = std::max(max[i], p[i]); }
  }
};

struct BvhNode {
  bool leaf = false;
  AabbI aabb;
  std::array<int32_t, 64> children;  // -1 = None
  std::vector<uint32_t> indices;     // leaf only
};

struct Bvh {
  uint32_t level = 0;
  std::vector<BvhNode> nodes;

  static AabbI aabb_from_offset_and_level(const Vec3i& off, uint32_t level) {
    AabbI a;
    int32_t side = 1;
    for (uint32_t i = 0; i < level; ++i) side *= 4;
    a.min = off;
    a.max = {off.x + side, off.y + side, off.z + side};
    return a;
  }

  // bounding_volume_hierarchy.rs:85-175
  void build(const std::vector<AabbI>& aabbs, uint32_t leaf_threshold) {
    level = 0;
    nodes.clear();
    AabbI aabb;
    for (auto& b : aabbs) { aabb.extend(b.min); aabb.extend(b.max); }
    if (aabbs.empty()) return;
    const Vec3i ext{aabb.max.x - aabb.min.x, aabb.max.y - aabb.min.y, aabb.max.z - aabb.min.z};
    const int32_t longest = std::max(ext.x, std::max(ext.y, ext.z));
    if (longest <= 0) return;  // "Bounding Volumes Hierarchy empty"
    uint32_t ilog4 = 0;
    for (int64_t v = longest; v >= 4; v /= 4) ++ilog4;
    level = ilog4 + 1;
    int32_t side = 1;
    for (uint32_t i = 0; i < level; ++i) side *= 4;
    // Rust integer division truncates toward zero, as does C++.
    const Vec3i offset{(aabb.min.x + aabb.max.x - side) / 2, (aabb.min.y + aabb.max.y - side) / 2, (aabb.min.z + aabb.max.z - side) / 2};
    std::vector<uint32_t> all(aabbs.size());
    for (uint32_t i = 0; i < aabbs.size(); ++i) all[i] = i;
    create(leaf_threshold, aabbs, level, aabb_from_offset_and_level(offset, level), std::move(all));
  }

  uint32_t create(uint32_t leaf_threshold, const std::vector<AabbI>& aabbs, uint32_t lvl, const AabbI& aabb, std::vector<uint32_t> indices) {
    const uint32_t index = (uint32_t)nodes.size();
    nodes.emplace_back();
    nodes[index].aabb = aabb;
    nodes[index].children.fill(-1);
    if (lvl == 0 || indices.size() < leaf_threshold) {
      nodes[index].leaf = true;
      nodes[index].indices = std::move(indices);
      return index;
    }
    const uint32_t child_level = lvl - 1;
    for (int child = 0; child < 64; ++child) {
      const Vec3i child_offset{aabb.min.x + (((child >> 4) & 3) << (2 * child_level)), aabb.min.y + (((child >> 2) & 3) << (2 * child_level)),
                               aabb.min.z + (((child >> 0) & 3) << (2 * child_level))};
      const AabbI child_aabb = aabb_from_offset_and_level(child_offset, child_level);
      std::vector<uint32_t> child_indices;
      for (uint32_t i : indices)
        if (aabbs[i].has_overlap(child_aabb)) child_indices.push_back(i);
      if (!child_indices.empty()) {
        const uint32_t ci = create(leaf_threshold, aabbs, child_level, child_aabb, std::move(child_indices));
        nodes[index].children[child] = (int32_t)ci;
      }
    }
    return index;
  }

  // bounding_volume_hierarchy.rs:177-217; returns nullptr for "empty"
  const std::vector<uint32_t>* query(const Vec3i& p) const {
    if (nodes.empty()) return nullptr;
    const BvhNode* cur = &nodes[0];
    if (!cur->aabb.contains(p)) return nullptr;
    if (cur->leaf) return &cur->indices;
    const uint32_t q[3] = {(uint32_t)(p.x - cur->aabb.min.x), (uint32_t)(p.y - cur->aabb.min.y), (uint32_t)(p.z - cur->aabb.min.z)};
    for (int lvl = (int)level - 1; lvl >= 0; --lvl) {
      // NOTE: a query on the root's inclusive max face gives (q >> 2*lvl) = 4 at the top level; the
      // reference masks with & 3 just the same (it wraps to child 0 of that axis).
      const uint32_t child = (((q[0] >> (2 * lvl)) & 3) << 4) | (((q[1] >> (2 * lvl)) & 3) << 2) | (((q[2] >> (2 * lvl)) & 3) << 0);
      const int32_t ci = cur->children[child];
      if (ci < 0) return nullptr;
      const BvhNode* next = &nodes[ci];
      if (next->leaf) return &next->indices;
      cur = next;
    }
    return nullptr;  // unreachable!() in the reference
  }
};

// bounding_volume_hierarchy.rs:219-237 and xpu/src/frame_input.rs:350-390 (per-triangle lattice AABB)
inline AabbI triangle_leaf_aabb(const Vec3f* pts, int n, float margin, float leaf_size) {
  Vec3f mn{std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
  Vec3f mx{std::numeric_limits<float>::lowest(), std::numeric_limits<float>::lowest(), std::numeric_limits<float>::lowest()};
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) { mn[k] = std::min(mn[k], pts[i][k]); mx[k] = std::max(mx[k], pts[i][k]); }
  AabbI r;
  for (int k = 0; k < 3; ++k) {
    r.min[k] = (int32_t)std::floor((mn[k] - margin) / leaf_size);
    r.max[k] = (int32_t)std::ceil((mx[k] + margin) / leaf_size);
  }
  return r;
}

struct TopologyInput {
  uint32_t collider;
  uint32_t num_vertices;
  std::vector<Triangle> triangles;
};

struct Topology {
  std::vector<std::vector<uint32_t>> vertex_triangle_lists;
  std::vector<Triangle> triangle_indices;
  std::vector<Opposites> triangle_opposites;
  std::vector<uint32_t> triangle_collider;

  // mesh_util/src/mesh.rs:30-142.  Returns "" on success, else the error text.
  std::string build(const std::vector<TopologyInput>& inputs) {
    vertex_triangle_lists.clear(); triangle_indices.clear(); triangle_opposites.clear(); triangle_collider.clear();
    uint32_t vertex_index_offset = 0;
    for (auto& input : inputs) {
      for (size_t t = 0; t < input.triangles.size(); ++t)
        for (int k = 0; k < 3; ++k)
          if (input.triangles[t][k] >= input.num_vertices) return "VertexIndexOutOfRange";
      std::map<std::pair<uint32_t, uint32_t>, std::vector<uint32_t>> edge_to_triangle;
      auto order_edge = [](uint32_t a, uint32_t b) { return a < b ? std::make_pair(a, b) : std::make_pair(b, a); };
      for (uint32_t t = 0; t < input.triangles.size(); ++t)
        for (int k = 0; k < 3; ++k) edge_to_triangle[order_edge(input.triangles[t][k], input.triangles[t][(k + 1) % 3])].push_back(t);
      for (auto& kv : edge_to_triangle)
        if (kv.second.size() > 2) return "NonManifoldEdge";
      const uint32_t triangle_index_offset = (uint32_t)triangle_indices.size();
      for (uint32_t t = 0; t < input.triangles.size(); ++t) {
        uint32_t opp[3];
        for (int k = 0; k < 3; ++k) {
          opp[k] = UINT32_MAX;
          for (uint32_t other : edge_to_triangle[order_edge(input.triangles[t][k], input.triangles[t][(k + 1) % 3])])
            if (other != t) { opp[k] = other + triangle_index_offset; break; }
        }
        triangle_opposites.push_back({opp[0], opp[1], opp[2]});
      }
      for (auto& t : input.triangles) triangle_indices.push_back({t.a + vertex_index_offset, t.b + vertex_index_offset, t.c + vertex_index_offset});
      triangle_collider.resize(triangle_collider.size() + input.triangles.size(), input.collider);
      vertex_index_offset += input.num_vertices;
    }
    vertex_triangle_lists.assign(vertex_index_offset, {});
    for (uint32_t t = 0; t < triangle_indices.size(); ++t)
      for (int k = 0; k < 3; ++k) vertex_triangle_lists[triangle_indices[t][k]].push_back(t);
    for (uint32_t v = 0; v < vertex_triangle_lists.size(); ++v) {
      std::map<uint32_t, int> neighbor_counts;
      for (uint32_t t : vertex_triangle_lists[v])
        for (int k = 0; k < 3; ++k)
          if (triangle_indices[t][k] != v) neighbor_counts[triangle_indices[t][k]]++;
      bool open = false;
      for (auto& kv : neighbor_counts) {
        if (kv.second > 2) return "missed non-manifoldness before";  // assert! in the reference
        if (kv.second != 2) open = true;
      }
      if (open) vertex_triangle_lists[v].clear();
    }
    return "";
  }
};

struct DistanceResult { float distance; Vec3f to_p; Vec3f normal; };

// mesh_util/src/mesh.rs:234-263
inline DistanceResult segment_distance_result(const Vec3f& p, const Vec3f& start, const Vec3f& end, const Vec3f& start_normal, const Vec3f& segment_normal,
                                              const Vec3f& end_normal) {
  const Vec3f segment = end - start;
  const float along = (p - start).dot(segment) / segment.dot(segment);
  if (along < 0.f) return {(p - start).norm(), p - start, start_normal};
  if (along < 1.f) return {(p - start - segment * along).norm(), p - start - segment * along, segment_normal};
  return {(p - end).norm(), p - end, end_normal};
}
// mesh_util/src/mesh.rs:265-275
inline float distance_to_segment(const Vec3f& p, const Vec3f& start, const Vec3f& end) {
  const Vec3f segment = end - start;
  const float along = (p - start).dot(segment) / segment.dot(segment);
  if (along < 0.f) return (p - start).norm();
  if (along < 1.f) return (p - start - segment * along).norm();
  return (p - end).norm();
}
// mesh_util/src/mesh.rs:277-309
inline float distance_to_triangle(const Vec3f& p, const Vec3f& a, const Vec3f& b, const Vec3f& c, const Vec3f& n) {
  const Vec3f ab = a - b, bc = b - c, ca = c - a;
  const bool sa = n.dot(bc.cross(c - p)) > 0.f;
  const bool sb = n.dot(ca.cross(a - p)) > 0.f;
  const bool sc = n.dot(ab.cross(b - p)) > 0.f;
  if (sa && sb && sc) return std::fabs((p - a).dot(n));
  float d = std::numeric_limits<float>::max();
  if (!sa) d = std::min(d, distance_to_segment(p, b, c));
  if (!sb) d = std::min(d, distance_to_segment(p, c, a));
  if (!sc) d = std::min(d, distance_to_segment(p, a, b));
  return d;
}

}  // namespace svo
