// svb_host.h — host-side (C++17) pieces of the back end that are not per-particle work:
//   * collider topology (edge opposites, closed-fan vertex lists)    mesh_util/src/mesh.rs:30-142
//   * integer-lattice 64-ary BVH build, flattened for the device     mesh_util/src/bounding_volume_hierarchy.rs:85-175
//                                                                     xpu/src/frame_input.rs:350-390
//   * adaptive time-step bookkeeping                                   cpu/src/adaptive_time_step_state.rs:13-64
// (reference paths relative to /root/reference/rust/crates).  These run once per keyframe / once per
// substep on a handful of scalars; the per-particle and per-node work is in svb_kernels.cu.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <limits>
#include <string>
#include <unordered_map>
#include <vector>

namespace svbh {

struct HostTopology {
  uint32_t n_vertices = 0, n_triangles = 0, n_colliders = 0;
  std::vector<uint32_t> tri;           // 3 per triangle, global vertex indices
  std::vector<uint32_t> opp;           // 3 per triangle (ab, bc, ca), UINT32_MAX = boundary edge
  std::vector<uint32_t> tri_collider;  // per triangle
  std::vector<uint32_t> fan_offsets;   // n_vertices + 1 (CSR); open fans are empty
  std::vector<uint32_t> fan_tris;

  // Returns "" or the TopologyError text.
  std::string build(uint32_t n_col, const uint32_t* num_vertices, const uint32_t* num_triangles, const uint32_t* triangles) {
    *this = HostTopology();
    n_colliders = n_col;
    size_t tri_base = 0;
    uint32_t vertex_base = 0;
    for (uint32_t c = 0; c < n_col; ++c) {
      const uint32_t nv = num_vertices[c], nt = num_triangles[c];
      const uint32_t* t = triangles + 3 * tri_base;
      for (uint32_t i = 0; i < 3 * nt; ++i)
        if (t[i] >= nv) return "VertexIndexOutOfRange";
      // undirected edge -> up to two incident triangles (local triangle ids)
      struct Inc { uint32_t t0 = UINT32_MAX, t1 = UINT32_MAX; uint32_t n = 0; };
      std::unordered_map<uint64_t, Inc> edges;
      edges.reserve((size_t)nt * 2);
      auto ekey = [](uint32_t a, uint32_t b) { return a < b ? ((uint64_t)a << 32) | b : ((uint64_t)b << 32) | a; };
      for (uint32_t ti = 0; ti < nt; ++ti)
        for (int k = 0; k < 3; ++k) {
          Inc& e = edges[ekey(t[3 * ti + k], t[3 * ti + (k + 1) % 3])];
          if (e.n == 0) e.t0 = ti;
          else if (e.n == 1) e.t1 = ti;
          ++e.n;
        }
      for (auto& kv : edges)
        if (kv.second.n > 2) return "NonManifoldEdge";
      const uint32_t tri_offset = (uint32_t)tri_collider.size();
      for (uint32_t ti = 0; ti < nt; ++ti) {
        for (int k = 0; k < 3; ++k) {
          const Inc& e = edges[ekey(t[3 * ti + k], t[3 * ti + (k + 1) % 3])];
          uint32_t other = UINT32_MAX;
          if (e.t0 != UINT32_MAX && e.t0 != ti) other = e.t0;
          else if (e.t1 != UINT32_MAX && e.t1 != ti) other = e.t1;
          opp.push_back(other == UINT32_MAX ? UINT32_MAX : other + tri_offset);
        }
        for (int k = 0; k < 3; ++k) tri.push_back(t[3 * ti + k] + vertex_base);
        tri_collider.push_back(c);
      }
      tri_base += nt;
      vertex_base += nv;
    }
    n_vertices = vertex_base;
    n_triangles = (uint32_t)tri_collider.size();
    // vertex -> incident triangles; keep only closed fans (every neighbour vertex seen exactly twice)
    std::vector<std::vector<uint32_t>> lists(n_vertices);
    for (uint32_t ti = 0; ti < n_triangles; ++ti)
      for (int k = 0; k < 3; ++k) lists[tri[3 * ti + k]].push_back(ti);
    fan_offsets.assign(n_vertices + 1, 0);
    for (uint32_t v = 0; v < n_vertices; ++v) {
      std::vector<uint32_t> nb;
      for (uint32_t ti : lists[v])
        for (int k = 0; k < 3; ++k)
          if (tri[3 * ti + k] != v) nb.push_back(tri[3 * ti + k]);
      std::sort(nb.begin(), nb.end());
      bool closed = true;
      for (size_t i = 0; i < nb.size();) {
        size_t j = i;
        while (j < nb.size() && nb[j] == nb[i]) ++j;
        if (j - i > 2) return "missed non-manifoldness before";
        if (j - i != 2) closed = false;
        i = j;
      }
      if (closed)
        for (uint32_t ti : lists[v]) fan_tris.push_back(ti);
      fan_offsets[v + 1] = (uint32_t)fan_tris.size();
    }
    return "";
  }
};

// Flattened BVH for the device: node = {min[3], level-independent side implied by depth, kind, first, count}.
struct FlatBvh {
  int32_t level = 0;                // root side = 4^level leaves
  std::vector<int32_t> node_min;    // 3 per node
  std::vector<int32_t> node_max;    // 3 per node (inclusive, = min + side)
  std::vector<int32_t> node_first;  // inner: index into children (64 entries); leaf: index into tri_indices
  std::vector<int32_t> node_count;  // inner: -1 ; leaf: number of triangles
  std::vector<int32_t> children;    // 64 per inner node, -1 = none
  std::vector<uint32_t> tri_indices;
  bool empty() const { return node_count.empty(); }
};

struct LatticeAabb {
  int32_t mn[3], mx[3];
};
inline bool overlaps(const LatticeAabb& a, const int32_t* mn, const int32_t* mx) {  // util/src/aabb.rs:160-166 (inclusive)
  for (int k = 0; k < 3; ++k)
    if (a.mn[k] > mx[k] || mn[k] > a.mx[k]) return false;
  return true;
}

class BvhBuilder {
 public:
  // xpu/src/frame_input.rs:350-390: lattice AABB of every triangle over keyframes a (and b), grown by
  // `margin` (= forget distance), quantised with floor/ceil on the UNSCALED leaf size.
  static FlatBvh build(const HostTopology& topo, const float* va, const float* vb, float margin, float leaf_size, uint32_t leaf_threshold) {
    FlatBvh out;
    const uint32_t nt = topo.n_triangles;
    if (nt == 0) return out;
    std::vector<LatticeAabb> boxes(nt);
    LatticeAabb all;
    for (int k = 0; k < 3; ++k) { all.mn[k] = INT32_MAX; all.mx[k] = INT32_MIN; }
    for (uint32_t t = 0; t < nt; ++t) {
      float mn[3] = {FLT_MAX_, FLT_MAX_, FLT_MAX_}, mx[3] = {-FLT_MAX_, -FLT_MAX_, -FLT_MAX_};
      for (int c = 0; c < 3; ++c) {
        const uint32_t v = topo.tri[3 * t + c];
        for (int k = 0; k < 3; ++k) {
          mn[k] = std::min(mn[k], va[3 * v + k]);
          mx[k] = std::max(mx[k], va[3 * v + k]);
          if (vb) {
            mn[k] = std::min(mn[k], vb[3 * v + k]);
            mx[k] = std::max(mx[k], vb[3 * v + k]);
          }
        }
      }
      for (int k = 0; k < 3; ++k) {
        boxes[t].mn[k] = (int32_t)std::floor((mn[k] - margin) / leaf_size);
        boxes[t].mx[k] = (int32_t)std::ceil((mx[k] + margin) / leaf_size);
        all.mn[k] = std::min(all.mn[k], boxes[t].mn[k]);
        all.mx[k] = std::max(all.mx[k], boxes[t].mx[k]);
      }
    }
    int64_t longest = 0;
    for (int k = 0; k < 3; ++k) longest = std::max<int64_t>(longest, (int64_t)all.mx[k] - all.mn[k]);
    if (longest <= 0) return out;  // the reference reports "empty"
    int32_t ilog4 = 0;
    for (int64_t v = longest; v >= 4; v /= 4) ++ilog4;
    out.level = ilog4 + 1;
    const int32_t side = 1 << (2 * out.level);
    int32_t root_min[3];
    for (int k = 0; k < 3; ++k) root_min[k] = (all.mn[k] + all.mx[k] - side) / 2;  // truncating division like Rust
    std::vector<uint32_t> idx(nt);
    for (uint32_t t = 0; t < nt; ++t) idx[t] = t;
    BvhBuilder b{out, boxes, leaf_threshold};
    b.create(out.level, root_min, std::move(idx));
    return out;
  }

 private:
  static constexpr float FLT_MAX_ = std::numeric_limits<float>::max();
  FlatBvh& out;
  const std::vector<LatticeAabb>& boxes;
  uint32_t leaf_threshold;

  int32_t create(int32_t level, const int32_t* mn, std::vector<uint32_t> idx) {
    const int32_t me = (int32_t)out.node_count.size();
    const int32_t side = 1 << (2 * level);
    for (int k = 0; k < 3; ++k) { out.node_min.push_back(mn[k]); out.node_max.push_back(mn[k] + side); }
    out.node_first.push_back(0);
    out.node_count.push_back(0);
    if (level == 0 || idx.size() < leaf_threshold) {
      out.node_first[me] = (int32_t)out.tri_indices.size();
      out.node_count[me] = (int32_t)idx.size();
      out.tri_indices.insert(out.tri_indices.end(), idx.begin(), idx.end());
      return me;
    }
    const int32_t cbase = (int32_t)out.children.size();
    out.children.resize(out.children.size() + 64, -1);
    out.node_first[me] = cbase;
    out.node_count[me] = -1;
    const int32_t cl = level - 1, cside = 1 << (2 * cl);
    for (int child = 0; child < 64; ++child) {
      const int32_t cmn[3] = {mn[0] + (((child >> 4) & 3) << (2 * cl)), mn[1] + (((child >> 2) & 3) << (2 * cl)), mn[2] + ((child & 3) << (2 * cl))};
      const int32_t cmx[3] = {cmn[0] + cside, cmn[1] + cside, cmn[2] + cside};
      std::vector<uint32_t> sub;
      for (uint32_t t : idx)
        if (overlaps(boxes[t], cmn, cmx)) sub.push_back(t);
      if (!sub.empty()) {
        const int32_t ci = create(cl, cmn, std::move(sub));
        out.children[cbase + child] = ci;
      }
    }
    return me;
  }
  BvhBuilder(FlatBvh& o, const std::vector<LatticeAabb>& b, uint32_t lt) : out(o), boxes(b), leaf_threshold(lt) {}
};

// f32::total_cmp minimum
inline float total_min(float a, float b) {
  auto key = [](float f) { int32_t i; std::memcpy(&i, &f, 4); return i ^ (int32_t)(((uint32_t)(i >> 31)) >> 1); };
  return key(b) < key(a) ? b : a;
}

struct AdaptiveTimeStep {  // cpu/src/adaptive_time_step_state.rs:13-64
  float max_time_step = std::numeric_limits<float>::max();
  bool has_velocity = false, has_deformation = false, has_isolated = false, has_sound = false;
  float by_velocity = 0, by_deformation = 0, by_isolated = 0, by_sound = 0;
  std::deque<float> prior;
  // after an adaptive svb_advance the state machine ran on the device: `allowed()` then reports the device's last value
  bool has_override = false;
  float allowed_override = 0;
  float allowed_without_prior() const {
    const float fmax = std::numeric_limits<float>::max();
    float r = max_time_step;
    r = total_min(r, has_velocity ? by_velocity : fmax);
    r = total_min(r, has_deformation ? by_deformation : fmax);
    r = total_min(r, has_sound ? by_sound : fmax);
    r = total_min(r, has_isolated ? by_isolated : fmax);
    return r;
  }
  float allowed() const {
    if (has_override) return allowed_override;
    float r = allowed_without_prior();
    for (float p : prior) r = total_min(r, p);
    return r;
  }
  void push_current_limit() {
    if (prior.size() > 10) prior.pop_front();
    prior.push_back(allowed_without_prior());
  }
};

}  // namespace svbh
