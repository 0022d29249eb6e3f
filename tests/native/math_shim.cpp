// Test-only shim: compiles the PRODUCT's device math header (squishy_volumes_b200/csrc/svb_math.cuh)
// for the host so its SVD / return mapping / stress / time-step bounds can be checked on CPU against
// numpy and against the oracle without a GPU.  Not part of the shipped library.
#include <cstring>
#include "../../squishy_volumes_b200/csrc/svb_math.cuh"
#include "../../squishy_volumes_b200/csrc/svb_host.h"
using namespace svb;
extern "C" {
uint32_t shim_node_id_to_murmur(int32_t x, int32_t y, int32_t z, uint32_t seed) { return node_id_to_murmur(x, y, z, seed); }
void shim_svd3(const float* F, float* U, float* S, float* V) {
  M3 f; std::memcpy(f.m, F, 36);
  Svd3 r = svd3(f);
  std::memcpy(U, r.u.m, 36); std::memcpy(V, r.v.m, 36);
  S[0] = r.s.x; S[1] = r.s.y; S[2] = r.s.z;
}
int shim_return_map(uint32_t flags, float p0, float p1, float alpha, float* F, float* energy) {
  M3 f; std::memcpy(f.m, F, 36);
  bool ok = return_map_and_energy(flags, p0, p1, alpha, f, *energy);
  std::memcpy(F, f.m, 36);
  return ok ? 1 : 0;
}
void shim_stress(int fluid, float p0, float p1, const float* F, float* P) {
  M3 f; std::memcpy(f.m, F, 36);
  M3 r = fluid ? first_piola_inviscid(p0, (int)p1, f) : first_piola_neo_hookean(p0, p1, f);
  std::memcpy(P, r.m, 36);
}
void shim_viscous(float dyn, float bulk, const float* C, float* out) {
  M3 c; std::memcpy(c.m, C, 36);
  M3 r = viscous_cauchy(dyn, bulk, c);
  std::memcpy(out, r.m, 36);
}
void shim_limits(int fluid, float p0, float p1, float mass, float vol, const float* F, float h, float* out2) {
  M3 f; std::memcpy(f.m, F, 36);
  ParticleLimits l = particle_time_step_limits(fluid != 0, p0, p1, mass, vol, f, h);
  out2[0] = l.by_sound; out2[1] = l.by_isolated;
}
float shim_kernel_quadratic(float x) { return kernel_quadratic(x); }
int shim_bits_get(uint32_t b, uint32_t c) { return bits_get(b, c); }
uint32_t shim_bits_set(uint32_t b, uint32_t c, int s) { return bits_set(b, c, s); }
int shim_bits_compatible(uint32_t a, uint32_t b) { return bits_compatible(a, b) ? 1 : 0; }

// host topology + BVH (product code in svb_host.h)
struct ShimMesh { svbh::HostTopology topo; svbh::FlatBvh bvh; };
ShimMesh* shim_mesh_build(uint32_t n_col, const uint32_t* nv, const uint32_t* nt, const uint32_t* tris, const float* va, const float* vb, float margin, float leaf_size,
                          uint32_t leaf_threshold, int* err) {
  auto* m = new ShimMesh();
  const std::string e = m->topo.build(n_col, nv, nt, tris);
  *err = e.empty() ? 0 : 1;
  if (e.empty()) m->bvh = svbh::BvhBuilder::build(m->topo, va, vb, margin, leaf_size, leaf_threshold);
  return m;
}
void shim_mesh_destroy(ShimMesh* m) { delete m; }
void shim_mesh_topology(const ShimMesh* m, uint32_t* tri, uint32_t* opp, uint32_t* collider, uint32_t* fan_sizes) {
  std::memcpy(tri, m->topo.tri.data(), m->topo.tri.size() * 4);
  std::memcpy(opp, m->topo.opp.data(), m->topo.opp.size() * 4);
  std::memcpy(collider, m->topo.tri_collider.data(), m->topo.tri_collider.size() * 4);
  for (uint32_t v = 0; v < m->topo.n_vertices; ++v) fan_sizes[v] = m->topo.fan_offsets[v + 1] - m->topo.fan_offsets[v];
}
int shim_mesh_level(const ShimMesh* m) { return m->bvh.level; }
// same integer descent as the device query (svb_kernels.cuh: bvh_query)
uint64_t shim_mesh_query(const ShimMesh* m, const int32_t* q, uint32_t* out, uint64_t cap) {
  const svbh::FlatBvh& B = m->bvh;
  if (B.empty()) return 0;
  for (int k = 0; k < 3; ++k)
    if (q[k] < B.node_min[k] || q[k] > B.node_max[k]) return 0;
  int cur = 0;
  auto leaf = [&](int node) -> uint64_t {
    const int first = B.node_first[node], count = B.node_count[node];
    for (int i = 0; i < count && (uint64_t)i < cap; ++i) out[i] = B.tri_indices[first + i];
    return (uint64_t)count;
  };
  if (B.node_count[0] >= 0) return leaf(0);
  const uint32_t u[3] = {(uint32_t)(q[0] - B.node_min[0]), (uint32_t)(q[1] - B.node_min[1]), (uint32_t)(q[2] - B.node_min[2])};
  for (int lvl = B.level - 1; lvl >= 0; --lvl) {
    const uint32_t child = (((u[0] >> (2 * lvl)) & 3u) << 4) | (((u[1] >> (2 * lvl)) & 3u) << 2) | ((u[2] >> (2 * lvl)) & 3u);
    const int ci = B.children[B.node_first[cur] + child];
    if (ci < 0) return 0;
    if (B.node_count[ci] >= 0) return leaf(ci);
    cur = ci;
  }
  return 0;
}
}
