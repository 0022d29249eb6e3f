// ORACLE — test infrastructure only (see svo_math.h header; parity unpinned end-to-end).
// Flat C API over the restatement so tests/ and bench.py's cpu_baseline leg can drive it with ctypes.
// The struct layouts are the ones include/svb200.h declares, so one set of ctypes mirrors serves both.
#include <omp.h>

#include <cstring>

#include "svo_state.h"

using namespace svo;

extern "C" {

struct SvoConsts {
  float grid_node_size, leaf_size;
  uint32_t leaf_threshold;
  float simulation_scale;
  uint32_t frames_per_second;
  float domain_min[3], domain_max[3];
};
struct SvoParticles {
  uint64_t n;
  uint32_t* flags;
  float *mass, *initial_volume, *mu_or_bulk_modulus, *lambda_or_exponent, *sand_alpha, *viscosity_dynamic, *viscosity_bulk;
  float *initial_positions, *positions, *position_gradients, *velocities, *velocity_gradients, *elastic_energies;
  uint32_t* collider_bits;
};
struct SvoKeyframe {
  float gravity[3];
  const uint32_t* particle_flags;
  const float* particle_goal_positions;
  const float* vertex_positions;
  const float* triangle_frictions;
  const float* triangle_dampings;
};
struct SvoGrid {
  uint64_t n;
  int32_t* node_ids;
  uint32_t* collider_bits;
  float* masses;
  float* velocities;
  uint32_t* contributor_counts;  // oracle-only extra: lets tests restrict to nodes with >= 1 contributor
};

struct SvoHandle {
  CpuState state;
  FrameInput fi;
  size_t n_vertices = 0, n_triangles = 0;
};

static InputConsts to_consts(const SvoConsts* c) {
  InputConsts k;
  k.grid_node_size = c->grid_node_size;
  k.leaf_size = c->leaf_size;
  k.leaf_threshold = c->leaf_threshold;
  k.simulation_scale = c->simulation_scale;
  k.frames_per_second = c->frames_per_second;
  for (int i = 0; i < 3; ++i) { k.domain_min[i] = c->domain_min[i]; k.domain_max[i] = c->domain_max[i]; }
  return k;
}

// CpuState::from_io_state (cpu/src/cpu_state.rs:26-69)
SvoHandle* svo_create(const SvoConsts* consts, const SvoParticles* p, double time) {
  auto* h = new SvoHandle();
  h->fi.consts = to_consts(consts);
  Particles& P = h->state.particles;
  const size_t n = p->n;
  P.sort_map.resize(n);
  for (size_t i = 0; i < n; ++i) P.sort_map[i] = (uint32_t)i;
  P.reverse_sort_map = P.sort_map;
  P.flags.assign(p->flags, p->flags + n);
  P.parameters.resize(n);
  for (size_t i = 0; i < n; ++i) {
    ParticleParameters& q = P.parameters[i];
    const uint32_t f = p->flags[i];
    q.mass = p->mass[i];
    q.initial_volume = p->initial_volume[i];
    q.has_viscosity = (f & USE_VISCOSITY) != 0;
    q.viscosity_dynamic = p->viscosity_dynamic ? p->viscosity_dynamic[i] : 0.f;
    q.viscosity_bulk = p->viscosity_bulk ? p->viscosity_bulk[i] : 0.f;
    q.is_fluid = (f & IS_FLUID) != 0;
    if (q.is_fluid) {
      q.bulk_modulus = p->mu_or_bulk_modulus[i];
      q.exponent = (int32_t)p->lambda_or_exponent[i];
    } else {
      q.mu = p->mu_or_bulk_modulus[i];
      q.lambda = p->lambda_or_exponent[i];
      q.has_sand_alpha = (f & USE_SAND_ALPHA) != 0;
      q.sand_alpha = p->sand_alpha ? p->sand_alpha[i] : 0.f;
    }
  }
  auto v3 = [&](const float* src, std::vector<Vec3f>& dst) {
    dst.resize(n);
    if (src) std::memcpy(dst.data(), src, n * 12);
  };
  auto m3 = [&](const float* src, std::vector<Mat3f>& dst) {
    dst.resize(n);
    if (src) std::memcpy(dst.data(), src, n * 36);
  };
  v3(p->initial_positions, P.initial_positions);
  v3(p->positions, P.positions);
  m3(p->position_gradients, P.position_gradients);
  v3(p->velocities, P.velocities);
  m3(p->velocity_gradients, P.velocity_gradients);
  P.elastic_energies.assign(n, 0.f);
  if (p->elastic_energies) std::memcpy(P.elastic_energies.data(), p->elastic_energies, n * 4);
  P.collider_bits.assign(n, 0u);
  if (p->collider_bits) std::memcpy(P.collider_bits.data(), p->collider_bits, n * 4);
  h->state.time = time;
  return h;
}
void svo_destroy(SvoHandle* h) { delete h; }
void svo_set_deterministic(SvoHandle* h, int on) { h->state.deterministic = on != 0; }
void svo_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int svo_max_threads() { return omp_get_max_threads(); }

// Topology::new over the collider inputs of frame 0 (xpu/src/frame_input.rs:176-184): triangles hold
// per-collider LOCAL vertex indices, concatenated in collider order.
int svo_set_topology(SvoHandle* h, uint32_t n_colliders, const uint32_t* num_vertices, const uint32_t* num_triangles, const uint32_t* triangles) {
  if (n_colliders > 16) return -5;  // InputError::TooManyColliders
  std::vector<TopologyInput> inputs(n_colliders);
  size_t off = 0, nv = 0;
  for (uint32_t c = 0; c < n_colliders; ++c) {
    inputs[c].collider = c;
    inputs[c].num_vertices = num_vertices[c];
    inputs[c].triangles.resize(num_triangles[c]);
    for (uint32_t t = 0; t < num_triangles[c]; ++t) inputs[c].triangles[t] = {triangles[3 * (off + t)], triangles[3 * (off + t) + 1], triangles[3 * (off + t) + 2]};
    off += num_triangles[c];
    nv += num_vertices[c];
  }
  const std::string err = h->fi.topology.build(inputs);
  h->n_vertices = nv;
  h->n_triangles = off;
  return err.empty() ? 0 : -6;
}

static Keyframe to_keyframe(const SvoHandle* h, size_t frame, const SvoKeyframe* k) {
  Keyframe f;
  f.frame = frame;
  f.gravity = {k->gravity[0], k->gravity[1], k->gravity[2]};
  const size_t n = h->state.particles.size();
  f.particle_flags.assign(n, 0u);
  if (k->particle_flags) std::memcpy(f.particle_flags.data(), k->particle_flags, n * 4);
  f.particle_goal_positions.assign(n, Vec3f::zeros());
  if (k->particle_goal_positions) std::memcpy(f.particle_goal_positions.data(), k->particle_goal_positions, n * 12);
  f.vertex_positions.resize(h->n_vertices);
  if (h->n_vertices) std::memcpy(f.vertex_positions.data(), k->vertex_positions, h->n_vertices * 12);
  f.triangle_frictions.assign(h->n_triangles, 0.f);
  f.triangle_dampings.assign(h->n_triangles, 0.f);
  if (h->n_triangles) {
    std::memcpy(f.triangle_frictions.data(), k->triangle_frictions, h->n_triangles * 4);
    std::memcpy(f.triangle_dampings.data(), k->triangle_dampings, h->n_triangles * 4);
  }
  return f;
}

// FrameInput::load result (xpu/src/frame_input.rs:206-232): a = keyframe `frame`, b = `frame+1` or null.
int svo_set_keyframes(SvoHandle* h, uint64_t frame, const SvoKeyframe* a, const SvoKeyframe* b) {
  std::optional<Keyframe> kb;
  if (b) kb = to_keyframe(h, frame + 1, b);
  h->fi.set_keyframes(frame, to_keyframe(h, frame, a), std::move(kb));
  return 0;
}

int svo_advance(SvoHandle* h, double target_time, float max_time_step, int adaptive, const volatile int* cancel) {
  return h->state.produce_next_state(h->fi, target_time, max_time_step, adaptive != 0, cancel);
}
double svo_time(const SvoHandle* h) { return h->state.time; }
uint64_t svo_substeps(const SvoHandle* h) { return h->state.substeps; }
float svo_allowed_time_step(const SvoHandle* h) { return h->state.adaptive.allowed_time_step(); }

// CpuState::to_io_state (cpu/src/cpu_state.rs:71-135): original particle order through reverse_sort_map.
void svo_download(const SvoHandle* h, SvoParticles* out) {
  const Particles& P = h->state.particles;
  const size_t n = P.size();
  for (size_t o = 0; o < n; ++o) {
    const size_t i = P.reverse_sort_map[o];
    if (out->flags) out->flags[o] = P.flags[i];
    if (out->elastic_energies) out->elastic_energies[o] = P.elastic_energies[i];
    if (out->collider_bits) out->collider_bits[o] = P.collider_bits[i];
    if (out->positions) std::memcpy(out->positions + 3 * o, &P.positions[i], 12);
    if (out->initial_positions) std::memcpy(out->initial_positions + 3 * o, &P.initial_positions[i], 12);
    if (out->velocities) std::memcpy(out->velocities + 3 * o, &P.velocities[i], 12);
    if (out->position_gradients) std::memcpy(out->position_gradients + 9 * o, &P.position_gradients[i], 36);
    if (out->velocity_gradients) std::memcpy(out->velocity_gradients + 9 * o, &P.velocity_gradients[i], 36);
  }
}
// current (sorted) order: sort_map[current] = original
void svo_sort_map(const SvoHandle* h, uint32_t* out) { std::memcpy(out, h->state.particles.sort_map.data(), h->state.particles.size() * 4); }

uint64_t svo_grid_count(const SvoHandle* h) { return h->state.grid.keys.size(); }
void svo_download_grid(const SvoHandle* h, SvoGrid* g) {
  const GridNodes& G = h->state.grid;
  for (size_t i = 0; i < G.keys.size(); ++i) {
    g->node_ids[3 * i] = G.keys[i].node_id.x;
    g->node_ids[3 * i + 1] = G.keys[i].node_id.y;
    g->node_ids[3 * i + 2] = G.keys[i].node_id.z;
    g->collider_bits[i] = G.keys[i].collider_bits;
    g->masses[i] = i < G.masses.size() ? G.masses[i] : 0.f;
    if (i < G.velocities.size()) std::memcpy(g->velocities + 3 * i, &G.velocities[i], 12);
    if (g->contributor_counts) g->contributor_counts[i] = (uint32_t)G.contributors[i].size();
  }
}

// ------------------------------------------------------------------ unit exports (known-answer tests)
float svo_kernel_linear(float x) { return kernel_linear(x); }
float svo_kernel_quadratic(float x) { return kernel_quadratic(x); }
float svo_kernel_cubic(float x) { return kernel_cubic(x); }
void svo_shift_quadratic(uint64_t n, const float* positions, float h, int32_t* out) {
  for (uint64_t i = 0; i < n; ++i) {
    const Vec3i s = position_to_shift_quadratic({positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]}, h);
    out[3 * i] = s.x; out[3 * i + 1] = s.y; out[3 * i + 2] = s.z;
  }
}
int svo_bits_get(uint32_t bits, uint32_t c) { return collider_bits::get(bits, c); }
uint32_t svo_bits_set(uint32_t bits, uint32_t c, int s) { collider_bits::set(bits, c, s); return bits; }
int svo_bits_compatible(uint32_t a, uint32_t b) { return collider_bits::compatible(a, b) ? 1 : 0; }

static Mat3<double> md(const double* F) { Mat3<double> m; for (int i = 0; i < 9; ++i) m.m[i] = F[i]; return m; }
static void mo(const Mat3<double>& m, double* o) { for (int i = 0; i < 9; ++i) o[i] = m.m[i]; }
double svo_lame_mu(double E, double nu) { return lame_mu(E, nu); }
double svo_lame_lambda(double E, double nu) { return lame_lambda(E, nu); }
double svo_energy_neo_hookean(double mu, double lambda, const double* F) { return elastic_energy_neo_hookean(mu, lambda, md(F)); }
void svo_stress_neo_hookean(double mu, double lambda, const double* F, double* P) { mo(first_piola_stress_neo_hookean(mu, lambda, md(F)), P); }
void svo_stress_neo_hookean_svd_diag(double mu, double lambda, const double* s, double* out) {
  const auto r = first_piola_stress_neo_hookean_svd_diag(mu, lambda, Vec3<double>{s[0], s[1], s[2]});
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void svo_second_neo_hookean_svd_diag(double mu, double lambda, const double* s, double* out) { mo(second_derivative_neo_hookean_svd_diag(mu, lambda, Vec3<double>{s[0], s[1], s[2]}), out); }
double svo_energy_inviscid(double K, int exponent, const double* F) { return elastic_energy_inviscid(K, exponent, md(F)); }
void svo_stress_inviscid(double K, int exponent, const double* F, double* P) { mo(first_piola_stress_inviscid(K, exponent, md(F)), P); }
void svo_stress_inviscid_svd_diag(double K, int exponent, const double* s, double* out) {
  const auto r = first_piola_stress_inviscid_svd_diag(K, exponent, Vec3<double>{s[0], s[1], s[2]});
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void svo_second_inviscid_svd_diag(double K, int exponent, const double* s, double* out) { mo(second_derivative_inviscid_svd_diag(K, exponent, Vec3<double>{s[0], s[1], s[2]}), out); }
void svo_viscosity_stress(double dynamic, double bulk, const double* C, double* out) { mo(cauchy_stress_general_viscosity(dynamic, bulk, md(C)), out); }
double svo_det3(const double* F) { return md(F).determinant(); }
void svo_svd3(const double* F, double* U, double* S, double* V) {
  Svd3d d;
  svd3(F, d);
  std::memcpy(U, d.U, 72); std::memcpy(S, d.S, 24); std::memcpy(V, d.V, 72);
}
float svo_distance_to_triangle(const float* p, const float* a, const float* b, const float* c, const float* n) {
  return distance_to_triangle({p[0], p[1], p[2]}, {a[0], a[1], a[2]}, {b[0], b[1], b[2]}, {c[0], c[1], c[2]}, {n[0], n[1], n[2]});
}

// BVH: build from float triangles exactly like triangles_to_leaf_aabbs (bounding_volume_hierarchy.rs:219-237)
struct SvoBvh { Bvh bvh; };
SvoBvh* svo_bvh_build(uint64_t n_tri, const float* tri_vertices /* n*9 */, float leaf_size, float margin, uint32_t leaf_threshold) {
  auto* b = new SvoBvh();
  std::vector<AabbI> aabbs(n_tri);
  for (uint64_t t = 0; t < n_tri; ++t) {
    Vec3f pts[3];
    for (int k = 0; k < 3; ++k) pts[k] = {tri_vertices[9 * t + 3 * k], tri_vertices[9 * t + 3 * k + 1], tri_vertices[9 * t + 3 * k + 2]};
    aabbs[t] = triangle_leaf_aabb(pts, 3, margin, leaf_size);
  }
  b->bvh.build(aabbs, leaf_threshold);
  return b;
}
void svo_bvh_destroy(SvoBvh* b) { delete b; }
uint32_t svo_bvh_level(const SvoBvh* b) { return b->bvh.level; }
uint64_t svo_bvh_num_nodes(const SvoBvh* b) { return b->bvh.nodes.size(); }
// returns count, copies up to cap indices
uint64_t svo_bvh_query(const SvoBvh* b, const int32_t* q, uint32_t* out, uint64_t cap) {
  const auto* r = b->bvh.query({q[0], q[1], q[2]});
  if (!r) return 0;
  for (uint64_t i = 0; i < r->size() && i < cap; ++i) out[i] = (*r)[i];
  return r->size();
}
uint64_t svo_handle_bvh_query(const SvoHandle* h, const int32_t* q, uint32_t* out, uint64_t cap) {
  const auto* r = h->fi.bvh.query({q[0], q[1], q[2]});
  if (!r) return 0;
  for (uint64_t i = 0; i < r->size() && i < cap; ++i) out[i] = (*r)[i];
  return r->size();
}
// Topology accessors for tests
uint64_t svo_topology_counts(const SvoHandle* h, uint64_t* n_vertices) { *n_vertices = h->n_vertices; return h->n_triangles; }
void svo_topology_get(const SvoHandle* h, uint32_t* tri, uint32_t* opp, uint32_t* collider, uint32_t* fan_sizes) {
  const Topology& T = h->fi.topology;
  for (size_t t = 0; t < T.triangle_indices.size(); ++t) {
    tri[3 * t] = T.triangle_indices[t].a; tri[3 * t + 1] = T.triangle_indices[t].b; tri[3 * t + 2] = T.triangle_indices[t].c;
    opp[3 * t] = T.triangle_opposites[t].ab; opp[3 * t + 1] = T.triangle_opposites[t].bc; opp[3 * t + 2] = T.triangle_opposites[t].ca;
    collider[t] = T.triangle_collider[t];
  }
  for (size_t v = 0; v < T.vertex_triangle_lists.size(); ++v) fan_sizes[v] = (uint32_t)T.vertex_triangle_lists[v].size();
}
// interpolated mesh of the most recent InterpolateInput phase (for stage-level parity tests)
int svo_interpolated_mesh(const SvoHandle* h, float* vertex_positions, float* vertex_normals, float* triangle_normals) {
  if (!h->state.interpolated) return -1;
  const InterpolatedInput& in = *h->state.interpolated;
  if (vertex_positions) std::memcpy(vertex_positions, in.vertex_positions.data(), in.vertex_positions.size() * 12);
  if (vertex_normals) std::memcpy(vertex_normals, in.vertex_normals.data(), in.vertex_normals.size() * 12);
  if (triangle_normals) std::memcpy(triangle_normals, in.triangle_normals.data(), in.triangle_normals.size() * 12);
  return 0;
}

}  // extern "C"
