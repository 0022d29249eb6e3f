// ORACLE — test infrastructure only (see svo_math.h header; parity unpinned end-to-end).
// Phase-by-phase CPU restatement of /root/reference/rust/crates/cpu/src/phase/*.rs.
#include "svo_state.h"

#include <omp.h>
#include <parallel/algorithm>

#include <cstring>
#include <numeric>

namespace svo {

// f32::total_cmp ordering
static inline int32_t total_key(float f) {
  int32_t b;
  std::memcpy(&b, &f, 4);
  return b ^ (int32_t)(((uint32_t)(b >> 31)) >> 1);
}
static inline bool total_less(float a, float b) { return total_key(a) < total_key(b); }
static inline float total_min(float a, float b) { return total_less(b, a) ? b : a; }
static inline float total_max(float a, float b) { return total_less(a, b) ? b : a; }

// ---------------------------------------------------------------- adaptive_time_step_state.rs
float AdaptiveTimeStepState::allowed_without_prior() const {  // :36-47
  const float fmax = std::numeric_limits<float>::max();
  float r = max_time_step;
  r = total_min(r, by_velocity.value_or(fmax));
  r = total_min(r, by_deformation.value_or(fmax));
  r = total_min(r, by_sound.value_or(fmax));
  r = total_min(r, by_isolated.value_or(fmax));
  return r;
}
float AdaptiveTimeStepState::allowed_time_step() const {  // :49-54
  float r = allowed_without_prior();
  for (float p : prior) r = total_min(r, p);
  return r;
}
void AdaptiveTimeStepState::push_current_limit() {  // :56-63
  if (prior.size() > 10) prior.pop_front();
  prior.push_back(allowed_without_prior());
}

// ---------------------------------------------------------------- xpu/src/frame_input.rs
void FrameInput::set_keyframes(size_t frame_, Keyframe a_, std::optional<Keyframe> b_) {
  frame = frame_;
  a = std::move(a_);
  b = std::move(b_);
  // linear_vertex_velocities :334-348
  vertex_velocities.assign(a.vertex_positions.size(), Vec3f::zeros());
  if (b)
    for (size_t i = 0; i < a.vertex_positions.size(); ++i) vertex_velocities[i] = (b->vertex_positions[i] - a.vertex_positions[i]) * (float)consts.frames_per_second;
  // update_bvh :350-390
  const float margin = consts.forget_distance();
  std::vector<AabbI> aabbs;
  aabbs.reserve(topology.triangle_indices.size());
  for (auto& t : topology.triangle_indices) {
    Vec3f pts[6];
    int n = 0;
    for (int k = 0; k < 3; ++k) {
      pts[n++] = a.vertex_positions[t[k]];
      if (b) pts[n++] = b->vertex_positions[t[k]];
    }
    aabbs.push_back(triangle_leaf_aabb(pts, n, margin, consts.leaf_size));
  }
  bvh.build(aabbs, consts.leaf_threshold);
}

// ---------------------------------------------------------------- cpu_state.rs:146-197
int CpuState::produce_next_state(const FrameInput& fi, double target_time, float max_time_step, bool adaptive_time_steps, const volatile int* cancel) {
  adaptive.max_time_step = max_time_step;
  while (time < target_time) {
    if (cancel && *cancel) return CANCELED;
    if (adaptive.allowed_time_step() == 0.f) return ZERO_TIME_STEP;
    const bool run = adaptive_time_steps || (phase != Phase::LimitTimeStepBeforeForce && phase != Phase::LimitTimeStepBeforeIntegrate);
    if (run) {
      const int rc = run_phase(fi);
      if (rc != OK) return rc;  // ENERGY_ERROR: caller still calls to_io_state (cpu_state.rs:178-184)
    }
    phase = (Phase)(((int)phase + 1) % (int)Phase::COUNT);
    if (phase == Phase::InterpolateInput) {
      time += (double)adaptive.allowed_time_step();
      ++substeps;
    }
  }
  return OK;
}

int CpuState::run_phase(const FrameInput& fi) {  // phase/mod.rs:52-75
  const float h = fi.consts.scaled_grid_node_size();
  switch (phase) {
    case Phase::InterpolateInput: return interpolate_input(fi);
    case Phase::Sort: sort(h); break;
    case Phase::Collide: collide(fi); break;
    case Phase::ExternalForce: return external_force(fi);
    case Phase::UpdateGridNodes: update_grid_nodes(h); break;
    case Phase::LimitTimeStepBeforeForce: limit_time_step_before_force(h); break;
    case Phase::ScatterMomentum: scatter_momentum(h); break;
    case Phase::MeldGrid: meld_grid(); break;
    case Phase::CollectVelocity: collect_velocity(h); break;
    case Phase::LimitTimeStepBeforeIntegrate: limit_time_step_before_integrate(h); break;
    case Phase::AdvanceParticles: return advance_particles();
    case Phase::CullParticles: cull_particles(fi); break;
    default: break;
  }
  return OK;
}

// ---------------------------------------------------------------- phase/interpolate_input.rs:18-107
int CpuState::interpolate_input(const FrameInput& fi) {
  const auto& tris = fi.topology.triangle_indices;
  const Keyframe& a = fi.a;
  const Keyframe& b = fi.b ? *fi.b : fi.a;
  float factor_b;
  if (!fi.frame_factor(time, factor_b)) return FRAME_INPUT;
  const float factor_a = 1.f - factor_b;

  InterpolatedInput out;
  out.gravity = factor_a * a.gravity + factor_b * b.gravity;
  const size_t np = a.particle_goal_positions.size();
  out.particle_goal_positions.resize(np);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < np; ++i) out.particle_goal_positions[i] = factor_a * a.particle_goal_positions[i] + factor_b * b.particle_goal_positions[i];
  const size_t nv = a.vertex_positions.size();
  out.vertex_positions.resize(nv);
  for (size_t i = 0; i < nv; ++i) out.vertex_positions[i] = factor_a * a.vertex_positions[i] + factor_b * b.vertex_positions[i];
  out.triangle_normals.resize(tris.size());
  for (size_t t = 0; t < tris.size(); ++t) {
    const Vec3f& pa = out.vertex_positions[tris[t].a];
    const Vec3f& pb = out.vertex_positions[tris[t].b];
    const Vec3f& pc = out.vertex_positions[tris[t].c];
    out.triangle_normals[t] = normalize_or_zero((pb - pa).cross(pc - pa), NORMALIZATION_EPS);
  }
  // angle-weighted vertex normals over closed fans only
  out.vertex_normals.resize(nv);
  for (size_t v = 0; v < nv; ++v) {
    Vec3f sum = Vec3f::zeros();
    for (uint32_t t : fi.topology.vertex_triangle_lists[v]) {
      uint32_t others[2];
      int n = 0;
      for (int k = 0; k < 3; ++k)
        if (tris[t][k] != (uint32_t)v && n < 2) others[n++] = tris[t][k];
      const Vec3f p = out.vertex_positions[v];
      const Vec3f pa = out.vertex_positions[others[0]];
      const Vec3f pb = out.vertex_positions[others[1]];
      const float ang = angle(pa - p, pb - p);
      sum += ang * out.triangle_normals[t];
    }
    out.vertex_normals[v] = normalize_or_zero(sum, NORMALIZATION_EPS);
  }
  out.triangle_frictions.resize(a.triangle_frictions.size());
  out.triangle_dampings.resize(a.triangle_dampings.size());
  for (size_t t = 0; t < a.triangle_frictions.size(); ++t) out.triangle_frictions[t] = factor_a * a.triangle_frictions[t] + factor_b * b.triangle_frictions[t];
  for (size_t t = 0; t < a.triangle_dampings.size(); ++t) out.triangle_dampings[t] = factor_a * a.triangle_dampings[t] + factor_b * b.triangle_dampings[t];
  interpolated = std::move(out);
  return OK;
}

// ---------------------------------------------------------------- phase/sort.rs:17-114
void CpuState::sort(float h) {
  const size_t n = particles.size();
  struct Entry { int32_t i, j, k; uint32_t idx; };
  std::vector<Entry> tmp(n);
#pragma omp parallel for schedule(static)
  for (size_t p = 0; p < n; ++p) {
    const Vec3i s = position_to_shift_quadratic(particles.positions[p], h);
    tmp[p] = {s.x, s.y, s.z, (uint32_t)p};
  }
  // par_sort_unstable_by_key: intra-cell order unspecified; (key, index) is one legal outcome.
  __gnu_parallel::sort(tmp.begin(), tmp.end(), [](const Entry& a, const Entry& b) {
    if (a.i != b.i) return a.i < b.i;
    if (a.j != b.j) return a.j < b.j;
    if (a.k != b.k) return a.k < b.k;
    return a.idx < b.idx;
  });
  auto permute = [&](auto& vec) {
    auto lookup = vec;
#pragma omp parallel for schedule(static)
    for (size_t p = 0; p < n; ++p) vec[p] = lookup[tmp[p].idx];
  };
  permute(particles.flags);
  permute(particles.positions);
  permute(particles.initial_positions);
  permute(particles.sort_map);
  permute(particles.parameters);
  permute(particles.position_gradients);
  permute(particles.velocities);
  permute(particles.velocity_gradients);
  permute(particles.collider_bits);
  particles.reverse_sort_map.resize(n);
  for (size_t cur = 0; cur < n; ++cur) particles.reverse_sort_map[particles.sort_map[cur]] = (uint32_t)cur;
}

// ---------------------------------------------------------------- phase/collide.rs:21-206
void CpuState::collide(const FrameInput& fi) {
  const float time_step = adaptive.allowed_time_step();
  const auto& tris = fi.topology.triangle_indices;
  const auto& opps_all = fi.topology.triangle_opposites;
  const auto& tri_collider = fi.topology.triangle_collider;
  const auto& vv = fi.vertex_velocities;
  const InterpolatedInput& in = *interpolated;
  const float leaf_size = fi.consts.leaf_size;
  const float forget = fi.consts.forget_distance();
  const float accept = fi.consts.accept_distance();
  const size_t n = particles.size();
#pragma omp parallel for schedule(dynamic, 1024)
  for (size_t pi = 0; pi < n; ++pi) {
    if (particles.flags[pi] & TOMBSTONED) continue;
    const Vec3f p = particles.positions[pi];
    Vec3f& velocity = particles.velocities[pi];
    uint32_t& bits = particles.collider_bits[pi];
    const Vec3i leaf{(int32_t)std::floor(p.x / leaf_size), (int32_t)std::floor(p.y / leaf_size), (int32_t)std::floor(p.z / leaf_size)};
    const std::vector<uint32_t>* to_check = fi.bvh.query(leaf);
    if (!to_check || to_check->empty()) { bits = 0; continue; }

    uint32_t closest[16];
    float min_dist[16];
    for (int c = 0; c < 16; ++c) { closest[c] = UINT32_MAX; min_dist[c] = std::numeric_limits<float>::max(); }
    for (uint32_t t : *to_check) {
      const Vec3f& n_ = in.triangle_normals[t];
      if (n_ == Vec3f::zeros()) continue;
      const float d = distance_to_triangle(p, in.vertex_positions[tris[t].a], in.vertex_positions[tris[t].b], in.vertex_positions[tris[t].c], n_);
      if (d >= forget) continue;
      const uint32_t c = tri_collider[t];
      if (d < min_dist[c]) { min_dist[c] = d; closest[c] = t; }
    }
    for (unsigned collider = 0; collider < 16; ++collider) {
      if (closest[collider] == UINT32_MAX) { collider_bits::set(bits, collider, -1); continue; }
      const uint32_t ct = closest[collider];
      const Triangle& tri = tris[ct];
      const Opposites& opps = opps_all[ct];
      const Vec3f& nrm = in.triangle_normals[ct];
      const Vec3f &a = in.vertex_positions[tri.a], &b = in.vertex_positions[tri.b], &c = in.vertex_positions[tri.c];
      const Vec3f &a_v = vv[tri.a], &b_v = vv[tri.b], &c_v = vv[tri.c];
      const Vec3f &a_n = in.vertex_normals[tri.a], &b_n = in.vertex_normals[tri.b], &c_n = in.vertex_normals[tri.c];
      const Vec3f ab_n = opps.ab != UINT32_MAX ? nrm + in.triangle_normals[opps.ab] : Vec3f::zeros();
      const Vec3f bc_n = opps.bc != UINT32_MAX ? nrm + in.triangle_normals[opps.bc] : Vec3f::zeros();
      const Vec3f ca_n = opps.ca != UINT32_MAX ? nrm + in.triangle_normals[opps.ca] : Vec3f::zeros();
      const Vec3f ab = a - b, bc = b - c, ca = c - a;
      const float area2 = nrm.dot(ca.cross(ab));
      const float a_bary = nrm.dot(bc.cross(c - p)) / area2;
      const float b_bary = nrm.dot(ca.cross(a - p)) / area2;
      const float c_bary = nrm.dot(ab.cross(b - p)) / area2;
      const bool sa = a_bary > 0.f, sb = b_bary > 0.f, sc = c_bary > 0.f;
      DistanceResult res;
      if (sa && sb && sc) {
        res = {std::fabs((p - a).dot(nrm)), nrm * (p - a).dot(nrm), nrm};
      } else {
        const DistanceResult r0 = segment_distance_result(p, a, b, a_n, ab_n, b_n);
        const DistanceResult r1 = segment_distance_result(p, b, c, b_n, bc_n, c_n);
        const DistanceResult r2 = segment_distance_result(p, c, a, c_n, ca_n, a_n);
        res = r0;  // min_by keeps the first of equal minima
        if (total_less(r1.distance, res.distance)) res = r1;
        if (total_less(r2.distance, res.distance)) res = r2;
      }
      if (res.normal == Vec3f::zeros()) { collider_bits::set(bits, collider, -1); continue; }
      const bool new_side = 0.f <= res.to_p.dot(res.normal);
      const int prior = collider_bits::get(bits, collider);
      if (prior < 0) {
        if (res.distance < accept) collider_bits::set(bits, collider, new_side ? 1 : 0);
        continue;
      }
      if ((prior == 1) == new_side) continue;
      if (res.distance > NORMALIZATION_EPS) {
        const Vec3f collider_velocity = a_v * a_bary + b_v * b_bary + c_v * c_bary;
        const Vec3f relative_velocity = velocity - collider_velocity;
        const Vec3f contact_normal = res.to_p / res.distance;
        const Vec3f normal_velocity = contact_normal * relative_velocity.dot(contact_normal);
        const Vec3f tangential_velocity = relative_velocity - normal_velocity;
        const float tnorm = tangential_velocity.norm();
        if (tnorm > NORMALIZATION_EPS) {
          const Vec3f tangent = tangential_velocity / tnorm;
          const Vec3f friction_impulse = tangent * std::fmin(in.triangle_frictions[ct] * res.distance / time_step, tnorm);
          velocity -= friction_impulse;
        }
        velocity -= std::fmin(in.triangle_dampings[ct], 1.f) * normal_velocity;
      }
      velocity -= res.to_p / time_step;
    }
  }
}

// ---------------------------------------------------------------- phase/external_force.rs:17-50
int CpuState::external_force(const FrameInput& fi) {
  const float time_step = adaptive.allowed_time_step();
  const Keyframe& a = fi.a;
  const Keyframe& b = fi.b ? *fi.b : fi.a;
  if (!interpolated) return INTERPOLATED_INPUT_MISSING;
  const InterpolatedInput& in = *interpolated;
  const size_t n = particles.size();
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) {
    if (particles.flags[i] & TOMBSTONED) continue;
    const size_t index = particles.sort_map[i];
    if ((a.particle_flags[index] & HAS_GOAL) && (b.particle_flags[index] & HAS_GOAL))
      particles.velocities[i] = (in.particle_goal_positions[index] - particles.positions[i]) / time_step;
    else
      particles.velocities[i] += time_step * in.gravity;
  }
  return OK;
}

// ---------------------------------------------------------------- phase/update_grid_nodes.rs:31-156
void CpuState::update_grid_nodes(float h) {
  // prune nodes whose contributor list was empty last substep, then re-index (:37-52)
  for (auto it = grid.map.begin(); it != grid.map.end();) {
    if (grid.contributors[it->second].empty()) it = grid.map.erase(it);
    else ++it;
  }
  {
    uint32_t i = 0;
    for (auto& kv : grid.map) kv.second = i++;
  }
  grid.contributors.resize(grid.map.size());
  for (auto& v : grid.contributors) v.clear();

  const size_t n = particles.size();
  const uint32_t existing = (uint32_t)grid.map.size();
  // existing nodes: parallel lookup + locked push (the reference's Mutex<SmallVec>); new keys go to
  // per-thread queues that are merged in thread order (the reference's collector thread, :72-91).
  std::vector<std::atomic_flag> locks(existing);
  for (auto& l : locks) l.clear();
  const int nthreads = omp_get_max_threads();
  std::vector<std::vector<std::pair<GridKey, uint32_t>>> fresh(nthreads);
#pragma omp parallel
  {
    auto& mine = fresh[omp_get_thread_num()];
#pragma omp for schedule(static)
    for (size_t pi = 0; pi < n; ++pi) {
      if (particles.flags[pi] & TOMBSTONED) continue;
      const Vec3i shift = position_to_shift_quadratic(particles.positions[pi], h);
      const uint32_t bits = particles.collider_bits[pi];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
          for (int k = 0; k < 3; ++k) {
            const GridKey key{{shift.x + i, shift.y + j, shift.z + k}, bits};
            auto it = grid.map.find(key);
            if (it != grid.map.end()) {
              const uint32_t gi = it->second;
              while (locks[gi].test_and_set(std::memory_order_acquire)) {}
              grid.contributors[gi].push_back((uint32_t)pi);
              locks[gi].clear(std::memory_order_release);
            } else {
              mine.emplace_back(key, (uint32_t)pi);
            }
          }
    }
  }
  for (auto& q : fresh)
    for (auto& kp : q) {
      auto ins = grid.map.emplace(kp.first, (uint32_t)grid.map.size());
      if (ins.second) grid.contributors.emplace_back();
      grid.contributors[ins.first->second].push_back(kp.second);
    }
  if (deterministic) {
#pragma omp parallel for schedule(dynamic, 256)
    for (size_t g = 0; g < grid.contributors.size(); ++g) std::sort(grid.contributors[g].begin(), grid.contributors[g].end());
  }
  // keys sorted by index (:133-143), multi map (:145-155)
  grid.keys.resize(grid.map.size());
  for (auto& kv : grid.map) grid.keys[kv.second] = kv.first;
  grid.multi_map.clear();
  for (uint32_t i = 0; i < grid.keys.size(); ++i) grid.multi_map[grid.keys[i].node_id].push_back(i);
}

// ---------------------------------------------------------------- phase/limit_time_step.rs:25-223
void CpuState::limit_time_step_before_force(float h) {
  const size_t n = particles.size();
  bool any = false;
  float by_sound = 0, by_isolated = 0;
#pragma omp parallel
  {
    bool l_any = false;
    float l_sound = 0, l_iso = 0;
#pragma omp for schedule(static)
    for (size_t pi = 0; pi < n; ++pi) {
      if (particles.flags[pi] & TOMBSTONED) continue;
      const ParticleParameters& prm = particles.parameters[pi];
      const Mat3f& F = particles.position_gradients[pi];
      // speed of sound (:37-114)
      const Vec3f s = svd3f(F).s;
      const float j = s.product();
      const bool xy_close = std::fabs(s.x - s.y) < SINGULAR_VALUE_SEPARATION;
      const bool yz_close = std::fabs(s.y - s.z) < SINGULAR_VALUE_SEPARATION;
      const bool zx_close = std::fabs(s.z - s.x) < SINGULAR_VALUE_SEPARATION;
      Vec3f first;
      Mat3f second;
      if (!prm.is_fluid) {
        first = first_piola_stress_neo_hookean_svd_diag(prm.mu, prm.lambda, s);
        second = second_derivative_neo_hookean_svd_diag(prm.mu, prm.lambda, s);
      } else {
        first = first_piola_stress_inviscid_svd_diag(prm.bulk_modulus, prm.exponent, s);
        second = second_derivative_inviscid_svd_diag(prm.bulk_modulus, prm.exponent, s);
      }
      // nalgebra mRC is 1-based (row, column)
      const float m11 = second(0, 0), m22 = second(1, 1), m33 = second(2, 2), m21 = second(1, 0), m32 = second(2, 1), m13 = second(0, 2);
      const float k[6] = {
          s.x * s.x * m11,
          s.y * s.y * m22,
          s.z * s.z * m33,
          s.y * s.y * (xy_close ? (first.x + s.x * m11 - s.y * m21) / 2.f / s.x : (s.x * first.x - s.y * first.y) / (s.x * s.x - s.y * s.y)),
          s.y * s.z * (yz_close ? (first.y + s.y * m22 - s.z * m32) / 2.f / s.y : (s.y * first.y - s.z * first.z) / (s.y * s.y - s.z * s.z)),
          s.z * s.x * (zx_close ? (first.z + s.x * m33 - s.x * m13) / 2.f / s.z : (s.z * first.z - s.x * first.x) / (s.z * s.z - s.x * s.x)),
      };
      float kmax = k[0];
      for (int q = 1; q < 6; ++q) kmax = total_max(kmax, k[q]);
      const float kappa = kmax / j;
      const float initial_density = prm.mass / prm.initial_volume;
      const float current_density = initial_density / j;
      const float c = std::sqrt(kappa / current_density);
      const float dt_sound = h / c;
      // isolated particle (:116-182)
      float dt_iso;
      if (!prm.is_fluid) {
        const float xi = 3.f / h / h;
        const float R = 1.f, K = 1.f, D = 3.f;
        dt_iso = std::sqrt(prm.mass / (prm.initial_volume * xi * (R - K / 2.f) * (prm.mu + D / 2.f * prm.lambda)));
      } else {
        const float jj = F.determinant();
        const float K = 6.f, D = 3.f;
        const float fst = d_inviscid_by_i3(prm.bulk_modulus, prm.exponent, jj);
        if (std::fabs(jj - 1.f) > SINGULAR_VALUE_SEPARATION) {
          dt_iso = h / jj * std::sqrt(initial_density * (jj - 1.f) / (K * fst * D));
        } else {
          const float snd = dd_inviscid_by_i3(prm.bulk_modulus, prm.exponent, jj);
          dt_iso = h * std::sqrt(initial_density / (K * snd * D));
        }
      }
      if (!l_any) { l_any = true; l_sound = dt_sound; l_iso = dt_iso; }
      else { l_sound = total_min(l_sound, dt_sound); l_iso = total_min(l_iso, dt_iso); }
    }
#pragma omp critical
    {
      if (l_any) {
        if (!any) { any = true; by_sound = l_sound; by_isolated = l_iso; }
        else { by_sound = total_min(by_sound, l_sound); by_isolated = total_min(by_isolated, l_iso); }
      }
    }
  }
  adaptive.by_sound = any ? std::optional<float>(by_sound) : std::nullopt;
  adaptive.by_isolated = any ? std::optional<float>(by_isolated) : std::nullopt;
  adaptive.push_current_limit();
}

void CpuState::limit_time_step_before_integrate(float h) {  // :187-223
  const size_t n = particles.size();
  bool any = false;
  float max_vel = 0, min_def = 0;
#pragma omp parallel
  {
    bool l_any = false;
    float l_vel = 0, l_def = 0;
#pragma omp for schedule(static)
    for (size_t pi = 0; pi < n; ++pi) {
      if (particles.flags[pi] & TOMBSTONED) continue;
      const float vel = particles.velocities[pi].norm();
      const Mat3f& C = particles.velocity_gradients[pi];
      float def = 0.2f / std::fmax(std::fabs(C.m[0]), 1e-8f);
      for (int q = 1; q < 9; ++q) def = total_min(def, 0.2f / std::fmax(std::fabs(C.m[q]), 1e-8f));
      if (!l_any) { l_any = true; l_vel = vel; l_def = def; }
      else { l_vel = total_max(l_vel, vel); l_def = total_min(l_def, def); }
    }
#pragma omp critical
    {
      if (l_any) {
        if (!any) { any = true; max_vel = l_vel; min_def = l_def; }
        else { max_vel = total_max(max_vel, l_vel); min_def = total_min(min_def, l_def); }
      }
    }
  }
  adaptive.by_velocity = (any && max_vel != 0.f) ? std::optional<float>(0.5f * h / max_vel) : std::nullopt;
  adaptive.by_deformation = any ? std::optional<float>(min_def) : std::nullopt;
  adaptive.push_current_limit();
}

// ---------------------------------------------------------------- phase/scatter_momentum.rs:22-93
void CpuState::scatter_momentum(float h) {
  const float scaling = adaptive.allowed_time_step() * 4.f / powi(h, 2);
  const size_t g_n = grid.map.size();
  grid.masses.assign(g_n, 0.f);
  grid.velocities.assign(g_n, Vec3f::zeros());
#pragma omp parallel for schedule(dynamic, 256)
  for (size_t g = 0; g < g_n; ++g) {
    const Vec3i node_id = grid.keys[g].node_id;
    float mass = 0.f;
    Vec3f velocity = Vec3f::zeros();
    for (uint32_t pidx : grid.contributors[g]) {
      const Vec3f normalized = particles.positions[pidx] / h;
      const Vec3f to_node_n = Vec3f{(float)node_id.x, (float)node_id.y, (float)node_id.z} - normalized;
      const float weight = kernel_quadratic(to_node_n.x) * kernel_quadratic(to_node_n.y) * kernel_quadratic(to_node_n.z);
      const Vec3f to_node = to_node_n * h;
      const ParticleParameters& prm = particles.parameters[pidx];
      const Mat3f& C = particles.velocity_gradients[pidx];
      const Mat3f& F = particles.position_gradients[pidx];
      Vec3f imparted = (particles.velocities[pidx] + C * to_node) * prm.mass;
      const Mat3f stress = prm.is_fluid ? first_piola_stress_inviscid(prm.bulk_modulus, prm.exponent, F) : first_piola_stress_neo_hookean(prm.mu, prm.lambda, F);
      if (prm.has_viscosity) {
        const Mat3f cauchy = cauchy_stress_general_viscosity(prm.viscosity_dynamic, prm.viscosity_bulk, C);
        imparted -= cauchy * (to_node * (scaling * F.determinant() * prm.initial_volume));
      }
      imparted -= stress * (F.transpose() * (to_node * (scaling * prm.initial_volume)));
      imparted *= weight;
      mass += weight * prm.mass;
      velocity += imparted;
    }
    grid.masses[g] = mass;
    grid.velocities[g] = velocity;
  }
}

// ---------------------------------------------------------------- phase/meld_grid.rs:16-69
void CpuState::meld_grid() {
  const std::vector<float> masses = grid.masses;
  const std::vector<Vec3f> velocities = grid.velocities;
  const size_t g_n = grid.keys.size();
#pragma omp parallel for schedule(static)
  for (size_t index = 0; index < g_n; ++index) {
    const GridKey& key = grid.keys[index];
    float mass = grid.masses[index];
    Vec3f velocity = grid.velocities[index];
    for (uint32_t other : grid.multi_map.find(key.node_id)->second) {
      if (other == index) continue;
      if (!collider_bits::compatible(key.collider_bits, grid.keys[other].collider_bits)) continue;
      mass += masses[other];
      velocity += velocities[other];
    }
    if (mass > 0.f) velocity /= mass;
    else velocity = Vec3f::zeros();
    grid.masses[index] = mass;
    grid.velocities[index] = velocity;
  }
}

// ---------------------------------------------------------------- phase/collect_velocity.rs:19-75
void CpuState::collect_velocity(float h) {
  const size_t n = particles.size();
#pragma omp parallel for schedule(static)
  for (size_t pi = 0; pi < n; ++pi) {
    if (particles.flags[pi] & TOMBSTONED) continue;
    const Vec3f position = particles.positions[pi];
    const uint32_t bits = particles.collider_bits[pi];
    Vec3f velocity = Vec3f::zeros();
    Mat3f vgrad = Mat3f::zeros();
    const Vec3f normalized = position / h;
    const Vec3f shift{std::floor(normalized.x - 0.5f), std::floor(normalized.y - 0.5f), std::floor(normalized.z - 0.5f)};
    const Vec3f shifted = normalized - shift;
    float xw[3], yw[3], zw[3];
    for (int i = 0; i < 3; ++i) {
      xw[i] = kernel_quadratic(shifted.x - (float)i);
      yw[i] = kernel_quadratic(shifted.y - (float)i);
      zw[i] = kernel_quadratic(shifted.z - (float)i);
    }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        for (int k = 0; k < 3; ++k) {
          const float weight = xw[i] * yw[j] * zw[k];
          const Vec3i node_id{(int32_t)shift.x + i, (int32_t)shift.y + j, (int32_t)shift.z + k};
          const Vec3f node_pos = Vec3f{(float)node_id.x, (float)node_id.y, (float)node_id.z} * h;
          const Vec3f to_node = node_pos - position;
          const auto it = grid.map.find(GridKey{node_id, bits});  // .expect("missing node")
          const Vec3f gv = grid.velocities[it->second];
          velocity += gv * weight;
          vgrad += outer(gv * weight, to_node);
        }
    vgrad *= 4.f / h / h;
    particles.velocities[pi] = velocity;
    particles.velocity_gradients[pi] = vgrad;
  }
}

// ---------------------------------------------------------------- phase/advance_particles.rs:17-93
int CpuState::advance_particles() {
  const float time_step = adaptive.allowed_time_step();
  const size_t n = particles.size();
  int failed = 0;
#pragma omp parallel for schedule(static) reduction(| : failed)
  for (size_t pi = 0; pi < n; ++pi) {
    if (particles.flags[pi] & TOMBSTONED) continue;
    const ParticleParameters& prm = particles.parameters[pi];
    Mat3f& F = particles.position_gradients[pi];
    particles.positions[pi] += particles.velocities[pi] * time_step;
    F += particles.velocity_gradients[pi] * F * time_step;
    if (!prm.is_fluid) {
      if (prm.has_sand_alpha) {
        Svd3f svd = svd3f(F);
        const Vec3f e{std::log(svd.s.x), std::log(svd.s.y), std::log(svd.s.z)};
        const float e_tr = e.sum();
        const Vec3f e_hat = e - Vec3f::repeat(e_tr / 3.f);
        const float e_hat_norm = e_hat.norm();
        if (e_tr < 0.f && e_hat_norm > 0.f) {
          // assert!(mu > 0.) in the reference
          if (e_hat_norm != 0.f) {
            const float delta_gamma = e_hat_norm + (3.f * prm.lambda + 2.f * prm.mu) / 2.f / prm.mu * e_tr * prm.sand_alpha;
            if (delta_gamma > 0.f) {
              const Vec3f big_h = e - delta_gamma / e_hat_norm * e_hat;
              svd.s = {std::exp(big_h.x), std::exp(big_h.y), std::exp(big_h.z)};
              F = svd.recompose();
            }
          }
        } else {
          F = svd.u * svd.v_t;
        }
      }
      float energy;
      if (try_elastic_energy_neo_hookean(prm.mu, prm.lambda, F, energy)) {
        particles.elastic_energies[pi] = energy;
      } else {
        particles.flags[pi] |= FAILED;
        failed |= 1;
      }
    } else {
      Svd3f svd = svd3f(F);
      const float iso = std::pow(svd.s.product(), 1.f / 3.f);
      svd.s = {iso, iso, iso};
      F = svd.recompose();
      particles.elastic_energies[pi] = elastic_energy_inviscid(prm.bulk_modulus, prm.exponent, F);
    }
  }
  return failed ? ENERGY_ERROR : OK;
}

// ---------------------------------------------------------------- phase/cull_particles.rs:17-41
void CpuState::cull_particles(const FrameInput& fi) {
  const Vec3f mn = fi.consts.scaled_domain_min(), mx = fi.consts.scaled_domain_max();
  const size_t n = particles.size();
#pragma omp parallel for schedule(static)
  for (size_t pi = 0; pi < n; ++pi) {
    if (particles.flags[pi] & TOMBSTONED) continue;
    const Vec3f& p = particles.positions[pi];
    const bool within = p.x > mn.x && p.x < mx.x && p.y > mn.y && p.y < mx.y && p.z > mn.z && p.z < mx.z;
    if (!within) particles.flags[pi] |= TOMBSTONED;
  }
}

}  // namespace svo
