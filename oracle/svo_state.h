// ORACLE — test infrastructure only (see svo_math.h header; parity unpinned end-to-end).
//
// CPU restatement of the reference's CPU back end (paths relative to /root/reference/rust/crates):
//   cpu/src/cpu_state.rs:15-197                 CpuState, from_io_state, to_io_state, produce_next_state
//   cpu/src/adaptive_time_step_state.rs:13-64   AdaptiveTimeStepState
//   cpu/src/phase/*.rs                          the 12 phases
//   xpu/src/frame_input.rs:33-52,266-390        FrameInput pieces used by the path
//   file_input/src/header.rs:10-71              InputConsts
//   file_frame/src/particles.rs:55-109          ParticleParameters / Particles
//
// Deterministic where the reference is not: stable sort, contributor lists in
// ascending particle order, new nodes indexed in ascending particle order.  Each is one legal
// outcome of the reference's unordered execution.
#pragma once
#include <atomic>
#include <deque>
#include <optional>
#include <unordered_map>
#include <vector>

#include "svo_math.h"
#include "svo_mesh.h"

namespace svo {

struct InputConsts {  // file_input/src/header.rs:10-18
  float grid_node_size = 0.5f;
  float leaf_size = 1.f;
  uint32_t leaf_threshold = 16;
  float simulation_scale = 1.f;
  uint32_t frames_per_second = 24;
  float domain_min[3] = {-100, -100, -100};
  float domain_max[3] = {100, 100, 100};
  float scaled_grid_node_size() const { return grid_node_size / simulation_scale; }       // :36-38
  Vec3f scaled_domain_min() const { return {domain_min[0] / simulation_scale, domain_min[1] / simulation_scale, domain_min[2] / simulation_scale}; }
  Vec3f scaled_domain_max() const { return {domain_max[0] / simulation_scale, domain_max[1] / simulation_scale, domain_max[2] / simulation_scale}; }
  float accept_distance() const { return scaled_grid_node_size() * 2.f; }                  // :60-62
  float forget_distance() const { return scaled_grid_node_size() * 2.2f; }                 // :64-66
  double seconds_per_frame() const { return 1. / (double)frames_per_second; }              // :68-70
};

struct Keyframe {  // xpu/src/frame_input.rs:54-66 (already divided by simulation_scale, :118-123)
  size_t frame = 0;
  Vec3f gravity;
  std::vector<uint32_t> particle_flags;
  std::vector<Vec3f> particle_goal_positions;
  std::vector<Vec3f> vertex_positions;
  std::vector<float> triangle_frictions;
  std::vector<float> triangle_dampings;
};

struct FrameInput {  // xpu/src/frame_input.rs:33-52
  size_t frame = 0;
  InputConsts consts;
  Topology topology;
  Bvh bvh;
  Keyframe a;
  std::optional<Keyframe> b;
  std::vector<Vec3f> vertex_velocities;

  // :266-278; returns false on WrongFrameLoaded
  bool frame_factor(double time, float& out) const {
    const double frame_time = time * (double)consts.frames_per_second;
    const size_t frame_low = (size_t)std::floor(frame_time);
    if (frame != frame_low) return false;
    out = (float)std::fmod(frame_time, 1.);
    return true;
  }
  void set_keyframes(size_t frame_, Keyframe a_, std::optional<Keyframe> b_);  // :334-390
};

struct ParticleParameters {  // file_frame/src/particles.rs:55-109, flattened
  float mass = 0, initial_volume = 0;
  bool has_viscosity = false;
  float viscosity_dynamic = 0, viscosity_bulk = 0;
  bool is_fluid = false;
  float mu = 0, lambda = 0;
  bool has_sand_alpha = false;
  float sand_alpha = 0;
  int32_t exponent = 0;
  float bulk_modulus = 0;
};

struct Particles {  // cpu/src/particles.rs:13-31
  std::vector<uint32_t> sort_map, reverse_sort_map;
  std::vector<uint32_t> flags;
  std::vector<ParticleParameters> parameters;
  std::vector<Vec3f> initial_positions, positions;
  std::vector<Mat3f> position_gradients;
  std::vector<Vec3f> velocities;
  std::vector<Mat3f> velocity_gradients;
  std::vector<float> elastic_energies;
  std::vector<uint32_t> collider_bits;
  size_t size() const { return flags.size(); }
};

struct GridKey {
  Vec3i node_id;
  uint32_t collider_bits;
  bool operator==(const GridKey& o) const { return node_id == o.node_id && collider_bits == o.collider_bits; }
};
struct GridKeyHash {
  size_t operator()(const GridKey& k) const {
    uint64_t h = 0;
    auto mix = [&](uint32_t v) { h = (((h << 5) | (h >> 59)) ^ v) * 0x517cc1b727220a95ull; };  // FxHash-style
    mix((uint32_t)k.node_id.x); mix((uint32_t)k.node_id.y); mix((uint32_t)k.node_id.z); mix(k.collider_bits);
    return (size_t)h;
  }
};
struct Vec3iHash {
  size_t operator()(const Vec3i& k) const { return GridKeyHash()(GridKey{k, 0}); }
};

struct GridNodes {  // cpu/src/grid_nodes.rs:21-33
  std::unordered_map<GridKey, uint32_t, GridKeyHash> map;
  std::unordered_map<Vec3i, std::vector<uint32_t>, Vec3iHash> multi_map;
  std::vector<GridKey> keys;
  std::vector<std::vector<uint32_t>> contributors;
  std::vector<float> masses;
  std::vector<Vec3f> velocities;
};

struct InterpolatedInput {  // cpu/src/interpolated_input.rs:11-23
  Vec3f gravity;
  std::vector<Vec3f> particle_goal_positions, vertex_positions, vertex_normals;
  std::vector<float> triangle_frictions, triangle_dampings;
  std::vector<Vec3f> triangle_normals;
};

struct AdaptiveTimeStepState {  // cpu/src/adaptive_time_step_state.rs:13-64
  float max_time_step = std::numeric_limits<float>::max();
  std::optional<float> by_velocity, by_deformation, by_isolated, by_sound;
  std::deque<float> prior;
  float allowed_without_prior() const;
  float allowed_time_step() const;
  void push_current_limit();
};

enum class Phase : int {  // cpu/src/phase/mod.rs:27-41
  InterpolateInput = 0, Sort, Collide, ExternalForce, UpdateGridNodes, LimitTimeStepBeforeForce, ScatterMomentum, MeldGrid, CollectVelocity,
  LimitTimeStepBeforeIntegrate, AdvanceParticles, CullParticles, COUNT
};

enum Status : int {  // cpu/src/errors.rs:9-36 mapped to the C-ABI convention
  OK = 0,
  ENERGY_ERROR = 8,  // simulation-level, state still valid (same bit as PARTICLE_CLOSE_TO_INVERTED)
  CANCELED = -1,
  ZERO_TIME_STEP = -2,
  FRAME_INPUT = -3,
  INTERPOLATED_INPUT_MISSING = -4,
};

struct CpuState {
  double time = 0;
  AdaptiveTimeStepState adaptive;
  Phase phase = Phase::InterpolateInput;
  Particles particles;
  GridNodes grid;
  std::optional<InterpolatedInput> interpolated;
  uint64_t substeps = 0;
  bool deterministic = true;  // sort contributor lists (a legal order of the reference's mutex pushes)

  int produce_next_state(const FrameInput& fi, double target_time, float max_time_step, bool adaptive_time_steps, const volatile int* cancel);
  int run_phase(const FrameInput& fi);

  int interpolate_input(const FrameInput& fi);
  void sort(float h);
  void collide(const FrameInput& fi);
  int external_force(const FrameInput& fi);
  void update_grid_nodes(float h);
  void limit_time_step_before_force(float h);
  void scatter_momentum(float h);
  void meld_grid();
  void collect_velocity(float h);
  void limit_time_step_before_integrate(float h);
  int advance_particles();
  void cull_particles(const FrameInput& fi);
};

}  // namespace svo
