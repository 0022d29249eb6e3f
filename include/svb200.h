/* svb200.h — C ABI of the B200-native MPM substep back end.
 *
 * This is the drop-in seam of Algebraic-UG/squishy_volumes' back-end boundary
 * (`enum ComputeState { Cpu(CpuState), Gpu(GpuState) }`, core/src/compute_thread.rs:100-163):
 * a third arm `B200(B200State)` binds exactly these entry points.  Plain pointers and sizes only;
 * no torch / CUDA types cross the boundary.  All reference paths are relative to
 * /root/reference/rust/crates.
 *
 * Status convention (every call returning int32_t):
 *    0  ok
 *   >0  simulation-level error, the state is still valid and downloadable (the reference's inner
 *       `Err`, cpu/src/cpu_state.rs:178-184).  Bit values follow gpu/src/status.rs:34-37.
 *   <0  fatal (the reference's outer `Err`, cpu/src/errors.rs:9-36).
 * `svb_last_error` returns a UTF-8 description of the most recent non-zero status.
 */
#ifndef SVB200_H
#define SVB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* gpu/src/status.rs:34-37 (simulation-level bits) */
#define SVB_TABLE_TRIES_EXCEEDED 1
#define SVB_TABLE_ENTRY_MISSING 2
#define SVB_INDIRECT_LIMIT_EXCEEDED 4
#define SVB_PARTICLE_CLOSE_TO_INVERTED 8 /* = cpu EnergyError::PositionGradientNonPositive */
/* cpu/src/errors.rs:9-36 (fatal) */
#define SVB_CANCELED (-1)
#define SVB_ZERO_TIME_STEP (-2)
#define SVB_FRAME_INPUT (-3)          /* xpu FrameInputError::WrongFrameLoaded */
#define SVB_INPUT_MISSING (-4)        /* keyframes / topology not set */
#define SVB_TOO_MANY_COLLIDERS (-5)   /* xpu InputError::TooManyColliders (> 16) */
#define SVB_BAD_MESH (-6)             /* mesh_util TopologyError */
#define SVB_CUDA_ERROR (-7)
#define SVB_BAD_ARGUMENT (-8)
#define SVB_KEY_RANGE (-9)            /* live particles span more grid blocks than a 63-bit bin key holds */
#define SVB_COMM_ERROR (-10)

/* file_input/src/header.rs:10-18 (InputConsts) */
typedef struct SvbConsts {
  float grid_node_size;
  float leaf_size;
  uint32_t leaf_threshold;
  float simulation_scale;
  uint32_t frames_per_second;
  float domain_min[3];
  float domain_max[3];
} SvbConsts;

/* file_frame/src/particles.rs:93-109 (Particles), struct of arrays in ORIGINAL particle order.
 * ParticleParameters (:55-91) is flattened: which of mu/lambda/sand_alpha or bulk_modulus/exponent
 * and viscosity_* are live follows the flag bits (:23-33).  Vectors are [f32;3], matrices
 * [[f32;3];3] = three columns (column-major nalgebra Matrix3).  NULL output pointers are skipped. */
typedef struct SvbParticles {
  uint64_t n;
  uint32_t* flags;
  float* mass;
  float* initial_volume;
  float* mu_or_bulk_modulus;
  float* lambda_or_exponent;
  float* sand_alpha;
  float* viscosity_dynamic;
  float* viscosity_bulk;
  float* initial_positions;   /* n*3 */
  float* positions;           /* n*3 */
  float* position_gradients;  /* n*9 */
  float* velocities;          /* n*3 */
  float* velocity_gradients;  /* n*9 */
  float* elastic_energies;    /* n   */
  uint32_t* collider_bits;    /* n   */
} SvbParticles;

/* xpu/src/frame_input.rs:54-66 (InputInterpolationPoint); positions already divided by
 * simulation_scale (:118-123).  Particle arrays are in original particle order. */
typedef struct SvbKeyframe {
  float gravity[3];
  const uint32_t* particle_flags;          /* n or NULL (= no goals) */
  const float* particle_goal_positions;    /* n*3 or NULL */
  const float* vertex_positions;           /* V*3 */
  const float* triangle_frictions;         /* T */
  const float* triangle_dampings;          /* T */
} SvbKeyframe;

/* file_frame/src/grid_nodes.rs:9-15 (GridNodes); caller-allocated, capacity from svb_grid_count.
 * contributor_counts may be NULL (diagnostic: 1 for every emitted node). */
typedef struct SvbGrid {
  uint64_t n;
  int32_t* node_ids;          /* n*3 */
  uint32_t* collider_bits;    /* n */
  float* masses;              /* n */
  float* velocities;          /* n*3 */
  uint32_t* contributor_counts;
} SvbGrid;

typedef struct SvbHandle SvbHandle;

/* core/src/api_impl/context.rs:10-12 + gpu/src/context.rs (GpuContext::available_gpus):
 * newline-separated "index: name" of the visible CUDA devices; returns the device count or <0. */
int32_t svb_available_devices(char* out, size_t cap);

/* CpuState::from_io_state (cpu/src/cpu_state.rs:26-69) / GpuState::from_io_state
 * (gpu/src/gpu_state.rs:38-45): copies the particle state to `device`; the caller keeps its buffers. */
int32_t svb_create(const SvbConsts* consts, const SvbParticles* particles, double time, int32_t device,
                   SvbHandle** out);
void svb_destroy(SvbHandle* h);
/* The same, over `n_dev` GPUs of the box behind ONE handle, for a caller that drives the back end from a single thread like
 * core/src/compute_thread.rs:100-163 does (`const int* devices, int n_dev`).  The state is cut into slabs along x
 * inside the library (by particle count, on grid-block planes); svb_upload / svb_set_topology / svb_set_keyframes / svb_advance
 * / svb_download / svb_destroy and the scalar getters work on the returned handle exactly as on a single-device one (download
 * re-assembles original particle order); every call fans out to one slab rank per device for its duration.  The ranks exchange
 * halo sums, migrating particles and the adaptive-step limits through each other's HBM (cudaDeviceEnablePeerAccess: the devices
 * must be peers).  The introspection entry points below (grid, binning, snapshot) are single-device only. */
int32_t svb_create_multi(const SvbConsts* consts, const SvbParticles* particles, double time, const int32_t* devices,
                         int32_t n_dev, SvbHandle** out);
/* Replaces the resident particle state of an existing handle (H2D of a new IoState) and restarts its clock,
 * step history and error words; constants, collider input and — on slab ranks — the communicator and the
 * peer mailboxes are kept.  The compute thread calls `from_io_state` whenever a frame is (re)loaded
 * (core/src/compute_thread.rs:100-117); with a session-lived handle that is this call. */
int32_t svb_upload(SvbHandle* h, const SvbParticles* particles, double time);

/* Topology::new over the collider inputs of frame 0 (mesh_util/src/mesh.rs:30-142,
 * xpu/src/frame_input.rs:176-184): per-collider vertex and triangle counts, triangles as LOCAL
 * vertex indices concatenated in collider order. */
int32_t svb_set_topology(SvbHandle* h, uint32_t n_colliders, const uint32_t* num_vertices,
                         const uint32_t* num_triangles, const uint32_t* triangles);

/* FrameInput::load result (xpu/src/frame_input.rs:206-232, 334-390): a = keyframe `frame`,
 * b = keyframe `frame+1` or NULL.  Rebuilds vertex velocities and the integer-lattice BVH. */
int32_t svb_set_keyframes(SvbHandle* h, uint64_t frame, const SvbKeyframe* a, const SvbKeyframe* b);

/* CpuState::produce_next_state loop (cpu/src/cpu_state.rs:146-197): substeps until
 * time >= target_time.  `cancel` (may be NULL) is polled every substep (Harness::check,
 * xpu/src/harness.rs:69-75); `progress` (may be NULL) receives the milliseconds into the frame
 * (Harness::step_to). */
int32_t svb_advance(SvbHandle* h, double target_time, float max_time_step, int32_t adaptive_time_steps,
                    const volatile int32_t* cancel, void (*progress)(void* user, size_t ms), void* user);

/* CpuState::to_io_state (cpu/src/cpu_state.rs:71-135): original particle order. */
int32_t svb_download(SvbHandle* h, SvbParticles* out);
int64_t svb_grid_count(SvbHandle* h);
int32_t svb_download_grid(SvbHandle* h, SvbGrid* out);

double svb_time(const SvbHandle* h);
uint64_t svb_substeps(const SvbHandle* h);            /* fills the `last_frame_substeps` TODO (core/src/compute_thread.rs:190) */
float svb_allowed_time_step(const SvbHandle* h);      /* AdaptiveTimeStepState::allowed_time_step */
uint32_t svb_status(const SvbHandle* h);              /* accumulated simulation-level status bits */
const char* svb_last_error(const SvbHandle* h);
uint64_t svb_kernel_launches(const SvbHandle* h);     /* kernels launched by this handle so far */
/* Device time of the last svb_advance call: CUDA events recorded on the handle's own stream around
 * the substep loop (gpu/src/gpu_state.rs `run_step` profiler scope analogue). */
float svb_last_advance_ms(const SvbHandle* h);

/* ---- introspection used by the parity tests (integer stages must be bit-exact) ---- */
/* Current (binned) order: sort_map[current] = original index (cpu/src/particles.rs:13-17),
 * cell[current*3..] = base node (i,j,k) = floor(x/h - 1/2) (cpu/src/kernels.rs:46-49). */
int32_t svb_binning(SvbHandle* h, uint32_t* sort_map, int32_t* cells);
/* Active (block, collider_bits) layers of the last substep: block = node_id >> 2 per axis. */
int64_t svb_active_block_count(SvbHandle* h);
int32_t svb_active_blocks(SvbHandle* h, int32_t* block_ids /* n*3 */, uint32_t* collider_bits /* n */);
/* Per-stage device times of the last svb_advance call, in milliseconds (gpu_profile.csv analogue,
 * gpu/src/profiler_output.rs:93-192).  Returns the number of stages; names are static strings. */
int32_t svb_stage_times(SvbHandle* h, const char** names, float* ms, int32_t cap);
void svb_enable_stage_timing(SvbHandle* h, int32_t on);
/* Slab ranks: milliseconds the exchange kernels of this rank's device have spent WAITING since the last reset, measured on the device
 * (%globaltimer around the spin loops): [0] halo sender for P2G's boundary tiles, [1] halo receiver for the neighbours' columns,
 * [2] migration sender for G2P's boundary tiles, [3] migration receiver for the neighbours' rows, [4] ... for every rank's error
 * word, [5] adaptive steps: the other ranks' limit reductions.  [0] and [2] are waits of the second stream (hidden behind the interior
 * tiles); [1], [3], [4], [5] stall the main stream.  With cap >= 8 also [6] the number of particle tiles and [7] of active grid tiles
 * of the last substep the host has looked at (plain counts).  Synchronises.  Returns the number of entries.  No reference counterpart. */
int32_t svb_exchange_waits(SvbHandle* h, double* ms, int32_t cap, int32_t reset);
/* Options by name: "store_grid" (CpuRunParameters::store_grid: exact contributor masks for svb_download_grid),
 * "global_particles" (slab ranks: size of the original-order keyframe arrays), "murmur_table_hash" (hash the tile table with the
 * reference's murmur node key instead of the 64-bit mixer; results do not depend on it). */
void svb_set_option(SvbHandle* h, const char* name, double value);

/* The reference's (dormant) `node_ids_to_murmur` stage, gpu/src/node_ids_to_murmur/mod.rs with the key of gpu/src/util.rs:79-100:
 * murmur3_x86_32 over the three ordered-u32 (x ^ 0x8000_0000) little-endian coordinates of every node id, seed 0
 * (`hashes_node_ids`) and seed = collider bits (`hashes_node_ids_and_bits`); either output may be NULL.  The same key hashes the
 * tile table when the option "murmur_table_hash" is on (svb_set_option). */
int32_t svb_node_ids_to_murmur(int32_t device, const int32_t* node_ids /* n*3 */, const uint32_t* collider_bits /* n or NULL */, uint64_t n,
                               uint32_t* hashes_node_ids, uint32_t* hashes_node_ids_and_bits);

/* ---- device-resident entry points (bench: inputs already in HBM) ---- */
/* Copies the current device state into a device-side snapshot / restores it (no host traffic). */
int32_t svb_snapshot(SvbHandle* h);
int32_t svb_restore(SvbHandle* h);

/* ---- multi-GPU slab decomposition (no reference counterpart) ---- */
/* 128-byte NCCL unique id created on rank 0 and distributed by the launcher. */
int32_t svb_comm_unique_id(uint8_t out[128]);
/* Joins this handle (already holding ITS slab's particles) to an n_ranks communicator. Slabs are
 * contiguous ranges of block-x; `original_offset` is added to local particle indices to form the
 * global original index carried by migrating particles. */
int32_t svb_comm_init(SvbHandle* h, const uint8_t unique_id[128], int32_t rank, int32_t n_ranks,
                      int32_t slab_lo_block_x, int32_t slab_hi_block_x, uint64_t original_offset);
uint64_t svb_particle_count(const SvbHandle* h);      /* particles currently resident on this rank */
/* Global original index of every uploaded row (call right after svb_create, before any substep) when the
 * rank's rows are not a contiguous range of the global particle order. */
int32_t svb_set_original_indices(SvbHandle* h, const uint32_t* original_index, uint64_t n);
/* Download in resident order together with the global original index of each row. */
int32_t svb_download_resident(SvbHandle* h, SvbParticles* out, uint64_t* original_index);
/* Live particles per block column on this rank: counts[0] = columns below first_col, counts[1 + k] = column first_col + k,
 * counts[n_cols + 1] = columns above (n_cols + 2 entries).  The planning input of a rebalance. */
int32_t svb_slab_histogram(SvbHandle* h, int32_t first_col, uint32_t n_cols, uint64_t* counts);
/* Collective, between two svb_advance calls (peer-memory path): this rank's block columns become [new_lo, new_hi).  The new
 * slabs must be contiguous, keep the outer ends, and every cut must stay strictly inside the two old slabs it separates, so
 * that rows only move to an adjacent rank; rows outside the new range are handed over at once ("rebalanced by
 * particle count every K substeps"). */
int32_t svb_slab_rebalance(SvbHandle* h, int32_t new_lo_block_x, int32_t new_hi_block_x);

#ifdef __cplusplus
}
#endif
#endif /* SVB200_H */
