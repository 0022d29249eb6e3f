/* svb_files.h — the data formats on either side of the substep path, host-only C ABI.
 *
 *   frame files    rust/crates/file_frame/src/io_state.rs:31-75      IoState::write / IoState::read
 *   input files    rust/crates/file_input/src/{lib,header,frame,collider_inputs,reading,writing}.rs
 *   scene set-up   rust/crates/core/src/initialization.rs:84-277     initialize_io_state
 *   keyframes      rust/crates/xpu/src/frame_input.rs:66-135         InputInterpolationPoint::new
 *
 * Container (rust/crates/file_util/src/lib.rs:26-101): 32 magic bytes, 64 version bytes (the crate version, zero padded;
 * a reader rejects any other version), then a bincode 1.3.3 body — little endian, fixed-width integers, u64 lengths for
 * Vec / String / map, u8 tag for Option, u32 variant index for enums, fixed-size arrays without a length, usize as u64.
 * An input file ends with the bincode Vec<u64> of frame offsets followed by the 8-byte little-endian offset of that index.
 *
 * All functions return 0 or a negative SVBF_* code; svbf_last_error() gives the message of the calling thread's last
 * failure.  Arrays are caller-allocated unless stated otherwise; particle arrays use the flattened SvbParticles layout of
 * svb200.h (which flags select mu/lambda/sand_alpha, bulk_modulus/exponent, viscosity: file_frame/src/particles.rs:36-54).
 */
#ifndef SVB_FILES_H
#define SVB_FILES_H

#include <stddef.h>
#include <stdint.h>

#include "svb200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SVBF_IO_ERROR (-20)             /* open / read / write / rename failed (file_frame Error::{Create,Open,Write,Move}) */
#define SVBF_MAGIC_MISMATCH (-21)       /* file_util Error::MagicMismatch */
#define SVBF_VERSION_MISMATCH (-22)     /* file_util Error::VersionMismatch */
#define SVBF_FORMAT (-23)               /* truncated or malformed bincode body (Error::Deserialize) */
#define SVBF_FRAME_NOT_AVAILABLE (-24)  /* file_input InputError::FrameNotAvailable */
#define SVBF_OBJECT_ERROR (-25)         /* ObjectError::{ObjectNotInHeader, ObjectChangedType}, StateInitializationError::Object* */
#define SVBF_PARTICLE_INVALID (-26)     /* StateInitializationError::ParticleInvalid (unknown flags, solid xor fluid, parameter bounds) */
#define SVBF_MISSING_INPUT (-27)        /* StateInitializationError::MissingInput / ParticleInvalid::MissingInput */
#define SVBF_LENGTH_MISMATCH (-28)      /* FrameVerifcationError::LengthMismatch, FrameInputError::AttributeLengthMismatch */
#define SVBF_COLLIDER_INPUT_MISSING (-29) /* FrameVerifcationError::ColliderInputMissing */
#define SVBF_BAD_ARGUMENT (-30)

const char* svbf_last_error(void);
/* the version string this build writes and accepts by default: rust/crates/file_util/Cargo.toml:3 */
const char* svbf_default_version(void);

/* ------------------------------------------------------------------ frame files ("frame_%05d.bin", cache/src/util.rs:9-11) */
/* IoState::write: writes `<dir>/temp.bin`, then renames it to `path`.  grid == NULL stores `grid_nodes: None`.
 * version == NULL uses svbf_default_version(). */
int32_t svbf_frame_write(const char* path, const char* version, double time, const SvbParticles* particles, const SvbGrid* grid, uint64_t* written_bytes);
/* cache/src/util.rs:9-11 */
int32_t svbf_frame_path(const char* cache_dir, uint64_t frame, char* out, size_t cap);

typedef struct SvbfFrame SvbfFrame;
/* IoState::read: parses the whole file; the counts size the caller's arrays for svbf_frame_copy. */
int32_t svbf_frame_open(const char* path, const char* version, SvbfFrame** out);
double svbf_frame_time(const SvbfFrame* f);
uint64_t svbf_frame_particle_count(const SvbfFrame* f);
int64_t svbf_frame_grid_count(const SvbfFrame* f);   /* -1: grid_nodes is None */
/* copies into caller arrays (NULL members are skipped; grid may be NULL) */
int32_t svbf_frame_copy(const SvbfFrame* f, SvbParticles* particles, SvbGrid* grid);
void svbf_frame_close(SvbfFrame* f);

/* ------------------------------------------------------------------ input files ("simulation_input.bin") */
#define SVBF_OBJECT_PARTICLES 0
#define SVBF_OBJECT_COLLIDER 1

typedef struct SvbfInput SvbfInput;
/* InputReader::new: checks magic + version, reads the frame index and the header. */
int32_t svbf_input_open(const char* path, const char* version, SvbfInput** out);
void svbf_input_close(SvbfInput* in);
uint64_t svbf_input_size(const SvbfInput* in);          /* bytes */
uint64_t svbf_input_frame_count(const SvbfInput* in);   /* InputReader::len */
int32_t svbf_input_consts(const SvbfInput* in, SvbConsts* out);
/* InputRanges::new (file_input/src/header.rs:118-160): objects in name order, particle / vertex / triangle ranges cumulative */
uint64_t svbf_input_total_particles(const SvbfInput* in);
uint64_t svbf_input_total_vertices(const SvbfInput* in);
uint64_t svbf_input_total_triangles(const SvbfInput* in);
uint32_t svbf_input_object_count(const SvbfInput* in);
/* kind: SVBF_OBJECT_*; particles: count = particles, start = first particle; collider: count = vertices, count2 = triangles,
 * start = first vertex, start2 = first triangle */
int32_t svbf_input_object(const SvbfInput* in, uint32_t index, char* name, size_t name_cap, int32_t* kind, uint64_t* count, uint64_t* count2, uint64_t* start,
                          uint64_t* start2);
/* collider topology from the collider inputs of frame 0, colliders in name order (xpu/src/frame_input.rs:168-184):
 * per collider vertex / triangle counts and the LOCAL triangle indices, concatenated — the arguments of svb_set_topology.
 * Pass NULL arrays to query n_colliders only. */
int32_t svbf_input_topology(SvbfInput* in, uint32_t* n_colliders, uint32_t* num_vertices, uint32_t* num_triangles, uint32_t* triangles);
/* initialize_io_state (core/src/initialization.rs:84-277) from the header and frame 0: every member array of `out` must hold
 * svbf_input_total_particles() rows (NULL members are skipped).  F = the upper 3x3 of each transform, x = its translation
 * / simulation_scale, V0 = (size / scale)^3, m = V0 * density, mu / lambda from (E, nu), C = 0, energies = 0, bits = 0. */
int32_t svbf_input_initialize(SvbfInput* in, SvbParticles* out);
/* InputInterpolationPoint::new for `frame` (positions divided by simulation_scale): flags[n], goals[n*3],
 * vertex_positions[V*3], frictions[T], dampings[T]; NULL outputs are skipped. */
int32_t svbf_input_keyframe(SvbfInput* in, uint64_t frame, float gravity[3], uint32_t* particle_flags, float* particle_goal_positions, float* vertex_positions,
                            float* triangle_frictions, float* triangle_dampings);

/* ---- recording side (InputWriter, file_input/src/writing.rs:27-73) */
typedef struct SvbfObjectDesc {
  const char* name;
  int32_t kind;        /* SVBF_OBJECT_* */
  uint64_t count;      /* particles, or vertices */
  uint64_t count2;     /* triangles (collider) */
} SvbfObjectDesc;
/* file_input/src/frame.rs:10-26; NULL = None.  Every present array has `n` rows. */
typedef struct SvbfParticlesInput {
  const char* name;
  uint64_t n;
  const uint32_t* flags;              /* required */
  const float* transforms;            /* n*16: [[f32;4];4], row [3] = translation */
  const float* sizes;
  const float* densities;
  const float* youngs_moduluses;
  const float* poissons_ratios;
  const float* initial_positions;     /* n*3 */
  const float* initial_velocities;    /* n*3 */
  const float* viscosities_dynamic;
  const float* viscosities_bulk;
  const uint32_t* exponents;
  const float* bulk_moduluses;
  const float* sand_alphas;
  const float* goal_positions;        /* n*3 */
} SvbfParticlesInput;
/* file_input/src/collider_inputs.rs:12-18 */
typedef struct SvbfColliderInput {
  const char* name;
  uint64_t num_vertices, num_triangles;
  const float* vertex_positions;      /* V*3 */
  const uint32_t* triangle_indices;   /* T*3 */
  const float* triangle_frictions;    /* T */
  const float* triangle_dampings;     /* T */
} SvbfColliderInput;

typedef struct SvbfInputWriter SvbfInputWriter;
int32_t svbf_input_writer_open(const char* path, const char* version, const SvbConsts* consts, const SvbfObjectDesc* objects, uint32_t n_objects, SvbfInputWriter** out);
/* record_frame: verifies the frame against the header (frame.rs:50-187: unknown objects, type changes, lengths, every
 * collider of the header present), then appends it. */
int32_t svbf_input_writer_frame(SvbfInputWriter* w, const float gravity[3], const SvbfParticlesInput* particles, uint32_t n_particles_inputs, const SvbfColliderInput* colliders,
                                uint32_t n_collider_inputs);
/* flush: writes the frame index and its offset, closes the file and frees the writer (also on failure) */
int32_t svbf_input_writer_finish(SvbfInputWriter* w);

#ifdef __cplusplus
}
#endif
#endif
