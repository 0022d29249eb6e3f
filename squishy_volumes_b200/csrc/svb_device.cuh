// svb_device.cuh — device-side data layout of the B200 MPM substep.
//
// Particles: nine arrays of 16-byte quads in HBM, back to back in one allocation (see Field below).  Two such buffers
// ping-pong across the physical re-bin (G2P reads a row through the inverse map of the counting sort and
// writes it to its binned slot in the other buffer).
// Grid: sparse set of 4x4x4-node blocks, one dense 64-node float4 (px,py,pz,m) tile per active
// (block, collider-bits layer), found through an open-addressing table of 64-bit tile keys.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "svb_math.cuh"

namespace svb {

// ---- particle words.  The 34 state words of a particle are stored in nine QUADS of four words (the last one holds two): quad q of
// particle i is ONE 16-byte element, ((float4*)base)[q * cap + i], so the kernels that touch every particle move a particle with
// 9 (P2G) or 6 + 9 (G2P) 16-byte accesses per lane instead of 31 / 30 + 34 four-byte ones.  Measured on B200 with a G2P-shaped
// gather / scatter pass (profiles/gather_bench.cu, 8 M particles, 5 CTAs per SM): 594 us by single words, 474 us by quads.  The words
// are ordered by use: G2P reads quads 0..5 (it replaces v and C), P2G reads all nine.
//   quad 0: x (3), flags | 1: F0..F3 | 2: F4..F7 | 3: F8, mass, V0, mu or K | 4: lambda or gamma, alpha, eta, zeta | 5: collider bits,
//   original index, v.x, v.y | 6: v.z, C0, C1, C2 | 7: C3..C6 | 8: C7, C8, -, -
enum Field : int {
  PX = 0,      // position            (3)
  PFLAGS = 3,  // u32
  PF = 4,      // position gradient F (9, column-major)
  PMASS = 13,
  PVOL = 14,   // initial volume
  PP0 = 15,    // mu | bulk modulus
  PP1 = 16,    // lambda | exponent
  PALPHA = 17, // sand alpha
  PVD = 18,    // viscosity dynamic
  PVB = 19,    // viscosity bulk
  PBITS = 20,  // u32 collider bits
  PORIG = 21,  // u32 original index (the reference's sort_map)
  PV = 22,     // velocity            (3)
  PC = 25,     // velocity gradient C (9, column-major)
  NFIELDS = 34,
  NQUADS = 9,
  NWORDS = 36, // words per particle in memory (two unused words in the last quad)
};

// word `field` of consecutive particles is 16 bytes apart
template <class T>
struct FieldRef {
  T* p;
  __host__ __device__ T& operator[](size_t i) const { return p[i * 4]; }
};
struct ParticleBuf {
  uint32_t* base;
  size_t cap;
  __host__ __device__ FieldRef<float> f(int field) const { return FieldRef<float>{reinterpret_cast<float*>(base) + (size_t)(field >> 2) * cap * 4 + (field & 3)}; }
  __host__ __device__ FieldRef<uint32_t> u(int field) const { return FieldRef<uint32_t>{base + (size_t)(field >> 2) * cap * 4 + (field & 3)}; }
  __host__ __device__ float4* q(int quad) const { return reinterpret_cast<float4*>(base) + (size_t)quad * cap; }
};

// ---- grid tiles.  A tile = one 4x4x4-node block of one collider-bits layer; its 64-bit key packs the
// absolute block coordinates (node >> 2, biased by 2^16, 17 bits per axis) and the layer id (13 bits):
//   key = bx | by | bz | layer      (most significant first)
// Tiles live in an open-addressing table (key -> dense tile id) that is rebuilt every substep; ids
// are handed out in insertion order, tiles that own particles first, halo-only tiles after.
// The layer id of a collider-bits pattern is 0 for "no collider near" and slot+1 in the small
// per-substep set of distinct patterns otherwise.
constexpr int BLOCK_BITS = 17;
constexpr int BLOCK_BIAS = 1 << 16;
constexpr int LAYER_BITS = 13;
constexpr int LAYER_SLOTS = 4096;     // distinct collider-bit patterns per substep (open addressing)
constexpr int SIB_MAX = 12;           // compatible sibling layers of one block a G2P tile load can sum
constexpr unsigned long long TILE_EMPTY = ~0ull;
constexpr uint32_t TILE_PENDING = 0xffffffffu;

__host__ __device__ inline unsigned long long tile_key_pack(int bx, int by, int bz, uint32_t layer) {
  return ((((unsigned long long)(uint32_t)(bx + BLOCK_BIAS) << BLOCK_BITS | (unsigned long long)(uint32_t)(by + BLOCK_BIAS)) << BLOCK_BITS |
           (unsigned long long)(uint32_t)(bz + BLOCK_BIAS))
          << LAYER_BITS) |
         layer;
}
__host__ __device__ inline void tile_key_unpack(unsigned long long k, int& bx, int& by, int& bz, uint32_t& layer) {
  layer = (uint32_t)(k & ((1u << LAYER_BITS) - 1));
  k >>= LAYER_BITS;
  bz = (int)(k & ((1u << BLOCK_BITS) - 1)) - BLOCK_BIAS;
  k >>= BLOCK_BITS;
  by = (int)(k & ((1u << BLOCK_BITS) - 1)) - BLOCK_BIAS;
  k >>= BLOCK_BITS;
  bx = (int)k - BLOCK_BIAS;
}
// key of the neighbour tile at block offset d (bit0 = +x, bit1 = +y, bit2 = +z), same layer
__host__ __device__ inline unsigned long long tile_key_offset(unsigned long long k, int d) {
  if (d & 1) k += 1ull << (2 * BLOCK_BITS + LAYER_BITS);
  if (d & 2) k += 1ull << (BLOCK_BITS + LAYER_BITS);
  if (d & 4) k += 1ull << LAYER_BITS;
  return k;
}

struct TileTable {
  ulonglong2* slots;          // [mask + 1] {key (TILE_EMPTY when free), tile id (TILE_PENDING until published)}: one 16-byte load per probe
  uint32_t mask;
  unsigned long long* tile_key;  // [tile_cap] key of tile id
  uint32_t* tile_slot;           // [tile_cap] table slot of tile id (the next substep clears exactly the used slots)
  uint32_t tile_cap;
  uint32_t murmur;               // hash the table with murmur3 of the ordered-u32 block coordinates, seed = layer (the reference's node key) instead of the 64-bit mixer
};

// ---- per-substep scalars that kernels read from HBM (so launches do not depend on host values)
constexpr uint32_t ST_KEY_RANGE = 0x80000000u;   // a live particle lies outside the +-2^18-cell key range
constexpr uint32_t ST_TILE_OVERFLOW = 0x40000000u;  // tile capacity exceeded: the host grows it and redoes the binning
constexpr uint32_t ST_COMM_TIMEOUT = 0x20000000u;   // slab ranks: a neighbour's message did not arrive
constexpr uint32_t ST_COMM_OVERFLOW = 0x10000000u;  // slab ranks: a mailbox or the particle buffer ran out of room
constexpr uint32_t ST_ZERO_DT = 0x08000000u;        // adaptive steps: the allowed time step became 0 inside a substep (the rest of it must not run)
constexpr uint32_t ST_ABORT_MASK = ST_KEY_RANGE | ST_TILE_OVERFLOW | ST_COMM_TIMEOUT | ST_COMM_OVERFLOW | ST_ZERO_DT;
constexpr int WORK_CLASSES = 16;
struct StepScalars {
  uint32_t n;          // resident particles (incl. tombstoned)
  uint32_t n_live;     // particles with a bin (not tombstoned)
  uint32_t n_tomb;
  uint32_t tomb_cursor; // k_invert_zero: next free row among the tombstoned (behind the live rows)
  uint32_t n_tiles;    // active (block, layer) grid tiles
  uint32_t n_ptiles;   // tiles that own particles = ids [0, n_ptiles): the work list of P2G / G2P
  uint32_t n_layers;   // distinct non-zero collider-bit patterns
  uint32_t n_tiles_zeroed;  // tiles cleared by k_invert_zero (tiles created later by a halo message are stored, not added)
  uint32_t status;     // SVB_* simulation-level bits | ST_*
  uint32_t work_counter[4];  // tile claims: [0] P2G, [1] G2P (all tiles, or a slab rank's boundary tiles), [2] / [3] the same for its interior tiles
  uint32_t n_work[2];        // slab ranks: particle tiles in the boundary / interior work list (k_offsets)
  uint32_t n_class[WORK_CLASSES];   // particle tiles by work class (k_offsets): boundary tiles in classes [0, 8), the others in [8, 16), inside
                             // each half by falling particle count (class = 7 - min(7, count / 128)): P2G / G2P claim the heavy tiles first
  uint32_t boundary_done[2]; // slab ranks: boundary tiles P2G / G2P have finished (the concurrent exchange senders wait for n_work[0])
  uint32_t bin_blocks_done;  // k_bin blocks finished: the last one publishes n_ptiles
  uint32_t n_candidates;     // particles whose BVH leaf holds a few triangles within reach (k_collide_query -> k_collide_cand, one thread each)
  uint32_t n_candidates_big; // ... whose leaf holds a long triangle run (back of the candidate array, four lanes each)
  // adaptive time step reductions (f32::total_cmp keys)
  int32_t min_sound_key, min_isolated_key, max_velocity_key, min_deformation_key;
  uint32_t live_count;
  // ---- the run's sticky words: three ADJACENT uint32 (the host clears them with one 12-byte memset at the start of an svb_advance call)
  uint32_t sticky;     // error / stop bits that end the run (SVB_PARTICLE_CLOSE_TO_INVERTED, ST_STOP_*), as carried INTO this substep: constant
                       // while the substep runs, so every CTA of every kernel takes the same "run / no-op" decision
  uint32_t sticky_new; // ... raised BY this substep (a FAILED particle in G2P / advance, a neighbour rank's error word, the device clock
                       // reaching its target): folded into the next substep's `sticky`, which makes that substep and all later ones no-ops
  uint32_t accum;      // simulation-level bits (SVB_TABLE_*) of the EARLIER substeps of this svb_advance call: the status word of every
                       // finished substep is folded in here, so a bit set in any substep's back half reaches the host at the end of the call
  uint32_t reset_done; // blocks of the set-reset that have finished (the last one re-initialises these scalars)
};
// abort bits that stay set once raised: every later substep of the call is a no-op and the host reports a fatal error.  (A tile
// overflow is not carried: the host grows the tables and redoes the binning.)
constexpr uint32_t ST_CARRY_MASK = ST_KEY_RANGE | ST_COMM_TIMEOUT | ST_COMM_OVERFLOW | ST_ZERO_DT;
#define SVB_ABORTED(S) (((S)->status & ST_ABORT_MASK) || (S)->sticky)

// "stop" bits that travel in the upper half of sticky / sticky_new (never reported as simulation status): they turn every later substep
// of the call into a no-op exactly like a FAILED particle does, which is how a device-side clock ends a run the host has queued ahead
constexpr uint32_t ST_STOP_DONE = 0x10000u;      // time >= target_time: produce_next_state's loop is over (cpu_state.rs:166)
constexpr uint32_t ST_STOP_FRAME = 0x20000u;     // floor(time * fps) != loaded frame (xpu/src/frame_input.rs:266-278: WrongFrameLoaded)
constexpr uint32_t ST_STOP_ZERO_DT = 0x40000u;   // allowed_time_step() == 0 (cpu_state.rs:170-172: ZeroTimeStep)

// Adaptive time stepping on the device (cpu/src/adaptive_time_step_state.rs:13-64, cpu/src/phase/limit_time_step.rs:25-33,187-195,
// cpu/src/cpu_state.rs:166-190): the clock, the four limits and the <= 11-entry history live in HBM; kernels read dt from here, three
// one-thread kernels (k_dt_open / k_dt_integrate / k_dt_tail) play AdaptiveTimeStepState, and the host only reads a lagged copy to
// learn when the run is over.  Fixed-dt runs keep the clock on the host and pass dt as a launch argument.
struct DtState {
  double time, target, fps;
  unsigned long long frame;
  float max_dt;
  float allowed;        // allowed_time_step() as of the last push (what the next phase reads)
  float dt_force;       // the step Collide / ExternalForce of the current substep see (before LimitTimeStepBeforeForce's push)
  float h;
  uint32_t has;         // bit 0 velocity, 1 deformation, 2 isolated, 3 sound: which Option<f32> limits are Some
  float by_velocity, by_deformation, by_isolated, by_sound;
  uint32_t prior_len;
  float prior[12];
  uint32_t substeps;    // completed by this svb_advance call
  uint32_t stop;        // ST_STOP_* raised so far (mirror of what went into sticky_new)
  float factor_b, g[3]; // frame factor and interpolated gravity of the CURRENT substep (interpolate_input.rs:25-33)
  float ga[3], gb[3];
  // LimitTimeStepBeforeForce of the NEXT substep, reduced by the kernel that already holds the advanced F (k_advance) or by k_limit_force
  int32_t next_min_sound_key, next_min_isolated_key;
  uint32_t next_live;
  uint32_t pad;
};

struct MeshDev {
  uint32_t n_vertices, n_triangles, n_colliders;
  const uint32_t* tri;          // 3 per triangle
  const uint32_t* opp;          // 3 per triangle
  const uint32_t* tri_collider;
  const uint32_t* fan_offsets;
  const uint32_t* fan_tris;
  const float* va;              // keyframe a vertex positions (3 per vertex)
  const float* vb;              // keyframe b (== va when there is no b)
  const float* vvel;            // vertex velocities
  const float* fric_a; const float* fric_b;
  const float* damp_a; const float* damp_b;
  // interpolated per substep
  float* vpos; float* vnormal; float* tnormal; float* tfric; float* tdamp;
  float* tbox;                  // 8 per triangle (two float4: min, max) bounding box of the interpolated triangle, a cheap reject before the exact distance
  // BVH
  int32_t bvh_level;
  int32_t bvh_nodes;
  const int32_t* node_min; const int32_t* node_max; const int32_t* node_first; const int32_t* node_count;
  const int32_t* children; const uint32_t* tri_indices;
};

struct SimConsts {
  float h;
  float leaf_size;
  float accept_distance, forget_distance;
  float domain_min[3], domain_max[3];
};

}  // namespace svb
