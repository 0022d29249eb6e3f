// svb_device.cuh — device-side data layout of the B200 MPM substep.
//
// Particles: struct of arrays in HBM, every field a contiguous run of `cap` 4-byte words, fields
// back to back in one allocation (field f of particle i = base[f*cap + i]).  Two such buffers
// ping-pong across the physical re-bin (gather by the radix-sorted index).
// Grid: sparse set of 4x4x4-node blocks, one dense 64-node float4 (px,py,pz,m) tile per active
// (block, collider-bits layer), addressed through the sorted list of active bin keys.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "svb_math.cuh"

namespace svb {

// ---- particle fields (4-byte words)
enum Field : int {
  PX = 0,    // position            (3)
  PV = 3,    // velocity            (3)
  PC = 6,    // velocity gradient C (9, column-major)
  PF = 15,   // position gradient F (9, column-major)
  PMASS = 24,
  PVOL = 25,   // initial volume
  PP0 = 26,    // mu | bulk modulus
  PP1 = 27,    // lambda | exponent
  PALPHA = 28, // sand alpha
  PVD = 29,    // viscosity dynamic
  PVB = 30,    // viscosity bulk
  PFLAGS = 31, // u32
  PBITS = 32,  // u32 collider bits
  PORIG = 33,  // u32 original index (the reference's sort_map)
  NFIELDS = 34,
};

struct ParticleBuf {
  uint32_t* base;
  size_t cap;
  __host__ __device__ float* f(int field) const { return reinterpret_cast<float*>(base) + (size_t)field * cap; }
  __host__ __device__ uint32_t* u(int field) const { return base + (size_t)field * cap; }
};

// ---- bin key layout, recomputed every substep from the live bounding box (device resident)
//   key = tomb | bx | by | bz | layer | cell(6)     (most significant first)
// bx/by/bz: block coordinates (node >> 2) relative to `block_min`; layer: rank of the particle's
// collider bits among the distinct values present this substep; cell: (cx<<4 | cy<<2 | cz) of the
// base node inside its block.
struct BinLayout {
  int32_t block_min[3];
  int32_t nb[3];        // bits per block axis
  int32_t nl;           // bits for the layer rank
  int32_t total_bits;   // nb[0]+nb[1]+nb[2]+nl+6 ; the tombstone bit is bit `total_bits`
  int32_t n_layers;
  int32_t cell_min[3], cell_max[3];  // live bounding box of base nodes (debug / multi-GPU)
};
constexpr int LAYER_CAP = 4096;       // distinct collider-bit patterns per substep
constexpr int LAYER_SLOTS = 8192;     // open-addressing set backing the rank table

// ---- per-substep scalars that kernels read from HBM (so launches do not depend on host values)
struct StepScalars {
  float dt_force;      // dt seen by collide / external force
  float dt_scatter;    // dt seen by P2G
  float dt_advance;    // dt seen by advance
  float factor_b;      // keyframe interpolation factor
  float gravity[3];    // interpolated
  uint32_t n;          // resident particles (incl. tombstoned)
  uint32_t n_live;     // particles with a bin (not tombstoned)
  uint32_t n_groups;   // (block, layer) runs of particles
  uint32_t n_cand;
  uint32_t n_active;   // active (block, layer) grid tiles
  uint32_t status;     // SVB_* simulation-level bits
  uint32_t work_counter[4];
  int32_t bbox_min[3], bbox_max[3];  // atomics target for the next layout
  // adaptive time step reductions (f32::total_cmp keys)
  int32_t min_sound_key, min_isolated_key, max_velocity_key, min_deformation_key;
  uint32_t live_count;
  uint32_t layer_count;
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
