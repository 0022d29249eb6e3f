// svb_kernels.cuh — hand-written CUDA kernels (sm_100a) of the MPM substep.
//
// Phase map to the reference's CPU back end (/root/reference/rust/crates/cpu/src/phase/*.rs):
//   k_mesh_*            interpolate_input.rs:18-107      collider mesh lerp, face / vertex normals
//   k_collide_*         collide.rs:21-206                 BVH query, closest triangle per collider, response (compacted candidate lists)
//   k_begin, k_bin      external_force.rs:17-50 + sort.rs:29-33 (cell of floor(x/h - 1/2)) + update_grid_nodes.rs:31-156
//                       (tile activation): single-pass counting sort on (tile, cell)
//   k_offsets           update_grid_nodes.rs:102-108      cell offsets, tile runs, neighbour tiles reached by each tile's particles
//   k_invert_zero       sort.rs:91-101                    inverse map of the binned order (the state is permuted by the G2P write) + grid clear
//   k_p2g               scatter_momentum.rs:22-93         particle -> grid, stress once per particle
//   k_g2p               meld_grid.rs:16-69 + collect_velocity.rs:19-75 + advance_particles.rs:17-93
//                       + cull_particles.rs:17-41 (+ limit_time_step.rs:187-223 reductions)
//   k_limit_force       limit_time_step.rs:25-182
//   k_halo_*2, k_migrate_*, k_note_outside, k_column_histogram     multi-GPU slabs (no reference counterpart)
#pragma once
#include "svb_device.cuh"

namespace svb {

#define SVB_FULL 0xffffffffu

// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may begin while the
// previous kernel of the stream is still draining (its blocks are scheduled onto the SMs that kernel's tail has already left); it
// must not touch anything that kernel wrote before this returns.  A no-op under an ordinary launch.
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// nanoseconds the exchange kernels of this device spent waiting, by kind (svb_exchange_waits): [0] halo sender for P2G's boundary tiles,
// [1] halo receiver for the neighbours' columns, [2] migration sender for G2P's boundary tiles, [3] migration receiver for the
// neighbours' rows, [4] ... for every rank's error word, [5] time-step reductions of the other ranks
__device__ unsigned long long g_wait_ns[8];
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
struct WaitClock {   // block 0 accounts for the whole kernel: its waiting thread sees the same counters as every other block's
  unsigned long long t0;
  int kind;
  __device__ __forceinline__ WaitClock(int k) : t0(blockIdx.x == 0 ? global_timer_ns() : 0ull), kind(k) {}
  __device__ __forceinline__ void stop() { if (blockIdx.x == 0) atomicAdd(&g_wait_ns[kind], global_timer_ns() - t0); }
};
__device__ __forceinline__ int floor_div4(int v) { return v >> 2; }
__device__ __forceinline__ int ceil_log2_u32(uint32_t v) {  // bits needed to represent values 0..v-1
  return v <= 1 ? 0 : 32 - __clz(v - 1);
}

// cpu/src/kernels.rs:46-49 — must be bit-exact: IEEE divide, subtract, floor, no contraction.
__device__ __forceinline__ int base_node(float x, float h) { return (int)floorf(__fsub_rn(__fdiv_rn(x, h), 0.5f)); }

// The same integer without the IEEE division when x * (1/h) - 1/2 is clearly inside a cell: x * (1/h) differs from the correctly
// rounded quotient by a few ulps of |x / h| (two roundings instead of one), so only a value within that distance of an integer
// takes the exact sequence.  Bit-exact by construction; used where the divisions are pure overhead (binning inside G2P).
__device__ __forceinline__ int base_node_fast(float x, float h, float inv_h) {
  const float t = fmaf(x, inv_h, -0.5f);
  const float fl = floorf(t);
  const float f = t - fl;
  const float margin = fmaf(fabsf(t), 4e-7f, 1e-6f);
  if (f > margin && f < 1.f - margin) return (int)fl;
  return base_node(x, h);
}

// ------------------------------------------------------------------------------------------------
// wire <-> particle quads
// (all words as u32: a flag / bit word must not pass through a float register as a signalling NaN pattern would still be preserved
// by LDG / STG, but typed loads keep this obvious)
template <int K>
__global__ void k_wire_to_soa(const uint32_t* __restrict__ src, ParticleBuf P, int field, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
#pragma unroll
  for (int k = 0; k < K; ++k) P.u(field + k)[i] = src[(size_t)i * K + k];
}
// out[(original index)*K + k] = word[field + k][i]   (to_io_state, cpu/src/cpu_state.rs:84-93)
template <int K>
__global__ void k_soa_to_wire(ParticleBuf P, int field, uint32_t* __restrict__ dst, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t o = P.u(PORIG)[i];
#pragma unroll
  for (int k = 0; k < K; ++k) dst[o * K + k] = P.u(field + k)[i];
}
// ... in row order (no scatter by original index)
__global__ void k_soa_to_wire_plain(ParticleBuf P, int field, uint32_t* __restrict__ dst, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = P.u(field)[i];
}
// a plain per-row array (the elastic energies) into original order
__global__ void k_array_to_wire(const float* __restrict__ src, ParticleBuf P, float* __restrict__ dst, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[P.u(PORIG)[i]] = src[i];
}
__global__ void k_iota_orig(ParticleBuf P, uint32_t n, uint32_t offset) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) P.u(PORIG)[i] = i + offset;
}

// ------------------------------------------------------------------------------------------------
// collider mesh interpolation (interpolate_input.rs:36-96)
__global__ void k_mesh_lerp(MeshDev M, float factor_b, const DtState* __restrict__ D) {
  if (D) factor_b = D->factor_b;   // adaptive steps: the clock lives on the device
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const float fa = 1.f - factor_b;
  if (i < M.n_vertices * 3) M.vpos[i] = fa * M.va[i] + factor_b * M.vb[i];
  if (i < M.n_triangles) {
    M.tfric[i] = fa * M.fric_a[i] + factor_b * M.fric_b[i];
    M.tdamp[i] = fa * M.damp_a[i] + factor_b * M.damp_b[i];
  }
}
__device__ __forceinline__ V3 ld3(const float* p, uint32_t i) { return V3{p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }
__global__ void k_mesh_tri_normals(MeshDev M) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M.n_triangles) return;
  const V3 a = ld3(M.vpos, M.tri[3 * t]), b = ld3(M.vpos, M.tri[3 * t + 1]), c = ld3(M.vpos, M.tri[3 * t + 2]);
  const V3 n = normalize_or_zero(cross(b - a, c - a), SVB_NORMALIZATION_EPS);
  M.tnormal[3 * t] = n.x; M.tnormal[3 * t + 1] = n.y; M.tnormal[3 * t + 2] = n.z;
  float4* box = reinterpret_cast<float4*>(M.tbox) + 2 * t;   // (min, -) and (max, -): two 16-byte loads per test
  box[0] = make_float4(fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z)), 0.f);
  box[1] = make_float4(fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z)), 0.f);
}
// true when the triangle's bounding box is farther from p than `reach`: its exact distance is then >= reach as well, so
// skipping it cannot change which triangle is closest within the forget distance (reach carries a 0.1 % rounding margin)
__device__ __forceinline__ bool triangle_out_of_reach(const MeshDev& M, uint32_t t, V3 p, float reach) {
  const float4 lo = __ldg(reinterpret_cast<const float4*>(M.tbox) + 2 * t), hi = __ldg(reinterpret_cast<const float4*>(M.tbox) + 2 * t + 1);
  const float dx = fmaxf(fmaxf(lo.x - p.x, p.x - hi.x), 0.f);
  const float dy = fmaxf(fmaxf(lo.y - p.y, p.y - hi.y), 0.f);
  const float dz = fmaxf(fmaxf(lo.z - p.z, p.z - hi.z), 0.f);
  return dx * dx + dy * dy + dz * dz > reach * reach;
}
__global__ void k_mesh_vertex_normals(MeshDev M) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= M.n_vertices) return;
  V3 sum = V3{0.f, 0.f, 0.f};
  const V3 p = ld3(M.vpos, v);
  for (uint32_t q = M.fan_offsets[v]; q < M.fan_offsets[v + 1]; ++q) {
    const uint32_t t = M.fan_tris[q];
    uint32_t others[2] = {0, 0};
    int n = 0;
    for (int k = 0; k < 3; ++k) {
      const uint32_t w = M.tri[3 * t + k];
      if (w != v && n < 2) others[n++] = w;
    }
    const float ang = angle_between(ld3(M.vpos, others[0]) - p, ld3(M.vpos, others[1]) - p);
    sum = sum + ang * ld3(M.tnormal, t);
  }
  const V3 r = normalize_or_zero(sum, SVB_NORMALIZATION_EPS);
  M.vnormal[3 * v] = r.x; M.vnormal[3 * v + 1] = r.y; M.vnormal[3 * v + 2] = r.z;
}

// ------------------------------------------------------------------------------------------------
// point / triangle distance (mesh_util/src/mesh.rs:234-309)
struct DistResult {
  float distance;
  V3 to_p, normal;
};
__device__ __forceinline__ DistResult segment_result(V3 p, V3 start, V3 end, V3 n_start, V3 n_seg, V3 n_end) {
  const V3 seg = end - start;
  const float along = dot(p - start, seg) / dot(seg, seg);
  if (along < 0.f) return DistResult{norm(p - start), p - start, n_start};
  if (along < 1.f) {
    const V3 d = p - start - seg * along;
    return DistResult{norm(d), d, n_seg};
  }
  return DistResult{norm(p - end), p - end, n_end};
}
__device__ __forceinline__ float segment_distance(V3 p, V3 start, V3 end) {
  const V3 seg = end - start;
  const float along = dot(p - start, seg) / dot(seg, seg);
  if (along < 0.f) return norm(p - start);
  if (along < 1.f) return norm(p - start - seg * along);
  return norm(p - end);
}
__device__ __forceinline__ float triangle_distance(V3 p, V3 a, V3 b, V3 c, V3 n) {
  const V3 ab = a - b, bc = b - c, ca = c - a;
  const bool sa = dot(n, cross(bc, c - p)) > 0.f;
  const bool sb = dot(n, cross(ca, a - p)) > 0.f;
  const bool sc = dot(n, cross(ab, b - p)) > 0.f;
  if (sa && sb && sc) return fabsf(dot(p - a, n));
  float d = 3.402823466e+38f;
  if (!sa) d = fminf(d, segment_distance(p, b, c));
  if (!sb) d = fminf(d, segment_distance(p, c, a));
  if (!sc) d = fminf(d, segment_distance(p, a, b));
  return d;
}

// BVH point query (mesh_util/src/bounding_volume_hierarchy.rs:177-217): returns the leaf's triangle
// run or count 0.
__device__ __forceinline__ void bvh_query(const MeshDev& M, int qx, int qy, int qz, int& first, int& count) {
  first = 0;
  count = 0;
  if (M.bvh_nodes == 0) return;
  int cur = 0;
  if (qx < M.node_min[0] || qy < M.node_min[1] || qz < M.node_min[2] || qx > M.node_max[0] || qy > M.node_max[1] || qz > M.node_max[2]) return;
  if (M.node_count[0] >= 0) { first = M.node_first[0]; count = M.node_count[0]; return; }
  const uint32_t ux = (uint32_t)(qx - M.node_min[0]), uy = (uint32_t)(qy - M.node_min[1]), uz = (uint32_t)(qz - M.node_min[2]);
  for (int lvl = M.bvh_level - 1; lvl >= 0; --lvl) {
    const uint32_t child = (((ux >> (2 * lvl)) & 3u) << 4) | (((uy >> (2 * lvl)) & 3u) << 2) | ((uz >> (2 * lvl)) & 3u);
    const int ci = M.children[M.node_first[cur] + child];
    if (ci < 0) return;
    if (M.node_count[ci] >= 0) { first = M.node_first[ci]; count = M.node_count[ci]; return; }
    cur = ci;
  }
}

// collide.rs:128-204 for one particle, given the closest triangle per collider (0xffffffff = none within the forget
// distance): feature classification, side bits, friction / damping / push-out.  Returns the new collider bits; edits `vel`.
// NC = number of colliders handled in registers (the scene has <= NC); the reference loops over all 16 slots and forgets the
// colliders it found no triangle for — for the slots >= NC that is the final mask.
template <int NC>
__device__ __forceinline__ uint32_t collide_respond(const MeshDev& M, const SimConsts& K, float dt, V3 p, V3& vel, uint32_t bits, const uint32_t* closest) {
#pragma unroll
  for (unsigned collider = 0; collider < (unsigned)NC; ++collider) {
    const uint32_t ct = closest[collider];
    if (ct == 0xffffffffu) { bits = bits_set(bits, collider, -1); continue; }
    const uint32_t ia = M.tri[3 * ct], ib = M.tri[3 * ct + 1], ic = M.tri[3 * ct + 2];
    const uint32_t oab = M.opp[3 * ct], obc = M.opp[3 * ct + 1], oca = M.opp[3 * ct + 2];
    const V3 n = ld3(M.tnormal, ct);
    const V3 a = ld3(M.vpos, ia), b = ld3(M.vpos, ib), c = ld3(M.vpos, ic);
    const V3 zero = V3{0.f, 0.f, 0.f};
    const V3 ab = a - b, bc = b - c, ca = c - a;
    const float area2 = dot(n, cross(ca, ab));
    const float a_bary = dot(n, cross(bc, c - p)) / area2;
    const float b_bary = dot(n, cross(ca, a - p)) / area2;
    const float c_bary = dot(n, cross(ab, b - p)) / area2;
    DistResult res;
    if (a_bary > 0.f && b_bary > 0.f && c_bary > 0.f) {
      const float s = dot(p - a, n);
      res = DistResult{fabsf(s), n * s, n};
    } else {
      const V3 a_n = ld3(M.vnormal, ia), b_n = ld3(M.vnormal, ib), c_n = ld3(M.vnormal, ic);
      const V3 ab_n = oab != 0xffffffffu ? n + ld3(M.tnormal, oab) : zero;
      const V3 bc_n = obc != 0xffffffffu ? n + ld3(M.tnormal, obc) : zero;
      const V3 ca_n = oca != 0xffffffffu ? n + ld3(M.tnormal, oca) : zero;
      res = segment_result(p, a, b, a_n, ab_n, b_n);
      const DistResult r1 = segment_result(p, b, c, b_n, bc_n, c_n);
      const DistResult r2 = segment_result(p, c, a, c_n, ca_n, a_n);
      if (total_key(r1.distance) < total_key(res.distance)) res = r1;  // min_by keeps the first minimum
      if (total_key(r2.distance) < total_key(res.distance)) res = r2;
    }
    if (is_zero(res.normal)) { bits = bits_set(bits, collider, -1); continue; }
    const bool new_side = 0.f <= dot(res.to_p, res.normal);
    const int prior = bits_get(bits, collider);
    if (prior < 0) {
      if (res.distance < K.accept_distance) bits = bits_set(bits, collider, new_side ? 1 : 0);
      continue;
    }
    if ((prior == 1) == new_side) continue;
    if (res.distance > SVB_NORMALIZATION_EPS) {
      const V3 cv = ld3(M.vvel, ia) * a_bary + ld3(M.vvel, ib) * b_bary + ld3(M.vvel, ic) * c_bary;
      const V3 rel = vel - cv;
      const V3 cn = res.to_p / res.distance;
      const V3 nv = cn * dot(rel, cn);
      const V3 tv = rel - nv;
      const float tn = norm(tv);
      if (tn > SVB_NORMALIZATION_EPS) {
        const V3 tangent = tv / tn;
        vel = vel - tangent * fminf(M.tfric[ct] * res.distance / dt, tn);
      }
      vel = vel - fminf(M.tdamp[ct], 1.f) * nv;
    }
    vel = vel - res.to_p / dt;
  }
  if (NC < 16) bits &= 0x00010001u * ((1u << NC) - 1u);   // slots of colliders that do not exist: forgotten (collide.rs:163-166 with no triangle)
  return bits;
}

// Collide (collide.rs:21-206) in two passes over the particles, so that the triangle loops of the particles near a collider do
// not stall the warps of the many that are not:
//   k_collide_query   one thread per particle: BVH point query of its leaf cell; an empty leaf clears the collider bits
//                     (collide.rs:57-61), a leaf with triangles puts the particle on the candidate list (warp-aggregated append:
//                     32 consecutive rows of the binned state stay together);
//   k_collide_cand    one thread per candidate: the reference's sequential scan of the leaf's triangle run for the closest triangle
//                     per collider (collide.rs:63-88; strict `<` keeps the first minimum, so the bits stay bit-exact), then the
//                     response.  Consecutive candidates are consecutive rows of the binned state — neighbours in space, mostly in
//                     the same BVH leaf — so the lanes of a warp walk the SAME triangle run: uniform trip counts, broadcast loads.
//                     (Round 1 gave a warp to every candidate of a long run and lane 0 the response: 16 of 32 lanes idle, 550
//                     warp instructions per candidate, 141 us for 1 M sand particles resting on a torus.)
//                     NC = colliders tracked in registers (instantiated for 1, 2, 4, 16).
constexpr int COLLIDE_SMALL_MAX = 24;
__global__ void __launch_bounds__(256) k_collide_query(ParticleBuf P, StepScalars* S, SimConsts K, MeshDev M, uint32_t* __restrict__ candidates, uint32_t cap, uint32_t n) {
  if (S->sticky) return;
  n = min(n, S->n);
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  int kind = 0;   // 1: a short triangle run (front of the list, one thread each), 2: a long one (back of the list, four lanes each)
  if (i < n) {
    const float4 pq0 = P.q(0)[i];   // position, flags
    const uint32_t flags = __float_as_uint(pq0.w);
    const float x0 = pq0.x, x1 = pq0.y, x2 = pq0.z;
    if (!(flags & (F_TOMBSTONED | F_GONE))) {
      int first, count;
      bvh_query(M, (int)floorf(x0 / K.leaf_size), (int)floorf(x1 / K.leaf_size), (int)floorf(x2 / K.leaf_size), first, count);
      if (count > COLLIDE_SMALL_MAX) kind = 2;
      else {
        // a leaf with few triangles (a coarse mesh: the whole box collider of a dam break is ONE leaf) — look for any triangle
        // whose bounding box is within reach; with none, every collider ends "not near": the same bits as an empty leaf
        const V3 p = V3{x0, x1, x2};
        const float reach = K.forget_distance * 1.001f;
        for (int r = 0; r < count && kind == 0; ++r)
          if (!triangle_out_of_reach(M, M.tri_indices[first + r], p, reach)) kind = 1;
      }
      if (kind == 0) P.u(PBITS)[i] = 0u;
    }
  }
  const uint32_t lane = threadIdx.x & 31;
  const unsigned ms = __ballot_sync(SVB_FULL, kind == 1), mb = __ballot_sync(SVB_FULL, kind == 2);
  uint32_t base_s = 0, base_b = 0;
  if (lane == 0) {
    if (ms) base_s = atomicAdd(&S->n_candidates, (uint32_t)__popc(ms));
    if (mb) base_b = atomicAdd(&S->n_candidates_big, (uint32_t)__popc(mb));
  }
  base_s = __shfl_sync(SVB_FULL, base_s, 0);
  base_b = __shfl_sync(SVB_FULL, base_b, 0);
  const unsigned below = (1u << lane) - 1u;
  if (kind == 1) candidates[base_s + __popc(ms & below)] = i;                    // the two lists cannot meet: together they hold <= n <= cap entries
  if (kind == 2) candidates[cap - 1 - (base_b + __popc(mb & below))] = i;
}
// LANES lanes (1 or 4) per candidate share the leaf's triangle run (entries sub, sub + LANES, ...).  The scan is a chain of dependent
// loads (run entry -> bounding box -> vertices): a particle resting on a finely meshed collider sees 40-100 triangles, and one thread
// per candidate left the SMs at 14 % active warps (105 us for 1 M sand particles on a torus; 62 us with four lanes) — while the
// candidates of a coarse mesh (the 12-triangle box of a dam break is ONE leaf) have a handful of triangles and are dominated by the
// response, which only one lane of a group computes (four lanes there: 113 -> 131 us).  Hence two lists.  Each lane keeps, per
// collider, the closest triangle of ITS entries as the key (distance bits << 32 | position in the run); distances are >= 0, so the
// unsigned order of the keys is the numeric order and ties go to the earlier entry: the minimum over the lanes is the FIRST minimum
// of the reference's sequential scan (collide.rs:82).  BIG: the list of long runs (back of the candidate array).
template <int NC, int LANES, bool BIG>
__global__ void __launch_bounds__(128) k_collide_cand(ParticleBuf P, const StepScalars* __restrict__ S, SimConsts K, MeshDev M, const uint32_t* __restrict__ candidates, uint32_t cap, float dt,
                                                      const DtState* __restrict__ D) {
  if (S->sticky) return;
  if (D) dt = D->dt_force;
  const uint32_t n_cand = BIG ? S->n_candidates_big : S->n_candidates;
  const float reach = K.forget_distance * 1.001f;
  const uint32_t lane = threadIdx.x & 31, sub = lane & (LANES - 1);
  const uint32_t groups = (gridDim.x * blockDim.x) / LANES;
  constexpr uint32_t PER_WARP = 32 / LANES;
  for (uint32_t q0 = (blockIdx.x * blockDim.x + threadIdx.x) / LANES; q0 < ((n_cand + PER_WARP - 1) / PER_WARP) * PER_WARP; q0 += groups) {   // warp-uniform trip count
    const bool have = q0 < n_cand;
    const uint32_t i = have ? candidates[BIG ? cap - 1 - q0 : q0] : 0u;
    V3 p = V3{0.f, 0.f, 0.f};
    int first = 0, count = 0;
    if (have) {
      const float4 pq0 = P.q(0)[i];
      p = V3{pq0.x, pq0.y, pq0.z};
      bvh_query(M, (int)floorf(p.x / K.leaf_size), (int)floorf(p.y / K.leaf_size), (int)floorf(p.z / K.leaf_size), first, count);
    }
    unsigned long long best[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) best[c] = ~0ull;
    for (int r = (int)sub; r < count; r += LANES) {
      const uint32_t t = M.tri_indices[first + r];
      if (triangle_out_of_reach(M, t, p, reach)) continue;
      const V3 n = ld3(M.tnormal, t);
      if (is_zero(n)) continue;
      const float d = triangle_distance(p, ld3(M.vpos, M.tri[3 * t]), ld3(M.vpos, M.tri[3 * t + 1]), ld3(M.vpos, M.tri[3 * t + 2]), n);
      if (!(d < K.forget_distance)) continue;   // (also drops a NaN distance, which the reference's `d < min` never selects)
      const uint32_t c = M.tri_collider[t] & 15u;
      const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (uint32_t)r;
#pragma unroll
      for (int k = 0; k < NC; ++k)
        if ((uint32_t)k == c && key < best[k]) best[k] = key;
    }
    uint32_t closest[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      unsigned long long b = best[c];
#pragma unroll
      for (int o = 1; o < LANES; o <<= 1) {
        const unsigned long long other = __shfl_xor_sync(SVB_FULL, b, o);
        b = other < b ? other : b;
      }
      closest[c] = b != ~0ull ? M.tri_indices[first + (uint32_t)b] : 0xffffffffu;
    }
    if (have && sub == 0) {
      float4 pq5 = P.q(5)[i];   // collider bits, original index, v.x, v.y
      float* vz = &P.f(PV + 2)[i];
      V3 vel = V3{pq5.z, pq5.w, *vz};
      const uint32_t bits = collide_respond<NC>(M, K, dt, p, vel, __float_as_uint(pq5.x), closest);
      pq5.x = __uint_as_float(bits); pq5.z = vel.x; pq5.w = vel.y;
      P.q(5)[i] = pq5;
      *vz = vel.z;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// layer set: distinct non-zero collider-bit patterns of the live particles of this substep.
// layer id = slot + 1 (0 = "no collider near"); slots hold (1<<32 | bits), 0 = free.
__device__ __forceinline__ uint32_t layer_hash(uint32_t b) {
  b ^= b >> 16; b *= 0x7feb352du; b ^= b >> 15; b *= 0x846ca68bu; b ^= b >> 16;
  return b & (LAYER_SLOTS - 1);
}
__device__ __forceinline__ uint32_t layer_find_or_insert(unsigned long long* slots, uint32_t* layer_list, uint32_t bits, StepScalars* S) {
  if (bits == 0u) return 0u;
  const unsigned long long want = (1ull << 32) | bits;
  uint32_t s = layer_hash(bits);
  for (int tries = 0; tries < LAYER_SLOTS; ++tries) {
    unsigned long long cur = *(volatile unsigned long long*)&slots[s];
    if (cur == 0ull) {
      cur = atomicCAS(&slots[s], 0ull, want);
      if (cur == 0ull) {
        const uint32_t at = atomicAdd(&S->n_layers, 1u);
        layer_list[at] = s + 1;
        return s + 1;
      }
    }
    if (cur == want) return s + 1;
    s = (s + 1) & (LAYER_SLOTS - 1);
  }
  atomicOr(&S->status, 1u /*SVB_TABLE_TRIES_EXCEEDED*/);
  return 0u;
}
__device__ __forceinline__ uint32_t layer_bits_of(const unsigned long long* __restrict__ slots, uint32_t layer) {
  return layer == 0u ? 0u : (uint32_t)slots[layer - 1];
}

// ------------------------------------------------------------------------------------------------
// tile table (update_grid_nodes.rs:31-156 without the hash map rebuild: the table is cleared and
// refilled every substep, so there are no stale nodes)
__device__ __forceinline__ uint32_t tile_hash_mix(unsigned long long k, uint32_t mask) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return (uint32_t)k & mask;
}
// option "murmur_table_hash": the table is hashed like the reference hashes its grid nodes (gpu/src/util.rs:79-100,
// build_hash_table_on_cpu :102-118): murmur3_x86_32 of the ordered-u32 coordinates — here of the tile's block, seeded with its layer
__device__ __forceinline__ uint32_t tile_hash(const TileTable& T, unsigned long long k) {
  if (!T.murmur) return tile_hash_mix(k, T.mask);
  int bx, by, bz;
  uint32_t layer;
  tile_key_unpack(k, bx, by, bz, layer);
  return node_id_to_murmur(bx, by, bz, layer) & T.mask;
}
__device__ __forceinline__ ulonglong2 slot_load(const ulonglong2* p) {
  ulonglong2 v;
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
  return v;
}
// returns the tile id, or TILE_PENDING when the tile capacity is exhausted (status bit set)
__device__ __forceinline__ uint32_t tile_find_or_insert(const TileTable& T, unsigned long long key, StepScalars* S) {
  uint32_t s = tile_hash(T, key);
  for (uint32_t tries = 0; tries <= T.mask; ++tries) {
    ulonglong2 cur = slot_load(&T.slots[s]);
    if (cur.x == TILE_EMPTY) {
      cur.x = atomicCAS(&T.slots[s].x, TILE_EMPTY, key);
      if (cur.x == TILE_EMPTY) {
        const uint32_t id = atomicAdd(&S->n_tiles, 1u);
        if (id < T.tile_cap) {
          T.tile_key[id] = key;
          T.tile_slot[id] = s;
          __threadfence();
          *(volatile unsigned long long*)&T.slots[s].y = id;
          return id;
        }
        atomicOr(&S->status, ST_TILE_OVERFLOW);
        *(volatile unsigned long long*)&T.slots[s].y = TILE_PENDING - 1;  // poisoned: waiters stop spinning
        return TILE_PENDING;
      }
      cur.y = ~0ull;
    }
    if (cur.x == key) {
      while ((uint32_t)cur.y == TILE_PENDING) cur.y = *(volatile unsigned long long*)&T.slots[s].y;
      return (uint32_t)cur.y == TILE_PENDING - 1 ? TILE_PENDING : (uint32_t)cur.y;
    }
    s = (s + 1) & T.mask;
  }
  atomicOr(&S->status, ST_TILE_OVERFLOW);
  return TILE_PENDING;
}
// k_bin's variant: returns the TABLE SLOT of the tile (or ~0u when the capacity is exhausted, status bit set).  The particle
// counters of a substep are indexed by slot, so no thread waits for a tile id: the thread that claims an empty slot takes the next
// particle-tile id and records (id -> key, slot) and slot -> id for the kernels that follow, nobody spins on it.  Particle tiles
// count in n_ptiles AND n_tiles, so that n_tiles == n_ptiles when the binning kernel ends without any block having to publish it.
__device__ __forceinline__ uint32_t tile_slot_find_or_insert(const TileTable& T, unsigned long long key, StepScalars* S) {
  uint32_t s = tile_hash(T, key);
  for (uint32_t tries = 0; tries <= T.mask; ++tries) {
    unsigned long long k = *(volatile unsigned long long*)&T.slots[s].x;
    if (k == TILE_EMPTY) {
      k = atomicCAS(&T.slots[s].x, TILE_EMPTY, key);
      if (k == TILE_EMPTY) {
        const uint32_t id = atomicAdd(&S->n_ptiles, 1u);
        atomicAdd(&S->n_tiles, 1u);   // result unused: a fire-and-forget RED
        if (id < T.tile_cap) {
          T.tile_key[id] = key;
          T.tile_slot[id] = s;
          *(volatile unsigned long long*)&T.slots[s].y = id;
          return s;
        }
        atomicOr(&S->status, ST_TILE_OVERFLOW);
        *(volatile unsigned long long*)&T.slots[s].y = TILE_PENDING - 1;
        return ~0u;
      }
    }
    if (k == key) return s;
    s = (s + 1) & T.mask;
  }
  atomicOr(&S->status, ST_TILE_OVERFLOW);
  return ~0u;
}
__device__ __forceinline__ int tile_find(const TileTable& T, unsigned long long key) {
  uint32_t s = tile_hash(T, key);
  for (uint32_t tries = 0; tries <= T.mask; ++tries) {
    const ulonglong2 cur = T.slots[s];
    if (cur.x == key) return (uint32_t)cur.y >= TILE_PENDING - 1 ? -1 : (int)(uint32_t)cur.y;
    if (cur.x == TILE_EMPTY) return -1;
    s = (s + 1) & T.mask;
  }
  return -1;
}

// ------------------------------------------------------------------------------------------------
// binning (sort.rs:29-33 cell keys + update_grid_nodes.rs:31-156 tile activation): a single-pass counting sort on (tile, cell)
struct GoalDev {
  const uint32_t* flags_a; const uint32_t* flags_b;   // original order, may be null
  const float* goal_a; const float* goal_b;           // 3 per particle, original order
};
// A "front set" = the per-substep tile bookkeeping (tile table, per-slot cell counters and touch masks).  Two sets alternate: while
// substep n runs on set n % 2, the fused G2P of substep n already bins the ADVANCED positions into the other set for substep n + 1
// (scenes without a collider mesh; with a mesh the collider bits — hence the layer of a particle's tile — are only known after the
// collide pass of substep n + 1, so those scenes bin at the start of the substep with k_bin).
//
// reset_set: undo exactly what the set's previous use left in it (its tiles' hash slots, touch masks and cell counters) instead of
// memset-ing whole allocations, then — last block done — re-initialise the set's scalars, carrying the run's sticky words over from
// `from` (the scalars of the substep before the one this set will serve).  `S` still describes the set's previous content while the
// blocks work, which is why its re-initialisation waits for the last of them.
__device__ __forceinline__ void reset_set(const StepScalars* __restrict__ from, StepScalars* __restrict__ S, const TileTable& T, uint32_t* __restrict__ cell_count, uint32_t* __restrict__ tile_touch,
                                          uint32_t n, const uint32_t* __restrict__ n_dev, int tables_fresh, uint32_t block, uint32_t n_blocks) {
  // tables_fresh: the host has just (re)allocated and memset the tables, tile_slot holds nothing to undo
  const uint32_t n_tiles = tables_fresh ? 0u : min(S->n_tiles, T.tile_cap), n_ptiles = tables_fresh ? 0u : min(S->n_ptiles, T.tile_cap);
  const uint32_t gtid = block * blockDim.x + threadIdx.x, gsz = n_blocks * blockDim.x;
  for (uint32_t t = gtid; t < n_tiles; t += gsz) {
    const uint32_t slot = T.tile_slot[t];
    T.slots[slot] = make_ulonglong2(TILE_EMPTY, ~0ull);
    tile_touch[slot] = 0u;
  }
  uint4* cc = reinterpret_cast<uint4*>(cell_count);
  for (uint32_t q = gtid; q < n_ptiles * 16u; q += gsz) cc[(size_t)T.tile_slot[q >> 4] * 16u + (q & 15u)] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&S->reset_done, 1u) == n_blocks - 1) {
      StepScalars z{};
      z.n = n_dev ? min(*n_dev, n) : n;   // slab ranks: the row count lives on the device (migration changes it without the host)
      z.min_sound_key = INT32_MAX; z.min_isolated_key = INT32_MAX; z.max_velocity_key = INT32_MIN; z.min_deformation_key = INT32_MAX;
      z.sticky = from->sticky | from->sticky_new;
      z.accum = from->accum | (from->status & 0xffffu);
      z.status = from->status & ST_CARRY_MASK;
      *S = z;
    }
  }
}
// Start of a substep that bins with k_bin: reset the set this substep uses and (mesh scenes) the layer set of the previous substep.
__global__ void __launch_bounds__(256) k_begin(const StepScalars* __restrict__ prev, StepScalars* __restrict__ cur, TileTable T, uint32_t* __restrict__ cell_count, uint32_t* __restrict__ tile_touch,
                                               unsigned long long* __restrict__ layer_slots, uint32_t n, const uint32_t* __restrict__ n_dev, int tables_fresh) {
  // an earlier substep stopped the run (a FAILED particle, the device clock reached its target, an exchange error): this substep is a
  // no-op — leave the tables (the grid of the last real substep is still downloadable) and only hand the stop on
  const uint32_t stop = prev->sticky | prev->sticky_new, carry = prev->status & ST_CARRY_MASK;
  if (stop | carry) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      cur->sticky = stop;
      cur->sticky_new = 0u;
      cur->accum = prev->accum | (prev->status & 0xffffu);
      cur->status = carry;
      cur->bin_blocks_done = 0u;
    }
    return;
  }
  if (prev->n_layers && !tables_fresh)
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < LAYER_SLOTS; q += gridDim.x * blockDim.x) layer_slots[q] = 0ull;
  reset_set(prev, cur, T, cell_count, tile_touch, n, n_dev, tables_fresh, blockIdx.x, gridDim.x);
}

// Binning of one particle, WARP-COLLECTIVE (all 32 lanes call it; `state`: 0 = live, 1 = tombstoned, 2 = gone / no particle):
//   * the particle's tile (block of its base node, layer of its collider bits) is found or created in the tile table,
//     warp-aggregated: lanes with equal keys elect one lane to touch the table;
//   * its cell is counted: cell_count[tile*64 + cell] += 1 (one reduction per distinct cell per warp) — the counting half of a
//     single-pass counting sort on (tile, cell); the RANK inside the cell is handed out later by k_invert_zero from the same
//     counters (turned into offsets by k_offsets), so nothing here waits for the return value of an atomic;
//   * the tile remembers which of its 8 neighbour blocks its particles' stencils reach.
// Writes pcell[row] when `row` is valid.  x must be the position the NEXT P2G will see.
// CACHED (the fused G2P, one CTA per tile): the table slots of the 27 blocks around the CTA's tile are cached in shared memory
// (`cache`, ~0u = not looked up yet; `cb*` = the tile's block), because a particle moves less than a cell per substep: after the
// first particle of a tile has found or created a neighbour's slot no other one probes the table in HBM for it.
struct BinArrays {
  uint32_t* pcell;        // [n] table slot * 64 + cell; 0xffffffff tombstoned, 0xfffffffd gone, 0xfffffffe unbinned
  uint32_t* cell_count;   // [table slots * 64]: indexed by the tile's TABLE SLOT, so counting never waits for a tile id
  uint32_t* tile_touch;   // [table slots] 8-bit mask over neighbour offsets d
  unsigned long long* layer_slots;
  uint32_t* layer_list;
};
template <bool HAS_MESH, bool CACHED>
__device__ __forceinline__ void bin_warp(int state, V3 x, uint32_t bits, bool has_row, uint32_t row, const SimConsts& K, const TileTable& T, const BinArrays& B, StepScalars* S, uint32_t lane,
                                         volatile uint32_t* cache = nullptr, int cbx = 0, int cby = 0, int cbz = 0) {
  bool live = false;
  unsigned long long key = TILE_EMPTY;
  uint32_t cell = 0, touch = 0;
  int cache_at = -1;
  if (state == 0) {
    const int s0 = base_node(x.x, K.h), s1 = base_node(x.y, K.h), s2 = base_node(x.z, K.h);
    const int b0 = floor_div4(s0), b1 = floor_div4(s1), b2 = floor_div4(s2);
    const int lim = BLOCK_BIAS - 2;
    if (b0 < -lim || b0 > lim || b1 < -lim || b1 > lim || b2 < -lim || b2 > lim) {
      atomicOr(&S->status, ST_KEY_RANGE);
    } else {
      live = true;
      const uint32_t layer = HAS_MESH ? layer_find_or_insert(B.layer_slots, B.layer_list, bits, S) : 0u;
      key = tile_key_pack(b0, b1, b2, layer);
      cell = ((uint32_t)(s0 & 3) << 4) | ((uint32_t)(s1 & 3) << 2) | (uint32_t)(s2 & 3);
      // neighbour offsets reached by the 3-node stencil: +1 on an axis iff the in-block coordinate >= 2
      touch = 1u;
      if (s0 & 2) touch |= touch << 1;
      if (s1 & 2) touch |= touch << 2;
      if (s2 & 2) touch |= touch << 4;
      if (CACHED) {
        const int d0 = b0 - cbx + 1, d1 = b1 - cby + 1, d2 = b2 - cbz + 1;
        if ((unsigned)d0 < 3u && (unsigned)d1 < 3u && (unsigned)d2 < 3u) cache_at = (d0 * 3 + d1) * 3 + d2;
      }
    }
  }
  // ---- tile: one table access per distinct key in the warp (none when the CTA's cache knows the slot)
  const unsigned peers = __match_any_sync(SVB_FULL, key);
  const int leader = __ffs(peers) - 1;
  uint32_t tile = ~0u;   // the tile's table slot
  if (live && (int)lane == leader) {
    if (CACHED && cache_at >= 0) tile = cache[cache_at];
    if (tile == ~0u) {
      tile = tile_slot_find_or_insert(T, key, S);
      if (CACHED && cache_at >= 0 && tile != ~0u) cache[cache_at] = tile;
    }
  }
  tile = __shfl_sync(SVB_FULL, tile, leader);
  const uint32_t tm = __reduce_or_sync(peers, touch);
  if (live && (int)lane == leader && tile != ~0u) atomicOr(&B.tile_touch[tile], tm);  // result unused: a fire-and-forget RED
  // ---- count the cell: one reduction per distinct (tile, cell) in the warp, result unused (RED)
  const bool binned = live && tile != ~0u;
  const bool tomb = state == 1;
  const uint32_t ci = binned ? tile * 64u + cell : (tomb ? 0xffffffffu : (state == 2 ? 0xfffffffdu : 0xfffffffeu));
  const unsigned cpeers = __match_any_sync(SVB_FULL, ci);
  if ((int)lane == __ffs(cpeers) - 1) {
    if (binned) atomicAdd(&B.cell_count[ci], (uint32_t)__popc(cpeers));
    else if (tomb) atomicAdd(&S->n_tomb, (uint32_t)__popc(cpeers));
  }
  if (has_row) B.pcell[row] = ci;
}

// k_bin — binning as its own pass, one thread per particle in the CURRENT order: the first substep after an upload, every substep
// of a collider scene (after its collide pass), and the redo after a tile-capacity overflow.  (The external force moved into P2G:
// the forced velocity is only ever read there — collect_velocity.rs overwrites v — so it never has to be stored.)
template <bool HAS_MESH>
__global__ void __launch_bounds__(256) k_bin(ParticleBuf P, StepScalars* S, SimConsts K, TileTable T, BinArrays B, uint32_t n) {
  if (S->sticky) return;  // an earlier substep hit a simulation-level error: leave the state as it is (every later kernel no-ops too)
  n = min(n, S->n);       // the launch covers an upper bound of the row count
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31;
  int state = 2;
  V3 x = V3{0.f, 0.f, 0.f};
  uint32_t bits = 0u;
  if (i < n) {
    // every load of the row is issued before the flags are looked at: one memory latency instead of two in the chain
    const float4 pq0 = P.q(0)[i];   // position, flags
    const uint32_t flags = __float_as_uint(pq0.w);
    x = V3{pq0.x, pq0.y, pq0.z};
    bits = HAS_MESH ? P.u(PBITS)[i] : 0u;
    const bool gone = (flags & F_GONE) != 0;   // migrated to a neighbour slab: the row is dropped by this re-bin
    state = gone ? 2 : ((flags & F_TOMBSTONED) ? 1 : 0);
  }
  bin_warp<HAS_MESH, false>(state, x, bits, i < n, i, K, T, B, S, lane);
  if (i == 0) S->bin_blocks_done = 1u;   // "this substep was binned" (a sticky error makes the kernel return at the top instead)
}

// per particle-owning tile (one warp each): exclusive scan of its 64 cell counts (in place), the tile's slot range
// [first, end) in the binned order (S->n_live ends up as the number of binned particles), and the halo: create the neighbour tiles its particles' stencils reach and record the 8 neighbour ids
// (update_grid_nodes.rs:102-108)
// The particle tiles a launch of P2G / G2P works through, claimed one at a time from `cursor`.  Slab ranks order their tiles:
// BOUNDARY tiles first (block columns lo and hi - 1: the only ones that write grid nodes a neighbour rank also writes, read nodes
// that receive a neighbour's sums, or hold particles that can leave the slab), then the INTERIOR tiles.  Every finished boundary tile
// ticks `done`; the sending kernel of the exchange runs CONCURRENTLY on a second stream, waits for `done` to reach the number of
// boundary tiles and ships the halo column / the leavers while the interior tiles are still being worked on — so the neighbour's
// message is long there when the receiving kernel that follows P2G / G2P on the main stream asks for it.
struct WorkList {
  const uint32_t* ids;       // WORK_CLASSES lists of tile ids, list c from c * split
  uint32_t split;            // the tile capacity
  const uint32_t* count;     // n_class[WORK_CLASSES] of the substep's scalars: final before P2G starts
  uint32_t* cursor;
  uint32_t* done;            // boundary work items finished so far (null: nobody waits)
  int tail;                  // G2P: this launch also carries the tombstoned rows over
  // A work item is a tile, or 1 / `parts` of its particle run (parts = 1, 2 or 4: particles [k, k + 1) * 512 / parts of the run, the
  // last part up to the run's end): at 1 M particles a GPU holds only ~3.5 tiles per resident P2G CTA and a tile keeps a CTA busy for
  // ~17 us, so whole-tile items leave SMs idle while the last CTAs finish; every item also pays ~3 us of dependent round trips, so
  // only the LAST `split_last` tiles a launch hands out are cut (its final wave), the others stay whole.
  uint32_t parts, split_last;
};
// Items are handed out class by class: boundary tiles first (the concurrent sender waits for them), and inside both halves the
// heaviest tiles first — longest-processing-time-first list scheduling, so the last items a launch hands out are its lightest and
// the CTAs finish together (a compressed region holds tiles of 700+ particles next to surface tiles of 64: in arrival order one
// heavy tile claimed last cost a slab rank at the collision interface 45 % more P2G / G2P time for 13 % more particles).
// `ends`: shared memory, inclusive prefix sums of the class counts (work_prefix, before the first claim).
__device__ __forceinline__ void work_prefix(const WorkList& W, uint32_t* ends) {
  if (threadIdx.x == 0) {
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < WORK_CLASSES; ++c) { acc += W.count[c]; ends[c] = acc; }
  }
}
// -> tile id (0xffffffff: no work left) and which part of its run; `boundary`: an item whose completion the concurrent sender counts
// `part` = which part | number of parts << 8; `tick`: what the item's completion adds to W.done when it belongs to a boundary tile (a
// whole tile counts `parts`, a part 1: the sender waits for boundary tiles * parts), 0 for an interior item
__device__ __forceinline__ uint32_t work_claim(const WorkList& W, const uint32_t* ends, uint32_t& tick, uint32_t& part) {
  const uint32_t item = atomicAdd(W.cursor, 1u);
  const uint32_t total = ends[WORK_CLASSES - 1];
  const uint32_t n_whole = total - min(W.parts > 1 ? W.split_last : 0u, total);
  uint32_t q = item, nparts = 1;
  part = 1u << 8;
  if (item >= n_whole) {
    const uint32_t r = item - n_whole;
    q = n_whole + r / W.parts;
    nparts = W.parts;
    part = (r - (q - n_whole) * W.parts) | (W.parts << 8);
  }
  tick = 0;
  if (q >= total) return 0xffffffffu;
  uint32_t cls = 0, base = 0;
#pragma unroll
  for (int c = 0; c < WORK_CLASSES - 1; ++c) {
    const uint32_t e = ends[c];
    if (q >= e) { cls = c + 1; base = e; }
  }
  if (cls < WORK_CLASSES / 2) tick = nparts == 1 ? W.parts : 1u;
  return W.ids[(size_t)cls * W.split + (q - base)];
}
// the particle sub-range of a work item
__device__ __forceinline__ uint2 work_range(uint2 run, uint32_t part) {
  const uint32_t nparts = part >> 8, k = part & 0xffu;
  if (nparts == 1) return run;
  const uint32_t span = 512u / nparts;
  const uint32_t s = run.x + k * span;
  const uint32_t e = k + 1 == nparts ? run.y : min(run.y, s + span);
  return make_uint2(min(s, run.y), max(e, min(s, run.y)));
}
// the sender's side: true once `*done` has reached `*count * parts` (gives up after ~2 s like wait_seq)
__device__ __forceinline__ bool wait_boundary_done(const uint32_t* done, const uint32_t* count, uint32_t parts) {
  const uint32_t want = *count * parts;
  for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(done) : "memory");
    if (v >= want) return true;
    __nanosleep(spin < 64 ? 64 : 400);
  }
  return false;
}
struct SlabColumns {
  int lo, hi;
  uint32_t* list;          // [WORK_CLASSES * tile_cap]: the work lists (WorkList)
  int slab;                // peer-memory slab rank: tiles of block columns lo and hi - 1 are boundary tiles
  int by_size;             // 0: one class per half, tiles in arrival order (SVB_WORK_ORDER=0, for A/B runs)
};
// `Sprev` != null: this substep's set was filled ahead of time by the previous substep's G2P, so this is the FIRST kernel of the substep
// and it folds what the previous substep's back half raised (a FAILED particle, table status bits, exchange errors) into this
// substep's scalars — they were initialised before that back half ran — plus, on slab ranks, the row count migration left behind.
__global__ void __launch_bounds__(256) k_offsets(StepScalars* S, const StepScalars* __restrict__ Sprev, uint32_t n_rows, const uint32_t* __restrict__ n_dev, TileTable T,
                                                 uint32_t* __restrict__ cell_count, uint2* __restrict__ tile_range, uint32_t* __restrict__ slot_first, const uint32_t* __restrict__ tile_touch,
                                                 int* __restrict__ nbr, SlabColumns cols) {
  grid_dependency_wait();
  bool aborted = SVB_ABORTED(S);
  if (Sprev) {
    const uint32_t carry = Sprev->status & ST_CARRY_MASK, stop = Sprev->sticky | Sprev->sticky_new;
    aborted = aborted || carry || stop;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      S->sticky |= stop;
      S->accum |= Sprev->accum | (Sprev->status & 0xffffu);
      if (carry) atomicOr(&S->status, carry);
      S->n = n_dev ? min(*n_dev, n_rows) : n_rows;
      if (!aborted) S->bin_blocks_done = 1u;   // "this substep was binned"
    }
  }
  if (aborted) return;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t n_ptiles = min(S->n_ptiles, T.tile_cap);   // final: only k_bin creates particle tiles
  for (uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n_ptiles; t += warps) {
    const uint32_t slot = T.tile_slot[t];   // the counters of a particle tile live at its table slot
    // halo first: its table probes overlap the scan below
    int r = -1;
    if (lane == 0) r = (int)t;
    else if (lane < 8 && ((tile_touch[slot] >> lane) & 1u)) {
      const uint32_t id = tile_find_or_insert(T, tile_key_offset(T.tile_key[t], (int)lane), S);
      r = id == TILE_PENDING ? -1 : (int)id;
    }
    const uint2 c = *reinterpret_cast<const uint2*>(cell_count + (size_t)slot * 64 + 2 * lane);
    const uint32_t mine = c.x + c.y;
    uint32_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(SVB_FULL, inc, o);
      if (lane >= (uint32_t)o) inc += v;
    }
    const uint32_t ex = inc - mine;
    *reinterpret_cast<uint2*>(cell_count + (size_t)slot * 64 + 2 * lane) = make_uint2(ex, ex + c.x);
    // the tile's run in the binned order: claimed from a cursor, so the runs are dense but in no particular tile order
    // (any order is a valid binning; this replaces a single-CTA scan over the tile totals and its launch)
    if (lane == 31) {
      const uint32_t first = atomicAdd(&S->n_live, inc);
      tile_range[t] = make_uint2(first, first + inc);
      slot_first[slot] = first;
      int which = 1;
      if (cols.slab) {
        int bx, by, bz;
        uint32_t layer;
        tile_key_unpack(T.tile_key[t], bx, by, bz, layer);
        which = (bx == cols.lo || bx == cols.hi - 1) ? 0 : 1;
        atomicAdd(&S->n_work[which], 1u);
      }
      const uint32_t cls = (uint32_t)which * (WORK_CLASSES / 2) + (cols.by_size ? 7u - min(7u, inc >> 7) : 0u);
      cols.list[(size_t)cls * T.tile_cap + atomicAdd(&S->n_class[cls], 1u)] = t;
    }
    if (lane < 8) nbr[(size_t)t * 8 + lane] = r;
  }
}

// single CTA: exclusive scan in place; a[n] = total; *total_out = total
__global__ void __launch_bounds__(1024) k_scan_tiles(uint32_t* a, const uint32_t* n_ptr, uint32_t n_fixed, uint32_t* total_out) {
  __shared__ uint32_t ws[32];
  __shared__ uint32_t carry;
  const uint32_t n = n_ptr ? *n_ptr : n_fixed;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < n ? a[i] : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(SVB_FULL, inc, o);
      if ((threadIdx.x & 31) >= o) inc += t;
    }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t w = ws[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(SVB_FULL, w, o);
        if (threadIdx.x >= o) w += t;
      }
      ws[threadIdx.x] = w;
    }
    __syncthreads();
    const uint32_t warp_off = (threadIdx.x >> 5) ? ws[(threadIdx.x >> 5) - 1] : 0u;
    const uint32_t c = carry;
    if (i < n) a[i] = c + warp_off + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + warp_off + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    a[n] = carry;
    if (total_out) *total_out = carry;
  }
}

// re-bin (sort.rs:91-101).  Slot of particle i in the binned order:
//   j = slot_first[slot] + cell_offset[slot*64 + cell] + rank      (slot = the tile's table slot; tombstoned: n_live + rank)
// Only the inverse map src_of[j] = i is materialised here (4 B per particle).  P2G gathers its inputs
// through it and G2P writes its results — and the fields it merely carries — to slot j of the other
// buffer, so the physical permutation of the 144-byte state costs no pass of its own: consecutive
// slots come from (nearly) consecutive rows of the previous order, the gathers stay coalesced.
// The same launch also clears the grid tiles of this substep (blocks beyond the particle range).
// With `prep.blocks` > 0 the last blocks also prepare the OTHER front set for the G2P of this substep, which bins the advanced
// positions into it (reset_set; its scalars start from this substep's, k_offsets of the next substep folds the rest).
struct PrepareNext {
  StepScalars* S; TileTable T; uint32_t* cell_count; uint32_t* tile_touch;
  uint32_t n; const uint32_t* n_dev; int tables_fresh; uint32_t blocks;
};
template <int R>
__global__ void __launch_bounds__(256) k_invert_zero(StepScalars* __restrict__ S, const uint32_t* __restrict__ pcell, uint32_t* __restrict__ cell_offset,
                                                     const uint32_t* __restrict__ slot_first, uint32_t* __restrict__ src_of, uint32_t n, uint32_t invert_blocks, float4* __restrict__ grid,
                                                     unsigned long long* __restrict__ node_mask, uint32_t tile_cap, PrepareNext prep) {
  if (SVB_ABORTED(S)) return;
  if (blockIdx.x >= gridDim.x - prep.blocks) {
    reset_set(S, prep.S, prep.T, prep.cell_count, prep.tile_touch, prep.n, prep.n_dev, prep.tables_fresh, blockIdx.x - (gridDim.x - prep.blocks), prep.blocks);
    return;
  }
  if (blockIdx.x >= invert_blocks) {
    const size_t total = (size_t)min(S->n_tiles, tile_cap) * 64;
    const size_t stride = (size_t)(gridDim.x - prep.blocks - invert_blocks) * blockDim.x;
    for (size_t q = (size_t)(blockIdx.x - invert_blocks) * blockDim.x + threadIdx.x; q < total; q += stride) {
      grid[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (node_mask && (q & 63) == 0) node_mask[q >> 6] = 0ull;
    }
    if (blockIdx.x == invert_blocks && threadIdx.x == 0) S->n_tiles_zeroed = (uint32_t)(total >> 6);
    return;
  }
  // the rank of a particle inside its cell is handed out HERE, from the cell's offset used as a cursor (one atomic per distinct
  // cell per warp: consecutive rows share cells) — the binning itself only counted, so the kernels that bin never wait on an atomic.
  // Every row is a chain of three dependent round trips (cell word -> cursor atomic -> the tile's first slot): a thread carries R
  // rows through it side by side, so that one wave of CTAs covers the rows instead of R.
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t n_rows = min(n, S->n);
  uint32_t row[R], ci[R], base[R];
  unsigned peers[R];
#pragma unroll
  for (int k = 0; k < R; ++k) {
    row[k] = (blockIdx.x * R + k) * blockDim.x + threadIdx.x;
    ci[k] = row[k] < n_rows ? pcell[row[k]] : 0xfffffffeu;
  }
#pragma unroll
  for (int k = 0; k < R; ++k) {
    peers[k] = __match_any_sync(SVB_FULL, ci[k]);
    base[k] = 0;
    if ((int)lane == __ffs(peers[k]) - 1) {
      if (ci[k] < 0xfffffffdu) base[k] = atomicAdd(&cell_offset[ci[k]], (uint32_t)__popc(peers[k]));
      else if (ci[k] == 0xffffffffu) base[k] = atomicAdd(&S->tomb_cursor, (uint32_t)__popc(peers[k]));
    }
  }
  uint32_t first[R];
#pragma unroll
  for (int k = 0; k < R; ++k) first[k] = ci[k] < 0xfffffffdu ? slot_first[ci[k] >> 6] : S->n_live;
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const uint32_t r = __shfl_sync(SVB_FULL, base[k], __ffs(peers[k]) - 1) + __popc(peers[k] & ((1u << lane) - 1u));
    if (ci[k] < 0xfffffffdu || ci[k] == 0xffffffffu) src_of[first[k] + r] = row[k];   // (migrated away, or unbinned after an abort: no slot)
  }
}
// ------------------------------------------------------------------------------------------------
// P2G (scatter_momentum.rs:22-93).  One CTA per (block, layer) run of particles, claimed from a
// work counter.  Each warp takes 32 consecutive particles: every lane evaluates ITS particle once
// (affine momentum matrix A = m C - s V0 P F^T - s J V0 sigma, so the stress is computed once per
// particle, not 27 times) and parks 16 numbers in shared memory: the fractional position x/h - base,
// the mass, m v and A.  Then the warp walks the 32 particles together with lane = one of the 27 stencil
// nodes: four warp-uniform LDS.128 per particle, the lane's own weight and offset derived in registers
// (packed fp32), accumulating in registers while consecutive particles share a cell (they do: the run is
// sorted by cell) — the warp-aggregated form of the scatter.  A cell change flushes 27 float4 into the
// warp's private 6x6x6 tile (plain read-modify-write, no shared atomics: fp32 shared atomics are CAS
// loops).  The walk is bound by shared-memory bandwidth; measured on B200 (profiles/lds_bench.cu) a
// warp-uniform LDS.128 costs 2.2 cycles, a 3-address LDS.64 2 cycles and a 3-address LDS.128 4 cycles,
// which is why the earlier layout (per-axis (w, d) pairs + 13 uniform floats = 13.6 cycles per particle)
// lost to this one (8.8 cycles): 95 -> 80 us at 1 M particles.  (Also measured slower: lane = z-column of
// the stencil with three interleaved particles per warp iteration — serialised flushes.)  At the end the
// warps' tiles are summed and every non-zero tile node goes to HBM with one red.global.add.v4.f32.
// Stencil weights from the base node: `shifted` = x/h - base in [1/2, 3/2).  Branch-free forms of
// kernel_quadratic(shifted - a), a = 0,1,2 (cpu/src/kernels.rs:17-26 evaluated on the branch each a falls in;
// both branches agree at the hand-over points).
__device__ __forceinline__ void quad_weights(float shifted, float* w) {
  const float a0 = fmaxf(1.5f - shifted, 0.f), a1 = shifted - 1.f, a2 = fmaxf(shifted - 0.5f, 0.f);
  w[0] = 0.5f * a0 * a0;
  w[1] = 0.75f - a1 * a1;
  w[2] = 0.5f * a2 * a2;
}
#ifndef SVB_P2G_CTAS_PER_SM
#define SVB_P2G_CTAS_PER_SM 7   // 71 registers, 24 KB of shared memory per CTA; measured: 6 -> 81 us, 7 -> 79 us, 8 (spills) -> 88 us at 1 M
#endif
#ifndef SVB_P2G_WARPS
#define SVB_P2G_WARPS 4
#endif
constexpr int P2G_WARPS = SVB_P2G_WARPS;
constexpr int P2G_CTAS_PER_SM = SVB_P2G_CTAS_PER_SM;
constexpr int TILE_NODES = 216;
constexpr int STAGE_STRIDE = 20;  // floats per staged particle (16 used + the cell id): 16-byte aligned rows, conflict-free float4 stores
constexpr int P2G_SMEM = P2G_WARPS * TILE_NODES * 16 + P2G_WARPS * 32 * STAGE_STRIDE * 4;

// Blackwell packed fp32: one FFMA2 / FADD2 issue slot does two lanes of work (SASS `FFMA2 Rd, Ra.F32x2.HI_LO, Rb.F32, Rc.F32x2.HI_LO`
// — the scalar operand is broadcast by the instruction, no register pair has to be built for it).
__device__ __forceinline__ float2 ffma2(float2 a, float s, float2 c) {
  unsigned long long d;
  const float2 b = make_float2(s, s);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)),
      "l"(*reinterpret_cast<const unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fmul2(float2 a, float s) {
  unsigned long long d;
  const float2 b = make_float2(s, s);
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)),
      "l"(*reinterpret_cast<const unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void red_add_v4(float4* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Shared memory through 32-bit window addresses (P2G's walk): the address arithmetic of a run is one integer add, and an `opaque`
// value cannot be rematerialised — ptxas otherwise recomputes cheap per-lane constants (and the shared window base) inside every run
// of the walk to keep them out of the prologue's register budget: 25 of the 66 instructions a run cost besides its particles.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t opaque(uint32_t v) { asm volatile("mov.b32 %0, %0;" : "+r"(v)); return v; }
template <int OFF>
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a), "n"(OFF) : "memory");
  return v;
}
template <int OFF>
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// The external force (external_force.rs:17-50) is applied here, in registers: v + dt g, or (goal - x) / dt for a goal particle.  The
// forced velocity is only ever read by the scatter (collect_velocity.rs overwrites v), so it is never stored.
struct ForceIn {
  GoalDev G;
  float dt, gx, gy, gz, factor_b;   // dt: the time step ExternalForce sees (with adaptive steps the one BEFORE LimitTimeStepBeforeForce)
  const DtState* D;                 // adaptive steps: dt (scatter and force), gravity and the frame factor come from the device clock
};
template <bool HAS_GOALS>
__global__ void __launch_bounds__(P2G_WARPS * 32, P2G_CTAS_PER_SM) k_p2g(ParticleBuf P, const uint32_t* __restrict__ src_of, const uint2* __restrict__ group_range, const int* __restrict__ nbr,
                                                           StepScalars* S, float4* __restrict__ grid, float h, float dt, ForceIn force, WorkList W) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* tiles = reinterpret_cast<float4*>(smem_raw);
  float* stage_all = reinterpret_cast<float*>(smem_raw + P2G_WARPS * TILE_NODES * 16);
  __shared__ uint32_t s_group, s_part, s_ends[WORK_CLASSES];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4* my_tile = tiles + warp * TILE_NODES;
  float* stage = stage_all + warp * 32 * STAGE_STRIDE;
  grid_dependency_wait();
  if (SVB_ABORTED(S)) return;
  work_prefix(W, s_ends);
  if (force.D) {
    dt = force.D->allowed;
    force.dt = force.D->dt_force; force.gx = force.D->g[0]; force.gy = force.D->g[1]; force.gz = force.D->g[2]; force.factor_b = force.D->factor_b;
  }
  const float scaling = dt * 4.f / (h * h);
  // node handled by this lane in the 3x3x3 stencil (k fastest); lanes 27..31 shadow node 0 and never flush
  const bool node_lane = lane < 27;
  const int li = node_lane ? lane / 9 : 0, lj = node_lane ? (lane / 3) % 3 : 0, lk = node_lane ? lane % 3 : 0;
  // staged row (floats): 0..3 t = x/h - base (3), mass | 4..7 b.x, b.y, B0, B1 | 8..11 B3, B4, B6, B7 | 12..15 b.z, B2, B5, B8 | 16 cell
  // with B = A h and b = m v - B t, so that the momentum a particle sends to the node at stencil offset a in {0,1,2}^3 is
  //   w (m v + A (a - t) h) = w (b + B a):   the lane's offset a is a constant, no per-particle distance has to be formed.
  // Every walk load is a warp-uniform LDS.128 (2.2 cycles of shared-memory bandwidth per 16 bytes, measured; a 3-address
  // LDS.64 costs 2 cycles per 8 bytes): the walk was bound by shared-memory bandwidth with staged per-axis (w, d) pairs, so
  // each lane derives the weight of ITS node from t with per-lane polynomial coefficients instead:
  //   kernel_quadratic on the branch node a falls in (cpu/src/kernels.rs:17-26), t in [1/2, 3/2):
  //   a = 0: 1/2 (3/2 - t)^2 = t^2/2 - 3t/2 + 9/8     a = 1: 3/4 - (t - 1)^2 = -t^2 + 2t - 1/4     a = 2: 1/2 (t - 1/2)^2 = t^2/2 - t/2 + 1/8
  const float2 kp2_xy = make_float2(li == 1 ? -1.f : 0.5f, lj == 1 ? -1.f : 0.5f);
  const float2 kp1_xy = make_float2(li == 0 ? -1.5f : li == 1 ? 2.f : -0.5f, lj == 0 ? -1.5f : lj == 1 ? 2.f : -0.5f);
  const float2 kp0_xy = make_float2(li == 0 ? 1.125f : li == 1 ? -0.25f : 0.125f, lj == 0 ? 1.125f : lj == 1 ? -0.25f : 0.125f);
  const float kp2_z = lk == 1 ? -1.f : 0.5f, kp1_z = lk == 0 ? -1.5f : lk == 1 ? 2.f : -0.5f, kp0_z = lk == 0 ? 1.125f : lk == 1 ? -0.25f : 0.125f;
  const float fli = (float)li, flj = (float)lj, flk = (float)lk;
  const int lane_tile_off = (li * 6 + lj) * 6 + lk;
  // the lane's polynomial coefficients and stencil offsets live in shared memory and are fetched after each chunk's prologue (three
  // LDS.128 per 32 particles), so they never compete with the prologue for registers
  __shared__ float4 s_lane_k[3 * 32];
  if (warp == 0) {
    s_lane_k[lane] = make_float4(kp2_xy.x, kp2_xy.y, kp2_z, fli);
    s_lane_k[32 + lane] = make_float4(kp1_xy.x, kp1_xy.y, kp1_z, flj);
    s_lane_k[64 + lane] = make_float4(kp0_xy.x, kp0_xy.y, kp0_z, flk);
  }
  const uint32_t lane_k_s = opaque(smem_addr(s_lane_k) + lane * 16);
  const uint32_t stage_s = opaque(smem_addr(stage));
  const uint32_t tile_s = opaque(smem_addr(my_tile) + (uint32_t)lane_tile_off * 16u);   // the lane's node of cell (0,0,0) in the warp's tile

  uint32_t ticking = 0;   // thread 0: the item just finished belonged to a boundary tile (its weight in W.done)
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) {
      // the previous tile's sums are in HBM (every thread's reductions precede the barrier above; one fence of the ticking thread
      // orders them before the tick): tell the concurrent halo sender
      if (ticking && W.done) { __threadfence(); atomicAdd(W.done, ticking); }
      s_group = work_claim(W, s_ends, ticking, s_part);
    }
    __syncthreads();
    const uint32_t g = s_group;
    if (g == 0xffffffffu) break;
    const uint2 range = work_range(group_range[g], s_part);
    const uint32_t start = range.x, end = range.y;
    if (start >= end) continue;   // (a short run has no particles for this part)
    for (int q = threadIdx.x; q < P2G_WARPS * TILE_NODES; q += blockDim.x) tiles[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    for (uint32_t chunk = start + warp * 32; chunk < end; chunk += P2G_WARPS * 32) {
      // ---- per-particle evaluation by the owning lane
      float4 st[4];
      int cell = -1;
      if (chunk + lane < end) {
        const uint32_t i = src_of[chunk + lane];  // row of this particle in the pre-bin order
        // the particle's nine quads (svb_device.cuh: Field): 16-byte gathers, consecutive slots come from (nearly) consecutive rows
        const float4 pq0 = P.q(0)[i], pq1 = P.q(1)[i], pq2 = P.q(2)[i], pq3 = P.q(3)[i], pq4 = P.q(4)[i], pq5 = P.q(5)[i], pq6 = P.q(6)[i], pq7 = P.q(7)[i];
        const float2 pq8 = *reinterpret_cast<const float2*>(P.q(8) + i);
        const float x0 = pq0.x, x1 = pq0.y, x2 = pq0.z;
        const float n0 = __fdiv_rn(x0, h), n1 = __fdiv_rn(x1, h), n2 = __fdiv_rn(x2, h);
        const int s0 = (int)floorf(__fsub_rn(n0, 0.5f)), s1 = (int)floorf(__fsub_rn(n1, 0.5f)), s2 = (int)floorf(__fsub_rn(n2, 0.5f));
        cell = ((s0 & 3) << 4) | ((s1 & 3) << 2) | (s2 & 3);
        M3 C, F;
        F.m[0] = pq1.x; F.m[1] = pq1.y; F.m[2] = pq1.z; F.m[3] = pq1.w; F.m[4] = pq2.x; F.m[5] = pq2.y; F.m[6] = pq2.z; F.m[7] = pq2.w; F.m[8] = pq3.x;
        C.m[0] = pq6.y; C.m[1] = pq6.z; C.m[2] = pq6.w; C.m[3] = pq7.x; C.m[4] = pq7.y; C.m[5] = pq7.z; C.m[6] = pq7.w; C.m[7] = pq8.x; C.m[8] = pq8.y;
        const float mass = pq3.y, vol = pq3.z, p0 = pq3.w, p1 = pq4.x;
        const uint32_t flags = __float_as_uint(pq0.w);
        const M3 stress = (flags & F_IS_FLUID) ? first_piola_inviscid(p0, (int)p1, F) : first_piola_neo_hookean(p0, p1, F);
        const M3 pft = mul_nt(stress, F);
        const float sv = scaling * vol;
        M3 A;
#pragma unroll
        for (int q = 0; q < 9; ++q) A.m[q] = mass * C.m[q] - sv * pft.m[q];
        if (flags & F_USE_VISCOSITY) {
          const M3 cauchy = viscous_cauchy(pq4.z, pq4.w, C);
          const float sj = scaling * det(F) * vol;
#pragma unroll
          for (int q = 0; q < 9; ++q) A.m[q] -= sj * cauchy.m[q];
        }
        V3 v = V3{pq5.z, pq5.w, pq6.x};
        {   // ExternalForce, after the collide pass of the same substep like phase/mod.rs:27-41
          bool goal = false;
          if (HAS_GOALS) {
            const uint32_t o = __float_as_uint(pq5.y);
            if ((force.G.flags_a[o] & F_HAS_GOAL) && (force.G.flags_b[o] & F_HAS_GOAL)) {
              const float fa = 1.f - force.factor_b;
              const V3 target = fa * ld3(force.G.goal_a, o) + force.factor_b * ld3(force.G.goal_b, o);
              v = (target - V3{x0, x1, x2}) / force.dt;
              goal = true;
            }
          }
          if (!goal) v = v + force.dt * V3{force.gx, force.gy, force.gz};
        }
        const float mv0 = mass * v.x, mv1 = mass * v.y, mv2 = mass * v.z;
        const float t0 = n0 - (float)s0, t1 = n1 - (float)s1, t2 = n2 - (float)s2;
#pragma unroll
        for (int q = 0; q < 9; ++q) A.m[q] *= h;
        const float b0 = mv0 - (A.m[0] * t0 + A.m[3] * t1 + A.m[6] * t2);
        const float b1 = mv1 - (A.m[1] * t0 + A.m[4] * t1 + A.m[7] * t2);
        const float b2 = mv2 - (A.m[2] * t0 + A.m[5] * t1 + A.m[8] * t2);
        st[0] = make_float4(t0, t1, t2, mass);
        st[1] = make_float4(b0, b1, A.m[0], A.m[1]);
        st[2] = make_float4(A.m[3], A.m[4], A.m[6], A.m[7]);
        st[3] = make_float4(b2, A.m[2], A.m[5], A.m[8]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) st[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      // runs of equal cells inside this chunk (the run is sorted by cell): heads as a warp-uniform mask
      const int prev_cell = __shfl_up_sync(SVB_FULL, cell, 1);
      uint32_t heads = __ballot_sync(SVB_FULL, cell >= 0 && (lane == 0 || prev_cell != cell));
      const uint32_t valid = __ballot_sync(SVB_FULL, cell >= 0);
      __syncwarp();
      float4* row = reinterpret_cast<float4*>(stage + lane * STAGE_STRIDE);
#pragma unroll
      for (int q = 0; q < 4; ++q) row[q] = st[q];
      stage[lane * STAGE_STRIDE + 16] = __int_as_float((((cell >> 4) * 6 + ((cell >> 2) & 3)) * 6 + (cell & 3)) * 16);   // byte offset of the cell's first node in a tile
      __syncwarp();

      // ---- cooperative walk: lane = stencil node, one register accumulator per run of equal cells
      const int count = __popc(valid);
      {
        const float4 k2 = lds128<0>(lane_k_s), k1 = lds128<512>(lane_k_s), k0 = lds128<1024>(lane_k_s);
        const float2 c2_xy = make_float2(k2.x, k2.y), c1_xy = make_float2(k1.x, k1.y), c0_xy = make_float2(k0.x, k0.y);
        while (heads) {
          const int first = __ffs(heads) - 1;
          heads &= heads - 1;
          const int last = heads ? __ffs(heads) - 1 : count;
          float2 acc_xy = make_float2(0.f, 0.f);
          float acc_z = 0.f, acc_m = 0.f;
          const uint32_t run_s = stage_s + (uint32_t)first * (STAGE_STRIDE * 4);
          uint32_t sp = run_s;
          auto node_update = [&](const float4 q0, const float4 q1, const float4 q2, const float4 q3) {
            const float2 t_xy = make_float2(q0.x, q0.y);
            const float2 w_xy = ffma2(ffma2(c2_xy, t_xy, c1_xy), t_xy, c0_xy);
            const float w_z = fmaf(fmaf(k2.z, q0.z, k1.z), q0.z, k0.z);
            const float wgt = w_xy.x * w_xy.y * w_z;
            const float2 m_xy = ffma2(make_float2(q2.z, q2.w), k0.w, ffma2(make_float2(q2.x, q2.y), k1.w, ffma2(make_float2(q1.z, q1.w), k2.w, make_float2(q1.x, q1.y))));
            const float m_z = fmaf(q3.w, k0.w, fmaf(q3.z, k1.w, fmaf(q3.y, k2.w, q3.x)));
            acc_xy = ffma2(m_xy, wgt, acc_xy);
            acc_z = fmaf(wgt, m_z, acc_z);
            acc_m = fmaf(wgt, q0.w, acc_m);
          };
          constexpr int RB = STAGE_STRIDE * 4;   // bytes per staged row
          int left = last - first;
          for (; left >= 4; left -= 4) {   // whole groups of four without a trip test in between
            node_update(lds128<0>(sp), lds128<16>(sp), lds128<32>(sp), lds128<48>(sp));
            node_update(lds128<RB>(sp), lds128<RB + 16>(sp), lds128<RB + 32>(sp), lds128<RB + 48>(sp));
            node_update(lds128<2 * RB>(sp), lds128<2 * RB + 16>(sp), lds128<2 * RB + 32>(sp), lds128<2 * RB + 48>(sp));
            node_update(lds128<3 * RB>(sp), lds128<3 * RB + 16>(sp), lds128<3 * RB + 32>(sp), lds128<3 * RB + 48>(sp));
            sp += 4 * RB;
          }
          if (left & 2) {
            node_update(lds128<0>(sp), lds128<16>(sp), lds128<32>(sp), lds128<48>(sp));
            node_update(lds128<RB>(sp), lds128<RB + 16>(sp), lds128<RB + 32>(sp), lds128<RB + 48>(sp));
            sp += 2 * RB;
          }
          if (left & 1) node_update(lds128<0>(sp), lds128<16>(sp), lds128<32>(sp), lds128<48>(sp));
          if (node_lane) {
            const uint32_t a = tile_s + lds32<64>(run_s);
            float4 o = lds128<0>(a);
            o.x += acc_xy.x; o.y += acc_xy.y; o.z += acc_z; o.w += acc_m;
            sts128(a, o);
          }
          __syncwarp();   // the next run's flush reads nodes another lane has just written (neighbouring cells share 18 of their 27 nodes)
        }
      }
      __syncwarp();
    }
    __syncthreads();
    // ---- tile -> HBM
    for (int t = threadIdx.x; t < TILE_NODES; t += blockDim.x) {
      float4 sum = tiles[t];
#pragma unroll
      for (int w = 1; w < P2G_WARPS; ++w) {
        const float4 o = tiles[w * TILE_NODES + t];
        sum.x += o.x; sum.y += o.y; sum.z += o.z; sum.w += o.w;
      }
      if (sum.x != 0.f || sum.y != 0.f || sum.z != 0.f || sum.w != 0.f) {
        const int ti = t / 36, tj = (t / 6) % 6, tk = t % 6;
        const int d = (ti >> 2) | ((tj >> 2) << 1) | ((tk >> 2) << 2);
        const int nb = nbr[(size_t)g * 8 + d];
        if (nb >= 0) red_add_v4(grid + (size_t)nb * 64 + (((ti & 3) << 4) | ((tj & 3) << 2) | (tk & 3)), sum);
        else atomicOr(&S->status, 2u /*SVB_TABLE_ENTRY_MISSING*/);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// meld (meld_grid.rs:16-69) evaluated while loading a tile: node value seen by a layer = sum over the
// layers of the same block whose collider bits are compatible; velocity = momentum / mass.
struct MeldInfo {
  const unsigned long long* layer_slots;
  const uint32_t* layer_list;
};
__device__ __forceinline__ float4 finish_node(float4 s) {
  if (s.w > 0.f) { const float r = __frcp_rn(s.w); s.x *= r; s.y *= r; s.z *= r; }  // p/m within 1 ulp of the division
  else { s.x = 0.f; s.y = 0.f; s.z = 0.f; }
  return s;
}
// sibling tiles (ids) of `key`'s block compatible with its layer, self first.  Returns the count.
__device__ __forceinline__ int sibling_tiles(const TileTable& T, const MeldInfo& Mi, uint32_t n_layers, unsigned long long key, int self, int* out, StepScalars* S) {
  int n = 0;
  out[n++] = self;
  if (n_layers == 0) return n;
  const uint32_t lmask = (1u << LAYER_BITS) - 1u;
  const uint32_t my_layer = (uint32_t)key & lmask;
  const uint32_t mine = layer_bits_of(Mi.layer_slots, my_layer);
  const unsigned long long block = key & ~(unsigned long long)lmask;
  for (uint32_t q = 0; q <= n_layers; ++q) {
    const uint32_t layer = q == 0 ? 0u : Mi.layer_list[q - 1];
    if (layer == my_layer) continue;
    if (!bits_compatible(mine, layer_bits_of(Mi.layer_slots, layer))) continue;
    const int t = tile_find(T, block | layer);
    if (t < 0) continue;
    if (n < SIB_MAX) out[n++] = t;
    else atomicOr(&S->status, 1u /*SVB_TABLE_TRIES_EXCEEDED*/);
  }
  return n;
}

// melded velocity grid (collider scenes only): per tile, sum the compatible sibling layers and divide by
// the mass once per node (meld_grid.rs:41-66), so that G2P can bulk-copy finished tiles.
__global__ void __launch_bounds__(256) k_meld(const StepScalars* __restrict__ S, const float4* __restrict__ grid, float4* __restrict__ melded, TileTable T, MeldInfo Mi) {
  if (SVB_ABORTED(S)) return;
  __shared__ int sib[4][SIB_MAX];
  __shared__ int nsib[4];
  const uint32_t n_tiles = min(S->n_tiles, T.tile_cap), n_layers = S->n_layers;
  const int sub = threadIdx.x >> 6, node = threadIdx.x & 63;
  for (uint32_t base = blockIdx.x * 4; base < n_tiles; base += gridDim.x * 4) {
    const uint32_t t = base + sub;
    __syncthreads();
    if (node == 0 && t < n_tiles) nsib[sub] = sibling_tiles(T, Mi, n_layers, T.tile_key[t], (int)t, sib[sub], const_cast<StepScalars*>(S));
    __syncthreads();
    if (t >= n_tiles) continue;
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < nsib[sub]; ++q) {
      const float4 v = grid[(size_t)sib[sub][q] * 64 + node];
      sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
    }
    melded[(size_t)t * 64 + node] = finish_node(sum);
  }
}

// G2P (+ advance, return mapping, energy, cull when FUSE).  One CTA per particle-owning tile ; the 6x6x6 velocity tile is staged in shared memory, then one thread per particle
// gathers with compile-time offsets.  Collider scenes read the melded velocity grid of k_meld.
// (Two software-pipelined variants — TMA bulk tile copies + cp.async particle staging, per CTA and per
// warp — were measured slower: 4-byte cp.async doubles the LSU instructions per particle and the
// staging buffers cost occupancy; see profiles/README.md r1c-r1g.)
#ifndef SVB_G2P_THREADS
#define SVB_G2P_THREADS 64   // two warps per CTA, 14 CTAs per SM: a CTA's serial sections (claim, tile load, barriers) stall two warps instead of four; measured 128 x 7 -> 64 x 14: G2P 471 -> 439 us at 8 M
#endif
constexpr int G2P_THREADS = SVB_G2P_THREADS;
#ifndef SVB_G2P_CTAS_PER_SM
#define SVB_G2P_CTAS_PER_SM 14   // (x 64 threads; 7 x 128 before) 72 registers (60 bytes of spills) since the particle words come in quads; measured 5 -> 7: G2P 496 -> 477 us at 8 M, 171 -> 157 us on a 2 M dam break; 6 is slower, 8 the same as 7
#endif
constexpr int G2P_CTAS_PER_SM = SVB_G2P_CTAS_PER_SM;
// Reads the particle through src_of from the pre-bin buffer `P`, writes every field of it to slot i of
// `D` (this is where the physical re-bin happens).
// slab ranks: a particle whose advanced position left the block columns [lo, hi) is noted in a per-side list of binned
// slots while G2P still holds its new position (no pass over all rows to find the few that migrate)
struct MigrateCut {
  int lo, hi, reach_lo, reach_hi;
  float node_lo, node_hi;   // 4 * lo, 4 * hi as floats
  uint32_t* list;     // [2 * cap]: slots leaving to the left, then to the right
  uint32_t* counts;   // [2]
  uint32_t cap;
};
// BIN: the thread that holds a particle's ADVANCED position also bins it for the next substep, into the other front set (bin_warp):
// the next substep then starts at k_offsets — no k_bin pass over the particles, no k_begin launch (scenes without a collider mesh).
struct BinNext {
  StepScalars* S; TileTable T; BinArrays B;
};
template <bool FUSE, bool REDUCE, bool MELDED, bool SLAB, bool BIN>
__global__ void __launch_bounds__(G2P_THREADS, G2P_CTAS_PER_SM) k_g2p(ParticleBuf P, ParticleBuf D, const uint32_t* __restrict__ src_of, float* __restrict__ energy,
                                                        const uint2* __restrict__ group_range, const int* __restrict__ nbr, StepScalars* S,
                                                        const float4* __restrict__ grid, SimConsts K, float dt, MigrateCut mc, BinNext bn,
                                                        const unsigned long long* __restrict__ tile_key, WorkList W) {
  __shared__ float4 tile[TILE_NODES];
  __shared__ uint32_t s_group, s_part, s_ends[WORK_CLASSES];
  __shared__ int s_nbr[8];
  // BIN: a particle moves less than a cell per substep, so its next tile is one of the 27 blocks around this CTA's tile: their table
  // slots (in the NEXT substep's set) are cached, the cell counts and touch masks of the tile's particles are collected in shared
  // memory (native integer atomics) and go to HBM once per tile — no table probe, no global atomic and no warp vote per particle
  __shared__ uint32_t s_cache[27];   // ~0u = not looked up yet
  __shared__ uint32_t s_touch[27];
  __shared__ uint32_t s_cnt[BIN ? 27 * 64 : 1];
  __shared__ int s_block[3];
  if (BIN) {   // (shared-memory set-up overlaps the previous kernel's tail under a programmatic dependent launch)
    for (int q = threadIdx.x; q < 27 * 64; q += blockDim.x) s_cnt[q] = 0u;
    if (threadIdx.x < 27) { s_touch[threadIdx.x] = 0u; s_cache[threadIdx.x] = ~0u; }
  }
  grid_dependency_wait();
  if (SVB_ABORTED(S)) return;
  work_prefix(W, s_ends);
  const float h = K.h, inv_h = 1.f / K.h;
  uint32_t ticking = 0;   // thread 0: the item just finished belonged to a boundary tile (its weight in W.done)
  int red_vel = INT32_MIN, red_def = INT32_MAX;
  uint32_t failed = 0;
  for (;;) {
    __syncthreads();
    if (BIN) {   // flush what the previous tile's particles counted (every thread flushes and clears its own counters)
      for (int q = threadIdx.x; q < 27 * 64; q += blockDim.x) {
        const uint32_t c = s_cnt[q];
        if (c) { atomicAdd(&bn.B.cell_count[(size_t)s_cache[q >> 6] * 64 + (q & 63)], c); s_cnt[q] = 0u; }
      }
      if (threadIdx.x < 27) {
        const uint32_t m = s_touch[threadIdx.x];
        if (m) { atomicOr(&bn.B.tile_touch[s_cache[threadIdx.x]], m); s_touch[threadIdx.x] = 0u; }
      }
    }
    if (threadIdx.x < 8) {
      uint32_t g0 = 0xffffffffu;
      if (threadIdx.x == 0) {
        // the previous tile's rows (and its leavers' list entries) are written: tell the concurrent migration sender
        if (ticking && W.done) { __threadfence(); atomicAdd(W.done, ticking); }
        s_group = g0 = work_claim(W, s_ends, ticking, s_part);
      }
      g0 = __shfl_sync(0xffu, g0, 0);
      if (g0 != 0xffffffffu) s_nbr[threadIdx.x] = nbr[(size_t)g0 * 8 + threadIdx.x];
      if (BIN && threadIdx.x == 0 && g0 != 0xffffffffu) {
        int bx, by, bz;
        uint32_t layer;
        tile_key_unpack(tile_key[g0], bx, by, bz, layer);
        s_block[0] = bx; s_block[1] = by; s_block[2] = bz;
      }
    }
    __syncthreads();
    if (BIN && threadIdx.x >= 32 && threadIdx.x < 32 + 27) s_cache[threadIdx.x - 32] = ~0u;   // (the flush above has read the old slots)
    const uint32_t g = s_group;
    if (g == 0xffffffffu) break;
    const uint2 range = work_range(group_range[g], s_part);
    const uint32_t start = range.x, end = range.y;
    if (start >= end) continue;   // (a short run has no particles for this part)
    for (int t = threadIdx.x; t < TILE_NODES; t += blockDim.x) {
      const int ti = t / 36, tj = (t / 6) % 6, tk = t % 6;
      const int nb = s_nbr[(ti >> 2) | ((tj >> 2) << 1) | ((tk >> 2) << 2)];
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (nb >= 0) v = grid[(size_t)nb * 64 + (((ti & 3) << 4) | ((tj & 3) << 2) | (tk & 3))];
      tile[t] = MELDED ? v : finish_node(v);
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    // a warp's 32 slots start on a 128-byte line of the destination arrays: every store of a warp is one line, not two (G2P 80 -> 77 us at 1 M)
    // the rows of the warp's NEXT iteration are requested into L2 while this one computes (row indices two iterations ahead)
    const uint32_t i_first = (start & ~31u) + threadIdx.x;
    uint32_t s_cur = (i_first >= start && i_first < end) ? src_of[i_first] : 0u;
    uint32_t s_nxt = i_first + blockDim.x < end ? src_of[i_first + blockDim.x] : 0u;
    for (uint32_t base = (start & ~31u) + (threadIdx.x & ~31u); base < end; base += blockDim.x) {   // warp-uniform trip count (bin_warp is warp-collective)
      const uint32_t i = base + lane;
      int bin_state = 2;
      V3 bin_x = V3{0.f, 0.f, 0.f};
      const uint32_t si_pf = s_cur;
      s_cur = s_nxt;
      s_nxt = i + 2 * blockDim.x < end ? src_of[i + 2 * blockDim.x] : 0u;
      if (i + blockDim.x < end) {
#pragma unroll
        for (int q = 0; q < 6; ++q) prefetch_l2(P.q(q) + s_cur);
      }
      if (i >= start && i < end) {
      const uint32_t si = si_pf;
      // quads 0..5 of the particle (svb_device.cuh: Field): position, flags, F and the words this thread merely carries — six 16-byte
      // gathers instead of thirty 4-byte ones (v and C are replaced, quads 6..8 are not read)
      const float4 pq0 = P.q(0)[si], pq1 = P.q(1)[si], pq2 = P.q(2)[si], pq3 = P.q(3)[si], pq4 = P.q(4)[si];
      const float2 pq5 = *reinterpret_cast<const float2*>(P.q(5) + si);
      V3 x = V3{pq0.x, pq0.y, pq0.z};
      uint32_t flags = __float_as_uint(pq0.w);
      M3 F;
      F.m[0] = pq1.x; F.m[1] = pq1.y; F.m[2] = pq1.z; F.m[3] = pq1.w; F.m[4] = pq2.x; F.m[5] = pq2.y; F.m[6] = pq2.z; F.m[7] = pq2.w; F.m[8] = pq3.x;
      const float carry[7] = {pq3.y, pq3.z, pq3.w, pq4.x, pq4.y, pq4.z, pq4.w};   // mass, V0, mu | K, lambda | gamma, alpha, eta, zeta
      // (deriving the base node from the binned cell id instead of three IEEE divisions was measured slower:
      // the extra gathered 4-byte load costs more than the divisions)
      const V3 nrm = V3{__fdiv_rn(x.x, h), __fdiv_rn(x.y, h), __fdiv_rn(x.z, h)};
      const V3 shiftf = V3{floorf(__fsub_rn(nrm.x, 0.5f)), floorf(__fsub_rn(nrm.y, 0.5f)), floorf(__fsub_rn(nrm.z, 0.5f))};
      const int s0 = (int)shiftf.x, s1 = (int)shiftf.y, s2 = (int)shiftf.z;
      const V3 shifted = nrm - shiftf;
      float wx[3], wy[3], wz[3], dx[3], dy[3], dz[3];
      quad_weights(shifted.x, wx); quad_weights(shifted.y, wy); quad_weights(shifted.z, wz);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        dx[a] = (float)(s0 + a) * h - x.x;   // collect_velocity.rs:52-53: node position - particle position
        dy[a] = (float)(s1 + a) * h - x.y;
        dz[a] = (float)(s2 + a) * h - x.z;
      }
      // v = sum w v_n ; C = sum (w v_n) (x_n - x)^T, factored along z and evaluated as packed fp32 pairs (FFMA2):
      //   txy = sum_c wz_c v_c.xy          zxy = sum_c (wz_c dz_c) v_c.xy          tz = sum_c (wz_c, wz_c dz_c) v_c.z
      // so a stencil column costs 9 + 8 packed instructions instead of 33 scalar ones.
      const float2 wzp[3] = {make_float2(wz[0], wz[0] * dz[0]), make_float2(wz[1], wz[1] * dz[1]), make_float2(wz[2], wz[2] * dz[2])};
      float2 vxy = make_float2(0.f, 0.f), vz_c8 = make_float2(0.f, 0.f);              // (v.x, v.y), (v.z, C22)
      float2 c01 = make_float2(0.f, 0.f), c34 = make_float2(0.f, 0.f), c67 = make_float2(0.f, 0.f), c25 = make_float2(0.f, 0.f);
      const int tb = ((s0 & 3) * 6 + (s1 & 3)) * 6 + (s2 & 3);
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          const float4 g0 = tile[tb + (a * 6 + b) * 6], g1 = tile[tb + (a * 6 + b) * 6 + 1], g2 = tile[tb + (a * 6 + b) * 6 + 2];
          const float2 txy = ffma2(make_float2(g2.x, g2.y), wzp[2].x, ffma2(make_float2(g1.x, g1.y), wzp[1].x, fmul2(make_float2(g0.x, g0.y), wzp[0].x)));
          const float2 zxy = ffma2(make_float2(g2.x, g2.y), wzp[2].y, ffma2(make_float2(g1.x, g1.y), wzp[1].y, fmul2(make_float2(g0.x, g0.y), wzp[0].y)));
          const float2 tz = ffma2(wzp[2], g2.z, ffma2(wzp[1], g1.z, fmul2(wzp[0], g0.z)));   // (sum wz v.z, sum wz dz v.z)
          const float wxy = wx[a] * wy[b];
          const float2 wd = fmul2(make_float2(dx[a], dy[b]), wxy);                           // (w dx, w dy)
          vxy = ffma2(txy, wxy, vxy);
          vz_c8 = ffma2(tz, wxy, vz_c8);
          c01 = ffma2(txy, wd.x, c01);
          c34 = ffma2(txy, wd.y, c34);
          c67 = ffma2(zxy, wxy, c67);
          c25 = ffma2(wd, tz.x, c25);
        }
      V3 v = V3{vxy.x, vxy.y, vz_c8.x};
      M3 C;
      C.m[0] = c01.x; C.m[1] = c01.y; C.m[2] = c25.x;
      C.m[3] = c34.x; C.m[4] = c34.y; C.m[5] = c25.y;
      C.m[6] = c67.x; C.m[7] = c67.y; C.m[8] = vz_c8.y;
      const float cs = 4.f * inv_h * inv_h;
#pragma unroll
      for (int q = 0; q < 9; ++q) C.m[q] *= cs;
      if (REDUCE) {
        red_vel = max(red_vel, total_key(norm(v)));
#pragma unroll
        for (int q = 0; q < 9; ++q) red_def = min(red_def, total_key(0.2f / fmaxf(fabsf(C.m[q]), 1e-8f)));
      }
      if (FUSE) {
        // advance_particles.rs:41-86, cull_particles.rs:31-39
        x = x + v * dt;
        const M3 CF = mul(C, F);
#pragma unroll
        for (int q = 0; q < 9; ++q) F.m[q] += CF.m[q] * dt;
        float e;
        if (return_map_and_energy(flags, carry[PP0 - PMASS], carry[PP1 - PMASS], (flags & F_USE_SAND_ALPHA) ? carry[PALPHA - PMASS] : 0.f, F, e)) energy[i] = e;
        else { flags |= F_FAILED; failed = 1; }
        const bool within = x.x > K.domain_min[0] && x.x < K.domain_max[0] && x.y > K.domain_min[1] && x.y < K.domain_max[1] && x.z > K.domain_min[2] && x.z < K.domain_max[2];
        if (!within) flags |= F_TOMBSTONED;
        bool leaving = false;
        // cheap filter first (x * (1/h) is within a small fraction of a cell of the exact x / h): only particles within half a
        // cell of a cut evaluate the bit-exact base node
        const float approx_node = x.x * inv_h - 0.5f;
        if (SLAB && within && (approx_node < mc.node_lo + 0.5f || approx_node > mc.node_hi - 0.5f)) {
          const int bx = floor_div4(base_node(x.x, h));
          if (bx < mc.lo || bx >= mc.hi) {
            if (bx < mc.reach_lo || bx >= mc.reach_hi) atomicOr(&S->status, ST_KEY_RANGE);   // crossed more than one slab in a substep
            else {
              const int side = bx < mc.lo ? 0 : 1;
              const uint32_t slot = atomicAdd(&mc.counts[side], 1u);
              if (slot < mc.cap) mc.list[(size_t)side * mc.cap + slot] = i;
              leaving = true;
            }
          }
        }
        bin_state = leaving ? 2 : ((flags & F_TOMBSTONED) ? 1 : 0);
        bin_x = x;
      }
      D.q(0)[i] = make_float4(x.x, x.y, x.z, __uint_as_float(flags));
      D.q(1)[i] = make_float4(F.m[0], F.m[1], F.m[2], F.m[3]);
      D.q(2)[i] = make_float4(F.m[4], F.m[5], F.m[6], F.m[7]);
      D.q(3)[i] = make_float4(F.m[8], carry[0], carry[1], carry[2]);
      D.q(4)[i] = make_float4(carry[3], carry[4], carry[5], carry[6]);
      D.q(5)[i] = make_float4(pq5.x, pq5.y, v.x, v.y);   // collider bits, original index
      D.q(6)[i] = make_float4(v.z, C.m[0], C.m[1], C.m[2]);
      D.q(7)[i] = make_float4(C.m[3], C.m[4], C.m[5], C.m[6]);
      D.q(8)[i] = make_float4(C.m[7], C.m[8], 0.f, 0.f);
      }
      if (BIN && i >= start && i < end) {
        uint32_t ci = bin_state == 1 ? 0xffffffffu : 0xfffffffdu;
        if (bin_state == 1) atomicAdd(&bn.S->n_tomb, 1u);   // culled just now (rare)
        if (bin_state == 0) {
          const int s0 = base_node_fast(bin_x.x, h, inv_h), s1 = base_node_fast(bin_x.y, h, inv_h), s2 = base_node_fast(bin_x.z, h, inv_h);
          const int b0 = floor_div4(s0), b1 = floor_div4(s1), b2 = floor_div4(s2);
          const uint32_t cell = ((uint32_t)(s0 & 3) << 4) | ((uint32_t)(s1 & 3) << 2) | (uint32_t)(s2 & 3);
          uint32_t touch = 1u;   // neighbour offsets reached by the 3-node stencil: +1 on an axis iff the in-block coordinate >= 2
          if (s0 & 2) touch |= touch << 1;
          if (s1 & 2) touch |= touch << 2;
          if (s2 & 2) touch |= touch << 4;
          const int d0 = b0 - s_block[0] + 1, d1 = b1 - s_block[1] + 1, d2 = b2 - s_block[2] + 1;
          const int lim = BLOCK_BIAS - 2;
          ci = 0xfffffffeu;
          if (b0 < -lim || b0 > lim || b1 < -lim || b1 > lim || b2 < -lim || b2 > lim) atomicOr(&bn.S->status, ST_KEY_RANGE);
          else if ((unsigned)d0 < 3u && (unsigned)d1 < 3u && (unsigned)d2 < 3u) {
            const int at = (d0 * 3 + d1) * 3 + d2;
            // (deliberately unsynchronised: every thread that finds ~0u looks the tile up itself and all of them store the same slot —
            //  compute-sanitizer's racecheck reports this read / write pair, profiles/README.md)
            uint32_t slot = ((volatile uint32_t*)s_cache)[at];
            if (slot == ~0u) {   // the first particle(s) of this tile that reach the block look its tile up (or create it)
              slot = tile_slot_find_or_insert(bn.T, tile_key_pack(b0, b1, b2, 0u), bn.S);
              if (slot != ~0u) ((volatile uint32_t*)s_cache)[at] = slot;
            }
            if (slot != ~0u) {
              atomicAdd(&s_cnt[at * 64 + cell], 1u);
              atomicOr(&s_touch[at], touch);
              ci = slot * 64u + cell;
            }
          } else {   // more than a block away (no CFL-respecting step does this): straight to the tables
            const uint32_t slot = tile_slot_find_or_insert(bn.T, tile_key_pack(b0, b1, b2, 0u), bn.S);
            if (slot != ~0u) {
              atomicAdd(&bn.B.cell_count[(size_t)slot * 64 + cell], 1u);
              atomicOr(&bn.B.tile_touch[slot], touch);
              ci = slot * 64u + cell;
            }
          }
        }
        bn.B.pcell[i] = ci;
      }
    }
  }
  // tombstoned particles take no part in P2G / G2P: carry their rows over as they are (behind the live ones)
  if (W.tail) {
    const uint32_t n_live = S->n_live, n_end = n_live + S->n_tomb;  // rows of migrated particles are gone
    for (uint32_t j = n_live + blockIdx.x * blockDim.x + threadIdx.x; j < n_end; j += gridDim.x * blockDim.x) {
      const uint32_t i = src_of[j];
#pragma unroll 1
      for (int q = 0; q < NQUADS; ++q) D.q(q)[j] = P.q(q)[i];
      if (BIN) {   // stays tombstoned: behind the live rows of the next substep as well
        bn.B.pcell[j] = 0xffffffffu;
        atomicAdd(&bn.S->n_tomb, 1u);
      }
    }
  }
  if (REDUCE) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      red_vel = max(red_vel, __shfl_xor_sync(SVB_FULL, red_vel, o));
      red_def = min(red_def, __shfl_xor_sync(SVB_FULL, red_def, o));
    }
    if ((threadIdx.x & 31) == 0) {
      if (red_vel != INT32_MIN) atomicMax(&S->max_velocity_key, red_vel);
      if (red_def != INT32_MAX) atomicMin(&S->min_deformation_key, red_def);
    }
  }
  if (FUSE && failed) atomicOr(&S->sticky_new, 8u /*SVB_PARTICLE_CLOSE_TO_INVERTED*/);
}

// ---- peer-memory mailbox header and system-scope accessors (used by the slab exchanges below and by the time-step kernels)
constexpr int SLAB_MAX_RANKS = 16;
constexpr uint32_t SLAB_OVERFLOW = 0x80000000u;   // count word: the sender ran out of mailbox capacity
struct SlabHeader {
  uint32_t halo_seq[2], halo_count[2];   // [0] written by the left neighbour, [1] by the right one
  uint32_t mig_seq[2], mig_count[2];
  uint32_t err_seq[SLAB_MAX_RANKS], err_val[SLAB_MAX_RANKS];  // every rank posts its sticky error word to every rank
  // adaptive time steps: every rank posts its limit reductions to every rank; two slots alternate (an exchange may only overwrite
  // the slot of the exchange before last, which every rank has provably finished reading)
  uint32_t dt_seq[2][SLAB_MAX_RANKS];
  int32_t dt_val[2][SLAB_MAX_RANKS][4];
};
__device__ __forceinline__ uint32_t ld_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
// true when *p reached `seq`; gives up after ~2 s so a lost neighbour cannot hang the GPU
__device__ __forceinline__ bool wait_seq(const uint32_t* p, uint32_t seq) {
  for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
    if ((int32_t)(ld_sys(p) - seq) >= 0) return true;
    __nanosleep(spin < 64 ? 32 : 400);
  }
  return false;
}

// slab ranks: the limit reductions of adaptive time stepping are global minima / maxima over all ranks.  Every rank posts its three
// words to every rank and waits for everybody else's; min / max are order-independent, so all ranks compute the same step.
struct DtPeers {
  int rank, n_ranks;                       // n_ranks <= 1: single device, nothing to exchange
  uint32_t* post_seq[SLAB_MAX_RANKS];      // [r]: &header_of_rank_r.dt_seq[0][rank]   (null for r == rank)
  int32_t* post_val[SLAB_MAX_RANKS];       // [r]: header_of_rank_r.dt_val[0][rank]
  const SlabHeader* mine;
  uint32_t exchange;                       // number of this exchange (same on every rank, strictly increasing)
};
// combine (a: min, b: min, c: max, d: max) over all ranks; false on a timeout
__device__ __forceinline__ bool dt_exchange(const DtPeers& P, int32_t& a, int32_t& b, int32_t& c, int32_t& d) {
  if (P.n_ranks <= 1) return true;
  const uint32_t slot = P.exchange & 1u, stride_seq = SLAB_MAX_RANKS, stride_val = SLAB_MAX_RANKS * 4;
  for (int r = 0; r < P.n_ranks; ++r)
    if (r != P.rank) {
      int32_t* v = P.post_val[r] + slot * stride_val;
      asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(v), "r"(a) : "memory");
      asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(v + 1), "r"(b) : "memory");
      asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(v + 2), "r"(c) : "memory");
      asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(v + 3), "r"(d) : "memory");
      st_sys(P.post_seq[r] + slot * stride_seq, P.exchange);
    }
  bool ok = true;
  const unsigned long long t0 = global_timer_ns();
  for (int r = 0; r < P.n_ranks; ++r)
    if (r != P.rank) {
      if (!wait_seq(&P.mine->dt_seq[slot][r], P.exchange)) { ok = false; continue; }
      const int32_t* v = P.mine->dt_val[slot][r];
      a = min(a, (int32_t)ld_sys((const uint32_t*)v));
      b = min(b, (int32_t)ld_sys((const uint32_t*)(v + 1)));
      c = max(c, (int32_t)ld_sys((const uint32_t*)(v + 2)));
      d = max(d, (int32_t)ld_sys((const uint32_t*)(v + 3)));
    }
  atomicAdd(&g_wait_ns[5], global_timer_ns() - t0);
  return ok;
}

// ------------------------------------------------------------------------------------------------
// adaptive time stepping with the state machine on the device (DtState, svb_device.cuh)
__device__ __forceinline__ float dt_total_min(float a, float b) { return total_key(b) < total_key(a) ? b : a; }   // min_by(f32::total_cmp) keeps the first minimum
__device__ __forceinline__ float dt_allowed_without_prior(const DtState& d) {   // adaptive_time_step_state.rs:36-47
  const float fmax = 3.402823466e+38f;
  float r = d.max_dt;
  r = dt_total_min(r, (d.has & 1u) ? d.by_velocity : fmax);
  r = dt_total_min(r, (d.has & 2u) ? d.by_deformation : fmax);
  r = dt_total_min(r, (d.has & 8u) ? d.by_sound : fmax);
  r = dt_total_min(r, (d.has & 4u) ? d.by_isolated : fmax);
  return r;
}
__device__ __forceinline__ float dt_allowed(const DtState& d) {   // :49-54
  float r = dt_allowed_without_prior(d);
  for (uint32_t q = 0; q < d.prior_len; ++q) r = dt_total_min(r, d.prior[q]);
  return r;
}
__device__ __forceinline__ void dt_push(DtState& d) {   // :56-63 (pop when len > 10, then push: at most 11 entries)
  if (d.prior_len > 10) {
    for (uint32_t q = 1; q < d.prior_len; ++q) d.prior[q - 1] = d.prior[q];
    --d.prior_len;
  }
  d.prior[d.prior_len++] = dt_allowed_without_prior(d);
}
// frame factor + gravity of the substep that starts at d.time; false when the loaded frame is the wrong one
__device__ __forceinline__ bool dt_interpolate_input(DtState& d) {
  const double frame_time = d.time * d.fps;
  if ((unsigned long long)floor(frame_time) != d.frame) return false;
  d.factor_b = (float)fmod(frame_time, 1.0);
  const float fa = 1.f - d.factor_b;
  for (int k = 0; k < 3; ++k) d.g[k] = fa * d.ga[k] + d.factor_b * d.gb[k];
  return true;
}
// LimitTimeStepBeforeForce (limit_time_step.rs:25-33) from the reductions in next_*; resets them for the next round.
// false: a slab neighbour did not answer
__device__ __forceinline__ bool dt_limit_before_force(DtState& d, const DtPeers& peers) {
  int32_t live = d.next_live > 0 ? 1 : 0, unused = 0;
  const bool ok = dt_exchange(peers, d.next_min_sound_key, d.next_min_isolated_key, live, unused);
  const bool any = live > 0;
  d.has = (d.has & ~12u) | (any ? 12u : 0u);
  if (any) {
    d.by_sound = total_unkey(d.next_min_sound_key);
    d.by_isolated = total_unkey(d.next_min_isolated_key);
  }
  dt_push(d);
  d.allowed = dt_allowed(d);
  d.next_min_sound_key = INT32_MAX; d.next_min_isolated_key = INT32_MAX; d.next_live = 0;
  return ok;
}
// Start of an svb_advance call (one thread).  The host has written time / target / max_dt / fps / frame / gravities and cleared `stop`.
// `S` = the scalars of the substep about to run (its sticky words were cleared by the host).
__global__ void k_dt_open(DtState* D, StepScalars* S, DtPeers peers) {
  DtState d = *D;
  uint32_t stop = 0;
  d.substeps = 0;
  d.allowed = dt_allowed(d);                              // max_time_step may have changed (cpu_state.rs:164)
  if (d.allowed == 0.f) stop |= ST_STOP_ZERO_DT;          // cpu_state.rs:170-172
  if (!(d.time < d.target)) stop |= ST_STOP_DONE;
  else if (!dt_interpolate_input(d)) stop |= ST_STOP_FRAME;
  d.dt_force = d.allowed;
  if (!stop) {   // (every rank takes the same branch: time, target, frame and the step history are the same everywhere)
    if (!dt_limit_before_force(d, peers)) atomicOr(&S->status, ST_COMM_TIMEOUT);
    if (d.allowed == 0.f) stop |= ST_STOP_ZERO_DT;
  }
  d.stop = stop;
  *D = d;
  if (stop) S->sticky |= stop;   // no kernel of this substep has started yet: the whole substep is a no-op
}
// LimitTimeStepBeforeIntegrate (limit_time_step.rs:187-223) from the G2P reductions of this substep
__global__ void k_dt_integrate(DtState* D, StepScalars* S, DtPeers peers) {
  if (SVB_ABORTED(S)) return;
  DtState d = *D;
  // global over all slab ranks: min deformation limit, max velocity, any live particle anywhere
  int32_t def_key = S->min_deformation_key, unused = INT32_MAX, vel_key = S->max_velocity_key, live = S->n_live > 0 ? 1 : 0;
  if (!dt_exchange(peers, def_key, unused, vel_key, live)) atomicOr(&S->status, ST_COMM_TIMEOUT);
  const bool any = live > 0;
  const float max_vel = any && vel_key != INT32_MIN ? total_unkey(vel_key) : 0.f;
  const bool has_vel = any && max_vel != 0.f;
  d.has = (d.has & ~3u) | (has_vel ? 1u : 0u) | (any ? 2u : 0u);
  if (has_vel) d.by_velocity = 0.5f * d.h / max_vel;
  if (any) d.by_deformation = total_unkey(def_key);
  dt_push(d);
  d.allowed = dt_allowed(d);
  *D = d;
  if (d.allowed == 0.f) { D->stop |= ST_STOP_ZERO_DT; atomicOr(&S->status, ST_ZERO_DT); }   // the advance of this substep must not run
}
// End of a substep: the clock moves on (cpu_state.rs:187-190), then everything the NEXT substep needs before its first kernel:
// the run-is-over test, the frame factor and gravity, and LimitTimeStepBeforeForce from the limits k_advance reduced on the advanced F.
__global__ void k_dt_tail(DtState* D, StepScalars* S, DtPeers peers) {
  if (SVB_ABORTED(S)) return;
  if (S->sticky_new & 0xffffu) return;   // a FAILED particle: the reference returns from the Advance phase without moving the clock (cpu_state.rs:176-184)
  DtState d = *D;
  d.time += (double)d.allowed;
  ++d.substeps;
  uint32_t stop = 0;
  if (d.allowed == 0.f) stop |= ST_STOP_ZERO_DT;
  if (!(d.time < d.target)) stop |= ST_STOP_DONE;
  else if (!dt_interpolate_input(d)) stop |= ST_STOP_FRAME;
  d.dt_force = d.allowed;
  if (!stop) {
    if (!dt_limit_before_force(d, peers)) atomicOr(&S->status, ST_COMM_TIMEOUT);
    if (d.allowed == 0.f) stop |= ST_STOP_ZERO_DT;
  }
  d.stop |= stop;
  *D = d;
  if (stop) atomicOr(&S->sticky_new, stop);
}

// advance + cull as its own pass (adaptive time stepping: dt is only known after the G2P reductions).  Also — it holds the advanced
// position and F — the binning of the next substep (BIN, scenes without a collider mesh) and the per-particle limits of the next
// substep's LimitTimeStepBeforeForce (limit_time_step.rs:35-182).  One thread per row of the binned buffer, tombstoned rows included.
#ifndef SVB_ADVANCE_BLOCKS
#define SVB_ADVANCE_BLOCKS 3   // 80 registers since the particle words come in quads (2 blocks / SM with the word layout)
#endif
template <bool BIN>
__global__ void __launch_bounds__(256, SVB_ADVANCE_BLOCKS) k_advance(ParticleBuf P, float* __restrict__ energy, StepScalars* S, SimConsts K, uint32_t n, DtState* D, BinNext bn, MigrateCut mc) {
  if (SVB_ABORTED(S)) return;
  n = min(n, S->n_live + S->n_tomb);
  const float dt = D->allowed;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31;
  int bin_state = 2;
  V3 x = V3{0.f, 0.f, 0.f};
  int ks = INT32_MAX, ki = INT32_MAX;
  uint32_t live = 0;
  if (i < n) {
    const float4 pq0 = P.q(0)[i];   // (every quad of the row is requested before the flags are looked at)
    const float4 pq1 = P.q(1)[i], pq2 = P.q(2)[i], pq3 = P.q(3)[i], pq4 = P.q(4)[i], pq5 = P.q(5)[i], pq6 = P.q(6)[i], pq7 = P.q(7)[i];
    const float2 pq8 = *reinterpret_cast<const float2*>(P.q(8) + i);
    uint32_t flags = __float_as_uint(pq0.w);
    bin_state = 1;
    if (!(flags & F_TOMBSTONED)) {
      x = V3{pq0.x, pq0.y, pq0.z};
      const V3 v = V3{pq5.z, pq5.w, pq6.x};
      M3 C, F;
      F.m[0] = pq1.x; F.m[1] = pq1.y; F.m[2] = pq1.z; F.m[3] = pq1.w; F.m[4] = pq2.x; F.m[5] = pq2.y; F.m[6] = pq2.z; F.m[7] = pq2.w; F.m[8] = pq3.x;
      C.m[0] = pq6.y; C.m[1] = pq6.z; C.m[2] = pq6.w; C.m[3] = pq7.x; C.m[4] = pq7.y; C.m[5] = pq7.z; C.m[6] = pq7.w; C.m[7] = pq8.x; C.m[8] = pq8.y;
      const float p0 = pq3.w, p1 = pq4.x;
      x = x + v * dt;
      const M3 CF = mul(C, F);
#pragma unroll
      for (int q = 0; q < 9; ++q) F.m[q] += CF.m[q] * dt;
      float e;
      if (return_map_and_energy(flags, p0, p1, (flags & F_USE_SAND_ALPHA) ? pq4.y : 0.f, F, e)) energy[i] = e;
      else { flags |= F_FAILED; atomicOr(&S->sticky_new, 8u); }
      const bool within = x.x > K.domain_min[0] && x.x < K.domain_max[0] && x.y > K.domain_min[1] && x.y < K.domain_max[1] && x.z > K.domain_min[2] && x.z < K.domain_max[2];
      if (!within) flags |= F_TOMBSTONED;
      P.q(0)[i] = make_float4(x.x, x.y, x.z, __uint_as_float(flags));
      P.q(1)[i] = make_float4(F.m[0], F.m[1], F.m[2], F.m[3]);
      P.q(2)[i] = make_float4(F.m[4], F.m[5], F.m[6], F.m[7]);
      P.f(PF + 8)[i] = F.m[8];
      if (within) {
        bin_state = 0;
        if (mc.list) {   // slab ranks: note the particles whose advanced position left the rank's block columns (like the fused G2P)
          const float approx_node = x.x * (1.f / K.h) - 0.5f;
          if (approx_node < mc.node_lo + 0.5f || approx_node > mc.node_hi - 0.5f) {
            const int bx = floor_div4(base_node(x.x, K.h));
            if (bx < mc.lo || bx >= mc.hi) {
              if (bx < mc.reach_lo || bx >= mc.reach_hi) atomicOr(&S->status, ST_KEY_RANGE);   // crossed more than one slab in a substep
              else {
                const int side = bx < mc.lo ? 0 : 1;
                const uint32_t slot = atomicAdd(&mc.counts[side], 1u);
                if (slot < mc.cap) mc.list[(size_t)side * mc.cap + slot] = i;
                bin_state = 2;
              }
            }
          }
        }
        const ParticleLimits l = particle_time_step_limits((flags & F_IS_FLUID) != 0, p0, p1, pq3.y, pq3.z, F, K.h);
        ks = total_key(l.by_sound);
        ki = total_key(l.by_isolated);
        live = 1;
      }
    }
  }
  if (BIN) bin_warp<false, false>(bin_state, x, 0u, i < n, i, K, bn.T, bn.B, bn.S, lane);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ks = min(ks, __shfl_xor_sync(SVB_FULL, ks, o));
    ki = min(ki, __shfl_xor_sync(SVB_FULL, ki, o));
    live += __shfl_xor_sync(SVB_FULL, live, o);
  }
  if (lane == 0 && live) {
    atomicMin(&D->next_min_sound_key, ks);
    atomicMin(&D->next_min_isolated_key, ki);
    atomicAdd(&D->next_live, live);
  }
}

// limit_time_step.rs:25-182 as its own pass (the first substep after an upload: no k_advance has reduced the limits yet):
// global minima of the sound-speed and isolated-particle bounds into DtState::next_*
__global__ void __launch_bounds__(256) k_limit_force(ParticleBuf P, DtState* D, float h, uint32_t n, const uint32_t* __restrict__ n_dev) {
  if (n_dev) n = min(n, *n_dev);
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  int ks = INT32_MAX, ki = INT32_MAX;
  uint32_t live = 0;
  if (i < n) {
    const uint32_t flags = P.u(PFLAGS)[i];
    if (!(flags & (F_TOMBSTONED | F_GONE))) {
      const float4 pq1 = P.q(1)[i], pq2 = P.q(2)[i], pq3 = P.q(3)[i];
      M3 F;
      F.m[0] = pq1.x; F.m[1] = pq1.y; F.m[2] = pq1.z; F.m[3] = pq1.w; F.m[4] = pq2.x; F.m[5] = pq2.y; F.m[6] = pq2.z; F.m[7] = pq2.w; F.m[8] = pq3.x;
      const ParticleLimits l = particle_time_step_limits((flags & F_IS_FLUID) != 0, pq3.w, P.f(PP1)[i], pq3.y, pq3.z, F, h);
      ks = total_key(l.by_sound);
      ki = total_key(l.by_isolated);
      live = 1;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ks = min(ks, __shfl_xor_sync(SVB_FULL, ks, o));
    ki = min(ki, __shfl_xor_sync(SVB_FULL, ki, o));
    live += __shfl_xor_sync(SVB_FULL, live, o);
  }
  if ((threadIdx.x & 31) == 0 && live) {
    atomicMin(&D->next_min_sound_key, ks);
    atomicMin(&D->next_min_isolated_key, ki);
    atomicAdd(&D->next_live, live);
  }
}

// progress markers of the exchange kernels (message numbers), read by the host when an exchange times out (diagnostics only)
__device__ uint32_t g_exchange_trace[16];
#define SVB_TRACE(k, v) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_exchange_trace[k] = (v); } while (0)

// ------------------------------------------------------------------------------------------------
// multi-GPU slabs (no reference counterpart).  A rank owns the particles whose base
// node lies in block columns [lo, hi) along x.
//   halo: after P2G the tiles of one block column are packed as (block key, collider bits, 64 x float4),
//   exchanged with the neighbour rank and added into (or created in) the receiver's tile table;
//   migration: after the advance, particles whose block column left [lo, hi) are packed as 35-word
//   rows, flagged F_GONE locally (dropped by the next re-bin) and appended on the neighbour.
struct HaloEntry {
  unsigned long long block_key;  // tile key with the layer field cleared
  uint32_t bits;                 // collider bits of the layer (layer ids are rank-local)
  uint32_t pad;
  float4 node[64];
};
__global__ void __launch_bounds__(256) k_pack_column(const StepScalars* __restrict__ S, TileTable T, const unsigned long long* __restrict__ layer_slots, const float4* __restrict__ grid,
                                                     int column, HaloEntry* __restrict__ out, uint32_t* __restrict__ count, uint32_t cap) {
  if (SVB_ABORTED(S)) return;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t n_tiles = min(S->n_tiles, T.tile_cap);
  for (uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n_tiles; t += warps) {
    const unsigned long long k = T.tile_key[t];
    int bx, by, bz;
    uint32_t layer;
    tile_key_unpack(k, bx, by, bz, layer);
    if (bx != column) continue;
    uint32_t slot = 0;
    if (lane == 0) slot = atomicAdd(count, 1u);
    slot = __shfl_sync(SVB_FULL, slot, 0);
    if (slot >= cap) continue;  // the host sees count > cap and grows the buffers
    HaloEntry* e = out + slot;
    if (lane == 0) {
      e->block_key = k & ~(unsigned long long)((1u << LAYER_BITS) - 1u);
      e->bits = layer_bits_of(layer_slots, layer);
      e->pad = 0;
    }
    e->node[lane] = grid[(size_t)t * 64 + lane];
    e->node[lane + 32] = grid[(size_t)t * 64 + lane + 32];
  }
}
__global__ void __launch_bounds__(256) k_unpack_add(StepScalars* S, TileTable T, unsigned long long* layer_slots, uint32_t* layer_list, float4* __restrict__ grid,
                                                    const HaloEntry* __restrict__ in, uint32_t count) {
  if (SVB_ABORTED(S)) return;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t zeroed = S->n_tiles_zeroed;
  for (uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < count; q += warps) {
    const HaloEntry* e = in + q;
    uint32_t id = TILE_PENDING;
    if (lane == 0) {
      const uint32_t layer = layer_find_or_insert(layer_slots, layer_list, e->bits, S);
      id = tile_find_or_insert(T, e->block_key | layer, S);
    }
    id = __shfl_sync(SVB_FULL, id, 0);
    if (id == TILE_PENDING) continue;
    float4* dst = grid + (size_t)id * 64;
    if (id >= zeroed) {  // created by this message: the tile was never zeroed, so store instead of add
      dst[lane] = e->node[lane];
      dst[lane + 32] = e->node[lane + 32];
    } else {
      red_add_v4(dst + lane, e->node[lane]);
      red_add_v4(dst + lane + 32, e->node[lane + 32]);
    }
  }
}
// ---- the same exchanges over peer memory (NVLink loads/stores into the neighbour's HBM, no NCCL call and no
// host round trip on the data path).  Every rank owns a mailbox that its neighbours map with CUDA IPC:
//   header | halo entries from the left | from the right | migrating rows from the left | from the right
// A sender writes its payload straight into the neighbour's mailbox; the last block of the sending kernel
// publishes the count and then the substep sequence number with system-scope fences.  The receiving kernel
// spins (bounded) on the sequence number in its own HBM.  Buffers are reused safely because a rank can only
// send message k+1 after it has consumed the neighbour's message k of the other kind.
// both directions in one launch: tiles of my first column go to the left neighbour, tiles of my halo column (== hi) to the
// right one; `local` = [0], [1] slot counters, [2] blocks done
struct HaloPeers {
  HaloEntry* entries[2];                 // the neighbours' mailbox regions for my messages (null at the domain ends)
  uint32_t* count[2]; uint32_t* seq[2];
};
__global__ void __launch_bounds__(256) k_halo_send2(const StepScalars* __restrict__ S, TileTable T, const unsigned long long* __restrict__ layer_slots, const float4* __restrict__ grid, int lo, int hi,
                                                    HaloPeers peers, uint32_t cap, uint32_t seq, uint32_t* __restrict__ local, const uint32_t* done, const uint32_t* n_boundary, uint32_t parts) {
  // a run that was stopped by an earlier substep (sticky is the same word on every rank: error words are exchanged, stop bits come
  // from identical clocks) exchanges nothing: every rank skips the same messages, however many no-op substeps its host queued
  if (S->sticky) return;
  SVB_TRACE(0, seq);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  if (done && !SVB_ABORTED(S)) {
    // launched next to P2G on a second stream: the halo columns are complete once every boundary tile has been scattered
    __shared__ int s_ok;
    if (threadIdx.x == 0) { WaitClock wc(0); s_ok = wait_boundary_done(done, n_boundary, parts) ? 1 : 0; wc.stop(); }
    __syncthreads();
    if (!s_ok && threadIdx.x == 0) atomicOr(const_cast<uint32_t*>(&S->status), ST_COMM_TIMEOUT);
  }
  SVB_TRACE(1, seq);
  if (!SVB_ABORTED(S)) {
    const uint32_t n_tiles = min(S->n_tiles, T.tile_cap);
    for (uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n_tiles; t += warps) {
      const unsigned long long k = T.tile_key[t];
      int bx, by, bz;
      uint32_t layer;
      tile_key_unpack(k, bx, by, bz, layer);
      int side;
      if (bx == lo && peers.entries[0]) side = 0;
      else if (bx == hi && peers.entries[1]) side = 1;
      else continue;
      uint32_t slot = 0;
      if (lane == 0) slot = atomicAdd(&local[side], 1u);
      slot = __shfl_sync(SVB_FULL, slot, 0);
      if (slot >= cap) continue;
      HaloEntry* e = peers.entries[side] + slot;
      if (lane == 0) {
        e->block_key = k & ~(unsigned long long)((1u << LAYER_BITS) - 1u);
        e->bits = layer_bits_of(layer_slots, layer);
        e->pad = 0;
      }
      e->node[lane] = grid[(size_t)t * 64 + lane];
      e->node[lane + 32] = grid[(size_t)t * 64 + lane + 32];
    }
  }
  // the block's peer-memory writes are ordered before its "done" tick by the barrier plus one system-scope fence of the
  // ticking thread (fence cumulativity), not by a fence per thread
  __syncthreads();
  if (threadIdx.x == 0) __threadfence_system();
  if (threadIdx.x == 0 && atomicAdd(&local[2], 1u) == gridDim.x - 1) {
    for (int side = 0; side < 2; ++side) {
      const uint32_t c = atomicAdd(&local[side], 0u);
      if (peers.count[side]) {
        st_sys(peers.count[side], c > cap ? (cap | SLAB_OVERFLOW) : c);
        st_sys(peers.seq[side], seq);
      }
      local[side] = 0;
    }
    local[2] = 0;
    g_exchange_trace[2] = seq;
  }
}
__global__ void __launch_bounds__(256) k_halo_recv2(StepScalars* S, TileTable T, unsigned long long* layer_slots, uint32_t* layer_list, float4* __restrict__ grid, const HaloEntry* __restrict__ in_left,
                                                    const HaloEntry* __restrict__ in_right, const SlabHeader* __restrict__ hdr, int has_left, int has_right, uint32_t seq) {
  __shared__ uint32_t s_count[2];
  if (S->sticky) return;   // stopped run: nothing was sent (see k_halo_send2)
  SVB_TRACE(3, seq);
  if (threadIdx.x == 0) {
    const int has[2] = {has_left, has_right};
    WaitClock wc(1);
    for (int side = 0; side < 2; ++side) {
      uint32_t c = 0;
      if (has[side]) {
        if (wait_seq(&hdr->halo_seq[side], seq)) {
          c = ld_sys(&hdr->halo_count[side]);
          if (c & SLAB_OVERFLOW) { atomicOr(&S->status, ST_COMM_OVERFLOW); c = 0; }
        } else atomicOr(&S->status, ST_COMM_TIMEOUT);   // in the abort mask: the rest of the substep no-ops
      }
      s_count[side] = c;
    }
    wc.stop();
  }
  __syncthreads();
  const uint32_t cl = s_count[0], total = cl + s_count[1];
  SVB_TRACE(4, seq);
  if (SVB_ABORTED(S)) return;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t zeroed = S->n_tiles_zeroed;
  for (uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < total; q += warps) {
    const HaloEntry* e = q < cl ? in_left + q : in_right + (q - cl);
    uint32_t id = TILE_PENDING;
    if (lane == 0) {
      const uint32_t layer = layer_find_or_insert(layer_slots, layer_list, e->bits, S);
      id = tile_find_or_insert(T, e->block_key | layer, S);
    }
    id = __shfl_sync(SVB_FULL, id, 0);
    if (id == TILE_PENDING) continue;
    float4* dst = grid + (size_t)id * 64;
    if (id >= zeroed) {  // created by this message: the tile was never zeroed, so store instead of add
      dst[lane] = e->node[lane];
      dst[lane + 32] = e->node[lane + 32];
    } else {
      red_add_v4(dst + lane, e->node[lane]);
      red_add_v4(dst + lane + 32, e->node[lane + 32]);
    }
  }
}

constexpr int MIG_WORDS = NFIELDS + 1;  // the 34 state words + the elastic energy
__global__ void __launch_bounds__(256) k_migrate_pack(ParticleBuf P, const float* __restrict__ energy, StepScalars* S, float h, int lo, int hi, int reach_lo, int reach_hi,
                                                      uint32_t* __restrict__ out_left, uint32_t* __restrict__ out_right, uint32_t* __restrict__ counts, uint32_t cap, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t flags = P.u(PFLAGS)[i];
  if (flags & (F_TOMBSTONED | F_GONE)) return;
  const int bx = floor_div4(base_node(P.f(PX)[i], h));
  if (bx >= lo && bx < hi) return;
  if (bx < reach_lo || bx >= reach_hi) { atomicOr(&S->status, ST_KEY_RANGE); return; }  // crossed more than one slab in a substep
  const int side = bx < lo ? 0 : 1;
  const uint32_t slot = atomicAdd(&counts[side], 1u);
  P.u(PFLAGS)[i] = flags | F_GONE;
  if (slot >= cap) return;
  uint32_t* row = (side ? out_right : out_left) + (size_t)slot * MIG_WORDS;
#pragma unroll
  for (int f = 0; f < NFIELDS; ++f) row[f] = P.u(f)[i];
  row[PFLAGS] = flags;  // the row travels without the local F_GONE mark
  row[NFIELDS] = __float_as_uint(energy[i]);
}
__global__ void __launch_bounds__(256) k_migrate_unpack(ParticleBuf P, float* __restrict__ energy, const uint32_t* __restrict__ in, uint32_t count, uint32_t base) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= count) return;
  const uint32_t* row = in + (size_t)q * MIG_WORDS;
  const uint32_t i = base + q;
#pragma unroll
  for (int f = 0; f < NFIELDS; ++f) P.u(f)[i] = row[f];
  energy[i] = __uint_as_float(row[NFIELDS]);
}
// migration over peer memory.  Sending side: rows that left [lo, hi) go straight into the neighbour's mailbox;
// the last block publishes both counts and the sequence numbers.
struct SlabPeers {
  uint32_t* rows[2];                     // neighbour mailbox: migrating rows (left / right neighbour), null at the domain ends
  uint32_t* count[2]; uint32_t* seq[2];
  uint32_t* err_seq[SLAB_MAX_RANKS]; uint32_t* err_val[SLAB_MAX_RANKS];   // slot `rank` of every rank's header (null for self)
  int n_ranks;
};
// Sending side, driven by the lists k_g2p<SLAB> filled (slots whose advanced position left the slab): one thread per
// migrating row; the last block publishes both counts and the sequence numbers.  (This rank's error word is posted by
// k_migrate_recv, which follows G2P on the main stream: this kernel may run BESIDE G2P, whose interior tiles can still raise one.)
__global__ void __launch_bounds__(256) k_migrate_send_list(ParticleBuf P, const float* __restrict__ energy, StepScalars* S, MigrateCut mc, SlabPeers peers, uint32_t cap, uint32_t seq,
                                                           uint32_t* __restrict__ blocks_done, int between_substeps, const uint32_t* done, const uint32_t* n_boundary, uint32_t parts) {
  __shared__ uint32_t s_c[2];
  if (!between_substeps && S->sticky) return;   // stopped run: no message (see k_halo_send2)
  SVB_TRACE(5, seq);
  if (done && !SVB_ABORTED(S)) {
    // launched next to G2P on a second stream: only boundary tiles hold particles that can leave the slab
    __shared__ int s_ok;
    if (threadIdx.x == 0) { WaitClock wc(2); s_ok = wait_boundary_done(done, n_boundary, parts) ? 1 : 0; wc.stop(); }
    __syncthreads();
    if (!s_ok && threadIdx.x == 0) atomicOr(&S->status, ST_COMM_TIMEOUT);
  }
  SVB_TRACE(6, seq);
  if (threadIdx.x < 2) s_c[threadIdx.x] = atomicAdd(&mc.counts[threadIdx.x], 0u);
  __syncthreads();
  const uint32_t c0 = s_c[0], c1 = s_c[1];
  const uint32_t l0 = min(c0, min(mc.cap, cap)), l1 = min(c1, min(mc.cap, cap));
  if (!SVB_ABORTED(S)) {
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < l0 + l1; q += gridDim.x * blockDim.x) {
      const int side = q < l0 ? 0 : 1;
      const uint32_t slot = side ? q - l0 : q;
      const uint32_t i = mc.list[(size_t)side * mc.cap + slot];
      const uint32_t flags = P.u(PFLAGS)[i];
      P.u(PFLAGS)[i] = flags | F_GONE;
      if (!peers.rows[side]) continue;
      uint32_t* row = peers.rows[side] + (size_t)slot * MIG_WORDS;
#pragma unroll
      for (int f = 0; f < NFIELDS; ++f) row[f] = P.u(f)[i];
      row[PFLAGS] = flags;  // the row travels without the local F_GONE mark
      row[NFIELDS] = __float_as_uint(energy[i]);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) __threadfence_system();
  if (threadIdx.x == 0 && atomicAdd(blocks_done, 1u) == gridDim.x - 1) {
    const uint32_t c[2] = {c0, c1};
    for (int side = 0; side < 2; ++side) {
      if (peers.count[side]) {
        st_sys(peers.count[side], c[side] > min(mc.cap, cap) ? (min(mc.cap, cap) | SLAB_OVERFLOW) : c[side]);
        st_sys(peers.seq[side], seq);
      }
      mc.counts[side] = 0;
    }
    *blocks_done = 0;
    g_exchange_trace[7] = seq;
  }
}
// receiving side: post this rank's sticky error word to every rank (G2P / the advance of this substep are complete: this kernel
// follows them on the main stream), append the neighbours' rows behind this rank's rows, publish the new row count on the device,
// and fold every rank's error word into this rank's (a FAILED particle anywhere stops every rank after this substep)
__global__ void __launch_bounds__(256) k_migrate_recv(ParticleBuf P, float* __restrict__ energy, StepScalars* S, const SlabHeader* __restrict__ hdr, const uint32_t* __restrict__ rows_left,
                                                      const uint32_t* __restrict__ rows_right, int has_left, int has_right, int rank, int n_ranks, uint32_t seq, uint32_t* __restrict__ n_dev,
                                                      int between_substeps, SimConsts K, BinNext bn, int bin, SlabPeers peers) {
  __shared__ uint32_t s_c[2];
  if (!between_substeps && S->sticky) return;   // stopped run: nothing was sent (see k_halo_send2)
  SVB_TRACE(8, seq);
  if (blockIdx.x == 0 && threadIdx.x < (unsigned)n_ranks && peers.err_val[threadIdx.x]) {
    const uint32_t err = (S->sticky | S->sticky_new) & 0xffffu;   // simulation-level errors only: stop bits are every rank's own
    st_sys(peers.err_val[threadIdx.x], err);
    st_sys(peers.err_seq[threadIdx.x], seq);
  }
  if (threadIdx.x == 0) {
    uint32_t c[2] = {0, 0};
    const int has[2] = {has_left, has_right};
    WaitClock wc(3);
    for (int side = 0; side < 2; ++side)
      if (has[side]) {
        if (wait_seq(&hdr->mig_seq[side], seq)) {
          c[side] = ld_sys(&hdr->mig_count[side]);
          if (c[side] & SLAB_OVERFLOW) { atomicOr(&S->status, ST_COMM_OVERFLOW); c[side] &= ~SLAB_OVERFLOW; }
        } else atomicOr(&S->status, ST_COMM_TIMEOUT);
      }
    wc.stop();
    s_c[0] = c[0]; s_c[1] = c[1];
    if (blockIdx.x == 0) {
      WaitClock we(4);
      uint32_t err = 0;
      for (int r = 0; r < n_ranks; ++r)
        if (r != rank) {
          if (wait_seq(&hdr->err_seq[r], seq)) err |= ld_sys(&hdr->err_val[r]);
          else atomicOr(&S->status, ST_COMM_TIMEOUT);
        }
      if (err) atomicOr(&S->sticky_new, err);
      we.stop();
    }
  }
  __syncthreads();
  SVB_TRACE(9, seq);
  const uint32_t cl = s_c[0], cr = s_c[1];
  // after a substep the rows are the re-binned ones; a rebalance between substeps appends behind whatever is resident
  const uint32_t base = between_substeps ? *n_dev : S->n_live + S->n_tomb;
  const uint32_t room = (uint32_t)P.cap > base ? (uint32_t)P.cap - base : 0u;
  if (cl + cr > room) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&S->status, ST_COMM_OVERFLOW); }
  const uint32_t total = min(cl + cr, room);
  const bool ran = S->bin_blocks_done != 0 && !(S->sticky);   // this substep was a real one (not the no-op after a stop)
  const uint32_t lane = threadIdx.x & 31;
  // warp-uniform trip count: bin_warp (the arrivals join the next substep's bins, like the rows G2P wrote) is warp-collective
  for (uint32_t q0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); q0 < total; q0 += gridDim.x * blockDim.x) {
    const uint32_t q = q0 + lane;
    int state = 2;
    V3 x = V3{0.f, 0.f, 0.f};
    if (q < total) {
      const uint32_t* row = q < cl ? rows_left + (size_t)q * MIG_WORDS : rows_right + (size_t)(q - cl) * MIG_WORDS;
      const uint32_t i = base + q;
#pragma unroll
      for (int f = 0; f < NFIELDS; ++f) P.u(f)[i] = row[f];
      energy[i] = __uint_as_float(row[NFIELDS]);
      x = V3{__uint_as_float(row[PX]), __uint_as_float(row[PX + 1]), __uint_as_float(row[PX + 2])};
      state = (row[PFLAGS] & F_TOMBSTONED) ? 1 : 0;
    }
    if (bin && ran) bin_warp<false, false>(state, x, 0u, q < total, base + q, K, bn.T, bn.B, bn.S, lane);
  }
  // a substep that was a no-op on every rank (an earlier substep failed: k_bin returned at once) keeps the row count
  if (between_substeps) {
    // every block reads *n_dev as its base: publish the new count only after all of them did (last block, like the senders)
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      uint32_t* done = n_dev + 1;   // scratch word next to the row count
      if (atomicAdd(done, 1u) == gridDim.x - 1) { *n_dev = base + total; *done = 0; }
    }
  } else if (blockIdx.x == 0 && threadIdx.x == 0 && S->bin_blocks_done != 0) *n_dev = base + total;
}
// rebalance (between substeps): every resident live row whose block column lies outside the rank's NEW range goes on the
// per-side list k_migrate_send_list consumes (the per-substep path fills that list from k_g2p<SLAB> instead)
__global__ void __launch_bounds__(256) k_note_outside(ParticleBuf P, StepScalars* S, float h, MigrateCut mc, const uint32_t* __restrict__ n_dev) {
  const uint32_t n = *n_dev;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t flags = P.u(PFLAGS)[i];
    if (flags & (F_TOMBSTONED | F_GONE)) continue;
    const int bx = floor_div4(base_node(P.f(PX)[i], h));
    if (bx >= mc.lo && bx < mc.hi) continue;
    if (bx < mc.reach_lo || bx >= mc.reach_hi) { atomicOr(&S->status, ST_KEY_RANGE); continue; }
    const int side = bx < mc.lo ? 0 : 1;
    const uint32_t slot = atomicAdd(&mc.counts[side], 1u);
    if (slot < mc.cap) mc.list[(size_t)side * mc.cap + slot] = i;
  }
}
// live particles per block column: counts[0] = below first_col, counts[1 + k] = column first_col + k, counts[n_cols + 1] = above
__global__ void __launch_bounds__(256) k_column_histogram(ParticleBuf P, float h, uint32_t n, int first_col, uint32_t n_cols, unsigned long long* __restrict__ counts) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (P.u(PFLAGS)[i] & (F_TOMBSTONED | F_GONE)) continue;
    const long long k = (long long)floor_div4(base_node(P.f(PX)[i], h)) - first_col;
    const uint32_t bin = k < 0 ? 0u : (k >= (long long)n_cols ? n_cols + 1u : (uint32_t)k + 1u);
    // consecutive rows share a column (the state is binned): one atomic per distinct bin per warp
    const unsigned peers = __match_any_sync(__activemask(), bin);
    if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&counts[bin], (unsigned long long)__popc(peers));
  }
}
__global__ void k_add_orig(ParticleBuf P, uint32_t n, uint32_t add) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) P.u(PORIG)[i] += add;
}
// rows that are still resident (not migrated away), compacted in arbitrary order
__global__ void __launch_bounds__(256) k_resident_rows(ParticleBuf P, uint32_t n, uint32_t* __restrict__ rows, uint32_t* __restrict__ count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool keep = i < n && !(P.u(PFLAGS)[i] & F_GONE);
  const unsigned m = __ballot_sync(SVB_FULL, keep);
  uint32_t base = 0;
  if ((threadIdx.x & 31) == 0 && m) base = atomicAdd(count, (uint32_t)__popc(m));
  base = __shfl_sync(SVB_FULL, base, 0);
  if (keep) rows[base + __popc(m & ((1u << (threadIdx.x & 31)) - 1u))] = i;
}
template <int K>
__global__ void k_rows_to_wire(ParticleBuf P, int field, const uint32_t* __restrict__ rows, uint32_t* __restrict__ dst, uint32_t n) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const uint32_t i = rows[q];
#pragma unroll
  for (int k = 0; k < K; ++k) dst[(size_t)q * K + k] = P.u(field + k)[i];
}
__global__ void k_array_rows_to_wire(const float* __restrict__ src, const uint32_t* __restrict__ rows, float* __restrict__ dst, uint32_t n) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) dst[q] = src[rows[q]];
}

// ------------------------------------------------------------------------------------------------
// grid download helpers: which nodes of an active tile have >= 1 contributor
__global__ void __launch_bounds__(256) k_touch_nodes(ParticleBuf P, const uint32_t* __restrict__ src_of, const uint2* __restrict__ group_range, const int* __restrict__ nbr,
                                                     const StepScalars* __restrict__ S, float h, unsigned long long* __restrict__ node_mask) {
  if (SVB_ABORTED(S)) return;
  const uint32_t n_groups = S->n_ptiles;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < n_groups; g += warps) {
    const uint2 range = group_range[g];
    const uint32_t start = range.x, end = range.y;
    for (uint32_t j = start + lane; j < end; j += 32) {
      const uint32_t i = src_of[j];
      const int s0 = base_node(P.f(PX)[i], h) & 3, s1 = base_node(P.f(PX + 1)[i], h) & 3, s2 = base_node(P.f(PX + 2)[i], h) & 3;
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
          for (int c = 0; c < 3; ++c) {
            const int ti = s0 + a, tj = s1 + b, tk = s2 + c;
            const int d = (ti >> 2) | ((tj >> 2) << 1) | ((tk >> 2) << 2);
            const int nb = nbr[(size_t)g * 8 + d];
            if (nb >= 0) atomicOr(&node_mask[nb], 1ull << (((ti & 3) << 4) | ((tj & 3) << 2) | (tk & 3)));
          }
    }
  }
}
__global__ void __launch_bounds__(64) k_emit_grid(const float4* __restrict__ grid, TileTable T, MeldInfo Mi, StepScalars* S, const unsigned long long* __restrict__ node_mask,
                                                  const uint32_t* __restrict__ node_offset, int32_t* __restrict__ node_ids, uint32_t* __restrict__ out_bits, float* __restrict__ masses,
                                                  float* __restrict__ velocities) {
  const uint32_t e = blockIdx.x;
  const int node = threadIdx.x;
  __shared__ int sib[SIB_MAX];
  __shared__ int nsib;
  const unsigned long long k = T.tile_key[e];
  if (threadIdx.x == 0) nsib = sibling_tiles(T, Mi, S->n_layers, k, (int)e, sib, S);
  __syncthreads();
  const unsigned long long mask = node_mask[e];
  if (!((mask >> node) & 1ull)) return;
  const uint32_t at = node_offset[e] + __popcll(mask & ((1ull << node) - 1ull));
  int bx, by, bz;
  uint32_t layer;
  tile_key_unpack(k, bx, by, bz, layer);
  node_ids[3 * at] = (bx << 2) + (node >> 4);
  node_ids[3 * at + 1] = (by << 2) + ((node >> 2) & 3);
  node_ids[3 * at + 2] = (bz << 2) + (node & 3);
  out_bits[at] = layer_bits_of(Mi.layer_slots, layer);
  float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int q = 0; q < nsib; ++q) {
    const float4 v = grid[(size_t)sib[q] * 64 + node];
    sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
  }
  sum = finish_node(sum);
  masses[at] = sum.w;
  velocities[3 * at] = sum.x; velocities[3 * at + 1] = sum.y; velocities[3 * at + 2] = sum.z;
}
__global__ void __launch_bounds__(64) k_mask_from_values(const float4* __restrict__ grid, unsigned long long* __restrict__ node_mask) {
  const float4 v = grid[(size_t)blockIdx.x * 64 + threadIdx.x];
  const bool on = v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f;
  const unsigned lo = __ballot_sync(SVB_FULL, on);
  __shared__ unsigned parts[2];
  if ((threadIdx.x & 31) == 0) parts[threadIdx.x >> 5] = lo;
  __syncthreads();
  if (threadIdx.x == 0) node_mask[blockIdx.x] = (unsigned long long)parts[0] | ((unsigned long long)parts[1] << 32);
}
__global__ void k_popc_masks(const unsigned long long* __restrict__ node_mask, uint32_t* __restrict__ counts, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) counts[i] = __popcll(node_mask[i]);
}
__global__ void k_decode_active(TileTable T, const unsigned long long* __restrict__ layer_slots, uint32_t n_active, int32_t* __restrict__ block_ids, uint32_t* __restrict__ out_bits) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_active) return;
  int bx, by, bz;
  uint32_t layer;
  tile_key_unpack(T.tile_key[e], bx, by, bz, layer);
  block_ids[3 * e] = bx; block_ids[3 * e + 1] = by; block_ids[3 * e + 2] = bz;
  out_bits[e] = layer_bits_of(layer_slots, layer);
}
// the reference's node_ids_to_murmur stage (gpu/src/node_ids_to_murmur/mod.rs; test.rs:15-96): both hashes of every (node id, bits)
__global__ void k_node_ids_to_murmur(const int32_t* __restrict__ node_ids, const uint32_t* __restrict__ bits, uint32_t n, uint32_t* __restrict__ plain, uint32_t* __restrict__ seeded) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t x = node_ids[3 * i], y = node_ids[3 * i + 1], z = node_ids[3 * i + 2];
  if (plain) plain[i] = node_id_to_murmur(x, y, z, 0u);
  if (seeded) seeded[i] = node_id_to_murmur(x, y, z, bits ? bits[i] : 0u);
}
__global__ void k_cells(ParticleBuf P, float h, uint32_t n, int32_t* __restrict__ cells) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  cells[3 * i] = base_node(P.f(PX)[i], h);
  cells[3 * i + 1] = base_node(P.f(PX + 1)[i], h);
  cells[3 * i + 2] = base_node(P.f(PX + 2)[i], h);
}

}  // namespace svb
