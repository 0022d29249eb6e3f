// svb_kernels.cuh — hand-written CUDA kernels (sm_100a) of the MPM substep.
//
// Phase map to the reference's CPU back end (/root/reference/rust/crates/cpu/src/phase/*.rs):
//   k_mesh_*            interpolate_input.rs:18-107      collider mesh lerp, face / vertex normals
//   k_force             collide.rs:21-206 + external_force.rs:17-50 (+ live bounding box, layer set)
//   k_layout, k_keys    sort.rs:29-33                     bin keys from floor(x/h - 1/2)
//   k_permute           sort.rs:91-101                    gather the SoA state into binned order
//   k_flag_*            update_grid_nodes.rs:31-156       block / layer activation (sorted, no hash map)
//   k_group_touch, k_neighbors                            which neighbour blocks a run of particles touches
//   k_p2g               scatter_momentum.rs:22-93         particle -> grid, stress once per particle
//   k_g2p               meld_grid.rs:16-69 + collect_velocity.rs:19-75 + advance_particles.rs:17-93
//                       + cull_particles.rs:17-41 (+ limit_time_step.rs:187-223 reductions)
//   k_limit_force       limit_time_step.rs:25-182
#pragma once
#include "svb_device.cuh"

namespace svb {

#define SVB_FULL 0xffffffffu

__device__ __forceinline__ int floor_div4(int v) { return v >> 2; }
__device__ __forceinline__ int ceil_log2_u32(uint32_t v) {  // bits needed to represent values 0..v-1
  return v <= 1 ? 0 : 32 - __clz(v - 1);
}

// cpu/src/kernels.rs:46-49 — must be bit-exact: IEEE divide, subtract, floor, no contraction.
__device__ __forceinline__ int base_node(float x, float h) { return (int)floorf(__fsub_rn(__fdiv_rn(x, h), 0.5f)); }

// ------------------------------------------------------------------------------------------------
// wire <-> SoA
template <int K>
__global__ void k_wire_to_soa(const float* __restrict__ src, float* __restrict__ dst, size_t cap, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
#pragma unroll
  for (int k = 0; k < K; ++k) dst[(size_t)k * cap + i] = src[(size_t)i * K + k];
}
// out[(original index)*K + k] = field[k][i]   (to_io_state, cpu/src/cpu_state.rs:84-93)
template <int K>
__global__ void k_soa_to_wire(const float* __restrict__ src, size_t cap, const uint32_t* __restrict__ orig, float* __restrict__ dst, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t o = orig ? orig[i] : i;
#pragma unroll
  for (int k = 0; k < K; ++k) dst[o * K + k] = src[(size_t)k * cap + i];
}
__global__ void k_iota(uint32_t* a, uint32_t n, uint32_t offset) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i + offset;
}

// ------------------------------------------------------------------------------------------------
// collider mesh interpolation (interpolate_input.rs:36-96)
__global__ void k_mesh_lerp(MeshDev M, float factor_b) {
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

// collide.rs:45-204 for one particle.  Returns the new collider bits; edits `vel`.
__device__ __noinline__ uint32_t collide_particle(const MeshDev& M, const SimConsts& K, float dt, V3 p, V3& vel, uint32_t bits) {
  const int lx = (int)floorf(p.x / K.leaf_size), ly = (int)floorf(p.y / K.leaf_size), lz = (int)floorf(p.z / K.leaf_size);
  int first, count;
  bvh_query(M, lx, ly, lz, first, count);
  if (count == 0) return 0u;
  uint32_t closest[16];
  float min_dist[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) { closest[c] = 0xffffffffu; min_dist[c] = 3.402823466e+38f; }
  for (int q = 0; q < count; ++q) {
    const uint32_t t = M.tri_indices[first + q];
    const V3 n = ld3(M.tnormal, t);
    if (is_zero(n)) continue;
    const float d = triangle_distance(p, ld3(M.vpos, M.tri[3 * t]), ld3(M.vpos, M.tri[3 * t + 1]), ld3(M.vpos, M.tri[3 * t + 2]), n);
    if (d >= K.forget_distance) continue;
    const uint32_t c = M.tri_collider[t] & 15u;
    if (d < min_dist[c]) { min_dist[c] = d; closest[c] = t; }
  }
  for (unsigned collider = 0; collider < 16; ++collider) {
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
  return bits;
}

// ------------------------------------------------------------------------------------------------
// layer set: distinct non-zero collider-bit patterns of the live particles of this substep
__device__ __forceinline__ uint32_t layer_hash(uint32_t b) {
  b ^= b >> 16; b *= 0x7feb352du; b ^= b >> 15; b *= 0x846ca68bu; b ^= b >> 16;
  return b & (LAYER_SLOTS - 1);
}
__device__ __forceinline__ void layer_insert(unsigned long long* slots, uint32_t bits, StepScalars* S) {
  const unsigned long long want = (1ull << 32) | bits;
  uint32_t s = layer_hash(bits);
  for (int tries = 0; tries < LAYER_SLOTS; ++tries) {
    unsigned long long cur = slots[s];
    if (cur == want) return;
    if (cur == 0ull) {
      cur = atomicCAS(&slots[s], 0ull, want);
      if (cur == 0ull || cur == want) return;
    }
    s = (s + 1) & (LAYER_SLOTS - 1);
  }
  atomicOr(&S->status, 1u /*SVB_TABLE_TRIES_EXCEEDED*/);
}
__device__ __forceinline__ uint32_t layer_rank(const unsigned long long* __restrict__ slots, const uint32_t* __restrict__ slot_rank, uint32_t bits) {
  if (bits == 0u) return 0u;
  const unsigned long long want = (1ull << 32) | bits;
  uint32_t s = layer_hash(bits);
  for (int tries = 0; tries < LAYER_SLOTS; ++tries) {
    const unsigned long long cur = slots[s];
    if (cur == want) return slot_rank[s];
    if (cur == 0ull) break;
    s = (s + 1) & (LAYER_SLOTS - 1);
  }
  return 0u;  // unreachable when k_force registered the value
}

// ------------------------------------------------------------------------------------------------
// collide + external force, one thread per particle (current order).  Also accumulates the live
// bounding box of base nodes for the next bin layout and registers collider-bit layers.
struct GoalDev {
  const uint32_t* flags_a; const uint32_t* flags_b;   // original order, may be null
  const float* goal_a; const float* goal_b;           // 3 per particle, original order
};
template <bool HAS_MESH>
__global__ void __launch_bounds__(256) k_force(ParticleBuf P, StepScalars* S, SimConsts K, MeshDev M, GoalDev G, unsigned long long* layer_slots,
                                               uint32_t n, float dt, float gx, float gy, float gz, float factor_b) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  int mn[3] = {INT32_MAX, INT32_MAX, INT32_MAX}, mx[3] = {INT32_MIN, INT32_MIN, INT32_MIN};
  if (i < n) {
    const uint32_t flags = P.u(PFLAGS)[i];
    if (!(flags & F_TOMBSTONED)) {
      const V3 x = V3{P.f(PX)[i], P.f(PX + 1)[i], P.f(PX + 2)[i]};
      V3 v = V3{P.f(PV)[i], P.f(PV + 1)[i], P.f(PV + 2)[i]};
      uint32_t bits = 0u;
      if (HAS_MESH) {
        const uint32_t old_bits = P.u(PBITS)[i];
        bits = collide_particle(M, K, dt, x, v, old_bits);
        if (bits != 0u) layer_insert(layer_slots, bits, S);
      }
      P.u(PBITS)[i] = bits;
      bool goal = false;
      if (G.flags_a) {
        const uint32_t o = P.u(PORIG)[i];
        if ((G.flags_a[o] & F_HAS_GOAL) && (G.flags_b[o] & F_HAS_GOAL)) {
          const float fa = 1.f - factor_b;
          const V3 ga = ld3(G.goal_a, o), gb = ld3(G.goal_b, o);
          const V3 target = fa * ga + factor_b * gb;
          v = (target - x) / dt;
          goal = true;
        }
      }
      if (!goal) v = v + dt * V3{gx, gy, gz};
      P.f(PV)[i] = v.x; P.f(PV + 1)[i] = v.y; P.f(PV + 2)[i] = v.z;
      mn[0] = mx[0] = base_node(x.x, K.h);
      mn[1] = mx[1] = base_node(x.y, K.h);
      mn[2] = mx[2] = base_node(x.z, K.h);
    }
  }
  // block-level min/max, then at most 6 atomics per CTA — and only when they would change the box
  // (same-address atomics from every warp serialise in L2: 32 k warps cost ~100 us at 1 M particles)
  __shared__ int s_lo[3][8], s_hi[3][8];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    int lo = mn[a], hi = mx[a];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = min(lo, __shfl_xor_sync(SVB_FULL, lo, o));
      hi = max(hi, __shfl_xor_sync(SVB_FULL, hi, o));
    }
    if ((threadIdx.x & 31) == 0) { s_lo[a][threadIdx.x >> 5] = lo; s_hi[a][threadIdx.x >> 5] = hi; }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int a = threadIdx.x;
    int lo = s_lo[a][0], hi = s_hi[a][0];
#pragma unroll
    for (int w = 1; w < 8; ++w) { lo = min(lo, s_lo[a][w]); hi = max(hi, s_hi[a][w]); }
    if (lo <= hi) {
      if (lo < *(volatile int*)&S->bbox_min[a]) atomicMin(&S->bbox_min[a], lo);
      if (hi > *(volatile int*)&S->bbox_max[a]) atomicMax(&S->bbox_max[a], hi);
    }
  }
}

// One CTA: bin-key layout from the live bounding box + ranks of the collider-bit layers.
__global__ void __launch_bounds__(1024) k_layout(StepScalars* S, BinLayout* L, const unsigned long long* __restrict__ layer_slots, uint32_t* slot_rank, uint32_t* layer_bits) {
  __shared__ uint32_t vals[LAYER_CAP];
  __shared__ uint32_t vslot[LAYER_CAP];
  __shared__ uint32_t count;
  if (threadIdx.x == 0) count = 0;
  __syncthreads();
  for (uint32_t s = threadIdx.x; s < LAYER_SLOTS; s += blockDim.x) {
    const unsigned long long e = layer_slots[s];
    if (e != 0ull) {
      const uint32_t at = atomicAdd(&count, 1u);
      if (at < LAYER_CAP - 1) { vals[at] = (uint32_t)e; vslot[at] = s; }
    }
  }
  __syncthreads();
  uint32_t nvals = count;
  if (nvals > LAYER_CAP - 1) {
    nvals = LAYER_CAP - 1;
    if (threadIdx.x == 0) atomicOr(&S->status, 1u);
  }
  for (uint32_t q = threadIdx.x; q < nvals; q += blockDim.x) {
    const uint32_t v = vals[q];
    uint32_t r = 1;
    for (uint32_t o = 0; o < nvals; ++o) r += vals[o] < v ? 1u : 0u;
    slot_rank[vslot[q]] = r;
    layer_bits[r] = v;
  }
  if (threadIdx.x == 0) {
    layer_bits[0] = 0u;
    BinLayout l;
    int total = 6;
    for (int a = 0; a < 3; ++a) {
      int lo = S->bbox_min[a], hi = S->bbox_max[a];
      if (lo > hi) { lo = 0; hi = 0; }  // no live particle
      l.cell_min[a] = lo; l.cell_max[a] = hi;
      const int bmin = floor_div4(lo), bmax = floor_div4(hi + 2);
      l.block_min[a] = bmin;
      l.nb[a] = ceil_log2_u32((uint32_t)(bmax - bmin + 1));
      total += l.nb[a];
    }
    l.n_layers = (int)nvals + 1;
    l.nl = ceil_log2_u32((uint32_t)l.n_layers);
    total += l.nl;
    l.total_bits = total;
    if (total > 62) atomicOr(&S->status, 0x80000000u);  // SVB_KEY_RANGE
    *L = l;
    S->layer_count = (uint32_t)l.n_layers;
    S->n_live = 0; S->n_groups = 0; S->n_cand = 0; S->n_active = 0;
    S->work_counter[0] = S->work_counter[1] = S->work_counter[2] = S->work_counter[3] = 0;
  }
}

// bin key of every particle (cpu/src/phase/sort.rs:29-33 — the cell (i,j,k) is bit-exact; the ORDER
// is block-major so that one run of the sorted array is one (block, layer) tile of work).
__global__ void __launch_bounds__(256) k_keys(ParticleBuf P, const BinLayout* __restrict__ Lp, const unsigned long long* __restrict__ layer_slots,
                                              const uint32_t* __restrict__ slot_rank, float h, uint32_t n, unsigned long long* __restrict__ keys, uint32_t* __restrict__ idx) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const BinLayout L = *Lp;
  const uint32_t flags = P.u(PFLAGS)[i];
  unsigned long long key;
  if (flags & F_TOMBSTONED) {
    key = 1ull << L.total_bits;
  } else {
    const int sx = base_node(P.f(PX)[i], h), sy = base_node(P.f(PX + 1)[i], h), sz = base_node(P.f(PX + 2)[i], h);
    const unsigned long long bx = (unsigned long long)(floor_div4(sx) - L.block_min[0]);
    const unsigned long long by = (unsigned long long)(floor_div4(sy) - L.block_min[1]);
    const unsigned long long bz = (unsigned long long)(floor_div4(sz) - L.block_min[2]);
    const uint32_t rank = L.nl ? layer_rank(layer_slots, slot_rank, P.u(PBITS)[i]) : 0u;
    const uint32_t cell = ((uint32_t)(sx & 3) << 4) | ((uint32_t)(sy & 3) << 2) | (uint32_t)(sz & 3);
    key = (((((bx << L.nb[1]) | by) << L.nb[2]) | bz) << L.nl | rank) << 6 | cell;
  }
  keys[i] = key;
  idx[i] = i;
}

// physical re-bin: dst[f][i] = src[f][idx[i]] for every field
__global__ void __launch_bounds__(256) k_permute(ParticleBuf src, ParticleBuf dst, const uint32_t* __restrict__ idx, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t j = idx[i];
  uint32_t v[NFIELDS];
#pragma unroll
  for (int f = 0; f < NFIELDS; ++f) v[f] = __ldg(src.base + (size_t)f * src.cap + j);
#pragma unroll
  for (int f = 0; f < NFIELDS; ++f) dst.base[(size_t)f * dst.cap + i] = v[f];
}

// ------------------------------------------------------------------------------------------------
// run heads by a 3-kernel exclusive scan.  MODE 0: heads of (block, layer) runs in the sorted
// particle keys (key >> 6 changes, tombstoned excluded) -> out_pos[g] = first particle of run g.
// MODE 1: unique values of the sorted candidate keys (~0 = empty) -> out_key[a] = key.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <int MODE>
__device__ __forceinline__ bool head_flag(const unsigned long long* __restrict__ keys, uint32_t i, uint32_t n, int total_bits) {
  if (i >= n) return false;
  const unsigned long long k = keys[i];
  if (MODE == 0) {
    if (k >> total_bits) return false;
    return i == 0 || (keys[i - 1] >> 6) != (k >> 6);
  } else {
    if (k >> (total_bits - 6)) return false;  // empty candidate slot (bit just above the tile-key range)
    return i == 0 || keys[i - 1] != k;
  }
}
template <int MODE>
__global__ void __launch_bounds__(SCAN_THREADS) k_flag_count(const unsigned long long* __restrict__ keys, const uint32_t* n_ptr, uint32_t n_mul, const BinLayout* __restrict__ L,
                                                             uint32_t* __restrict__ tile_count, StepScalars* S) {
  const uint32_t n = n_ptr ? *n_ptr * n_mul : n_mul;
  const int tb = L->total_bits;
  uint32_t c = 0;
  const uint32_t base = blockIdx.x * SCAN_TILE;
#pragma unroll
  for (int q = 0; q < SCAN_ITEMS; ++q) {
    const uint32_t i = base + q * SCAN_THREADS + threadIdx.x;
    c += head_flag<MODE>(keys, i, n, tb) ? 1u : 0u;
    if (MODE == 0 && i < n && !(keys[i] >> tb) && (i + 1 == n || (keys[i + 1] >> tb))) S->n_live = i + 1;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(SVB_FULL, c, o);
  __shared__ uint32_t ws[SCAN_THREADS / 32];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < SCAN_THREADS / 32; ++w) t += ws[w];
    tile_count[blockIdx.x] = t;
  }
}
// single CTA: exclusive scan of the tile counts in place; total -> *total_out
__global__ void __launch_bounds__(1024) k_scan_tiles(uint32_t* tile_count, uint32_t n_tiles, uint32_t* total_out) {
  __shared__ uint32_t ws[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n_tiles; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < n_tiles ? tile_count[i] : 0u;
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
    if (i < n_tiles) tile_count[i] = c + warp_off + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + warp_off + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry;
}
template <int MODE>
__global__ void __launch_bounds__(SCAN_THREADS) k_flag_write(const unsigned long long* __restrict__ keys, const uint32_t* n_ptr, uint32_t n_mul, const BinLayout* __restrict__ L,
                                                             const uint32_t* __restrict__ tile_offset, uint32_t* __restrict__ out_pos, unsigned long long* __restrict__ out_key,
                                                             uint32_t out_cap, StepScalars* S) {
  const uint32_t n = n_ptr ? *n_ptr * n_mul : n_mul;
  const int tb = L->total_bits;
  // blocked arrangement so output order = input order
  const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  bool fl[SCAN_ITEMS];
  uint32_t c = 0;
#pragma unroll
  for (int q = 0; q < SCAN_ITEMS; ++q) {
    fl[q] = head_flag<MODE>(keys, base + q, n, tb);
    c += fl[q] ? 1u : 0u;
  }
  uint32_t inc = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(SVB_FULL, inc, o);
    if ((threadIdx.x & 31) >= o) inc += t;
  }
  __shared__ uint32_t ws[SCAN_THREADS / 32];
  if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = inc;
  __syncthreads();
  uint32_t off = tile_offset[blockIdx.x] + inc - c;
  for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) off += ws[w];
#pragma unroll
  for (int q = 0; q < SCAN_ITEMS; ++q)
    if (fl[q]) {
      if (off < out_cap) {
        if (MODE == 0) out_pos[off] = base + q;
        else out_key[off] = keys[base + q];
      } else {
        atomicOr(&S->status, 4u /*SVB_INDIRECT_LIMIT_EXCEEDED*/);
      }
      ++off;
    }
}

// ------------------------------------------------------------------------------------------------
// Which of the 8 neighbour blocks (offsets d in {0,1}^3) does a run of particles touch?  A particle
// whose base node sits at in-block coordinate c touches block +1 on an axis iff c >= 2 (stencil
// c..c+2).  One warp per run; emits up to 8 candidate tile keys per run.
__device__ __forceinline__ unsigned long long group_delta(const BinLayout& L, int d) {
  unsigned long long r = 0;
  if (d & 1) r += 1ull << (L.nb[1] + L.nb[2] + L.nl);
  if (d & 2) r += 1ull << (L.nb[2] + L.nl);
  if (d & 4) r += 1ull << L.nl;
  return r;
}
__global__ void __launch_bounds__(256) k_group_touch(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ group_start, const BinLayout* __restrict__ Lp,
                                                     const StepScalars* __restrict__ S, unsigned long long* __restrict__ cand, uint32_t* __restrict__ group_touch) {
  const BinLayout L = *Lp;
  const uint32_t n_groups = S->n_groups, n_live = S->n_live;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < n_groups; g += warps) {
    const uint32_t start = group_start[g];
    const uint32_t end = g + 1 < n_groups ? group_start[g + 1] : n_live;
    uint32_t t = 0;
    for (uint32_t i = start + lane; i < end; i += 32) {
      const uint32_t cell = (uint32_t)keys[i] & 63u;
      const uint32_t m = ((cell >> 5) & 1u) | (((cell >> 3) & 1u) << 1) | (((cell >> 1) & 1u) << 2);  // c >= 2 per axis
      // subsets of m as a mask over d (bit0 = +x, bit1 = +y, bit2 = +z)
      uint32_t sub = 1u;
      if (m & 1u) sub |= sub << 1;
      if (m & 2u) sub |= sub << 2;
      if (m & 4u) sub |= sub << 4;
      t |= sub;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t |= __shfl_xor_sync(SVB_FULL, t, o);
    if (lane < 8) {
      const unsigned long long gk = keys[start] >> 6;
      cand[(size_t)g * 8 + lane] = ((t >> lane) & 1u) ? gk + group_delta(L, (int)lane) : (1ull << (L.total_bits - 6));
    }
    if (lane == 0) group_touch[g] = t;
  }
}
__device__ __forceinline__ int find_key(const unsigned long long* __restrict__ a, uint32_t n, unsigned long long k) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (a[mid] < k) lo = mid + 1; else hi = mid;
  }
  return (lo < n && a[lo] == k) ? (int)lo : -1;
}
__global__ void __launch_bounds__(256) k_neighbors(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ group_start, const uint32_t* __restrict__ group_touch,
                                                   const BinLayout* __restrict__ Lp, const StepScalars* __restrict__ S, const unsigned long long* __restrict__ active_keys,
                                                   int* __restrict__ nbr) {
  const BinLayout L = *Lp;
  const uint32_t n_groups = S->n_groups, n_active = S->n_active;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n_groups * 8; q += gridDim.x * blockDim.x) {
    const uint32_t g = q >> 3, d = q & 7;
    int r = -1;
    if ((group_touch[g] >> d) & 1u) r = find_key(active_keys, n_active, (keys[group_start[g]] >> 6) + group_delta(L, (int)d));
    nbr[q] = r;
  }
}

// ------------------------------------------------------------------------------------------------
// P2G (scatter_momentum.rs:22-93).  One CTA per (block, layer) run of particles, claimed from a
// work counter.  Each warp takes 32 consecutive particles: every lane evaluates ITS particle once
// (weights, affine momentum matrix A = m C - s V0 P F^T - s J V0 sigma, so the stress is computed
// once per particle, not 27 times) and parks the 32 numbers a node needs in shared memory; then the
// warp walks the 32 particles together with lane = one of the 27 stencil nodes, accumulating in
// registers while consecutive particles share a cell (they do: the run is sorted by cell) — the
// warp-aggregated form of the scatter.  A cell change flushes 27 float4 into the warp's private
// 6x6x6 tile (plain read-modify-write, no shared atomics: fp32 shared atomics are CAS loops).  At
// the end the warps' tiles are summed and every non-zero tile node goes to HBM with one
// red.global.add.v4.f32.
constexpr int P2G_WARPS = 4;
constexpr int TILE_NODES = 216;
constexpr int STAGE_STRIDE = 36;  // floats per staged particle: 16-byte aligned rows, conflict-free float4 stores
constexpr int P2G_SMEM = P2G_WARPS * TILE_NODES * 16 + P2G_WARPS * 32 * STAGE_STRIDE * 4;

__device__ __forceinline__ void red_add_v4(float4* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(P2G_WARPS * 32, 6) k_p2g(ParticleBuf P, const uint32_t* __restrict__ group_start, const int* __restrict__ nbr, StepScalars* S, float4* __restrict__ grid,
                                                           float h, float dt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* tiles = reinterpret_cast<float4*>(smem_raw);
  float* stage_all = reinterpret_cast<float*>(smem_raw + P2G_WARPS * TILE_NODES * 16);
  __shared__ uint32_t s_group;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4* my_tile = tiles + warp * TILE_NODES;
  float* stage = stage_all + warp * 32 * STAGE_STRIDE;
  const uint32_t n_groups = S->n_groups, n_live = S->n_live;
  const float scaling = dt * 4.f / (h * h);
  // node handled by this lane in the 3x3x3 stencil (k fastest); lanes 27..31 shadow node 0 and never flush
  const bool node_lane = lane < 27;
  const int li = node_lane ? lane / 9 : 0, lj = node_lane ? (lane / 3) % 3 : 0, lk = node_lane ? lane % 3 : 0;
  // staged row (floats): [0..17] (w,d) pairs x0 x1 x2 y0 y1 y2 z0 z1 z2 | 18 cell | 19 mass | 20..22 m*v | 23 A8 | 24..31 A0..A7
  const float* lane_x = stage + 2 * li;
  const float* lane_y = stage + 6 + 2 * lj;
  const float* lane_z = stage + 12 + 2 * lk;
  const int lane_tile_off = (li * 6 + lj) * 6 + lk;

  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_group = atomicAdd(&S->work_counter[0], 1u);
    __syncthreads();
    const uint32_t g = s_group;
    if (g >= n_groups) break;
    const uint32_t start = group_start[g];
    const uint32_t end = g + 1 < n_groups ? group_start[g + 1] : n_live;
    for (int q = threadIdx.x; q < P2G_WARPS * TILE_NODES; q += blockDim.x) tiles[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    for (uint32_t chunk = start + warp * 32; chunk < end; chunk += P2G_WARPS * 32) {
      const uint32_t i = chunk + lane;
      // ---- per-particle evaluation by the owning lane
      float4 st[8];
      int cell = -1;
      if (i < end) {
        const float x0 = P.f(PX)[i], x1 = P.f(PX + 1)[i], x2 = P.f(PX + 2)[i];
        const float n0 = __fdiv_rn(x0, h), n1 = __fdiv_rn(x1, h), n2 = __fdiv_rn(x2, h);
        const int s0 = (int)floorf(__fsub_rn(n0, 0.5f)), s1 = (int)floorf(__fsub_rn(n1, 0.5f)), s2 = (int)floorf(__fsub_rn(n2, 0.5f));
        cell = ((s0 & 3) << 4) | ((s1 & 3) << 2) | (s2 & 3);
        float w[9], d[9];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const float dn0 = (float)(s0 + a) - n0, dn1 = (float)(s1 + a) - n1, dn2 = (float)(s2 + a) - n2;
          w[a] = kernel_quadratic(dn0); w[3 + a] = kernel_quadratic(dn1); w[6 + a] = kernel_quadratic(dn2);
          d[a] = dn0 * h; d[3 + a] = dn1 * h; d[6 + a] = dn2 * h;
        }
        M3 C, F;
#pragma unroll
        for (int q = 0; q < 9; ++q) { C.m[q] = P.f(PC + q)[i]; F.m[q] = P.f(PF + q)[i]; }
        const float mass = P.f(PMASS)[i], vol = P.f(PVOL)[i], p0 = P.f(PP0)[i], p1 = P.f(PP1)[i];
        const uint32_t flags = P.u(PFLAGS)[i];
        const M3 stress = (flags & F_IS_FLUID) ? first_piola_inviscid(p0, (int)p1, F) : first_piola_neo_hookean(p0, p1, F);
        const M3 pft = mul_nt(stress, F);
        const float sv = scaling * vol;
        M3 A;
#pragma unroll
        for (int q = 0; q < 9; ++q) A.m[q] = mass * C.m[q] - sv * pft.m[q];
        if (flags & F_USE_VISCOSITY) {
          const M3 cauchy = viscous_cauchy(P.f(PVD)[i], P.f(PVB)[i], C);
          const float sj = scaling * det(F) * vol;
#pragma unroll
          for (int q = 0; q < 9; ++q) A.m[q] -= sj * cauchy.m[q];
        }
        const float mv0 = mass * P.f(PV)[i], mv1 = mass * P.f(PV + 1)[i], mv2 = mass * P.f(PV + 2)[i];
        st[0] = make_float4(w[0], d[0], w[1], d[1]);
        st[1] = make_float4(w[2], d[2], w[3], d[3]);
        st[2] = make_float4(w[4], d[4], w[5], d[5]);
        st[3] = make_float4(w[6], d[6], w[7], d[7]);
        st[4] = make_float4(w[8], d[8], __int_as_float(cell), mass);
        st[5] = make_float4(mv0, mv1, mv2, A.m[8]);
        st[6] = make_float4(A.m[0], A.m[1], A.m[2], A.m[3]);
        st[7] = make_float4(A.m[4], A.m[5], A.m[6], A.m[7]);
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) st[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      // runs of equal cells inside this chunk (the run is sorted by cell): heads as a warp-uniform mask
      const int prev_cell = __shfl_up_sync(SVB_FULL, cell, 1);
      uint32_t heads = __ballot_sync(SVB_FULL, cell >= 0 && (lane == 0 || prev_cell != cell));
      const uint32_t valid = __ballot_sync(SVB_FULL, cell >= 0);
      __syncwarp();
      float4* row = reinterpret_cast<float4*>(stage + lane * STAGE_STRIDE);
#pragma unroll
      for (int q = 0; q < 8; ++q) row[q] = st[q];
      __syncwarp();

      // ---- cooperative walk: lane = stencil node, one register accumulator per run of equal cells
      const int count = __popc(valid);
      while (heads) {
        const int first = __ffs(heads) - 1;
        heads &= heads - 1;
        const int last = heads ? __ffs(heads) - 1 : count;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int off = first * STAGE_STRIDE;
        const float* px = lane_x + off;
        const float* py = lane_y + off;
        const float* pz = lane_z + off;
        const float* sp = stage + off;
#pragma unroll 4
        for (int p = first; p < last; ++p) {
          const float2 wx = *reinterpret_cast<const float2*>(px);
          const float2 wy = *reinterpret_cast<const float2*>(py);
          const float2 wz = *reinterpret_cast<const float2*>(pz);
          const float mass = sp[19];
          const float4 mvA = *reinterpret_cast<const float4*>(sp + 20);  // mv0 mv1 mv2 A8
          const float4 A0 = *reinterpret_cast<const float4*>(sp + 24);   // A0..A3
          const float4 A1 = *reinterpret_cast<const float4*>(sp + 28);   // A4..A7
          const float wgt = wx.x * wy.x * wz.x;
          const float m0 = mvA.x + (A0.x * wx.y + A0.w * wy.y + A1.z * wz.y);
          const float m1 = mvA.y + (A0.y * wx.y + A1.x * wy.y + A1.w * wz.y);
          const float m2 = mvA.z + (A0.z * wx.y + A1.y * wy.y + mvA.w * wz.y);
          acc.x += wgt * m0; acc.y += wgt * m1; acc.z += wgt * m2; acc.w += wgt * mass;
          px += STAGE_STRIDE; py += STAGE_STRIDE; pz += STAGE_STRIDE; sp += STAGE_STRIDE;
        }
        if (node_lane) {
          const int c = __float_as_int(stage[first * STAGE_STRIDE + 18]);
          const int t = ((c >> 4) * 6 + ((c >> 2) & 3)) * 6 + (c & 3) + lane_tile_off;
          float4 o = my_tile[t];
          o.x += acc.x; o.y += acc.y; o.z += acc.z; o.w += acc.w;
          my_tile[t] = o;
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
// meld (meld_grid.rs:16-69) evaluated while loading a tile: node value seen by layer `bits` =
// sum over the layers of the same block whose bits are compatible; velocity = momentum / mass.
__device__ __forceinline__ float4 melded_node(const float4* __restrict__ grid, const unsigned long long* __restrict__ active_keys, const uint32_t* __restrict__ layer_bits,
                                              uint32_t n_active, int nl, int e, int node) {
  float4 s = grid[(size_t)e * 64 + node];
  if (nl > 0) {
    const unsigned long long k = active_keys[e];
    const unsigned long long block = k >> nl;
    const uint32_t mask = (1u << nl) - 1u;
    const uint32_t mine = layer_bits[(uint32_t)k & mask];
    for (int o = e - 1; o >= 0; --o) {
      const unsigned long long ko = active_keys[o];
      if ((ko >> nl) != block) break;
      if (bits_compatible(mine, layer_bits[(uint32_t)ko & mask])) {
        const float4 v = grid[(size_t)o * 64 + node];
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
    }
    for (uint32_t o = e + 1; o < n_active; ++o) {
      const unsigned long long ko = active_keys[o];
      if ((ko >> nl) != block) break;
      if (bits_compatible(mine, layer_bits[(uint32_t)ko & mask])) {
        const float4 v = grid[(size_t)o * 64 + node];
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
    }
  }
  if (s.w > 0.f) { s.x /= s.w; s.y /= s.w; s.z /= s.w; }
  else { s.x = 0.f; s.y = 0.f; s.z = 0.f; }
  return s;
}

// G2P (+ advance, return mapping, energy, cull when FUSE).  One CTA per (block, layer) run; the
// melded 6x6x6 velocity tile is staged in shared memory, then one thread per particle gathers.
constexpr int G2P_THREADS = 128;
template <bool FUSE, bool REDUCE>
__global__ void __launch_bounds__(G2P_THREADS) k_g2p(ParticleBuf P, float* __restrict__ energy, const uint32_t* __restrict__ group_start, const int* __restrict__ nbr, StepScalars* S,
                                                     const float4* __restrict__ grid, const unsigned long long* __restrict__ active_keys, const uint32_t* __restrict__ layer_bits,
                                                     const BinLayout* __restrict__ Lp, SimConsts K, float dt) {
  __shared__ float4 tile[TILE_NODES];
  __shared__ uint32_t s_group;
  const uint32_t n_groups = S->n_groups, n_live = S->n_live, n_active = S->n_active;
  const int nl = Lp->nl;
  const float h = K.h;
  int red_vel = INT32_MIN, red_def = INT32_MAX;
  uint32_t failed = 0;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_group = atomicAdd(&S->work_counter[1], 1u);
    __syncthreads();
    const uint32_t g = s_group;
    if (g >= n_groups) break;
    const uint32_t start = group_start[g];
    const uint32_t end = g + 1 < n_groups ? group_start[g + 1] : n_live;
    for (int t = threadIdx.x; t < TILE_NODES; t += blockDim.x) {
      const int ti = t / 36, tj = (t / 6) % 6, tk = t % 6;
      const int d = (ti >> 2) | ((tj >> 2) << 1) | ((tk >> 2) << 2);
      const int nb = nbr[(size_t)g * 8 + d];
      tile[t] = nb >= 0 ? melded_node(grid, active_keys, layer_bits, n_active, nl, nb, ((ti & 3) << 4) | ((tj & 3) << 2) | (tk & 3)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    for (uint32_t i = start + threadIdx.x; i < end; i += blockDim.x) {
      V3 x = V3{P.f(PX)[i], P.f(PX + 1)[i], P.f(PX + 2)[i]};
      const V3 nrm = V3{__fdiv_rn(x.x, h), __fdiv_rn(x.y, h), __fdiv_rn(x.z, h)};
      const V3 shift = V3{floorf(__fsub_rn(nrm.x, 0.5f)), floorf(__fsub_rn(nrm.y, 0.5f)), floorf(__fsub_rn(nrm.z, 0.5f))};
      const int s0 = (int)shift.x, s1 = (int)shift.y, s2 = (int)shift.z;
      const V3 shifted = nrm - shift;
      float wx[3], wy[3], wz[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        wx[a] = kernel_quadratic(shifted.x - (float)a);
        wy[a] = kernel_quadratic(shifted.y - (float)a);
        wz[a] = kernel_quadratic(shifted.z - (float)a);
      }
      V3 v = V3{0.f, 0.f, 0.f};
      M3 C;
#pragma unroll
      for (int q = 0; q < 9; ++q) C.m[q] = 0.f;
      const int tb = ((s0 & 3) * 6 + (s1 & 3)) * 6 + (s2 & 3);
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float w = wx[a] * wy[b] * wz[c];
            const float4 gv = tile[tb + (a * 6 + b) * 6 + c];
            const V3 to_node = V3{(float)(s0 + a) * h - x.x, (float)(s1 + b) * h - x.y, (float)(s2 + c) * h - x.z};
            const V3 wv = V3{gv.x * w, gv.y * w, gv.z * w};
            v = v + wv;
            C.m[0] += wv.x * to_node.x; C.m[1] += wv.y * to_node.x; C.m[2] += wv.z * to_node.x;
            C.m[3] += wv.x * to_node.y; C.m[4] += wv.y * to_node.y; C.m[5] += wv.z * to_node.y;
            C.m[6] += wv.x * to_node.z; C.m[7] += wv.y * to_node.z; C.m[8] += wv.z * to_node.z;
          }
      const float cs = 4.f / h / h;
#pragma unroll
      for (int q = 0; q < 9; ++q) C.m[q] *= cs;
      P.f(PV)[i] = v.x; P.f(PV + 1)[i] = v.y; P.f(PV + 2)[i] = v.z;
#pragma unroll
      for (int q = 0; q < 9; ++q) P.f(PC + q)[i] = C.m[q];
      if (REDUCE) {
        red_vel = max(red_vel, total_key(norm(v)));
#pragma unroll
        for (int q = 0; q < 9; ++q) red_def = min(red_def, total_key(0.2f / fmaxf(fabsf(C.m[q]), 1e-8f)));
      }
      if (FUSE) {
        // advance_particles.rs:41-86, cull_particles.rs:31-39
        uint32_t flags = P.u(PFLAGS)[i];
        M3 F;
#pragma unroll
        for (int q = 0; q < 9; ++q) F.m[q] = P.f(PF + q)[i];
        x = x + v * dt;
        const M3 CF = mul(C, F);
#pragma unroll
        for (int q = 0; q < 9; ++q) F.m[q] += CF.m[q] * dt;
        float e;
        if (return_map_and_energy(flags, P.f(PP0)[i], P.f(PP1)[i], (flags & F_USE_SAND_ALPHA) ? P.f(PALPHA)[i] : 0.f, F, e)) energy[i] = e;
        else { flags |= F_FAILED; failed = 1; }
        const bool within = x.x > K.domain_min[0] && x.x < K.domain_max[0] && x.y > K.domain_min[1] && x.y < K.domain_max[1] && x.z > K.domain_min[2] && x.z < K.domain_max[2];
        if (!within) flags |= F_TOMBSTONED;
        P.f(PX)[i] = x.x; P.f(PX + 1)[i] = x.y; P.f(PX + 2)[i] = x.z;
#pragma unroll
        for (int q = 0; q < 9; ++q) P.f(PF + q)[i] = F.m[q];
        P.u(PFLAGS)[i] = flags;
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
  if (FUSE && failed) atomicOr(&S->status, 8u /*SVB_PARTICLE_CLOSE_TO_INVERTED*/);
}

// advance + cull as its own pass (adaptive time stepping: dt is only known after the G2P reductions)
__global__ void __launch_bounds__(256) k_advance(ParticleBuf P, float* __restrict__ energy, StepScalars* S, SimConsts K, uint32_t n, float dt) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t flags = P.u(PFLAGS)[i];
  if (flags & F_TOMBSTONED) return;
  V3 x = V3{P.f(PX)[i], P.f(PX + 1)[i], P.f(PX + 2)[i]};
  const V3 v = V3{P.f(PV)[i], P.f(PV + 1)[i], P.f(PV + 2)[i]};
  M3 C, F;
#pragma unroll
  for (int q = 0; q < 9; ++q) { C.m[q] = P.f(PC + q)[i]; F.m[q] = P.f(PF + q)[i]; }
  x = x + v * dt;
  const M3 CF = mul(C, F);
#pragma unroll
  for (int q = 0; q < 9; ++q) F.m[q] += CF.m[q] * dt;
  float e;
  if (return_map_and_energy(flags, P.f(PP0)[i], P.f(PP1)[i], (flags & F_USE_SAND_ALPHA) ? P.f(PALPHA)[i] : 0.f, F, e)) energy[i] = e;
  else { flags |= F_FAILED; atomicOr(&S->status, 8u); }
  const bool within = x.x > K.domain_min[0] && x.x < K.domain_max[0] && x.y > K.domain_min[1] && x.y < K.domain_max[1] && x.z > K.domain_min[2] && x.z < K.domain_max[2];
  if (!within) flags |= F_TOMBSTONED;
  P.f(PX)[i] = x.x; P.f(PX + 1)[i] = x.y; P.f(PX + 2)[i] = x.z;
#pragma unroll
  for (int q = 0; q < 9; ++q) P.f(PF + q)[i] = F.m[q];
  P.u(PFLAGS)[i] = flags;
}

// limit_time_step.rs:25-182: global minima of the sound-speed and isolated-particle bounds
__global__ void __launch_bounds__(256) k_limit_force(ParticleBuf P, StepScalars* S, float h, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  int ks = INT32_MAX, ki = INT32_MAX;
  uint32_t live = 0;
  if (i < n) {
    const uint32_t flags = P.u(PFLAGS)[i];
    if (!(flags & F_TOMBSTONED)) {
      M3 F;
#pragma unroll
      for (int q = 0; q < 9; ++q) F.m[q] = P.f(PF + q)[i];
      const ParticleLimits l = particle_time_step_limits((flags & F_IS_FLUID) != 0, P.f(PP0)[i], P.f(PP1)[i], P.f(PMASS)[i], P.f(PVOL)[i], F, h);
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
    atomicMin(&S->min_sound_key, ks);
    atomicMin(&S->min_isolated_key, ki);
    atomicAdd(&S->live_count, live);
  }
}

// ------------------------------------------------------------------------------------------------
// grid download helpers: which nodes of an active tile have >= 1 contributor
__global__ void __launch_bounds__(256) k_touch_nodes(ParticleBuf P, const uint32_t* __restrict__ group_start, const int* __restrict__ nbr, const StepScalars* __restrict__ S, float h,
                                                     unsigned long long* __restrict__ node_mask) {
  const uint32_t n_groups = S->n_groups, n_live = S->n_live;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < n_groups; g += warps) {
    const uint32_t start = group_start[g];
    const uint32_t end = g + 1 < n_groups ? group_start[g + 1] : n_live;
    for (uint32_t i = start + lane; i < end; i += 32) {
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
__global__ void __launch_bounds__(64) k_emit_grid(const float4* __restrict__ grid, const unsigned long long* __restrict__ active_keys, const uint32_t* __restrict__ layer_bits,
                                                  const BinLayout* __restrict__ Lp, const unsigned long long* __restrict__ node_mask, const uint32_t* __restrict__ node_offset,
                                                  uint32_t n_active, int32_t* __restrict__ node_ids, uint32_t* __restrict__ out_bits, float* __restrict__ masses, float* __restrict__ velocities) {
  const uint32_t e = blockIdx.x;
  const int node = threadIdx.x;
  const unsigned long long mask = node_mask[e];
  if (!((mask >> node) & 1ull)) return;
  const BinLayout L = *Lp;
  const uint32_t at = node_offset[e] + __popcll(mask & ((1ull << node) - 1ull));
  const unsigned long long k = active_keys[e];
  const uint32_t rank = (uint32_t)k & ((1u << L.nl) - 1u);
  unsigned long long b = k >> L.nl;
  const int bz = (int)(b & ((1ull << L.nb[2]) - 1ull)); b >>= L.nb[2];
  const int by = (int)(b & ((1ull << L.nb[1]) - 1ull)); b >>= L.nb[1];
  const int bx = (int)b;
  node_ids[3 * at] = ((bx + L.block_min[0]) << 2) + (node >> 4);
  node_ids[3 * at + 1] = ((by + L.block_min[1]) << 2) + ((node >> 2) & 3);
  node_ids[3 * at + 2] = ((bz + L.block_min[2]) << 2) + (node & 3);
  out_bits[at] = layer_bits[rank];
  const float4 v = melded_node(grid, active_keys, layer_bits, n_active, L.nl, (int)e, node);
  masses[at] = v.w;
  velocities[3 * at] = v.x; velocities[3 * at + 1] = v.y; velocities[3 * at + 2] = v.z;
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
__global__ void k_decode_active(const unsigned long long* __restrict__ active_keys, const uint32_t* __restrict__ layer_bits, const BinLayout* __restrict__ Lp, uint32_t n_active,
                                int32_t* __restrict__ block_ids, uint32_t* __restrict__ out_bits) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_active) return;
  const BinLayout L = *Lp;
  const unsigned long long k = active_keys[e];
  const uint32_t rank = (uint32_t)k & ((1u << L.nl) - 1u);
  unsigned long long b = k >> L.nl;
  const int bz = (int)(b & ((1ull << L.nb[2]) - 1ull)); b >>= L.nb[2];
  const int by = (int)(b & ((1ull << L.nb[1]) - 1ull)); b >>= L.nb[1];
  block_ids[3 * e] = (int)b + L.block_min[0];
  block_ids[3 * e + 1] = by + L.block_min[1];
  block_ids[3 * e + 2] = bz + L.block_min[2];
  out_bits[e] = layer_bits[rank];
}
__global__ void k_cells(ParticleBuf P, float h, uint32_t n, int32_t* __restrict__ cells) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  cells[3 * i] = base_node(P.f(PX)[i], h);
  cells[3 * i + 1] = base_node(P.f(PX + 1)[i], h);
  cells[3 * i + 2] = base_node(P.f(PX + 2)[i], h);
}

}  // namespace svb
