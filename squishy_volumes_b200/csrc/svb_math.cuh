// svb_math.cuh — leaf math of the MPM substep, usable from device code and (for unit tests) host code.
//
// Restates, for registers instead of nalgebra types (reference paths relative to
// /root/reference/rust/crates):
//   cpu/src/kernels.rs:17-26,46-49        quadratic B-spline and base-node shift
//   util/src/collider_bits.rs:9-38        collider side / near bits
//   util/src/elastic.rs:287-322,564-601   Neo-Hookean and weakly compressible first Piola stress
//   util/src/elastic.rs:246-283,554-588   elastic energies
//   util/src/elastic.rs:669-688           viscous Cauchy stress
//   nalgebra Matrix3::svd call sites      cpu/src/phase/advance_particles.rs:51,81, limit_time_step.rs:45
// Matrices are column-major like nalgebra's Matrix3 and the wire format: m[c*3+r].
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SVB_HD __host__ __device__ __forceinline__
#else
#define SVB_HD inline
#endif

namespace svb {

// file_frame/src/particles.rs:23-33
enum : uint32_t {
  F_IS_SOLID = 1u << 0,
  F_IS_FLUID = 1u << 1,
  F_USE_VISCOSITY = 1u << 2,
  F_USE_SAND_ALPHA = 1u << 3,
  F_HAS_GOAL = 1u << 4,
  F_TOMBSTONED = 1u << 5,
  F_FAILED = 1u << 6,
  F_GONE = 1u << 30,  // internal (never leaves the device): the row migrated to a neighbour slab
};

// util/src/consts.rs:11-14
#define SVB_NORMALIZATION_EPS 1e-5f
#define SVB_SINGULAR_VALUE_SEPARATION 1e-5f

struct V3 {
  float x, y, z;
};
struct M3 {
  float m[9];
  SVB_HD float& operator()(int r, int c) { return m[c * 3 + r]; }
  SVB_HD float operator()(int r, int c) const { return m[c * 3 + r]; }
};

SVB_HD V3 v3(float x, float y, float z) { return V3{x, y, z}; }
SVB_HD V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
SVB_HD V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
SVB_HD V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
SVB_HD V3 operator*(float s, V3 a) { return V3{a.x * s, a.y * s, a.z * s}; }
SVB_HD V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
SVB_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
SVB_HD V3 cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
SVB_HD float norm(V3 a) { return sqrtf(dot(a, a)); }
SVB_HD bool is_zero(V3 a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f; }
// nalgebra try_normalize(min_norm): None when norm <= min_norm; the callers map None to zero.
SVB_HD V3 normalize_or_zero(V3 a, float min_norm) {
  const float n = norm(a);
  if (n <= min_norm) return V3{0.f, 0.f, 0.f};
  return a / n;
}
// nalgebra Matrix::angle: 0 when either vector is zero, clamped acos otherwise.
SVB_HD float angle_between(V3 a, V3 b) {
  const float prod = dot(a, b);
  const float n1 = norm(a), n2 = norm(b);
  if (n1 == 0.f || n2 == 0.f) return 0.f;
  float c = prod / (n1 * n2);
  c = c < -1.f ? -1.f : (c > 1.f ? 1.f : c);
  return acosf(c);
}

SVB_HD V3 col(const M3& a, int c) { return V3{a.m[c * 3], a.m[c * 3 + 1], a.m[c * 3 + 2]}; }
SVB_HD V3 mul(const M3& a, V3 v) {
  return V3{a.m[0] * v.x + a.m[3] * v.y + a.m[6] * v.z, a.m[1] * v.x + a.m[4] * v.y + a.m[7] * v.z,
            a.m[2] * v.x + a.m[5] * v.y + a.m[8] * v.z};
}
SVB_HD M3 mul(const M3& a, const M3& b) {
  M3 r;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int row = 0; row < 3; ++row)
      r.m[c * 3 + row] = a.m[row] * b.m[c * 3] + a.m[3 + row] * b.m[c * 3 + 1] + a.m[6 + row] * b.m[c * 3 + 2];
  return r;
}
// a * b^T
SVB_HD M3 mul_nt(const M3& a, const M3& b) {
  M3 r;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int row = 0; row < 3; ++row)
      r.m[c * 3 + row] = a.m[row] * b.m[c] + a.m[3 + row] * b.m[3 + c] + a.m[6 + row] * b.m[6 + c];
  return r;
}
// nalgebra 3x3 determinant: cofactor expansion along the first row.
SVB_HD float det(const M3& a) {
  const float m11 = a.m[0], m21 = a.m[1], m31 = a.m[2];
  const float m12 = a.m[3], m22 = a.m[4], m32 = a.m[5];
  const float m13 = a.m[6], m23 = a.m[7], m33 = a.m[8];
  const float minor_m12_m23 = m22 * m33 - m32 * m23;
  const float minor_m11_m23 = m21 * m33 - m31 * m23;
  const float minor_m11_m22 = m21 * m32 - m31 * m22;
  return m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22;
}
// util/src/elastic.rs:115-120 — d det / dF = cofactor matrix, columns (c1 x c2, c2 x c0, c0 x c1)
SVB_HD M3 cofactor(const M3& a) {
  const V3 c0 = col(a, 0), c1 = col(a, 1), c2 = col(a, 2);
  const V3 r0 = cross(c1, c2), r1 = cross(c2, c0), r2 = cross(c0, c1);
  M3 r;
  r.m[0] = r0.x; r.m[1] = r0.y; r.m[2] = r0.z;
  r.m[3] = r1.x; r.m[4] = r1.y; r.m[5] = r1.z;
  r.m[6] = r2.x; r.m[7] = r2.y; r.m[8] = r2.z;
  return r;
}
// Rust f32::powi = compiler-rt __powisf2 (square and multiply).
SVB_HD float powi(float a, int b) {
  const bool recip = b < 0;
  float r = 1.f;
  while (true) {
    if (b & 1) r *= a;
    b /= 2;
    if (b == 0) break;
    a *= a;
  }
  return recip ? 1.f / r : r;
}

// cpu/src/kernels.rs:17-26
SVB_HD float kernel_quadratic(float x) {
  x = fabsf(x);
  if (x < 0.5f) return 0.75f - x * x;
  if (x < 1.5f) return 0.5f * (1.5f - x) * (1.5f - x);
  return 0.f;
}

// util/src/collider_bits.rs:9-38 ; state: -1 = None, 0 = Some(false), 1 = Some(true)
SVB_HD int bits_get(uint32_t bits, unsigned c) {
  if (!(bits & (0x00010000u << c))) return -1;
  return (bits & (0x00000001u << c)) ? 1 : 0;
}
SVB_HD uint32_t bits_set(uint32_t bits, unsigned c, int s) {
  bits &= ~(0x00010001u << c);
  if (s == 1) bits |= 0x00010001u << c;
  else if (s == 0) bits |= 0x00010000u << c;
  return bits;
}
SVB_HD bool bits_compatible(uint32_t a, uint32_t b) { return ((((a & b) >> 16) & (a ^ b))) == 0u; }

// util/src/elastic.rs:293-295,310-322
SVB_HD M3 first_piola_neo_hookean(float mu, float lambda, const M3& F) {
  const float i3 = det(F);
  const float d3 = (lambda * logf(i3) - mu) / i3;
  const float d2 = mu / 2.f;
  const M3 cf = cofactor(F);
  M3 r;
#pragma unroll
  for (int i = 0; i < 9; ++i) r.m[i] = d2 * (2.f * F.m[i]) + d3 * cf.m[i];
  return r;
}
// util/src/elastic.rs:564-570,591-601
SVB_HD float d_inviscid_by_i3(float K, int exponent, float i3) { return K * (1.f - 1.f / powi(i3, exponent)); }
SVB_HD float dd_inviscid_by_i3(float K, int exponent, float i3) { return (float)exponent * K / powi(i3, exponent + 1); }
SVB_HD M3 first_piola_inviscid(float K, int exponent, const M3& F) {
  const float d3 = d_inviscid_by_i3(K, exponent, det(F));
  const M3 cf = cofactor(F);
  M3 r;
#pragma unroll
  for (int i = 0; i < 9; ++i) r.m[i] = d3 * cf.m[i];
  return r;
}
// util/src/elastic.rs:669-688
SVB_HD M3 viscous_cauchy(float dynamic, float bulk, const M3& C) {
  M3 r;
  const float tr = C.m[0] + C.m[4] + C.m[8];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int row = 0; row < 3; ++row) {
      const float rate = 0.5f * (C.m[c * 3 + row] + C.m[row * 3 + c]);
      r.m[c * 3 + row] = 2.f * dynamic * rate + (row == c ? bulk * tr : 0.f);
    }
  return r;
}
// util/src/elastic.rs:246-283: returns false for EnergyError::PositionGradientNonPositive
SVB_HD bool try_energy_neo_hookean(float mu, float lambda, const M3& F, float& out) {
  const float i3 = det(F);
  if (!(i3 > 0.f)) return false;
  float i2 = 0.f;
#pragma unroll
  for (int i = 0; i < 9; ++i) i2 += F.m[i] * F.m[i];
  const float l = logf(i3);
  out = mu / 2.f * (i2 - 3.f) - mu * l + lambda / 2.f * (l * l);
  return true;
}
// util/src/elastic.rs:554-561,582-588
SVB_HD float energy_inviscid(float K, int exponent, const M3& F) {
  const float i3 = det(F);
  const float one_minus = 1.f - (float)exponent;
  const float at_rest = K * (1.f - 1.f / one_minus);
  return K * (i3 - powi(i3, 1 - exponent) / one_minus) - at_rest;
}

// ---------------------------------------------------------------------------------------------
// 3x3 SVD in registers: one-sided (Hestenes) Jacobi on the columns of F.  F = U diag(s) V^T with
// s >= 0 sorted descending — the properties of nalgebra's Matrix3::svd the path relies on
// (only s, U f(s) V^T and U V^T are consumed, all convention-independent).
struct Svd3 {
  M3 u, v;  // v holds V (not V^T)
  V3 s;
};
// Two columns count as orthogonal when |cos| <= 3e-7, the level an f32 dot product resolves.  (A first version asked for 1e-8,
// below f32 rounding: the test never fired and every particle ran all 8 sweeps — 3.4x the elastic G2P time on sand.)
constexpr float SVD_ORTHOGONAL = 1e-13f;
SVB_HD void jacobi_pair(M3& b, M3& v, int p, int q) {
  const float a0 = b.m[p * 3], a1 = b.m[p * 3 + 1], a2 = b.m[p * 3 + 2];
  const float c0 = b.m[q * 3], c1 = b.m[q * 3 + 1], c2 = b.m[q * 3 + 2];
  const float alpha = a0 * a0 + a1 * a1 + a2 * a2;
  const float beta = c0 * c0 + c1 * c1 + c2 * c2;
  const float gamma = a0 * c0 + a1 * c1 + a2 * c2;
  if (gamma * gamma <= SVD_ORTHOGONAL * alpha * beta) return;  // already orthogonal to f32 precision
  const float zeta = (beta - alpha) / (2.f * gamma);
  const float t = copysignf(1.f, zeta) / (fabsf(zeta) + sqrtf(1.f + zeta * zeta));
  const float c = 1.f / sqrtf(1.f + t * t);
  const float s = c * t;
  b.m[p * 3] = c * a0 - s * c0; b.m[p * 3 + 1] = c * a1 - s * c1; b.m[p * 3 + 2] = c * a2 - s * c2;
  b.m[q * 3] = s * a0 + c * c0; b.m[q * 3 + 1] = s * a1 + c * c1; b.m[q * 3 + 2] = s * a2 + c * c2;
  const float v0 = v.m[p * 3], v1 = v.m[p * 3 + 1], v2 = v.m[p * 3 + 2];
  const float w0 = v.m[q * 3], w1 = v.m[q * 3 + 1], w2 = v.m[q * 3 + 2];
  v.m[p * 3] = c * v0 - s * w0; v.m[p * 3 + 1] = c * v1 - s * w1; v.m[p * 3 + 2] = c * v2 - s * w2;
  v.m[q * 3] = s * v0 + c * w0; v.m[q * 3 + 1] = s * v1 + c * w1; v.m[q * 3 + 2] = s * v2 + c * w2;
}
SVB_HD void swap_cols(M3& a, int p, int q) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float t = a.m[p * 3 + r];
    a.m[p * 3 + r] = a.m[q * 3 + r];
    a.m[q * 3 + r] = t;
  }
}
SVB_HD Svd3 svd3(const M3& F) {
  M3 b = F;
  M3 v;
#pragma unroll
  for (int i = 0; i < 9; ++i) v.m[i] = (i % 4 == 0) ? 1.f : 0.f;
#pragma unroll 1
  for (int sweep = 0; sweep < 8; ++sweep) {
    jacobi_pair(b, v, 0, 1);
    jacobi_pair(b, v, 0, 2);
    jacobi_pair(b, v, 1, 2);
    // converged when every pair passes jacobi_pair's own skip test
    const V3 b0 = col(b, 0), b1 = col(b, 1), b2 = col(b, 2);
    const float n0 = dot(b0, b0), n1 = dot(b1, b1), n2 = dot(b2, b2);
    const float g01 = dot(b0, b1), g02 = dot(b0, b2), g12 = dot(b1, b2);
    if (g01 * g01 <= SVD_ORTHOGONAL * n0 * n1 && g02 * g02 <= SVD_ORTHOGONAL * n0 * n2 && g12 * g12 <= SVD_ORTHOGONAL * n1 * n2) break;
  }
  float s0 = norm(col(b, 0)), s1 = norm(col(b, 1)), s2 = norm(col(b, 2));
  if (s0 < s1) { swap_cols(b, 0, 1); swap_cols(v, 0, 1); const float t = s0; s0 = s1; s1 = t; }
  if (s0 < s2) { swap_cols(b, 0, 2); swap_cols(v, 0, 2); const float t = s0; s0 = s2; s2 = t; }
  if (s1 < s2) { swap_cols(b, 1, 2); swap_cols(v, 1, 2); const float t = s1; s1 = s2; s2 = t; }
  Svd3 r;
  r.v = v;
  r.s = V3{s0, s1, s2};
  // U = B diag(1/s); rank-deficient columns are completed to an orthonormal basis.
  const float tiny = 1e-30f;
  V3 u0 = s0 > tiny ? col(b, 0) / s0 : V3{1.f, 0.f, 0.f};
  V3 u1;
  if (s1 > tiny && s1 > 1e-7f * s0) u1 = col(b, 1) / s1;
  else {
    const V3 a = fabsf(u0.x) < 0.6f ? V3{1.f, 0.f, 0.f} : V3{0.f, 1.f, 0.f};
    u1 = a - u0 * dot(a, u0);
    u1 = u1 / norm(u1);
  }
  V3 u2;
  if (s2 > tiny && s2 > 1e-7f * s0) u2 = col(b, 2) / s2;
  else u2 = cross(u0, u1);
  r.u.m[0] = u0.x; r.u.m[1] = u0.y; r.u.m[2] = u0.z;
  r.u.m[3] = u1.x; r.u.m[4] = u1.y; r.u.m[5] = u1.z;
  r.u.m[6] = u2.x; r.u.m[7] = u2.y; r.u.m[8] = u2.z;
  return r;
}
// singular values only (sorted descending): the same rotations as svd3 applied to B alone — bit-identical values without carrying V
// and without forming U (the time-step limits need nothing else, cpu/src/phase/limit_time_step.rs:45: svd(false, false))
SVB_HD void jacobi_pair_values(M3& b, int p, int q) {
  const float a0 = b.m[p * 3], a1 = b.m[p * 3 + 1], a2 = b.m[p * 3 + 2];
  const float c0 = b.m[q * 3], c1 = b.m[q * 3 + 1], c2 = b.m[q * 3 + 2];
  const float alpha = a0 * a0 + a1 * a1 + a2 * a2;
  const float beta = c0 * c0 + c1 * c1 + c2 * c2;
  const float gamma = a0 * c0 + a1 * c1 + a2 * c2;
  if (gamma * gamma <= SVD_ORTHOGONAL * alpha * beta) return;
  const float zeta = (beta - alpha) / (2.f * gamma);
  const float t = copysignf(1.f, zeta) / (fabsf(zeta) + sqrtf(1.f + zeta * zeta));
  const float c = 1.f / sqrtf(1.f + t * t);
  const float s = c * t;
  b.m[p * 3] = c * a0 - s * c0; b.m[p * 3 + 1] = c * a1 - s * c1; b.m[p * 3 + 2] = c * a2 - s * c2;
  b.m[q * 3] = s * a0 + c * c0; b.m[q * 3 + 1] = s * a1 + c * c1; b.m[q * 3 + 2] = s * a2 + c * c2;
}
SVB_HD V3 singular_values3(const M3& F) {
  M3 b = F;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int sweep = 0; sweep < 8; ++sweep) {
    jacobi_pair_values(b, 0, 1);
    jacobi_pair_values(b, 0, 2);
    jacobi_pair_values(b, 1, 2);
    const V3 b0 = col(b, 0), b1 = col(b, 1), b2 = col(b, 2);
    const float n0 = dot(b0, b0), n1 = dot(b1, b1), n2 = dot(b2, b2);
    const float g01 = dot(b0, b1), g02 = dot(b0, b2), g12 = dot(b1, b2);
    if (g01 * g01 <= SVD_ORTHOGONAL * n0 * n1 && g02 * g02 <= SVD_ORTHOGONAL * n0 * n2 && g12 * g12 <= SVD_ORTHOGONAL * n1 * n2) break;
  }
  float s0 = norm(col(b, 0)), s1 = norm(col(b, 1)), s2 = norm(col(b, 2));
  if (s0 < s1) { const float t = s0; s0 = s1; s1 = t; }
  if (s0 < s2) { const float t = s0; s0 = s2; s2 = t; }
  if (s1 < s2) { const float t = s1; s1 = s2; s2 = t; }
  return V3{s0, s1, s2};
}
// U diag(d) V^T
SVB_HD M3 recompose(const Svd3& s, V3 d) {
  M3 ud;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    ud.m[r] = s.u.m[r] * d.x;
    ud.m[3 + r] = s.u.m[3 + r] * d.y;
    ud.m[6 + r] = s.u.m[6 + r] * d.z;
  }
  return mul_nt(ud, s.v);
}

// murmur3_x86_32 of three little-endian u32 words (12 bytes, no tail) — the node key of the reference's dormant sort path
// (gpu/src/util.rs:79-100: the node id mapped to ordered u32 with x ^ 0x8000_0000, seed 0 or the collider bits; crate murmur3 0.5.2).
SVB_HD uint32_t ordered_u32(int32_t x) { return (uint32_t)x ^ 0x80000000u; }   // gpu/src/util.rs:71-73 (i32_to_u32_offset)
SVB_HD uint32_t murmur3_rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
SVB_HD uint32_t murmur3_32_words3(uint32_t a, uint32_t b, uint32_t c, uint32_t seed) {
  uint32_t h = seed;
  const uint32_t w[3] = {a, b, c};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int q = 0; q < 3; ++q) {
    uint32_t k = w[q] * 0xcc9e2d51u;
    k = murmur3_rotl(k, 15) * 0x1b873593u;
    h ^= k;
    h = murmur3_rotl(h, 13) * 5u + 0xe6546b64u;
  }
  h ^= 12u;   // length in bytes
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  return h;
}
SVB_HD uint32_t node_id_to_murmur(int32_t x, int32_t y, int32_t z, uint32_t seed) { return murmur3_32_words3(ordered_u32(x), ordered_u32(y), ordered_u32(z), seed); }

// f32::total_cmp as a signed-integer key, so min/max reductions can use integer atomics
// (cpu/src/phase/limit_time_step.rs uses total_cmp for every reduction).
SVB_HD int32_t total_key(float f) {
#if defined(__CUDA_ARCH__)
  int32_t b = __float_as_int(f);
#else
  int32_t b;
  memcpy(&b, &f, 4);
#endif
  return b ^ (int32_t)(((uint32_t)(b >> 31)) >> 1);
}
SVB_HD float total_unkey(int32_t k) {
  const int32_t b = k ^ (int32_t)(((uint32_t)(k >> 31)) >> 1);
#if defined(__CUDA_ARCH__)
  return __int_as_float(b);
#else
  float f;
  memcpy(&f, &b, 4);
  return f;
#endif
}

// ---- cpu/src/phase/limit_time_step.rs:37-182 : per-particle bounds for the adaptive time step
// util/src/elastic.rs:325-351,640-666 (SVD-space first/second derivatives)
struct ParticleLimits {
  float by_sound, by_isolated;
};
SVB_HD ParticleLimits particle_time_step_limits(bool is_fluid, float p0, float p1, float mass, float initial_volume,
                                                const M3& F, float h) {
  const V3 s = singular_values3(F);
  const float j = s.x * s.y * s.z;
  const bool xy_close = fabsf(s.x - s.y) < SVB_SINGULAR_VALUE_SEPARATION;
  const bool yz_close = fabsf(s.y - s.z) < SVB_SINGULAR_VALUE_SEPARATION;
  const bool zx_close = fabsf(s.z - s.x) < SVB_SINGULAR_VALUE_SEPARATION;
  const V3 g = V3{s.y * s.z, s.x * s.z, s.x * s.y};  // d(det)/ds
  V3 first;
  float d3, dd3, diag_add;
  if (!is_fluid) {
    const float mu = p0, lambda = p1;
    d3 = (lambda * logf(j) - mu) / j;
    dd3 = (lambda * (1.f - logf(j)) + mu) / (j * j);
    first = (mu / 2.f) * (2.f * s) + d3 * g;
    diag_add = mu / 2.f * 2.f;
  } else {
    const int e = (int)p1;
    d3 = d_inviscid_by_i3(p0, e, j);
    dd3 = dd_inviscid_by_i3(p0, e, j);
    first = d3 * g;
    diag_add = 0.f;
  }
  // second derivative wrt singular values: diag_add*I + dd3 * g g^T + d3 * [[0,z,y],[z,0,x],[y,x,0]]
  const float m11 = diag_add + dd3 * g.x * g.x;
  const float m22 = diag_add + dd3 * g.y * g.y;
  const float m33 = diag_add + dd3 * g.z * g.z;
  const float m21 = dd3 * g.y * g.x + d3 * s.z;
  const float m32 = dd3 * g.z * g.y + d3 * s.x;
  const float m13 = dd3 * g.x * g.z + d3 * s.y;
  float k[6];
  k[0] = s.x * s.x * m11;
  k[1] = s.y * s.y * m22;
  k[2] = s.z * s.z * m33;
  k[3] = s.y * s.y * (xy_close ? (first.x + s.x * m11 - s.y * m21) / 2.f / s.x : (s.x * first.x - s.y * first.y) / (s.x * s.x - s.y * s.y));
  k[4] = s.y * s.z * (yz_close ? (first.y + s.y * m22 - s.z * m32) / 2.f / s.y : (s.y * first.y - s.z * first.z) / (s.y * s.y - s.z * s.z));
  k[5] = s.z * s.x * (zx_close ? (first.z + s.x * m33 - s.x * m13) / 2.f / s.z : (s.z * first.z - s.x * first.x) / (s.z * s.z - s.x * s.x));
  int32_t kk = total_key(k[0]);
#pragma unroll
  for (int q = 1; q < 6; ++q) {
    const int32_t t = total_key(k[q]);
    kk = t > kk ? t : kk;
  }
  const float kappa = total_unkey(kk) / j;
  const float initial_density = mass / initial_volume;
  const float current_density = initial_density / j;
  ParticleLimits r;
  r.by_sound = h / sqrtf(kappa / current_density);
  if (!is_fluid) {
    const float xi = 3.f / h / h;
    r.by_isolated = sqrtf(mass / (initial_volume * xi * (1.f - 1.f / 2.f) * (p0 + 3.f / 2.f * p1)));
  } else {
    const int e = (int)p1;
    const float jj = det(F);
    const float fst = d_inviscid_by_i3(p0, e, jj);
    if (fabsf(jj - 1.f) > SVB_SINGULAR_VALUE_SEPARATION) r.by_isolated = h / jj * sqrtf(initial_density * (jj - 1.f) / (6.f * fst * 3.f));
    else r.by_isolated = h * sqrtf(initial_density / (6.f * dd_inviscid_by_i3(p0, e, jj) * 3.f));
  }
  return r;
}

// ---- cpu/src/phase/advance_particles.rs:47-86 : plasticity return mapping + energy.
// Returns false when the solid energy is undefined (det F <= 0  =>  FAILED).
SVB_HD bool return_map_and_energy(uint32_t flags, float p0, float p1, float sand_alpha, M3& F, float& energy) {
  if (!(flags & F_IS_FLUID)) {
    const float mu = p0, lambda = p1;
    if (flags & F_USE_SAND_ALPHA) {
      const Svd3 svd = svd3(F);
      const V3 e = V3{logf(svd.s.x), logf(svd.s.y), logf(svd.s.z)};
      const float e_tr = e.x + e.y + e.z;
      const V3 e_hat = e - V3{e_tr / 3.f, e_tr / 3.f, e_tr / 3.f};
      const float e_hat_norm = norm(e_hat);
      if (e_tr < 0.f && e_hat_norm > 0.f) {
        const float delta_gamma = e_hat_norm + (3.f * lambda + 2.f * mu) / 2.f / mu * e_tr * sand_alpha;
        if (delta_gamma > 0.f) {
          const V3 big_h = e - (delta_gamma / e_hat_norm) * e_hat;
          F = recompose(svd, V3{expf(big_h.x), expf(big_h.y), expf(big_h.z)});
        }
      } else {
        F = mul_nt(svd.u, svd.v);
      }
    }
    return try_energy_neo_hookean(mu, lambda, F, energy);
  }
  const Svd3 svd = svd3(F);
  const float iso = powf(svd.s.x * svd.s.y * svd.s.z, 1.f / 3.f);
  F = recompose(svd, V3{iso, iso, iso});
  energy = energy_inviscid(p0, (int)p1, F);
  return true;
}

}  // namespace svb
