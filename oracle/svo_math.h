// ORACLE — test infrastructure only.  PARITY UNPINNED end-to-end (see oracle/README.md):
// the reference's Rust CPU path cannot be built in this environment and the reference
// holds no test that pins a whole substep; the pieces below are pinned against the
// reference's own known-answer material (tests/test_oracle_*.py).
//
// CPU restatement of the leaf math of Algebraic-UG/squishy_volumes (reference paths are
// relative to /root/reference/rust/crates):
//   cpu/src/kernels.rs:17-49            quadratic B-spline + base-node shift
//   util/src/collider_bits.rs:9-38      collider side/near bits
//   util/src/consts.rs:11-14            eps constants
//   util/src/elastic.rs                 Neo-Hookean / weakly compressible fluid / viscosity
//   nalgebra 0.33.3 (not vendored)      Matrix3 determinant, try_normalize, angle, svd
//
// Everything is templated on the scalar so the property tests can run in f64 exactly as the
// reference's `type T = f64` under cfg(test) (util/src/lib.rs:27-30).
#pragma once
#include <cmath>
#include <cstdint>
#include <algorithm>
#include <limits>

namespace svo {

// ---- util/src/consts.rs:11-14
constexpr float NORMALIZATION_EPS = 1e-5f;
constexpr float INVERSE_EPS = 1e-7f;
constexpr float SINGULAR_VALUE_SEPARATION = 1e-5f;

// ---- file_frame/src/particles.rs:23-33
enum ParticleFlags : uint32_t {
  IS_SOLID = 1u << 0,
  IS_FLUID = 1u << 1,
  USE_VISCOSITY = 1u << 2,
  USE_SAND_ALPHA = 1u << 3,
  HAS_GOAL = 1u << 4,
  TOMBSTONED = 1u << 5,
  FAILED = 1u << 6,
};

template <class T>
struct Vec3 {
  T x{}, y{}, z{};
  Vec3() = default;
  Vec3(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
  T& operator[](int i) { return (&x)[i]; }
  const T& operator[](int i) const { return (&x)[i]; }
  Vec3 operator+(const Vec3& o) const { return {x + o.x, y + o.y, z + o.z}; }
  Vec3 operator-(const Vec3& o) const { return {x - o.x, y - o.y, z - o.z}; }
  Vec3 operator-() const { return {-x, -y, -z}; }
  Vec3 operator*(T s) const { return {x * s, y * s, z * s}; }
  Vec3 operator/(T s) const { return {x / s, y / s, z / s}; }
  Vec3& operator+=(const Vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
  Vec3& operator-=(const Vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
  Vec3& operator*=(T s) { x *= s; y *= s; z *= s; return *this; }
  Vec3& operator/=(T s) { x /= s; y /= s; z /= s; return *this; }
  bool operator==(const Vec3& o) const { return x == o.x && y == o.y && z == o.z; }
  T dot(const Vec3& o) const { return x * o.x + y * o.y + z * o.z; }
  Vec3 cross(const Vec3& o) const {
    return {y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x};
  }
  T norm_squared() const { return x * x + y * y + z * z; }
  T norm() const { return std::sqrt(norm_squared()); }
  T product() const { return x * y * z; }
  T sum() const { return x + y + z; }
  static Vec3 zeros() { return {T(0), T(0), T(0)}; }
  static Vec3 repeat(T v) { return {v, v, v}; }
};
template <class T>
inline Vec3<T> operator*(T s, const Vec3<T>& v) { return v * s; }

using Vec3f = Vec3<float>;
using Vec3i = Vec3<int32_t>;

// nalgebra try_normalize(min_norm): None when norm <= min_norm.
template <class T>
inline bool try_normalize(const Vec3<T>& v, T min_norm, Vec3<T>& out) {
  T n = v.norm();
  if (n <= min_norm) return false;
  out = v / n;
  return true;
}
template <class T>
inline Vec3<T> normalize_or_zero(const Vec3<T>& v, T min_norm) {
  Vec3<T> o;
  return try_normalize(v, min_norm, o) ? o : Vec3<T>::zeros();
}
// nalgebra Matrix::angle: 0 for zero vectors, clamped acos otherwise.
template <class T>
inline T angle(const Vec3<T>& a, const Vec3<T>& b) {
  T prod = a.dot(b);
  T n1 = a.norm(), n2 = b.norm();
  if (n1 == T(0) || n2 == T(0)) return T(0);
  T c = prod / (n1 * n2);
  c = c < T(-1) ? T(-1) : (c > T(1) ? T(1) : c);
  return std::acos(c);
}

// Column-major 3x3 exactly like nalgebra's Matrix3 (and the wire format [[f32;3];3] = 3 columns,
// file_frame/src/particles.rs:103-106).  m[c*3+r].
template <class T>
struct Mat3 {
  T m[9]{};
  T& operator()(int r, int c) { return m[c * 3 + r]; }
  const T& operator()(int r, int c) const { return m[c * 3 + r]; }
  static Mat3 zeros() { return Mat3{}; }
  static Mat3 identity() { Mat3 a; a(0, 0) = a(1, 1) = a(2, 2) = T(1); return a; }
  static Mat3 from_diagonal(const Vec3<T>& d) { Mat3 a; a(0, 0) = d.x; a(1, 1) = d.y; a(2, 2) = d.z; return a; }
  static Mat3 from_diagonal_element(T d) { return from_diagonal({d, d, d}); }
  static Mat3 from_columns(const Vec3<T>& a, const Vec3<T>& b, const Vec3<T>& c) {
    Mat3 r;
    for (int i = 0; i < 3; ++i) { r(i, 0) = a[i]; r(i, 1) = b[i]; r(i, 2) = c[i]; }
    return r;
  }
  Vec3<T> column(int c) const { return {m[c * 3], m[c * 3 + 1], m[c * 3 + 2]}; }
  Mat3 operator+(const Mat3& o) const { Mat3 r; for (int i = 0; i < 9; ++i) r.m[i] = m[i] + o.m[i]; return r; }
  Mat3 operator-(const Mat3& o) const { Mat3 r; for (int i = 0; i < 9; ++i) r.m[i] = m[i] - o.m[i]; return r; }
  Mat3 operator*(T s) const { Mat3 r; for (int i = 0; i < 9; ++i) r.m[i] = m[i] * s; return r; }
  Mat3& operator+=(const Mat3& o) { for (int i = 0; i < 9; ++i) m[i] += o.m[i]; return *this; }
  Mat3& operator*=(T s) { for (int i = 0; i < 9; ++i) m[i] *= s; return *this; }
  Mat3 operator*(const Mat3& o) const {
    Mat3 r;
    for (int c = 0; c < 3; ++c)
      for (int row = 0; row < 3; ++row) {
        T acc = T(0);
        for (int k = 0; k < 3; ++k) acc += (*this)(row, k) * o(k, c);
        r(row, c) = acc;
      }
    return r;
  }
  Vec3<T> operator*(const Vec3<T>& v) const {
    Vec3<T> r;
    for (int row = 0; row < 3; ++row) r[row] = (*this)(row, 0) * v.x + (*this)(row, 1) * v.y + (*this)(row, 2) * v.z;
    return r;
  }
  Mat3 transpose() const { Mat3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = (*this)(j, i); return r; }
  T trace() const { return m[0] + m[4] + m[8]; }
  T norm_squared() const { T s = T(0); for (int i = 0; i < 9; ++i) s += m[i] * m[i]; return s; }
  // nalgebra 0.33 Matrix3 determinant (base/matrix.rs, 3x3 arm): cofactor expansion along row 1.
  T determinant() const {
    const T m11 = (*this)(0, 0), m12 = (*this)(0, 1), m13 = (*this)(0, 2);
    const T m21 = (*this)(1, 0), m22 = (*this)(1, 1), m23 = (*this)(1, 2);
    const T m31 = (*this)(2, 0), m32 = (*this)(2, 1), m33 = (*this)(2, 2);
    const T minor_m12_m23 = m22 * m33 - m32 * m23;
    const T minor_m11_m23 = m21 * m33 - m31 * m23;
    const T minor_m11_m22 = m21 * m32 - m31 * m22;
    return m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22;
  }
};
template <class T>
inline Mat3<T> operator*(T s, const Mat3<T>& a) { return a * s; }
template <class T>
inline Mat3<T> outer(const Vec3<T>& a, const Vec3<T>& b) {
  Mat3<T> r;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = a[i] * b[j];
  return r;
}
using Mat3f = Mat3<float>;

// Rust's f32::powi lowers to the llvm.powi intrinsic = compiler-rt __powisf2 (square & multiply).
template <class T>
inline T powi(T a, int b) {
  const bool recip = b < 0;
  T r = T(1);
  while (true) {
    if (b & 1) r *= a;
    b /= 2;
    if (b == 0) break;
    a *= a;
  }
  return recip ? T(1) / r : r;
}

// ---- cpu/src/kernels.rs:17-26
template <class T>
inline T kernel_quadratic(T x) {
  x = std::fabs(x);
  if (x < T(1) / T(2)) return T(3) / T(4) - x * x;
  if (x < T(3) / T(2)) return T(1) / T(2) * (T(3) / T(2) - x) * (T(3) / T(2) - x);
  return T(0);
}
// cpu/src/kernels.rs:10-13 / :28-38 (unused by the path; kept for the known-answer test)
template <class T>
inline T kernel_linear(T x) { x = std::fabs(x); return x < T(1) ? T(1) - x : T(0); }
template <class T>
inline T kernel_cubic(T x) {
  x = std::fabs(x);
  if (x < T(1)) return T(1) / T(2) * x * x * x - x * x + T(2) / T(3);
  if (x < T(2)) return T(1) / T(6) * (T(2) - x) * (T(2) - x) * (T(2) - x);
  return T(0);
}
// ---- cpu/src/kernels.rs:46-49  (also the sort key of cpu/src/phase/sort.rs:29-33)
inline Vec3i position_to_shift_quadratic(const Vec3f& p, float h) {
  return {(int32_t)std::floor(p.x / h - 0.5f), (int32_t)std::floor(p.y / h - 0.5f), (int32_t)std::floor(p.z / h - 0.5f)};
}

// ---- util/src/collider_bits.rs:9-38
namespace collider_bits {
inline bool near_(uint32_t bits, unsigned c) { return (bits & (0x00010000u << c)) != 0; }
inline bool side(uint32_t bits, unsigned c) { return (bits & (0x00000001u << c)) != 0; }
// returns -1 = None, 0 = Some(false), 1 = Some(true)
inline int get(uint32_t bits, unsigned c) { return near_(bits, c) ? (side(bits, c) ? 1 : 0) : -1; }
inline void set(uint32_t& bits, unsigned c, int s) {
  bits &= ~(0x00010001u << c);
  if (s == 1) bits |= 0x00010001u << c;
  else if (s == 0) bits |= 0x00010000u << c;
}
inline bool compatible(uint32_t a, uint32_t b) {
  const uint32_t mask = (a & b) >> 16;
  const uint32_t diff = a ^ b;
  return (mask & diff) == 0;
}
}  // namespace collider_bits

// ---- util/src/elastic.rs
template <class T> inline T lame_mu(T E, T nu) { return E / T(2) / (T(1) + nu); }                       // :52-56
template <class T> inline T lame_lambda(T E, T nu) { return E * nu / (T(1) + nu) / (T(1) - T(2) * nu); }  // :60-64
template <class T> inline T invariant_2(const Mat3<T>& F) { return F.norm_squared(); }                     // :80-82
template <class T> inline T invariant_3(const Mat3<T>& F) { return F.determinant(); }                      // :103-105
template <class T> inline Mat3<T> partial_invariant_2_by_F(const Mat3<T>& F) { return T(2) * F; }          // :92-94
// :115-120  columns (c1 x c2, c2 x c0, c0 x c1) = cofactor matrix
template <class T>
inline Mat3<T> partial_invariant_3_by_F(const Mat3<T>& F) {
  auto c = [&](int i) { return F.column(i); };
  return Mat3<T>::from_columns(c(1).cross(c(2)), c(2).cross(c(0)), c(0).cross(c(1)));
}
template <class T> inline Vec3<T> partial_invariant_2_by_svd(const Vec3<T>& s) { return T(2) * s; }       // :97-99
template <class T> inline Vec3<T> partial_invariant_3_by_svd(const Vec3<T>& s) { return {s.y * s.z, s.x * s.z, s.x * s.y}; }  // :123-129
template <class T>
inline Mat3<T> double_partial_invariant_3_by_svd(const Vec3<T>& s) {  // :132-141
  Mat3<T> r;  // from_column_slice [0,z,y, z,0,x, y,x,0]
  const T v[9] = {T(0), s.z, s.y, s.z, T(0), s.x, s.y, s.x, T(0)};
  for (int i = 0; i < 9; ++i) r.m[i] = v[i];
  return r;
}
// :246-257 energy by invariants (asserts invariant_3 > 0 in the reference)
template <class T>
inline T elastic_energy_neo_hookean_by_invariants(T mu, T lambda, T i2, T i3) {
  return mu / T(2) * (i2 - T(3)) - mu * std::log(i3) + lambda / T(2) * powi(std::log(i3), 2);
}
template <class T>
inline T elastic_energy_neo_hookean(T mu, T lambda, const Mat3<T>& F) {  // :259-267
  return elastic_energy_neo_hookean_by_invariants(mu, lambda, invariant_2(F), invariant_3(F));
}
// :270-283  returns false = EnergyError::PositionGradientNonPositive
template <class T>
inline bool try_elastic_energy_neo_hookean(T mu, T lambda, const Mat3<T>& F, T& out) {
  const T i3 = invariant_3(F);
  if (!(i3 > T(0))) return false;
  out = elastic_energy_neo_hookean_by_invariants(mu, lambda, invariant_2(F), i3);
  return true;
}
template <class T> inline T d_nh_by_i2(T mu) { return mu / T(2); }                                          // :287-289
template <class T> inline T d_nh_by_i3(T mu, T lambda, T i3) { return (lambda * std::log(i3) - mu) / i3; }   // :293-295
template <class T> inline T dd_nh_by_i3(T mu, T lambda, T i3) {                                             // :299-306
  return (lambda * (T(1) - std::log(i3)) + mu) / powi(i3, 2);
}
template <class T>
inline Mat3<T> first_piola_stress_neo_hookean(T mu, T lambda, const Mat3<T>& F) {  // :310-322
  return d_nh_by_i2(mu) * partial_invariant_2_by_F(F) + d_nh_by_i3(mu, lambda, invariant_3(F)) * partial_invariant_3_by_F(F);
}
template <class T>
inline Vec3<T> first_piola_stress_neo_hookean_svd_diag(T mu, T lambda, const Vec3<T>& s) {  // :325-333
  return d_nh_by_i2(mu) * partial_invariant_2_by_svd(s) + d_nh_by_i3(mu, lambda, s.product()) * partial_invariant_3_by_svd(s);
}
template <class T>
inline Mat3<T> second_derivative_neo_hookean_svd_diag(T mu, T lambda, const Vec3<T>& s) {  // :336-351
  const Vec3<T> g = partial_invariant_3_by_svd(s);
  return Mat3<T>::from_diagonal_element(d_nh_by_i2(mu) * T(2)) + dd_nh_by_i3(mu, lambda, s.product()) * outer(g, g) +
         d_nh_by_i3(mu, lambda, s.product()) * double_partial_invariant_3_by_svd(s);
}
// :554-561
template <class T>
inline T elastic_energy_inviscid_by_invariant(T K, int exponent, T i3) {
  const T at_rest = K * (T(1) - T(1) / (T(1) - (T)exponent));
  return K * (i3 - powi(i3, 1 - exponent) / (T(1) - (T)exponent)) - at_rest;
}
template <class T> inline T d_inviscid_by_i3(T K, int exponent, T i3) { return K * (T(1) - T(1) / powi(i3, exponent)); }        // :564-570
template <class T> inline T dd_inviscid_by_i3(T K, int exponent, T i3) { return (T)exponent * K / powi(i3, exponent + 1); }     // :573-579
template <class T>
inline T elastic_energy_inviscid(T K, int exponent, const Mat3<T>& F) {  // :582-588
  return elastic_energy_inviscid_by_invariant(K, exponent, invariant_3(F));
}
template <class T>
inline Mat3<T> first_piola_stress_inviscid(T K, int exponent, const Mat3<T>& F) {  // :591-601
  return d_inviscid_by_i3(K, exponent, invariant_3(F)) * partial_invariant_3_by_F(F);
}
template <class T>
inline Vec3<T> first_piola_stress_inviscid_svd_diag(T K, int exponent, const Vec3<T>& s) {  // :640-647
  return d_inviscid_by_i3(K, exponent, s.product()) * partial_invariant_3_by_svd(s);
}
template <class T>
inline Mat3<T> second_derivative_inviscid_svd_diag(T K, int exponent, const Vec3<T>& s) {  // :650-666
  const Vec3<T> g = partial_invariant_3_by_svd(s);
  return dd_inviscid_by_i3(K, exponent, s.product()) * outer(g, g) + d_inviscid_by_i3(K, exponent, s.product()) * double_partial_invariant_3_by_svd(s);
}
// :669-688
template <class T>
inline Mat3<T> cauchy_stress_general_viscosity(T dynamic, T bulk, const Mat3<T>& C) {
  const Mat3<T> rate = T(0.5) * (C + C.transpose());
  return T(2) * dynamic * rate + bulk * Mat3<T>::from_diagonal_element(C.trace());
}

// ---- 3x3 SVD.  nalgebra's Matrix3::svd (Golub-Kahan in f32) is not vendored; the path only
// consumes sorted non-negative singular values, U*diag(f(s))*V^T and U*V^T, which are invariant
// to the SVD's sign/order conventions, so any accurate SVD reproduces them.  This
// one runs a cyclic Jacobi eigen-solve of F^T F in double and re-orthogonalises U = F V / s.
struct Svd3d {
  double U[9], S[3], V[9];  // column-major, F = U diag(S) V^T, S descending, S >= 0
};
inline void svd3(const double* F /*col-major*/, Svd3d& out) {
  auto at = [](const double* a, int r, int c) { return a[c * 3 + r]; };
  double A[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += at(F, k, i) * at(F, k, j);
      A[i][j] = s;
    }
  double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    double diag = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
    if (off <= 1e-60 || off <= 1e-34 * diag) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (A[p][q] == 0.0) continue;
        double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {  // A <- A J
          double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {  // A <- J^T A
          double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int idx[3] = {0, 1, 2};
  double lam[3] = {A[0][0], A[1][1], A[2][2]};
  std::sort(idx, idx + 3, [&](int a, int b) { return lam[a] > lam[b]; });
  double Vs[3][3], B[3][3];
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) Vs[r][c] = V[r][idx[c]];
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += at(F, r, k) * Vs[k][c];
      B[r][c] = s;
    }
  // Gram-Schmidt on the columns of B (largest first) -> U, S
  double Um[3][3];
  double S[3];
  for (int c = 0; c < 3; ++c) {
    double v[3] = {B[0][c], B[1][c], B[2][c]};
    for (int pc = 0; pc < c; ++pc) {
      double d = v[0] * Um[0][pc] + v[1] * Um[1][pc] + v[2] * Um[2][pc];
      for (int r = 0; r < 3; ++r) v[r] -= d * Um[r][pc];
    }
    double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    double ref = std::sqrt(B[0][0] * B[0][0] + B[1][0] * B[1][0] + B[2][0] * B[2][0]);
    if (n > 1e-150 && n > 1e-14 * ref) {
      for (int r = 0; r < 3; ++r) Um[r][c] = v[r] / n;
      S[c] = n;
    } else {  // rank deficient: complete the basis
      S[c] = 0;
      double e[3];
      if (c == 0) { e[0] = 1; e[1] = 0; e[2] = 0; }
      else if (c == 1) {
        int k = std::fabs(Um[0][0]) < 0.6 ? 0 : 1;
        double a[3] = {0, 0, 0};
        a[k] = 1;
        double d = a[0] * Um[0][0] + a[1] * Um[1][0] + a[2] * Um[2][0];
        for (int r = 0; r < 3; ++r) e[r] = a[r] - d * Um[r][0];
        double en = std::sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
        for (int r = 0; r < 3; ++r) e[r] /= en;
      } else {
        e[0] = Um[1][0] * Um[2][1] - Um[2][0] * Um[1][1];
        e[1] = Um[2][0] * Um[0][1] - Um[0][0] * Um[2][1];
        e[2] = Um[0][0] * Um[1][1] - Um[1][0] * Um[0][1];
      }
      for (int r = 0; r < 3; ++r) Um[r][c] = e[r];
    }
  }
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) {
      out.U[c * 3 + r] = Um[r][c];
      out.V[c * 3 + r] = Vs[r][c];
    }
  for (int c = 0; c < 3; ++c) out.S[c] = S[c];
}

struct Svd3f {
  Mat3f u, v_t;
  Vec3f s;
  Mat3f recompose() const { return u * Mat3f::from_diagonal(s) * v_t; }
};
inline Svd3f svd3f(const Mat3f& F) {
  double Fd[9];
  for (int i = 0; i < 9; ++i) Fd[i] = F.m[i];
  Svd3d d;
  svd3(Fd, d);
  Svd3f r;
  for (int i = 0; i < 9; ++i) r.u.m[i] = (float)d.U[i];
  for (int c = 0; c < 3; ++c)
    for (int row = 0; row < 3; ++row) r.v_t(c, row) = (float)d.V[c * 3 + row];  // V^T(c,row) = V(row,c)
  r.s = {(float)d.S[0], (float)d.S[1], (float)d.S[2]};
  return r;
}

}  // namespace svo
