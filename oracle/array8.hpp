// TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.
//
// oracle/array8.hpp: an 8-lane value type that stands in for the reference's
//   fub::Array1d  = Eigen::Array<double, 1, kDefaultChunkSize = 8>
//   fub::MaskArray = Eigen::Array<bool, 1, 8>
// (reference: include/fub/ext/Eigen.hpp:36-61).  Every operator is a plain
// IEEE-754 per-lane operation, exactly what Eigen 3.3.9 array expressions
// evaluate to (no reassociation, no reciprocal tricks).  The oracle is built
// with -ffp-contract=off so that a*b+c is never fused.
//
// min/max follow the x86 packet semantics Eigen uses for Array::min/max
// (pmin = _mm*_min_pd(a, b): returns b when either operand is NaN); NaN inputs
// are outside the contract of the hot path (the reference asserts on them,
// FluxMethodAdapter.hpp:80,132).
#pragma once
#include <cmath>

namespace fvso {

constexpr int kChunk = 8;

struct M8 {
  bool v[kChunk];
};

struct A8 {
  double v[kChunk];
  A8() = default;
  explicit A8(double x) {
    for (int l = 0; l < kChunk; ++l) v[l] = x;
  }
};

#define FVSO_BINOP(op)                                                         \
  inline A8 operator op(const A8& a, const A8& b) {                            \
    A8 r;                                                                      \
    for (int l = 0; l < kChunk; ++l) r.v[l] = a.v[l] op b.v[l];                \
    return r;                                                                  \
  }                                                                            \
  inline A8 operator op(const A8& a, double b) {                               \
    A8 r;                                                                      \
    for (int l = 0; l < kChunk; ++l) r.v[l] = a.v[l] op b;                     \
    return r;                                                                  \
  }                                                                            \
  inline A8 operator op(double a, const A8& b) {                               \
    A8 r;                                                                      \
    for (int l = 0; l < kChunk; ++l) r.v[l] = a op b.v[l];                     \
    return r;                                                                  \
  }
FVSO_BINOP(+)
FVSO_BINOP(-)
FVSO_BINOP(*)
FVSO_BINOP(/)
#undef FVSO_BINOP

inline A8 operator-(const A8& a) {
  A8 r;
  for (int l = 0; l < kChunk; ++l) r.v[l] = -a.v[l];
  return r;
}
inline A8& operator+=(A8& a, const A8& b) { return a = a + b; }
inline A8& operator-=(A8& a, const A8& b) { return a = a - b; }
inline A8& operator*=(A8& a, const A8& b) { return a = a * b; }
inline A8& operator*=(A8& a, double b) { return a = a * b; }

#define FVSO_CMP(op)                                                           \
  inline M8 operator op(const A8& a, const A8& b) {                            \
    M8 r;                                                                      \
    for (int l = 0; l < kChunk; ++l) r.v[l] = a.v[l] op b.v[l];                \
    return r;                                                                  \
  }                                                                            \
  inline M8 operator op(const A8& a, double b) {                               \
    M8 r;                                                                      \
    for (int l = 0; l < kChunk; ++l) r.v[l] = a.v[l] op b;                     \
    return r;                                                                  \
  }
FVSO_CMP(>)
FVSO_CMP(<)
FVSO_CMP(>=)
FVSO_CMP(<=)
FVSO_CMP(==)
#undef FVSO_CMP

inline M8 operator&&(const M8& a, const M8& b) {
  M8 r;
  for (int l = 0; l < kChunk; ++l) r.v[l] = a.v[l] && b.v[l];
  return r;
}
inline M8 operator!(const M8& a) {
  M8 r;
  for (int l = 0; l < kChunk; ++l) r.v[l] = !a.v[l];
  return r;
}
inline bool any(const M8& m) {
  bool r = false;
  for (int l = 0; l < kChunk; ++l) r = r || m.v[l];
  return r;
}
inline bool any(bool m) { return m; }

// mask.select(a, b)
inline A8 select(const M8& m, const A8& a, const A8& b) {
  A8 r;
  for (int l = 0; l < kChunk; ++l) r.v[l] = m.v[l] ? a.v[l] : b.v[l];
  return r;
}
inline A8 select(const M8& m, const A8& a, double b) {
  A8 r;
  for (int l = 0; l < kChunk; ++l) r.v[l] = m.v[l] ? a.v[l] : b;
  return r;
}
inline A8 select(const M8& m, double a, const A8& b) {
  A8 r;
  for (int l = 0; l < kChunk; ++l) r.v[l] = m.v[l] ? a : b.v[l];
  return r;
}
inline double select(bool m, double a, double b) { return m ? a : b; }

inline A8 vsqrt(const A8& a) {
  A8 r;
  for (int l = 0; l < kChunk; ++l) r.v[l] = std::sqrt(a.v[l]);
  return r;
}
inline double vsqrt(double a) { return std::sqrt(a); }

inline A8 vabs(const A8& a) {
  A8 r;
  for (int l = 0; l < kChunk; ++l) r.v[l] = std::fabs(a.v[l]);
  return r;
}
inline double vabs(double a) { return std::fabs(a); }

inline double vmin(double a, double b) { return a < b ? a : b; }
inline double vmax(double a, double b) { return a > b ? a : b; }
inline A8 vmin(const A8& a, const A8& b) {
  A8 r;
  for (int l = 0; l < kChunk; ++l) r.v[l] = vmin(a.v[l], b.v[l]);
  return r;
}
inline A8 vmax(const A8& a, const A8& b) {
  A8 r;
  for (int l = 0; l < kChunk; ++l) r.v[l] = vmax(a.v[l], b.v[l]);
  return r;
}
inline A8 vmin(const A8& a, double b) { return vmin(a, A8(b)); }
inline A8 vmax(const A8& a, double b) { return vmax(a, A8(b)); }

inline double min_coeff(const A8& a) {
  double r = a.v[0];
  for (int l = 1; l < kChunk; ++l) r = vmin(r, a.v[l]);
  return r;
}

template <typename T> struct MaskOf;
template <> struct MaskOf<double> { using type = bool; };
template <> struct MaskOf<A8> { using type = M8; };
template <typename T> using mask_t = typename MaskOf<T>::type;

inline double splat(double, double x) { return x; }
inline A8 splat(const A8&, double x) { return A8(x); }

} // namespace fvso
