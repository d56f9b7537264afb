// TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.
//
// oracle/numerics.hpp: CPU restatement of the reference's per-face / per-cell
// numerics for the PerfectGas Euler equations.  Every function cites the
// reference file:line it follows (paths relative to /root/reference).  The
// code is templated on the lane type T: T = A8 restates the reference's
// *array* (SIMD, default OpenMpSimdTag) code path, T = double restates the
// *scalar* path where the two are the same formula.
//
// Operation order is kept exactly as the reference writes it, because the
// parity contract is bit-exactness of the CUDA path against this file when
// both are built without FMA contraction.
//
// Third-party arithmetic restated from its published algorithm (dependency
// absent from /root/reference): Eigen 3.3.9 (conanfile.txt:3).
//  * Array expressions are per-lane IEEE operations.
//  * `momentum.matrix().colwise().squaredNorm()` on a column of D entries is a
//    completely unrolled non-vectorised reduction (Eigen/src/Core/Redux.h,
//    redux_novec_unroller): it splits the range in halves recursively, i.e.
//    D=1: x0;  D=2: x0+x1;  D=3: x0+(x1+x2).
#pragma once
#include "array8.hpp"

#include <array>
#include <cmath>
#include <limits>
#include <stdexcept>

namespace fvso {

// ---------------------------------------------------------------------------
// Equation and states.  include/fub/equations/PerfectGas.hpp:40-45,132-137
// (component order rho, momentum[D], energy, pressure, speed_of_sound) and
// :216-219 (gamma = 1.4, gamma_minus_1_inv = 1.0 / (gamma - 1.0)).
struct PerfectGasEq {
  double gamma{1.4};
  double gamma_minus_1_inv{1.0 / (gamma - 1.0)};
  explicit PerfectGasEq(double g = 1.4)
      : gamma(g), gamma_minus_1_inv(1.0 / (g - 1.0)) {}
};

template <typename T, int D> struct Cons {
  T rho;
  T mom[D];
  T E;
};
template <typename T, int D> struct Complete : Cons<T, D> {
  T p;
  T a;
};

template <typename T, int D, typename F>
inline void for_each_cons(F&& f, Cons<T, D>& a) {
  f(a.rho);
  for (int d = 0; d < D; ++d) f(a.mom[d]);
  f(a.E);
}

// Eigen redux_novec_unroller order (see header comment).
template <typename T, int D> inline T squared_norm(const T (&m)[D]) {
  if constexpr (D == 1) {
    return m[0] * m[0];
  } else if constexpr (D == 2) {
    return m[0] * m[0] + m[1] * m[1];
  } else {
    return m[0] * m[0] + (m[1] * m[1] + m[2] * m[2]);
  }
}

// src/equations/PerfectGas.cpp:31-37
template <typename T, int D>
inline T KineticEnergy(const T& density, const T (&momentum)[D]) {
  T squaredMomentum = squared_norm<T, D>(momentum);
  return 0.5 * squaredMomentum / density;
}

// src/equations/PerfectGas.cpp:39-49 (masked)
template <typename T, int D>
inline T KineticEnergy(const T& density, const T (&momentum)[D],
                       mask_t<T> mask) {
  mask = mask && (density > 0.0);
  T squaredMomentum = squared_norm<T, D>(momentum);
  T safe_density = select(mask, density, 1.0);
  return 0.5 * squaredMomentum / safe_density;
}

// src/equations/PerfectGas.cpp:66-77 (array) == :52-64 (scalar)
template <typename T, int D>
inline void Flux(const PerfectGasEq&, Cons<T, D>& flux,
                 const Complete<T, D>& state, int d0) {
  const T velocity = state.mom[d0] / state.rho;
  flux.rho = state.mom[d0];
  for (int d = 0; d < D; ++d) {
    flux.mom[d] = velocity * state.mom[d];
  }
  flux.mom[d0] = flux.mom[d0] + state.p;
  flux.E = velocity * (state.E + state.p);
}

// src/equations/PerfectGas.cpp:79-93 (masked)
template <typename T, int D>
inline void Flux(const PerfectGasEq&, Cons<T, D>& flux,
                 const Complete<T, D>& state, mask_t<T> mask, int d0) {
  mask = mask && (state.rho > 0.0);
  const T density = select(mask, state.rho, 1.0);
  const T velocity = state.mom[d0] / density;
  flux.rho = state.mom[d0];
  for (int d = 0; d < D; ++d) {
    flux.mom[d] = velocity * state.mom[d];
  }
  flux.mom[d0] = flux.mom[d0] + state.p;
  flux.E = velocity * (state.E + state.p);
}

// src/equations/PerfectGas.cpp:112-124 (array) == :96-110 (scalar)
template <typename T, int D>
inline void CompleteFromCons(const PerfectGasEq& eq, Complete<T, D>& complete,
                             const Cons<T, D>& cons) {
  // copy first: `complete` may alias `cons`
  const T rho = cons.rho;
  T mom[D];
  for (int d = 0; d < D; ++d) mom[d] = cons.mom[d];
  const T E = cons.E;
  complete.rho = rho;
  for (int d = 0; d < D; ++d) complete.mom[d] = mom[d];
  complete.E = E;
  const T e_kin = KineticEnergy<T, D>(rho, mom);
  const T e_int = E - e_kin;
  complete.p = e_int / eq.gamma_minus_1_inv;
  complete.a = vsqrt(eq.gamma * complete.p / complete.rho);
}

// src/equations/PerfectGas.cpp:126-144 (masked)
template <typename T, int D>
inline void CompleteFromCons(const PerfectGasEq& eq, Complete<T, D>& complete,
                             const Cons<T, D>& cons, mask_t<T> mask) {
  const T rho = cons.rho;
  T mom[D];
  for (int d = 0; d < D; ++d) mom[d] = cons.mom[d];
  const T E = cons.E;
  mask = mask && (rho > 0.0);
  complete.rho = select(mask, rho, 0.0);
  for (int d = 0; d < D; ++d) {
    complete.mom[d] = select(mask, mom[d], 0.0);
  }
  complete.E = select(mask, E, 0.0);
  const T e_kin = KineticEnergy<T, D>(rho, mom, mask);
  const T e_int = E - e_kin;
  complete.p = select(mask, e_int / eq.gamma_minus_1_inv, 0.0);
  T safe_density = select(mask, complete.rho, 1.0);
  complete.a = select(mask, vsqrt(eq.gamma * complete.p / safe_density), 0.0);
}

// src/equations/PerfectGas.cpp:146-158
template <int D>
inline Complete<double, D> CompleteFromPrim(const PerfectGasEq& eq, double rho,
                                            const double (&v)[D], double p) {
  Complete<double, D> q;
  q.rho = rho;
  for (int d = 0; d < D; ++d) q.mom[d] = rho * v[d];
  q.p = p;
  const double e_kin = KineticEnergy<double, D>(q.rho, q.mom);
  q.E = e_kin + p * eq.gamma_minus_1_inv;
  q.a = std::sqrt(eq.gamma * q.p / q.rho);
  return q;
}

// ---------------------------------------------------------------------------
// Slope limiters.  include/fub/flux_method/MusclHancockMethod.hpp:31-187
struct NoLimiter {
  template <typename T> T operator()(const T& qL, const T& qM, const T& qR) const {
    const T sL = qM - qL; // :49-51
    const T sR = qR - qM;
    return 0.5 * (sL + sR);
  }
  template <typename T>
  T operator()(const T& qL, const T& qM, const T& qR, mask_t<T> mask) const {
    const T sL = select(mask, qM - qL, 0.0); // :62-65
    const T sR = select(mask, qR - qM, 0.0);
    return 0.5 * (sL + sR);
  }
};

struct MinMod {
  static A8 lim(const A8& sL, const A8& sR) {
    // :117  (sR > 0).select(zero.max(sL.min(sR)), zero.min(sL.max(sR)))
    const A8 zero(0.0);
    return select(sR > 0.0, vmax(zero, vmin(sL, sR)), vmin(zero, vmax(sL, sR)));
  }
  static double lim(double sL, double sR) {
    // :100-104 (scalar path, std::min/std::max)
    if (sR > 0) {
      return std::max(0.0, std::min(sL, sR));
    } else {
      return std::min(0.0, std::max(sL, sR));
    }
  }
  template <typename T> T operator()(const T& qL, const T& qM, const T& qR) const {
    const T sL = qM - qL;
    const T sR = qR - qM;
    return lim(sL, sR);
  }
  template <typename T>
  T operator()(const T& qL, const T& qM, const T& qR, mask_t<T> mask) const {
    const T sL = select(mask, qM - qL, 0.0); // :129-131
    const T sR = select(mask, qR - qM, 0.0);
    return lim(sL, sR);
  }
};

struct VanLeer {
  static A8 lim(const A8& sL, const A8& sR) {
    A8 r = 2.0 * sL * sR; // :144-147
    return select(r > 0.0, r / (sL + sR), 0.0);
  }
  static double lim(double sL, double sR) {
    double c = 2.0 * sL * sR; // :138-141
    return (c > 0.0) ? c / (sL + sR) : 0.0;
  }
  template <typename T> T operator()(const T& qL, const T& qM, const T& qR) const {
    const T sL = qM - qL; // :181-183
    const T sR = qR - qM;
    return lim(sL, sR);
  }
  template <typename T>
  T operator()(const T& qL, const T& qM, const T& qR, mask_t<T> mask) const {
    const T sL = select(mask, qM - qL, 0.0); // :169-171
    const T sR = select(mask, qR - qM, 0.0);
    return lim(sL, sR);
  }
};

// ---------------------------------------------------------------------------
// Einfeldt signal velocities (normal Roe velocity only in roeA^2).
// src/equations/perfect_gas/EinfeldtSignalVelocities.hpp:64-97 (array),
// :25-62 (scalar, same formula).
template <typename T, int D>
inline std::array<T, 2> EinfeldtSignals(const PerfectGasEq& eq,
                                        const Complete<T, D>& left,
                                        const Complete<T, D>& right, int dir) {
  const T rhoL = left.rho;
  const T rhoR = right.rho;
  const T rhoUL = left.mom[dir];
  const T rhoUR = right.mom[dir];
  const T aL = left.a;
  const T aR = right.a;
  const T sqRhoL = vsqrt(rhoL);
  const T sqRhoR = vsqrt(rhoR);
  const T uL = rhoUL / rhoL;
  const T uR = rhoUR / rhoR;
  const T oosqRhoSum = 1.0 / (sqRhoL + sqRhoR);
  auto roeAvg = [&](const T& xL, const T& xR) {
    return (xL * sqRhoL + xR * sqRhoR) * oosqRhoSum;
  };
  const T roeU = roeAvg(uL, uR);
  const T hL = (left.E + left.p) / rhoL;
  const T hR = (right.E + right.p) / rhoR;
  const T roeH = roeAvg(hL, hR);
  const T roeA2 = (eq.gamma - 1.0) * (roeH - 0.5 * roeU * roeU);
  const T roeA = vsqrt(roeA2);
  const T sL1 = uL - aL;
  const T sL2 = roeU - roeA;
  const T sR1 = roeU + roeA;
  const T sR2 = uR + aR;
  return {vmin(sL1, sL2), vmax(sR1, sR2)};
}

// src/equations/perfect_gas/EinfeldtSignalVelocities.hpp:99-142 (masked; note
// the different Roe-average formula and the *unmasked* rho in hL, hR).
template <typename T, int D>
inline std::array<T, 2> EinfeldtSignals(const PerfectGasEq& eq,
                                        const Complete<T, D>& left,
                                        const Complete<T, D>& right,
                                        const mask_t<T>& mask, int dir) {
  const T rhoL = left.rho;
  const T rhoR = right.rho;
  const T rhoUL = select(mask, left.mom[dir], 0.0);
  const T rhoUR = select(mask, right.mom[dir], 0.0);
  const T aL = select(mask, left.a, 0.0);
  const T aR = select(mask, right.a, 0.0);
  const T rhoLs = select(mask, rhoL, 1.0);
  const T rhoRs = select(mask, rhoR, 1.0);
  const T sqRhoL = vsqrt(rhoLs);
  const T sqRhoR = vsqrt(rhoRs);
  const T sqRho = sqRhoL + sqRhoR;
  const T uL = rhoUL / rhoLs;
  const T uR = rhoUR / rhoRs;
  const T sqRhoL_over_sqRho = sqRhoL / sqRho;
  const T sqRhoR_over_sqRho = sqRhoR / sqRho;
  auto roeAvg = [&](const T& xL, const T& xR) {
    return (xL * sqRhoL_over_sqRho + xR * sqRhoR_over_sqRho);
  };
  const T roeU = roeAvg(uL, uR);
  const T rhoEL = select(mask, left.E, 0.0);
  const T rhoER = select(mask, right.E, 0.0);
  const T pL = select(mask, left.p, 0.0);
  const T pR = select(mask, right.p, 0.0);
  const T hL = (rhoEL + pL) / rhoL;
  const T hR = (rhoER + pR) / rhoR;
  const T roeH = roeAvg(hL, hR);
  const T roeA2 = (eq.gamma - 1.0) * (roeH - 0.5 * roeU * roeU);
  const T roeA = vsqrt(roeA2);
  const T sL1 = uL - aL;
  const T sL2 = roeU - roeA;
  const T sR1 = roeU + roeA;
  const T sR2 = uR + aR;
  return {vmin(sL1, sL2), vmax(sR1, sR2)};
}

// ---------------------------------------------------------------------------
// HLL with Einfeldt signals.  include/fub/flux_method/HllMethod.hpp
template <typename T, int D> struct HllEinfeldt {
  static constexpr int kWidth = 1;
  static constexpr int kDim = D;
  using Lane = T;
  PerfectGasEq eq;
  explicit HllEinfeldt(const PerfectGasEq& e) : eq(e) {}

  // :260-293 (array); scalar twin :177-205 is the same formula
  void ComputeNumericFlux(Cons<T, D>& nf, const Complete<T, D>* states,
                          double /*dt*/, double /*dx*/, int dir) const {
    const Complete<T, D>& left = states[0];
    const Complete<T, D>& right = states[1];
    const auto signals = EinfeldtSignals<T, D>(eq, left, right, dir);
    Cons<T, D> fL, fR;
    Flux<T, D>(eq, fL, left, dir);
    Flux<T, D>(eq, fR, right, dir);
    const T sL = vmin(signals[0], 0.0);
    const T sR = vmax(signals[1], 0.0);
    const T sLsR = sL * sR;
    T ds = sR - sL;
    ds = select(ds > 0.0, ds, 1.0);
    auto one = [&](const T& fl, const T& fr, const T& ql, const T& qr) {
      return (sR * fl - sL * fr + sLsR * (qr - ql)) / ds;
    };
    nf.rho = one(fL.rho, fR.rho, left.rho, right.rho);
    for (int d = 0; d < D; ++d) {
      nf.mom[d] = one(fL.mom[d], fR.mom[d], left.mom[d], right.mom[d]);
    }
    nf.E = one(fL.E, fR.E, left.E, right.E);
  }

  // :295-335 (masked array)
  void ComputeNumericFlux(Cons<T, D>& nf, const T& face_fraction,
                          const Complete<T, D>* states, const T* /*alphas*/,
                          double /*dt*/, double /*dx*/, int dir) const {
    const Complete<T, D>& left = states[0];
    const Complete<T, D>& right = states[1];
    const mask_t<T> mask = face_fraction > 0.0;
    const auto signals = EinfeldtSignals<T, D>(eq, left, right, mask, dir);
    Cons<T, D> fL, fR;
    Flux<T, D>(eq, fL, left, mask, dir);
    Flux<T, D>(eq, fR, right, mask, dir);
    const T sL = vmin(signals[0], 0.0);
    const T sR = vmax(signals[1], 0.0);
    const T sLsR = sL * sR;
    const T ds = sR - sL;
    const T safe_ds = select(mask, ds, 1.0);
    auto one = [&](const T& fl, const T& fr, const T& ql, const T& qr) {
      T qL_s = select(mask, ql, 0.0);
      T qR_s = select(mask, qr, 0.0);
      return (sR * fl - sL * fr + sLsR * (qR_s - qL_s)) / safe_ds;
    };
    nf.rho = one(fL.rho, fR.rho, left.rho, right.rho);
    for (int d = 0; d < D; ++d) {
      nf.mom[d] = one(fL.mom[d], fR.mom[d], left.mom[d], right.mom[d]);
    }
    nf.E = one(fL.E, fR.E, left.E, right.E);
  }

  // :337-347
  T ComputeStableDt(const Complete<T, D>* states, double dx, int dir) const {
    const auto signals = EinfeldtSignals<T, D>(eq, states[0], states[1], dir);
    T max = splat(signals[0], 0.0);
    max = vmax(max, vabs(signals[0]));
    max = vmax(max, vabs(signals[1]));
    return splat(max, dx) / max;
  }

  // :349-365 (masked)
  T ComputeStableDt(const Complete<T, D>* states, const T& face_fraction,
                    const T* /*alphas*/, double dx, int dir) const {
    const mask_t<T> mask = face_fraction > 0.0;
    const auto signals =
        EinfeldtSignals<T, D>(eq, states[0], states[1], mask, dir);
    T max = splat(signals[0], 0.0);
    max = vmax(max, vabs(signals[0]));
    max = vmax(max, vabs(signals[1]));
    const T safe_max = select(mask, max, 1.0);
    T stable_dts = splat(max, dx) / safe_max;
    return select(mask, stable_dts, std::numeric_limits<double>::max());
  }

  // :240-258 (array SolveRiemannProblem; no CompleteFromCons -- it is
  // commented out in the reference, so p and a of `sol` stay untouched)
  void SolveRiemannProblem(Complete<T, D>& sol, const Complete<T, D>& left,
                           const Complete<T, D>& right, int dir) const {
    const auto signals = EinfeldtSignals<T, D>(eq, left, right, dir);
    Cons<T, D> fL, fR;
    Flux<T, D>(eq, fL, left, dir);
    Flux<T, D>(eq, fR, right, dir);
    const T sL = signals[0];
    const T sR = signals[1];
    T ds = sR - sL;
    ds = select(ds > 0.0, ds, 1.0);
    auto one = [&](const T& fl, const T& fr, const T& ql, const T& qr) {
      T s = (sR * qr - sL * ql + fl - fr) / ds;
      return select(sL > 0.0, ql, select(sR < 0.0, qr, s));
    };
    sol.rho = one(fL.rho, fR.rho, left.rho, right.rho);
    for (int d = 0; d < D; ++d) {
      sol.mom[d] = one(fL.mom[d], fR.mom[d], left.mom[d], right.mom[d]);
    }
    sol.E = one(fL.E, fR.E, left.E, right.E);
  }
};

// ---------------------------------------------------------------------------
// HLLEM (array code path).  src/equations/perfect_gas/HllemMethod.hpp
// Msq = 1 and Minv = 1/sqrt(1) = 1 for PerfectGas
// (include/fub/equations/EulerEquation.hpp:53-62): the multiplications by
// Msq / Minv in the reference are exact identities and are omitted.
template <int D, bool Larrouturou> struct Hllem {
  static constexpr int kWidth = 1;
  static constexpr int kDim = D;
  using T = A8;
  using Lane = A8;
  PerfectGasEq eq;
  explicit Hllem(const PerfectGasEq& e) : eq(e) {}

  // :711-1072
  void ComputeNumericFlux(Cons<T, D>& flux, const Complete<T, D>* states,
                          double /*dt*/, double /*dx*/, int dir) const {
    Complete<T, D> left = states[0];
    Complete<T, D> right = states[1];
    const T rhoL = left.rho;
    const T rhoR = right.rho;
    T vL_[D];
    T vR_[D];
    T eKinRL(0.0);
    T eKinRR(0.0);
    if constexpr (Larrouturou) {
      // :729-745
      for (int i = 0; i < D; ++i) {
        const int dim_mask = (i == dir);
        vL_[i] = left.mom[i] / rhoL;
        vR_[i] = right.mom[i] / rhoR;
        eKinRL += double(1 - dim_mask) * vL_[i] * vL_[i];
        eKinRR += double(1 - dim_mask) * vR_[i] * vR_[i];
        left.mom[i] *= double(dim_mask);
        right.mom[i] *= double(dim_mask);
      }
      eKinRL *= 0.5;
      eKinRR *= 0.5;
      left.E -= rhoL * eKinRL;
      right.E -= rhoR * eKinRR;
    }
    Cons<T, D> fluxL, fluxR;
    Flux<T, D>(eq, fluxL, left, dir);
    Flux<T, D>(eq, fluxR, right, dir);

    // :752-797 Einfeldt signal velocities with the full Roe velocity
    const T aL = left.a;
    const T aR = right.a;
    const T rhoEL = left.E;
    const T rhoER = right.E;
    const T pL = left.p;
    const T pR = right.p;
    const T hL = (rhoEL + pL) / rhoL;
    const T hR = (rhoER + pR) / rhoR;
    const T sqRhoL = vsqrt(rhoL);
    const T sqRhoR = vsqrt(rhoR);
    const T sqRhoSum = sqRhoL + sqRhoR;
    const T sqRhoL_over_Sum = sqRhoL / sqRhoSum;
    const T sqRhoR_over_Sum = sqRhoR / sqRhoSum;
    T uL[D], uR[D], roeU[D];
    for (int i = 0; i < D; ++i) {
      uL[i] = left.mom[i] / rhoL;
      uR[i] = right.mom[i] / rhoR;
      roeU[i] = sqRhoL_over_Sum * uL[i] + sqRhoR_over_Sum * uR[i];
    }
    const T roeH = sqRhoL_over_Sum * hL + sqRhoR_over_Sum * hR;
    T squaredNormRoeU(0.0);
    for (int i = 0; i < D; ++i) {
      squaredNormRoeU += roeU[i] * roeU[i];
    }
    const T squaredNormRoeU_half = 0.5 * squaredNormRoeU;
    const T roeGamma = (sqRhoL * eq.gamma + sqRhoR * eq.gamma) / sqRhoSum;
    const T gm1 = roeGamma - 1.0;
    const T roeA2 = gm1 * (roeH - squaredNormRoeU_half);
    const T roeA = vsqrt(roeA2);
    const T maxA = vmax(aL, aR);
    const T sL1 = uL[dir] - maxA;
    const T sL2 = roeU[dir] - roeA;
    const T sR1 = roeU[dir] + roeA;
    const T sR2 = uR[dir] + maxA;
    const T sL = vmin(sL1, sL2);
    const T sR = vmax(sR1, sR2);

    // :799-811 HLLE flux
    const T bL = vmin(sL, 0.0);
    const T bR = vmax(sR, 0.0);
    const T bLbR = bL * bR;
    const T db = bR - bL;
    const T db_positive = select(db > 0.0, db, 1.0);
    auto hlle = [&](const T& fL, const T& fR, const T& qL, const T& qR) {
      return (bR * fL - bL * fR + bLbR * (qR - qL)) / db_positive;
    };
    Cons<T, D> flux_hlle;
    flux_hlle.rho = hlle(fluxL.rho, fluxR.rho, left.rho, right.rho);
    for (int d = 0; d < D; ++d) {
      flux_hlle.mom[d] = hlle(fluxL.mom[d], fluxR.mom[d], left.mom[d],
                              right.mom[d]);
    }
    flux_hlle.E = hlle(fluxL.E, fluxR.E, left.E, right.E);

    // :813-823
    const T u_bar = 0.5 * (sR + sL);
    const T u_bar_abs = vabs(u_bar);
    const T delta = roeA / (roeA + u_bar_abs);
    const T b = bLbR / db_positive;
    const T deltaRho = rhoR - rhoL;
    T deltaRhoU[D];
    for (int i = 0; i < D; ++i) deltaRhoU[i] = right.mom[i] - left.mom[i];
    const T deltaRhoE = rhoER - rhoEL;
    const T gm1_over_roeA2 = gm1 / roeA2;

    if constexpr (D == 1) {
      // :825-835
      const T l21 = gm1_over_roeA2 * (roeH - squaredNormRoeU);
      const T l22 = gm1_over_roeA2 * roeU[0];
      const T l23 = gm1_over_roeA2 * (-1.0);
      const T alpha_2 = l21 * deltaRho + l22 * deltaRhoU[0] + l23 * deltaRhoE;
      const T b_delta_alpha_2 = b * delta * alpha_2;
      flux.rho = flux_hlle.rho - b_delta_alpha_2;
      flux.mom[0] = flux_hlle.mom[0] - b_delta_alpha_2 * roeU[0];
      flux.E = flux_hlle.E - b_delta_alpha_2 * squaredNormRoeU_half;
    } else if constexpr (D == 2) {
      // :892-923
      const int ix = dir;
      const int iy = (ix == 0);
      const T l21 = gm1_over_roeA2 * (roeH - squaredNormRoeU);
      const T l22 = gm1_over_roeA2 * roeU[0];
      const T l23 = gm1_over_roeA2 * roeU[1];
      const T l24 = -gm1_over_roeA2;
      const T alpha_2 = l21 * deltaRho + l22 * deltaRhoU[0] +
                        l23 * deltaRhoU[1] + l24 * deltaRhoE;
      const T b_delta = b * delta;
      const T b_delta_alpha_2 = b_delta * alpha_2;
      flux.rho = flux_hlle.rho - b_delta_alpha_2;
      flux.mom[ix] = flux_hlle.mom[ix] - b_delta_alpha_2 * roeU[ix];
      if constexpr (Larrouturou) {
        const M8 upwind = (flux.rho > 0.0);
        const T f_hllem_rhov = flux.rho * select(upwind, vL_[iy], vR_[iy]);
        const T f_hllem_rhoE = flux.rho * select(upwind, eKinRL, eKinRR);
        flux.mom[iy] = f_hllem_rhov;
        flux.E = flux_hlle.E - b_delta_alpha_2 * squaredNormRoeU_half +
                 f_hllem_rhoE;
      } else {
        const T l31 = -roeU[iy];
        const T alpha_3 = l31 * deltaRho + deltaRhoU[iy];
        const T b_delta_alpha_3 = b_delta * alpha_3;
        flux.mom[iy] =
            flux_hlle.mom[iy] - b_delta_alpha_2 * roeU[iy] - b_delta_alpha_3;
        flux.E = flux_hlle.E - b_delta_alpha_2 * squaredNormRoeU_half -
                 b_delta_alpha_3 * roeU[iy];
      }
    } else {
      // :974-1019
      static_assert(D == 3);
      const int ix = dir;
      const int iy = (ix + 1) % 3;
      const int iz = (iy + 1) % 3;
      const T l21 = gm1_over_roeA2 * (roeH - squaredNormRoeU);
      const T l22 = gm1_over_roeA2 * roeU[0];
      const T l23 = gm1_over_roeA2 * roeU[1];
      const T l24 = gm1_over_roeA2 * roeU[2];
      const T l25 = -gm1_over_roeA2;
      const T alpha_2 = l21 * deltaRho + l22 * deltaRhoU[0] +
                        l23 * deltaRhoU[1] + l24 * deltaRhoU[2] +
                        l25 * deltaRhoE;
      const T b_delta = b * delta;
      const T b_delta_alpha_2 = b_delta * alpha_2;
      flux.rho = flux_hlle.rho - b_delta_alpha_2;
      // :991-992 -- roeU.row(iy), NOT row(ix): reference quirk Q1, reproduced.
      flux.mom[ix] = flux_hlle.mom[ix] - b_delta_alpha_2 * roeU[iy];
      if constexpr (Larrouturou) {
        const M8 upwind = (flux.rho > 0.0);
        const T f_hllem_rhov = flux.rho * select(upwind, vL_[iy], vR_[iy]);
        const T f_hllem_rhow = flux.rho * select(upwind, vL_[iz], vR_[iz]);
        const T f_hllem_rhoE = flux.rho * select(upwind, eKinRL, eKinRR);
        flux.mom[iy] = f_hllem_rhov;
        flux.mom[iz] = f_hllem_rhow;
        flux.E = flux_hlle.E - b_delta_alpha_2 * squaredNormRoeU_half +
                 f_hllem_rhoE;
      } else {
        const T l31 = -roeU[iy];
        const T l41 = -roeU[iz];
        const T alpha_3 = l31 * deltaRho + deltaRhoU[iy];
        const T alpha_4 = l41 * deltaRho + deltaRhoU[iz];
        const T b_delta_alpha_3 = b_delta * alpha_3;
        const T b_delta_alpha_4 = b_delta * alpha_4;
        flux.mom[iy] =
            flux_hlle.mom[iy] - b_delta_alpha_2 * roeU[iy] - b_delta_alpha_3;
        flux.mom[iz] =
            flux_hlle.mom[iz] - b_delta_alpha_2 * roeU[iz] - b_delta_alpha_4;
        flux.E = flux_hlle.E - b_delta_alpha_2 * squaredNormRoeU_half -
                 b_delta_alpha_3 * roeU[iy] - b_delta_alpha_4 * roeU[iz];
      }
    }
  }

  // :1074-1093 (masked: plain flux, zeroed where the face fraction is <= 0)
  void ComputeNumericFlux(Cons<T, D>& flux, const T& face_fractions,
                          const Complete<T, D>* states, const T* /*alphas*/,
                          double dt, double dx, int dir) const {
    ComputeNumericFlux(flux, states, dt, dx, dir);
    const M8 face_mask = face_fractions > 0.0;
    for_each_cons<T, D>([&](T& f) { f = select(face_mask, f, 0.0); }, flux);
  }

  // :1152-1216
  T ComputeStableDt(const Complete<T, D>* states, double dx, int dir) const {
    const Complete<T, D>& left = states[0];
    const Complete<T, D>& right = states[1];
    const double gm1 = eq.gamma - 1.0;
    const T rhoL = left.rho;
    const T rhoR = right.rho;
    const T aL = left.a;
    const T aR = right.a;
    const T hL = (left.E + left.p) / rhoL;
    const T hR = (right.E + right.p) / rhoR;
    const T sqRhoL = vsqrt(rhoL);
    const T sqRhoR = vsqrt(rhoR);
    const T sqRhoSum = sqRhoL + sqRhoR;
    const T sqRhoL_over_Sum = sqRhoL / sqRhoSum;
    const T sqRhoR_over_Sum = sqRhoR / sqRhoSum;
    T uL[D], uR[D], roeU[D];
    for (int i = 0; i < D; ++i) {
      uL[i] = left.mom[i] / rhoL;
      uR[i] = right.mom[i] / rhoR;
      roeU[i] = sqRhoL_over_Sum * uL[i] + sqRhoR_over_Sum * uR[i];
    }
    const T roeH = sqRhoL_over_Sum * hL + sqRhoR_over_Sum * hR;
    T squaredNormRoeU(0.0);
    for (int i = 0; i < D; ++i) {
      squaredNormRoeU += roeU[i] * roeU[i];
    }
    const T squaredNormRoeU_half = 0.5 * squaredNormRoeU;
    const T roeA2 = gm1 * (roeH - squaredNormRoeU_half);
    const T roeA = vsqrt(roeA2);
    const T maxA = vmax(aL, aR);
    const T sL1 = uL[dir] - maxA;
    const T sL2 = roeU[dir] - roeA;
    const T sR1 = roeU[dir] + roeA;
    const T sR2 = uR[dir] + maxA;
    const T sL = vmin(sL1, sL2);
    const T sR = vmax(sR1, sR2);
    const T maxS = vmax(vabs(sL), vabs(sR));
    return A8(dx) / maxS;
  }

  // :1218-1293 (masked)
  T ComputeStableDt(const Complete<T, D>* states, const T& face_fraction,
                    const T* /*alphas*/, double dx, int dir) const {
    const Complete<T, D>& left = states[0];
    const Complete<T, D>& right = states[1];
    const M8 face_mask = face_fraction > 0.0;
    const double gm1 = eq.gamma - 1.0;
    const T rhoL = select(face_mask, left.rho, 1.0);
    const T rhoR = select(face_mask, right.rho, 1.0);
    const T aL = select(face_mask, left.a, 0.0);
    const T aR = select(face_mask, right.a, 0.0);
    const T rhoEL = select(face_mask, left.E, 0.0);
    const T rhoER = select(face_mask, right.E, 0.0);
    const T pL = select(face_mask, left.p, 0.0);
    const T pR = select(face_mask, right.p, 0.0);
    const T hL = (rhoEL + pL) / rhoL;
    const T hR = (rhoER + pR) / rhoR;
    const T sqRhoL = vsqrt(rhoL);
    const T sqRhoR = vsqrt(rhoR);
    const T sqRhoSum = sqRhoL + sqRhoR;
    const T sqRhoL_over_Sum = sqRhoL / sqRhoSum;
    const T sqRhoR_over_Sum = sqRhoR / sqRhoSum;
    T uL[D], uR[D], roeU[D];
    for (int i = 0; i < D; ++i) {
      uL[i] = left.mom[i] / rhoL;
      uR[i] = right.mom[i] / rhoR;
      roeU[i] = sqRhoL_over_Sum * uL[i] + sqRhoR_over_Sum * uR[i];
    }
    const T roeH = sqRhoL_over_Sum * hL + sqRhoR_over_Sum * hR;
    T squaredNormRoeU(0.0);
    for (int i = 0; i < D; ++i) {
      squaredNormRoeU += roeU[i] * roeU[i];
    }
    const T squaredNormRoeU_half = 0.5 * squaredNormRoeU;
    const T roeA2 = gm1 * (roeH - squaredNormRoeU_half);
    const T roeA = vsqrt(roeA2);
    const T maxA = vmax(aL, aR);
    const T sL1 = uL[dir] - maxA;
    const T sL2 = roeU[dir] - roeA;
    const T sR1 = roeU[dir] + roeA;
    const T sR2 = uR[dir] + maxA;
    const T sL = vmin(sL1, sL2);
    const T sR = vmax(sR1, sR2);
    const T maxS = vmax(vabs(sL), vabs(sR));
    const T maxS_positive = select(face_mask, maxS, 1.0);
    const T max_dt = A8(dx) / maxS_positive;
    return select(face_mask, max_dt, std::numeric_limits<double>::max());
  }
};

// ---------------------------------------------------------------------------
// Exact Riemann solver (scalar; the reference's array versions are per-lane
// loops over the scalar code, ExactRiemannSolver.hpp:172-202).
// include/fub/NewtonIteration.hpp:28-52
template <typename F, typename DF>
inline double NewtonIteration(F&& f, DF&& Df, double x0,
                              double tolerance = 1e-7,
                              int max_iterations = 5000) {
  double x = x0;
  double fx = f(x);
  int counter = 0;
  double err = std::abs(fx);
  while (err > tolerance && counter < max_iterations) {
    double dfx = Df(x);
    double offset = fx / dfx;
    double next = x - offset;
    x = ((next - next) != (next - next) || next < 0.0) ? 0.5 * x : next;
    fx = f(x);
    err = std::abs(fx);
    counter += 1;
  }
  if (counter == max_iterations && fx > tolerance) {
    throw std::runtime_error("Newton Iteration did not converge.");
  }
  return x;
}

template <int D> struct ExactRiemannSolver {
  PerfectGasEq eq;
  explicit ExactRiemannSolver(const PerfectGasEq& e) : eq(e) {}
  using State = Complete<double, D>;

  // src/equations/perfect_gas/ExactRiemannSolver.hpp:34-97
  std::array<double, 2> ComputeMiddleState(const State& left,
                                           const State& right, int dir) const {
    const double g = eq.gamma;
    const double gp1 = g + 1.0;
    const double gm1 = g - 1.0;
    const double rhoL = left.rho;
    const double pL = left.p;
    const double aL = left.a;
    const double A_l = 2.0 / (gp1 * rhoL);
    const double B_l = gm1 / gp1 * pL;
    const double exponent = 0.5 * gm1 / g;
    const double d_exponent = -0.5 * gp1 / g;
    auto f_left = [=](double p) {
      if (p > pL) {
        return (p - pL) * std::sqrt(A_l / (p + B_l));
      } else {
        return 2.0 * aL / gm1 * (std::pow(p / pL, exponent) - 1.0);
      }
    };
    auto df_left = [=](double p) {
      if (p > pL) {
        return std::sqrt(A_l / (B_l + p)) * (1.0 - 0.5 * (p - pL) / (B_l + p));
      } else {
        return std::pow(p / pL, d_exponent) / (rhoL * aL);
      }
    };
    const double rhoR = right.rho;
    const double pR = right.p;
    const double aR = right.a;
    const double A_r = 2.0 / (gp1 * rhoR);
    const double B_r = gm1 / gp1 * pR;
    auto f_right = [=](double p) {
      if (p > pR) {
        return (p - pR) * std::sqrt(A_r / (p + B_r));
      } else {
        return 2 * aR / gm1 * (std::pow(p / pR, exponent) - 1.0);
      }
    };
    auto df_right = [=](double p) {
      if (p > pR) {
        return std::sqrt(A_r / (B_r + p)) * (1.0 - 0.5 * (p - pR) / (B_r + p));
      } else {
        return std::pow(p / pR, d_exponent) / (rhoR * aR);
      }
    };
    const double uL = left.mom[dir] / left.rho;
    const double uR = right.mom[dir] / right.rho;
    const double du = uR - uL;
    auto f = [=](double h) { return f_right(h) + f_left(h) + du; };
    auto df = [=](double h) { return df_right(h) + df_left(h); };
    const double pM = NewtonIteration(f, df, 0.5 * (pL + pR));
    const double fL = f_left(pM);
    const double fR = f_right(pM);
    const double uM = 0.5 * (uL + uR) + 0.5 * (fR - fL);
    return {pM, uM};
  }

  // :105-148
  std::array<double, 2> ComputeSignals(const State& left, const State& right,
                                       int dir) const {
    std::array<double, 2> states = ComputeMiddleState(left, right, dir);
    const double pM = states[0];
    const double g = eq.gamma;
    const double gp1 = g + 1.0;
    const double gm1 = g - 1.0;
    std::array<double, 2> signals{};
    const double uL = left.mom[dir] / left.rho;
    const double aL = left.a;
    const double pL = left.p;
    const double pL_ratio = pM / pL;
    if (pM > pL) {
      signals[0] = uL - aL * std::sqrt((gp1 * pL_ratio + gm1) / (2.0 * g));
    } else {
      signals[0] = uL - aL;
    }
    const double uR = right.mom[dir] / right.rho;
    const double aR = right.a;
    const double pR = right.p;
    const double pR_ratio = pM / pR;
    if (pM > pR) {
      signals[1] = uR + aR * std::sqrt((gp1 * pR_ratio + gm1) / (2.0 * g));
    } else {
      signals[1] = uR + aR;
    }
    return signals;
  }

  // :204-353.  `state` keeps its previous content when a density is <= 0.
  void SolveRiemannProblem(State& state, const State& left, const State& right,
                           int dir) const {
    auto set_momentum = [dir](State& s, double momentum, const State& side) {
      if constexpr (D == 1) {
        (void)dir;
        s.mom[0] = momentum;
      } else {
        for (int d = 0; d < D; ++d) {
          s.mom[d] = s.rho * (side.mom[d] / side.rho);
        }
        s.mom[dir] = momentum;
      }
    };
    auto from_prim = [this](State& s) {
      // ExactRiemannSolver.hpp:28-31: scalar squaredNorm of a fixed-size
      // vector -> same unrolled reduction order as squared_norm().
      const double rhoE_kin = 0.5 * squared_norm<double, D>(s.mom) / s.rho;
      const double rhoE_internal = s.p * eq.gamma_minus_1_inv;
      s.E = rhoE_internal + rhoE_kin;
      s.a = std::sqrt(eq.gamma * s.p / s.rho);
    };
    if (0.0 < left.rho && 0.0 < right.rho) {
      std::array<double, 2> states = ComputeMiddleState(left, right, dir);
      const double pM = states[0];
      const double uM = states[1];
      const double g = eq.gamma;
      const double gp1 = g + 1.0;
      const double gm1 = g - 1.0;
      if (0.0 <= uM) {
        const double uL = left.mom[dir] / left.rho;
        const double aL = left.a;
        const double pL = left.p;
        const double p_ratio = pM / pL;
        const double g_ratio = gm1 / gp1;
        if (pM > pL) {
          const double shock_speed =
              uL - aL * std::sqrt((gp1 * p_ratio + gm1) / (2.0 * g));
          if (shock_speed < 0.0) {
            state.rho = left.rho * (p_ratio + g_ratio) / (g_ratio * p_ratio + 1.0);
            set_momentum(state, state.rho * uM, left);
            state.p = pM;
            from_prim(state);
          } else {
            state = left;
          }
        } else {
          const double aM = aL * std::pow(p_ratio, gm1 / (2.0 * g));
          const double shock_speed_TL = uM - aM;
          const double shock_speed_HL = uL - aL;
          if (0.0 < shock_speed_HL) {
            state = left;
          } else if (shock_speed_TL < 0.0) {
            state.rho = left.rho * std::pow(p_ratio, 1 / g);
            set_momentum(state, state.rho * uM, left);
            state.p = pM;
            from_prim(state);
          } else {
            state.rho =
                left.rho * std::pow(2.0 / gp1 + g_ratio * uL / aL, 2.0 / gm1);
            const double u = 2.0 / gp1 * (aL + 0.5 * gm1 * uL);
            set_momentum(state, state.rho * u, left);
            state.p = pL * std::pow(2.0 / gp1 + g_ratio * uL / aL, 2.0 * g / gm1);
            from_prim(state);
          }
        }
      } else {
        const double uR = right.mom[dir] / right.rho;
        const double aR = right.a;
        const double pR = right.p;
        const double p_ratio = pM / pR;
        const double g_ratio = gm1 / gp1;
        if (pM > pR) {
          const double shock_speed =
              uR + aR * std::sqrt((gp1 * p_ratio + gm1) / (2.0 * g));
          if (0.0 < shock_speed) {
            state.rho =
                right.rho * (p_ratio + g_ratio) / (g_ratio * p_ratio + 1.0);
            set_momentum(state, state.rho * uM, right);
            state.p = pM;
            from_prim(state);
          } else {
            state = right;
          }
        } else {
          const double aM = aR * std::pow(p_ratio, gm1 / (2.0 * g));
          const double shock_speed_TR = uM + aM;
          const double shock_speed_HR = uR + aR;
          if (shock_speed_HR <= 0.0) {
            state = right;
          } else if (shock_speed_TR >= 0.0) {
            state.rho = right.rho * std::pow(p_ratio, 1 / g);
            set_momentum(state, state.rho * uM, right);
            state.p = pM;
            from_prim(state);
          } else {
            state.rho =
                right.rho * std::pow(2.0 / gp1 - g_ratio * uR / aR, 2.0 / gm1);
            const double u = 2.0 / gp1 * (-aR + 0.5 * gm1 * uR);
            set_momentum(state, state.rho * u, right);
            state.p = pR * std::pow(2.0 / gp1 - g_ratio * uR / aR, 2.0 * g / gm1);
            from_prim(state);
          }
        }
      }
    }
  }
};

template <int D> inline Complete<double, D> get_lane(const Complete<A8, D>& a, int l) {
  Complete<double, D> s;
  s.rho = a.rho.v[l];
  for (int d = 0; d < D; ++d) s.mom[d] = a.mom[d].v[l];
  s.E = a.E.v[l];
  s.p = a.p.v[l];
  s.a = a.a.v[l];
  return s;
}
template <int D> inline void set_lane(Complete<A8, D>& a, const Complete<double, D>& s, int l) {
  a.rho.v[l] = s.rho;
  for (int d = 0; d < D; ++d) a.mom[d].v[l] = s.mom[d];
  a.E.v[l] = s.E;
  a.p.v[l] = s.p;
  a.a.v[l] = s.a;
}
template <int D> inline Complete<double, D> zero_state() {
  Complete<double, D> s;
  s.rho = 0;
  for (int d = 0; d < D; ++d) s.mom[d] = 0;
  s.E = 0;
  s.p = 0;
  s.a = 0;
  return s;
}

// Godunov method over the exact solver.
// include/fub/flux_method/GodunovMethod.hpp:129-199
//
// NOTE: in the reference the array ComputeSignals of the exact solver throws
// (ExactRiemannSolver.hpp:99-103), so Godunov<Exact>::ComputeStableDt has no
// working regular-grid array path; the oracle defines it through the scalar
// ComputeSignals per lane (what the masked path :185-199 and the seq tag do).
template <int D> struct GodunovExact {
  static constexpr int kWidth = 1;
  static constexpr int kDim = D;
  using T = A8;
  using Lane = A8;
  PerfectGasEq eq;
  ExactRiemannSolver<D> rs;
  explicit GodunovExact(const PerfectGasEq& e) : eq(e), rs(e) {}

  void ComputeNumericFlux(Cons<T, D>& nf, const Complete<T, D>* states,
                          double, double, int dir) const {
    Complete<T, D> sol;
    for (int l = 0; l < kChunk; ++l) {
      Complete<double, D> s = zero_state<D>();
      rs.SolveRiemannProblem(s, get_lane<D>(states[0], l),
                             get_lane<D>(states[1], l), dir);
      set_lane<D>(sol, s, l);
    }
    Flux<T, D>(eq, nf, sol, dir);
  }

  void ComputeNumericFlux(Cons<T, D>& nf, const T& face_fraction,
                          const Complete<T, D>* states, const T*, double,
                          double, int dir) const {
    const M8 mask = face_fraction > 0.0;
    Complete<T, D> sol;
    for (int l = 0; l < kChunk; ++l) {
      Complete<double, D> s = zero_state<D>();
      if (mask.v[l]) {
        rs.SolveRiemannProblem(s, get_lane<D>(states[0], l),
                               get_lane<D>(states[1], l), dir);
      }
      set_lane<D>(sol, s, l);
    }
    Flux<T, D>(eq, nf, sol, dir);
    for_each_cons<T, D>([&](T& f) { f = select(mask, f, 0.0); }, nf);
  }

  double StableDtLane(const Complete<double, D>& l, const Complete<double, D>& r,
                      double dx, int dir) const {
    auto signals = rs.ComputeSignals(l, r, dir);
    double s_max = 0.0;
    s_max = std::max(s_max, std::abs(signals[0]));
    s_max = std::max(s_max, std::abs(signals[1]));
    return dx / s_max;
  }

  T ComputeStableDt(const Complete<T, D>* states, double dx, int dir) const {
    T dts;
    for (int l = 0; l < kChunk; ++l) {
      dts.v[l] = StableDtLane(get_lane<D>(states[0], l),
                              get_lane<D>(states[1], l), dx, dir);
    }
    return dts;
  }

  T ComputeStableDt(const Complete<T, D>* states, const T& face_fraction,
                    const T*, double dx, int dir) const {
    T dts(std::numeric_limits<double>::infinity());
    for (int l = 0; l < kChunk; ++l) {
      if (face_fraction.v[l] > 0.0) {
        dts.v[l] = StableDtLane(get_lane<D>(states[0], l),
                                get_lane<D>(states[1], l), dx, dir);
      }
    }
    return dts;
  }
};

// ---------------------------------------------------------------------------
// MUSCL-Hancock.  include/fub/flux_method/MusclHancockMethod.hpp:189-571
template <typename Base, typename Limiter> struct MusclHancock {
  static constexpr int kWidth = 2;
  static constexpr int kDim = Base::kDim;
  static constexpr int D = Base::kDim;
  using T = typename Base::Lane;
  using Lane = T;
  PerfectGasEq eq;
  Base base;
  Limiter limiter;
  explicit MusclHancock(const PerfectGasEq& e) : eq(e), base(e) {}

  // :208-218
  T ComputeStableDt(const Complete<T, D>* states, double dx, int dir) const {
    return base.ComputeStableDt(states + 1, dx, dir);
  }
  // :220-227
  T ComputeStableDt(const Complete<T, D>* states, const T& face_fraction,
                    const T* alphas, double dx, int dir) const {
    return base.ComputeStableDt(states + 1, face_fraction, alphas + 1, dx, dir);
  }

  // one half of :428-451 / :456-479
  template <bool kRight>
  void Reconstruct(Complete<T, D>& rec, const Complete<T, D>* st,
                   const T& lambda_half, int dir) const {
    Cons<T, D> slope;
    slope.rho = limiter(st[0].rho, st[1].rho, st[2].rho);
    for (int d = 0; d < D; ++d) {
      slope.mom[d] = limiter(st[0].mom[d], st[1].mom[d], st[2].mom[d]);
    }
    slope.E = limiter(st[0].E, st[1].E, st[2].E);
    Complete<T, D> qL, qR;
    qL.rho = st[1].rho - 0.5 * slope.rho;
    qR.rho = st[1].rho + 0.5 * slope.rho;
    for (int d = 0; d < D; ++d) {
      qL.mom[d] = st[1].mom[d] - 0.5 * slope.mom[d];
      qR.mom[d] = st[1].mom[d] + 0.5 * slope.mom[d];
    }
    qL.E = st[1].E - 0.5 * slope.E;
    qR.E = st[1].E + 0.5 * slope.E;
    CompleteFromCons<T, D>(eq, qL, qL);
    CompleteFromCons<T, D>(eq, qR, qR);
    Cons<T, D> fL, fR;
    Flux<T, D>(eq, fL, qL, dir);
    Flux<T, D>(eq, fR, qR, dir);
    const Cons<T, D>& q = kRight ? static_cast<const Cons<T, D>&>(qR)
                                 : static_cast<const Cons<T, D>&>(qL);
    rec.rho = q.rho + lambda_half * (fL.rho - fR.rho);
    for (int d = 0; d < D; ++d) {
      rec.mom[d] = q.mom[d] + lambda_half * (fL.mom[d] - fR.mom[d]);
    }
    rec.E = q.E + lambda_half * (fL.E - fR.E);
    CompleteFromCons<T, D>(eq, rec, rec);
  }

  // :419-485 (array) == :294-356 (scalar)
  void ComputeNumericFlux(Cons<T, D>& flux, const Complete<T, D>* stencil,
                          double dt, double dx, int dir) const {
    const T lambda_half = splat(stencil[0].rho, 0.5 * dt / dx);
    Complete<T, D> rec[2];
    Reconstruct<true>(rec[0], stencil, lambda_half, dir);
    Reconstruct<false>(rec[1], stencil + 1, lambda_half, dir);
    base.ComputeNumericFlux(flux, rec, dt, dx, dir);
  }

  template <bool kRight>
  void ReconstructMasked(Complete<T, D>& rec, const Complete<T, D>* st,
                         const mask_t<T>& slope_mask,
                         const mask_t<T>& valid_face, const T& lambda_half,
                         int dir) const {
    Cons<T, D> slope;
    slope.rho = limiter(st[0].rho, st[1].rho, st[2].rho, slope_mask);
    for (int d = 0; d < D; ++d) {
      slope.mom[d] =
          limiter(st[0].mom[d], st[1].mom[d], st[2].mom[d], slope_mask);
    }
    slope.E = limiter(st[0].E, st[1].E, st[2].E, slope_mask);
    Complete<T, D> qL, qR;
    auto split = [&](T& l, T& r, const T& state, const T& s) {
      T q0 = select(valid_face, state, 0.0);
      l = q0 - 0.5 * s;
      r = q0 + 0.5 * s;
    };
    split(qL.rho, qR.rho, st[1].rho, slope.rho);
    for (int d = 0; d < D; ++d) {
      split(qL.mom[d], qR.mom[d], st[1].mom[d], slope.mom[d]);
    }
    split(qL.E, qR.E, st[1].E, slope.E);
    CompleteFromCons<T, D>(eq, qL, qL, valid_face);
    CompleteFromCons<T, D>(eq, qR, qR, valid_face);
    Cons<T, D> fL, fR;
    Flux<T, D>(eq, fL, qL, valid_face, dir);
    Flux<T, D>(eq, fR, qR, valid_face, dir);
    const Cons<T, D>& q = kRight ? static_cast<const Cons<T, D>&>(qR)
                                 : static_cast<const Cons<T, D>&>(qL);
    rec.rho = q.rho + lambda_half * (fL.rho - fR.rho);
    for (int d = 0; d < D; ++d) {
      rec.mom[d] = q.mom[d] + lambda_half * (fL.mom[d] - fR.mom[d]);
    }
    rec.E = q.E + lambda_half * (fL.E - fR.E);
    CompleteFromCons<T, D>(eq, rec, rec, valid_face);
  }

  // :487-571 (masked array)
  void ComputeNumericFlux(Cons<T, D>& flux, const T& face_fractions,
                          const Complete<T, D>* stencil, const T* alphas,
                          double dt, double dx, int dir) const {
    const mask_t<T> valid_face = face_fractions > 0.0;
    if (any(valid_face)) {
      const T lambda_half = splat(stencil[0].rho, 0.5 * dt / dx);
      Complete<T, D> rec[2];
      const mask_t<T> left_slope_mask =
          (alphas[0] > 0.0) && (alphas[1] > 0.0) && (alphas[2] > 0.0);
      ReconstructMasked<true>(rec[0], stencil, left_slope_mask, valid_face,
                              lambda_half, dir);
      const mask_t<T> right_slope_mask =
          (alphas[1] > 0.0) && (alphas[2] > 0.0) && (alphas[3] > 0.0);
      ReconstructMasked<false>(rec[1], stencil + 1, right_slope_mask,
                               valid_face, lambda_half, dir);
      base.ComputeNumericFlux(flux, face_fractions, rec, alphas + 1, dt, dx,
                              dir);
    } else {
      for_each_cons<T, D>([&](T& f) { f = splat(f, 0.0); }, flux);
    }
  }
};

// ---------------------------------------------------------------------------
// perfect_gas::MusclHancockPrim: MUSCL-Hancock on primitive variables with its
// own limiter over HLLEM (Larrouturou).
// src/equations/perfect_gas/MusclHancockPrimMethod.hpp:25-333
template <int D> struct MusclHancockPrim {
  static constexpr int kWidth = 2;
  static constexpr int kDim = D;
  using T = A8;
  using Lane = A8;
  PerfectGasEq eq;
  Hllem<D, true> hllem;
  explicit MusclHancockPrim(const PerfectGasEq& e) : eq(e), hllem(e) {}

  // :70-86
  static T limiter(const T& qL, const T& qM, const T& qR) {
    const T delta_q = 0.5 * (qR - qL);
    const M8 is_relevant = vabs(delta_q) > 1e-12;
    const T delta_q_if_relevant = select(is_relevant, delta_q, 1.0);
    const T rL = select(is_relevant, (qM - qL) / delta_q_if_relevant, 0.0);
    const T rR = select(is_relevant, (qR - qM) / delta_q_if_relevant, 0.0);
    const M8 is_positive = (rL > 0.0) && (rR > 0.0);
    const T r = select(is_positive, vmin(rL, rR), 0.0);
    const M8 less_than_one = r < 1.0;
    const T r_1 = 2 * r / (A8(1.0) + r);
    const T r_2 = 0.5 * (r + A8(1.0));
    const T sigma = select(less_than_one, r_1, r_2);
    return sigma * delta_q;
  }

  // :33-38
  T ComputeStableDt(const Complete<T, D>* states, double dx, int dir) const {
    return hllem.ComputeStableDt(states + 1, dx, dir);
  }
  // :40-47
  T ComputeStableDt(const Complete<T, D>* states, const T& face_fraction, const T* vols,
                    double dx, int dir) const {
    return hllem.ComputeStableDt(states + 1, face_fraction, vols + 1, dx, dir);
  }

  // the part of :59-179 / :185-329 after the (masked) inputs have been chosen
  void Reconstruct(Complete<T, D> (&w)[2], const T& rho_ll, const T& rho_l, const T& rho_r,
                   const T& rho_rr, const T& p_ll, const T& p_l, const T& p_r, const T& p_rr,
                   const T& a_l, const T& a_r, const T (&u_ll)[D], const T (&u_l)[D],
                   const T (&u_r)[D], const T (&u_rr)[D], double dt, double dx, int d) const {
    const T lambda(dt / dx);
    const T drho_l = limiter(rho_ll, rho_l, rho_r);
    const T dp_l = limiter(p_ll, p_l, p_r);
    T du_l[D];
    for (int i = 0; i < D; ++i) du_l[i] = limiter(u_ll[i], u_l[i], u_r[i]);
    const T drho_r = limiter(rho_l, rho_r, rho_rr);
    const T dp_r = limiter(p_l, p_r, p_rr);
    T du_r[D];
    for (int i = 0; i < D; ++i) du_r[i] = limiter(u_l[i], u_r[i], u_rr[i]);

    const T rec_rho_l = rho_l + 0.5 * (drho_l - lambda * (u_l[d] * drho_l + rho_l * du_l[d]));
    const T rec_p_l = p_l + 0.5 * (dp_l - lambda * (rho_l * a_l * a_l * du_l[d] + u_l[d] * dp_l));
    T rec_u_l[D];
    for (int i = 0; i < D; ++i) {
      rec_u_l[i] = u_l[i] + 0.5 * (du_l[i] - lambda * (u_l[d] * du_l[i]));
    }
    rec_u_l[d] = rec_u_l[d] - 0.5 * lambda * dp_l / rho_l;

    const T rec_rho_r = rho_r - 0.5 * (drho_r + lambda * (u_r[d] * drho_r + rho_r * du_r[d]));
    const T rec_p_r = p_r - 0.5 * (dp_r + lambda * (rho_r * a_r * a_r * du_r[d] + u_r[d] * dp_r));
    T rec_u_r[D];
    for (int i = 0; i < D; ++i) {
      rec_u_r[i] = u_r[i] - 0.5 * (du_r[i] + lambda * (u_r[d] * du_r[i]));
    }
    rec_u_r[d] = rec_u_r[d] - 0.5 * lambda * dp_r / rho_r;

    auto fill = [&](Complete<T, D>& q, const T& rho, const T (&u)[D], const T& p) {
      q.rho = rho;
      for (int i = 0; i < D; ++i) q.mom[i] = rho * u[i];
      q.p = p;
      q.a = vsqrt(eq.gamma * p / rho);
      q.E = 0.5 * rho * squared_norm<T, D>(u) + p * eq.gamma_minus_1_inv;
    };
    fill(w[0], rec_rho_l, rec_u_l, rec_p_l);
    fill(w[1], rec_rho_r, rec_u_r, rec_p_r);
  }

  // :59-179
  void ComputeNumericFlux(Cons<T, D>& flux, const Complete<T, D>* st, double dt, double dx,
                          int dir) const {
    T u[4][D];
    for (int s = 0; s < 4; ++s)
      for (int i = 0; i < D; ++i) u[s][i] = st[s].mom[i] / st[s].rho;
    Complete<T, D> w[2];
    Reconstruct(w, st[0].rho, st[1].rho, st[2].rho, st[3].rho, st[0].p, st[1].p, st[2].p, st[3].p,
                st[1].a, st[2].a, u[0], u[1], u[2], u[3], dt, dx, dir);
    hllem.ComputeNumericFlux(flux, w, dt, dx, dir);
  }

  // :185-329 (masked)
  void ComputeNumericFlux(Cons<T, D>& flux, const T& face_fractions, const Complete<T, D>* st,
                          const T* volume_fractions, double dt, double dx, int dir) const {
    const M8 mask = face_fractions > 0.0;
    if (!any(mask)) {
      // The reference returns here and leaves `flux` -- the method's zero-initialised
      // member array, KbnStabilisation.hpp:399-407 -- holding the fluxes of whichever chunk
      // this thread computed last; they are stored into faces whose face fraction is 0 and
      // every later use multiplies them by a zero fraction.  Which values end up there
      // depends on the OpenMP schedule, so there is nothing to reproduce: a face of a chunk
      // without any open face gets 0, like the closed faces of every other chunk (:1082-1083
      // of HllemMethod.hpp).  DESIGN.md section 7, "Deliberate restrictions".
      flux.rho = T(0.0);
      for (int i = 0; i < D; ++i) flux.mom[i] = T(0.0);
      flux.E = T(0.0);
      return;
    }
    const M8 left_mask = volume_fractions[0] > 0.0;
    const M8 right_mask = volume_fractions[3] > 0.0;
    const T rho_l = select(mask, st[1].rho, 1.0);
    const T rho_r = select(mask, st[2].rho, 1.0);
    const T rho_ll = select(left_mask, st[0].rho, rho_l);
    const T rho_rr = select(right_mask, st[3].rho, rho_r);
    const T p_l = select(mask, st[1].p, 1.0);
    const T p_r = select(mask, st[2].p, 1.0);
    const T p_ll = select(left_mask, st[0].p, p_l);
    const T p_rr = select(right_mask, st[3].p, p_r);
    const T a_l = select(mask, st[1].a, 1.0);
    const T a_r = select(mask, st[2].a, 1.0);
    T u_ll[D], u_l[D], u_r[D], u_rr[D];
    for (int i = 0; i < D; ++i) {
      const T rho_u_l = select(mask, st[1].mom[i], 0.0);
      const T rho_u_r = select(mask, st[2].mom[i], 0.0);
      const T rho_u_ll = select(left_mask, st[0].mom[i], rho_u_l);
      const T rho_u_rr = select(right_mask, st[3].mom[i], rho_u_r);
      u_ll[i] = rho_u_ll / rho_ll;
      u_l[i] = rho_u_l / rho_l;
      u_r[i] = rho_u_r / rho_r;
      u_rr[i] = rho_u_rr / rho_rr;
    }
    Complete<T, D> w[2];
    Reconstruct(w, rho_ll, rho_l, rho_r, rho_rr, p_ll, p_l, p_r, p_rr, a_l, a_r, u_ll, u_l, u_r,
                u_rr, dt, dx, dir);
    hllem.ComputeNumericFlux(flux, face_fractions, w, volume_fractions + 1, dt, dx, dir);
  }
};

// ---------------------------------------------------------------------------
// perfect_gas::MusclHancockCharacteristic (MusclHancockCharMethod): MUSCL-Hancock
// on characteristic variables, VanLeer limiter on the wave amplitudes, base
// method HLLEM (Larrouturou).
// src/equations/perfect_gas/MusclHancockCharacteristicMethod.hpp:28-359,
// include/fub/equations/perfect_gas/MusclHancockCharactersticMethod.hpp:34-175
template <int D> struct MusclHancockChar {
  static constexpr int kWidth = 2;
  static constexpr int kDim = D;
  using T = A8;
  using Lane = A8;
  PerfectGasEq eq;
  Hllem<D, true> hllem;
  explicit MusclHancockChar(const PerfectGasEq& e) : eq(e), hllem(e) {}

  struct Prim {
    T rho, p;
    T u[3];
  };
  struct Chars {
    T minus, zero, plus, v, w;
  };
  static Prim prim_of(const Complete<T, D>& q) { // PrimitivesArray, header :88-113
    Prim w;
    w.rho = q.rho;
    w.p = q.p;
    for (int d = 0; d < 3; ++d) w.u[d] = A8(0.0);
    for (int d = 0; d < D; ++d) w.u[d] = q.mom[d] / q.rho;
    return w;
  }
  // fub::VanLeer::operator()(Array1d, Array1d), MusclHancockMethod.hpp:144-149
  static T limiter(const T& sL, const T& sR) {
    const T r = 2.0 * sL * sR;
    return select(r > 0.0, r / (sL + sR), 0.0);
  }
  static Chars amplitudes(const Prim& l, const Prim& r, const T& rhoc, const T& ooc2, int ix,
                          int iy, int iz) { // :50-66
    const T dp = r.p - l.p;
    const T du = r.u[ix] - l.u[ix];
    const T drho = r.rho - l.rho;
    Chars a;
    a.minus = dp - rhoc * du;
    a.zero = drho - ooc2 * dp;
    a.plus = dp + rhoc * du;
    a.v = r.u[iy] - l.u[iy];
    a.w = r.u[iz] - l.u[iz];
    return a;
  }
  // Reconstruct, :228-253 (masked :255-285): `slope_mask` zeroes the left
  // amplitudes, `valid_face` masks CompleteFromPrim; both null = unmasked
  void Reconstruct(Complete<T, D>& rec, int is_right_side, const Complete<T, D>* st,
                   const T& lambda, int dir, const M8* slope_mask, const M8* valid_face) const {
    const int ix = dir, iy = (ix + 1) % 3, iz = (iy + 1) % 3;
    const Complete<T, D>& q = st[1];
    const T rho = q.rho;
    const T u = q.mom[dir] / rho;
    const T c = q.a;
    const T rhoc = rho * c;
    const T ooc2 = 1.0 / (c * c);
    const Prim pL = prim_of(st[0]), pM = prim_of(st[1]), pR = prim_of(st[2]);
    Chars al = amplitudes(pL, pM, rhoc, ooc2, ix, iy, iz);
    const Chars ar = amplitudes(pM, pR, rhoc, ooc2, ix, iy, iz);
    if (slope_mask) {
      al.minus = select(*slope_mask, al.minus, 0.0);
      al.zero = select(*slope_mask, al.zero, 0.0);
      al.plus = select(*slope_mask, al.plus, 0.0);
      al.v = select(*slope_mask, al.v, 0.0);
      al.w = select(*slope_mask, al.w, 0.0);
    }
    Chars slope;
    slope.minus = limiter(al.minus, ar.minus);
    slope.zero = limiter(al.zero, ar.zero);
    slope.plus = limiter(al.plus, ar.plus);
    slope.v = limiter(al.v, ar.v);
    slope.w = limiter(al.w, ar.w);
    // DoTimeStep, :178-190
    const int sign = 1 - 2 * is_right_side;
    Chars am;
    am.minus = sign * 0.5 * slope.minus * (1.0 - sign * lambda * (u - c));
    am.zero = sign * 0.5 * slope.zero * (1.0 - sign * lambda * u);
    am.plus = sign * 0.5 * slope.plus * (1.0 - sign * lambda * (u + c));
    am.v = sign * 0.5 * slope.v * (1.0 - sign * lambda * u);
    am.w = sign * 0.5 * slope.w * (1.0 - sign * lambda * u);
    // Recombine, :79-88
    Prim diffs;
    diffs.p = 0.5 * (am.minus + am.plus);
    diffs.rho = ooc2 * diffs.p + am.zero;
    for (int d = 0; d < 3; ++d) diffs.u[d] = A8(0.0);
    diffs.u[ix] = 0.5 * (am.plus - am.minus) / rhoc;
    diffs.u[iy] = am.v;
    diffs.u[iz] = am.w;
    diffs.rho = diffs.rho + rho;
    for (int d = 0; d < D; ++d) diffs.u[d] = diffs.u[d] + (q.mom[d] / q.rho);
    diffs.p = diffs.p + q.p;
    // PerfectGas::CompleteFromPrim (array), src/equations/PerfectGas.cpp:160-191
    if (!valid_face) {
      rec.rho = diffs.rho;
      for (int d = 0; d < D; ++d) rec.mom[d] = diffs.rho * diffs.u[d];
      rec.p = diffs.p;
      const T e_kin = KineticEnergy<T, D>(rec.rho, rec.mom);
      rec.E = e_kin + diffs.p * eq.gamma_minus_1_inv;
      rec.a = vsqrt(eq.gamma * rec.p / rec.rho);
    } else {
      const M8& mask = *valid_face;
      rec.rho = select(mask, diffs.rho, 0.0);
      for (int d = 0; d < D; ++d) rec.mom[d] = select(mask, diffs.rho * diffs.u[d], 0.0);
      rec.p = select(mask, diffs.p, 0.0);
      const T e_kin = select(mask, KineticEnergy<T, D>(diffs.rho, rec.mom), 0.0);
      rec.E = e_kin + rec.p * eq.gamma_minus_1_inv;
      rec.a = vsqrt(select(mask, eq.gamma * rec.p / rec.rho, 0.0));
    }
  }

  T ComputeStableDt(const Complete<T, D>* states, double dx, int dir) const {
    return hllem.ComputeStableDt(states + 1, dx, dir);
  }
  T ComputeStableDt(const Complete<T, D>* states, const T& face_fraction, const T* vols,
                    double dx, int dir) const {
    return hllem.ComputeStableDt(states + 1, face_fraction, vols + 1, dx, dir);
  }
  // :301-314
  void ComputeNumericFlux(Cons<T, D>& flux, const Complete<T, D>* st, double dt, double dx,
                          int dir) const {
    const T lambda(dt / dx);
    Complete<T, D> rec[2];
    Reconstruct(rec[0], 0, st, lambda, dir, nullptr, nullptr);
    Reconstruct(rec[1], 1, st + 1, lambda, dir, nullptr, nullptr);
    hllem.ComputeNumericFlux(flux, rec, dt, dx, dir);
  }
  // :316-334
  void ComputeNumericFlux(Cons<T, D>& flux, const T& face_fractions, const Complete<T, D>* st,
                          const T* vols, double dt, double dx, int dir) const {
    const T lambda(dt / dx);
    const M8 valid_face = face_fractions > 0.0;
    const M8 left_mask = (vols[0] > 0.0) && (vols[1] > 0.0) && (vols[2] > 0.0);
    const M8 right_mask = (vols[1] > 0.0) && (vols[2] > 0.0) && (vols[3] > 0.0);
    Complete<T, D> rec[2];
    Reconstruct(rec[0], 0, st, lambda, dir, &left_mask, &valid_face);
    Reconstruct(rec[1], 1, st + 1, lambda, dir, &right_mask, &valid_face);
    hllem.ComputeNumericFlux(flux, face_fractions, rec, vols + 1, dt, dx, dir);
  }
};

// ---------------------------------------------------------------------------
// MusclHancock2<Equation, Gradient, Reconstruction, Base>
// include/fub/flux_method/MusclHancockMethod2.hpp:35-239 (array path), with
//   include/fub/flux_method/Gradient.hpp     (limiters :31-59, CentralDifferenceGradient
//       :211-217, Conservative/Primitive/CharacteristicsGradient :247-302, :322-368)
//   include/fub/flux_method/Reconstruct.hpp  (Conservative :188-214, Primitive :343-388,
//       Characteristics :458-516)
// KIND selects the variables of both the gradient and the reconstruction (the three
// pairings of include/fub/AMReX/FluxMethodFactory.hpp:161-176); the limiter of the
// CentralDifferenceGradient is a run-time value here.
enum Mh2Kind { kMh2Conservative = 0, kMh2Primitive = 1, kMh2Characteristics = 2 };
enum Limiter2Kind { kLimiter2None = 0, kLimiter2Upwind = 1, kLimiter2MinMod = 2, kLimiter2VanLeer = 3 };

template <typename Base, int KIND> struct MusclHancock2 {
  static constexpr int kWidth = 2;
  static constexpr int kDim = Base::kDim;
  static constexpr int D = Base::kDim;
  using T = A8;
  using Lane = A8;
  PerfectGasEq eq;
  Base base;
  int limiter;
  MusclHancock2(const PerfectGasEq& e, int lim) : eq(e), base(e), limiter(lim) {}

  // Gradient.hpp:31-59
  T limit(const T& R) const {
    switch (limiter) {
    case kLimiter2None: return T(1.0);
    case kLimiter2Upwind: return T(0.0);
    case kLimiter2MinMod: {
      const T two_over_one_plus_R = 2.0 / (1.0 + R);
      return vmin(two_over_one_plus_R, R * two_over_one_plus_R);
    }
    default: return 4.0 * R / ((1.0 + R) * (1.0 + R));
    }
  }
  // CentralDifferenceGradient::ComputeGradient(Array1d, Array1d), Gradient.hpp:211-217
  T central(const T& sL, const T& sR) const {
    const M8 mask = (sL * sR) > 0.0;
    const T R = sR / sL;
    return select(mask, 0.5 * limit(R) * (sL + sR), 0.0);
  }

  // gradient variables: Conservative -> (rho, mom, E); Primitive -> (rho, u, p);
  // Characteristics -> (minus, zero[0..D-1], plus) stored as g[0], g[1..D], g[D+1]
  struct Grad {
    T g[D + 2];
  };
  struct Prim {
    T rho, p;
    T u[D];
  };
  static Prim prim_of(const Complete<T, D>& q) { // PrimFromComplete, PerfectGas.hpp:339-347
    Prim w;
    w.rho = q.rho;
    w.p = q.p;
    for (int i = 0; i < D; ++i) w.u[i] = q.mom[i] / q.rho;
    return w;
  }

  void ComputeGradient(Grad& out, const Complete<T, D>* q, double dx, int dir) const {
    if constexpr (KIND == kMh2Conservative) { // Gradient.hpp:259-270
      auto one = [&](const T& uL, const T& uM, const T& uR) {
        return central((uM - uL) / dx, (uR - uM) / dx);
      };
      out.g[0] = one(q[0].rho, q[1].rho, q[2].rho);
      for (int i = 0; i < D; ++i) out.g[1 + i] = one(q[0].mom[i], q[1].mom[i], q[2].mom[i]);
      out.g[D + 1] = one(q[0].E, q[1].E, q[2].E);
    } else if constexpr (KIND == kMh2Primitive) { // :287-302
      const Prim wL = prim_of(q[0]), wM = prim_of(q[1]), wR = prim_of(q[2]);
      auto one = [&](const T& a, const T& b, const T& c) {
        return central((b - a) / dx, (c - b) / dx);
      };
      out.g[0] = one(wL.rho, wM.rho, wR.rho);
      for (int i = 0; i < D; ++i) out.g[1 + i] = one(wL.u[i], wM.u[i], wR.u[i]);
      out.g[D + 1] = one(wL.p, wM.p, wR.p);
    } else { // :354-368 with ComputeAmplitudes :322-352 (MaSq == 1 for PerfectGas)
      const Prim wL = prim_of(q[0]), wM = prim_of(q[1]), wR = prim_of(q[2]);
      const T rho = q[1].rho;
      const T c = q[1].a;
      const T ooc2 = 1.0 / (c * c);
      const T rhoc = rho * c;
      auto amplitudes = [&](Grad& a, const Prim& left, const Prim& right) {
        const T Ma(std::sqrt(1.0));
        const T dp = (right.p - left.p) / dx;
        const T du = (right.u[dir] - left.u[dir]) / dx;
        const T drho = (right.rho - left.rho) / dx;
        a.g[0] = dp - Ma * rhoc * du;     // minus
        a.g[D + 1] = dp + Ma * rhoc * du; // plus
        a.g[1] = drho - ooc2 * dp;        // zero[0]
        for (int i = 1; i < D; ++i) {
          const int iy = (dir + i) % D;
          a.g[1 + i] = (right.u[iy] - left.u[iy]) / dx; // zero[i]
        }
      };
      Grad aL, aR;
      amplitudes(aL, wL, wM);
      amplitudes(aR, wM, wR);
      for (int k = 0; k < D + 2; ++k) out.g[k] = central(aL.g[k], aR.g[k]);
    }
  }

  // side: -1 = Side::Lower, +1 = Side::Upper
  void Reconstruct(Complete<T, D>& rec, const Complete<T, D>& q0, const Grad& grad, double dt,
                   double dx, int dir, int sign) const {
    if constexpr (KIND == kMh2Conservative) { // Reconstruct.hpp:188-214
      Complete<T, D> qL, qR;
      auto split = [&](T& uL, T& uR, const T& u, const T& du_dx) {
        const T du = (0.5 * dx) * du_dx;
        uL = u - du;
        uR = u + du;
      };
      split(qL.rho, qR.rho, q0.rho, grad.g[0]);
      for (int i = 0; i < D; ++i) split(qL.mom[i], qR.mom[i], q0.mom[i], grad.g[1 + i]);
      split(qL.E, qR.E, q0.E, grad.g[D + 1]);
      CompleteFromCons<T, D>(eq, qL, qL);
      CompleteFromCons<T, D>(eq, qR, qR);
      Cons<T, D> fL, fR;
      Flux<T, D>(eq, fL, qL, dir);
      Flux<T, D>(eq, fR, qR, dir);
      const double lambda_half = 0.5 * dt / dx;
      const Complete<T, D>& q = sign < 0 ? qL : qR;
      auto add = [&](T& r, const T& u, const T& fl, const T& fr) {
        const T dF = fl - fr;
        const T dU = lambda_half * dF;
        r = u + dU;
      };
      add(rec.rho, q.rho, fL.rho, fR.rho);
      for (int i = 0; i < D; ++i) add(rec.mom[i], q.mom[i], fL.mom[i], fR.mom[i]);
      add(rec.E, q.E, fL.E, fR.E);
      CompleteFromCons<T, D>(eq, rec, rec);
    } else if constexpr (KIND == kMh2Primitive) { // :343-388
      const Prim w = prim_of(q0);
      const T lambda(sign * dt / dx);
      const T dx_half(sign * 0.5 * dx);
      const T u = w.u[dir];
      const T rho = q0.rho;
      const T a = q0.a;
      const T a2 = a * a;
      const T rho_a2 = rho * a2;
      const T& d_rho = grad.g[0];
      const T& d_p = grad.g[D + 1];
      const T* d_u = &grad.g[1];
      Prim r;
      r.rho = w.rho + dx_half * (d_rho - lambda * (u * d_rho + rho * d_u[dir]));
      r.p = w.p + dx_half * (d_p - lambda * (rho_a2 * d_u[dir] + u * d_p));
      r.u[dir] = w.u[dir] + dx_half * (d_u[dir] - lambda * (u * d_u[dir] + d_p / rho));
      for (int i = 1; i < D; ++i) {
        const int iy = (dir + i) % D;
        r.u[iy] = w.u[iy] + dx_half * (d_u[iy] - lambda * u * d_u[iy]);
      }
      from_prim(rec, r);
    } else { // :458-516
      Prim w = prim_of(q0);
      const T u = w.u[dir];
      const T c = q0.a;
      const T lambda(sign * dt / dx);
      const T dx_half(sign * 0.5 * dx);
      const T one(1.0);
      Grad k;
      k.g[0] = dx_half * grad.g[0] * (one - lambda * (u - c));         // minus
      k.g[D + 1] = dx_half * grad.g[D + 1] * (one - lambda * (u + c)); // plus
      for (int i = 0; i < D; ++i) k.g[1 + i] = dx_half * grad.g[1 + i] * (one - lambda * u);
      const T rho = w.rho;
      const T rhoc = rho * c;
      const T c2 = c * c;
      Prim dw;
      dw.p = 0.5 * (k.g[0] + k.g[D + 1]);
      dw.rho = dw.p / c2 + k.g[1];
      dw.u[dir] = 0.5 * (k.g[D + 1] - k.g[0]) / rhoc;
      for (int i = 1; i < D; ++i) {
        const int iy = (dir + i) % D;
        dw.u[iy] = k.g[1 + i];
      }
      w.rho = w.rho + dw.rho;
      w.p = w.p + dw.p;
      for (int i = 0; i < D; ++i) w.u[i] = w.u[i] + dw.u[i];
      from_prim(rec, w);
    }
  }
  // PerfectGas::CompleteFromPrim (array), src/equations/PerfectGas.cpp:160-174
  void from_prim(Complete<T, D>& q, const Prim& w) const {
    q.rho = w.rho;
    for (int i = 0; i < D; ++i) q.mom[i] = w.rho * w.u[i];
    q.p = w.p;
    const T e_kin = KineticEnergy<T, D>(q.rho, q.mom);
    q.E = e_kin + w.p * eq.gamma_minus_1_inv;
    q.a = vsqrt(eq.gamma * q.p / q.rho);
  }

  // MusclHancockMethod2.hpp:55-75
  T ComputeStableDt(const Complete<T, D>* states, double dx, int dir) const {
    return base.ComputeStableDt(states + 1, dx, dir);
  }
  T ComputeStableDt(const Complete<T, D>* states, const T& face_fraction, const T* vols,
                    double dx, int dir) const {
    return base.ComputeStableDt(states + 1, face_fraction, vols + 1, dx, dir);
  }
  // :164-191
  void ComputeNumericFlux(Cons<T, D>& flux, const Complete<T, D>* st, double dt, double dx,
                          int dir) const {
    Grad grad[2];
    ComputeGradient(grad[0], st, dx, dir);
    ComputeGradient(grad[1], st + 1, dx, dir);
    Complete<T, D> rec[2];
    Reconstruct(rec[0], st[1], grad[0], dt, dx, dir, +1);
    Reconstruct(rec[1], st[2], grad[1], dt, dx, dir, -1);
    base.ComputeNumericFlux(flux, rec, dt, dx, dir);
  }
  // :193-239 (masked: only the gradients see the volume fractions)
  void ComputeNumericFlux(Cons<T, D>& flux, const T& face_fractions, const Complete<T, D>* st,
                          const T* vols, double dt, double dx, int dir) const {
    Grad grad[2];
    ComputeGradient(grad[0], st, dx, dir);
    ComputeGradient(grad[1], st + 1, dx, dir);
    const M8 left_mask = (vols[0] > 0.0) && (vols[1] > 0.0) && (vols[2] > 0.0);
    const M8 right_mask = (vols[1] > 0.0) && (vols[2] > 0.0) && (vols[3] > 0.0);
    for (int k = 0; k < D + 2; ++k) {
      grad[0].g[k] = select(left_mask, grad[0].g[k], 0.0);
      grad[1].g[k] = select(right_mask, grad[1].g[k], 0.0);
    }
    Complete<T, D> rec[2];
    Reconstruct(rec[0], st[1], grad[0], dt, dx, dir, +1);
    Reconstruct(rec[1], st[2], grad[1], dt, dx, dir, -1);
    base.ComputeNumericFlux(flux, face_fractions, rec, vols + 1, dt, dx, dir);
  }
};

} // namespace fvso
