// TEST INFRASTRUCTURE -- not part of the product path.
//
// The reference's SCALAR HLLEM code path, restated independently of the array
// restatement in numerics.hpp, from
//   src/equations/perfect_gas/HllemMethod.hpp:371-640  Hllem<Eq, Larrouturou>::ComputeNumericFlux
//                                                       (Conservative&, span<const Complete, 2>, ...)
//   src/equations/PerfectGas.cpp:66-77                  Flux
// for Rank 1 and 2 (the scalar 3-D branch never assigns flux.density and
// flux.momentum[ix], HllemMethod.hpp:613-640, SURVEY quirk Q1, and is left out).
//
// Purpose: pin the array restatement.  The reference's own cross-check of its two
// code paths (tests/test.simd_FluxMethod.cpp:54-105) exists for HLL only; for HLLEM
// the two paths are NOT bit-identical upstream and this file reproduces the
// differences instead of hiding them:
//   * Roe averages: the scalar path computes (sqrt(rhoL) x_L + sqrt(rhoR) x_R) / sum
//     (:427-429, :432), the array path (sqrt(rhoL) / sum) x_L + (sqrt(rhoR) / sum) x_R
//     (:779-786) -- equal up to rounding;
//   * gm1 / roeA2 * X is evaluated left to right per use (:470-472), the array path
//     shares gm1_over_roeA2 -- the same operations, so bit-identical;
//   * Larrouturou upwinding (quirk Q6): the scalar path takes
//     f_rho * (upwind * vL - (1 - upwind) * vR) (:543-546), i.e. MINUS f_rho * vR for a
//     flux towards the left, the array path f_rho * select(upwind, vL, vR) (:907-910).
//     The array path is the one the reference runs (FluxMethodAdapter.hpp:54-56
//     defaults to the SIMD tag) and the one the product reproduces.
// tests/test_oracle_golden.py compares the two restatements: bit-identical where the
// operation sequences coincide, within a few ulp where only the Roe-average order
// differs, and with the documented sign where Q6 applies.
#pragma once
#include <algorithm>
#include <cmath>

namespace fvso_scalar {

template <int D> struct State {
  double rho, mom[D], E, p, a;
};
template <int D> struct Flux {
  double rho, mom[D], E;
};

// PerfectGas<D>::Flux, src/equations/PerfectGas.cpp:66-77
template <int D> inline void EulerFlux(Flux<D>& f, const State<D>& s, int d0) {
  const double velocity = s.mom[d0] / s.rho;
  f.rho = s.mom[d0];
  for (int d = 0; d < D; ++d) f.mom[d] = velocity * s.mom[d];
  f.mom[d0] += s.p;
  f.E = velocity * (s.E + s.p);
}

template <int D> inline double SquaredNorm(const double (&v)[D]) {
  // Eigen's squaredNorm on a fixed-size vector: D = 1: x0^2; D = 2: x0^2 + x1^2
  double s = v[0] * v[0];
  for (int i = 1; i < D; ++i) s += v[i] * v[i];
  return s;
}

// HllemMethod.hpp:371-640, Msq = 1 (PerfectGas, EulerEquation.hpp:53-62)
template <int D>
inline void HllemScalar(Flux<D>& flux, State<D> left, State<D> right, double gamma, int d,
                        bool larrouturou) {
  static_assert(D == 1 || D == 2, "scalar path: Rank 1 and 2");
  const double Msq = 1.0;
  const double Minv = 1.0 / std::sqrt(Msq);
  const double rhoL = left.rho;
  const double rhoR = right.rho;
  double vL_[D], vR_[D];
  for (int i = 0; i < D; ++i) {
    vL_[i] = left.mom[i] / rhoL;
    vR_[i] = right.mom[i] / rhoR;
  }
  double eKinRL = 0.0, eKinRR = 0.0;
  if (larrouturou) { // :391-404
    for (int i = 0; i < D; ++i) {
      const int mask = i == d;
      eKinRL += (1 - mask) * vL_[i] * vL_[i];
      eKinRR += (1 - mask) * vR_[i] * vR_[i];
      left.mom[i] *= mask;
      right.mom[i] *= mask;
    }
    eKinRL *= 0.5 * Msq;
    eKinRR *= 0.5 * Msq;
    left.E -= rhoL * eKinRL;
    right.E -= rhoR * eKinRR;
  }
  Flux<D> fluxL, fluxR;
  EulerFlux<D>(fluxL, left, d);
  EulerFlux<D>(fluxR, right, d);
  // :410-447 Einfeldt signal velocities
  const double uL = left.mom[d] / rhoL;
  const double uR = right.mom[d] / rhoR;
  const double aL = left.a, aR = right.a;
  const double rhoEL = left.E, rhoER = right.E;
  const double pL = left.p, pR = right.p;
  const double hL = (rhoEL + pL) / rhoL;
  const double hR = (rhoER + pR) / rhoR;
  const double sqRhoL = std::sqrt(rhoL);
  const double sqRhoR = std::sqrt(rhoR);
  const double sqRhoSum = sqRhoL + sqRhoR;
  double vL[D], vR[D], roeU[D];
  for (int i = 0; i < D; ++i) {
    vL[i] = left.mom[i] / rhoL;
    vR[i] = right.mom[i] / rhoR;
    roeU[i] = (sqRhoL * vL[i] + sqRhoR * vR[i]) / sqRhoSum;
  }
  const double roeU0 = roeU[d];
  const double roeH = (sqRhoL * hL + sqRhoR * hR) / sqRhoSum;
  const double roeGamma = (sqRhoL * gamma + sqRhoR * gamma) / sqRhoSum;
  const double gm1 = roeGamma - 1.0;
  const double roeA2 = gm1 * (roeH - 0.5 * Msq * SquaredNorm<D>(roeU));
  const double roeA = std::sqrt(roeA2);
  const double maxA = std::max(aL, aR);
  const double sL1 = uL - maxA * Minv;
  const double sL2 = roeU0 - roeA * Minv;
  const double sR1 = roeU0 + roeA * Minv;
  const double sR2 = uR + maxA * Minv;
  const double sL = std::min(sL1, sL2);
  const double sR = std::max(sR1, sR2);
  const double bL = std::min(sL, 0.0);
  const double bR = std::max(sR, 0.0);
  const double bLbR = bL * bR;
  const double db = bR - bL;
  const double db_positive = int(db <= 0) + int(db > 0) * db;
  // :453-457 HLLE flux, component by component
  Flux<D> hlle;
  hlle.rho = (bR * fluxL.rho - bL * fluxR.rho + bLbR * (right.rho - left.rho)) / db_positive;
  for (int i = 0; i < D; ++i) {
    hlle.mom[i] =
        (bR * fluxL.mom[i] - bL * fluxR.mom[i] + bLbR * (right.mom[i] - left.mom[i])) / db_positive;
  }
  hlle.E = (bR * fluxL.E - bL * fluxR.E + bLbR * (right.E - left.E)) / db_positive;
  // :459-468
  const double squaredNormRoeU = Msq * SquaredNorm<D>(roeU);
  const double squaredNormRoeU_half = 0.5 * squaredNormRoeU;
  const double u_bar = 0.5 * (sR + sL);
  const double u_bar_abs = std::abs(u_bar);
  const double delta = roeA / (roeA + u_bar_abs);
  const double b = bLbR / db_positive;
  const double deltaRho = rhoR - rhoL;
  double deltaRhoU[D];
  for (int i = 0; i < D; ++i) deltaRhoU[i] = right.mom[i] - left.mom[i];
  const double deltaRhoE = rhoER - rhoEL;
  if constexpr (D == 1) { // :470-479
    const double l21 = gm1 / roeA2 * (roeH - Msq * SquaredNorm<D>(roeU));
    const double l22 = gm1 / roeA2 * Msq * roeU[0];
    const double l23 = gm1 / roeA2 * (-1.0);
    const double alpha_2 = l21 * deltaRho + l22 * deltaRhoU[0] + l23 * deltaRhoE;
    const double b_delta_alpha_2 = b * delta * alpha_2;
    flux.rho = hlle.rho - b_delta_alpha_2 * 1.0;
    flux.mom[0] = hlle.mom[0] - b_delta_alpha_2 * roeU[0];
    flux.E = hlle.E - b_delta_alpha_2 * squaredNormRoeU_half;
  } else { // :525-563
    const int ix = d;
    const int iy = (ix + 1) % 2;
    const double l21 = gm1 / roeA2 * (roeH - Msq * SquaredNorm<D>(roeU));
    const double l22 = gm1 / roeA2 * Msq * roeU[0];
    const double l23 = gm1 / roeA2 * roeU[1];
    const double l24 = gm1 / roeA2 * (-1.0);
    const double alpha_2 =
        l21 * deltaRho + l22 * deltaRhoU[0] + l23 * deltaRhoU[1] + l24 * deltaRhoE;
    const double b_delta = b * delta;
    const double b_delta_alpha_2 = b_delta * alpha_2;
    flux.rho = hlle.rho - b_delta_alpha_2;
    flux.mom[ix] = hlle.mom[ix] - b_delta_alpha_2 * roeU[ix];
    if (larrouturou) {
      const double f_rho = flux.rho;
      const int upwind = (f_rho > 0.0);
      const double f_hllem_rhov = f_rho * (upwind * vL_[iy] - (1 - upwind) * vR_[iy]);
      const double f_hllem_rhoE = f_rho * (upwind * eKinRL - (1 - upwind) * eKinRR);
      flux.mom[iy] = f_hllem_rhov;
      flux.E = hlle.E - b_delta_alpha_2 * squaredNormRoeU_half + f_hllem_rhoE;
    } else {
      const double l31 = -roeU[iy];
      const double l33 = 1;
      const double alpha_3 = l31 * deltaRho + l33 * deltaRhoU[iy];
      const double b_delta_alpha_3 = b_delta * alpha_3;
      flux.mom[iy] = hlle.mom[iy] - b_delta_alpha_2 * roeU[iy] - b_delta_alpha_3 * 1.0;
      flux.E = hlle.E - b_delta_alpha_2 * squaredNormRoeU_half - b_delta_alpha_3 * roeU[iy];
    }
  }
}

} // namespace fvso_scalar
