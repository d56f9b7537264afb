// numerics.cuh -- per-cell / per-face fp64 numerics of the PerfectGas split
// sweep, one value per thread.  Compiled for sm_100a with -fmad=false: every
// + - * / sqrt below is a single IEEE-754 round-to-nearest operation in the
// order the reference's array (Eigen) code path evaluates it, so results are
// bit-identical to the reference's CPU path (DESIGN.md "Bitwise strategy").
//
// Space dimension D and sweep direction DIR are template parameters so that
// every momentum index is a compile-time constant and states stay in
// registers.
//
// Reference (paths relative to maikel/FiniteVolumeSolver):
//   src/equations/PerfectGas.cpp:31-144              KineticEnergy, Flux, CompleteFromCons
//   include/fub/flux_method/MusclHancockMethod.hpp:31-187,419-485  limiters, MUSCL-Hancock
//   src/equations/perfect_gas/HllemMethod.hpp:711-1072,1152-1216   HLLEM flux, stable dt
//   src/equations/perfect_gas/EinfeldtSignalVelocities.hpp:64-97   Einfeldt signals
//   include/fub/flux_method/HllMethod.hpp:260-293,337-347          HLL flux, stable dt
//   src/equations/perfect_gas/ExactRiemannSolver.hpp:34-353        exact solver
//   include/fub/flux_method/GodunovMethod.hpp:129-199              Godunov
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace fvs {

enum : int { kBaseHllemLarrouturou = 0, kBaseHllem = 1, kBaseHllEinfeldt = 2, kBaseGodunovExact = 3 };
enum : int {
  kReconFirstOrder = 0,
  kReconMusclHancock = 1,
  kReconMusclHancockPrim = 2,
  kReconMusclHancockChar = 3,
  // MusclHancock2<Equation, Gradient, Reconstruction, Base>
  // (include/fub/flux_method/MusclHancockMethod2.hpp:35-239): gradient and reconstruction
  // in conservative, primitive or characteristic variables; the limiter of the
  // CentralDifferenceGradient is a run-time value (EqConst::limiter2)
  kReconMh2Conservative = 4,
  kReconMh2Primitive = 5,
  kReconMh2Characteristics = 6
};
enum : int { kLimiterNone = 0, kLimiterMinMod = 1, kLimiterVanLeer = 2 };
// include/fub/flux_method/Gradient.hpp:31-59
enum : int { kLimiter2None = 0, kLimiter2Upwind = 1, kLimiter2MinMod = 2, kLimiter2VanLeer = 3 };
__host__ __device__ constexpr bool is_mh2(int recon) { return recon >= kReconMh2Conservative; }

// PerfectGas equation constants (include/fub/equations/PerfectGas.hpp:216-219):
// gamma_minus_1_inv = 1.0 / (gamma - 1.0) evaluated in fp64 on the host.
struct EqConst {
  double gamma;
  double gamma_minus_1_inv;
  double gamma_minus_1; // gamma - 1.0
  // per launch, read by MusclHancock2 only: cell size along the sweep and the limiter
  double dx;
  int limiter2;
};

template <int D> struct Cons {
  double rho;
  double mom[D];
  double E;
};
template <int D> struct Complete {
  double rho;
  double mom[D];
  double E;
  double p;
  double a;
};

#define FVS_DEV __device__ __forceinline__
} // namespace fvs
#include "fastmath.cuh"
namespace fvs {

// Equation constants plus the prepared reciprocal of gamma_minus_1_inv (the
// reference divides by it, PerfectGas.cpp:121).
struct EqM {
  double gamma;
  double gamma_minus_1;
  Rcp gm1_inv;
  FVS_DEV explicit EqM(const EqConst& e)
      : gamma(e.gamma), gamma_minus_1(e.gamma_minus_1), gm1_inv(FastMath::rcp(e.gamma_minus_1_inv)) {}
};

// Eigen-style min/max (array8 semantics of the oracle): a < b ? a : b
FVS_DEV double vmin(double a, double b) { return a < b ? a : b; }
FVS_DEV double vmax(double a, double b) { return a > b ? a : b; }

// Unrolled reduction order of Eigen's squaredNorm on a fixed-size column:
// D=1: x0; D=2: x0+x1; D=3: x0+(x1+x2).
template <int D> FVS_DEV double squared_norm(const double (&m)[D]) {
  if constexpr (D == 1) {
    return m[0] * m[0];
  } else if constexpr (D == 2) {
    return m[0] * m[0] + m[1] * m[1];
  } else {
    return m[0] * m[0] + (m[1] * m[1] + m[2] * m[2]);
  }
}

// src/equations/PerfectGas.cpp:66-77; `velocity` = s.mom[DIR] / s.rho
template <int D, int DIR, typename State>
FVS_DEV void euler_flux_v(Cons<D>& f, const State& s, double velocity) {
  f.rho = s.mom[DIR];
#pragma unroll
  for (int d = 0; d < D; ++d) f.mom[d] = velocity * s.mom[d];
  f.mom[DIR] = f.mom[DIR] + s.p;
  f.E = velocity * (s.E + s.p);
}

// src/equations/PerfectGas.cpp:31-37,112-124: fills p and a of `q` from its
// conservative part; rc = rcp(q.rho), checked with M::check_window by the caller
// (the kinetic energy of a gas at rest is a zero numerator).
template <class M, int D>
FVS_DEV void complete_from_cons(const EqM& eq, Complete<D>& q, const Rcp& rc, Flag& fl) {
  const double e_kin = M::divz(0.5 * squared_norm<D>(q.mom), rc, fl);
  const double e_int = q.E - e_kin;
  q.p = M::divw(e_int, eq.gm1_inv, fl); // 2^-60 <= gamma - 1 <= 2^60, checked by the host
  q.a = M::sqrt(M::divw(eq.gamma * q.p, rc, fl), fl);
}

// ---------------------------------------------------------------------------
// Limiters, include/fub/flux_method/MusclHancockMethod.hpp:31-187 (array path)
template <class M, int LIM>
FVS_DEV double limited_slope(double qL, double qM, double qR, Flag& fl) {
  const double sL = qM - qL;
  const double sR = qR - qM;
  if constexpr (LIM == kLimiterNone) {
    return 0.5 * (sL + sR); // :49-51
  } else if constexpr (LIM == kLimiterMinMod) {
    // :117 (sR > 0).select(zero.max(sL.min(sR)), zero.min(sL.max(sR)))
    return sR > 0.0 ? vmax(0.0, vmin(sL, sR)) : vmin(0.0, vmax(sL, sR));
  } else {
    const double r = 2.0 * sL * sR; // :144-147
    const bool pos = r > 0.0;
    const double q = M::div(r, sL + sR, fl, pos);
    return pos ? q : 0.0;
  }
}

// A predicted edge state together with the reciprocal of its density (the
// base flux divides several numerators by it).
template <int D> struct Edge {
  Complete<D> q;
  Rcp rc;
};

// MUSCL-Hancock predictor of ONE cell (MusclHancockMethod.hpp:428-451 and its
// mirror :456-479).  The reference evaluates this once for each of the two
// faces of the cell; all operations that lead to the two edge states are
// identical, so both are produced here in one go:
//   recL: predicted state at the cell's lower face  (q_L + lambda/2 (f(q_L) - f(q_R)))
//   recR: predicted state at the cell's upper face  (q_R + lambda/2 (f(q_L) - f(q_R)))
// Masked CompleteFromCons of the cut-cell path (PerfectGas.cpp:126-144 on a
// valid face): a state whose density is not positive becomes all zeros, and
// the masked Flux (PerfectGas.cpp:79-93) then divides by 1 instead.
// NEED_A = false: the caller only uses the pressure (the trial states q_L, q_R
// of the predictor feed Flux(), which does not read the speed of sound), so
// the division and square root behind `a` are not evaluated at all.
template <class M, int D, bool MASKED, bool NEED_A = true>
FVS_DEV void complete_edge(const EqM& eq, Complete<D>& q, Rcp& rc, Flag& fl) {
  if constexpr (MASKED) {
    if (!(q.rho > 0.0)) {
      q.rho = 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) q.mom[d] = 0.0;
      q.E = 0.0;
      q.p = 0.0;
      q.a = 0.0;
      rc = M::rcp(1.0);
      return;
    }
  }
  M::check_window(rc, fl);
  if constexpr (NEED_A) {
    complete_from_cons<M, D>(eq, q, rc, fl);
  } else {
    const double e_kin = M::divz(0.5 * squared_norm<D>(q.mom), rc, fl);
    const double e_int = q.E - e_kin;
    q.p = M::divw(e_int, eq.gm1_inv, fl);
    q.a = 0.0;
  }
}

template <class M, int D, int DIR, int LIM, bool MASKED = false, typename S0, typename S1,
          typename S2>
FVS_DEV void muscl_predict(const EqM& eq, const S0& qm, const S1& q0, const S2& qp,
                           double lambda_half, Edge<D>& recL, Edge<D>& recR, Flag& fl,
                           bool slope_ok = true) {
  Complete<D> qL, qR;
  {
    double s = limited_slope<M, LIM>(qm.rho, q0.rho, qp.rho, fl);
    if constexpr (MASKED) s = slope_ok ? s : 0.0; // MusclHancockMethod.hpp:499-503
    qL.rho = q0.rho - 0.5 * s;
    qR.rho = q0.rho + 0.5 * s;
  }
#pragma unroll
  for (int d = 0; d < D; ++d) {
    double s = limited_slope<M, LIM>(qm.mom[d], q0.mom[d], qp.mom[d], fl);
    if constexpr (MASKED) s = slope_ok ? s : 0.0;
    qL.mom[d] = q0.mom[d] - 0.5 * s;
    qR.mom[d] = q0.mom[d] + 0.5 * s;
  }
  {
    double s = limited_slope<M, LIM>(qm.E, q0.E, qp.E, fl);
    if constexpr (MASKED) s = slope_ok ? s : 0.0;
    qL.E = q0.E - 0.5 * s;
    qR.E = q0.E + 0.5 * s;
  }
  Rcp rcL = M::rcp(qL.rho);
  Rcp rcR = M::rcp(qR.rho);
  complete_edge<M, D, MASKED, false>(eq, qL, rcL, fl);
  complete_edge<M, D, MASKED, false>(eq, qR, rcR, fl);
  Cons<D> fL, fR;
  euler_flux_v<D, DIR>(fL, qL, M::divz(qL.mom[DIR], rcL, fl));
  euler_flux_v<D, DIR>(fR, qR, M::divz(qR.mom[DIR], rcR, fl));
  {
    const double c = lambda_half * (fL.rho - fR.rho);
    recL.q.rho = qL.rho + c;
    recR.q.rho = qR.rho + c;
  }
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double c = lambda_half * (fL.mom[d] - fR.mom[d]);
    recL.q.mom[d] = qL.mom[d] + c;
    recR.q.mom[d] = qR.mom[d] + c;
  }
  {
    const double c = lambda_half * (fL.E - fR.E);
    recL.q.E = qL.E + c;
    recR.q.E = qR.E + c;
  }
  recL.rc = M::rcp(recL.q.rho);
  recR.rc = M::rcp(recR.q.rho);
  complete_edge<M, D, MASKED>(eq, recL.q, recL.rc, fl);
  complete_edge<M, D, MASKED>(eq, recR.q, recR.rc, fl);
}

// ---------------------------------------------------------------------------
// perfect_gas::MusclHancockPrim, src/equations/perfect_gas/MusclHancockPrimMethod.hpp:59-179:
// MUSCL-Hancock on the primitive variables (rho, u, p) with its own limiter;
// the base method is always HLLEM with the Larrouturou modification (:175).
template <int D> struct PrimCell {
  double rho, p, a;
  double u[D];
};
template <class M, int D, typename State>
FVS_DEV void prim_cell(PrimCell<D>& c, const State& q, Flag& fl) {
  const Rcp rc = M::rcp(q.rho);
  M::check_window(rc, fl);
  c.rho = q.rho;
  c.p = q.p;
  c.a = q.a;
#pragma unroll
  for (int i = 0; i < D; ++i) c.u[i] = M::divz(q.mom[i], rc, fl);
}
// the limiter lambda of :70-86
template <class M> FVS_DEV double prim_limited_slope(double qL, double qM, double qR, Flag& fl) {
  const double delta_q = 0.5 * (qR - qL);
  const bool is_relevant = fabs(delta_q) > 1e-12;
  const double dq = is_relevant ? delta_q : 1.0;
  const Rcp rc = M::rcp(dq);
  M::check(rc, fl);
  const double qLq = M::div_any(qM - qL, rc, fl, is_relevant);
  const double qRq = M::div_any(qR - qM, rc, fl, is_relevant);
  const double rL = is_relevant ? qLq : 0.0;
  const double rR = is_relevant ? qRq : 0.0;
  const bool is_positive = rL > 0.0 && rR > 0.0;
  const double r = is_positive ? vmin(rL, rR) : 0.0;
  const Rcp rc1 = M::rcp(1.0 + r); // in [1, 2): inside the window
  const double r_1 = M::divz(2 * r, rc1, fl);
  const double r_2 = 0.5 * (r + 1.0);
  const double sigma = r < 1.0 ? r_1 : r_2;
  return sigma * delta_q;
}
template <int D> struct PrimSlopes {
  double drho, dp;
  double du[D];
};
template <class M, int D>
FVS_DEV void prim_slopes(PrimSlopes<D>& s, const PrimCell<D>& m, const PrimCell<D>& c,
                         const PrimCell<D>& p, Flag& fl) {
  s.drho = prim_limited_slope<M>(m.rho, c.rho, p.rho, fl);
  s.dp = prim_limited_slope<M>(m.p, c.p, p.p, fl);
#pragma unroll
  for (int i = 0; i < D; ++i) s.du[i] = prim_limited_slope<M>(m.u[i], c.u[i], p.u[i], fl);
}
// One predicted edge state of cell `c`: RIGHT = its upper edge (w[0] of the
// face above, :108-119), else its lower edge (w[1] of the face below, :121-132);
// then the Complete state of :139-158.
template <class M, int D, int DIR, bool RIGHT>
FVS_DEV void prim_edge(const EqM& eq, Edge<D>& out, const PrimCell<D>& c, const PrimSlopes<D>& s,
                       double lambda, Flag& fl) {
  const double ud = c.u[DIR];
  double rec_rho, rec_p, rec_u[D];
  if constexpr (RIGHT) {
    rec_rho = c.rho + 0.5 * (s.drho - lambda * (ud * s.drho + c.rho * s.du[DIR]));
    rec_p = c.p + 0.5 * (s.dp - lambda * (c.rho * c.a * c.a * s.du[DIR] + ud * s.dp));
#pragma unroll
    for (int i = 0; i < D; ++i) rec_u[i] = c.u[i] + 0.5 * (s.du[i] - lambda * (ud * s.du[i]));
  } else {
    rec_rho = c.rho - 0.5 * (s.drho + lambda * (ud * s.drho + c.rho * s.du[DIR]));
    rec_p = c.p - 0.5 * (s.dp + lambda * (c.rho * c.a * c.a * s.du[DIR] + ud * s.dp));
#pragma unroll
    for (int i = 0; i < D; ++i) rec_u[i] = c.u[i] - 0.5 * (s.du[i] + lambda * (ud * s.du[i]));
  }
  {
    const Rcp rcc = M::rcp(c.rho);
    M::check_window(rcc, fl);
    rec_u[DIR] = rec_u[DIR] - M::divz(0.5 * lambda * s.dp, rcc, fl);
  }
  out.q.rho = rec_rho;
#pragma unroll
  for (int i = 0; i < D; ++i) out.q.mom[i] = rec_rho * rec_u[i];
  out.q.p = rec_p;
  out.rc = M::rcp(rec_rho);
  M::check(out.rc, fl);
  out.q.a = M::sqrt(M::div(eq.gamma * rec_p, out.rc, fl), fl);
  out.q.E = 0.5 * rec_rho * squared_norm<D>(rec_u) + rec_p * eq.gm1_inv.b;
}
// both edge states of the cell with neighbours qm, qp (regular grid: every
// cell predictor is shared by the two faces of the cell)
template <class M, int D, int DIR, typename S0, typename S1, typename S2>
FVS_DEV void muscl_prim_predict(const EqM& eq, const S0& qm, const S1& q0, const S2& qp,
                                double lambda, Edge<D>& recL, Edge<D>& recR, Flag& fl) {
  PrimCell<D> m, c, p;
  prim_cell<M, D>(m, qm, fl);
  prim_cell<M, D>(c, q0, fl);
  prim_cell<M, D>(p, qp, fl);
  PrimSlopes<D> s;
  prim_slopes<M, D>(s, m, c, p, fl);
  prim_edge<M, D, DIR, false>(eq, recL, c, s, lambda, fl);
  prim_edge<M, D, DIR, true>(eq, recR, c, s, lambda, fl);
}
// masked face of a cut-cell fab (:185-329): an outer stencil cell whose volume
// fraction is not positive is replaced by its inner neighbour
template <class M, int D, int DIR, typename S>
FVS_DEV void muscl_prim_face_masked(const EqM& eq, const S& s0, const S& s1, const S& s2,
                                    const S& s3, bool left_ok, bool right_ok, double lambda,
                                    Edge<D>& left, Edge<D>& right, Flag& fl) {
  PrimCell<D> ll, l, r, rr;
  prim_cell<M, D>(l, s1, fl);
  prim_cell<M, D>(r, s2, fl);
  if (left_ok) {
    prim_cell<M, D>(ll, s0, fl);
  } else {
    ll = l;
  }
  if (right_ok) {
    prim_cell<M, D>(rr, s3, fl);
  } else {
    rr = r;
  }
  PrimSlopes<D> sl, sr;
  prim_slopes<M, D>(sl, ll, l, r, fl);
  prim_slopes<M, D>(sr, l, r, rr, fl);
  prim_edge<M, D, DIR, true>(eq, left, l, sl, lambda, fl);
  prim_edge<M, D, DIR, false>(eq, right, r, sr, lambda, fl);
}

// ---------------------------------------------------------------------------
// perfect_gas::MusclHancockCharacteristic,
// src/equations/perfect_gas/MusclHancockCharacteristicMethod.hpp:28-359: slopes of
// the characteristic wave amplitudes (VanLeer), half time step, recombination
// to primitive variables; base method HLLEM with the Larrouturou modification.
template <int D> struct CharAmpl {
  double minus, zero, plus, t[2]; // t: the transverse velocity differences (iy, iz)
};
// VanLeer::operator()(sL, sR), MusclHancockMethod.hpp:144-149
template <class M> FVS_DEV double vanleer2(double sL, double sR, Flag& fl) {
  const double r = 2.0 * sL * sR;
  const bool pos = r > 0.0;
  const double q = M::div(r, sL + sR, fl, pos);
  return pos ? q : 0.0;
}
template <int D, int DIR>
FVS_DEV void char_amplitudes(CharAmpl<D>& a, const PrimCell<D>& l, const PrimCell<D>& r,
                             double rhoc, double ooc2) {
  constexpr int iy = (DIR + 1) % 3, iz = (iy + 1) % 3;
  const double dp = r.p - l.p;
  const double du = r.u[DIR] - l.u[DIR];
  const double drho = r.rho - l.rho;
  a.minus = dp - rhoc * du;
  a.zero = drho - ooc2 * dp;
  a.plus = dp + rhoc * du;
  if constexpr (iy < D) a.t[0] = r.u[iy] - l.u[iy];
  if constexpr (iz < D) a.t[1] = r.u[iz] - l.u[iz];
}
// one edge state: SIGN = +1 upper edge (is_right_side = 0), -1 lower edge
template <class M, int D, int DIR, int SIGN, bool MASKED>
FVS_DEV void char_edge(const EqM& eq, Edge<D>& out, const PrimCell<D>& c, const CharAmpl<D>& slope,
                       double lambda, double rhoc, double ooc2, Flag& fl) {
  constexpr int iy = (DIR + 1) % 3, iz = (iy + 1) % 3;
  const double u = c.u[DIR];
  const double sg = (double)SIGN;
  const double minus = sg * 0.5 * slope.minus * (1.0 - sg * lambda * (u - c.a));
  const double zero = sg * 0.5 * slope.zero * (1.0 - sg * lambda * u);
  const double plus = sg * 0.5 * slope.plus * (1.0 - sg * lambda * (u + c.a));
  double dp = 0.5 * (minus + plus);
  double drho = ooc2 * dp + zero;
  double du[D];
  {
    const Rcp rcc = M::rcp(rhoc);
    M::check_window(rcc, fl);
    du[DIR] = M::divz(0.5 * (plus - minus), rcc, fl);
  }
  if constexpr (iy < D) du[iy] = sg * 0.5 * slope.t[0] * (1.0 - sg * lambda * u);
  if constexpr (iz < D) du[iz] = sg * 0.5 * slope.t[1] * (1.0 - sg * lambda * u);
  drho = drho + c.rho;
#pragma unroll
  for (int d = 0; d < D; ++d) du[d] = du[d] + c.u[d];
  dp = dp + c.p;
  // PerfectGas::CompleteFromPrim, src/equations/PerfectGas.cpp:160-174
  out.q.rho = drho;
#pragma unroll
  for (int d = 0; d < D; ++d) out.q.mom[d] = drho * du[d];
  out.q.p = dp;
  out.rc = M::rcp(drho);
  M::check_window(out.rc, fl);
  const double e_kin = M::divz(0.5 * squared_norm<D>(out.q.mom), out.rc, fl);
  out.q.E = e_kin + dp * eq.gm1_inv.b;
  out.q.a = M::sqrt(M::div(eq.gamma * dp, out.rc, fl), fl);
  (void)MASKED;
}
template <class M, int D, int DIR, bool MASKED = false, typename S0, typename S1, typename S2>
FVS_DEV void muscl_char_predict(const EqM& eq, const S0& qm, const S1& q0, const S2& qp,
                                double lambda, Edge<D>& recL, Edge<D>& recR, Flag& fl,
                                bool slope_ok = true) {
  PrimCell<D> m, c, p;
  prim_cell<M, D>(m, qm, fl);
  prim_cell<M, D>(c, q0, fl);
  prim_cell<M, D>(p, qp, fl);
  const double rhoc = c.rho * c.a;
  const double ooc2 = M::div(1.0, c.a * c.a, fl);
  CharAmpl<D> al, ar, slope;
  char_amplitudes<D, DIR>(al, m, c, rhoc, ooc2);
  char_amplitudes<D, DIR>(ar, c, p, rhoc, ooc2);
  if constexpr (MASKED) {
    if (!slope_ok) { // :147-152: only the LEFT amplitudes are zeroed
      al.minus = 0.0;
      al.zero = 0.0;
      al.plus = 0.0;
      al.t[0] = 0.0;
      al.t[1] = 0.0;
    }
  }
  constexpr int iy = (DIR + 1) % 3, iz = (iy + 1) % 3;
  slope.minus = vanleer2<M>(al.minus, ar.minus, fl);
  slope.zero = vanleer2<M>(al.zero, ar.zero, fl);
  slope.plus = vanleer2<M>(al.plus, ar.plus, fl);
  if constexpr (iy < D) slope.t[0] = vanleer2<M>(al.t[0], ar.t[0], fl);
  if constexpr (iz < D) slope.t[1] = vanleer2<M>(al.t[1], ar.t[1], fl);
  char_edge<M, D, DIR, -1, MASKED>(eq, recL, c, slope, lambda, rhoc, ooc2, fl);
  char_edge<M, D, DIR, +1, MASKED>(eq, recR, c, slope, lambda, rhoc, ooc2, fl);
}

// ---------------------------------------------------------------------------
// HLLEM, array code path.  src/equations/perfect_gas/HllemMethod.hpp:711-1072.
// Msq = Minv = 1 for PerfectGas (include/fub/equations/EulerEquation.hpp:53-62).
// Returns the numeric flux between the edge states `eL` and `eR`.
template <class M, int D, int DIR, bool LARROUTUROU>
FVS_DEV void hllem_flux(const EqM& eq, Cons<D>& flux, const Edge<D>& eL, const Edge<D>& eR,
                        Flag& fl) {
  Complete<D> left = eL.q;
  Complete<D> right = eR.q;
  const Rcp& rcL = eL.rc;
  const Rcp& rcR = eR.rc;
  M::check_window(rcL, fl);
  M::check_window(rcR, fl);
  const double rhoL = left.rho;
  const double rhoR = right.rho;
  double vL_[D], vR_[D]; // velocities of the unmodified states
  double uL[D], uR[D];   // velocities after the Larrouturou modification
  double eKinRL = 0.0, eKinRR = 0.0;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    vL_[i] = M::divz(left.mom[i], rcL, fl);
    vR_[i] = M::divz(right.mom[i], rcR, fl);
    uL[i] = vL_[i];
    uR[i] = vR_[i];
  }
  // Tolerance arithmetic with the Larrouturou modification: the tangential
  // momenta, velocities, Roe velocities and momentum jumps are zero, so every
  // term they enter is dropped instead of being evaluated to a signed zero.
  constexpr bool kDropTangential = LARROUTUROU && !M::kExact;
  if constexpr (LARROUTUROU) {
    // :729-745 -- tangential momenta are multiplied by 0.0 (keeps their sign
    // bit) and their kinetic energy is taken out of E.
#pragma unroll
    for (int i = 0; i < D; ++i) {
      if (i != DIR) {
        eKinRL += vL_[i] * vL_[i];
        eKinRR += vR_[i] * vR_[i];
        if constexpr (kDropTangential) {
          left.mom[i] = 0.0;
          right.mom[i] = 0.0;
          uL[i] = 0.0;
          uR[i] = 0.0;
        } else {
          left.mom[i] *= 0.0;
          right.mom[i] *= 0.0;
          uL[i] = M::div_zero(left.mom[i], rcL); // (+-0) / rho, :774-775
          uR[i] = M::div_zero(right.mom[i], rcR);
        }
      }
    }
    eKinRL *= 0.5;
    eKinRR *= 0.5;
    left.E -= rhoL * eKinRL;
    right.E -= rhoR * eKinRR;
  }
  Cons<D> fluxL, fluxR;
  euler_flux_v<D, DIR>(fluxL, left, uL[DIR]);
  euler_flux_v<D, DIR>(fluxR, right, uR[DIR]);

  // :752-797 Einfeldt signal velocities with the full Roe velocity
  const double aL = left.a;
  const double aR = right.a;
  const double rhoEL = left.E;
  const double rhoER = right.E;
  const double pL = left.p;
  const double pR = right.p;
  const double hL = M::divw(rhoEL + pL, rcL, fl);
  const double hR = M::divw(rhoER + pR, rcR, fl);
  const double sqRhoL = M::sqrt(rhoL, fl);
  const double sqRhoR = M::sqrt(rhoR, fl);
  const double sqRhoSum = sqRhoL + sqRhoR;
  const Rcp rcSum = M::rcp(sqRhoSum);
  M::check_window(rcSum, fl);
  const double sqRhoL_over_Sum = M::divw(sqRhoL, rcSum, fl);
  const double sqRhoR_over_Sum = M::divw(sqRhoR, rcSum, fl);
  double roeU[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    if (kDropTangential && i != DIR) {
      roeU[i] = 0.0;
    } else {
      roeU[i] = sqRhoL_over_Sum * uL[i] + sqRhoR_over_Sum * uR[i];
    }
  }
  const double roeH = sqRhoL_over_Sum * hL + sqRhoR_over_Sum * hR;
  double squaredNormRoeU = 0.0;
  if constexpr (kDropTangential) {
    squaredNormRoeU = roeU[DIR] * roeU[DIR];
  } else {
#pragma unroll
    for (int i = 0; i < D; ++i) squaredNormRoeU += roeU[i] * roeU[i];
  }
  const double squaredNormRoeU_half = 0.5 * squaredNormRoeU;
  // (sqrt(rhoL) gamma + sqrt(rhoR) gamma) / (sqrt(rhoL) + sqrt(rhoR)) is gamma up to rounding
  double gm1;
  if constexpr (M::kExact) {
    const double roeGamma = M::divw(sqRhoL * eq.gamma + sqRhoR * eq.gamma, rcSum, fl);
    gm1 = roeGamma - 1.0;
  } else {
    gm1 = eq.gamma_minus_1;
  }
  const double roeA2 = gm1 * (roeH - squaredNormRoeU_half);
  const double roeA = M::sqrt(roeA2, fl);
  const double maxA = vmax(aL, aR);
  const double sL1 = uL[DIR] - maxA;
  const double sL2 = roeU[DIR] - roeA;
  const double sR1 = roeU[DIR] + roeA;
  const double sR2 = uR[DIR] + maxA;
  const double sL = vmin(sL1, sL2);
  const double sR = vmax(sR1, sR2);

  // :799-811 HLLE flux
  const double bL = vmin(sL, 0.0);
  const double bR = vmax(sR, 0.0);
  const double bLbR = bL * bR;
  const double db = bR - bL;
  const double db_positive = db > 0.0 ? db : 1.0;
  const Rcp rcDb = M::rcp(db_positive);
  M::check_window(rcDb, fl);
  Cons<D> hlle;
  // the mass and energy fluxes of a uniform gas at rest are zero numerators; the
  // normal momentum flux carries the pressure
  hlle.rho =
      M::divz(bR * fluxL.rho - bL * fluxR.rho + bLbR * (right.rho - left.rho), rcDb, fl);
#pragma unroll
  for (int d = 0; d < D; ++d) {
    // with the Larrouturou modification the tangential momentum fluxes are
    // upwinded below (:994-1004) and the HLLE values of those components,
    // which the reference computes and discards, are never read
    if (LARROUTUROU && D > 1 && d != DIR) {
      hlle.mom[d] = 0.0;
    } else {
      const double num =
          bR * fluxL.mom[d] - bL * fluxR.mom[d] + bLbR * (right.mom[d] - left.mom[d]);
      hlle.mom[d] = d == DIR ? M::divw(num, rcDb, fl) : M::divz(num, rcDb, fl);
    }
  }
  hlle.E = M::divz(bR * fluxL.E - bL * fluxR.E + bLbR * (right.E - left.E), rcDb, fl);

  // :813-823
  const double u_bar = 0.5 * (sR + sL);
  const double u_bar_abs = fabs(u_bar);
  const Rcp rcDelta = M::rcp(roeA + u_bar_abs);
  M::check_window(rcDelta, fl);
  const double delta = M::divw(roeA, rcDelta, fl);
  const double b = M::divz(bLbR, rcDb, fl);
  const double deltaRho = rhoR - rhoL;
  double deltaRhoU[D];
#pragma unroll
  for (int i = 0; i < D; ++i) deltaRhoU[i] = right.mom[i] - left.mom[i];
  const double deltaRhoE = rhoER - rhoEL;
  const Rcp rcA2 = M::rcp(roeA2);
  M::check_window(rcA2, fl);
  const double gm1_over_roeA2 = M::divw(gm1, rcA2, fl);
  const double l21 = gm1_over_roeA2 * (roeH - squaredNormRoeU);

  if constexpr (D == 1) {
    // :825-835
    const double l22 = gm1_over_roeA2 * roeU[0];
    const double l23 = gm1_over_roeA2 * (-1.0);
    const double alpha_2 = l21 * deltaRho + l22 * deltaRhoU[0] + l23 * deltaRhoE;
    const double b_delta_alpha_2 = b * delta * alpha_2;
    flux.rho = hlle.rho - b_delta_alpha_2;
    flux.mom[0] = hlle.mom[0] - b_delta_alpha_2 * roeU[0];
    flux.E = hlle.E - b_delta_alpha_2 * squaredNormRoeU_half;
  } else if constexpr (D == 2) {
    // :892-923
    constexpr int ix = DIR;
    constexpr int iy = (ix == 0) ? 1 : 0;
    const double l24 = -gm1_over_roeA2;
    double alpha_2;
    if constexpr (kDropTangential) {
      alpha_2 = l21 * deltaRho + (gm1_over_roeA2 * roeU[ix]) * deltaRhoU[ix] + l24 * deltaRhoE;
    } else {
      const double l22 = gm1_over_roeA2 * roeU[0];
      const double l23 = gm1_over_roeA2 * roeU[1];
      alpha_2 = l21 * deltaRho + l22 * deltaRhoU[0] + l23 * deltaRhoU[1] + l24 * deltaRhoE;
    }
    const double b_delta = b * delta;
    const double b_delta_alpha_2 = b_delta * alpha_2;
    flux.rho = hlle.rho - b_delta_alpha_2;
    flux.mom[ix] = hlle.mom[ix] - b_delta_alpha_2 * roeU[ix];
    if constexpr (LARROUTUROU) {
      const bool upwind = flux.rho > 0.0;
      flux.mom[iy] = flux.rho * (upwind ? vL_[iy] : vR_[iy]);
      const double f_rhoE = flux.rho * (upwind ? eKinRL : eKinRR);
      flux.E = hlle.E - b_delta_alpha_2 * squaredNormRoeU_half + f_rhoE;
    } else {
      const double l31 = -roeU[iy];
      const double alpha_3 = l31 * deltaRho + deltaRhoU[iy];
      const double b_delta_alpha_3 = b_delta * alpha_3;
      flux.mom[iy] = hlle.mom[iy] - b_delta_alpha_2 * roeU[iy] - b_delta_alpha_3;
      flux.E = hlle.E - b_delta_alpha_2 * squaredNormRoeU_half - b_delta_alpha_3 * roeU[iy];
    }
  } else {
    // :974-1019
    constexpr int ix = DIR;
    constexpr int iy = (ix + 1) % 3;
    constexpr int iz = (iy + 1) % 3;
    const double l25 = -gm1_over_roeA2;
    double alpha_2;
    if constexpr (kDropTangential) {
      alpha_2 = l21 * deltaRho + (gm1_over_roeA2 * roeU[ix]) * deltaRhoU[ix] + l25 * deltaRhoE;
    } else {
      const double l22 = gm1_over_roeA2 * roeU[0];
      const double l23 = gm1_over_roeA2 * roeU[1];
      const double l24 = gm1_over_roeA2 * roeU[2];
      alpha_2 = l21 * deltaRho + l22 * deltaRhoU[0] + l23 * deltaRhoU[1] + l24 * deltaRhoU[2] +
                l25 * deltaRhoE;
    }
    const double b_delta = b * delta;
    const double b_delta_alpha_2 = b_delta * alpha_2;
    flux.rho = hlle.rho - b_delta_alpha_2;
    // :991-992 uses roeU.row(iy), not row(ix): reference behaviour, reproduced
    // (with the Larrouturou modification roeU[iy] is zero)
    if constexpr (kDropTangential) {
      flux.mom[ix] = hlle.mom[ix];
    } else {
      flux.mom[ix] = hlle.mom[ix] - b_delta_alpha_2 * roeU[iy];
    }
    if constexpr (LARROUTUROU) {
      const bool upwind = flux.rho > 0.0;
      flux.mom[iy] = flux.rho * (upwind ? vL_[iy] : vR_[iy]);
      flux.mom[iz] = flux.rho * (upwind ? vL_[iz] : vR_[iz]);
      const double f_rhoE = flux.rho * (upwind ? eKinRL : eKinRR);
      flux.E = hlle.E - b_delta_alpha_2 * squaredNormRoeU_half + f_rhoE;
    } else {
      const double l31 = -roeU[iy];
      const double l41 = -roeU[iz];
      const double alpha_3 = l31 * deltaRho + deltaRhoU[iy];
      const double alpha_4 = l41 * deltaRho + deltaRhoU[iz];
      const double b_delta_alpha_3 = b_delta * alpha_3;
      const double b_delta_alpha_4 = b_delta * alpha_4;
      flux.mom[iy] = hlle.mom[iy] - b_delta_alpha_2 * roeU[iy] - b_delta_alpha_3;
      flux.mom[iz] = hlle.mom[iz] - b_delta_alpha_2 * roeU[iz] - b_delta_alpha_4;
      flux.E = hlle.E - b_delta_alpha_2 * squaredNormRoeU_half -
               b_delta_alpha_3 * roeU[iy] - b_delta_alpha_4 * roeU[iz];
    }
  }
}

// Per-cell part of the stable-dt signal evaluation: quantities that depend on
// one stored state only (HllemMethod.hpp:1152-1216 /
// EinfeldtSignalVelocities.hpp:64-97).  A cell takes part in two faces; they
// are computed once per cell and carried along the sweep.
template <int D> struct DtCell {
  double u[D];
  double h;
  double sq_rho;
  double a;
};
template <class M, int D, typename State>
FVS_DEV void dt_cell(DtCell<D>& c, const State& q, Flag& fl) {
  const Rcp rc = M::rcp(q.rho);
  M::check_window(rc, fl);
  c.h = M::div(q.E + q.p, rc, fl);
  c.sq_rho = M::sqrt(q.rho, fl);
#pragma unroll
  for (int i = 0; i < D; ++i) c.u[i] = M::divz(q.mom[i], rc, fl);
  c.a = q.a;
}

// max(|sL|, |sR|) of HllemMethod.hpp:1152-1216; the stable dt of the face is
// dx / (this value).
template <class M, int D, int DIR>
FVS_DEV double hllem_max_signal(const EqM& eq, const DtCell<D>& L, const DtCell<D>& R,
                                Flag& fl) {
  const double sqRhoSum = L.sq_rho + R.sq_rho;
  const Rcp rcSum = M::rcp(sqRhoSum);
  M::check(rcSum, fl);
  const double wL = M::div(L.sq_rho, rcSum, fl);
  const double wR = M::div(R.sq_rho, rcSum, fl);
  double roeU[D];
#pragma unroll
  for (int i = 0; i < D; ++i) roeU[i] = wL * L.u[i] + wR * R.u[i];
  const double roeH = wL * L.h + wR * R.h;
  double sq = 0.0;
#pragma unroll
  for (int i = 0; i < D; ++i) sq += roeU[i] * roeU[i];
  const double sq_half = 0.5 * sq;
  const double roeA2 = eq.gamma_minus_1 * (roeH - sq_half);
  const double roeA = M::sqrt(roeA2, fl);
  const double maxA = vmax(L.a, R.a);
  const double sL = vmin(L.u[DIR] - maxA, roeU[DIR] - roeA);
  const double sR = vmax(roeU[DIR] + roeA, R.u[DIR] + maxA);
  return vmax(fabs(sL), fabs(sR));
}

// ---------------------------------------------------------------------------
// Lazy stable dt.  The stable dt of a sweep is dx / max over faces of the signal
// speed S = max(|sL|, |sR|); only the faces near the maximum decide it.  For
// HLLEM and HLL-Einfeldt
//   |sL|, |sR| <= max(|uL|, |uR|) + max(aL, aR, roeA),  roeA^2 <= (gamma-1) max(hL, hR)
// up to a few roundings (the Roe averages are convex combinations whose weights
// sum to 1 within 2 ulp), so  B = max(|u|) + max(a, sqrt((gamma-1) h))  over the
// two cells bounds S from above.  B is evaluated from the 20-bit MUFU seeds
// (6 fp64 instructions per cell instead of ~85) and compared in the high words;
// the exact signal speed is only evaluated where B exceeds kLazyKeep times a
// known lower bound of the final maximum (the thread's running maximum, or the
// hint the host derives from the previous evaluation and verifies afterwards,
// context.cu: verify_hint).  All approximation errors (seeds 2^-19, high-word
// truncation 2^-19, roundings) stay far inside the 2^-10 margin, and anything
// that is not an ordinary positive number (NaN, infinity, negative density)
// turns B into NaN or infinity, which takes the exact path: the maximum is
// bit-identical to the eager evaluation.
struct SigBound {
  int u; // high word of an approximation of |u_DIR|
  int c; // high word of an approximation of max(a, sqrt((gamma-1) h))
};
template <int D> struct SigBoundAll {
  int u[D];
  int c;
};
constexpr double kLazyKeep = 1.0 - 0.0009765625; // 1 - 2^-10

FVS_DEV int imax(int a, int b) { return a > b ? a : b; }

// rs * rs ~ 1 / rho; NaN for a negative density, infinity for zero
template <int D, typename State>
FVS_DEV int sig_bound_c(const EqM& eq, const State& q, double r0) {
  const double hs = ((q.E + q.p) * r0) * eq.gamma_minus_1;
  const double ch = hs * mufu_rsq64h(hs); // sqrt((gamma-1) h); NaN for h <= 0
  return imax(hi32(q.a) & 0x7fffffff, hi32(ch));
}
template <int D, int DIR, typename State>
FVS_DEV SigBound sig_bound(const EqM& eq, const State& q) {
  const double rs = mufu_rsq64h(q.rho);
  const double r0 = rs * rs;
  SigBound b;
  b.u = hi32(fabs(q.mom[DIR]) * r0);
  b.c = sig_bound_c<D>(eq, q, r0);
  return b;
}
template <int D, typename State>
FVS_DEV SigBoundAll<D> sig_bound_all(const EqM& eq, const State& q) {
  const double rs = mufu_rsq64h(q.rho);
  const double r0 = rs * rs;
  SigBoundAll<D> b;
#pragma unroll
  for (int d = 0; d < D; ++d) b.u[d] = hi32(fabs(q.mom[d]) * r0);
  b.c = sig_bound_c<D>(eq, q, r0);
  return b;
}
// true iff the face between two cells may carry a signal speed above thr / kLazyKeep
FVS_DEV bool sig_bound_exceeds(int ul, int cl, int ur, int cr, double thr) {
  const double s = hilo(imax(ul, ur), 0) + hilo(imax(cl, cr), 0);
  return !(s <= thr);
}

// ---------------------------------------------------------------------------
// Einfeldt signal velocities (normal Roe velocity only),
// src/equations/perfect_gas/EinfeldtSignalVelocities.hpp:64-97, from the
// per-cell quantities uL = m_d / rho, hL = (E + p) / rho, sqrt(rho), a.
template <class M>
FVS_DEV void einfeldt_signals(const EqM& eq, double uL, double hL, double sqRhoL, double aL,
                              double uR, double hR, double sqRhoR, double aR, double& s0,
                              double& s1, Flag& fl) {
  const double oosqRhoSum = M::div(1.0, sqRhoL + sqRhoR, fl);
  const double roeU = (uL * sqRhoL + uR * sqRhoR) * oosqRhoSum;
  const double roeH = (hL * sqRhoL + hR * sqRhoR) * oosqRhoSum;
  const double roeA2 = (eq.gamma - 1.0) * (roeH - 0.5 * roeU * roeU);
  const double roeA = M::sqrt(roeA2, fl);
  s0 = vmin(uL - aL, roeU - roeA);
  s1 = vmax(roeU + roeA, uR + aR);
}

// Masked variant, EinfeldtSignalVelocities.hpp:99-142: the Roe weights are
// sqrt(rho_K) / (sqrt(rho_L) + sqrt(rho_R)) (two divisions) instead of a
// multiplication with the reciprocal of the sum.
template <class M>
FVS_DEV void einfeldt_signals_masked(const EqM& eq, double uL, double hL, double sqRhoL,
                                     double aL, double uR, double hR, double sqRhoR, double aR,
                                     double& s0, double& s1, Flag& fl) {
  const double sqRho = sqRhoL + sqRhoR;
  const double wL = M::div(sqRhoL, sqRho, fl);
  const double wR = M::div(sqRhoR, sqRho, fl);
  const double roeU = uL * wL + uR * wR;
  const double roeH = hL * wL + hR * wR;
  const double roeA2 = (eq.gamma - 1.0) * (roeH - 0.5 * roeU * roeU);
  const double roeA = M::sqrt(roeA2, fl);
  s0 = vmin(uL - aL, roeU - roeA);
  s1 = vmax(roeU + roeA, uR + aR);
}

// include/fub/flux_method/HllMethod.hpp:260-293 (MASKED: :295-335 on a face
// with positive face fraction)
template <class M, int D, int DIR, bool MASKED = false>
FVS_DEV void hll_einfeldt_flux(const EqM& eq, Cons<D>& nf, const Edge<D>& eL,
                               const Edge<D>& eR, Flag& fl) {
  const Complete<D>& left = eL.q;
  const Complete<D>& right = eR.q;
  M::check_window(eL.rc, fl);
  M::check_window(eR.rc, fl);
  const double uL = M::divz(left.mom[DIR], eL.rc, fl);
  const double uR = M::divz(right.mom[DIR], eR.rc, fl);
  const double hL = M::div(left.E + left.p, eL.rc, fl);
  const double hR = M::div(right.E + right.p, eR.rc, fl);
  double s0, s1;
  if constexpr (MASKED) {
    einfeldt_signals_masked<M>(eq, uL, hL, M::sqrt(left.rho, fl), left.a, uR, hR,
                               M::sqrt(right.rho, fl), right.a, s0, s1, fl);
  } else {
    einfeldt_signals<M>(eq, uL, hL, M::sqrt(left.rho, fl), left.a, uR, hR,
                        M::sqrt(right.rho, fl), right.a, s0, s1, fl);
  }
  Cons<D> fL, fR;
  euler_flux_v<D, DIR>(fL, left, uL);
  euler_flux_v<D, DIR>(fR, right, uR);
  const double sL = vmin(s0, 0.0);
  const double sR = vmax(s1, 0.0);
  const double sLsR = sL * sR;
  double ds = sR - sL;
  if constexpr (!MASKED) ds = ds > 0.0 ? ds : 1.0;
  const Rcp rcDs = M::rcp(ds);
  M::check_window(rcDs, fl);
  nf.rho = M::divz(sR * fL.rho - sL * fR.rho + sLsR * (right.rho - left.rho), rcDs, fl);
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double num = sR * fL.mom[d] - sL * fR.mom[d] + sLsR * (right.mom[d] - left.mom[d]);
    nf.mom[d] = d == DIR ? M::div(num, rcDs, fl) : M::divz(num, rcDs, fl);
  }
  nf.E = M::divz(sR * fL.E - sL * fR.E + sLsR * (right.E - left.E), rcDs, fl);
}

// include/fub/flux_method/HllMethod.hpp:337-347: max(0, |s0|, |s1|)
template <class M, int D, int DIR, bool MASKED = false>
FVS_DEV double hll_einfeldt_max_signal(const EqM& eq, const DtCell<D>& L, const DtCell<D>& R,
                                       Flag& fl) {
  double s0, s1;
  if constexpr (MASKED) {
    einfeldt_signals_masked<M>(eq, L.u[DIR], L.h, L.sq_rho, L.a, R.u[DIR], R.h, R.sq_rho, R.a,
                               s0, s1, fl);
  } else {
    einfeldt_signals<M>(eq, L.u[DIR], L.h, L.sq_rho, L.a, R.u[DIR], R.h, R.sq_rho, R.a, s0, s1,
                        fl);
  }
  double m = 0.0;
  m = vmax(m, fabs(s0));
  m = vmax(m, fabs(s1));
  return m;
}

// ---------------------------------------------------------------------------
// Exact Riemann solver, src/equations/perfect_gas/ExactRiemannSolver.hpp.
// Uses pow(); libm and CUDA pow differ in the last ulp, so every method built
// on it is compared with a tolerance, not bitwise.
struct NewtonStatus {
  int failed;
};

template <int D, int DIR> struct ExactRiemann {
  const EqConst& eq;
  FVS_DEV explicit ExactRiemann(const EqConst& e) : eq(e) {}

  struct Side {
    double rho, p, a, A, B;
  };
  FVS_DEV double f_side(const Side& s, double p, double gm1, double exponent) const {
    if (p > s.p) {
      return (p - s.p) * sqrt(s.A / (p + s.B));
    }
    return 2.0 * s.a / gm1 * (pow(p / s.p, exponent) - 1.0);
  }
  FVS_DEV double df_side(const Side& s, double p, double d_exponent) const {
    if (p > s.p) {
      return sqrt(s.A / (s.B + p)) * (1.0 - 0.5 * (p - s.p) / (s.B + p));
    }
    return pow(p / s.p, d_exponent) / (s.rho * s.a);
  }

  // :34-97 + include/fub/NewtonIteration.hpp:28-52
  template <typename SL, typename SR>
  FVS_DEV void middle_state(const SL& left, const SR& right, double& pM, double& uM,
                            int& failed) const {
    const double g = eq.gamma;
    const double gp1 = g + 1.0;
    const double gm1 = g - 1.0;
    Side L{left.rho, left.p, left.a, 2.0 / (gp1 * left.rho), gm1 / gp1 * left.p};
    Side R{right.rho, right.p, right.a, 2.0 / (gp1 * right.rho), gm1 / gp1 * right.p};
    const double exponent = 0.5 * gm1 / g;
    const double d_exponent = -0.5 * gp1 / g;
    const double uL = left.mom[DIR] / left.rho;
    const double uR = right.mom[DIR] / right.rho;
    const double du = uR - uL;
    double x = 0.5 * (L.p + R.p);
    double fx = f_side(R, x, gm1, exponent) + f_side(L, x, gm1, exponent) + du;
    int counter = 0;
    double err = fabs(fx);
    while (err > 1e-7 && counter < 5000) {
      const double dfx = df_side(R, x, d_exponent) + df_side(L, x, d_exponent);
      const double offset = fx / dfx;
      const double next = x - offset;
      x = ((next - next) != (next - next) || next < 0.0) ? 0.5 * x : next;
      fx = f_side(R, x, gm1, exponent) + f_side(L, x, gm1, exponent) + du;
      err = fabs(fx);
      counter += 1;
    }
    if (counter == 5000 && fx > 1e-7) failed = 1;
    pM = x;
    const double fL = f_side(L, pM, gm1, exponent);
    const double fR = f_side(R, pM, gm1, exponent);
    uM = 0.5 * (uL + uR) + 0.5 * (fR - fL);
  }

  // :105-148
  template <typename SL, typename SR>
  FVS_DEV void signals(const SL& left, const SR& right, double& s0, double& s1,
                       int& failed) const {
    double pM, uM;
    middle_state(left, right, pM, uM, failed);
    const double g = eq.gamma;
    const double gp1 = g + 1.0;
    const double gm1 = g - 1.0;
    const double uL = left.mom[DIR] / left.rho;
    const double pL_ratio = pM / left.p;
    if (pM > left.p) {
      s0 = uL - left.a * sqrt((gp1 * pL_ratio + gm1) / (2.0 * g));
    } else {
      s0 = uL - left.a;
    }
    const double uR = right.mom[DIR] / right.rho;
    const double pR_ratio = pM / right.p;
    if (pM > right.p) {
      s1 = uR + right.a * sqrt((gp1 * pR_ratio + gm1) / (2.0 * g));
    } else {
      s1 = uR + right.a;
    }
  }

  FVS_DEV void set_momentum(Complete<D>& s, double momentum, const Complete<D>& side) const {
    if constexpr (D > 1) {
#pragma unroll
      for (int d = 0; d < D; ++d) s.mom[d] = s.rho * (side.mom[d] / side.rho);
    }
    s.mom[DIR] = momentum;
  }
  FVS_DEV void from_prim(Complete<D>& s) const {
    const double rhoE_kin = 0.5 * squared_norm<D>(s.mom) / s.rho;
    const double rhoE_internal = s.p * eq.gamma_minus_1_inv;
    s.E = rhoE_internal + rhoE_kin;
    s.a = sqrt(eq.gamma * s.p / s.rho);
  }

  // :204-353, sampled at x/t = 0.  `state` must be zero-initialised by the
  // caller (GodunovMethod.hpp:129-145 keeps a zeroed member).
  FVS_DEV void solve(Complete<D>& state, const Complete<D>& left, const Complete<D>& right,
                     int& failed) const {
    if (!(0.0 < left.rho && 0.0 < right.rho)) return;
    double pM, uM;
    middle_state(left, right, pM, uM, failed);
    const double g = eq.gamma;
    const double gp1 = g + 1.0;
    const double gm1 = g - 1.0;
    const double g_ratio = gm1 / gp1;
    if (0.0 <= uM) {
      const double uL = left.mom[DIR] / left.rho;
      const double aL = left.a;
      const double pL = left.p;
      const double p_ratio = pM / pL;
      if (pM > pL) {
        const double shock_speed = uL - aL * sqrt((gp1 * p_ratio + gm1) / (2.0 * g));
        if (shock_speed < 0.0) {
          state.rho = left.rho * (p_ratio + g_ratio) / (g_ratio * p_ratio + 1.0);
          set_momentum(state, state.rho * uM, left);
          state.p = pM;
          from_prim(state);
        } else {
          state = left;
        }
      } else {
        const double aM = aL * pow(p_ratio, gm1 / (2.0 * g));
        const double speed_TL = uM - aM;
        const double speed_HL = uL - aL;
        if (0.0 < speed_HL) {
          state = left;
        } else if (speed_TL < 0.0) {
          state.rho = left.rho * pow(p_ratio, 1 / g);
          set_momentum(state, state.rho * uM, left);
          state.p = pM;
          from_prim(state);
        } else {
          state.rho = left.rho * pow(2.0 / gp1 + g_ratio * uL / aL, 2.0 / gm1);
          const double u = 2.0 / gp1 * (aL + 0.5 * gm1 * uL);
          set_momentum(state, state.rho * u, left);
          state.p = pL * pow(2.0 / gp1 + g_ratio * uL / aL, 2.0 * g / gm1);
          from_prim(state);
        }
      }
    } else {
      const double uR = right.mom[DIR] / right.rho;
      const double aR = right.a;
      const double pR = right.p;
      const double p_ratio = pM / pR;
      if (pM > pR) {
        const double shock_speed = uR + aR * sqrt((gp1 * p_ratio + gm1) / (2.0 * g));
        if (0.0 < shock_speed) {
          state.rho = right.rho * (p_ratio + g_ratio) / (g_ratio * p_ratio + 1.0);
          set_momentum(state, state.rho * uM, right);
          state.p = pM;
          from_prim(state);
        } else {
          state = right;
        }
      } else {
        const double aM = aR * pow(p_ratio, gm1 / (2.0 * g));
        const double speed_TR = uM + aM;
        const double speed_HR = uR + aR;
        if (speed_HR <= 0.0) {
          state = right;
        } else if (speed_TR >= 0.0) {
          state.rho = right.rho * pow(p_ratio, 1 / g);
          set_momentum(state, state.rho * uM, right);
          state.p = pM;
          from_prim(state);
        } else {
          state.rho = right.rho * pow(2.0 / gp1 - g_ratio * uR / aR, 2.0 / gm1);
          const double u = 2.0 / gp1 * (-aR + 0.5 * gm1 * uR);
          set_momentum(state, state.rho * u, right);
          state.p = pR * pow(2.0 / gp1 - g_ratio * uR / aR, 2.0 * g / gm1);
          from_prim(state);
        }
      }
    }
  }
};

// include/fub/flux_method/GodunovMethod.hpp:129-145 over the exact solver
template <int D, int DIR>
FVS_DEV void godunov_exact_flux(const EqConst& eq, Cons<D>& nf, const Complete<D>& left,
                                const Complete<D>& right, int& failed) {
  Complete<D> sol;
  sol.rho = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) sol.mom[d] = 0.0;
  sol.E = 0.0;
  sol.p = 0.0;
  sol.a = 0.0;
  ExactRiemann<D, DIR>(eq).solve(sol, left, right, failed);
  euler_flux_v<D, DIR>(nf, sol, sol.mom[DIR] / sol.rho);
}

template <int D, int DIR, typename SL, typename SR>
FVS_DEV double godunov_exact_max_signal(const EqConst& eq, const SL& left, const SR& right,
                                        int& failed) {
  double s0, s1;
  ExactRiemann<D, DIR>(eq).signals(left, right, s0, s1, failed);
  double m = 0.0;
  m = fmax(m, fabs(s0));
  m = fmax(m, fabs(s1));
  return m;
}

// ---------------------------------------------------------------------------
// MusclHancock2 (include/fub/flux_method/MusclHancockMethod2.hpp:35-239, array path):
// ComputeGradient of the cell from its two neighbours (Gradient.hpp), Reconstruct at the
// lower and the upper face (Reconstruct.hpp), the base method between the edge states of
// two cells.  One call evaluates the cell's gradient once and both of its edge states.
// CentralDifferenceGradient<Limiter>::ComputeGradient, Gradient.hpp:211-217:
//   mask = sL sR > 0;  R = sR / sL;  mask.select(0.5 limiter(R) (sL + sR), 0)
template <class M> FVS_DEV double mh2_central(double sL, double sR, int limiter, Flag& fl) {
  const bool mask = sL * sR > 0.0;
  double lim;
  if (limiter == kLimiter2None) {
    lim = 1.0;
  } else if (limiter == kLimiter2Upwind) {
    lim = 0.0;
  } else {
    const double R = M::div(sR, sL, fl, mask);
    if (limiter == kLimiter2MinMod) { // :41-50
      const double two = M::div(2.0, 1.0 + R, fl, mask);
      lim = vmin(two, R * two);
    } else { // VanLeerLimiter :52-59
      lim = M::div(4.0 * R, (1.0 + R) * (1.0 + R), fl, mask);
    }
  }
  const double g = 0.5 * lim * (sL + sR);
  return mask ? g : 0.0;
}

// (b - a) / dx for a difference that may be zero; rdx from M::rcp(dx), window-checked
template <class M> FVS_DEV double mh2_slope(double a, double b, const Rcp& rdx, Flag& fl) {
  return M::divz(b - a, rdx, fl);
}

// PerfectGas::CompleteFromPrim (array), src/equations/PerfectGas.cpp:160-174
template <class M, int D>
FVS_DEV void mh2_complete_from_prim(const EqM& eq, Edge<D>& out, double rho, const double (&u)[D],
                                    double p, Flag& fl) {
  out.q.rho = rho;
#pragma unroll
  for (int d = 0; d < D; ++d) out.q.mom[d] = rho * u[d];
  out.q.p = p;
  out.rc = M::rcp(rho);
  M::check_window(out.rc, fl);
  const double e_kin = M::divz(0.5 * squared_norm<D>(out.q.mom), out.rc, fl);
  out.q.E = e_kin + p * eq.gm1_inv.b;
  out.q.a = M::sqrt(M::div(eq.gamma * p, out.rc, fl), fl);
}

// ConservativeGradient + ConservativeReconstruction (Gradient.hpp:247-270,
// Reconstruct.hpp:188-214)
template <class M, int D, int DIR, typename S0, typename S1, typename S2>
FVS_DEV void mh2_cons_predict(const EqM& eq, const S0& qm, const S1& q0, const S2& qp, double dx,
                              int limiter, double lambda_half, Edge<D>& recL, Edge<D>& recR,
                              Flag& fl, bool slope_ok = true) {
  const Rcp rdx = M::rcp(dx);
  M::check_window(rdx, fl);
  const double half_dx = 0.5 * dx;
  Complete<D> qL, qR;
  auto split = [&](double um, double u, double up, double& l, double& r) {
    double g = mh2_central<M>(mh2_slope<M>(um, u, rdx, fl), mh2_slope<M>(u, up, rdx, fl), limiter, fl);
    g = slope_ok ? g : 0.0; // MusclHancockMethod2.hpp:200-213
    const double du = half_dx * g;
    l = u - du;
    r = u + du;
  };
  split(qm.rho, q0.rho, qp.rho, qL.rho, qR.rho);
#pragma unroll
  for (int d = 0; d < D; ++d) split(qm.mom[d], q0.mom[d], qp.mom[d], qL.mom[d], qR.mom[d]);
  split(qm.E, q0.E, qp.E, qL.E, qR.E);
  Rcp rcL = M::rcp(qL.rho);
  Rcp rcR = M::rcp(qR.rho);
  complete_edge<M, D, false, false>(eq, qL, rcL, fl);
  complete_edge<M, D, false, false>(eq, qR, rcR, fl);
  Cons<D> fL, fR;
  euler_flux_v<D, DIR>(fL, qL, M::divz(qL.mom[DIR], rcL, fl));
  euler_flux_v<D, DIR>(fR, qR, M::divz(qR.mom[DIR], rcR, fl));
  {
    const double c = lambda_half * (fL.rho - fR.rho);
    recL.q.rho = qL.rho + c;
    recR.q.rho = qR.rho + c;
  }
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double c = lambda_half * (fL.mom[d] - fR.mom[d]);
    recL.q.mom[d] = qL.mom[d] + c;
    recR.q.mom[d] = qR.mom[d] + c;
  }
  {
    const double c = lambda_half * (fL.E - fR.E);
    recL.q.E = qL.E + c;
    recR.q.E = qR.E + c;
  }
  recL.rc = M::rcp(recL.q.rho);
  recR.rc = M::rcp(recR.q.rho);
  complete_edge<M, D, false>(eq, recL.q, recL.rc, fl);
  complete_edge<M, D, false>(eq, recR.q, recR.rc, fl);
}

// PrimitiveGradient + PrimitiveReconstruction (Gradient.hpp:287-302, Reconstruct.hpp:343-388)
template <class M, int D, int DIR, typename S0, typename S1, typename S2>
FVS_DEV void mh2_prim_predict(const EqM& eq, const S0& qm, const S1& q0, const S2& qp, double dx,
                              int limiter, double lambda, Edge<D>& recL, Edge<D>& recR, Flag& fl,
                              bool slope_ok = true) {
  PrimCell<D> m, c, p;
  prim_cell<M, D>(m, qm, fl);
  prim_cell<M, D>(c, q0, fl);
  prim_cell<M, D>(p, qp, fl);
  const Rcp rdx = M::rcp(dx);
  M::check_window(rdx, fl);
  auto grad = [&](double a, double b, double cc) {
    const double g =
        mh2_central<M>(mh2_slope<M>(a, b, rdx, fl), mh2_slope<M>(b, cc, rdx, fl), limiter, fl);
    return slope_ok ? g : 0.0;
  };
  const double drho = grad(m.rho, c.rho, p.rho);
  const double dp = grad(m.p, c.p, p.p);
  double du[D];
#pragma unroll
  for (int i = 0; i < D; ++i) du[i] = grad(m.u[i], c.u[i], p.u[i]);
  const double u = c.u[DIR];
  const double rho = q0.rho;
  const double a2 = q0.a * q0.a;
  const double rho_a2 = rho * a2;
  const Rcp rcc = M::rcp(rho);
  M::check_window(rcc, fl);
  const double dp_over_rho = M::divz(dp, rcc, fl);
  auto edge = [&](double sign, Edge<D>& out) {
    const double lam = sign * lambda;
    const double dx_half = sign * (0.5 * dx);
    const double rec_rho = c.rho + dx_half * (drho - lam * (u * drho + rho * du[DIR]));
    const double rec_p = c.p + dx_half * (dp - lam * (rho_a2 * du[DIR] + u * dp));
    double rec_u[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      if (i == DIR) {
        rec_u[i] = c.u[i] + dx_half * (du[i] - lam * (u * du[i] + dp_over_rho));
      } else {
        rec_u[i] = c.u[i] + dx_half * (du[i] - lam * u * du[i]);
      }
    }
    mh2_complete_from_prim<M, D>(eq, out, rec_rho, rec_u, rec_p, fl);
  };
  edge(-1.0, recL);
  edge(+1.0, recR);
}

// CharacteristicsGradient + CharacteristicsReconstruction (Gradient.hpp:354-368 with
// ComputeAmplitudes :322-352, Reconstruct.hpp:458-516)
template <class M, int D, int DIR, typename S0, typename S1, typename S2>
FVS_DEV void mh2_char_predict(const EqM& eq, const S0& qm, const S1& q0, const S2& qp, double dx,
                              int limiter, double lambda, Edge<D>& recL, Edge<D>& recR, Flag& fl,
                              bool slope_ok = true) {
  PrimCell<D> m, c, p;
  prim_cell<M, D>(m, qm, fl);
  prim_cell<M, D>(c, q0, fl);
  prim_cell<M, D>(p, qp, fl);
  const Rcp rdx = M::rcp(dx);
  M::check_window(rdx, fl);
  const double cs = q0.a;
  const double ooc2 = M::div(1.0, cs * cs, fl);
  const double rhoc = q0.rho * cs;
  // amplitudes: minus, plus, zero[0] (entropy), zero[1..] (transverse velocities)
  struct Amp {
    double minus, plus, zero[D];
  };
  auto amplitudes = [&](const PrimCell<D>& l, const PrimCell<D>& r, Amp& a) {
    const double dpx = mh2_slope<M>(l.p, r.p, rdx, fl);
    const double dux = mh2_slope<M>(l.u[DIR], r.u[DIR], rdx, fl);
    const double drx = mh2_slope<M>(l.rho, r.rho, rdx, fl);
    a.minus = dpx - (1.0 * rhoc) * dux;
    a.plus = dpx + (1.0 * rhoc) * dux;
    a.zero[0] = drx - ooc2 * dpx;
#pragma unroll
    for (int i = 1; i < D; ++i) {
      const int iy = (DIR + i) % D;
      a.zero[i] = mh2_slope<M>(l.u[iy], r.u[iy], rdx, fl);
    }
  };
  Amp aL, aR, g;
  amplitudes(m, c, aL);
  amplitudes(c, p, aR);
  auto lim = [&](double sl, double sr) {
    const double v = mh2_central<M>(sl, sr, limiter, fl);
    return slope_ok ? v : 0.0;
  };
  g.minus = lim(aL.minus, aR.minus);
  g.plus = lim(aL.plus, aR.plus);
#pragma unroll
  for (int i = 0; i < D; ++i) g.zero[i] = lim(aL.zero[i], aR.zero[i]);
  const double u = c.u[DIR];
  const double c2 = cs * cs;
  const Rcp rc2 = M::rcp(c2);
  M::check_window(rc2, fl);
  const Rcp rrhoc = M::rcp(rhoc);
  M::check_window(rrhoc, fl);
  auto edge = [&](double sign, Edge<D>& out) {
    const double lam = sign * lambda;
    const double dx_half = sign * (0.5 * dx);
    const double kminus = dx_half * g.minus * (1.0 - lam * (u - cs));
    const double kplus = dx_half * g.plus * (1.0 - lam * (u + cs));
    double kzero[D];
#pragma unroll
    for (int i = 0; i < D; ++i) kzero[i] = dx_half * g.zero[i] * (1.0 - lam * u);
    const double dwp = 0.5 * (kminus + kplus);
    const double dwrho = M::divz(dwp, rc2, fl) + kzero[0];
    double dwu[D];
    dwu[DIR] = M::divz(0.5 * (kplus - kminus), rrhoc, fl);
#pragma unroll
    for (int i = 1; i < D; ++i) dwu[(DIR + i) % D] = kzero[i];
    double rec_u[D];
#pragma unroll
    for (int i = 0; i < D; ++i) rec_u[i] = c.u[i] + dwu[i];
    mh2_complete_from_prim<M, D>(eq, out, c.rho + dwrho, rec_u, c.p + dwp, fl);
  };
  edge(-1.0, recL);
  edge(+1.0, recR);
}

template <class M, int D, int DIR, int RECON, typename S0, typename S1, typename S2>
FVS_DEV void mh2_predict(const EqM& eq, const EqConst& c, const S0& qm, const S1& q0, const S2& qp,
                         double lambda_half, Edge<D>& recL, Edge<D>& recR, Flag& fl,
                         bool slope_ok = true) {
  if constexpr (RECON == kReconMh2Conservative) {
    mh2_cons_predict<M, D, DIR>(eq, qm, q0, qp, c.dx, c.limiter2, lambda_half, recL, recR, fl, slope_ok);
  } else if constexpr (RECON == kReconMh2Primitive) {
    mh2_prim_predict<M, D, DIR>(eq, qm, q0, qp, c.dx, c.limiter2, 2.0 * lambda_half, recL, recR, fl,
                                slope_ok);
  } else {
    mh2_char_predict<M, D, DIR>(eq, qm, q0, qp, c.dx, c.limiter2, 2.0 * lambda_half, recL, recR, fl,
                                slope_ok);
  }
}

// ---------------------------------------------------------------------------
// Dispatch on the base method.
struct Eqs {
  EqConst c;
  EqM m;
  FVS_DEV explicit Eqs(const EqConst& e) : c(e), m(e) {}
};

template <class M, int D, int DIR, int BASE, bool MASKED = false>
FVS_DEV void base_flux(const Eqs& eq, Cons<D>& f, const Edge<D>& left, const Edge<D>& right,
                       Flag& fl, int& failed) {
  if constexpr (BASE == kBaseHllemLarrouturou) {
    hllem_flux<M, D, DIR, true>(eq.m, f, left, right, fl);
  } else if constexpr (BASE == kBaseHllem) {
    hllem_flux<M, D, DIR, false>(eq.m, f, left, right, fl);
  } else if constexpr (BASE == kBaseHllEinfeldt) {
    hll_einfeldt_flux<M, D, DIR, MASKED>(eq.m, f, left, right, fl);
  } else {
    godunov_exact_flux<D, DIR>(eq.c, f, left.q, right.q, failed);
  }
}

// Per-cell data of the stable-dt evaluation for a base method: the signal
// speeds of a face need these from its two STORED neighbour states.
template <int D, int BASE> struct DtData {
  DtCell<D> c;
};
template <int D> struct DtData<D, kBaseGodunovExact> {
  Complete<D> c; // the exact solver needs the whole state
};

template <class M, int D, int BASE>
FVS_DEV void dt_prepare(DtData<D, BASE>& out, const Complete<D>& q, Flag& fl) {
  if constexpr (BASE == kBaseGodunovExact) {
    out.c = q;
  } else {
    dt_cell<M, D>(out.c, q, fl);
  }
}

// max signal speed of the face between two stored states (the stable dt of a
// face is dx / value; min over faces == dx / max over faces because IEEE
// division is monotone in its divisor).
template <class M, int D, int DIR, int BASE, bool MASKED = false>
FVS_DEV double base_max_signal(const Eqs& eq, const DtData<D, BASE>& left,
                               const DtData<D, BASE>& right, Flag& fl, int& failed) {
  if constexpr (BASE == kBaseHllemLarrouturou || BASE == kBaseHllem) {
    return hllem_max_signal<M, D, DIR>(eq.m, left.c, right.c, fl);
  } else if constexpr (BASE == kBaseHllEinfeldt) {
    return hll_einfeldt_max_signal<M, D, DIR, MASKED>(eq.m, left.c, right.c, fl);
  } else {
    return godunov_exact_max_signal<D, DIR>(eq.c, left.c, right.c, failed);
  }
}

// Runs `f(policy, flag)` with the branch-free arithmetic and, if any
// operation left the fast-path domain, again with the plain operators.
#ifndef FVS_FASTMATH
#define FVS_FASTMATH 1
#endif
#ifdef FVS_EXPERIMENT_NO_FALLBACK
__device__ int g_experiment_sink;
#endif
template <typename F> FVS_DEV void with_fallback(F&& f) {
#if FVS_FASTMATH
  Flag fl;
#if FVS_FASTMATH == 2
  f(ToleranceMath{}, fl);
#else
  f(FastMath{}, fl);
#endif
#ifdef FVS_EXPERIMENT_NO_FALLBACK
  // measurement only (tools/ab.sh): the tests are evaluated, the recompute is not compiled in
  if (fl.any()) atomicAdd(&g_experiment_sink, 1);
#else
  if (fl.any()) {
    Flag unused;
    f(SafeMath{}, unused);
  }
#endif
#else
  Flag unused;
  f(SafeMath{}, unused);
#endif
}

// Fast path only, validity accumulated into the caller's flag (the fused y/z sweep tests it
// once per line segment and redoes the segment with the checked stages if it is not clean).
#if FVS_FASTMATH == 2
using FastPolicy = ToleranceMath;
#elif FVS_FASTMATH
using FastPolicy = FastMath;
#else
using FastPolicy = SafeMath;
#endif
template <typename F> FVS_DEV void fast_only(F&& f, Flag& fl) { f(FastPolicy{}, fl); }

} // namespace fvs
