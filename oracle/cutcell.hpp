// TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.
//
// oracle/cutcell.hpp: CPU restatement of the reference's embedded-boundary
// (cut-cell) path: CutCellData + FillCutCellData + IsCutCell, the Klein-2009
// KbnCutCellMethod (reference states, boundary fluxes, masked regular fluxes,
// stabilised fluxes, masked stable dt), the cut-cell conservative update, the
// masked reconstruction, and the patch-level tagging methods.
// Citations are relative to /root/reference.
//
// PARITY UNPINNED: the reference has no test with numeric assertions for any
// of these functions, and the geometry arrays (alpha, beta, centroids,
// normals) come from AMReX EB2, which is absent here.  Oracle and CUDA path
// are fed the same arrays; this file follows the reference line by line.
// Eigen fixed-size products / dot products use the redux_novec_unroller
// order (x0 + (x1 + x2) for three terms, see numerics.hpp).
#pragma once
#include "patch.hpp"

#include <algorithm>
#include <cstdint>

namespace fvso {

enum FabType { kFabCovered = -1, kFabRegular = 0, kFabSingleValued = 1 };
enum RiemannKind { kRiemannHllEinfeldt = 0, kRiemannExact = 1 };

// include/fub/CutCellData.hpp:48-75 for one fab.  Cell-centred arrays live
// on the scratch box (incl. ghosts), beta[d] on surroundingNodes(box, d); the
// derived fractions have two components (0 absolute, 1 relative) like the
// MultiFabs of src/AMReX/cutcell/PatchHierarchy.cpp:380-431.
struct CutCellData {
  int fab_type{kFabRegular};
  Fab alpha;                      // volume_fractions
  Fab beta[3];                    // face_fractions
  Fab us[3], sl[3], sr[3], ds[3]; // unshielded, shielded left/right, doubly shielded
  Fab bcent;                      // boundary_centeroids (units of dx, rel. to cell centre)
  Fab bnorm;                      // boundary_normals as AMReX stores them
  void define(const Box& cells, int rank) {
    alpha.define(cells, 1);
    bcent.define(cells, rank);
    bnorm.define(cells, rank);
    for (int d = 0; d < rank; ++d) {
      const Box fb = surrounding_nodes(cells, d);
      beta[d].define(fb, 1);
      us[d].define(fb, 2);
      sl[d].define(fb, 2);
      sr[d].define(fb, 2);
      ds[d].define(fb, 2);
    }
  }
};

// src/AMReX/cutcell/FillCutCellData.cpp:28-118 on every face that has both
// neighbours in direction d.
inline void FillCutCellData(CutCellData& g, int rank) {
  for (int d = 0; d < rank; ++d) {
    Box box = g.beta[d].box;
    box.lo[d] += 1;
    box.hi[d] -= 1;
    if (!box.ok()) continue;
    int e[3] = {0, 0, 0};
    e[d] = 1;
    for (int k = box.lo[2]; k <= box.hi[2]; ++k)
      for (int j = box.lo[1]; j <= box.hi[1]; ++j)
        for (int i = box.lo[0]; i <= box.hi[0]; ++i) {
          const double betaL = g.beta[d](i - e[0], j - e[1], k - e[2], 0);
          double betaM = g.beta[d](i, j, k, 0);
          const double betaR = g.beta[d](i + e[0], j + e[1], k + e[2], 0);
          const double dBetaL = std::max(0.0, betaM - betaL);
          const double dBetaR = std::max(0.0, betaM - betaR);
          const double unshielded = std::max(0.0, betaM - std::max(dBetaL, dBetaR));
          const double doubly_shielded = std::min(dBetaL, dBetaR);
          const double shielded_left = dBetaL - doubly_shielded;
          const double shielded_right = dBetaR - doubly_shielded;
          betaM = betaM > 0 ? betaM : 1.0;
          g.us[d](i, j, k, 0) = unshielded;
          g.us[d](i, j, k, 1) = unshielded / betaM;
          g.ds[d](i, j, k, 0) = doubly_shielded;
          g.ds[d](i, j, k, 1) = doubly_shielded / betaM;
          g.sl[d](i, j, k, 0) = shielded_left;
          g.sl[d](i, j, k, 1) = shielded_left / betaM;
          g.sr[d](i, j, k, 0) = shielded_right;
          g.sr[d](i, j, k, 1) = shielded_right / betaM;
        }
  }
}

// src/CutCellData.cpp:225-256
inline bool IsCutCell(const CutCellData& g, int rank, int i, int j, int k) {
  const double a = g.alpha(i, j, k, 0);
  if (a <= 0.0) return false;
  if (a < 1.0) return true;
  for (int d = 0; d < rank; ++d) {
    int e[3] = {0, 0, 0};
    e[d] = 1;
    if (g.beta[d](i, j, k, 0) < 1.0 || g.beta[d](i + e[0], j + e[1], k + e[2], 0) < 1.0) {
      return true;
    }
  }
  return false;
}

// AMReX EBCellFlagFab::getType(box) modelled on alpha / beta: a cell is
// covered if alpha == 0, regular if it is not a cut cell, else single valued.
inline int CellFlagType(const CutCellData& g, int rank, const Box& cells) {
  long ncov = 0, nreg = 0, n = 0;
  for (int k = cells.lo[2]; k <= cells.hi[2]; ++k)
    for (int j = cells.lo[1]; j <= cells.hi[1]; ++j)
      for (int i = cells.lo[0]; i <= cells.hi[0]; ++i) {
        ++n;
        if (g.alpha(i, j, k, 0) <= 0.0) {
          ++ncov;
        } else if (!IsCutCell(g, rank, i, j, k)) {
          ++nreg;
        }
      }
  if (nreg == n) return kFabRegular;
  if (ncov == n) return kFabCovered;
  return kFabSingleValued;
}

// ---------------------------------------------------------------------------
// include/fub/ext/Eigen.hpp:104-130 and src/equations/PerfectGas.cpp:230-300
template <int D> struct Mat {
  double m[D][D];
};
template <int D> inline double dotD(const double (&a)[D], const double (&b)[D]) {
  if constexpr (D == 1) {
    return a[0] * b[0];
  } else if constexpr (D == 2) {
    return a[0] * b[0] + a[1] * b[1];
  } else {
    return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]);
  }
}
template <int D> inline Mat<D> MakeRotation(const double (&a)[D], const double (&b)[D]) {
  Mat<D> R{};
  bool opposite = true;
  for (int d = 0; d < D; ++d) opposite = opposite && (a[d] == -b[d]);
  if (opposite) {
    for (int r = 0; r < D; ++r)
      for (int c = 0; c < D; ++c) R.m[r][c] = (r == c) ? -1.0 : -0.0;
    return R;
  }
  if constexpr (D == 1) {
    R.m[0][0] = 1.0;
  } else if constexpr (D == 2) {
    const double s = a[0] * b[1] - a[1] * b[0];
    const double c = dotD<2>(a, b);
    R.m[0][0] = c;
    R.m[0][1] = -s;
    R.m[1][0] = s;
    R.m[1][1] = c;
  } else {
    const double v[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2],
                         a[0] * b[1] - a[1] * b[0]};
    const double c = dotD<3>(a, b);
    const double skew[3][3] = {{0, -v[2], +v[1]}, {+v[2], 0, -v[0]}, {-v[1], +v[0], 0}};
    const double one_plus_c = 1 + c;
    for (int r = 0; r < 3; ++r)
      for (int cc = 0; cc < 3; ++cc) {
        const double prod[3] = {skew[r][0] * skew[0][cc], skew[r][1] * skew[1][cc],
                                skew[r][2] * skew[2][cc]};
        const double sq = prod[0] + (prod[1] + prod[2]);
        const double I = (r == cc) ? 1.0 : 0.0;
        R.m[r][cc] = (I + skew[r][cc]) + sq / one_plus_c;
      }
  }
  return R;
}
template <int D> inline void RotateMomentum(double (&out)[D], const Mat<D>& R, const double (&m)[D]) {
  double tmp[D];
  for (int r = 0; r < D; ++r) {
    if constexpr (D == 1) {
      tmp[r] = R.m[r][0] * m[0];
    } else if constexpr (D == 2) {
      tmp[r] = R.m[r][0] * m[0] + R.m[r][1] * m[1];
    } else {
      tmp[r] = R.m[r][0] * m[0] + (R.m[r][1] * m[1] + R.m[r][2] * m[2]);
    }
  }
  for (int r = 0; r < D; ++r) out[r] = tmp[r];
}
template <int D>
inline void Rotate(Complete<double, D>& out, const Complete<double, D>& s, const Mat<D>& R) {
  Complete<double, D> r = s;
  RotateMomentum<D>(r.mom, R, s.mom);
  out = r;
}
// Reflect: momentum - 2 * ((momentum . normal) * normal)
template <int D>
inline void Reflect(Complete<double, D>& out, const Complete<double, D>& s,
                    const double (&normal)[D]) {
  Complete<double, D> r = s;
  const double mn = dotD<D>(s.mom, normal);
  for (int d = 0; d < D; ++d) r.mom[d] = s.mom[d] - 2 * (mn * normal[d]);
  out = r;
}

// ---------------------------------------------------------------------------
// Scalar HLL Riemann solver with Einfeldt signals:
// include/fub/flux_method/HllMethod.hpp:162-184,
// src/equations/perfect_gas/EinfeldtSignalVelocities.hpp:26-62
template <int D>
inline void HllSolveRiemannProblem(const PerfectGasEq& eq, Complete<double, D>& solution,
                                   const Complete<double, D>& left,
                                   const Complete<double, D>& right, int dir) {
  const double rhoL = left.rho;
  const double rhoR = right.rho;
  const double rhoUL = left.mom[dir];
  const double rhoUR = right.mom[dir];
  const double aL = left.a;
  const double aR = right.a;
  const double sqRhoL = std::sqrt(rhoL);
  const double sqRhoR = std::sqrt(rhoR);
  const double oosqRhoSum = 1.0 / (sqRhoL + sqRhoR);
  const double uL = rhoUL / rhoL;
  const double uR = rhoUR / rhoR;
  const double roeU = (uL * sqRhoL + uR * sqRhoR) * oosqRhoSum;
  const double hL = (left.E + left.p) / rhoL;
  const double hR = (right.E + right.p) / rhoR;
  const double roeH = (hL * sqRhoL + hR * sqRhoR) * oosqRhoSum;
  const double roeA2 = (eq.gamma - 1.0) * (roeH - 0.5 * roeU * roeU);
  const double roeA = std::sqrt(roeA2);
  const double sL = std::min(uL - aL, roeU - roeA);
  const double sR = std::max(roeU + roeA, uR + aR);
  Cons<double, D> fL, fR;
  Flux<double, D>(eq, fL, left, dir);
  Flux<double, D>(eq, fR, right, dir);
  const double ds = sL != sR ? sR - sL : 1.0;
  if (0.0 < sL) {
    solution = left;
  } else if (sR < 0.0) {
    solution = right;
  } else {
    Cons<double, D> c;
    c.rho = (sR * right.rho - sL * left.rho + fL.rho - fR.rho) / ds;
    for (int d = 0; d < D; ++d) {
      c.mom[d] = (sR * right.mom[d] - sL * left.mom[d] + fL.mom[d] - fR.mom[d]) / ds;
    }
    c.E = (sR * right.E - sL * left.E + fL.E - fR.E) / ds;
    CompleteFromCons<double, D>(eq, solution, c);
  }
}

template <int D> inline Complete<double, D> load_state(const Fab& f, int i, int j, int k) {
  Complete<double, D> q;
  q.rho = f(i, j, k, 0);
  for (int d = 0; d < D; ++d) q.mom[d] = f(i, j, k, 1 + d);
  q.E = f(i, j, k, 1 + D);
  q.p = f(i, j, k, 2 + D);
  q.a = f(i, j, k, 3 + D);
  return q;
}
template <int D> inline void store_state(Fab& f, int i, int j, int k, const Complete<double, D>& q) {
  f(i, j, k, 0) = q.rho;
  for (int d = 0; d < D; ++d) f(i, j, k, 1 + d) = q.mom[d];
  f(i, j, k, 1 + D) = q.E;
  f(i, j, k, 2 + D) = q.p;
  f(i, j, k, 3 + D) = q.a;
}

// ---------------------------------------------------------------------------
// KbnCutCellMethod, include/fub/cutcell_method/KbnStabilisation.hpp
template <int D> struct Kbn {
  PerfectGasEq eq;
  int riemann{kRiemannHllEinfeldt};
  ExactRiemannSolver<D> exact;
  Kbn(const PerfectGasEq& e, int rk) : eq(e), riemann(rk), exact(e) {}

  void Solve(Complete<double, D>& sol, const Complete<double, D>& l,
             const Complete<double, D>& r) const {
    if (riemann == kRiemannExact) {
      exact.SolveRiemannProblem(sol, l, r, 0);
    } else {
      HllSolveRiemannProblem<D>(eq, sol, l, r, 0);
    }
  }
  static void normal_of(double (&n)[D], const CutCellData& g, int i, int j, int k) {
    // src/CutCellData.cpp:90-105: minus the AMReX normal
    for (int d = 0; d < D; ++d) n[d] = -g.bnorm(i, j, k, d);
  }
  // Wall Riemann problem in the frame where the boundary normal is e_x:
  // rotate, reflect, solve.  Returns the solution in the rotated frame.
  void WallProblem(Complete<double, D>& solution, Complete<double, D>& state,
                   const double (&normal)[D]) const {
    double unit[D] = {};
    unit[0] = 1.0;
    Rotate<D>(state, state, MakeRotation<D>(normal, unit));
    Complete<double, D> reflected;
    Reflect<D>(reflected, state, unit);
    Solve(solution, reflected, state);
  }

  // :188-210
  void PreAdvanceHierarchy(Fab& references, const Fab& states, const CutCellData& g,
                           const Box& box) const {
    for (int k = box.lo[2]; k <= box.hi[2]; ++k)
      for (int j = box.lo[1]; j <= box.hi[1]; ++j)
        for (int i = box.lo[0]; i <= box.hi[0]; ++i) {
          if (!IsCutCell(g, D, i, j, k)) continue;
          Complete<double, D> state = load_state<D>(states, i, j, k);
          if (state.rho > 0.0) {
            double normal[D];
            normal_of(normal, g, i, j, k);
            Complete<double, D> solution = zero_state<D>();
            WallProblem(solution, state, normal);
            double unit[D] = {};
            unit[0] = 1.0;
            Rotate<D>(solution, solution, MakeRotation<D>(unit, normal));
            store_state<D>(references, i, j, k, solution);
          }
        }
  }

  // :212-273 with mask == 0 (reference_masks are all zero unless a boundary
  // condition outside this path sets them)
  void ComputeBoundaryFluxes(Fab& boundary_fluxes, const Fab& states, const Fab& references,
                             const CutCellData& g, const Box& box, int dir) const {
    for (int k = box.lo[2]; k <= box.hi[2]; ++k)
      for (int j = box.lo[1]; j <= box.hi[1]; ++j)
        for (int i = box.lo[0]; i <= box.hi[0]; ++i) {
          if (!IsCutCell(g, D, i, j, k)) continue;
          Complete<double, D> state = load_state<D>(states, i, j, k);
          const Complete<double, D> ref = load_state<D>(references, i, j, k);
          double normal[D];
          normal_of(normal, g, i, j, k);
          Cons<double, D> flux;
          Flux<double, D>(eq, flux, ref, dir);
          Complete<double, D> solution = zero_state<D>();
          WallProblem(solution, state, normal);
          flux.mom[dir] += solution.p - ref.p;
          boundary_fluxes(i, j, k, 0) = flux.rho;
          for (int d = 0; d < D; ++d) boundary_fluxes(i, j, k, 1 + d) = flux.mom[d];
          boundary_fluxes(i, j, k, 1 + D) = flux.E;
        }
  }

  // :277-345.  `cells` = the cell box handed in by the adapter.
  template <typename Method>
  double ComputeStableDt(const Method& m, const Fab& states, const CutCellData& g,
                         const Box& cells, double dx, int dir) const {
    constexpr int W = Method::kWidth;
    constexpr int S = 2 * W;
    double min_dt = std::numeric_limits<double>::infinity();
    Box faces = cells; // face f lies between cells f-1 and f
    faces.lo[dir] += W;
    faces.hi[dir] -= (W - 1);
    if (!faces.ok()) return min_dt;
    Complete<A8, D> stencil[S];
    for (int s = 0; s < S; ++s) init_benign<D>(stencil[s]);
    A8 alphas[S];
    const long sstride = states.stride(dir);
    const long astride = g.alpha.stride(dir);
    for (int k = faces.lo[2]; k <= faces.hi[2]; ++k)
      for (int j = faces.lo[1]; j <= faces.hi[1]; ++j) {
        int i = faces.lo[0];
        int n = faces.hi[0] - i + 1;
        do {
          const int nn = std::min(n, kChunk);
          A8 betas(0.0);
          for (int l = 0; l < nn; ++l) betas.v[l] = g.beta[dir](i + l, j, k, 0);
          const long c0 = states.index(i, j, k) - W * sstride;
          const long a0 = g.alpha.index(i, j, k) - W * astride;
          for (int s = 0; s < S; ++s) {
            load_complete<D>(stencil[s], states, c0 + s * sstride, nn);
            alphas[s] = A8(0.0);
            for (int l = 0; l < nn; ++l) alphas[s].v[l] = g.alpha.data[size_t(a0 + s * astride + l)];
          }
          const A8 dts = m.ComputeStableDt(stencil, betas, alphas, dx, dir);
          min_dt = std::min(min_dt, min_coeff(dts));
          i += nn;
          n -= nn;
        } while (n > 0);
      }
    return min_dt;
  }

  // :347-429
  template <typename Method>
  void ComputeRegularFluxes(const Method& m, Fab& fluxes, const Box& faces, const Fab& states,
                            const CutCellData& g, double dt, double dx, int dir) const {
    constexpr int W = Method::kWidth;
    constexpr int S = 2 * W;
    if (!faces.ok()) return;
    Complete<A8, D> stencil[S];
    for (int s = 0; s < S; ++s) init_benign<D>(stencil[s]);
    A8 alphas[S];
    Cons<A8, D> flux;
    const long sstride = states.stride(dir);
    const long astride = g.alpha.stride(dir);
    for (int k = faces.lo[2]; k <= faces.hi[2]; ++k)
      for (int j = faces.lo[1]; j <= faces.hi[1]; ++j) {
        int i = faces.lo[0];
        int n = faces.hi[0] - i + 1;
        while (n > 0) {
          const int nn = std::min(n, kChunk);
          A8 betas(0.0);
          for (int l = 0; l < nn; ++l) betas.v[l] = g.beta[dir](i + l, j, k, 0);
          const long c0 = states.index(i, j, k) - W * sstride;
          const long a0 = g.alpha.index(i, j, k) - W * astride;
          for (int s = 0; s < S; ++s) {
            load_complete<D>(stencil[s], states, c0 + s * sstride, nn);
            alphas[s] = A8(0.0);
            for (int l = 0; l < nn; ++l) alphas[s].v[l] = g.alpha.data[size_t(a0 + s * astride + l)];
          }
          m.ComputeNumericFlux(flux, betas, stencil, alphas, dt, dx, dir);
          store_cons<D>(fluxes, fluxes.index(i, j, k), flux, nn);
          i += nn;
          n -= nn;
        }
      }
  }

  // :431-450 + src/cutcell_method/KbnStabilisation.cpp:55-159
  void ComputeCutCellFluxes(Fab& stabilised, Fab& shielded_left, Fab& shielded_right,
                            const Fab& boundary, const Fab& regular, const Box& faces,
                            const CutCellData& g, int dir) const {
    if (!faces.ok()) return;
    int e[3] = {0, 0, 0};
    e[dir] = 1;
    for (int c = 0; c < D + 2; ++c)
      for (int k = faces.lo[2]; k <= faces.hi[2]; ++k)
        for (int j = faces.lo[1]; j <= faces.hi[1]; ++j)
          for (int i = faces.lo[0]; i <= faces.hi[0]; ++i) {
            const double centerL = g.bcent(i - e[0], j - e[1], k - e[2], dir);
            const double centerR = g.bcent(i, j, k, dir);
            const double dL = std::clamp(0.5 - centerL, 0.0, 1.0);
            const double dR = std::clamp(0.5 + centerR, 0.0, 1.0);
            const double f = regular(i, j, k, c);
            const double fsL = dL * f + (1.0 - dL) * boundary(i - e[0], j - e[1], k - e[2], c);
            const double fsR = dR * f + (1.0 - dR) * boundary(i, j, k, c);
            const double betaL = g.sl[dir](i, j, k, 1);
            const double betaR = g.sr[dir](i, j, k, 1);
            const double betaUS = g.us[dir](i, j, k, 1);
            shielded_left(i, j, k, c) = fsL;
            shielded_right(i, j, k, c) = fsR;
            stabilised(i, j, k, c) = betaUS * f + betaL * fsL + betaR * fsR;
          }
  }
};

// src/AMReX/cutcell/TimeIntegrator.cpp:65-221 (scalar semantics: alpha == 0
// stores 0; the Vc lanes give the same values wherever alpha > 0).
inline void CutCellUpdateConservatively(Fab& scratch, const Fab& unshielded, const Fab& stabilized,
                                        const Fab& shielded_left, const Fab& shielded_right,
                                        const Fab& boundary, const CutCellData& g,
                                        const Box& cells, int ncons, double dt_over_dx, int dir) {
  if (!cells.ok()) return;
  int e[3] = {0, 0, 0};
  e[dir] = 1;
  for (int c = 0; c < ncons; ++c)
    for (int k = cells.lo[2]; k <= cells.hi[2]; ++k)
      for (int j = cells.lo[1]; j <= cells.hi[1]; ++j)
        for (int i = cells.lo[0]; i <= cells.hi[0]; ++i) {
          const double alpha = g.alpha(i, j, k, 0);
          if (alpha == 0.0) {
            scratch(i, j, k, c) = 0.0;
            continue;
          }
          const int ir = i + e[0], jr = j + e[1], kr = k + e[2];
          const double betaL = g.beta[dir](i, j, k, 0);
          const double betaR = g.beta[dir](ir, jr, kr, 0);
          const double prev = scratch(i, j, k, c);
          const double fLs = stabilized(i, j, k, c);
          const double fRs = stabilized(ir, jr, kr, c);
          const double fB = boundary(i, j, k, c);
          const double center = g.bcent(i, j, k, dir);
          const double regular = prev + dt_over_dx * (fLs - fRs);
          double next;
          if (betaL == betaR) {
            next = regular;
          } else if (betaL > betaR) {
            const double betaL_us_over_alpha = g.us[dir](i, j, k, 0) / alpha;
            const double betaL_sL_over_alpha = g.sl[dir](i, j, k, 0) / alpha;
            const double distance_to_boundary = 0.5 + center;
            const double betaL_sR = g.sr[dir](i, j, k, 0);
            const double dBeta_over_alpha =
                std::clamp((betaL_sR * distance_to_boundary) / alpha, 0.0, 1.0);
            const double fLus = unshielded(i, j, k, c);
            const double dF_us = fLus - fRs;
            const double dF_sL = shielded_left(i, j, k, c) - fRs;
            const double dB = fLus - fB;
            next = prev + dt_over_dx * betaL_us_over_alpha * dF_us +
                   dt_over_dx * betaL_sL_over_alpha * dF_sL + dt_over_dx * dBeta_over_alpha * dB;
          } else {
            const double betaR_us_over_alpha = g.us[dir](ir, jr, kr, 0) / alpha;
            const double betaR_sR_over_alpha = g.sr[dir](ir, jr, kr, 0) / alpha;
            const double betaR_sL = g.sl[dir](ir, jr, kr, 0);
            const double distance_to_boundary = 0.5 - center;
            const double dBeta_over_alpha =
                std::clamp((betaR_sL * distance_to_boundary) / alpha, 0.0, 1.0);
            const double fRus = unshielded(ir, jr, kr, c);
            const double dF_us = fLs - fRus;
            const double dF_sR = fLs - shielded_right(ir, jr, kr, c);
            const double dB = fB - fRus;
            next = prev + dt_over_dx * betaR_us_over_alpha * dF_us +
                   dt_over_dx * betaR_sR_over_alpha * dF_sR + dt_over_dx * dBeta_over_alpha * dB;
          }
          scratch(i, j, k, c) = next;
        }
}

// include/fub/AMReX/cutcell/Reconstruction.hpp:124-157 (masked by alpha > 0)
template <int D>
void CompleteFromConsBoxMasked(const PerfectGasEq& eq, Fab& fab, const Fab& alpha, const Box& cells) {
  if (!cells.ok()) return;
  Cons<A8, D> cons;
  Complete<A8, D> complete;
  std::memset(static_cast<void*>(&cons), 0, sizeof(cons));
  for (int k = cells.lo[2]; k <= cells.hi[2]; ++k)
    for (int j = cells.lo[1]; j <= cells.hi[1]; ++j) {
      int i = cells.lo[0];
      int n = cells.hi[0] - i + 1;
      while (n > 0) {
        const int nn = std::min(n, kChunk);
        load_cons<D>(cons, fab, fab.index(i, j, k), nn);
        A8 a(0.0);
        for (int l = 0; l < nn; ++l) a.v[l] = alpha(i + l, j, k, 0);
        const M8 mask = a > 0.0;
        CompleteFromCons<A8, D>(eq, complete, cons, mask);
        store_complete<D>(fab, fab.index(i, j, k), complete, nn);
        i += nn;
        n -= nn;
      }
    }
}

// ---------------------------------------------------------------------------
// Tagging (patch level).
// include/fub/tagging_method/GradientDetector.hpp:156-232: `states` is read on
// `cells` grown by one, tags are written on `cells`; `alpha` may be null.
inline void TagGradient(char* tags, const Box& cells, const Fab& states, const Fab* alpha,
                        int rank, int ncond, const int* comps, const double* tols) {
  const long nx = cells.len(0), ny = cells.len(1);
  for (int dir = 0; dir < rank; ++dir) {
    int e[3] = {0, 0, 0};
    e[dir] = 1;
    for (int k = cells.lo[2]; k <= cells.hi[2]; ++k)
      for (int j = cells.lo[1]; j <= cells.hi[1]; ++j)
        for (int i = cells.lo[0]; i <= cells.hi[0]; ++i) {
          if (alpha) {
            if (!((*alpha)(i - e[0], j - e[1], k - e[2], 0) > 0.0 && (*alpha)(i, j, k, 0) > 0.0 &&
                  (*alpha)(i + e[0], j + e[1], k + e[2], 0) > 0.0)) {
              continue;
            }
          }
          char& t = tags[(i - cells.lo[0]) + nx * ((j - cells.lo[1]) + ny * (k - cells.lo[2]))];
          for (int c = 0; c < ncond; ++c) {
            const double xL = states(i - e[0], j - e[1], k - e[2], comps[c]);
            const double xM = states(i, j, k, comps[c]);
            const double xR = states(i + e[0], j + e[1], k + e[2], comps[c]);
            if (xM != xR || xM != xL) {
              const double left = std::abs(xM - xL) / (std::abs(xM) + std::abs(xL));
              const double right = std::abs(xM - xR) / (std::abs(xM) + std::abs(xR));
              t |= static_cast<char>(left > tols[c] || right > tols[c]);
            }
          }
        }
  }
}
// src/tagging_method/TagCutCells.cpp:27-35
inline void TagCutCells(char* tags, const Box& cells, const Fab& alpha) {
  const long nx = cells.len(0), ny = cells.len(1);
  for (int k = cells.lo[2]; k <= cells.hi[2]; ++k)
    for (int j = cells.lo[1]; j <= cells.hi[1]; ++j)
      for (int i = cells.lo[0]; i <= cells.hi[0]; ++i) {
        const double a = alpha(i, j, k, 0);
        if (0.0 < a && a < 1.0) {
          tags[(i - cells.lo[0]) + nx * ((j - cells.lo[1]) + ny * (k - cells.lo[2]))] = '\1';
        }
      }
}
// src/tagging_method/TagBuffer.cpp:27-86 on a dense view of extents n[].
// (1-D: the reference reads index -1 when a tag sits at i == width; guarded.)
inline void TagBuffer(char* data, int rank, const int n[3], int width) {
  const std::ptrdiff_t w = width;
  if (rank == 1) {
    const std::ptrdiff_t size = n[0];
    for (std::ptrdiff_t i = w; i < size - w; ++i) {
      if (data[i] == '\1') {
        for (std::ptrdiff_t j = 0; j < w; ++j) data[i + j] = static_cast<char>('\2' - data[i + j]);
        for (std::ptrdiff_t j = w; j > 0; --j) {
          if (i - j - 1 >= 0) data[i - j - 1] = static_cast<char>('\2' - data[i - j - 1]);
        }
      }
    }
  } else if (rank == 2) {
    const std::ptrdiff_t nx = n[0], ny = n[1];
    for (std::ptrdiff_t j = 0; j < ny; ++j)
      for (std::ptrdiff_t i = 0; i < nx; ++i) {
        if (data[i + nx * j] == '\1' || data[i + nx * j] == '\2') {
          const std::ptrdiff_t k0 = std::max(i - w, std::ptrdiff_t{0});
          const std::ptrdiff_t k1 = std::min(i + w, nx);
          const std::ptrdiff_t l0 = std::max(j - w, std::ptrdiff_t{0});
          const std::ptrdiff_t l1 = std::min(j + w, ny);
          for (std::ptrdiff_t l = l0; l < l1; ++l)
            for (std::ptrdiff_t k = k0; k < k1; ++k) {
              char& d = data[k + nx * l];
              d = std::max(d, static_cast<char>('\3' - d));
            }
        }
      }
  } else {
    const std::ptrdiff_t nx = n[0], ny = n[1], nz = n[2];
    for (std::ptrdiff_t k0 = w; k0 < nz - w; ++k0)
      for (std::ptrdiff_t j0 = w; j0 < ny - w; ++j0)
        for (std::ptrdiff_t i0 = w; i0 < nx - w; ++i0)
          for (std::ptrdiff_t k = k0 - w; k < k0 + w; ++k)
            for (std::ptrdiff_t j = j0 - w; j < j0 + w; ++j)
              for (std::ptrdiff_t i = i0 - w; i < i0 + w; ++i) {
                char& d = data[i + nx * (j + ny * k)];
                d = std::max(d, static_cast<char>('\2' - d));
              }
  }
}

} // namespace fvso
