// cutcell.cu -- method-independent kernels of the embedded-boundary path
// (Klein-2009 KbnCutCellMethod and its adapters) and the tagging kernels.
// Plain IEEE operators, -fmad=false: bit-identical to the oracle's restatement
// of the reference (oracle/cutcell.hpp).
//
// Reference (paths relative to maikel/FiniteVolumeSolver):
//   src/AMReX/cutcell/FillCutCellData.cpp:28-118          fill_cutcell_data_kernel
//   src/CutCellData.cpp:90-105,225-256                    is_cut_cell, boundary normal
//   include/fub/ext/Eigen.hpp:104-130                     make_rotation
//   src/equations/PerfectGas.cpp:230-300                  rotate, reflect
//   include/fub/flux_method/HllMethod.hpp:162-184         hll_solve (scalar Riemann solver)
//   include/fub/cutcell_method/KbnStabilisation.hpp:188-273   reference states, boundary fluxes
//   src/cutcell_method/KbnStabilisation.cpp:55-159        stabilised fluxes
//   src/AMReX/cutcell/TimeIntegrator.cpp:65-411           cut-cell conservative update
//   include/fub/AMReX/cutcell/Reconstruction.hpp:124-193  masked CompleteFromCons
//   include/fub/tagging_method/GradientDetector.hpp:156-232, src/tagging_method/*.cpp
#include "launch.h"

namespace fvs {
namespace {

inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

// std::min / std::max / std::clamp semantics of the reference's scalar code
__device__ __forceinline__ double std_min(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double std_max(double a, double b) { return (a < b) ? b : a; }
__device__ __forceinline__ double std_clamp(double v, double lo, double hi) {
  return (v < lo) ? lo : ((hi < v) ? hi : v);
}

// src/CutCellData.cpp:225-256; (i, j, k) array-relative, beta[d] has extents
// n + e_d
template <int D>
__device__ bool is_cut_cell(const double* alpha, const double* const* beta, const int* n, int i,
                            int j, int k) {
  const double a = alpha[i + (long long)n[0] * (j + (long long)n[1] * k)];
  if (a <= 0.0) return false;
  if (a < 1.0) return true;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const int bn0 = n[0] + (d == 0 ? 1 : 0);
    const int bn1 = n[1] + (d == 1 ? 1 : 0);
    const long long left = i + (long long)bn0 * (j + (long long)bn1 * k);
    const long long step = d == 0 ? 1 : (d == 1 ? (long long)bn0 : (long long)bn0 * bn1);
    if (beta[d][left] < 1.0 || beta[d][left + step] < 1.0) return true;
  }
  return false;
}

template <int D> struct Mat {
  double m[D][D];
};
template <int D> __device__ double dotD(const double* a, const double* b) {
  if constexpr (D == 2) {
    return a[0] * b[0] + a[1] * b[1];
  } else {
    return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]);
  }
}
// include/fub/ext/Eigen.hpp:104-130
template <int D> __device__ Mat<D> make_rotation(const double* a, const double* b) {
  Mat<D> R;
  bool opposite = true;
#pragma unroll
  for (int d = 0; d < D; ++d) opposite = opposite && (a[d] == -b[d]);
  if (opposite) {
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int c = 0; c < D; ++c) R.m[r][c] = (r == c) ? -1.0 : -0.0;
    return R;
  }
  if constexpr (D == 2) {
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
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        const double p0 = skew[r][0] * skew[0][cc];
        const double p1 = skew[r][1] * skew[1][cc];
        const double p2 = skew[r][2] * skew[2][cc];
        const double sq = p0 + (p1 + p2);
        const double I = (r == cc) ? 1.0 : 0.0;
        R.m[r][cc] = (I + skew[r][cc]) + sq / one_plus_c;
      }
  }
  return R;
}
template <int D> __device__ void rotate(Complete<D>& s, const Mat<D>& R) {
  double tmp[D];
#pragma unroll
  for (int r = 0; r < D; ++r) {
    if constexpr (D == 2) {
      tmp[r] = R.m[r][0] * s.mom[0] + R.m[r][1] * s.mom[1];
    } else {
      tmp[r] = R.m[r][0] * s.mom[0] + (R.m[r][1] * s.mom[1] + R.m[r][2] * s.mom[2]);
    }
  }
#pragma unroll
  for (int r = 0; r < D; ++r) s.mom[r] = tmp[r];
}
template <int D> __device__ Complete<D> reflect(const Complete<D>& s, const double* normal) {
  Complete<D> r = s;
  const double mn = dotD<D>(s.mom, normal);
#pragma unroll
  for (int d = 0; d < D; ++d) r.mom[d] = s.mom[d] - 2 * (mn * normal[d]);
  return r;
}

// scalar Flux (PerfectGas.cpp:52-64), runtime direction
template <int D> __device__ void flux_scalar(Cons<D>& f, const Complete<D>& s, int dir) {
  const double velocity = s.mom[dir] / s.rho;
  f.rho = s.mom[dir];
#pragma unroll
  for (int d = 0; d < D; ++d) f.mom[d] = velocity * s.mom[d];
  f.mom[dir] = f.mom[dir] + s.p;
  f.E = velocity * (s.E + s.p);
}
// scalar CompleteFromCons (PerfectGas.cpp:96-110)
template <int D> __device__ void complete_scalar(const EqConst& eq, Complete<D>& q) {
  const double e_kin = 0.5 * squared_norm<D>(q.mom) / q.rho;
  const double e_int = q.E - e_kin;
  q.p = e_int / eq.gamma_minus_1_inv;
  q.a = ::sqrt(eq.gamma * q.p / q.rho);
}

// HllBase::SolveRiemannProblem with EinfeldtSignalVelocities (scalar), x direction
template <int D>
__device__ void hll_solve(const EqConst& eq, Complete<D>& solution, const Complete<D>& left,
                          const Complete<D>& right) {
  const double rhoL = left.rho, rhoR = right.rho;
  const double sqRhoL = ::sqrt(rhoL);
  const double sqRhoR = ::sqrt(rhoR);
  const double oosqRhoSum = 1.0 / (sqRhoL + sqRhoR);
  const double uL = left.mom[0] / rhoL;
  const double uR = right.mom[0] / rhoR;
  const double roeU = (uL * sqRhoL + uR * sqRhoR) * oosqRhoSum;
  const double hL = (left.E + left.p) / rhoL;
  const double hR = (right.E + right.p) / rhoR;
  const double roeH = (hL * sqRhoL + hR * sqRhoR) * oosqRhoSum;
  const double roeA2 = (eq.gamma - 1.0) * (roeH - 0.5 * roeU * roeU);
  const double roeA = ::sqrt(roeA2);
  const double sL = std_min(uL - left.a, roeU - roeA);
  const double sR = std_max(roeU + roeA, uR + right.a);
  Cons<D> fL, fR;
  flux_scalar<D>(fL, left, 0);
  flux_scalar<D>(fR, right, 0);
  const double ds = sL != sR ? sR - sL : 1.0;
  if (0.0 < sL) {
    solution = left;
  } else if (sR < 0.0) {
    solution = right;
  } else {
    solution.rho = (sR * right.rho - sL * left.rho + fL.rho - fR.rho) / ds;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      solution.mom[d] = (sR * right.mom[d] - sL * left.mom[d] + fL.mom[d] - fR.mom[d]) / ds;
    }
    solution.E = (sR * right.E - sL * left.E + fL.E - fR.E) / ds;
    complete_scalar<D>(eq, solution);
  }
}

template <int D> __device__ Complete<D> load_state(const double* p, long long sc) {
  Complete<D> q;
  q.rho = p[0];
#pragma unroll
  for (int d = 0; d < D; ++d) q.mom[d] = p[sc * (1 + d)];
  q.E = p[sc * (1 + D)];
  q.p = p[sc * (2 + D)];
  q.a = p[sc * (3 + D)];
  return q;
}
template <int D> __device__ Complete<D> zero_state() {
  Complete<D> q;
  q.rho = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) q.mom[d] = 0.0;
  q.E = q.p = q.a = 0.0;
  return q;
}

// rotate the state so that the boundary normal becomes e_x, reflect it and
// solve the wall Riemann problem (KbnStabilisation.hpp:199-203, 229-236)
template <int D>
__device__ void wall_problem(const EqConst& eq, int riemann, Complete<D>& solution,
                             Complete<D>& state, const double* normal, int& failed) {
  double unit[D];
#pragma unroll
  for (int d = 0; d < D; ++d) unit[d] = d == 0 ? 1.0 : 0.0;
  rotate<D>(state, make_rotation<D>(normal, unit));
  const Complete<D> reflected = reflect<D>(state, unit);
  if (riemann == 1) {
    ExactRiemann<D, 0>(eq).solve(solution, reflected, state, failed);
  } else {
    hll_solve<D>(eq, solution, reflected, state);
  }
}

// a18  KbnCutCellMethod::PreAdvanceHierarchy: thread per cut cell of the
// compacted list
template <int D> __global__ void __launch_bounds__(128) reference_states_kernel(const EbArgs A) {
  const EbDesc& B = A.boxes[blockIdx.y];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (B.fab_type != 1 || t >= B.n_cut) return;
  const long long cell = B.cut_cells[t];
  Complete<D> state = load_state<D>(B.state + cell, B.sc);
  int failed = 0;
  if (state.rho > 0.0) {
    double normal[D];
#pragma unroll
    for (int d = 0; d < D; ++d) normal[d] = -B.bnorm[cell + B.sc * d];
    Complete<D> solution = zero_state<D>();
    wall_problem<D>(A.eq, A.riemann, solution, state, normal, failed);
    double unit[D];
#pragma unroll
    for (int d = 0; d < D; ++d) unit[d] = d == 0 ? 1.0 : 0.0;
    rotate<D>(solution, make_rotation<D>(unit, normal));
    double* r = B.reference + cell;
    r[0] = solution.rho;
#pragma unroll
    for (int d = 0; d < D; ++d) r[B.sc * (1 + d)] = solution.mom[d];
    r[B.sc * (1 + D)] = solution.E;
    r[B.sc * (2 + D)] = solution.p;
    r[B.sc * (3 + D)] = solution.a;
  }
  if (failed) *A.fail_flag = 1;
}

// a19  KbnCutCellMethod::ComputeBoundaryFluxes (mask == 0 branch)
template <int D> __global__ void __launch_bounds__(128) boundary_fluxes_kernel(const EbArgs A) {
  const EbDesc& B = A.boxes[blockIdx.y];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (B.fab_type != 1 || t >= B.n_cut) return;
  const long long cell = B.cut_cells[t];
  Complete<D> state = load_state<D>(B.state + cell, B.sc);
  const Complete<D> ref = load_state<D>(B.reference + cell, B.sc);
  double normal[D];
#pragma unroll
  for (int d = 0; d < D; ++d) normal[d] = -B.bnorm[cell + B.sc * d];
  Cons<D> flux;
  flux_scalar<D>(flux, ref, A.dir);
  Complete<D> solution = zero_state<D>();
  int failed = 0;
  wall_problem<D>(A.eq, A.riemann, solution, state, normal, failed);
  flux.mom[A.dir] += solution.p - ref.p;
  double* f = B.boundary_flux + cell;
  f[0] = flux.rho;
#pragma unroll
  for (int d = 0; d < D; ++d) f[B.sc * (1 + d)] = flux.mom[d];
  f[B.sc * (1 + D)] = flux.E;
  if (failed) *A.fail_flag = 1;
}

// a21  ComputeStableFluxes_Row: thread per face, loop over components
__global__ void __launch_bounds__(256) stabilised_fluxes_kernel(const EbArgs A) {
  const EbDesc& B = A.boxes[blockIdx.y];
  if (B.fab_type != 1) return;
  const int dir = A.dir;
  int len[3] = {B.n[0], B.n[1], B.n[2]};
  len[dir] = B.fhi - B.flo + 1;
  if (len[dir] <= 0) return;
  const long long total = (long long)len[0] * len[1] * len[2];
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int idx[3];
  idx[0] = (int)(t % len[0]);
  idx[1] = (int)((t / len[0]) % len[1]);
  idx[2] = (int)(t / ((long long)len[0] * len[1]));
  idx[dir] += B.flo;
  const long long stride = dir == 0 ? 1 : (dir == 1 ? (long long)B.n[0] : (long long)B.n[0] * B.n[1]);
  const long long cellR = idx[0] + (long long)B.n[0] * (idx[1] + (long long)B.n[1] * idx[2]);
  const long long cellL = cellR - stride;
  const long long face = idx[0] + (long long)B.bn[0] * (idx[1] + (long long)B.bn[1] * idx[2]);
  const long long fcell = (idx[0] - B.foff[0]) +
                          (long long)B.fn[0] * ((idx[1] - B.foff[1]) +
                                                (long long)B.fn[1] * (idx[2] - B.foff[2]));
  const double centerL = B.bcent[cellL + B.sc * dir];
  const double centerR = B.bcent[cellR + B.sc * dir];
  const double dL = std_clamp(0.5 - centerL, 0.0, 1.0);
  const double dR = std_clamp(0.5 + centerR, 0.0, 1.0);
  const double betaL = B.sl[face + B.bsc];
  const double betaR = B.sr[face + B.bsc];
  const double betaUS = B.us[face + B.bsc];
  const int ncons = A.rank + 2;
  for (int c = 0; c < ncons; ++c) {
    const double f = B.flux[fcell + B.fsc * c];
    const double fsL = dL * f + (1.0 - dL) * B.boundary_flux[cellL + B.sc * c];
    const double fsR = dR * f + (1.0 - dR) * B.boundary_flux[cellR + B.sc * c];
    B.shl[fcell + B.fsc * c] = fsL;
    B.shr[fcell + B.fsc * c] = fsR;
    B.stab[fcell + B.fsc * c] = betaUS * f + betaL * fsL + betaR * fsR;
  }
}

// a23 (+ a13 for the regular fabs of an EB level): thread per cell
__global__ void __launch_bounds__(256) cc_update_kernel(const EbArgs A) {
  const EbDesc& B = A.boxes[blockIdx.y];
  if (B.fab_type < 0) return;
  const int dir = A.dir;
  int len[3] = {B.n[0], B.n[1], B.n[2]};
  len[dir] = B.uhi - B.ulo + 1;
  if (len[dir] <= 0) return;
  const long long total = (long long)len[0] * len[1] * len[2];
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int idx[3];
  idx[0] = (int)(t % len[0]);
  idx[1] = (int)((t / len[0]) % len[1]);
  idx[2] = (int)(t / ((long long)len[0] * len[1]));
  idx[dir] += B.ulo;
  const long long cell = idx[0] + (long long)B.n[0] * (idx[1] + (long long)B.n[1] * idx[2]);
  const long long fL = (idx[0] - B.foff[0]) +
                       (long long)B.fn[0] * ((idx[1] - B.foff[1]) +
                                             (long long)B.fn[1] * (idx[2] - B.foff[2]));
  const long long fstride =
      dir == 0 ? 1 : (dir == 1 ? (long long)B.fn[0] : (long long)B.fn[0] * B.fn[1]);
  const long long fR = fL + fstride;
  const int ncons = A.rank + 2;
  const double dt_over_dx = A.lambda;
  double* u = B.state + cell;
  if (B.fab_type == 0) {
    for (int c = 0; c < ncons; ++c) {
      u[B.sc * c] = u[B.sc * c] + dt_over_dx * (B.flux[fL + B.fsc * c] - B.flux[fR + B.fsc * c]);
    }
    return;
  }
  const double alpha = B.alpha[cell];
  if (alpha == 0.0) {
    for (int c = 0; c < ncons; ++c) u[B.sc * c] = 0.0;
    return;
  }
  const long long faceL = idx[0] + (long long)B.bn[0] * (idx[1] + (long long)B.bn[1] * idx[2]);
  const long long bstride =
      dir == 0 ? 1 : (dir == 1 ? (long long)B.bn[0] : (long long)B.bn[0] * B.bn[1]);
  const long long faceR = faceL + bstride;
  const double betaL = B.beta[faceL];
  const double betaR = B.beta[faceR];
  const double center = B.bcent[cell + B.sc * dir];
  if (betaL == betaR) {
    for (int c = 0; c < ncons; ++c) {
      u[B.sc * c] = u[B.sc * c] + dt_over_dx * (B.stab[fL + B.fsc * c] - B.stab[fR + B.fsc * c]);
    }
  } else if (betaL > betaR) {
    const double betaL_us_over_alpha = B.us[faceL] / alpha;
    const double betaL_sL_over_alpha = B.sl[faceL] / alpha;
    const double distance_to_boundary = 0.5 + center;
    const double dBeta_over_alpha =
        std_clamp((B.sr[faceL] * distance_to_boundary) / alpha, 0.0, 1.0);
    for (int c = 0; c < ncons; ++c) {
      const double fRs = B.stab[fR + B.fsc * c];
      const double fLus = B.flux[fL + B.fsc * c];
      const double dF_us = fLus - fRs;
      const double dF_sL = B.shl[fL + B.fsc * c] - fRs;
      const double dB = fLus - B.boundary_flux[cell + B.sc * c];
      u[B.sc * c] = u[B.sc * c] + dt_over_dx * betaL_us_over_alpha * dF_us +
                    dt_over_dx * betaL_sL_over_alpha * dF_sL + dt_over_dx * dBeta_over_alpha * dB;
    }
  } else {
    const double betaR_us_over_alpha = B.us[faceR] / alpha;
    const double betaR_sR_over_alpha = B.sr[faceR] / alpha;
    const double distance_to_boundary = 0.5 - center;
    const double dBeta_over_alpha =
        std_clamp((B.sl[faceR] * distance_to_boundary) / alpha, 0.0, 1.0);
    for (int c = 0; c < ncons; ++c) {
      const double fLs = B.stab[fL + B.fsc * c];
      const double fRus = B.flux[fR + B.fsc * c];
      const double dF_us = fLs - fRus;
      const double dF_sR = fLs - B.shr[fR + B.fsc * c];
      const double dB = B.boundary_flux[cell + B.sc * c] - fRus;
      u[B.sc * c] = u[B.sc * c] + dt_over_dx * betaR_us_over_alpha * dF_us +
                    dt_over_dx * betaR_sR_over_alpha * dF_sR + dt_over_dx * dBeta_over_alpha * dB;
    }
  }
}

// a24  cutcell::Reconstruction::CompleteFromCons; complete_mode holds the
// flag type of the cell's grown tile box (0 covered: skip, 1 regular: plain,
// 2 single valued: masked by alpha > 0 and rho > 0, zeros otherwise)
template <int D> __global__ void __launch_bounds__(256) cc_complete_kernel(const EbArgs A) {
  const EbDesc& B = A.boxes[blockIdx.y];
  const long long total = (long long)B.n[0] * B.n[1] * B.n[2];
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int mode = B.complete_mode[t];
  if (mode == 0) return;
  double* u = B.state + t;
  Complete<D> q;
  q.rho = u[0];
#pragma unroll
  for (int d = 0; d < D; ++d) q.mom[d] = u[B.sc * (1 + d)];
  q.E = u[B.sc * (1 + D)];
  if (mode == 2 && !(B.alpha[t] > 0.0 && q.rho > 0.0)) {
    u[0] = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) u[B.sc * (1 + d)] = 0.0;
    u[B.sc * (1 + D)] = 0.0;
    u[B.sc * (2 + D)] = 0.0;
    u[B.sc * (3 + D)] = 0.0;
    return;
  }
  complete_scalar<D>(A.eq, q);
  u[B.sc * (2 + D)] = q.p;
  u[B.sc * (3 + D)] = q.a;
}

// a25  FillCutCellData__: thread per face with both neighbours along dir
__global__ void __launch_bounds__(256)
fill_cutcell_data_kernel(const double* beta, double* us, double* sl, double* sr, double* ds, int b0,
                         int b1, int b2, int dir) {
  const long long total = (long long)b0 * b1 * b2;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int idx[3];
  idx[0] = (int)(t % b0);
  idx[1] = (int)((t / b0) % b1);
  idx[2] = (int)(t / ((long long)b0 * b1));
  const int bn[3] = {b0, b1, b2};
  if (idx[dir] < 1 || idx[dir] > bn[dir] - 2) return;
  const long long stride = dir == 0 ? 1 : (dir == 1 ? (long long)b0 : (long long)b0 * b1);
  const double betaL = beta[t - stride];
  double betaM = beta[t];
  const double betaR = beta[t + stride];
  const double dBetaL = std_max(0.0, betaM - betaL);
  const double dBetaR = std_max(0.0, betaM - betaR);
  const double unshielded = std_max(0.0, betaM - std_max(dBetaL, dBetaR));
  const double doubly_shielded = std_min(dBetaL, dBetaR);
  const double shielded_left = dBetaL - doubly_shielded;
  const double shielded_right = dBetaR - doubly_shielded;
  betaM = betaM > 0 ? betaM : 1.0;
  us[t] = unshielded;
  us[t + total] = unshielded / betaM;
  ds[t] = doubly_shielded;
  ds[t + total] = doubly_shielded / betaM;
  sl[t] = shielded_left;
  sl[t + total] = shielded_left / betaM;
  sr[t] = shielded_right;
  sr[t + total] = shielded_right / betaM;
}

struct BetaPtrs {
  const double* p[3];
};
template <int D>
__global__ void __launch_bounds__(256)
cut_cell_flags_kernel(const double* alpha, BetaPtrs beta, int n0, int n1, int n2,
                      unsigned char* flags) {
  const long long total = (long long)n0 * n1 * n2;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int n[3] = {n0, n1, n2};
  const int i = (int)(t % n0);
  const int j = (int)((t / n0) % n1);
  const int k = (int)(t / ((long long)n0 * n1));
  flags[t] = is_cut_cell<D>(alpha, beta.p, n, i, j, k) ? 1 : 0;
}

// ---- tagging ---------------------------------------------------------------
struct TagConds {
  int n;
  int comps[8];
  double tols[8];
};
// a26  ScalarGradientDetector_::TagCellsForRefinement (with / without alpha)
__global__ void __launch_bounds__(256)
tag_gradient_kernel(const double* state, long long sc, int n0, int n1, int n2, const double* alpha,
                    int rank, int lo0, int lo1, int lo2, int l0, int l1, int l2, TagConds conds,
                    char* tags) {
  const long long total = (long long)l0 * l1 * l2;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int i = lo0 + (int)(t % l0);
  const int j = lo1 + (int)((t / l0) % l1);
  const int k = lo2 + (int)(t / ((long long)l0 * l1));
  const long long cell = i + (long long)n0 * (j + (long long)n1 * k);
  (void)n2;
  char tag = tags[t];
  for (int dir = 0; dir < rank; ++dir) {
    const long long stride = dir == 0 ? 1 : (dir == 1 ? (long long)n0 : (long long)n0 * n1);
    if (alpha) {
      if (!(alpha[cell - stride] > 0.0 && alpha[cell] > 0.0 && alpha[cell + stride] > 0.0)) continue;
    }
    for (int c = 0; c < conds.n; ++c) {
      const double* x = state + sc * conds.comps[c];
      const double xL = x[cell - stride];
      const double xM = x[cell];
      const double xR = x[cell + stride];
      if (xM != xR || xM != xL) {
        const double left = fabs(xM - xL) / (fabs(xM) + fabs(xL));
        const double right = fabs(xM - xR) / (fabs(xM) + fabs(xR));
        tag |= (char)(left > conds.tols[c] || right > conds.tols[c]);
      }
    }
  }
  tags[t] = tag;
}
// src/tagging_method/TagCutCells.cpp:27-35
__global__ void __launch_bounds__(256)
tag_cutcells_kernel(const double* alpha, int n0, int n1, int lo0, int lo1, int lo2, int l0, int l1,
                    int l2, char* tags) {
  const long long total = (long long)l0 * l1 * l2;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int i = lo0 + (int)(t % l0);
  const int j = lo1 + (int)((t / l0) % l1);
  const int k = lo2 + (int)(t / ((long long)l0 * l1));
  const double a = alpha[i + (long long)n0 * (j + (long long)n1 * k)];
  if (0.0 < a && a < 1.0) tags[t] = 1;
}
// src/tagging_method/TagBuffer.cpp:27-86.  The 2-D and 3-D loops of the
// reference are order independent, so every cell
// computes its final value from the ORIGINAL tags; 1-D is inherently
// sequential and runs on one thread.
__global__ void tag_buffer_1d_kernel(char* data, int size, int w) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  for (int i = w; i < size - w; ++i) {
    if (data[i] == 1) {
      for (int j = 0; j < w; ++j) data[i + j] = (char)(2 - data[i + j]);
      for (int j = w; j > 0; --j) {
        if (i - j - 1 >= 0) data[i - j - 1] = (char)(2 - data[i - j - 1]);
      }
    }
  }
}
__global__ void __launch_bounds__(256)
tag_buffer_2d_kernel(const char* in, char* out, int nx, int ny, int w) {
  const long long total = (long long)nx * ny;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int k = (int)(t % nx);
  const int l = (int)(t / nx);
  const char orig = in[t];
  // a source (i, j) with tag 1 or 2 reaches k in [i - w, i + w), l likewise
  bool hit = false;
  const int i0 = max(k - w + 1, 0), i1 = min(k + w, nx - 1);
  const int j0 = max(l - w + 1, 0), j1 = min(l + w, ny - 1);
  for (int j = j0; j <= j1 && !hit; ++j)
    for (int i = i0; i <= i1; ++i) {
      const char v = in[i + (long long)nx * j];
      if (v == 1 || v == 2) {
        hit = true;
        break;
      }
    }
  char d = orig;
  if (hit) {
    const char alt = (char)(3 - d);
    d = d > alt ? d : alt;
  }
  out[t] = d;
}
__global__ void __launch_bounds__(256)
tag_buffer_3d_kernel(char* data, int nx, int ny, int nz, int w) {
  const long long total = (long long)nx * ny * nz;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  if (nx <= 2 * w || ny <= 2 * w || nz <= 2 * w) return;
  const int i = (int)(t % nx);
  const int j = (int)((t / nx) % ny);
  const int k = (int)(t / ((long long)nx * ny));
  // union of the neighbourhoods [c - w, c + w) of all interior centres
  if (i > nx - 2 || j > ny - 2 || k > nz - 2) return;
  const char d = data[t];
  const char alt = (char)(2 - d);
  data[t] = d > alt ? d : alt;
}

} // namespace

#define FVS_EB_RANK(kernel, grid, block)                                                      \
  switch (A.rank) {                                                                           \
  case 2: kernel<2><<<grid, block, 0, s>>>(A); return 1;                                      \
  case 3: kernel<3><<<grid, block, 0, s>>>(A); return 1;                                      \
  }                                                                                           \
  return -1;

int launch_cc_reference_states(const EbArgs& A, cudaStream_t s) {
  if (A.max_cut <= 0) return 0;
  dim3 grid(cdiv(A.max_cut, 128), A.nboxes, 1);
  FVS_EB_RANK(reference_states_kernel, grid, 128)
}
int launch_cc_boundary_fluxes(const EbArgs& A, cudaStream_t s) {
  if (A.max_cut <= 0) return 0;
  dim3 grid(cdiv(A.max_cut, 128), A.nboxes, 1);
  FVS_EB_RANK(boundary_fluxes_kernel, grid, 128)
}
int launch_cc_stabilised_fluxes(const EbArgs& A, cudaStream_t s) {
  dim3 grid(cdiv(A.max_cells, 256), A.nboxes, 1);
  stabilised_fluxes_kernel<<<grid, 256, 0, s>>>(A);
  return 1;
}
int launch_cc_update(const EbArgs& A, cudaStream_t s) {
  dim3 grid(cdiv(A.max_cells, 256), A.nboxes, 1);
  cc_update_kernel<<<grid, 256, 0, s>>>(A);
  return 1;
}
int launch_cc_complete(const EbArgs& A, cudaStream_t s) {
  dim3 grid(cdiv(A.max_cells, 256), A.nboxes, 1);
  FVS_EB_RANK(cc_complete_kernel, grid, 256)
}
int launch_fill_cutcell_data(const double* beta, double* us, double* sl, double* sr, double* ds,
                             const int bn[3], int dir, cudaStream_t s) {
  const long long total = (long long)bn[0] * bn[1] * bn[2];
  fill_cutcell_data_kernel<<<cdiv(total, 256), 256, 0, s>>>(beta, us, sl, sr, ds, bn[0], bn[1],
                                                             bn[2], dir);
  return 1;
}
int launch_cut_cell_flags(const double* alpha, const double* const beta[3], const int n[3],
                          int rank, unsigned char* flags, cudaStream_t s) {
  const long long total = (long long)n[0] * n[1] * n[2];
  BetaPtrs b{{beta[0], beta[1], beta[2]}};
  if (rank == 2) {
    cut_cell_flags_kernel<2><<<cdiv(total, 256), 256, 0, s>>>(alpha, b, n[0], n[1], n[2], flags);
  } else if (rank == 3) {
    cut_cell_flags_kernel<3><<<cdiv(total, 256), 256, 0, s>>>(alpha, b, n[0], n[1], n[2], flags);
  } else {
    return -1;
  }
  return 1;
}
int launch_tag_gradient(const double* state, long long sc, const int n[3], const double* alpha,
                        int rank, const int lo[3], const int len[3], int ncond, const int* comps,
                        const double* tols, char* tags, cudaStream_t s) {
  if (ncond < 0 || ncond > 8) return -1;
  TagConds c{};
  c.n = ncond;
  for (int i = 0; i < ncond; ++i) {
    c.comps[i] = comps[i];
    c.tols[i] = tols[i];
  }
  const long long total = (long long)len[0] * len[1] * len[2];
  tag_gradient_kernel<<<cdiv(total, 256), 256, 0, s>>>(state, sc, n[0], n[1], n[2], alpha, rank,
                                                        lo[0], lo[1], lo[2], len[0], len[1],
                                                        len[2], c, tags);
  return 1;
}
int launch_tag_cutcells(const double* alpha, const int n[3], const int lo[3], const int len[3],
                        char* tags, cudaStream_t s) {
  const long long total = (long long)len[0] * len[1] * len[2];
  tag_cutcells_kernel<<<cdiv(total, 256), 256, 0, s>>>(alpha, n[0], n[1], lo[0], lo[1], lo[2],
                                                        len[0], len[1], len[2], tags);
  return 1;
}
int launch_tag_buffer(char* tags, char* scratch, int rank, const int n[3], int width,
                      cudaStream_t s) {
  const long long total = (long long)n[0] * n[1] * n[2];
  if (rank == 1) {
    tag_buffer_1d_kernel<<<1, 1, 0, s>>>(tags, n[0], width);
  } else if (rank == 2) {
    tag_buffer_2d_kernel<<<cdiv(total, 256), 256, 0, s>>>(tags, scratch, n[0], n[1], width);
    cudaMemcpyAsync(tags, scratch, (size_t)total, cudaMemcpyDeviceToDevice, s);
  } else {
    tag_buffer_3d_kernel<<<cdiv(total, 256), 256, 0, s>>>(tags, n[0], n[1], n[2], width);
  }
  return 1;
}

} // namespace fvs
