// amr.cu -- coarse-fine kernels between two levels held by one rank.
//
// Reference (paths relative to maikel/FiniteVolumeSolver): the fub side is
// src/AMReX/IntegratorContext.cpp:380-420 (FillGhostLayerTwoLevels: pc_interp),
// :451-470 (CoarsenConservatively: amrex::average_down; cut-cell levels:
// amrex::EB_average_down, src/AMReX/cutcell/IntegratorContext.cpp:662-677), :472-497
// (AccumulateCoarseFineFluxes: FluxRegister::FineAdd / CrseInit(ADD)), :499-512
// (ResetCoarseFineFluxes), :514-531 (ApplyFluxCorrection: FluxRegister::Reflux).
// The arithmetic itself is AMReX's (unpinned dependency, absent here): restated
// from the published algorithms, identical to oracle/amr.hpp ("parity
// unpinned" against a real AMReX build).
#include <algorithm>

#include "launch.h"

namespace fvs {
namespace {

inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

__device__ __forceinline__ long long at(const ArrayRef& a, int i, int j, int k) {
  return (i - a.lo[0]) + (long long)a.n[0] * ((j - a.lo[1]) + (long long)a.n[1] * (k - a.lo[2]));
}
__device__ __forceinline__ int coarsen_index(int i, int r) {
  return i >= 0 ? i / r : -((-i + r - 1) / r);
}
struct Ratio {
  int r[3];
};

__global__ void __launch_bounds__(256)
average_down_kernel(const AvgDownDesc* descs, int ncons, Ratio R) {
  const AvgDownDesc& D = descs[blockIdx.y];
  const long long cells = (long long)D.len[0] * D.len[1] * D.len[2];
  const double volfrac = 1.0 / double(R.r[0] * R.r[1] * R.r[2]);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < cells;
       t += (long long)gridDim.x * blockDim.x) {
    const int i = D.lo[0] + (int)(t % D.len[0]);
    const int j = D.lo[1] + (int)((t / D.len[0]) % D.len[1]);
    const int k = D.lo[2] + (int)(t / ((long long)D.len[0] * D.len[1]));
    const long long c_at = at(D.crse, i, j, k);
    if (D.vfrac) { // amrex_eb_avgdown: src/AMReX/cutcell/IntegratorContext.cpp:662-677
      for (int n = 0; n < ncons; ++n) {
        double c = 0.0;
        double cv = 0.0;
        for (int kr = 0; kr < R.r[2]; ++kr)
          for (int jr = 0; jr < R.r[1]; ++jr)
            for (int ir = 0; ir < R.r[0]; ++ir) {
              const long long f_at = at(D.fine, i * R.r[0] + ir, j * R.r[1] + jr, k * R.r[2] + kr);
              const double tmp = D.vfrac[f_at];
              cv += tmp;
              c += D.fine.p[f_at + D.fine.sc * n] * tmp;
            }
        if (cv > 1.e-30) {
          c /= cv;
        } else {
          c = D.fine.p[at(D.fine, i * R.r[0], j * R.r[1], k * R.r[2]) + D.fine.sc * n];
        }
        D.crse.p[c_at + D.crse.sc * n] = c;
      }
      continue;
    }
    for (int n = 0; n < ncons; ++n) {
      double c = 0.0;
      for (int kr = 0; kr < R.r[2]; ++kr)
        for (int jr = 0; jr < R.r[1]; ++jr)
          for (int ir = 0; ir < R.r[0]; ++ir) {
            c += D.fine.p[at(D.fine, i * R.r[0] + ir, j * R.r[1] + jr, k * R.r[2] + kr) +
                          D.fine.sc * n];
          }
      D.crse.p[c_at + D.crse.sc * n] = volfrac * c;
    }
  }
}

__global__ void __launch_bounds__(256)
pc_interp_kernel(const InterpDesc* descs, int ncomp, Ratio R) {
  const InterpDesc& D = descs[blockIdx.y];
  const long long cells = (long long)D.len[0] * D.len[1] * D.len[2];
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < cells;
       t += (long long)gridDim.x * blockDim.x) {
    const int i = D.lo[0] + (int)(t % D.len[0]);
    const int j = D.lo[1] + (int)((t / D.len[0]) % D.len[1]);
    const int k = D.lo[2] + (int)(t / ((long long)D.len[0] * D.len[1]));
    if (i >= D.vlo[0] && i <= D.vhi[0] && j >= D.vlo[1] && j <= D.vhi[1] && k >= D.vlo[2] &&
        k <= D.vhi[2]) {
      continue;
    }
    const long long f_at = at(D.fine, i, j, k);
    const long long c_at = at(D.crse, coarsen_index(i, R.r[0]), coarsen_index(j, R.r[1]),
                              coarsen_index(k, R.r[2]));
    for (int n = 0; n < ncomp; ++n) D.fine.p[f_at + D.fine.sc * n] = D.crse.p[c_at + D.crse.sc * n];
  }
}

__global__ void __launch_bounds__(128)
flux_register_kernel(const RegisterDesc* descs, int op, int ncons, int dir, Ratio R, double mult,
                     double vol) {
  const RegisterDesc& D = descs[blockIdx.y];
  const long long faces = (long long)D.len[0] * D.len[1] * D.len[2];
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < faces;
       t += (long long)gridDim.x * blockDim.x) {
    const int i = D.lo[0] + (int)(t % D.len[0]);
    const int j = D.lo[1] + (int)((t / D.len[0]) % D.len[1]);
    const int k = D.lo[2] + (int)(t / ((long long)D.len[0] * D.len[1]));
    const long long r_at = at(D.reg, i, j, k);
    if (op == 0) { // FineAdd: every fine face of the coarse face, (k, j, i) order
      int lo[3] = {i * R.r[0], j * R.r[1], k * R.r[2]};
      int hi[3] = {lo[0] + R.r[0] - 1, lo[1] + R.r[1] - 1, lo[2] + R.r[2] - 1};
      hi[dir] = lo[dir];
      for (int c = 0; c < ncons; ++c) {
        double acc = D.reg.p[r_at + D.reg.sc * c];
        for (int kk = lo[2]; kk <= hi[2]; ++kk)
          for (int jj = lo[1]; jj <= hi[1]; ++jj)
            for (int ii = lo[0]; ii <= hi[0]; ++ii) {
              acc += mult * D.other.p[at(D.other, ii, jj, kk) + D.other.sc * c];
            }
        D.reg.p[r_at + D.reg.sc * c] = acc;
      }
      continue;
    }
    int out[3] = {i, j, k};
    if (D.side == 0) out[dir] -= 1;
    const bool inside = out[0] >= D.box_lo[0] && out[0] <= D.box_hi[0] && out[1] >= D.box_lo[1] &&
                        out[1] <= D.box_hi[1] && out[2] >= D.box_lo[2] && out[2] <= D.box_hi[2];
    if (!inside) continue;
    if (op == 1) { // CrseInit(ADD)
      const long long f_at = at(D.other, i, j, k);
      for (int c = 0; c < ncons; ++c) {
        D.reg.p[r_at + D.reg.sc * c] += mult * D.other.p[f_at + D.other.sc * c];
      }
    } else { // Reflux: mult = scale, vol = coarse cell volume
      if (!D.open[r_at]) continue;
      const long long s_at = at(D.other, out[0], out[1], out[2]);
      const double sign = D.side == 0 ? -1.0 : 1.0;
      for (int c = 0; c < ncons; ++c) {
        D.other.p[s_at + D.other.sc * c] += sign * mult * D.reg.p[r_at + D.reg.sc * c] / vol;
      }
    }
  }
}

} // namespace

int launch_average_down(const AvgDownDesc* d, int n, int ncons, const int ratio[3],
                        long long max_cells, cudaStream_t s) {
  if (n <= 0) return 0;
  Ratio R{{ratio[0], ratio[1], ratio[2]}};
  dim3 grid(std::min(cdiv(max_cells, 256), 4096u), n, 1);
  average_down_kernel<<<grid, 256, 0, s>>>(d, ncons, R);
  return 1;
}
int launch_pc_interp(const InterpDesc* d, int n, int ncomp, const int ratio[3], long long max_cells,
                     cudaStream_t s) {
  if (n <= 0) return 0;
  Ratio R{{ratio[0], ratio[1], ratio[2]}};
  dim3 grid(std::min(cdiv(max_cells, 256), 4096u), n, 1);
  pc_interp_kernel<<<grid, 256, 0, s>>>(d, ncomp, R);
  return 1;
}
int launch_flux_register(const RegisterDesc* d, int n, int op, int ncons, int dir,
                         const int ratio[3], double mult, double vol, long long max_faces,
                         cudaStream_t s) {
  if (n <= 0) return 0;
  Ratio R{{ratio[0], ratio[1], ratio[2]}};
  dim3 grid(std::min(cdiv(max_faces, 128), 1024u), n, 1);
  flux_register_kernel<<<grid, 128, 0, s>>>(d, op, ncons, dir, R, mult, vol);
  return 1;
}

} // namespace fvs
