// generic_kernels.cu -- method-independent kernels: conservative update (K3),
// complete-from-cons (K4), ghost-cell copies, halo pack/unpack and the
// physical boundary conditions, plus the dispatch from (rank, base) to the
// per-unit launchers of sweep_inst.cu.
#include "sweep_kernels.cuh"

namespace fvs {

#define FVS_DECL(D, B)                                                                        \
  int launch_fused_sweep_D##D##_##B(const MethodKey&, int, const SweepArgs&, cudaStream_t);   \
  int launch_numeric_fluxes_D##D##_##B(const MethodKey&, int, const SweepArgs&, cudaStream_t);\
  int launch_stable_dt_D##D##_##B(const MethodKey&, int, const SweepArgs&, cudaStream_t);     \
  int launch_stable_dt_all_D##D##_##B(const MethodKey&, const DtAllArgs&, cudaStream_t);
#define FVS_DECL_D(D) FVS_DECL(D, 0) FVS_DECL(D, 1) FVS_DECL(D, 2) FVS_DECL(D, 3)
FVS_DECL_D(1)
FVS_DECL_D(2)
FVS_DECL_D(3)
#define FVS_DECL_CC(D, B)                                                                     \
  int launch_cc_numeric_fluxes_D##D##_##B(const MethodKey&, int, const SweepArgs&,           \
                                          cudaStream_t);                                     \
  int launch_cc_stable_dt_D##D##_##B(const MethodKey&, int, const SweepArgs&, cudaStream_t);
#define FVS_DECL_CC_D(D) FVS_DECL_CC(D, 0) FVS_DECL_CC(D, 1) FVS_DECL_CC(D, 2) FVS_DECL_CC(D, 3)
FVS_DECL_CC_D(2)
FVS_DECL_CC_D(3)
#define FVS_ALL_CC(fn, ...)                                                                   \
  FVS_CASE(2, 0, fn, __VA_ARGS__) FVS_CASE(2, 1, fn, __VA_ARGS__) FVS_CASE(2, 2, fn, __VA_ARGS__) \
  FVS_CASE(2, 3, fn, __VA_ARGS__) FVS_CASE(3, 0, fn, __VA_ARGS__) FVS_CASE(3, 1, fn, __VA_ARGS__) \
  FVS_CASE(3, 2, fn, __VA_ARGS__) FVS_CASE(3, 3, fn, __VA_ARGS__)

#define FVS_CASE(D, B, fn, ...)                                                               \
  if (k.rank == D && k.base == B) return fn##_D##D##_##B(__VA_ARGS__);
#define FVS_ALL(fn, ...)                                                                      \
  FVS_CASE(1, 0, fn, __VA_ARGS__) FVS_CASE(1, 1, fn, __VA_ARGS__) FVS_CASE(1, 2, fn, __VA_ARGS__) \
  FVS_CASE(1, 3, fn, __VA_ARGS__) FVS_CASE(2, 0, fn, __VA_ARGS__) FVS_CASE(2, 1, fn, __VA_ARGS__) \
  FVS_CASE(2, 2, fn, __VA_ARGS__) FVS_CASE(2, 3, fn, __VA_ARGS__) FVS_CASE(3, 0, fn, __VA_ARGS__) \
  FVS_CASE(3, 1, fn, __VA_ARGS__) FVS_CASE(3, 2, fn, __VA_ARGS__) FVS_CASE(3, 3, fn, __VA_ARGS__)

int launch_fused_sweep(const MethodKey& k, int dir, const SweepArgs& A, cudaStream_t s) {
  FVS_ALL(launch_fused_sweep, k, dir, A, s)
  return -1;
}
// tolerance-arithmetic build of the fused sweep (sweep_inst.cu, FVS_TOLERANCE_UNIT);
// the exact Riemann solver (libm pow, Newton) has no such build
#define FVS_DECL_TOL(D, B)                                                                    \
  int launch_fused_sweep_tol_D##D##_##B(const void*, int, const void*, cudaStream_t);         \
  int launch_stable_dt_all_tol_D##D##_##B(const void*, const void*, cudaStream_t);
#define FVS_DECL_TOL_D(D) FVS_DECL_TOL(D, 0) FVS_DECL_TOL(D, 1) FVS_DECL_TOL(D, 2)
FVS_DECL_TOL_D(1)
FVS_DECL_TOL_D(2)
FVS_DECL_TOL_D(3)
int launch_fused_sweep_tolerance(const MethodKey& k, int dir, const SweepArgs& A, cudaStream_t s) {
#define FVS_CASE_TOL(D, B)                                                                    \
  if (k.rank == D && k.base == B) return launch_fused_sweep_tol_D##D##_##B(&k, dir, &A, s);
  FVS_CASE_TOL(1, 0) FVS_CASE_TOL(1, 1) FVS_CASE_TOL(1, 2)
  FVS_CASE_TOL(2, 0) FVS_CASE_TOL(2, 1) FVS_CASE_TOL(2, 2)
  FVS_CASE_TOL(3, 0) FVS_CASE_TOL(3, 1) FVS_CASE_TOL(3, 2)
  return -1;
}
int launch_stable_dt_all_tolerance(const MethodKey& k, const DtAllArgs& A, cudaStream_t s) {
#define FVS_CASE_TOL_DT(D, B)                                                                 \
  if (k.rank == D && k.base == B) return launch_stable_dt_all_tol_D##D##_##B(&k, &A, s);
  FVS_CASE_TOL_DT(1, 0) FVS_CASE_TOL_DT(1, 1) FVS_CASE_TOL_DT(1, 2)
  FVS_CASE_TOL_DT(2, 0) FVS_CASE_TOL_DT(2, 1) FVS_CASE_TOL_DT(2, 2)
  FVS_CASE_TOL_DT(3, 0) FVS_CASE_TOL_DT(3, 1) FVS_CASE_TOL_DT(3, 2)
  return -1;
}
int launch_numeric_fluxes(const MethodKey& k, int dir, const SweepArgs& A, cudaStream_t s) {
  FVS_ALL(launch_numeric_fluxes, k, dir, A, s)
  return -1;
}
int launch_stable_dt(const MethodKey& k, int dir, const SweepArgs& A, cudaStream_t s) {
  FVS_ALL(launch_stable_dt, k, dir, A, s)
  return -1;
}
int launch_cc_numeric_fluxes(const MethodKey& k, int dir, const SweepArgs& A, cudaStream_t s) {
  FVS_ALL_CC(launch_cc_numeric_fluxes, k, dir, A, s)
  return -1;
}
int launch_cc_stable_dt(const MethodKey& k, int dir, const SweepArgs& A, cudaStream_t s) {
  FVS_ALL_CC(launch_cc_stable_dt, k, dir, A, s)
  return -1;
}
int launch_stable_dt_all(const MethodKey& k, const DtAllArgs& A, cudaStream_t s) {
  FVS_ALL(launch_stable_dt_all, k, A, s)
  return -1;
}

namespace {
inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

// K3  HyperbolicPatchIntegrator::UpdateConservatively,
// src/HyperbolicPatchIntegrator.cpp:35-85 on the cells
// src/AMReX/TimeIntegrator.cpp:41-46 selects; in place.
template <int D, int DIR>
__global__ void __launch_bounds__(256) update_kernel(const SweepArgs A) {
  const BoxDesc& B = A.boxes[blockIdx.y];
  int len[3] = {B.n[0], B.n[1], B.n[2]};
  len[DIR] = B.uhi - B.ulo + 1;
  const long long total = (long long)len[0] * len[1] * len[2];
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (len[DIR] <= 0 || t >= total) return;
  int idx[3];
  idx[0] = (int)(t % len[0]);
  idx[1] = (int)((t / len[0]) % len[1]);
  idx[2] = (int)(t / ((long long)len[0] * len[1]));
  idx[DIR] += B.ulo;
  const long long cell = idx[0] + (long long)B.n[0] * (idx[1] + (long long)B.n[1] * idx[2]);
  const long long fcell =
      (idx[0] - B.foff[0]) +
      (long long)B.fn[0] * ((idx[1] - B.foff[1]) + (long long)B.fn[1] * (idx[2] - B.foff[2]));
  const long long fstride =
      DIR == 0 ? 1 : (DIR == 1 ? (long long)B.fn[0] : (long long)B.fn[0] * B.fn[1]);
  double* u = B.dst + cell;
  const double* f = B.flux + fcell;
#pragma unroll
  for (int c = 0; c < D + 2; ++c) {
    const double fl = f[B.fsc * c];
    const double fr = f[B.fsc * c + fstride];
    u[B.sc * c] = u[B.sc * c] + A.lambda * (fl - fr);
  }
}

// K4  CompleteFromConsFn, include/fub/CompleteFromCons.hpp:141-204; in place
template <int D> __global__ void __launch_bounds__(256) complete_kernel(const SweepArgs A) {
  const BoxDesc& B = A.boxes[blockIdx.y];
  const long long total = (long long)B.n[0] * B.n[1] * B.n[2];
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  double* u = B.dst + t;
  Complete<D> q;
  q.rho = u[0];
#pragma unroll
  for (int d = 0; d < D; ++d) q.mom[d] = u[B.sc * (1 + d)];
  q.E = u[B.sc * (1 + D)];
  const Eqs eq(A.eq);
  with_fallback([&](auto M, Flag& fl) {
    const Rcp rc = decltype(M)::rcp(q.rho);
    decltype(M)::check_window(rc, fl);
    complete_from_cons<decltype(M), D>(eq.m, q, rc, fl);
  });
  u[B.sc * (2 + D)] = q.p;
  u[B.sc * (3 + D)] = q.a;
}

__global__ void __launch_bounds__(256) copy_regions_kernel(const CopyDesc* regions, int ncomp) {
  const CopyDesc& R = regions[blockIdx.y];
  const long long cells = (long long)R.len[0] * R.len[1] * R.len[2];
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < cells;
       t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t % R.len[0]);
    const int j = (int)((t / R.len[0]) % R.len[1]);
    const int k = (int)(t / ((long long)R.len[0] * R.len[1]));
    const long long s = (R.slo[0] + i) +
                        (long long)R.sn[0] * ((R.slo[1] + j) + (long long)R.sn[1] * (R.slo[2] + k));
    const long long d = (R.dlo[0] + i) +
                        (long long)R.dn[0] * ((R.dlo[1] + j) + (long long)R.dn[1] * (R.dlo[2] + k));
    for (int c = 0; c < ncomp; ++c) R.dst[d + R.dsc * c] = R.src[s + R.ssc * c];
  }
}

// The same copy for destinations in a PEER's memory (halo exchange by direct stores over
// NVLink): remote stores are fire-and-forget, what limits them is how many bytes a thread has
// in flight.  Every thread moves four independent 16-byte pieces per round (all loads first,
// then all stores), components and rows flattened into one index so that also an 8-cell wide
// x slab (64-byte runs) keeps every lane busy.  Falls back to 8-byte pieces where the rows are
// not 16-byte aligned.
template <typename V> __device__ __forceinline__ void copy_wide(const CopyDesc& R, int ncomp) {
  constexpr int W = (int)(sizeof(V) / sizeof(double));
  const unsigned lx = (unsigned)R.len[0] / W;
  const unsigned ly = (unsigned)R.len[1];
  const unsigned per_comp = lx * ly * (unsigned)R.len[2];
  const unsigned total = per_comp * (unsigned)ncomp;
  const unsigned nthreads = gridDim.x * blockDim.x;
  for (unsigned t0 = blockIdx.x * blockDim.x + threadIdx.x; t0 < total; t0 += 4 * nthreads) {
    V v[4];
    long long d[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned t = t0 + u * nthreads;
      d[u] = -1;
      if (t < total) {
        const unsigned c = t / per_comp;
        const unsigned r = t - c * per_comp;
        const unsigned row = r / lx;
        const unsigned i = (r - row * lx) * W;
        const unsigned k = row / ly;
        const unsigned j = row - k * ly;
        const long long s = (R.slo[0] + i) + (long long)R.sn[0] * ((R.slo[1] + j) +
                                                                   (long long)R.sn[1] * (R.slo[2] + k));
        d[u] = (R.dlo[0] + i) +
               (long long)R.dn[0] * ((R.dlo[1] + j) + (long long)R.dn[1] * (R.dlo[2] + k)) +
               R.dsc * c;
        v[u] = __ldg(reinterpret_cast<const V*>(R.src + s + R.ssc * c));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (d[u] >= 0) *reinterpret_cast<V*>(R.dst + d[u]) = v[u];
    }
  }
}
__global__ void __launch_bounds__(256) copy_regions_wide_kernel(const CopyDesc* regions, int ncomp) {
  const CopyDesc& R = regions[blockIdx.y];
  const bool even = (((R.len[0] | R.slo[0] | R.dlo[0] | R.sn[0] | R.dn[0]) & 1) == 0) &&
                    ((R.ssc | R.dsc) & 1) == 0 &&
                    ((((unsigned long long)R.src) | ((unsigned long long)R.dst)) & 15) == 0;
  if (even) {
    copy_wide<double2>(R, ncomp);
  } else {
    copy_wide<double>(R, ncomp);
  }
}

__global__ void __launch_bounds__(256)
pack_kernel(const PackDesc* regions, int ncomp, double* buffer, int unpack) {
  const PackDesc& R = regions[blockIdx.y];
  const long long cells = (long long)R.len[0] * R.len[1] * R.len[2];
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < cells;
       t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t % R.len[0]);
    const int j = (int)((t / R.len[0]) % R.len[1]);
    const int k = (int)(t / ((long long)R.len[0] * R.len[1]));
    const long long a = (R.lo[0] + i) +
                        (long long)R.n[0] * ((R.lo[1] + j) + (long long)R.n[1] * (R.lo[2] + k));
    for (int c = 0; c < ncomp; ++c) {
      if (unpack) {
        R.array[a + R.sc * c] = buffer[R.offset + cells * c + t];
      } else {
        buffer[R.offset + cells * c + t] = R.array[a + R.sc * c];
      }
    }
  }
}

// Physical boundaries.
//  transmissive: src/AMReX/boundary_condition/TransmissiveBoundary.cpp:31-93
//                ghost <- the boundary cell of the same line
//  reflective:   include/fub/AMReX/boundary_condition/ReflectiveBoundary.hpp:67-94,
//                include/fub/boundary_condition/ReflectiveBoundary.hpp:57-70,
//                src/equations/PerfectGas.cpp:266-300
//                ghost <- mirror cell with m -> m - 2 (m . e_d) e_d
__global__ void __launch_bounds__(256)
boundary_kernel(const BcDesc* regions, int ncomp, int dir, int kind) {
  const BcDesc& R = regions[blockIdx.y];
  const long long cells = (long long)R.len[0] * R.len[1] * R.len[2];
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < cells;
       t += (long long)gridDim.x * blockDim.x) {
    int idx[3];
    idx[0] = R.lo[0] + (int)(t % R.len[0]);
    idx[1] = R.lo[1] + (int)((t / R.len[0]) % R.len[1]);
    idx[2] = R.lo[2] + (int)(t / ((long long)R.len[0] * R.len[1]));
    int src[3] = {idx[0], idx[1], idx[2]};
    src[dir] = (kind == 0) ? R.src_plane : (R.mirror_sum - idx[dir]);
    const long long d = idx[0] + (long long)R.n[0] * (idx[1] + (long long)R.n[1] * idx[2]);
    const long long s = src[0] + (long long)R.n[0] * (src[1] + (long long)R.n[1] * src[2]);
    for (int c = 0; c < ncomp; ++c) {
      double v = R.array[s + R.sc * c];
      if (kind == 1 && c == 1 + dir) v = v - 2 * (v * 1.0 * 1.0);
      R.array[d + R.sc * c] = v;
    }
  }
}

} // namespace

int launch_update_conservatively(int rank, int dir, const SweepArgs& A, cudaStream_t s) {
  const long long total = (long long)A.max_n[0] * A.max_n[1] * A.max_n[2];
  dim3 grid(cdiv(total, 256), A.nboxes, 1);
#define FVS_UPD(D, DIR)                                                                       \
  if (rank == D && dir == DIR) {                                                              \
    update_kernel<D, DIR><<<grid, 256, 0, s>>>(A);                                           \
    return 1;                                                                                 \
  }
  FVS_UPD(1, 0) FVS_UPD(2, 0) FVS_UPD(2, 1) FVS_UPD(3, 0) FVS_UPD(3, 1) FVS_UPD(3, 2)
#undef FVS_UPD
  return -1;
}

int launch_complete_from_cons(int rank, const SweepArgs& A, cudaStream_t s) {
  const long long total = (long long)A.max_n[0] * A.max_n[1] * A.max_n[2];
  dim3 grid(cdiv(total, 256), A.nboxes, 1);
  switch (rank) {
  case 1: complete_kernel<1><<<grid, 256, 0, s>>>(A); return 1;
  case 2: complete_kernel<2><<<grid, 256, 0, s>>>(A); return 1;
  case 3: complete_kernel<3><<<grid, 256, 0, s>>>(A); return 1;
  }
  return -1;
}

int launch_copy_regions(const CopyDesc* regions, int nregions, int ncomp, long long max_cells,
                        cudaStream_t s) {
  if (nregions <= 0) return 0;
  const unsigned gx = (unsigned)min((long long)cdiv(max_cells, 256), 4096LL);
  for (int r0 = 0; r0 < nregions; r0 += 65535) {
    dim3 grid(gx, (unsigned)min(nregions - r0, 65535), 1);
    copy_regions_kernel<<<grid, 256, 0, s>>>(regions + r0, ncomp);
  }
  return (nregions + 65534) / 65535;
}

int launch_copy_regions_wide(const CopyDesc* regions, int nregions, int ncomp, long long max_cells,
                             cudaStream_t s) {
  if (nregions <= 0) return 0;
  // enough threads for four 16-byte pieces each, at most 8 blocks per SM of a B200
  const long long pieces = (max_cells * ncomp + 1) / 2;
  const unsigned gx = (unsigned)min((long long)cdiv(pieces, 256 * 4), 148LL * 8);
  for (int r0 = 0; r0 < nregions; r0 += 65535) {
    dim3 grid(max(gx, 1u), (unsigned)min(nregions - r0, 65535), 1);
    copy_regions_wide_kernel<<<grid, 256, 0, s>>>(regions + r0, ncomp);
  }
  return (nregions + 65534) / 65535;
}

int launch_pack(const PackDesc* regions, int nregions, int ncomp, double* buffer,
                long long max_cells, int unpack, cudaStream_t s) {
  if (nregions <= 0) return 0;
  const unsigned gx = (unsigned)min((long long)cdiv(max_cells, 256), 4096LL);
  for (int r0 = 0; r0 < nregions; r0 += 65535) {
    dim3 grid(gx, (unsigned)min(nregions - r0, 65535), 1);
    pack_kernel<<<grid, 256, 0, s>>>(regions + r0, ncomp, buffer, unpack);
  }
  return (nregions + 65534) / 65535;
}

int launch_boundary(const BcDesc* regions, int nregions, int ncomp, int dir, int kind,
                    long long max_cells, cudaStream_t s) {
  if (nregions <= 0) return 0;
  const unsigned gx = (unsigned)min((long long)cdiv(max_cells, 256), 4096LL);
  dim3 grid(gx, (unsigned)nregions, 1);
  boundary_kernel<<<grid, 256, 0, s>>>(regions, ncomp, dir, kind);
  return 1;
}

} // namespace fvs
