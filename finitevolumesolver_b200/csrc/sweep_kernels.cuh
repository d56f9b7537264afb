// sweep_kernels.cuh -- the split-sweep kernels.
//
// Fused sweep (the hot path): ONE pass over the state does what the reference
// does in four (DimensionalSplitLevelIntegrator.hpp:235-256):
//   K1 ComputeStableDt      (FluxMethod.hpp:341-382 over the stored states)
//   K2 ComputeNumericFluxes (FluxMethod.hpp:246-293)
//   K3 UpdateConservatively (src/HyperbolicPatchIntegrator.cpp:35-85)
//   K4 CompleteFromCons     (CompleteFromCons.hpp:141-204)
// reading every Complete state once and writing it once (112 B per 3-D
// cell-update).  The state is read from `src` and written to `dst` (the
// context ping-pongs buffers), so any decomposition of the box is race free.
//
// Work decomposition: a software pipeline that marches along the sweep
// direction and evaluates every MUSCL-Hancock cell predictor exactly once and
// every face flux exactly once:
//   stage A (cell  j-1): slopes + half-step predictor -> edge states recL, recR
//   stage B (face  j-1): base flux(recR(j-2), recL(j-1))
//   stage C (cell  j-2): u += dt/dx (F(j-2) - F(j-1)); p, a from cons; store
//   stage D (face j-1|j): signal speeds of the STORED states -> stable dt
// * y / z sweeps: one thread per grid line; threads of a warp are adjacent in
//   x, so every load and store is a coalesced 256-byte row segment; the
//   loop-carried pipeline state lives in thread-private shared-memory slots
//   (MUSCL-Hancock) or in registers (first order, MusclHancockPrim).
// * x sweeps: one warp per grid row; lane l holds cell c0+l; a lane hands its
//   stage results to its right neighbour through per-warp shared-memory slots
//   (slot 0 carries the last lane of the previous 32-cell chunk).
#pragma once
#include <type_traits>

#include "launch.h"

namespace fvs {

template <int D> FVS_DEV void load_complete(Complete<D>& q, const double* p, long long sc) {
  q.rho = __ldg(p);
#pragma unroll
  for (int d = 0; d < D; ++d) q.mom[d] = __ldg(p + sc * (1 + d));
  q.E = __ldg(p + sc * (1 + D));
  q.p = __ldg(p + sc * (2 + D));
  q.a = __ldg(p + sc * (3 + D));
}
// prefetch.global.L1 of the D + 4 component values of one cell: hides the
// load latency of the next pipeline step without holding registers
template <int D> FVS_DEV void prefetch_complete(const double* p, long long sc) {
#pragma unroll
  for (int c = 0; c < D + 4; ++c) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p + sc * c));
  }
}
// cp.async (LDGSTS) of the D + 4 component values of one cell into a lane-private
// shared-memory column: the load of the NEXT chunk is in flight during a whole chunk of
// arithmetic and costs no registers
template <int D> FVS_DEV void stage_complete_async(double* smem_col, int stride, const double* p,
                                                   long long sc) {
#pragma unroll
  for (int c = 0; c < D + 4; ++c) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_col + c * stride);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(p + sc * c)
                 : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int D> FVS_DEV void staged_complete(Complete<D>& q, const double* smem_col, int stride) {
  q.rho = smem_col[0];
#pragma unroll
  for (int d = 0; d < D; ++d) q.mom[d] = smem_col[(1 + d) * stride];
  q.E = smem_col[(1 + D) * stride];
  q.p = smem_col[(2 + D) * stride];
  q.a = smem_col[(3 + D) * stride];
}
template <int D> FVS_DEV void store_complete(double* p, long long sc, const Complete<D>& q) {
  p[0] = q.rho;
#pragma unroll
  for (int d = 0; d < D; ++d) p[sc * (1 + d)] = q.mom[d];
  p[sc * (1 + D)] = q.E;
  p[sc * (2 + D)] = q.p;
  p[sc * (3 + D)] = q.a;
}
template <int D> FVS_DEV void copy_cons(Cons<D>& c, const Complete<D>& q) {
  c.rho = q.rho;
#pragma unroll
  for (int d = 0; d < D; ++d) c.mom[d] = q.mom[d];
  c.E = q.E;
}

// Block-wide max of a non-negative double, then one atomic per block.  The
// bit pattern of a non-negative IEEE double is monotone as an unsigned
// integer, so atomicMax on the bits is an exact max.
FVS_DEV void reduce_max_signal(double v, unsigned long long* target) {
  // NaN never enters: callers pass `v` built with `x > v ? x : v`.
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w > v ? w : v;
  }
  __shared__ double warp_max[32];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  if (lane == 0) warp_max[warp] = v;
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    v = lane < nw ? warp_max[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double w = __shfl_xor_sync(0xffffffffu, v, o);
      v = w > v ? w : v;
    }
    if (lane == 0 && v > 0.0) {
      atomicMax(target, (unsigned long long)__double_as_longlong(v));
    }
  }
}

// Pipeline stages -----------------------------------------------------------
// stage A: the two predicted edge states of the cell with neighbours qm, qp
template <int D, int DIR, int RECON, int LIM, typename S0, typename S1, typename S2>
FVS_DEV void stage_predict(const Eqs& eq, const S0& qm, const S1& q0, const S2& qp,
                           double lambda_half, Edge<D>& recL, Edge<D>& recR) {
  if constexpr (RECON == kReconMusclHancock) {
    with_fallback([&](auto M, Flag& fl) {
      muscl_predict<decltype(M), D, DIR, LIM>(eq.m, qm, q0, qp, lambda_half, recL, recR, fl);
    });
  } else if constexpr (RECON == kReconMusclHancockPrim) {
    // MusclHancockPrim uses lambda = dt / dx (MusclHancockPrimMethod.hpp:101);
    // 2 * (0.5 * dt / dx) is that value exactly
    with_fallback([&](auto M, Flag& fl) {
      muscl_prim_predict<decltype(M), D, DIR>(eq.m, qm, q0, qp, 2.0 * lambda_half, recL, recR, fl);
    });
  } else if constexpr (RECON == kReconMusclHancockChar) {
    with_fallback([&](auto M, Flag& fl) {
      muscl_char_predict<decltype(M), D, DIR>(eq.m, qm, q0, qp, 2.0 * lambda_half, recL, recR, fl);
    });
  } else if constexpr (is_mh2(RECON)) {
    with_fallback([&](auto M, Flag& fl) {
      mh2_predict<decltype(M), D, DIR, RECON>(eq.m, eq.c, qm, q0, qp, lambda_half, recL, recR, fl);
    });
  } else {
    recL.q = q0;
    recL.rc = FastMath::rcp(q0.rho);
    recR = recL;
  }
}
// the same under the math policy M with the caller's validity flag (mark and repair)
template <class M, int D, int DIR, int RECON, int LIM, typename S0, typename S1, typename S2>
FVS_DEV void predict_with(const Eqs& eq, const S0& qm, const S1& q0, const S2& qp,
                          double lambda_half, Edge<D>& recL, Edge<D>& recR, Flag& fl) {
  if constexpr (RECON == kReconMusclHancock) {
    muscl_predict<M, D, DIR, LIM>(eq.m, qm, q0, qp, lambda_half, recL, recR, fl);
  } else if constexpr (RECON == kReconMusclHancockPrim) {
    muscl_prim_predict<M, D, DIR>(eq.m, qm, q0, qp, 2.0 * lambda_half, recL, recR, fl);
  } else if constexpr (RECON == kReconMusclHancockChar) {
    muscl_char_predict<M, D, DIR>(eq.m, qm, q0, qp, 2.0 * lambda_half, recL, recR, fl);
  } else if constexpr (is_mh2(RECON)) {
    mh2_predict<M, D, DIR, RECON>(eq.m, eq.c, qm, q0, qp, lambda_half, recL, recR, fl);
  } else {
    recL.q = q0;
    recL.rc = M::rcp(q0.rho);
    recR = recL;
  }
}
// stage B: base flux through the face between two edge states
template <int D, int DIR, int BASE>
FVS_DEV void stage_flux(const Eqs& eq, Cons<D>& F, const Edge<D>& left, const Edge<D>& right,
                        int& failed) {
  with_fallback([&](auto M, Flag& fl) {
    base_flux<decltype(M), D, DIR, BASE>(eq, F, left, right, fl, failed);
  });
}
// stage C: cons_new = cons + lambda * (F_left - F_right), then p and a
// (src/HyperbolicPatchIntegrator.cpp:35-61, PerfectGas.cpp:112-124)
template <int D, typename S>
FVS_DEV void stage_update(const Eqs& eq, Complete<D>& out, const S& q, const Cons<D>& Fl,
                          const Cons<D>& Fr, double lambda, bool update) {
  if (update) {
    out.rho = q.rho + lambda * (Fl.rho - Fr.rho);
#pragma unroll
    for (int d = 0; d < D; ++d) out.mom[d] = q.mom[d] + lambda * (Fl.mom[d] - Fr.mom[d]);
    out.E = q.E + lambda * (Fl.E - Fr.E);
  } else {
    out.rho = q.rho;
#pragma unroll
    for (int d = 0; d < D; ++d) out.mom[d] = q.mom[d];
    out.E = q.E;
  }
  with_fallback([&](auto M, Flag& fl) {
    const Rcp rc = decltype(M)::rcp(out.rho);
    decltype(M)::check_window(rc, fl);
    complete_from_cons<decltype(M), D>(eq.m, out, rc, fl);
  });
}
// stage D: per-cell stable-dt data and the signal speed of a face
template <int D, int BASE>
FVS_DEV void stage_dt_cell(DtData<D, BASE>& out, const Complete<D>& q) {
  with_fallback([&](auto M, Flag& fl) { dt_prepare<decltype(M), D, BASE>(out, q, fl); });
}
template <int D, int DIR, int BASE>
FVS_DEV double stage_dt_face(const Eqs& eq, const DtData<D, BASE>& l, const DtData<D, BASE>& r,
                             int& failed) {
  double s = 0.0;
  with_fallback([&](auto M, Flag& fl) {
    s = base_max_signal<decltype(M), D, DIR, BASE>(eq, l, r, fl, failed);
  });
  return s;
}

// stage D, lazy (numerics.cuh "Lazy stable dt"): the exact signal speed of the
// face between the stored states at pl and pr is only evaluated where the
// cheap bound exceeds `thr`; the exact path reloads the two states (cache hits).
template <int BASE> constexpr bool kLazyDt = BASE != kBaseGodunovExact;
template <int D, int DIR, int BASE>
FVS_DEV void stage_dt_exact(const Eqs& eq, const double* pl, const double* pr, long long sc,
                            double& maxsig, double& thr, int& failed) {
  Complete<D> l, r;
  load_complete<D>(l, pl, sc);
  load_complete<D>(r, pr, sc);
  DtData<D, BASE> dl, dr;
  stage_dt_cell<D, BASE>(dl, l);
  stage_dt_cell<D, BASE>(dr, r);
  const double s = stage_dt_face<D, DIR, BASE>(eq, dl, dr, failed);
  if (s > maxsig) {
    maxsig = s;
    const double t = s * kLazyKeep;
    thr = t > thr ? t : thr;
  }
}

#ifndef FVS_SWEEP_THREADS
#define FVS_SWEEP_THREADS 128
#endif
// 3 resident blocks (168 registers, 12 warps per SM): measured best for the
// shared-memory pipeline (profiles/r01_tuning.md)
#ifndef FVS_SWEEP_MIN_BLOCKS
#define FVS_SWEEP_MIN_BLOCKS 3
#endif
#ifndef FVS_SWEEP_UNROLL
#define FVS_SWEEP_UNROLL 1
#endif
// tuning builds may cap the registers directly instead of through the resident-block count
#ifdef FVS_SWEEP_MAXNREG
#define FVS_SWEEP_BOUNDS __maxnreg__(FVS_SWEEP_MAXNREG)
#else
#define FVS_SWEEP_BOUNDS __launch_bounds__(kSweepThreads, FVS_SWEEP_MIN_BLOCKS)
#endif
constexpr int kSweepThreads = FVS_SWEEP_THREADS;
constexpr int kSweepUnroll = FVS_SWEEP_UNROLL;

// ---------------------------------------------------------------------------
// Fused sweep along y or z: thread per line, register pipeline.
template <int D, int DIR, int BASE, int RECON, int LIM>
__global__ void __launch_bounds__(kSweepThreads, FVS_SWEEP_MIN_BLOCKS)
fused_sweep_yz_kernel(const SweepArgs A) {
  static_assert(DIR >= 1 && DIR < D, "y/z sweep");
  const BoxDesc& B = A.boxes[blockIdx.y];
  const int n0 = B.n[0];
  const int nother = (DIR == 1) ? B.n[2] : B.n[1];
  const long long nlines = (long long)n0 * nother;
  const long long line = (long long)blockIdx.x * kSweepThreads + threadIdx.x;
  const int nd = B.n[DIR];
  const int o_first = blockIdx.z * A.seg_len;
  double maxsig = 0.0;
  int failed = 0;
  if (line < nlines && o_first < nd) {
    const Eqs eq(A.eq);
    const int o_last = min(nd, o_first + A.seg_len) - 1;
    const long long stride = (DIR == 1) ? (long long)n0 : (long long)n0 * B.n[1];
    long long base;
    if (DIR == 1) {
      const int x = (int)(line % n0);
      const int z = (int)(line / n0);
      base = x + (long long)n0 * B.n[1] * z;
    } else {
      base = line;
    }
    const double* src = B.src + base;
    double* dst = B.dst + base;
    const long long sc = B.sc;
    const int ulo = B.ulo, uhi = B.uhi;
    const unsigned char* dtmask = B.dtmask;

    Complete<D> q0, q1, qn; // cells j, j-1 and the prefetched cell j+1
    // cell j-2: the primitive-variable predictor also needs its pressure
    std::conditional_t<RECON == kReconMusclHancock || RECON == kReconFirstOrder, Cons<D>,
                       Complete<D>>
        q2;
    Edge<D> recR_prev;      // upper edge state of cell j-2
    Cons<D> F_prev;         // flux through face j-2
    DtData<D, BASE> dt_prev, dt_cur;
    const int jstart = max(o_first - 2, 0);
    load_complete<D>(q0, src + stride * jstart, sc);
    qn = q0;
    if (jstart + 1 < nd) load_complete<D>(qn, src + stride * (jstart + 1), sc);
    q1 = q0;
    if constexpr (std::is_same_v<decltype(q2), Complete<D>>) {
      q2 = q0;
    } else {
      copy_cons<D>(q2, q0);
    }
    recR_prev.q = q0;
    recR_prev.rc = FastMath::rcp(q0.rho);
    copy_cons<D>(F_prev, q0);
    stage_dt_cell<D, BASE>(dt_cur, q0);
    dt_prev = dt_cur;

    // One pipeline step.  STEADY: every stage is active (interior of the
    // update range), so A, B and C share one fast-path validity test and the
    // step has no range tests.
    const int jend = o_last + 2;
    // steady range: all of A (cell j-1), B (face j-1), C (cell j-2) active
    const int js = max(max(jstart + 3, ulo + 2), o_first + 2);
    const int je = min(min(uhi + 2, nd - 1), jend); // inclusive
#pragma unroll kSweepUnroll
    for (int j = jstart; j <= jend; ++j) {
      if (j > jstart) {
        q0 = qn;
        if (j < nd) stage_dt_cell<D, BASE>(dt_cur, q0);
      }
      if (j + 1 < nd) load_complete<D>(qn, src + stride * (j + 1), sc); // prefetch
      Edge<D> recL, recR;
      Cons<D> F;
      if (j >= js && j <= je) {
        Complete<D> out;
        with_fallback([&](auto M, Flag& fl) {
          using MM = decltype(M);
          if constexpr (RECON == kReconMusclHancock) {
            muscl_predict<MM, D, DIR, LIM>(eq.m, q2, q1, q0, A.lambda_half, recL, recR, fl);
          } else if constexpr (RECON == kReconMusclHancockPrim) {
            muscl_prim_predict<MM, D, DIR>(eq.m, q2, q1, q0, A.lambda, recL, recR, fl);
          } else if constexpr (RECON == kReconMusclHancockChar) {
            muscl_char_predict<MM, D, DIR>(eq.m, q2, q1, q0, A.lambda, recL, recR, fl);
          } else if constexpr (is_mh2(RECON)) {
            mh2_predict<MM, D, DIR, RECON>(eq.m, eq.c, q2, q1, q0, A.lambda_half, recL, recR, fl);
          } else {
            recL.q = q1;
            recL.rc = MM::rcp(q1.rho);
            recR = recL;
          }
          base_flux<MM, D, DIR, BASE>(eq, F, recR_prev, recL, fl, failed);
          out.rho = q2.rho + A.lambda * (F_prev.rho - F.rho);
#pragma unroll
          for (int d = 0; d < D; ++d) {
            out.mom[d] = q2.mom[d] + A.lambda * (F_prev.mom[d] - F.mom[d]);
          }
          out.E = q2.E + A.lambda * (F_prev.E - F.E);
          const Rcp rc = MM::rcp(out.rho);
          MM::check_window(rc, fl);
          complete_from_cons<MM, D>(eq.m, out, rc, fl);
        });
        store_complete<D>(dst + stride * (j - 2), sc, out);
      } else {
        // stage A: cell j-1
        if (j >= jstart + (RECON != kReconFirstOrder ? 2 : 1) && j - 1 >= ulo - 1 &&
            j - 1 <= uhi + 1) {
          stage_predict<D, DIR, RECON, LIM>(eq, q2, q1, q0, A.lambda_half, recL, recR);
        } else {
          recL = recR_prev;
          recR = recR_prev;
        }
        // stage B: face j-1
        if (j >= jstart + (RECON != kReconFirstOrder ? 3 : 2) && j - 1 >= ulo &&
            j - 1 <= uhi + 1) {
          stage_flux<D, DIR, BASE>(eq, F, recR_prev, recL, failed);
        } else {
          F = F_prev;
        }
        // stage C: cell j-2
        if (j - 2 >= o_first) {
          Complete<D> out;
          stage_update<D>(eq, out, q2, F_prev, F, A.lambda, j - 2 >= ulo && j - 2 <= uhi);
          store_complete<D>(dst + stride * (j - 2), sc, out);
        }
      }
      // stage D: face (j-1 | j) of the stored states
      if (j > jstart && j < nd && dtmask[j - 1]) {
        const double s = stage_dt_face<D, DIR, BASE>(eq, dt_prev, dt_cur, failed);
        maxsig = s > maxsig ? s : maxsig;
      }
      if constexpr (std::is_same_v<decltype(q2), Complete<D>>) {
        q2 = q1;
      } else {
        copy_cons<D>(q2, q1);
      }
      q1 = q0;
      recR_prev = recR;
      F_prev = F;
      dt_prev = dt_cur;
    }
  }
  reduce_max_signal(maxsig, A.max_signal);
  if (failed) *A.fail_flag = 1;
}

// ---------------------------------------------------------------------------
// Fused sweep along y or z, shared-memory pipeline: the same march as above,
// but the loop-carried pipeline state (the conservative parts of cells j-1 and
// j-2, the upper edge state of cell j-2, the flux through face j-2 and the dt
// data of cell j-1) lives in thread-private shared-memory slots instead of
// registers, and the stages are separated by compiler barriers.  The register
// budget is then the largest single stage, not the whole pipeline, which
// doubles the resident warps per SM; the kernel is latency bound on dependent
// fp64 chains (profiles/), so occupancy is what buys throughput.  Slots are
// padded to an odd number of doubles: a warp's 64-bit accesses are then
// bank-conflict free.
template <typename T, bool kOdd = (sizeof(T) / 8) % 2 == 1> struct Slot {
  T v;
};
template <typename T> struct Slot<T, false> {
  T v;
  double pad;
};
#ifdef FVS_NO_STAGE_BARRIER
#define FVS_STAGE_BARRIER()
#else
#define FVS_STAGE_BARRIER() asm volatile("" ::: "memory")
#endif

// 0: none, 1: cell j+1 held in registers, 2: prefetch.global.L1 of cell
// j + FVS_SWEEP_PREFETCH_DIST
// 3 = cp.async staging of cell j + 1 through a thread-private shared-memory column (no
// registers held across the step; measured best for the bit-exact arithmetic: 6.67 -> 6.47 ms
// per 512^3 sweep together with the x kernel's staging), 1 = register prefetch (measured
// best for the tolerance arithmetic), 2 = prefetch.global.L1
#ifndef FVS_SWEEP_PREFETCH
#ifdef FVS_TOLERANCE_UNIT
#define FVS_SWEEP_PREFETCH 1
#else
#define FVS_SWEEP_PREFETCH 3
#endif
#endif
#ifndef FVS_SWEEP_PREFETCH_DIST
#define FVS_SWEEP_PREFETCH_DIST 1
#endif

// Stages A, B, C of one steady step of the y/z march under the math policy MM:
// predictor of cell j-1 (from cells j-2, j-1, j), flux through face j-1, update of cell j-2.
template <class MM, int D, int DIR, int BASE, int RECON, int LIM, typename QT, typename Q0>
FVS_DEV void steady_stages(const Eqs& eq, const QT& q2, const QT& q1, const Q0& q0,
                           const Edge<D>& recR_prev, Edge<D>& recR_out, const Cons<D>& F_prev,
                           Cons<D>& F, Complete<D>& out, double lambda, double lambda_half,
                           Flag& fl, int& failed) {
  Edge<D> recL;
  {
    Edge<D> recR;
    if constexpr (RECON == kReconMusclHancock) {
      muscl_predict<MM, D, DIR, LIM>(eq.m, q2, q1, q0, lambda_half, recL, recR, fl);
    } else {
      recL.q = q1;
      recL.rc = MM::rcp(q1.rho);
      recR = recL;
    }
    recR_out = recR;
  }
  FVS_STAGE_BARRIER();
  base_flux<MM, D, DIR, BASE>(eq, F, recR_prev, recL, fl, failed);
  FVS_STAGE_BARRIER();
  out.rho = q2.rho + lambda * (F_prev.rho - F.rho);
#pragma unroll
  for (int d = 0; d < D; ++d) {
    out.mom[d] = q2.mom[d] + lambda * (F_prev.mom[d] - F.mom[d]);
  }
  out.E = q2.E + lambda * (F_prev.E - F.E);
  const Rcp rc = MM::rcp(out.rho);
  MM::check_window(rc, fl);
  complete_from_cons<MM, D>(eq.m, out, rc, fl);
}

// Mark and repair.  The marches never branch on the validity of their fast paths: a warp that
// did would have to drain the longest dependency chain of the step (update -> 1 / rho -> p -> a)
// before it could issue anything else, once per cell.  Instead every steady step stores its
// results unconditionally and a step whose fast paths left their domain (vacuum, tiny or
// denormal momenta, ...) is entered into the sweep's repair list (one atomic, predicated).
// Such a step j can only have spoiled
//   the updated cell j-2, the flux through face j-1 (cells j-2, j-1) and the upper edge state of
//   cell j-1 (flux through face j, cells j-1 and j),
// everything later is computed from the cell states again -- the pipeline heals itself.  The
// repair pass that follows every fused sweep on the stream (repair_kernel; a handful of blocks
// that return at once when the list is empty) recomputes cells j-2 .. j of the listed steps
// from global memory (`src` is never written by the sweep) with the checked, unfused stage
// functions.  A list longer than its capacity means "every cell of the sweep".  Keeping the
// recompute out of the sweep kernels altogether matters: as an out-of-line call behind the
// march it made the x sweep spill 112 B at 128 registers (profiles/r02_tuning.md).
template <int D, int DIR, int BASE, int RECON, int LIM>
FVS_DEV int repair_cells(const Eqs& eq, const double* src, double* dst, long long stride,
                         long long sc, int nd, int c_lo, int c_hi, int ulo, int uhi,
                         double lambda, double lambda_half) {
  int failed = 0;
#pragma unroll 1
  for (int c = c_lo; c <= c_hi; ++c) {
    Complete<D> q[5]; // cells c-2 .. c+2
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      load_complete<D>(q[k], src + stride * min(max(c - 2 + k, 0), nd - 1), sc);
    }
    Edge<D> eL[3], eR[3]; // edge states of cells c-1, c, c+1
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      stage_predict<D, DIR, RECON, LIM>(eq, q[k], q[k + 1], q[k + 2], lambda_half, eL[k], eR[k]);
    }
    Cons<D> Flo, Fhi;
    stage_flux<D, DIR, BASE>(eq, Flo, eR[0], eL[1], failed);
    stage_flux<D, DIR, BASE>(eq, Fhi, eR[1], eL[2], failed);
    Complete<D> out;
    stage_update<D>(eq, out, q[2], Flo, Fhi, lambda, c >= ulo && c <= uhi);
    store_complete<D>(dst + stride * c, sc, out);
  }
  return failed;
}

// enters cells lo..hi of the line at element offset `line` of box `box` into the repair list
FVS_DEV void mark_for_repair(const SweepArgs& A, int box, long long line, int lo, int hi) {
  if (lo > hi) return;
  const unsigned i = atomicAdd(A.repair_count, 1u);
  if (i < A.repair_cap) A.repair_list[i] = RepairEntry{line, box, lo, hi};
}

// The marches collect the marks of up to 32 steps in a register (no branch per step: the flag of
// a step is the last thing its dependency chains produce) and enter them here in one go; bit b
// stands for step j0 + b * jstep.
FVS_DEV void flush_marks(const SweepArgs& A, unsigned mask, int box, long long line, int j0,
                         int jstep, int clamp_lo, int clamp_hi) {
  while (mask) {
    const int b = __ffs((int)mask) - 1;
    mask &= mask - 1;
    const int j = j0 + b * jstep;
    mark_for_repair(A, box, line, max(j - 2, clamp_lo), min(j, clamp_hi));
  }
}

constexpr int kRepairBlocks = 148;
template <int D, int DIR, int BASE, int RECON, int LIM>
__global__ void __launch_bounds__(128) repair_kernel(const SweepArgs A) {
  const unsigned n = *reinterpret_cast<volatile unsigned*>(A.repair_count);
  if (n == 0) return; // the same on every block: nothing else writes the counter now
  const Eqs eq(A.eq);
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned nthreads = gridDim.x * blockDim.x;
  int failed = 0;
  if (n <= A.repair_cap) {
    for (unsigned i = tid; i < n; i += nthreads) {
      const RepairEntry e = A.repair_list[i];
      const BoxDesc& B = A.boxes[e.box];
      const long long stride =
          DIR == 0 ? 1 : DIR == 1 ? (long long)B.n[0] : (long long)B.n[0] * B.n[1];
      failed |= repair_cells<D, DIR, BASE, RECON, LIM>(eq, B.src + e.line, B.dst + e.line, stride,
                                                       B.sc, B.n[DIR], e.lo, e.hi, B.ulo, B.uhi,
                                                       A.lambda, A.lambda_half);
    }
  } else {
    // the list overflowed: every cell of every line of every box
    for (int b = 0; b < A.nboxes; ++b) {
      const BoxDesc& B = A.boxes[b];
      const int n0 = B.n[0], n1 = B.n[1], n2 = B.n[2];
      const long long stride = DIR == 0 ? 1 : DIR == 1 ? (long long)n0 : (long long)n0 * n1;
      const long long nlines = DIR == 0 ? (long long)n1 * n2 : DIR == 1 ? (long long)n0 * n2
                                                                        : (long long)n0 * n1;
      for (long long l = tid; l < nlines; l += nthreads) {
        long long line;
        if (DIR == 0) {
          line = l * n0;
        } else if (DIR == 1) {
          line = (l % n0) + (long long)n0 * n1 * (l / n0);
        } else {
          line = l;
        }
        failed |= repair_cells<D, DIR, BASE, RECON, LIM>(eq, B.src + line, B.dst + line, stride,
                                                         B.sc, B.n[DIR], 0, B.n[DIR] - 1, B.ulo,
                                                         B.uhi, A.lambda, A.lambda_half);
      }
    }
  }
  if (failed) *A.fail_flag = 1;
  // the last block to finish empties the list for the next sweep
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicInc(A.repair_count + 1, gridDim.x - 1) == gridDim.x - 1) *A.repair_count = 0u;
  }
}

template <int D, int DIR, int BASE, int RECON, int LIM>
__global__ void FVS_SWEEP_BOUNDS
fused_sweep_yz_smem_kernel(const SweepArgs A) {
  static_assert(DIR >= 1 && DIR < D, "y/z sweep");
  // first order needs the stored p and a of the neighbour, MUSCL only its
  // conservative part
  using QT = std::conditional_t<RECON == kReconMusclHancock, Cons<D>, Complete<D>>;
  __shared__ Slot<QT> s_q[3][kSweepThreads];        // ring: slot j % 3 holds cell j
  __shared__ Slot<Edge<D>> s_rec[2][kSweepThreads]; // ring: slot j & 1 holds recR of cell j-1
  __shared__ Slot<Cons<D>> s_F[kSweepThreads];      // flux through face j-2
  // eager stable dt (exact Riemann solver): dt data of cell j-1
  __shared__ Slot<DtData<D, BASE>> s_dt[kLazyDt<BASE> ? 1 : kSweepThreads];
  // the eager stable dt (exact Riemann solver) needs the shared memory for its dt data
  constexpr int kPrefetch = (FVS_SWEEP_PREFETCH == 3 && !kLazyDt<BASE>) ? 1 : FVS_SWEEP_PREFETCH;
  constexpr bool kStaged = kPrefetch == 3;
  // cell j+1, staged by cp.async while cell j is being worked on: [component][thread].  One
  // buffer (static shared memory ends at 48 KB): the slot is thread-private, the thread has read
  // cell j out of it (LDS, in program order through the same load/store unit) before it
  // issues the copy of cell j+1, whose write happens a global-memory latency later
  __shared__ double s_in[kStaged ? D + 4 : 1][kStaged ? kSweepThreads : 1];
  const BoxDesc& B = A.boxes[blockIdx.y];
  const int n0 = B.n[0];
  const long long nlines = (long long)n0 * ((DIR == 1) ? B.n[2] : B.n[1]);
  const long long line = (long long)blockIdx.x * kSweepThreads + threadIdx.x;
  const int nd = B.n[DIR];
  const int o_first = blockIdx.z * A.seg_len;
  const int tid = threadIdx.x;
  double maxsig = 0.0;
  int failed = 0;
  if (line < nlines && o_first < nd) {
    const Eqs eq(A.eq);
    const int o_last = min(nd, o_first + A.seg_len) - 1;
    const long long stride = (DIR == 1) ? (long long)n0 : (long long)n0 * B.n[1];
    long long base;
    if (DIR == 1) {
      const int x = (int)(line % n0);
      const int z = (int)(line / n0);
      base = x + (long long)n0 * B.n[1] * z;
    } else {
      base = line;
    }
    const double* src = B.src + base;
    double* dst = B.dst + base;
    const long long sc = B.sc;
    const int ulo = B.ulo, uhi = B.uhi;
    const unsigned char* dtmask = B.dtmask;

    Complete<D> q0; // cell j
    Complete<D> qn; // prefetched cell j+1 (kPrefetch == 1 only)
    DtData<D, BASE> dt_cur;
    SigBound sb_prev{0, 0}, sb_cur{0, 0};
    double thr = A.sig_hint * kLazyKeep;
    const int jstart = max(o_first - 2, 0);
    load_complete<D>(q0, src + stride * jstart, sc);
    if constexpr (kPrefetch == 1) {
      qn = q0;
      if (jstart + 1 < nd) load_complete<D>(qn, src + stride * (jstart + 1), sc);
    } // staged: the first pass of the loop stages cell jstart + 1
    {
      Cons<D> c0;
      copy_cons<D>(c0, q0);
      QT r0;
      if constexpr (RECON == kReconMusclHancock) {
        r0 = c0;
      } else {
        r0 = q0;
      }
      s_q[0][tid].v = r0;
      s_q[1][tid].v = r0;
      s_q[2][tid].v = r0;
      Edge<D> e0;
      e0.q = q0;
      e0.rc = FastMath::rcp(q0.rho);
      s_rec[0][tid].v = e0;
      s_rec[1][tid].v = e0;
      s_F[tid].v = c0;
      if constexpr (kLazyDt<BASE>) {
        sb_cur = sig_bound<D, DIR>(eq.m, q0);
      } else {
        stage_dt_cell<D, BASE>(dt_cur, q0);
        s_dt[tid].v = dt_cur;
      }
    }
    const int jend = o_last + 2;
    const int js = max(max(jstart + 3, ulo + 2), o_first + 2);
    const int je = min(min(uhi + 2, nd - 1), jend); // inclusive
    int m3 = jstart % 3; // j % 3
    unsigned marks = 0;  // bit j - jg: steady step j of the current group left the fast-path domain
    // groups of 32 steps: the marks of a group are entered into the repair list between the
    // groups, the march itself has no branch on them
#pragma unroll 1
    for (int jg = jstart; jg <= jend; jg += 32) {
    const int jg_end = min(jg + 31, jend);
#pragma unroll 1
    for (int j = jg; j <= jg_end; ++j) {
      if (j > jstart) {
        if constexpr (kPrefetch == 1) {
          q0 = qn;
        } else if constexpr (kStaged) {
          if (j < nd) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            staged_complete<D>(q0, &s_in[0][tid], kSweepThreads);
          }
        } else {
          if (j < nd) load_complete<D>(q0, src + stride * j, sc);
        }
        if (j < nd) {
          if constexpr (kLazyDt<BASE>) {
            sb_prev = sb_cur;
            sb_cur = sig_bound<D, DIR>(eq.m, q0);
          } else {
            stage_dt_cell<D, BASE>(dt_cur, q0);
          }
        }
      }
      if constexpr (kPrefetch == 1) {
        if (j + 1 < nd) load_complete<D>(qn, src + stride * (j + 1), sc); // prefetch
      } else if constexpr (kPrefetch == 2) {
        if (j + FVS_SWEEP_PREFETCH_DIST < nd) {
          prefetch_complete<D>(src + stride * (j + FVS_SWEEP_PREFETCH_DIST), sc);
        }
      } else if constexpr (kStaged) {
        if (j + 1 < nd) {
          stage_complete_async<D>(&s_in[0][tid], kSweepThreads, src + stride * (j + 1), sc);
        }
      }
      const int i1 = m3 == 0 ? 2 : m3 - 1;  // (j - 1) % 3
      const int i2 = i1 == 0 ? 2 : i1 - 1;  // (j - 2) % 3
      const int cr = j & 1;
      QT& q1 = s_q[i1][tid].v;
      QT& q2 = s_q[i2][tid].v;
      Edge<D>& recR_prev = s_rec[cr ^ 1][tid].v;
      Edge<D>& recR_out = s_rec[cr][tid].v;
      Cons<D>& F_prev = s_F[tid].v;
      Edge<D> recL;
      Cons<D> F;
      const bool steady = j >= js && j <= je;
      if (steady) {
        // steady state: stages A, B, C, fast paths only; results are stored unconditionally and
        // the step is marked if a fast path left its domain (repair_marked_steps)
        Complete<D> out;
        Flag fl;
        int newton = 0;
        steady_stages<FastPolicy, D, DIR, BASE, RECON, LIM>(eq, q2, q1, q0, recR_prev, recR_out,
                                                            F_prev, F, out, A.lambda,
                                                            A.lambda_half, fl, newton);
        const bool bad = FastPolicy::kFast && fl.any();
        marks |= (unsigned)bad << (j - jg);
        failed |= bad ? 0 : newton;
        store_complete<D>(dst + stride * (j - 2), sc, out);
      } else {
        // stage A: cell j-1
        if (j >= jstart + (RECON == kReconMusclHancock ? 2 : 1) && j - 1 >= ulo - 1 &&
            j - 1 <= uhi + 1) {
          Edge<D> recR;
          if constexpr (RECON == kReconMusclHancock) {
            stage_predict<D, DIR, RECON, LIM>(eq, q2, q1, q0, A.lambda_half, recL, recR);
          } else {
            recL.q = q1;
            recL.rc = FastMath::rcp(q1.rho);
            recR = recL;
          }
          recR_out = recR;
        } else {
          recL = recR_prev;
          recR_out = recR_prev;
        }
        FVS_STAGE_BARRIER();
        // stage B: face j-1
        if (j >= jstart + (RECON == kReconMusclHancock ? 3 : 2) && j - 1 >= ulo &&
            j - 1 <= uhi + 1) {
          stage_flux<D, DIR, BASE>(eq, F, recR_prev, recL, failed);
        } else {
          F = F_prev;
        }
        FVS_STAGE_BARRIER();
        // stage C: cell j-2
        if (j - 2 >= o_first) {
          Complete<D> out;
          stage_update<D>(eq, out, q2, F_prev, F, A.lambda, j - 2 >= ulo && j - 2 <= uhi);
          store_complete<D>(dst + stride * (j - 2), sc, out);
        }
      }
      FVS_STAGE_BARRIER();
      // stage D: face (j-1 | j) of the stored states
      if (j > jstart && j < nd && dtmask[j - 1]) {
        if constexpr (kLazyDt<BASE>) {
          if (sig_bound_exceeds(sb_prev.u, sb_prev.c, sb_cur.u, sb_cur.c, thr)) {
            stage_dt_exact<D, DIR, BASE>(eq, src + stride * (j - 1), src + stride * j, sc, maxsig,
                                         thr, failed);
          }
        } else {
          const double s = stage_dt_face<D, DIR, BASE>(eq, s_dt[tid].v, dt_cur, failed);
          maxsig = s > maxsig ? s : maxsig;
        }
      }
      // rotate: slot j % 3 <- cell j, flux and dt slots
      if constexpr (RECON == kReconMusclHancock) {
        copy_cons<D>(s_q[m3][tid].v, q0);
      } else {
        s_q[m3][tid].v = q0;
      }
      F_prev = F;
      if constexpr (!kLazyDt<BASE>) s_dt[tid].v = dt_cur;
      FVS_STAGE_BARRIER();
      m3 = m3 == 2 ? 0 : m3 + 1;
    }
    if (marks) {
      flush_marks(A, marks, blockIdx.y, base, jg, 1, o_first, min(uhi, o_last));
      marks = 0;
    }
    }
  }
  reduce_max_signal(maxsig, A.max_signal);
  if (failed) *A.fail_flag = 1;
}

// ---------------------------------------------------------------------------
// Fused sweep along x, shared-memory exchange: warp per row, lane = cell, like
// fused_sweep_x_kernel, but a lane hands its results to its right neighbour
// through per-warp shared-memory slots instead of warp shuffles, and the
// stages are separated by __syncwarp().  Slot i + 1 of a stage array belongs to
// lane i, slot 0 carries lane 31 of the previous 32-cell chunk (two carry slots
// for the cell states): no shuffle / select carry logic, and -- as in the y/z
// kernel -- a register budget of one stage instead of the whole pipeline.
#ifndef FVS_SWEEP_XS_MIN_BLOCKS
#define FVS_SWEEP_XS_MIN_BLOCKS 4
#endif
// 3 = the next chunk's cells staged by cp.async (hides the global-load latency that was
// 12.6 % of the kernel's stall samples), 2 = held in registers one chunk ahead (tolerance
// units, build.py), 1 = prefetch.global.L1
#ifndef FVS_SWEEP_XS_PREFETCH
#define FVS_SWEEP_XS_PREFETCH 3
#endif
template <int D, int BASE, int RECON, int LIM>
__global__ void __launch_bounds__(kSweepThreads, FVS_SWEEP_XS_MIN_BLOCKS)
fused_sweep_x_smem_kernel(const SweepArgs A) {
  constexpr int DIR = 0;
  constexpr int kWarps = kSweepThreads / 32;
  using QT = std::conditional_t<RECON == kReconMusclHancock, Cons<D>, Complete<D>>;
  __shared__ Slot<QT> s_q[kWarps][34];              // slot i + 2 <- cell of lane i
  __shared__ Slot<Edge<D>> s_rec[kWarps][33];       // slot i + 1 <- recR of lane i (cell c - 1)
  __shared__ Slot<Cons<D>> s_F[kWarps][33];         // slot i + 1 <- flux of lane i (face c - 1)
  // eager stable dt (exact Riemann solver): slot i + 1 <- dt data of lane i's cell
  __shared__ Slot<DtData<D, BASE>> s_dt[kLazyDt<BASE> ? 1 : kWarps][kLazyDt<BASE> ? 1 : 33];
#if FVS_SWEEP_XS_PREFETCH == 3
  // next chunk's cells, staged by cp.async: [warp][buffer][component][lane]
  __shared__ double s_in[kWarps][2][D + 4][32];
#endif
  const BoxDesc& B = A.boxes[blockIdx.y];
  const int n0 = B.n[0];
  const long long nrows = (long long)B.n[1] * B.n[2];
  const int nseg = (n0 + A.seg_len - 1) / A.seg_len;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long w = (long long)blockIdx.x * kWarps + warp;
  double maxsig = 0.0;
  int failed = 0;
  if (w < nrows * nseg) {
    const Eqs eq(A.eq);
    const long long row = w / nseg;
    const int seg = (int)(w % nseg);
    const int o_first = seg * A.seg_len;
    const int o_last = min(n0, o_first + A.seg_len) - 1;
    const double* src = B.src + row * n0;
    double* dst = B.dst + row * n0;
    const long long sc = B.sc;
    const int ulo = B.ulo, uhi = B.uhi;
    const unsigned char* dtmask = B.dtmask;
    bool first = true;
    unsigned marks = 0; // bit b: this lane's step of chunk b of the group left the fast-path domain
    double thr = A.sig_hint * kLazyKeep;
    int carry_u = 0, carry_c = 0; // signal bound of lane 31 of the previous chunk
#if FVS_SWEEP_XS_PREFETCH == 2
    // the next chunk's cell is loaded into registers one chunk ahead
    Complete<D> qn;
    load_complete<D>(qn, src + min(max(o_first - 2 + lane, 0), n0 - 1), sc);
#endif
#if FVS_SWEEP_XS_PREFETCH == 3
    int stage_buf = 0;
    stage_complete_async<D>(&s_in[warp][0][0][lane], 32,
                            src + min(max(o_first - 2 + lane, 0), n0 - 1), sc);
    // the stable-dt mask byte of the face (c-1 | c), fetched one chunk ahead as well
    unsigned char mask_next = dtmask[min(max(o_first - 3 + lane, 0), n0 - 1)];
#endif
    // groups of 32 chunks: marks entered into the repair list between the groups
#pragma unroll 1
    for (int cg = o_first - 2; cg - 2 <= o_last; cg += 32 * 32) {
    int chunk = 0;
#pragma unroll 1
    for (int c0 = cg; c0 - 2 <= o_last && chunk < 32; c0 += 32, ++chunk) {
      const int c = c0 + lane;
      const int cl = min(max(c, 0), n0 - 1);
      // ---- load: cell c, its dt data
      Complete<D> q0;
#if FVS_SWEEP_XS_PREFETCH == 2
      q0 = qn;
      (void)cl;
      if (c0 + 32 - 2 <= o_last) load_complete<D>(qn, src + min(c + 32, n0 - 1), sc);
#elif FVS_SWEEP_XS_PREFETCH == 3
      (void)cl;
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      staged_complete<D>(q0, &s_in[warp][stage_buf][0][lane], 32);
      const unsigned char mask_cur = mask_next;
      if (c0 + 32 - 2 <= o_last) {
        stage_buf ^= 1;
        stage_complete_async<D>(&s_in[warp][stage_buf][0][lane], 32, src + min(c + 32, n0 - 1), sc);
        mask_next = dtmask[min(max(c + 31, 0), n0 - 1)];
      }
#else
      load_complete<D>(q0, src + cl, sc);
#endif
#if FVS_SWEEP_XS_PREFETCH == 1
      if (c0 + 32 - 2 <= o_last) prefetch_complete<D>(src + min(c + 32, n0 - 1), sc);
#endif
      DtData<D, BASE> dt_cur;
      SigBound sb{0, 0};
      if constexpr (kLazyDt<BASE>) {
        sb = sig_bound<D, DIR>(eq.m, q0);
      } else {
        stage_dt_cell<D, BASE>(dt_cur, q0);
      }
      {
        QT mine;
        if constexpr (RECON == kReconMusclHancock) {
          copy_cons<D>(mine, q0);
        } else {
          mine = q0;
        }
        s_q[warp][lane + 2].v = mine;
        if constexpr (!kLazyDt<BASE>) s_dt[warp][lane + 1].v = dt_cur;
        if (first && lane == 0) {
          // no carry yet: lanes 0..3 produce no output in the first chunk; give
          // their missing neighbours real data to keep the arithmetic benign
          s_q[warp][0].v = mine;
          s_q[warp][1].v = mine;
          if constexpr (!kLazyDt<BASE>) s_dt[warp][0].v = dt_cur;
        }
      }
      __syncwarp();
      // Stages A to C run the branch-free fast paths only and share one validity flag; a lane
      // whose flag is not clean enters its cell into the repair list (mark and repair, see
      // repair_cells: the spoiled cells are c-2 .. c here as well)
      Flag fl;
      int newton = 0;
      // ---- stage A: cell c-1 from cells c-2, c-1, c
      Edge<D> recL;
      {
        Edge<D> recR;
        predict_with<FastPolicy, D, DIR, RECON, LIM>(eq, s_q[warp][lane].v, s_q[warp][lane + 1].v,
                                                     q0, A.lambda_half, recL, recR, fl);
        s_rec[warp][lane + 1].v = recR;
        if (first && lane == 0) s_rec[warp][0].v = recR;
      }
      __syncwarp();
      // ---- stage B: face c-1 = (c-2 | c-1)
      Cons<D> F;
      base_flux<FastPolicy, D, DIR, BASE>(eq, F, s_rec[warp][lane].v, recL, fl, newton);
      s_F[warp][lane + 1].v = F;
      if (first && lane == 0) s_F[warp][0].v = F;
      __syncwarp();
      // ---- stage C: cell c-2
      const int o = c - 2;
      if ((!first || lane >= 4) && o >= o_first && o <= o_last) {
        Complete<D> out;
        const QT& qo = s_q[warp][lane].v;
        const Cons<D>& Fl = s_F[warp][lane].v;
        if (o >= ulo && o <= uhi) {
          out.rho = qo.rho + A.lambda * (Fl.rho - F.rho);
#pragma unroll
          for (int d = 0; d < D; ++d) out.mom[d] = qo.mom[d] + A.lambda * (Fl.mom[d] - F.mom[d]);
          out.E = qo.E + A.lambda * (Fl.E - F.E);
        } else {
          out.rho = qo.rho;
#pragma unroll
          for (int d = 0; d < D; ++d) out.mom[d] = qo.mom[d];
          out.E = qo.E;
        }
        const Rcp rc = FastPolicy::rcp(out.rho);
        FastPolicy::check_window(rc, fl);
        complete_from_cons<FastPolicy, D>(eq.m, out, rc, fl);
        store_complete<D>(dst + o, sc, out);
      }
      {
        const bool bad = FastPolicy::kFast && fl.any();
        marks |= (unsigned)bad << chunk;
        failed |= bad ? 0 : newton;
      }
      // ---- stage D: face (c-1 | c) of the stored states
      if constexpr (kLazyDt<BASE>) {
        int lu = __shfl_up_sync(0xffffffffu, sb.u, 1);
        int lc = __shfl_up_sync(0xffffffffu, sb.c, 1);
        if (lane == 0) {
          lu = carry_u;
          lc = carry_c;
        }
        carry_u = __shfl_sync(0xffffffffu, sb.u, 31);
        carry_c = __shfl_sync(0xffffffffu, sb.c, 31);
#if FVS_SWEEP_XS_PREFETCH == 3
        const bool in_dt = mask_cur != 0;
#else
        const bool in_dt = c - 1 >= 0 && c < n0 && dtmask[c - 1];
#endif
        if ((!first || lane >= 1) && c - 1 >= 0 && c < n0 && in_dt) {
          if (sig_bound_exceeds(lu, lc, sb.u, sb.c, thr)) {
            stage_dt_exact<D, DIR, BASE>(eq, src + (c - 1), src + c, sc, maxsig, thr, failed);
          }
        }
      } else {
#if FVS_SWEEP_XS_PREFETCH == 3
        const bool in_dt = mask_cur != 0;
#else
        const bool in_dt = c - 1 >= 0 && c < n0 && dtmask[c - 1];
#endif
        if ((!first || lane >= 1) && c - 1 >= 0 && c < n0 && in_dt) {
          const double sgn = stage_dt_face<D, DIR, BASE>(eq, s_dt[warp][lane].v, dt_cur, failed);
          maxsig = sgn > maxsig ? sgn : maxsig;
        }
      }
      __syncwarp();
      // ---- carry of lanes 30, 31 into slots 0, 1 for the next chunk
      if (lane >= 30) s_q[warp][lane - 30].v = s_q[warp][lane + 2].v;
      if (lane == 31) {
        s_rec[warp][0].v = s_rec[warp][32].v;
        s_F[warp][0].v = F;
        if constexpr (!kLazyDt<BASE>) s_dt[warp][0].v = dt_cur;
      }
      __syncwarp();
      first = false;
    }
    if (marks) {
      // bit b: this lane's cell of chunk b of the group
      flush_marks(A, marks, blockIdx.y, row * n0, cg + lane, 32, o_first, o_last);
      marks = 0;
    }
    }
  }
  reduce_max_signal(maxsig, A.max_signal);
  if (failed) *A.fail_flag = 1;
}

// ---------------------------------------------------------------------------
// Unfused policies (thread per face / cell): the per-policy drop-in behind
// fub::HyperbolicMethod's three members.

// K2  FluxMethod<Base>::ComputeNumericFluxes, FluxMethod.hpp:246-293
template <int D, int DIR, int BASE, int RECON, int LIM>
__global__ void __launch_bounds__(128) numeric_flux_kernel(const SweepArgs A) {
  const BoxDesc& B = A.boxes[blockIdx.y];
  // faces: along DIR flo..fhi, other dims the whole state array
  int len[3] = {B.n[0], B.n[1], B.n[2]};
  len[DIR] = B.fhi - B.flo + 1;
  const long long total = (long long)len[0] * len[1] * len[2];
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int failed = 0;
  if (len[DIR] > 0 && t < total) {
    int idx[3];
    idx[0] = (int)(t % len[0]);
    idx[1] = (int)((t / len[0]) % len[1]);
    idx[2] = (int)(t / ((long long)len[0] * len[1]));
    idx[DIR] += B.flo;
    const long long stride = DIR == 0 ? 1 : (DIR == 1 ? (long long)B.n[0] : (long long)B.n[0] * B.n[1]);
    const long long cell = idx[0] + (long long)B.n[0] * (idx[1] + (long long)B.n[1] * idx[2]);
    const double* p = B.src + cell; // cell idx = right neighbour of the face
    const Eqs eq(A.eq);
    Cons<D> F;
    Edge<D> unused, left, right;
    if constexpr (RECON != kReconFirstOrder) {
      Complete<D> s0, s1, s2, s3;
      load_complete<D>(s0, p - 2 * stride, B.sc);
      load_complete<D>(s1, p - stride, B.sc);
      load_complete<D>(s2, p, B.sc);
      load_complete<D>(s3, p + stride, B.sc);
      stage_predict<D, DIR, RECON, LIM>(eq, s0, s1, s2, A.lambda_half, unused, left);
      stage_predict<D, DIR, RECON, LIM>(eq, s1, s2, s3, A.lambda_half, right, unused);
    } else {
      Complete<D> s0, s1;
      load_complete<D>(s0, p - stride, B.sc);
      load_complete<D>(s1, p, B.sc);
      stage_predict<D, DIR, RECON, LIM>(eq, s0, s0, s0, A.lambda_half, unused, left);
      stage_predict<D, DIR, RECON, LIM>(eq, s1, s1, s1, A.lambda_half, right, unused);
    }
    stage_flux<D, DIR, BASE>(eq, F, left, right, failed);
    const long long fcell = (idx[0] - B.foff[0]) +
                            (long long)B.fn[0] * ((idx[1] - B.foff[1]) +
                                                  (long long)B.fn[1] * (idx[2] - B.foff[2]));
    double* f = B.flux + fcell;
    f[0] = F.rho;
#pragma unroll
    for (int d = 0; d < D; ++d) f[B.fsc * (1 + d)] = F.mom[d];
    f[B.fsc * (1 + D)] = F.E;
  }
  if (failed) *A.fail_flag = 1;
}

// K1  FluxMethodAdapter::ComputeStableDt, FluxMethodAdapter.hpp:69-114
template <int D, int DIR, int BASE>
__global__ void __launch_bounds__(256) stable_dt_kernel(const SweepArgs A) {
  const BoxDesc& B = A.boxes[blockIdx.y];
  const long long total = (long long)B.n[0] * B.n[1] * B.n[2];
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double maxsig = 0.0;
  int failed = 0;
  if (t < total) {
    int idx[3];
    idx[0] = (int)(t % B.n[0]);
    idx[1] = (int)((t / B.n[0]) % B.n[1]);
    idx[2] = (int)(t / ((long long)B.n[0] * B.n[1]));
    if (idx[DIR] + 1 < B.n[DIR] && B.dtmask[idx[DIR]]) {
      const long long stride = DIR == 0 ? 1 : (DIR == 1 ? (long long)B.n[0] : (long long)B.n[0] * B.n[1]);
      const Eqs eq(A.eq);
      Complete<D> l, r;
      load_complete<D>(l, B.src + t, B.sc);
      load_complete<D>(r, B.src + t + stride, B.sc);
      DtData<D, BASE> dl, dr;
      stage_dt_cell<D, BASE>(dl, l);
      stage_dt_cell<D, BASE>(dr, r);
      maxsig = stage_dt_face<D, DIR, BASE>(eq, dl, dr, failed);
      maxsig = maxsig > 0.0 ? maxsig : 0.0; // drops NaN
    }
  }
  reduce_max_signal(maxsig, A.max_signal);
  if (failed) *A.fail_flag = 1;
}

// stable dt in all directions, one pass (DimensionalSplitLevelIntegrator.hpp:204-219).
// A thread visits kDtAllCellsPerThread cells (block-strided, so a warp still reads
// consecutive cells) before the three block reductions: with one cell per thread
// their barriers were 15 % of the kernel's stall samples.
#ifndef FVS_DT_ALL_CELLS
#define FVS_DT_ALL_CELLS 4
#endif
constexpr int kDtAllCellsPerThread = FVS_DT_ALL_CELLS;
template <int D, int BASE>
__global__ void __launch_bounds__(256) stable_dt_all_kernel(const DtAllArgs A) {
  const BoxDesc& B = A.boxes[0][blockIdx.y];
  const long long total = (long long)B.n[0] * B.n[1] * B.n[2];
  double maxsig[D];
#pragma unroll
  for (int d = 0; d < D; ++d) maxsig[d] = 0.0;
  int failed = 0;
  const Eqs eq(A.eq);
#pragma unroll 1
  for (int k = 0; k < kDtAllCellsPerThread; ++k) {
    const long long t =
        ((long long)blockIdx.x * kDtAllCellsPerThread + k) * blockDim.x + threadIdx.x;
    if (t >= total) break;
    int idx[3];
    idx[0] = (int)(t % B.n[0]);
    idx[1] = (int)((t / B.n[0]) % B.n[1]);
    idx[2] = (int)(t / ((long long)B.n[0] * B.n[1]));
    Complete<D> l;
    load_complete<D>(l, B.src + t, B.sc);
    DtData<D, BASE> dl, dr;
    stage_dt_cell<D, BASE>(dl, l);
    if (idx[0] + 1 < B.n[0] && A.boxes[0][blockIdx.y].dtmask[idx[0]]) {
      Complete<D> r;
      load_complete<D>(r, B.src + t + 1, B.sc);
      stage_dt_cell<D, BASE>(dr, r);
      const double s = stage_dt_face<D, 0, BASE>(eq, dl, dr, failed);
      maxsig[0] = s > maxsig[0] ? s : maxsig[0];
    }
    if constexpr (D >= 2) {
      if (idx[1] + 1 < B.n[1] && A.boxes[1][blockIdx.y].dtmask[idx[1]]) {
        Complete<D> r;
        load_complete<D>(r, B.src + t + B.n[0], B.sc);
        stage_dt_cell<D, BASE>(dr, r);
        const double s = stage_dt_face<D, 1, BASE>(eq, dl, dr, failed);
        maxsig[1] = s > maxsig[1] ? s : maxsig[1];
      }
    }
    if constexpr (D >= 3) {
      if (idx[2] + 1 < B.n[2] && A.boxes[2][blockIdx.y].dtmask[idx[2]]) {
        Complete<D> r;
        load_complete<D>(r, B.src + t + (long long)B.n[0] * B.n[1], B.sc);
        stage_dt_cell<D, BASE>(dr, r);
        const double s = stage_dt_face<D, 2, BASE>(eq, dl, dr, failed);
        maxsig[2] = s > maxsig[2] ? s : maxsig[2];
      }
    }
  }
#pragma unroll
  for (int d = 0; d < D; ++d) {
    reduce_max_signal(maxsig[d], A.max_signal + d);
    __syncthreads();
  }
  if (failed) *A.fail_flag = 1;
}

// The same with the lazy evaluation (numerics.cuh "Lazy stable dt"), HLLEM and
// HLL-Einfeldt: a read pass at HBM speed instead of 800 instructions per cell.
// A block covers a tile of 32 x 8 (x, y) columns and marches kDtAllSegment cells
// along z.  Per step every thread
// evaluates the bound of its own cell ONCE and publishes it in shared memory
// (double buffered, one barrier per step); the +x / +y faces read the
// neighbour's bound from there, the +z face keeps the previous step's.  Tiles
// overlap by one column and one row (the last column / row of a tile only
// supplies bounds), so no thread ever evaluates a neighbour's bound.  The exact
// signal speeds are evaluated only where a bound exceeds the threshold, out of
// line so that the march stays small.
constexpr int kDtAllSegment = 16;
constexpr int kDtTileX = 32, kDtTileY = 8, kDtPlanes = 2;
// exact signal speed of the face between the stored states at pl and pr (the
// lazy bases never set the Newton failure flag)
template <int D, int DIR, int BASE>
__device__ __noinline__ double dt_all_exact(const EqConst ec, const double* pl, const double* pr,
                                            long long sc) {
  const Eqs eq(ec);
  Complete<D> l, r;
  load_complete<D>(l, pl, sc);
  load_complete<D>(r, pr, sc);
  DtData<D, BASE> dl, dr;
  stage_dt_cell<D, BASE>(dl, l);
  stage_dt_cell<D, BASE>(dr, r);
  int failed = 0;
  return stage_dt_face<D, DIR, BASE>(eq, dl, dr, failed);
}
FVS_DEV void dt_all_take(double s, double& maxsig, double& thr) {
  if (s > maxsig) {
    maxsig = s;
    const double t = s * kLazyKeep;
    thr = t > thr ? t : thr;
  }
}
#ifndef FVS_DT_ALL_MIN_BLOCKS
#define FVS_DT_ALL_MIN_BLOCKS 4
#endif
template <int D, int BASE>
__global__ void __launch_bounds__(256, FVS_DT_ALL_MIN_BLOCKS) stable_dt_all_lazy_kernel(const DtAllArgs A) {
  __shared__ int4 s_b[2][kDtPlanes][kDtTileY][kDtTileX];
  const BoxDesc& B = A.boxes[0][blockIdx.y];
  const int n0 = B.n[0], n1 = B.n[1], n2 = B.n[2];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  // tiles advance by 31 / 7 cells (2-D and 3-D; a 1-D level has a single row)
  const int step_y = D >= 2 ? kDtTileY - 1 : 1;
  const int tiles_x = (n0 + kDtTileX - 2) / (kDtTileX - 1);
  const int x = (blockIdx.x % tiles_x) * (kDtTileX - 1) + tx;
  const int y = (blockIdx.x / tiles_x) * step_y + ty;
  const int z0 = blockIdx.z * kDtAllSegment;
  double maxsig[D], thr[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    maxsig[d] = 0.0;
    thr[d] = A.sig_hint[d] * kLazyKeep;
  }
  if (z0 < n2) { // block uniform
    const EqM eq(A.eq);
    const long long sc = B.sc;
    const long long plane = (long long)n0 * n1;
    const bool in_box = x < n0 && y < n1;
    // cells outside the box read the last cell (their bounds are never used)
    const double* p = B.src + min(x, n0 - 1) + (long long)n0 * min(y, n1 - 1) + plane * z0;
    const unsigned char* mx = A.boxes[0][blockIdx.y].dtmask;
    const unsigned char* my = D >= 2 ? A.boxes[1][blockIdx.y].dtmask : nullptr;
    const unsigned char* mz = D >= 3 ? A.boxes[2][blockIdx.y].dtmask : nullptr;
    // faces owned by this thread: not in the last column / row of the tile
    const bool owner = in_box && tx < kDtTileX - 1 && (D < 2 || ty < kDtTileY - 1);
    const bool face_x = owner && x + 1 < n0 && mx[x];
    const bool face_y = D >= 2 && owner && y + 1 < n1 && my[y];
    const int z_end = min(z0 + kDtAllSegment, n2 - 1); // last cell read: the +z neighbour of the segment
    SigBoundAll<D> prev{};
    // kDtPlanes planes per step: their loads are issued together (memory-level
    // parallelism without a register prefetch), one barrier per step
#pragma unroll 1
    for (int zb = z0; zb <= z_end; zb += kDtPlanes, p += plane * kDtPlanes) {
      Complete<D> q[kDtPlanes];
#pragma unroll
      for (int i = 0; i < kDtPlanes; ++i) {
        // planes past the end re-read the last one; their results are not used
        load_complete<D>(q[i], p + plane * min(i, z_end - zb), sc);
      }
      SigBoundAll<D> cur[kDtPlanes];
      const int buf = ((zb - z0) / kDtPlanes) & 1;
#pragma unroll
      for (int i = 0; i < kDtPlanes; ++i) {
        cur[i] = sig_bound_all<D>(eq, q[i]);
        int4 mine;
        mine.x = cur[i].u[0];
        mine.y = D >= 2 ? cur[i].u[1] : 0;
        mine.z = D >= 3 ? cur[i].u[2] : 0;
        mine.w = cur[i].c;
        s_b[buf][i][ty][tx] = mine;
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < kDtPlanes; ++i) {
        const int z = zb + i;
        if (z > z_end) break;
        const double* pz = p + plane * i;
        if constexpr (D >= 3) {
          if (owner && z > z0 && mz[z - 1] &&
              sig_bound_exceeds(prev.u[2], prev.c, cur[i].u[2], cur[i].c, thr[2])) {
            dt_all_take(dt_all_exact<D, 2, BASE>(A.eq, pz - plane, pz, sc), maxsig[2], thr[2]);
          }
        }
        prev = cur[i];
        if (z < z0 + kDtAllSegment) { // the +z neighbour of the segment only closes the z face
          if (face_x) {
            const int4 r = s_b[buf][i][ty][tx + 1];
            if (sig_bound_exceeds(cur[i].u[0], cur[i].c, r.x, r.w, thr[0])) {
              dt_all_take(dt_all_exact<D, 0, BASE>(A.eq, pz, pz + 1, sc), maxsig[0], thr[0]);
            }
          }
          if constexpr (D >= 2) {
            if (face_y) {
              const int4 r = s_b[buf][i][ty + 1][tx];
              if (sig_bound_exceeds(cur[i].u[1], cur[i].c, r.y, r.w, thr[1])) {
                dt_all_take(dt_all_exact<D, 1, BASE>(A.eq, pz, pz + n0, sc), maxsig[1], thr[1]);
              }
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int d = 0; d < D; ++d) {
    reduce_max_signal(maxsig[d], A.max_signal + d);
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// Cut-cell (embedded boundary) levels.  Thread per face, plain operators: the
// cut-cell fabs are the small, branchy part of an EB run (regular fabs of an
// EB level use the same kernels with fab_type == 0).

// KbnCutCellMethod::ComputeRegularFluxes, KbnStabilisation.hpp:347-429 ->
// masked MusclHancock (MusclHancockMethod.hpp:487-571) -> masked base method
// (HllemMethod.hpp:1074-1093, HllMethod.hpp:295-335, GodunovMethod.hpp:147-160);
// regular fabs: FluxMethod::ComputeNumericFluxes as in numeric_flux_kernel.
template <int D, int DIR, int BASE, int RECON, int LIM>
__global__ void __launch_bounds__(128) cc_numeric_flux_kernel(const SweepArgs A) {
  const BoxDesc& B = A.boxes[blockIdx.y];
  if (B.fab_type < 0) return; // covered fab: nothing is computed
  int len[3] = {B.n[0], B.n[1], B.n[2]};
  len[DIR] = B.fhi - B.flo + 1;
  const long long total = (long long)len[0] * len[1] * len[2];
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int failed = 0;
  if (len[DIR] > 0 && t < total) {
    int idx[3];
    idx[0] = (int)(t % len[0]);
    idx[1] = (int)((t / len[0]) % len[1]);
    idx[2] = (int)(t / ((long long)len[0] * len[1]));
    idx[DIR] += B.flo;
    const long long stride = DIR == 0 ? 1 : (DIR == 1 ? (long long)B.n[0] : (long long)B.n[0] * B.n[1]);
    const long long cell = idx[0] + (long long)B.n[0] * (idx[1] + (long long)B.n[1] * idx[2]);
    const double* p = B.src + cell; // right neighbour of the face
    const Eqs eq(A.eq);
    Cons<D> F;
    const bool single = B.fab_type == 1;
    bool valid = true;
    if (single) {
      const int bn0 = B.n[0] + (DIR == 0 ? 1 : 0);
      const int bn1 = B.n[1] + (DIR == 1 ? 1 : 0);
      const double beta = B.beta[idx[0] + (long long)bn0 * (idx[1] + (long long)bn1 * idx[2])];
      valid = beta > 0.0;
    }
    if (!valid) {
      F.rho = 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) F.mom[d] = 0.0;
      F.E = 0.0;
    } else {
      Flag fl;
      Edge<D> unused, left, right;
      if constexpr (RECON == kReconMusclHancockPrim) {
        Complete<D> s0, s1, s2, s3;
        load_complete<D>(s0, p - 2 * stride, B.sc);
        load_complete<D>(s1, p - stride, B.sc);
        load_complete<D>(s2, p, B.sc);
        load_complete<D>(s3, p + stride, B.sc);
        if (single) {
          const double* al = B.alpha + cell;
          muscl_prim_face_masked<SafeMath, D, DIR>(eq.m, s0, s1, s2, s3, al[-2 * stride] > 0.0,
                                                   al[stride] > 0.0, A.lambda, left, right, fl);
        } else {
          muscl_prim_predict<SafeMath, D, DIR>(eq.m, s0, s1, s2, A.lambda, unused, left, fl);
          muscl_prim_predict<SafeMath, D, DIR>(eq.m, s1, s2, s3, A.lambda, right, unused, fl);
        }
      } else if constexpr (RECON == kReconMusclHancockChar) {
        Complete<D> s0, s1, s2, s3;
        load_complete<D>(s0, p - 2 * stride, B.sc);
        load_complete<D>(s1, p - stride, B.sc);
        load_complete<D>(s2, p, B.sc);
        load_complete<D>(s3, p + stride, B.sc);
        if (single) {
          const double* al = B.alpha + cell;
          const bool a0 = al[-2 * stride] > 0.0, a1 = al[-stride] > 0.0, a2 = al[0] > 0.0,
                     a3 = al[stride] > 0.0;
          muscl_char_predict<SafeMath, D, DIR, true>(eq.m, s0, s1, s2, A.lambda, unused, left, fl,
                                                     a0 && a1 && a2);
          muscl_char_predict<SafeMath, D, DIR, true>(eq.m, s1, s2, s3, A.lambda, right, unused, fl,
                                                     a1 && a2 && a3);
        } else {
          muscl_char_predict<SafeMath, D, DIR>(eq.m, s0, s1, s2, A.lambda, unused, left, fl);
          muscl_char_predict<SafeMath, D, DIR>(eq.m, s1, s2, s3, A.lambda, right, unused, fl);
        }
      } else if constexpr (is_mh2(RECON)) {
        // MusclHancockMethod2.hpp:193-216: only the gradients are masked
        Complete<D> s0, s1, s2, s3;
        load_complete<D>(s0, p - 2 * stride, B.sc);
        load_complete<D>(s1, p - stride, B.sc);
        load_complete<D>(s2, p, B.sc);
        load_complete<D>(s3, p + stride, B.sc);
        bool okL = true, okR = true;
        if (single) {
          const double* al = B.alpha + cell;
          const bool a0 = al[-2 * stride] > 0.0, a1 = al[-stride] > 0.0, a2 = al[0] > 0.0,
                     a3 = al[stride] > 0.0;
          okL = a0 && a1 && a2;
          okR = a1 && a2 && a3;
        }
        mh2_predict<SafeMath, D, DIR, RECON>(eq.m, eq.c, s0, s1, s2, A.lambda_half, unused, left, fl,
                                             okL);
        mh2_predict<SafeMath, D, DIR, RECON>(eq.m, eq.c, s1, s2, s3, A.lambda_half, right, unused,
                                             fl, okR);
      } else if constexpr (RECON == kReconMusclHancock) {
        Complete<D> s0, s1, s2, s3;
        load_complete<D>(s0, p - 2 * stride, B.sc);
        load_complete<D>(s1, p - stride, B.sc);
        load_complete<D>(s2, p, B.sc);
        load_complete<D>(s3, p + stride, B.sc);
        if (single) {
          const double* al = B.alpha + cell;
          const bool a0 = al[-2 * stride] > 0.0, a1 = al[-stride] > 0.0, a2 = al[0] > 0.0,
                     a3 = al[stride] > 0.0;
          muscl_predict<SafeMath, D, DIR, LIM, true>(eq.m, s0, s1, s2, A.lambda_half, unused, left,
                                                     fl, a0 && a1 && a2);
          muscl_predict<SafeMath, D, DIR, LIM, true>(eq.m, s1, s2, s3, A.lambda_half, right,
                                                     unused, fl, a1 && a2 && a3);
        } else {
          muscl_predict<SafeMath, D, DIR, LIM>(eq.m, s0, s1, s2, A.lambda_half, unused, left, fl);
          muscl_predict<SafeMath, D, DIR, LIM>(eq.m, s1, s2, s3, A.lambda_half, right, unused, fl);
        }
      } else {
        load_complete<D>(left.q, p - stride, B.sc);
        load_complete<D>(right.q, p, B.sc);
        left.rc = SafeMath::rcp(left.q.rho);
        right.rc = SafeMath::rcp(right.q.rho);
      }
      if (single) {
        base_flux<SafeMath, D, DIR, BASE, true>(eq, F, left, right, fl, failed);
      } else {
        base_flux<SafeMath, D, DIR, BASE, false>(eq, F, left, right, fl, failed);
      }
    }
    const long long fcell = (idx[0] - B.foff[0]) +
                            (long long)B.fn[0] * ((idx[1] - B.foff[1]) +
                                                  (long long)B.fn[1] * (idx[2] - B.foff[2]));
    double* f = B.flux + fcell;
    f[0] = F.rho;
#pragma unroll
    for (int d = 0; d < D; ++d) f[B.fsc * (1 + d)] = F.mom[d];
    f[B.fsc * (1 + D)] = F.E;
  }
  if (failed) *A.fail_flag = 1;
}

// cutcell::FluxMethod::ComputeStableDt (include/fub/AMReX/cutcell/FluxMethod.hpp:258-289):
// the faces are those of the flux box whose stencil fits (flo..fhi), not the
// grown-tile set of the regular adapter; single-valued fabs take the masked
// signal formulas on faces with a positive face fraction
// (KbnStabilisation.hpp:277-345).
template <int D, int DIR, int BASE>
__global__ void __launch_bounds__(256) cc_stable_dt_kernel(const SweepArgs A) {
  const BoxDesc& B = A.boxes[blockIdx.y];
  int len[3] = {B.n[0], B.n[1], B.n[2]};
  len[DIR] = B.fhi - B.flo + 1;
  const long long total = (long long)len[0] * len[1] * len[2];
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double maxsig = 0.0;
  int failed = 0;
  if (B.fab_type >= 0 && len[DIR] > 0 && t < total) {
    int idx[3];
    idx[0] = (int)(t % len[0]);
    idx[1] = (int)((t / len[0]) % len[1]);
    idx[2] = (int)(t / ((long long)len[0] * len[1]));
    idx[DIR] += B.flo;
    const long long stride = DIR == 0 ? 1 : (DIR == 1 ? (long long)B.n[0] : (long long)B.n[0] * B.n[1]);
    const long long cell = idx[0] + (long long)B.n[0] * (idx[1] + (long long)B.n[1] * idx[2]);
    const bool single = B.fab_type == 1;
    bool valid = true;
    if (single) {
      const int bn0 = B.n[0] + (DIR == 0 ? 1 : 0);
      const int bn1 = B.n[1] + (DIR == 1 ? 1 : 0);
      valid = B.beta[idx[0] + (long long)bn0 * (idx[1] + (long long)bn1 * idx[2])] > 0.0;
    }
    if (valid) {
      const Eqs eq(A.eq);
      Complete<D> l, r;
      load_complete<D>(l, B.src + cell - stride, B.sc);
      load_complete<D>(r, B.src + cell, B.sc);
      DtData<D, BASE> dl, dr;
      Flag fl;
      dt_prepare<SafeMath, D, BASE>(dl, l, fl);
      dt_prepare<SafeMath, D, BASE>(dr, r, fl);
      const double s = single
                           ? base_max_signal<SafeMath, D, DIR, BASE, true>(eq, dl, dr, fl, failed)
                           : base_max_signal<SafeMath, D, DIR, BASE, false>(eq, dl, dr, fl, failed);
      maxsig = s > 0.0 ? s : 0.0; // drops NaN
    }
  }
  reduce_max_signal(maxsig, A.max_signal);
  if (failed) *A.fail_flag = 1;
}

} // namespace fvs
