// sweep_inst.cu -- instantiates the method-dependent kernels for ONE space
// dimension (FVS_INST_D) and ONE base method (FVS_INST_BASE); compiled twelve
// times by the build (3 dimensions x 4 base methods) so the translation
// units build in parallel.
//
// FVS_TOLERANCE_UNIT: the same fused-sweep kernels built a second time with the
// tolerance arithmetic (fastmath.cuh ToleranceMath, FMA contraction on) for
// contexts that opted in with fvs_set_arithmetic(FVS_ARITH_TOLERANCE).  The
// whole header stack is compiled into namespace fvs_tol so the two builds of
// the templates never meet at link time; only the fused launcher is exported,
// through plain pointers (the descriptors are the same PODs in both
// namespaces).
#ifdef FVS_TOLERANCE_UNIT
#define FVS_FASTMATH 2
#define fvs fvs_tol
#endif
#include "sweep_kernels.cuh"

#ifndef FVS_SWEEP_SMEM
#define FVS_SWEEP_SMEM 1
#endif
#include <algorithm>
#include <type_traits>
#ifndef FVS_INST_D
#error "define FVS_INST_D"
#endif
#ifndef FVS_INST_BASE
#error "define FVS_INST_BASE"
#endif

#define FVS_CAT3(a, b, c) a##b##_##c
#define FVS_CAT(a, b, c) FVS_CAT3(a, b, c)
#define FVS_FN(name) FVS_CAT(name##_D, FVS_INST_D, FVS_INST_BASE)

namespace fvs {
namespace {

constexpr int D = FVS_INST_D;
constexpr int BASE = FVS_INST_BASE;

inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

template <int DIR, int RECON, int LIM>
int fused(const SweepArgs& A, cudaStream_t s) {
  if constexpr (DIR == 0) {
    const long long rows = (long long)A.max_n[1] * A.max_n[2];
    const int nseg = (A.max_n[0] + A.seg_len - 1) / A.seg_len;
    dim3 grid(cdiv(rows * nseg, kSweepThreads / 32), A.nboxes, 1);
    fused_sweep_x_smem_kernel<D, BASE, RECON, LIM><<<grid, kSweepThreads, 0, s>>>(A);
    repair_kernel<D, DIR, BASE, RECON, LIM><<<kRepairBlocks, 128, 0, s>>>(A);
    return 2;
  } else {
    const long long lines = (long long)A.max_n[0] * (DIR == 1 ? A.max_n[2] : A.max_n[1]);
    const int nseg = (A.max_n[DIR] + A.seg_len - 1) / A.seg_len;
    dim3 grid(cdiv(lines, kSweepThreads), A.nboxes, nseg);
    // the shared-memory pipeline pays off where the register pipeline is at
    // the 255-register limit: the MUSCL-Hancock variants
    if constexpr (FVS_SWEEP_SMEM != 0 && RECON == kReconMusclHancock) {
      fused_sweep_yz_smem_kernel<D, DIR, BASE, RECON, LIM><<<grid, kSweepThreads, 0, s>>>(A);
      repair_kernel<D, DIR, BASE, RECON, LIM><<<kRepairBlocks, 128, 0, s>>>(A);
      return 2;
    } else {
      fused_sweep_yz_kernel<D, DIR, BASE, RECON, LIM><<<grid, kSweepThreads, 0, s>>>(A);
    }
  }
  return 1;
}

#ifndef FVS_TOLERANCE_UNIT
template <int DIR, int RECON, int LIM>
int fluxes(const SweepArgs& A, cudaStream_t s) {
  long long total = (long long)A.max_n[0] * A.max_n[1] * A.max_n[2];
  // along DIR there is at most one more face than cells
  total += total / A.max_n[DIR];
  dim3 grid(cdiv(total, 128), A.nboxes, 1);
  numeric_flux_kernel<D, DIR, BASE, RECON, LIM><<<grid, 128, 0, s>>>(A);
  return 1;
}

template <int DIR>
int stable_dt(const SweepArgs& A, cudaStream_t s) {
  const long long total = (long long)A.max_n[0] * A.max_n[1] * A.max_n[2];
  dim3 grid(cdiv(total, 256), A.nboxes, 1);
  stable_dt_kernel<D, DIR, BASE><<<grid, 256, 0, s>>>(A);
  return 1;
}

template <int DIR, int RECON, int LIM>
int cc_fluxes(const SweepArgs& A, cudaStream_t s) {
  long long total = (long long)A.max_n[0] * A.max_n[1] * A.max_n[2];
  total += total / A.max_n[DIR];
  dim3 grid(cdiv(total, 128), A.nboxes, 1);
  cc_numeric_flux_kernel<D, DIR, BASE, RECON, LIM><<<grid, 128, 0, s>>>(A);
  return 1;
}

template <int DIR>
int cc_stable_dt(const SweepArgs& A, cudaStream_t s) {
  long long total = (long long)A.max_n[0] * A.max_n[1] * A.max_n[2];
  total += total / A.max_n[DIR];
  dim3 grid(cdiv(total, 256), A.nboxes, 1);
  cc_stable_dt_kernel<D, DIR, BASE><<<grid, 256, 0, s>>>(A);
  return 1;
}

#endif // !FVS_TOLERANCE_UNIT

template <int DIR, typename F0, typename F1, typename F2, typename F3, typename F4, typename F5>
int by_recon(const MethodKey& k, F0 fo, F1 none, F2 minmod, F3 vanleer, F4 prim, F5 chr) {
  if (k.recon == kReconFirstOrder) return fo();
  if (k.recon == kReconMusclHancockPrim) return prim();
  if (k.recon == kReconMusclHancockChar) return chr();
  switch (k.limiter) {
  case kLimiterNone: return none();
  case kLimiterMinMod: return minmod();
  case kLimiterVanLeer: return vanleer();
  }
  return -1;
}

// MusclHancock2: the gradient / reconstruction variables are a template argument, the
// limiter of the CentralDifferenceGradient a run-time value (EqConst::limiter2)
template <typename F> int by_mh2(const MethodKey& k, F f) {
  switch (k.recon) {
  case kReconMh2Conservative: return f(std::integral_constant<int, kReconMh2Conservative>{});
  case kReconMh2Primitive: return f(std::integral_constant<int, kReconMh2Primitive>{});
  case kReconMh2Characteristics: return f(std::integral_constant<int, kReconMh2Characteristics>{});
  }
  return -1;
}

template <int DIR> int fused_dir(const MethodKey& k, const SweepArgs& A, cudaStream_t s) {
#ifdef FVS_INST_MINIMAL
  // tuning builds (tools/build_variant.py --minimal, tools/count_steady.sh): only the
  // bench configuration, MUSCL-Hancock with VanLeer, is instantiated
  if (k.recon == kReconMusclHancock && k.limiter == kLimiterVanLeer) {
    return fused<DIR, kReconMusclHancock, kLimiterVanLeer>(A, s);
  }
  return -1;
#else
  if (is_mh2(k.recon)) {
#ifdef FVS_TOLERANCE_UNIT
    return -1; // no tolerance-arithmetic build of MusclHancock2
#else
    return by_mh2(k, [&](auto R) { return fused<DIR, decltype(R)::value, kLimiterNone>(A, s); });
#endif
  }
  return by_recon<DIR>(
      k, [&] { return fused<DIR, kReconFirstOrder, kLimiterNone>(A, s); },
      [&] { return fused<DIR, kReconMusclHancock, kLimiterNone>(A, s); },
      [&] { return fused<DIR, kReconMusclHancock, kLimiterMinMod>(A, s); },
      [&] { return fused<DIR, kReconMusclHancock, kLimiterVanLeer>(A, s); },
      [&] {
        // perfect_gas::MusclHancockPrim exists over HLLEM-Larrouturou only
        if constexpr (BASE == kBaseHllemLarrouturou) {
          return fused<DIR, kReconMusclHancockPrim, kLimiterNone>(A, s);
        } else {
          return -1;
        }
      },
      [&] {
        // perfect_gas::MusclHancockCharacteristic: VanLeer amplitudes over HLLEM-Larrouturou
        if constexpr (BASE == kBaseHllemLarrouturou) {
          return fused<DIR, kReconMusclHancockChar, kLimiterVanLeer>(A, s);
        } else {
          return -1;
        }
      });
#endif
}
#ifndef FVS_TOLERANCE_UNIT
template <int DIR> int fluxes_dir(const MethodKey& k, const SweepArgs& A, cudaStream_t s) {
#ifdef FVS_INST_MINIMAL
  return -1;
#else
  if (is_mh2(k.recon)) {
    return by_mh2(k, [&](auto R) { return fluxes<DIR, decltype(R)::value, kLimiterNone>(A, s); });
  }
  return by_recon<DIR>(
      k, [&] { return fluxes<DIR, kReconFirstOrder, kLimiterNone>(A, s); },
      [&] { return fluxes<DIR, kReconMusclHancock, kLimiterNone>(A, s); },
      [&] { return fluxes<DIR, kReconMusclHancock, kLimiterMinMod>(A, s); },
      [&] { return fluxes<DIR, kReconMusclHancock, kLimiterVanLeer>(A, s); },
      [&] {
        // perfect_gas::MusclHancockPrim exists over HLLEM-Larrouturou only
        if constexpr (BASE == kBaseHllemLarrouturou) {
          return fluxes<DIR, kReconMusclHancockPrim, kLimiterNone>(A, s);
        } else {
          return -1;
        }
      },
      [&] {
        // perfect_gas::MusclHancockCharacteristic: VanLeer amplitudes over HLLEM-Larrouturou
        if constexpr (BASE == kBaseHllemLarrouturou) {
          return fluxes<DIR, kReconMusclHancockChar, kLimiterVanLeer>(A, s);
        } else {
          return -1;
        }
      });
#endif
}

template <int DIR> int cc_fluxes_dir(const MethodKey& k, const SweepArgs& A, cudaStream_t s) {
#ifdef FVS_INST_MINIMAL
  return -1;
#else
  if (is_mh2(k.recon)) {
    return by_mh2(k, [&](auto R) { return cc_fluxes<DIR, decltype(R)::value, kLimiterNone>(A, s); });
  }
  return by_recon<DIR>(
      k, [&] { return cc_fluxes<DIR, kReconFirstOrder, kLimiterNone>(A, s); },
      [&] { return cc_fluxes<DIR, kReconMusclHancock, kLimiterNone>(A, s); },
      [&] { return cc_fluxes<DIR, kReconMusclHancock, kLimiterMinMod>(A, s); },
      [&] { return cc_fluxes<DIR, kReconMusclHancock, kLimiterVanLeer>(A, s); },
      [&] {
        // perfect_gas::MusclHancockPrim exists over HLLEM-Larrouturou only
        if constexpr (BASE == kBaseHllemLarrouturou) {
          return cc_fluxes<DIR, kReconMusclHancockPrim, kLimiterNone>(A, s);
        } else {
          return -1;
        }
      },
      [&] {
        // perfect_gas::MusclHancockCharacteristic: VanLeer amplitudes over HLLEM-Larrouturou
        if constexpr (BASE == kBaseHllemLarrouturou) {
          return cc_fluxes<DIR, kReconMusclHancockChar, kLimiterVanLeer>(A, s);
        } else {
          return -1;
        }
      });
#endif
}

#endif // !FVS_TOLERANCE_UNIT

} // namespace

#if FVS_INST_D >= 2 && !defined(FVS_TOLERANCE_UNIT)
int FVS_FN(launch_cc_numeric_fluxes)(const MethodKey& k, int dir, const SweepArgs& A,
                                     cudaStream_t s) {
  if (dir == 0) return cc_fluxes_dir<0>(k, A, s);
  if (dir == 1) return cc_fluxes_dir<1>(k, A, s);
  if constexpr (D >= 3) {
    if (dir == 2) return cc_fluxes_dir<2>(k, A, s);
  }
  return -1;
}
int FVS_FN(launch_cc_stable_dt)(const MethodKey&, int dir, const SweepArgs& A, cudaStream_t s) {
  if (dir == 0) return cc_stable_dt<0>(A, s);
  if (dir == 1) return cc_stable_dt<1>(A, s);
  if constexpr (D >= 3) {
    if (dir == 2) return cc_stable_dt<2>(A, s);
  }
  return -1;
}
#endif

int FVS_FN(launch_fused_sweep)(const MethodKey& k, int dir, const SweepArgs& A,
                               cudaStream_t s) {
  if (dir == 0) return fused_dir<0>(k, A, s);
  if constexpr (D >= 2) {
    if (dir == 1) return fused_dir<1>(k, A, s);
  }
  if constexpr (D >= 3) {
    if (dir == 2) return fused_dir<2>(k, A, s);
  }
  return -1;
}

#ifndef FVS_TOLERANCE_UNIT
int FVS_FN(launch_numeric_fluxes)(const MethodKey& k, int dir, const SweepArgs& A,
                                  cudaStream_t s) {
  if (dir == 0) return fluxes_dir<0>(k, A, s);
  if constexpr (D >= 2) {
    if (dir == 1) return fluxes_dir<1>(k, A, s);
  }
  if constexpr (D >= 3) {
    if (dir == 2) return fluxes_dir<2>(k, A, s);
  }
  return -1;
}

int FVS_FN(launch_stable_dt)(const MethodKey&, int dir, const SweepArgs& A, cudaStream_t s) {
  if (dir == 0) return stable_dt<0>(A, s);
  if constexpr (D >= 2) {
    if (dir == 1) return stable_dt<1>(A, s);
  }
  if constexpr (D >= 3) {
    if (dir == 2) return stable_dt<2>(A, s);
  }
  return -1;
}

#endif // !FVS_TOLERANCE_UNIT

namespace {
// a template, so that only the kernel of the taken branch is instantiated
template <int B> int dt_all(const DtAllArgs& A, cudaStream_t s) {
  if constexpr (kLazyDt<B>) {
    // overlapping tiles: 31 (x) by 7 (y; 1 for a 1-D level) owned columns per block
    const int step_y = D >= 2 ? kDtTileY - 1 : 1;
    dim3 grid(cdiv(A.max_n[0], kDtTileX - 1) * cdiv(A.max_n[1], step_y), A.nboxes,
              cdiv(A.max_n[2], kDtAllSegment));
    stable_dt_all_lazy_kernel<D, B><<<grid, 256, 0, s>>>(A);
  } else {
    const long long total = (long long)A.max_n[0] * A.max_n[1] * A.max_n[2];
    dim3 grid(cdiv(total, 256 * kDtAllCellsPerThread), A.nboxes, 1);
    stable_dt_all_kernel<D, B><<<grid, 256, 0, s>>>(A);
  }
  return 1;
}
} // namespace

int FVS_FN(launch_stable_dt_all)(const MethodKey&, const DtAllArgs& A, cudaStream_t s) {
  return dt_all<BASE>(A, s);
}

} // namespace fvs

#ifdef FVS_TOLERANCE_UNIT
#undef fvs
namespace fvs {
// called by launch_fused_sweep_tolerance / launch_stable_dt_all_tolerance (generic_kernels.cu)
int FVS_FN(launch_fused_sweep_tol)(const void* key, int dir, const void* args, cudaStream_t s) {
  return fvs_tol::FVS_FN(launch_fused_sweep)(*static_cast<const fvs_tol::MethodKey*>(key), dir,
                                             *static_cast<const fvs_tol::SweepArgs*>(args), s);
}
int FVS_FN(launch_stable_dt_all_tol)(const void* key, const void* args, cudaStream_t s) {
  return fvs_tol::FVS_FN(launch_stable_dt_all)(*static_cast<const fvs_tol::MethodKey*>(key),
                                               *static_cast<const fvs_tol::DtAllArgs*>(args), s);
}
} // namespace fvs
#endif
