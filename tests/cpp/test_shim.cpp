// tests/cpp/test_shim.cpp -- the C++17 policy shim (include/fub_b200/) driven
// the way the reference drives its policies, checked against the CPU oracle.
//
// Mirrors the set-up of examples/AMReX/SodShockTube2d.cpp:93-161 (context
// options 4w/3w ghost cells, HyperbolicMethod{flux, ForwardIntegrator,
// Reconstruction}, DimensionalSplitLevelIntegrator) for Rank 1..3 and checks,
// bit for bit, that
//   (a) the three unfused policies, sequenced by DimensionalSplitLevelIntegrator,
//   (b) the same integrator with the fused single-pass sweep
// reproduce the oracle's states and per-step stable dt.
//
// Links libfvs_b200.so (product) and liboracle.so (checker, test only).
// Usage: test_shim            (exit code 0 = all checks passed)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "fub_b200/HyperbolicMethod.hpp"

extern "C" {
// oracle/oracle_capi.cpp
void* fvso_level_create(int rank, const int* dom_lo, const int* dom_hi, const double* dx,
                        const int* periodic, int nboxes, const int* box_lo, const int* box_hi,
                        int scratch_gcw, int flux_gcw, double gamma);
void fvso_level_destroy(void* h);
void fvso_level_set_method(void* h, int base, int recon, int limiter);
void fvso_level_add_bc(void* h, int dir, int side, int kind);
void fvso_level_set_cycles(void* h, long cycles, double time_point);
double* fvso_level_scratch(void* h, int box, int* lo, int* hi, int* ncomp);
int fvso_fill_ghost_layer(void* h);
int fvso_compute_stable_dt(void* h, int dir, double* out);
int fvso_advance_hierarchy(void* h, double dt, int split_kind, double* too_large_dt,
                           double* sweep_dts, int* n_sweep_dts);
}

using namespace fub;
using namespace fub::b200;

static int g_failures = 0;
#define REQUIRE(cond)                                                         \
  do {                                                                        \
    if (!(cond)) {                                                            \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);           \
      ++g_failures;                                                           \
    }                                                                         \
  } while (0)

// Riemann-problem initial data (sphere / disc / interval of radius 0.4)
static void InitialData(int rank, const FabView& v, const std::array<double, 3>& dx,
                        std::vector<double>& fab) {
  const int n[3] = {v.box.hi[0] - v.box.lo[0] + 1, v.box.hi[1] - v.box.lo[1] + 1,
                    v.box.hi[2] - v.box.lo[2] + 1};
  const std::size_t sc = std::size_t(n[0]) * n[1] * n[2];
  fab.assign(sc * v.ncomp, 0.0);
  const double gm1_inv = 1.0 / (1.4 - 1.0);
  for (int k = 0; k < n[2]; ++k)
    for (int j = 0; j < n[1]; ++j)
      for (int i = 0; i < n[0]; ++i) {
        const int idx[3] = {i + v.box.lo[0], j + v.box.lo[1], k + v.box.lo[2]};
        double r2 = 0.0;
        for (int d = 0; d < rank; ++d) {
          const double x = -1.0 + (idx[d] + 0.5) * dx[d];
          r2 += x * x;
        }
        const bool in = r2 < 0.16;
        const double rho = in ? 1.0 : 0.125;
        const double p = in ? 1.0 : 0.1;
        const std::size_t c = i + std::size_t(n[0]) * (j + std::size_t(n[1]) * k);
        fab[c] = rho;
        // a small shear so tangential momenta are exercised
        for (int d = 0; d < rank; ++d) fab[c + sc * (1 + d)] = rho * 0.05 * (d + 1);
        double ekin = 0.0;
        for (int d = 0; d < rank; ++d) ekin += 0.5 * rho * (0.05 * (d + 1)) * (0.05 * (d + 1));
        fab[c + sc * (1 + rank)] = p * gm1_inv + ekin;
        fab[c + sc * (2 + rank)] = p;
        fab[c + sc * (3 + rank)] = std::sqrt(1.4 * p / rho);
      }
}

template <int Rank, typename Split>
static void RunCase(int split_kind, bool fused, int ncells, int nsteps) {
  const int w = 2;
  IntegratorContext::Options opts;
  opts.scratch_gcw = 4 * w; // examples/AMReX/SodShockTube2d.cpp:151-153
  opts.flux_gcw = 3 * w;
  PerfectGas<Rank> eq;
  IntegratorContext context(eq, BaseMethod::HllemLarrouturou, Reconstruct::MusclHancock,
                            Limiter::VanLeer, opts);
  Box domain;
  std::array<double, 3> dx{1.0, 1.0, 1.0};
  for (int d = 0; d < Rank; ++d) {
    domain.hi[d] = ncells - 1;
    dx[d] = 2.0 / ncells;
  }
  // two boxes split along x
  Box b0 = domain, b1 = domain;
  b0.hi[0] = ncells / 2 - 1;
  b1.lo[0] = ncells / 2;
  const std::vector<Box> boxes{b0, b1};
  context.DefineLevel(0, domain, dx, {0, 0, 0}, boxes);

  int olo[6], ohi[6];
  for (int b = 0; b < 2; ++b)
    for (int d = 0; d < 3; ++d) {
      olo[3 * b + d] = boxes[b].lo[d];
      ohi[3 * b + d] = boxes[b].hi[d];
    }
  const int periodic[3] = {0, 0, 0};
  void* oracle = fvso_level_create(Rank, domain.lo.data(), domain.hi.data(), dx.data(), periodic,
                                   2, olo, ohi, opts.scratch_gcw, opts.flux_gcw, eq.gamma);
  fvso_level_set_method(oracle, FVS_BASE_HLLEM_LARROUTUROU, FVS_RECON_MUSCL_HANCOCK,
                        FVS_LIMITER_VANLEER);
  for (int d = 0; d < Rank; ++d)
    for (int side = 0; side < 2; ++side) {
      const Boundary kind = d == 0 ? Boundary::Reflective : Boundary::Transmissive;
      context.AddBoundary(0, static_cast<Direction>(d), side, kind);
      fvso_level_add_bc(oracle, d, side, static_cast<int>(kind));
    }

  std::vector<double> fab;
  for (int b = 0; b < 2; ++b) {
    const FabView v = context.ScratchView(0, b);
    REQUIRE(v.ncomp == PerfectGas<Rank>::n_complete);
    InitialData(Rank, v, dx, fab);
    context.UploadScratch(0, b, fab.data());
    int lo[3], hi[3], nc;
    double* o = fvso_level_scratch(oracle, b, lo, hi, &nc);
    REQUIRE(nc == v.ncomp);
    for (int d = 0; d < 3; ++d) REQUIRE(lo[d] == v.box.lo[d] && hi[d] == v.box.hi[d]);
    std::memcpy(o, fab.data(), fab.size() * sizeof(double));
  }
  context.FillGhostLayerSingleLevel(0);
  REQUIRE(fvso_fill_ghost_layer(oracle) == 0);

  HyperbolicMethod method{FluxMethod{}, TimeIntegrator{}, Reconstruction{}};
  REQUIRE(method.flux_method.GetStencilWidth() == fvs_stencil_width(context.Handle()));
  DimensionalSplitLevelIntegrator<Rank, Split> integrator(context, method, Split(), fused);

  for (int step = 0; step < nsteps; ++step) {
    const Duration stable = integrator.ComputeStableDt(0);
    double odt = std::numeric_limits<double>::max();
    for (int d = 0; d < Rank; ++d) {
      double v = 0.0;
      REQUIRE(fvso_compute_stable_dt(oracle, d, &v) == 0);
      odt = std::min(odt, v);
    }
    REQUIRE(stable.count() == odt); // identical dt sequence
    const Duration dt(0.5 * stable.count());
    AdvanceResult r = integrator.AdvanceLevelNonRecursively(0, dt);
    REQUIRE(!r);
    double too_large = 0.0;
    REQUIRE(fvso_advance_hierarchy(oracle, dt.count(), split_kind, &too_large, nullptr,
                                   nullptr) == 0);
    // the oracle's AdvanceHierarchy also bumps cycles and refills ghosts
    // (SubcycleFineFirstSolver.hpp:177-193); do the same on the shim side
    context.SetCycles(context.GetCycles() + 1, Duration(context.GetTimePoint().count() + dt.count()));
    context.FillGhostLayerSingleLevel(0);
  }
  long long mismatches = 0;
  for (int b = 0; b < 2; ++b) {
    const FabView v = context.ScratchView(0, b);
    fab.resize(v.size());
    context.DownloadScratch(0, b, fab.data());
    int lo[3], hi[3], nc;
    const double* o = fvso_level_scratch(oracle, b, lo, hi, &nc);
    for (std::size_t i = 0; i < fab.size(); ++i) {
      if (std::memcmp(&fab[i], &o[i], sizeof(double)) != 0) ++mismatches;
    }
  }
  std::printf("Rank %d split %d %s: %d steps, %lld mismatching doubles, %ld kernel launches\n",
              Rank, split_kind, fused ? "fused" : "policies", nsteps, mismatches,
              context.KernelLaunches());
  REQUIRE(mismatches == 0);
  REQUIRE(context.KernelLaunches() > 0);
  fvso_level_destroy(oracle);
}

// A too large dt must come back as TimeStepTooLarge, not as an exception, and
// leave the state untouched (DimensionalSplitLevelIntegrator.hpp:243-245).
static void RejectsTooLargeDt() {
  IntegratorContext::Options opts;
  opts.scratch_gcw = 8;
  opts.flux_gcw = 6;
  PerfectGas<2> eq;
  IntegratorContext context(eq, BaseMethod::HllemLarrouturou, Reconstruct::MusclHancock,
                            Limiter::VanLeer, opts);
  Box domain;
  domain.hi = {31, 31, 0};
  const std::array<double, 3> dx{2.0 / 32, 2.0 / 32, 1.0};
  context.DefineLevel(0, domain, dx, {0, 0, 0}, {domain});
  std::vector<double> fab, after;
  const FabView v = context.ScratchView(0, 0);
  InitialData(2, v, dx, fab);
  context.UploadScratch(0, 0, fab.data());
  for (bool fused : {false, true}) {
    DimensionalSplitLevelIntegrator<2, GodunovSplitting> integrator(
        context, HyperbolicMethod{}, GodunovSplitting(), fused);
    const Duration stable = integrator.ComputeStableDt(0);
    AdvanceResult r = integrator.AdvanceLevelNonRecursively(0, Duration(4.0 * stable.count()));
    REQUIRE(bool(r));
    if (r) {
      REQUIRE(r->level == 0);
      REQUIRE(r->dt.count() < 4.0 * stable.count());
    }
    after.resize(fab.size());
    context.DownloadScratch(0, 0, after.data());
    REQUIRE(std::memcmp(after.data(), fab.data(), fab.size() * sizeof(double)) == 0);
  }
  // hard errors throw std::runtime_error
  bool threw = false;
  try {
    context.ScratchView(0, 99);
  } catch (const std::runtime_error&) {
    threw = true;
  }
  REQUIRE(threw);
}

int main() {
  if (fvs_device_count() <= 0) {
    std::printf("no CUDA device: the shim has no CPU fallback\n");
    return 2;
  }
  for (bool fused : {false, true}) {
    RunCase<1, GodunovSplitting>(FVS_SPLIT_GODUNOV, fused, 64, 4);
    RunCase<2, GodunovSplitting>(FVS_SPLIT_GODUNOV, fused, 32, 3);
    RunCase<2, StrangSplitting>(FVS_SPLIT_STRANG, fused, 32, 3);
    RunCase<2, StrangSplittingLumped>(FVS_SPLIT_STRANG_LUMPED, fused, 32, 2);
    RunCase<3, GodunovSplitting>(FVS_SPLIT_GODUNOV, fused, 16, 2);
    RunCase<3, StrangSplitting>(FVS_SPLIT_STRANG, fused, 16, 3);
    RunCase<3, StrangSplittingLumped>(FVS_SPLIT_STRANG_LUMPED, fused, 16, 2);
  }
  RejectsTooLargeDt();
  if (g_failures == 0) std::printf("all shim checks passed\n");
  return g_failures == 0 ? 0 : 1;
}
