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
#include <algorithm>
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
int fvso_level_enable_eb(void* h, int kbn_riemann);
double* fvso_level_eb_array(void* h, int box, int kind, int dir, int* lo, int* hi, int* ncomp);
int fvso_level_eb_finalize(void* h, int box, int fab_type);
int fvso_level_is_cut_cell(void* h, int box, unsigned char* out);
int fvso_tag_gradient(void* h, int box, int ncond, const int* comps, const double* tols,
                      int use_alpha, char* tags);
int fvso_tag_cutcells(void* h, int box, char* tags);
int fvso_tag_buffer(char* tags, int rank, const int* n, int width);
void* fvso_hierarchy_create(int nlevels, void** levels, const int* ratio);
void fvso_hierarchy_destroy(void* h);
int fvso_hierarchy_stable_dt(void* h, double* out);
int fvso_hierarchy_advance(void* h, double dt, int split_kind, double* too_large_dt,
                           double* sweep_dts, int* n_sweep_dts);
void* fvso_cf_create(void* fine, void* coarse, const int* ratio);
void fvso_cf_destroy(void* cf);
int fvso_cf_fill_ghost_layer_two_levels(void* cf, void* fine, void* coarse);
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
static void RunCase(int split_kind, bool fused, int ncells, int nsteps, bool tolerance = false) {
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
  // opt-in arithmetic of the fused sweep: within 1e-12 per step instead of bit-identical
  if (tolerance) context.SetArithmetic(FVS_ARITH_TOLERANCE);

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
    if (tolerance) {
      REQUIRE(std::fabs(stable.count() - odt) <= 1e-12 * odt);
    } else {
      REQUIRE(stable.count() == odt); // identical dt sequence
    }
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
  double max_diff = 0.0, max_ref = 0.0;
  for (int b = 0; b < 2; ++b) {
    const FabView v = context.ScratchView(0, b);
    fab.resize(v.size());
    context.DownloadScratch(0, b, fab.data());
    int lo[3], hi[3], nc;
    const double* o = fvso_level_scratch(oracle, b, lo, hi, &nc);
    for (std::size_t i = 0; i < fab.size(); ++i) {
      if (std::memcmp(&fab[i], &o[i], sizeof(double)) != 0) ++mismatches;
      max_diff = std::max(max_diff, std::fabs(fab[i] - o[i]));
      max_ref = std::max(max_ref, std::fabs(o[i]));
    }
  }
  if (tolerance) {
    std::printf("Rank %d split %d tolerance arithmetic: %d steps, rel Linf %.2e\n", Rank, split_kind,
                nsteps, max_diff / max_ref);
    REQUIRE(max_diff <= 1e-12 * nsteps * max_ref);
    REQUIRE(mismatches > 0); // otherwise the switch did nothing
    fvso_level_destroy(oracle);
    return;
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

// ---------------------------------------------------------------------------
// Cut cells through the shim: an inclined wall (half plane) in a 2-D box, the
// set-up of examples/AMReX/EB/2D/Schardin.cpp:87-161 in miniature
// (KbnCutCellMethod over MusclHancock<HLLEM, VanLeer>, HLL wall solver, 4w/3w
// ghost cells).  Geometry by sub-sampling the level set, the same arrays on
// both sides.
static void CutCellCase(bool fused) {
  constexpr int Rank = 2;
  cutcell::IntegratorContext::Options opts;
  opts.scratch_gcw = 8;
  opts.flux_gcw = 6;
  PerfectGas<Rank> eq;
  cutcell::IntegratorContext context(eq, BaseMethod::HllemLarrouturou, Reconstruct::MusclHancock,
                                     Limiter::VanLeer, opts);
  Box domain;
  domain.hi = {39, 27, 0};
  const std::array<double, 3> dx{0.15 / 40, 0.10 / 28, 1.0};
  context.DefineLevel(0, domain, dx, {0, 0, 0}, {domain});
  const int periodic[3] = {0, 0, 0};
  void* oracle = fvso_level_create(Rank, domain.lo.data(), domain.hi.data(), dx.data(), periodic,
                                   1, domain.lo.data(), domain.hi.data(), opts.scratch_gcw,
                                   opts.flux_gcw, eq.gamma);
  fvso_level_set_method(oracle, FVS_BASE_HLLEM_LARROUTUROU, FVS_RECON_MUSCL_HANCOCK,
                        FVS_LIMITER_VANLEER);
  REQUIRE(fvso_level_enable_eb(oracle, FVS_RIEMANN_HLL_EINFELDT) == 0);
  context.SetKbnRiemannSolver(cutcell::RiemannSolver::HllEinfeldt);
  for (int d = 0; d < Rank; ++d)
    for (int side = 0; side < 2; ++side) {
      context.AddBoundary(0, static_cast<Direction>(d), side, Boundary::Transmissive);
      fvso_level_add_bc(oracle, d, side, FVS_BC_TRANSMISSIVE);
    }
  const FabView v = context.ScratchView(0, 0);
  const int n[2] = {v.box.hi[0] - v.box.lo[0] + 1, v.box.hi[1] - v.box.lo[1] + 1};
  // fluid where phi > 0: below the inclined wall through (0.09, 0) with normal (0.6, 0.8)
  auto phi = [](double x, double y) { return -((x - 0.09) * 0.6 + (y - 0.06) * 0.8); };
  auto fraction = [&](double x0, double x1, double y0, double y1) {
    const int ns = 8;
    int inside = 0, total = 0;
    const int nx = x1 > x0 ? ns : 1, ny = y1 > y0 ? ns : 1;
    for (int b = 0; b < ny; ++b)
      for (int a = 0; a < nx; ++a) {
        const double x = x1 > x0 ? x0 + (a + 0.5) / ns * (x1 - x0) : x0;
        const double y = y1 > y0 ? y0 + (b + 0.5) / ns * (y1 - y0) : y0;
        inside += phi(x, y) > 0.0;
        ++total;
      }
    return double(inside) / total;
  };
  std::vector<double> alpha(std::size_t(n[0]) * n[1]), bcent(2 * alpha.size(), 0.0),
      bnorm(2 * alpha.size(), 0.0);
  std::vector<double> beta0(std::size_t(n[0] + 1) * n[1]), beta1(std::size_t(n[0]) * (n[1] + 1));
  for (int j = 0; j < n[1]; ++j)
    for (int i = 0; i < n[0]; ++i) {
      const double x0 = (v.box.lo[0] + i) * dx[0], y0 = (v.box.lo[1] + j) * dx[1];
      alpha[i + std::size_t(n[0]) * j] = fraction(x0, x0 + dx[0], y0, y0 + dx[1]);
    }
  auto covered = [&](int i, int j) {
    return i >= 0 && i < n[0] && j >= 0 && j < n[1] && alpha[i + std::size_t(n[0]) * j] == 0.0;
  };
  for (int j = 0; j < n[1]; ++j)
    for (int i = 0; i <= n[0]; ++i) {
      const double x = (v.box.lo[0] + i) * dx[0], y0 = (v.box.lo[1] + j) * dx[1];
      double b = fraction(x, x, y0, y0 + dx[1]);
      if (covered(i - 1, j) || covered(i, j)) b = 0.0;
      beta0[i + std::size_t(n[0] + 1) * j] = b;
    }
  for (int j = 0; j <= n[1]; ++j)
    for (int i = 0; i < n[0]; ++i) {
      const double x0 = (v.box.lo[0] + i) * dx[0], y = (v.box.lo[1] + j) * dx[1];
      double b = fraction(x0, x0 + dx[0], y, y);
      if (covered(i, j - 1) || covered(i, j)) b = 0.0;
      beta1[i + std::size_t(n[0]) * j] = b;
    }
  for (int j = 0; j < n[1]; ++j)
    for (int i = 0; i < n[0]; ++i) {
      const std::size_t c = i + std::size_t(n[0]) * j;
      if (alpha[c] > 0.0 && alpha[c] < 1.0) {
        const double xc = (v.box.lo[0] + i + 0.5) * dx[0], yc = (v.box.lo[1] + j + 0.5) * dx[1];
        const double dist = phi(xc, yc); // the wall normal (into the body) is (0.6, 0.8)
        bnorm[c] = 0.6;
        bnorm[c + alpha.size()] = 0.8;
        bcent[c] = std::min(0.5, std::max(-0.5, dist * 0.6 / dx[0]));
        bcent[c + alpha.size()] = std::min(0.5, std::max(-0.5, dist * 0.8 / dx[1]));
      }
    }
  cutcell::CutCellGeometry geom;
  geom.volume_fractions = alpha.data();
  geom.face_fractions = {beta0.data(), beta1.data(), nullptr};
  geom.boundary_centeroids = bcent.data();
  geom.boundary_normals = bnorm.data();
  geom.fab_type = FVS_FAB_SINGLEVALUED;
  context.SetEmbeddedBoundary(0, 0, geom);
  {
    int lo[3], hi[3], nc;
    std::memcpy(fvso_level_eb_array(oracle, 0, 0, 0, lo, hi, &nc), alpha.data(), alpha.size() * 8);
    std::memcpy(fvso_level_eb_array(oracle, 0, 1, 0, lo, hi, &nc), beta0.data(), beta0.size() * 8);
    std::memcpy(fvso_level_eb_array(oracle, 0, 1, 1, lo, hi, &nc), beta1.data(), beta1.size() * 8);
    std::memcpy(fvso_level_eb_array(oracle, 0, 2, 0, lo, hi, &nc), bcent.data(), bcent.size() * 8);
    std::memcpy(fvso_level_eb_array(oracle, 0, 3, 0, lo, hi, &nc), bnorm.data(), bnorm.size() * 8);
    REQUIRE(fvso_level_eb_finalize(oracle, 0, FVS_FAB_SINGLEVALUED) == 0);
  }
  // bit-exact cut-cell indexing
  std::vector<unsigned char> flags(alpha.size()), oflags(alpha.size());
  const long ncut = context.CutCellFlags(0, 0, flags.data());
  REQUIRE(fvso_level_is_cut_cell(oracle, 0, oflags.data()) == 0);
  REQUIRE(flags == oflags);
  REQUIRE(ncut > 10);
  // shock at x = 0.015 moving towards the wall
  std::vector<double> fab(v.size(), 0.0);
  const std::size_t sc = alpha.size();
  const double gm1_inv = 1.0 / (1.4 - 1.0);
  for (int j = 0; j < n[1]; ++j)
    for (int i = 0; i < n[0]; ++i) {
      const std::size_t c = i + std::size_t(n[0]) * j;
      if (alpha[c] == 0.0) continue;
      const double x = (v.box.lo[0] + i + 0.5) * dx[0];
      const double rho = 1.0, p = x < 0.015 ? 5 * 101325.0 : 101325.0;
      fab[c] = rho;
      fab[c + 3 * sc] = p * gm1_inv;
      fab[c + 4 * sc] = p;
      fab[c + 5 * sc] = std::sqrt(1.4 * p / rho);
    }
  context.UploadScratch(0, 0, fab.data());
  {
    int lo[3], hi[3], nc;
    std::memcpy(fvso_level_scratch(oracle, 0, lo, hi, &nc), fab.data(), fab.size() * 8);
  }
  context.FillGhostLayerSingleLevel(0);
  REQUIRE(fvso_fill_ghost_layer(oracle) == 0);
  HyperbolicMethod method{cutcell::FluxMethod{}, cutcell::TimeIntegrator{}, cutcell::Reconstruction{}};
  DimensionalSplitLevelIntegrator<Rank, StrangSplitting> integrator(context, method,
                                                                    StrangSplitting(), fused);
  for (int step = 0; step < 8; ++step) {
    const Duration stable = integrator.ComputeStableDt(0);
    double odt = std::numeric_limits<double>::max();
    for (int d = 0; d < Rank; ++d) {
      double t = 0.0;
      REQUIRE(fvso_compute_stable_dt(oracle, d, &t) == 0);
      odt = std::min(odt, t);
    }
    REQUIRE(stable.count() == odt);
    const Duration dt(0.4 * stable.count());
    REQUIRE(!integrator.AdvanceLevelNonRecursively(0, dt));
    double too_large = 0.0;
    REQUIRE(fvso_advance_hierarchy(oracle, dt.count(), FVS_SPLIT_STRANG, &too_large, nullptr,
                                   nullptr) == 0);
    context.SetCycles(context.GetCycles() + 1, Duration(context.GetTimePoint().count() + dt.count()));
    context.FillGhostLayerSingleLevel(0);
  }
  std::vector<double> got(v.size());
  context.DownloadScratch(0, 0, got.data());
  int lo[3], hi[3], nc;
  const double* o = fvso_level_scratch(oracle, 0, lo, hi, &nc);
  long long mismatches = 0;
  for (std::size_t i = 0; i < got.size(); ++i) mismatches += std::memcmp(&got[i], &o[i], 8) != 0;
  std::printf("cut cells %s: 8 steps, %ld cut cells, %lld mismatching doubles\n",
              fused ? "fused entry point" : "policies", ncut, mismatches);
  REQUIRE(mismatches == 0);
  // tagging: GradientDetector (EB variant) + TagCutCells + TagBuffer, bit-exact
  const int nv[3] = {40, 28, 1};
  std::vector<char> tags(40 * 28, 0), otags(40 * 28, 0);
  GradientDetector detector{{{0, 1e-2}, {4, 0.05}}, true};
  detector.TagCellsForRefinement(tags.data(), context, 0, 0);
  TagCutCells().TagCellsForRefinement(tags.data(), context, 0, 0);
  TagBuffer{4}.TagCellsForRefinement(tags.data(), context, Rank, {40, 28, 1});
  const int comps[2] = {0, 4};
  const double tols[2] = {1e-2, 0.05};
  REQUIRE(fvso_tag_gradient(oracle, 0, 2, comps, tols, 1, otags.data()) == 0);
  REQUIRE(fvso_tag_cutcells(oracle, 0, otags.data()) == 0);
  REQUIRE(fvso_tag_buffer(otags.data(), Rank, nv, 4) == 0);
  REQUIRE(tags == otags);
  long ntag = 0;
  for (char t : tags) ntag += t != 0;
  REQUIRE(ntag > 0);
  fvso_level_destroy(oracle);
}

// Two levels through SubcycleFineFirstSolver (fine-first subcycling, flux
// correction, conservative coarsening) against the oracle's hierarchy.
static void HierarchyCase() {
  constexpr int Rank = 2;
  IntegratorContext::Options opts;
  opts.scratch_gcw = 8;
  opts.flux_gcw = 6;
  PerfectGas<Rank> eq;
  IntegratorContext context(eq, BaseMethod::HllemLarrouturou, Reconstruct::MusclHancock,
                            Limiter::VanLeer, opts);
  Box d0, d1, f0;
  d0.hi = {31, 23, 0};
  d1.hi = {63, 47, 0};
  f0.lo = {16, 8, 0};
  f0.hi = {47, 39, 0};
  const std::array<double, 3> dx0{2.0 / 32, 2.0 / 24, 1.0}, dx1{1.0 / 32, 1.0 / 24, 1.0};
  context.DefineLevel(0, d0, dx0, {0, 0, 0}, {d0});
  context.DefineLevel(1, d1, dx1, {0, 0, 0}, {f0});
  REQUIRE(context.LevelExists(1) && !context.LevelExists(2));
  const int periodic[3] = {0, 0, 0};
  void* lv[2];
  lv[0] = fvso_level_create(Rank, d0.lo.data(), d0.hi.data(), dx0.data(), periodic, 1, d0.lo.data(),
                            d0.hi.data(), 8, 6, eq.gamma);
  lv[1] = fvso_level_create(Rank, d1.lo.data(), d1.hi.data(), dx1.data(), periodic, 1, f0.lo.data(),
                            f0.hi.data(), 8, 6, eq.gamma);
  for (int l = 0; l < 2; ++l) {
    fvso_level_set_method(lv[l], FVS_BASE_HLLEM_LARROUTUROU, FVS_RECON_MUSCL_HANCOCK,
                          FVS_LIMITER_VANLEER);
    for (int d = 0; d < Rank; ++d)
      for (int side = 0; side < 2; ++side) {
        context.AddBoundary(l, static_cast<Direction>(d), side, Boundary::Reflective);
        fvso_level_add_bc(lv[l], d, side, FVS_BC_REFLECTIVE);
      }
    const FabView v = context.ScratchView(l, 0);
    std::vector<double> fab;
    InitialData(Rank, v, l == 0 ? dx0 : dx1, fab);
    context.UploadScratch(l, 0, fab.data());
    int lo[3], hi[3], nc;
    std::memcpy(fvso_level_scratch(lv[l], 0, lo, hi, &nc), fab.data(), fab.size() * 8);
  }
  const int ratio[3] = {2, 2, 1};
  void* hier = fvso_hierarchy_create(2, lv, ratio);
  void* cf = fvso_cf_create(lv[1], lv[0], ratio);
  context.FillGhostLayerSingleLevel(0);
  context.FillGhostLayerTwoLevels(1, 0);
  REQUIRE(fvso_fill_ghost_layer(lv[0]) == 0);
  REQUIRE(fvso_cf_fill_ghost_layer_two_levels(cf, lv[1], lv[0]) == 0);
  SubcycleFineFirstSolver<StrangSplitting> solver(context);
  for (int step = 0; step < 3; ++step) {
    const Duration stable = solver.ComputeStableDt();
    double odt = 0.0;
    REQUIRE(fvso_hierarchy_stable_dt(hier, &odt) == 0);
    REQUIRE(stable.count() == odt);
    REQUIRE(!solver.AdvanceHierarchy(Duration(0.5 * odt)));
    double too_large = 0.0;
    REQUIRE(fvso_hierarchy_advance(hier, 0.5 * odt, FVS_SPLIT_STRANG, &too_large, nullptr,
                                   nullptr) == 0);
  }
  long long mismatches = 0;
  for (int l = 0; l < 2; ++l) {
    const FabView v = context.ScratchView(l, 0);
    std::vector<double> got(v.size());
    context.DownloadScratch(l, 0, got.data());
    int lo[3], hi[3], nc;
    const double* o = fvso_level_scratch(lv[l], 0, lo, hi, &nc);
    for (std::size_t i = 0; i < got.size(); ++i) mismatches += std::memcmp(&got[i], &o[i], 8) != 0;
  }
  std::printf("two-level hierarchy: 3 coarse steps, %lld mismatching doubles\n", mismatches);
  REQUIRE(mismatches == 0);
  // a too large step is reported, not thrown, and scaled to the coarse level
  AdvanceResult r = solver.AdvanceHierarchy(Duration(50.0 * solver.ComputeStableDt().count()));
  REQUIRE(bool(r));
  fvso_cf_destroy(cf);
  fvso_hierarchy_destroy(hier);
  fvso_level_destroy(lv[0]);
  fvso_level_destroy(lv[1]);
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
  RunCase<2, StrangSplitting>(FVS_SPLIT_STRANG, true, 32, 3, true);
  RunCase<3, StrangSplitting>(FVS_SPLIT_STRANG, true, 16, 3, true);
  RejectsTooLargeDt();
  CutCellCase(false);
  CutCellCase(true);
  HierarchyCase();
  if (g_failures == 0) std::printf("all shim checks passed\n");
  return g_failures == 0 ? 0 : 1;
}
