// fub_b200/HyperbolicMethod.hpp -- C++17 host-side mirror of the reference's
// policy surface for the split-sweep path, forwarding to the C ABI of
// libfvs_b200.so (include/fvs_b200.h).  Header only; no CUDA, no torch.
//
// What it mirrors (paths relative to maikel/FiniteVolumeSolver):
//   fub::Direction / fub::Duration           include/fub/Direction.hpp:29, include/fub/Duration.hpp:32
//   fub::HyperbolicMethod<Context>           include/fub/HyperbolicMethod.hpp:155-159
//   FluxMethod policy                        include/fub/AMReX/FluxMethodAdapter.hpp:40-56
//       Duration ComputeStableDt(Context&, int level, Direction)
//       void ComputeNumericFluxes(Context&, int level, Duration dt, Direction)
//       int GetStencilWidth() const
//   TimeIntegrator policy                    include/fub/AMReX/TimeIntegrator.hpp  (src/AMReX/TimeIntegrator.cpp:29-67)
//       void UpdateConservatively(Context&, int level, Duration dt, Direction)
//   Reconstruction policy                    include/fub/AMReX/Reconstruction.hpp:44-60
//       void CompleteFromCons(Context&, int level, Duration dt)
//   DimensionalSplitLevelIntegrator          include/fub/solver/DimensionalSplitLevelIntegrator.hpp:204-276
//   split methods                            src/split_method/{Godunov,Strang,StrangSplittingLumped}Splitting.cpp
//   TimeStepTooLarge                         include/fub/TimeStepError.hpp:71-78
//
// The policies are value types holding nothing but their configuration; the
// context is a ref-counted handle (the reference deep-copies policies and the
// whole solver every step, include/fub/RunSimulation.hpp:177-181, so a GPU
// policy must be cheap to copy -- SURVEY 8b "Ownership").
//
// Errors: a negative status of the C ABI becomes std::runtime_error carrying
// fvs_last_error() (the reference's policies throw std::runtime_error too);
// a CFL violation is NOT an exception, it is returned as TimeStepTooLarge
// like DimensionalSplitLevelIntegrator.hpp:243-245 does.
#ifndef FUB_B200_HYPERBOLIC_METHOD_HPP
#define FUB_B200_HYPERBOLIC_METHOD_HPP

#include <algorithm>
#include <array>
#include <chrono>
#include <cstddef>
#include <limits>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../fvs_b200.h"

namespace fub {

// When this header is compiled next to the reference's own headers their
// definitions win (same include guards as the reference).
#ifndef FUB_DIRECTION_HPP
#define FUB_DIRECTION_HPP
enum class Direction : std::size_t { X, Y, Z };
enum class Side : int { Lower, Upper };
#endif
#ifndef FUB_DURATION_HPP
#define FUB_DURATION_HPP
struct Duration : std::chrono::duration<double> {
  using duration::duration;
  Duration(const std::chrono::duration<double>& other) noexcept : duration(other) {}
};
#endif

namespace b200 {

/// fub::TimeStepTooLarge, include/fub/TimeStepError.hpp:71-78
struct TimeStepTooLarge {
  Duration dt;
  int level;
};
/// Result<void, TimeStepTooLarge>: empty optional == success
using AdvanceResult = std::optional<TimeStepTooLarge>;

inline void Check(int status) {
  if (status < 0) throw std::runtime_error(std::string("fvs_b200: ") + fvs_last_error());
}

enum class BaseMethod : int {
  HllemLarrouturou = FVS_BASE_HLLEM_LARROUTUROU, ///< perfect_gas::HllemMethod<Eq, true>
  Hllem = FVS_BASE_HLLEM,                        ///< perfect_gas::HllemMethod<Eq, false>
  HllEinfeldt = FVS_BASE_HLL_EINFELDT,           ///< HllMethod<Eq, EinfeldtSignalVelocities<Eq>>
  GodunovExact = FVS_BASE_GODUNOV_EXACT          ///< GodunovMethod<Eq, ExactRiemannSolver<Eq>>
};
enum class Reconstruct : int {
  FirstOrder = FVS_RECON_FIRST_ORDER,
  MusclHancock = FVS_RECON_MUSCL_HANCOCK, ///< MusclHancockMethod<Eq, Base, Limiter>
  MusclHancockPrim = FVS_RECON_MUSCL_HANCOCK_PRIM, ///< perfect_gas::MusclHancockPrim<Rank>
  MusclHancockChar = FVS_RECON_MUSCL_HANCOCK_CHAR, ///< perfect_gas::MusclHancockCharacteristic<Rank>
  /// MusclHancock2<Eq, XGradient<Eq, CentralDifferenceGradient<L>>, XReconstruction<Eq>, Base>
  /// (include/fub/flux_method/MusclHancockMethod2.hpp); the limiter is a Limiter2 value
  Mh2Conservative = FVS_RECON_MH2_CONSERVATIVE,
  Mh2Primitive = FVS_RECON_MH2_PRIMITIVE,
  Mh2Characteristics = FVS_RECON_MH2_CHARACTERISTICS
};
/// include/fub/flux_method/Gradient.hpp:31-59
enum class Limiter2 : int {
  None = FVS_LIMITER2_NONE,
  Upwind = FVS_LIMITER2_UPWIND,
  MinMod = FVS_LIMITER2_MINMOD,
  VanLeer = FVS_LIMITER2_VANLEER
};
enum class Limiter : int {
  None = FVS_LIMITER_NONE,
  MinMod = FVS_LIMITER_MINMOD,
  VanLeer = FVS_LIMITER_VANLEER
};
enum class Boundary : int {
  Transmissive = FVS_BC_TRANSMISSIVE,
  Reflective = FVS_BC_REFLECTIVE
};

/// fub::PerfectGas<Rank>, include/fub/equations/PerfectGas.hpp:199-222
template <int Rank> struct PerfectGas {
  static constexpr int Rank_v = Rank;
  static constexpr int n_complete = Rank + 4;
  static constexpr int n_conservative = Rank + 2;
  double gamma{1.4};
  double gamma_minus_1_inv{1.0 / (gamma - 1.0)};
};

/// Index box [lo, hi] (inclusive, like amrex::Box); unused dims hold 0.
struct Box {
  std::array<int, 3> lo{0, 0, 0};
  std::array<int, 3> hi{0, 0, 0};
  long long NumCells() const {
    long long n = 1;
    for (int d = 0; d < 3; ++d) n *= (hi[d] - lo[d] + 1);
    return n;
  }
};

/// Non-owning description of a host array in FAB layout.
struct FabView {
  Box box;
  int ncomp{0};
  std::size_t size() const { return static_cast<std::size_t>(box.NumCells()) * ncomp; }
};

/// The part of fub::amrex::IntegratorContext the three policies use
/// (src/AMReX/IntegratorContext.cpp:229-236,292-345,347-359,422-449), backed by
/// device memory of one fvs_ctx.  Copies share the context.
class IntegratorContext {
public:
  struct Options {
    int scratch_gcw{4}; ///< IntegratorContextOptions::scratch_gcw
    int flux_gcw{2};    ///< IntegratorContextOptions::flux_gcw
    int regrid_frequency{1};
    int device{-1};
    int comm_rank{0};
    int comm_size{1};
  };

  template <int Rank>
  IntegratorContext(const PerfectGas<Rank>& eq, BaseMethod base, Reconstruct recon,
                    Limiter limiter, const Options& opts = Options())
      : rank_(Rank), options_(opts) {
    fvs_config cfg{};
    cfg.rank = Rank;
    cfg.gamma = eq.gamma;
    cfg.base_method = static_cast<int>(base);
    cfg.reconstruction = static_cast<int>(recon);
    cfg.limiter = static_cast<int>(limiter);
    cfg.device = opts.device;
    cfg.comm_rank = opts.comm_rank;
    cfg.comm_size = opts.comm_size;
    fvs_ctx* raw = nullptr;
    Check(fvs_create(&cfg, &raw));
    ctx_ = std::shared_ptr<fvs_ctx>(raw, [](fvs_ctx* p) { fvs_destroy(p); });
    Check(fvs_set_regrid_frequency(raw, opts.regrid_frequency));
  }

  /// IntegratorContext::ResetHierarchyConfiguration for one level:
  /// `boxes` is the BoxArray, `owner` the DistributionMapping.
  void DefineLevel(int level, const Box& domain, const std::array<double, 3>& dx,
                   const std::array<int, 3>& periodic, const std::vector<Box>& boxes,
                   const std::vector<int>& owner = {}) {
    std::vector<std::array<int, 3>> lo(boxes.size()), hi(boxes.size());
    for (std::size_t b = 0; b < boxes.size(); ++b) {
      lo[b] = boxes[b].lo;
      hi[b] = boxes[b].hi;
    }
    Check(fvs_level_define(ctx_.get(), level, domain.lo.data(), domain.hi.data(), dx.data(),
                           periodic.data(), static_cast<int>(boxes.size()),
                           reinterpret_cast<const int(*)[3]>(lo.data()),
                           reinterpret_cast<const int(*)[3]>(hi.data()),
                           owner.empty() ? nullptr : owner.data(), options_.scratch_gcw,
                           options_.flux_gcw));
    dx_ = dx;
    nboxes_ = static_cast<int>(boxes.size());
  }
  void AddBoundary(int level, Direction dir, int side, Boundary kind) {
    Check(fvs_level_add_boundary(ctx_.get(), level, static_cast<int>(dir), side,
                                 static_cast<int>(kind)));
  }

  /// Arithmetic of the fused sweep: FVS_ARITH_IEEE (default, bit-identical to the
  /// reference's array code path) or FVS_ARITH_TOLERANCE (within 1e-12 per step).
  void SetArithmetic(int arithmetic) { Check(fvs_set_arithmetic(Handle(), arithmetic)); }

  int Rank() const { return rank_; }
  int NumBoxes() const { return nboxes_; }
  double GetDx(int /*level*/, Direction dir) const { return dx_[static_cast<int>(dir)]; }
  fvs_ctx* Handle() const { return ctx_.get(); }

  /// Geometry of GetScratch(level)[box] / GetFluxes(level, dir)[box]
  FabView ScratchView(int level, int box) const {
    FabView v;
    Check(fvs_state_box(ctx_.get(), level, box, v.box.lo.data(), v.box.hi.data(), &v.ncomp));
    return v;
  }
  FabView FluxView(int level, int box, Direction dir) const {
    FabView v;
    Check(fvs_flux_box(ctx_.get(), level, box, static_cast<int>(dir), v.box.lo.data(),
                       v.box.hi.data(), &v.ncomp));
    return v;
  }
  /// host FAB <-> device scratch (what a MultiFab-backed caller does at the boundary)
  void UploadScratch(int level, int box, const double* fab) {
    Check(fvs_upload_state(ctx_.get(), level, box, fab));
  }
  void DownloadScratch(int level, int box, double* fab) const {
    Check(fvs_download_state(ctx_.get(), level, box, fab));
  }
  void UploadFluxes(int level, int box, Direction dir, const double* fab) {
    Check(fvs_upload_fluxes(ctx_.get(), level, box, static_cast<int>(dir), fab));
  }
  void DownloadFluxes(int level, int box, Direction dir, double* fab) const {
    Check(fvs_download_fluxes(ctx_.get(), level, box, static_cast<int>(dir), fab));
  }

  /// src/AMReX/IntegratorContext.cpp:422-449 and :347-359
  void FillGhostLayerSingleLevel(int level) { Check(fvs_fill_ghost_layer(ctx_.get(), level)); }
  void ApplyBoundaryCondition(int level, Direction dir) {
    Check(fvs_apply_boundary_condition(ctx_.get(), level, static_cast<int>(dir)));
  }

  long GetCycles(int /*level*/ = 0) const { return fvs_cycles(ctx_.get()); }
  Duration GetTimePoint(int /*level*/ = 0) const { return Duration(fvs_time_point(ctx_.get())); }
  void SetCycles(long cycles, Duration time_point) {
    Check(fvs_set_cycles(ctx_.get(), cycles, time_point.count()));
  }
  /// fub::MinAll over the ranks that share the level (src/ext/Mpi.cpp:54-59)
  Duration MinAll(Duration local) const {
    double v = local.count();
    Check(fvs_min_all(ctx_.get(), &v));
    return Duration(v);
  }
  long KernelLaunches() const { return fvs_kernel_launches(ctx_.get()); }

  // ---- coarse-fine (src/AMReX/IntegratorContext.cpp:380-533) -----------------
  bool LevelExists(int level) const {
    int lo[3], hi[3], nc;
    return fvs_state_box(ctx_.get(), level, 0, lo, hi, &nc) == FVS_OK;
  }
  void SetRefineRatio(const std::array<int, 3>& ratio) {
    Check(fvs_set_refine_ratio(ctx_.get(), ratio.data()));
  }
  void FillGhostLayerTwoLevels(int fine, int coarse) {
    Check(fvs_fill_ghost_layer_two_levels(ctx_.get(), fine, coarse));
  }
  void ResetCoarseFineFluxes(int fine, int coarse) {
    Check(fvs_reset_coarse_fine_fluxes(ctx_.get(), fine, coarse));
  }
  void AccumulateCoarseFineFluxes(int level, double scale, Direction dir) {
    Check(fvs_accumulate_coarse_fine_fluxes(ctx_.get(), level, scale, static_cast<int>(dir)));
  }
  void ApplyFluxCorrection(int fine, int coarse, Duration dt) {
    Check(fvs_apply_flux_correction(ctx_.get(), fine, coarse, dt.count()));
  }
  void CoarsenConservatively(int fine, int coarse) {
    Check(fvs_coarsen_conservatively(ctx_.get(), fine, coarse));
  }

private:
  int rank_;
  Options options_;
  std::shared_ptr<fvs_ctx> ctx_;
  std::array<double, 3> dx_{1.0, 1.0, 1.0};
  int nboxes_{0};
};

/// FluxMethod policy: drop-in for fub::amrex::FluxMethodAdapter<Tag, FM>.
/// The numerical method itself is a property of the context's device kernels
/// (fvs_set_method); the policy is an empty, trivially copyable value.
struct FluxMethod {
  Duration ComputeStableDt(IntegratorContext& context, int level, Direction dir) {
    double dt = std::numeric_limits<double>::max();
    Check(fvs_compute_stable_dt(context.Handle(), level, static_cast<int>(dir), &dt));
    return Duration(dt);
  }
  void ComputeNumericFluxes(IntegratorContext& context, int level, Duration dt, Direction dir) {
    Check(fvs_compute_numeric_fluxes(context.Handle(), level, dt.count(),
                                     static_cast<int>(dir)));
  }
  int GetStencilWidth() const { return stencil_width; }
  /// cutcell::FluxMethod::PreAdvanceHierarchy (include/fub/AMReX/cutcell/FluxMethod.hpp:108-133);
  /// a no-op on levels without an embedded boundary
  void PreAdvanceHierarchy(IntegratorContext& context) {
    Check(fvs_pre_advance_hierarchy(context.Handle()));
  }
  int stencil_width{2};
};

/// TimeIntegrator policy: drop-in for fub::amrex::ForwardIntegrator<Tag>.
struct TimeIntegrator {
  void UpdateConservatively(IntegratorContext& context, int level, Duration dt, Direction dir) {
    Check(fvs_update_conservatively(context.Handle(), level, dt.count(),
                                    static_cast<int>(dir)));
  }
};

/// Reconstruction policy: drop-in for fub::amrex::Reconstruction<Tag, Eq>.
struct Reconstruction {
  void CompleteFromCons(IntegratorContext& context, int level, Duration /*dt*/) {
    Check(fvs_complete_from_cons(context.Handle(), level));
  }
};

/// fub::HyperbolicMethod<IntegratorContext>, include/fub/HyperbolicMethod.hpp:155-159
struct HyperbolicMethod {
  FluxMethod flux_method;
  TimeIntegrator time_integrator;
  Reconstruction reconstruction;
};

/// fub::MakeSplitDirections, include/fub/split_method/SplittingMethod.hpp:115-126
template <int Rank>
std::array<Direction, Rank> MakeSplitDirections(int cycle, std::pair<int, int> subcycle) {
  std::array<Direction, Rank> dirs;
  for (int d = 0; d < Rank; ++d) dirs[d] = static_cast<Direction>(d);
  if ((cycle % 2 + subcycle.first) % 2 == 1) std::reverse(dirs.begin(), dirs.end());
  return dirs;
}

/// Splitting methods: Advance(dt, op...) applies the directional operators
/// with the reference's order and fractions of dt.
struct GodunovSplitting { // src/split_method/GodunovSplitting.cpp:25-34
  template <typename Op> AdvanceResult Advance(Duration dt, std::vector<Op>& ops) const {
    for (Op& op : ops) {
      if (AdvanceResult r = op(dt)) return r;
    }
    return std::nullopt;
  }
};
struct StrangSplitting { // src/split_method/StrangSplitting.cpp:25-69
  template <typename Op> AdvanceResult Advance(Duration dt, std::vector<Op>& ops) const {
    const Duration half(0.5 * dt.count());
    const std::size_t n = ops.size();
    if (n == 1) return ops[0](dt);
    for (std::size_t i = 0; i < n; ++i) {
      if (AdvanceResult r = ops[i](half)) return r;
    }
    for (std::size_t i = n; i-- > 0;) {
      if (AdvanceResult r = ops[i](half)) return r;
    }
    return std::nullopt;
  }
};
struct StrangSplittingLumped { // src/split_method/StrangSplittingLumped.cpp:25-39
  template <typename Op> AdvanceResult Advance(Duration dt, std::vector<Op>& ops) const {
    const Duration half(0.5 * dt.count());
    const std::size_t n = ops.size();
    if (n == 1) return ops[0](dt);
    if (n == 2) {
      if (AdvanceResult r = ops[0](half)) return r;
      if (AdvanceResult r = ops[1](dt)) return r;
      return ops[0](half);
    }
    // 3-D: (X half, Y half), Z full, (Y half, X half) lumped as in the reference
    if (AdvanceResult r = ops[0](half)) return r;
    if (AdvanceResult r = ops[1](half)) return r;
    if (AdvanceResult r = ops[2](dt)) return r;
    if (AdvanceResult r = ops[1](half)) return r;
    return ops[0](half);
  }
};

/// fub::DimensionalSplitLevelIntegrator<Rank, Context, SplitMethod>
/// (include/fub/solver/DimensionalSplitLevelIntegrator.hpp:204-276) over the
/// three policies.  `fused = true` replaces the four policy calls of a sweep
/// by the single-pass fvs_sweep kernel (same results, one read + one write of
/// the state).
template <int Rank, typename SplitMethod = GodunovSplitting>
class DimensionalSplitLevelIntegrator {
public:
  DimensionalSplitLevelIntegrator(IntegratorContext context, HyperbolicMethod method,
                                  SplitMethod split = SplitMethod(), bool fused = false)
      : context_(std::move(context)), method_(method), split_(split), fused_(fused) {}

  IntegratorContext& GetContext() { return context_; }
  HyperbolicMethod& GetHyperbolicMethod() { return method_; }

  /// :204-219
  Duration ComputeStableDt(int level) {
    Duration min_dt(std::numeric_limits<double>::max());
    for (int d = 0; d < Rank; ++d) {
      const Direction dir = static_cast<Direction>(d);
      min_dt = std::min(min_dt, method_.flux_method.ComputeStableDt(context_, level, dir));
    }
    return min_dt;
  }

  /// :221-276
  AdvanceResult AdvanceLevelNonRecursively(int level, Duration dt,
                                           std::pair<int, int> subcycle = {0, 1}) {
    struct Op {
      DimensionalSplitLevelIntegrator* self;
      int level;
      Direction dir;
      int count_split_steps;
      AdvanceResult operator()(Duration split_dt) {
        IntegratorContext& ctx = self->context_;
        if (count_split_steps > 0) ctx.ApplyBoundaryCondition(level, dir);
        count_split_steps += 1;
        if (self->fused_) {
          double level_dt = 0.0;
          const int rc = fvs_sweep(ctx.Handle(), level, split_dt.count(),
                                   static_cast<int>(dir), &level_dt);
          Check(rc);
          if (rc == FVS_TIME_STEP_TOO_LARGE) return TimeStepTooLarge{Duration(level_dt), level};
          return std::nullopt;
        }
        HyperbolicMethod& m = self->method_;
        const Duration level_dt = m.flux_method.ComputeStableDt(ctx, level, dir);
        if (level_dt < split_dt) return TimeStepTooLarge{level_dt, level};
        m.flux_method.ComputeNumericFluxes(ctx, level, split_dt, dir);
        m.time_integrator.UpdateConservatively(ctx, level, split_dt, dir);
        m.reconstruction.CompleteFromCons(ctx, level, split_dt);
        // DimensionalSplitLevelIntegrator.hpp:262-264 (a no-op on a level without registers;
        // fvs_sweep does the same for the fused path)
        ctx.AccumulateCoarseFineFluxes(level, split_dt.count(), dir);
        return std::nullopt;
      }
    };
    // cutcell::IntegratorContext::PreAdvanceLevel (src/AMReX/cutcell/IntegratorContext.cpp:802-806)
    method_.flux_method.PreAdvanceHierarchy(context_);
    const auto dirs =
        MakeSplitDirections<Rank>(static_cast<int>(context_.GetCycles(0)), subcycle);
    std::vector<Op> ops;
    for (Direction dir : dirs) ops.push_back(Op{this, level, dir, 0});
    return split_.Advance(dt, ops);
  }

private:
  IntegratorContext context_;
  HyperbolicMethod method_;
  SplitMethod split_;
  bool fused_;
};

/// fub::SubcycleFineFirstSolver over the device hierarchy
/// (include/fub/solver/SubcycleFineFirstSolver.hpp:100-193): stable dt over all
/// levels and one coarse step with fine-first subcycling, flux correction and
/// conservative coarsening; TimeStepTooLarge leaves every level untouched.
template <typename SplitKind> struct SplitKindOf;
template <> struct SplitKindOf<GodunovSplitting> {
  static constexpr int value = FVS_SPLIT_GODUNOV;
};
template <> struct SplitKindOf<StrangSplitting> {
  static constexpr int value = FVS_SPLIT_STRANG;
};
template <> struct SplitKindOf<StrangSplittingLumped> {
  static constexpr int value = FVS_SPLIT_STRANG_LUMPED;
};
template <typename SplitMethod = GodunovSplitting> class SubcycleFineFirstSolver {
public:
  explicit SubcycleFineFirstSolver(IntegratorContext context) : context_(std::move(context)) {}
  IntegratorContext& GetContext() { return context_; }
  Duration ComputeStableDt() {
    double dt = 0.0;
    Check(fvs_compute_stable_dt_hierarchy(context_.Handle(), &dt));
    return Duration(dt);
  }
  AdvanceResult AdvanceHierarchy(Duration dt) {
    double failure = 0.0;
    const int rc = fvs_advance_hierarchy(context_.Handle(), dt.count(),
                                         SplitKindOf<SplitMethod>::value, &failure, nullptr, nullptr);
    Check(rc);
    if (rc == FVS_TIME_STEP_TOO_LARGE) return TimeStepTooLarge{Duration(failure), 0};
    return std::nullopt;
  }

private:
  IntegratorContext context_;
};

// ---- embedded boundary -------------------------------------------------------
namespace cutcell {

/// The arrays of fub::CutCellData<Rank> that come from amrex::EBFArrayBoxFactory
/// (include/fub/CutCellData.hpp:48-75, src/AMReX/cutcell/PatchHierarchy.cpp:507-547);
/// the shielded / unshielded fractions are derived on the device.
struct CutCellGeometry {
  const double* volume_fractions{nullptr};            ///< getVolFrac, scratch box
  std::array<const double*, 3> face_fractions{};      ///< getAreaFrac()[d], one more face along d
  const double* boundary_centeroids{nullptr};         ///< getBndryCent, Rank components
  const double* boundary_normals{nullptr};            ///< getBndryNormal, Rank components
  int fab_type{FVS_FAB_SINGLEVALUED};                 ///< amrex::FabType of the box
};
enum class RiemannSolver : int { HllEinfeldt = FVS_RIEMANN_HLL_EINFELDT, Exact = FVS_RIEMANN_EXACT };

/// fub::amrex::cutcell::IntegratorContext: the same policies, cut-cell adapters
/// once a box has been given its geometry.
class IntegratorContext : public ::fub::b200::IntegratorContext {
public:
  using ::fub::b200::IntegratorContext::IntegratorContext;
  void SetKbnRiemannSolver(RiemannSolver rs) {
    Check(fvs_set_cutcell_riemann_solver(Handle(), static_cast<int>(rs)));
  }
  /// PatchHierarchy::GetCutCellData(level, mfi) handed to the device
  void SetEmbeddedBoundary(int level, int box, const CutCellGeometry& g) {
    const double* beta[3] = {g.face_fractions[0], g.face_fractions[1], g.face_fractions[2]};
    Check(fvs_level_set_eb(Handle(), level, box, g.volume_fractions, beta, g.boundary_centeroids,
                           g.boundary_normals, g.fab_type));
  }
  /// fub::IsCutCell for every cell of the scratch box; returns the number of cut cells
  long CutCellFlags(int level, int box, unsigned char* flags) {
    const long n = fvs_cut_cell_flags(Handle(), level, box, flags);
    if (n < 0) throw std::runtime_error(fvs_last_error());
    return n;
  }
};
using FluxMethod = ::fub::b200::FluxMethod;         ///< KbnCutCellMethod behind cutcell::FluxMethod
using TimeIntegrator = ::fub::b200::TimeIntegrator; ///< cutcell::TimeIntegrator
using Reconstruction = ::fub::b200::Reconstruction; ///< cutcell::Reconstruction

} // namespace cutcell

// ---- tagging ---------------------------------------------------------------------
/// fub::GradientDetector(equation, {projection, tolerance}...) with projections
/// given as component indices of the Complete state
/// (include/fub/tagging_method/GradientDetector.hpp:156-232).
struct GradientDetector {
  std::vector<std::pair<int, double>> conditions;
  bool use_volume_fractions{false};
  void TagCellsForRefinement(char* tags, IntegratorContext& context, int level, int box) const {
    std::vector<int> comps;
    std::vector<double> tols;
    for (const auto& c : conditions) {
      comps.push_back(c.first);
      tols.push_back(c.second);
    }
    Check(fvs_tag_gradient(context.Handle(), level, box, static_cast<int>(comps.size()),
                           comps.data(), tols.data(), use_volume_fractions ? 1 : 0, tags));
  }
};
/// fub::TagCutCells (src/tagging_method/TagCutCells.cpp:27-35)
struct TagCutCells {
  void TagCellsForRefinement(char* tags, IntegratorContext& context, int level, int box) const {
    Check(fvs_tag_cutcells(context.Handle(), level, box, tags));
  }
};
/// fub::TagBuffer (src/tagging_method/TagBuffer.cpp:27-86) on a dense view
struct TagBuffer {
  int buffer_width{1};
  void TagCellsForRefinement(char* tags, IntegratorContext& context, int rank,
                             const std::array<int, 3>& extents) const {
    Check(fvs_tag_buffer(context.Handle(), rank, tags, extents.data(), buffer_width));
  }
};

} // namespace b200
} // namespace fub

#endif
