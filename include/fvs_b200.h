/* fvs_b200.h -- C ABI of libfvs_b200.so, the B200 (sm_100a) split-sweep
 * finite-volume patch integrator.
 *
 * This is the drop-in boundary: plain C, plain pointers
 * and sizes, no C++/torch types.  Every entry point names the reference
 * interface it replaces (paths relative to the reference repository
 * maikel/FiniteVolumeSolver).  All functions return 0 on success and a
 * negative value on error (text via fvs_last_error()); nothing throws across
 * the boundary.  A context is driven by one host thread, like the reference's
 * IntegratorContext (one driver thread per MPI rank).  Device work is
 * stream-ordered on the context's stream; a call synchronises only where it
 * returns a host value (dt) or fills a host buffer.
 *
 * Array layout ("FAB layout", include/fub/core/mdspan.hpp:731-732,
 * include/fub/AMReX/ViewFArrayBox.hpp:74-92): Fortran order, x fastest, then
 * y, z, component slowest; the box of a state array includes its ghost cells.
 * State components (include/fub/equations/PerfectGas.hpp:40-45,132-137):
 *   rho, momentum[rank], energy, pressure, speed_of_sound   (rank + 4)
 * Flux components: rho, momentum[rank], energy               (rank + 2)
 * Index triples are always int[3]; unused dimensions hold lo = hi = 0.
 */
#ifndef FVS_B200_H
#define FVS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FVS_VERSION 1

/* Base (Riemann) method.  src/equations/perfect_gas/HllemMethod.hpp,
 * include/fub/flux_method/HllMethod.hpp + EinfeldtSignalVelocities.hpp,
 * include/fub/flux_method/GodunovMethod.hpp + ExactRiemannSolver.hpp */
enum {
  FVS_BASE_HLLEM_LARROUTUROU = 0,
  FVS_BASE_HLLEM = 1,
  FVS_BASE_HLL_EINFELDT = 2,
  FVS_BASE_GODUNOV_EXACT = 3
};
/* include/fub/flux_method/MusclHancockMethod.hpp:189-571;
 * perfect_gas::MusclHancockPrim (src/equations/perfect_gas/MusclHancockPrimMethod.hpp:25-333):
 * primitive-variable MUSCL-Hancock with its own limiter, base method HLLEM
 * (Larrouturou) only, the limiter argument is ignored;
 * perfect_gas::MusclHancockCharacteristic (MusclHancockCharMethod,
 * src/equations/perfect_gas/MusclHancockCharacteristicMethod.hpp:28-359): characteristic
 * amplitudes limited with VanLeer, HLLEM (Larrouturou) only */
enum {
  FVS_RECON_FIRST_ORDER = 0,
  FVS_RECON_MUSCL_HANCOCK = 1,
  FVS_RECON_MUSCL_HANCOCK_PRIM = 2,
  FVS_RECON_MUSCL_HANCOCK_CHAR = 3,
  /* include/fub/flux_method/MusclHancockMethod2.hpp:35-239 MusclHancock2<Equation, Gradient,
   * Reconstruction, Base> with the gradient and the reconstruction in the same variables
   * (include/fub/flux_method/Gradient.hpp, Reconstruct.hpp; the three combinations of
   * include/fub/AMReX/FluxMethodFactory.hpp:161-176); `limiter` is then a FVS_LIMITER2_*
   * value, any base method */
  FVS_RECON_MH2_CONSERVATIVE = 4,
  FVS_RECON_MH2_PRIMITIVE = 5,
  FVS_RECON_MH2_CHARACTERISTICS = 6
};
/* limiters of CentralDifferenceGradient<Limiter>, include/fub/flux_method/Gradient.hpp:31-59
 * (NoLimiter2, UpwindLimiter, MinModLimiter, VanLeerLimiter) */
enum {
  FVS_LIMITER2_NONE = 0,
  FVS_LIMITER2_UPWIND = 1,
  FVS_LIMITER2_MINMOD = 2,
  FVS_LIMITER2_VANLEER = 3
};
/* include/fub/flux_method/MusclHancockMethod.hpp:31-187 */
enum { FVS_LIMITER_NONE = 0, FVS_LIMITER_MINMOD = 1, FVS_LIMITER_VANLEER = 2 };
/* src/AMReX/boundary_condition/TransmissiveBoundary.cpp,
 * include/fub/AMReX/boundary_condition/ReflectiveBoundary.hpp */
enum { FVS_BC_TRANSMISSIVE = 0, FVS_BC_REFLECTIVE = 1 };
/* src/split_method/{Godunov,Strang,StrangLumped}Splitting.cpp */
enum { FVS_SPLIT_GODUNOV = 0, FVS_SPLIT_STRANG = 1, FVS_SPLIT_STRANG_LUMPED = 2 };

/* amrex::FabType of a box (include/fub/AMReX/cutcell/FluxMethod.hpp:180,205) */
enum { FVS_FAB_COVERED = -1, FVS_FAB_REGULAR = 0, FVS_FAB_SINGLEVALUED = 1 };
/* RiemannSolver of KbnCutCellMethod (include/fub/cutcell_method/KbnStabilisation.hpp:90-100) */
enum { FVS_RIEMANN_HLL_EINFELDT = 0, FVS_RIEMANN_EXACT = 1 };
/* arrays of an EB level readable with fvs_download_eb_array */
enum {
  FVS_EB_ALPHA = 0,          /* volume fractions                (1 comp, cells)  */
  FVS_EB_BETA = 1,           /* face fractions of dir           (1 comp, faces)  */
  FVS_EB_BCENT = 2,          /* boundary centroids              (rank, cells)    */
  FVS_EB_BNORM = 3,          /* boundary normals (AMReX sign)   (rank, cells)    */
  FVS_EB_UNSHIELDED = 4,     /* abs + rel fractions of dir      (2 comps, faces) */
  FVS_EB_SHIELDED_LEFT = 5,
  FVS_EB_SHIELDED_RIGHT = 6,
  FVS_EB_DOUBLY_SHIELDED = 7,
  FVS_EB_REFERENCE_STATES = 8, /* rank + 4 comps, cells */
  FVS_EB_BOUNDARY_FLUXES = 9,  /* rank + 2 comps, cells */
  FVS_EB_STABILIZED_FLUXES = 10, /* rank + 2 comps, flux box of dir */
  FVS_EB_SHIELDED_LEFT_FLUXES = 11,
  FVS_EB_SHIELDED_RIGHT_FLUXES = 12
};

/* Arithmetic of the fused sweep (fvs_set_arithmetic).
 * FVS_ARITH_IEEE (default): every fp64 operation correctly rounded, no FMA
 *   contraction, operation order of the reference's array code path -- results
 *   are bit-identical to the CPU restatement of the reference.
 * FVS_ARITH_TOLERANCE: quotients as a * (1 / b) with a reciprocal refined to
 *   below one ulp, uncorrected square-root refinement, FMA contraction (what
 *   `g++ -O3 -march=native` does to the reference too): each operation within
 *   ~2 ulp of IEEE, conserved variables within 1e-12 relative L-inf per step of
 *   the reference (BASELINE.json north_star), the dt sequence equal to that
 *   relative accuracy rather than bit for bit. */
enum { FVS_ARITH_IEEE = 0, FVS_ARITH_TOLERANCE = 1 };

/* status codes */
enum {
  FVS_OK = 0,
  FVS_ERROR = -1,           /* see fvs_last_error() */
  FVS_TIME_STEP_TOO_LARGE = 1 /* fub::TimeStepTooLarge, include/fub/TimeStepError.hpp:71-78 */
};

typedef struct fvs_ctx fvs_ctx;

typedef struct fvs_config {
  int rank;           /* space dimension 1, 2 or 3 (PerfectGas<rank>) */
  double gamma;       /* PerfectGas::gamma, include/fub/equations/PerfectGas.hpp:216-219 */
  int base_method;    /* FVS_BASE_* */
  int reconstruction; /* FVS_RECON_* */
  int limiter;        /* FVS_LIMITER_* (ignored for first order); FVS_LIMITER2_* for FVS_RECON_MH2_* */
  int device;         /* CUDA device ordinal, -1 = current device */
  int comm_rank;      /* this process' index among the ranks sharing the level (0 if single) */
  int comm_size;      /* number of ranks (1 if single) */
} fvs_config;

const char* fvs_last_error(void);
int fvs_version(void);

/* Number of CUDA devices visible (0 if none); never fails. */
int fvs_device_count(void);

/* replaces: construction of fub::amrex::IntegratorContext + HyperbolicMethod
 * (src/AMReX/IntegratorContext.cpp:60-120, include/fub/HyperbolicMethod.hpp:155-159) */
int fvs_create(const fvs_config* config, fvs_ctx** out);
void fvs_destroy(fvs_ctx* ctx);

/* Change the numerical method of an existing context (HyperbolicMethod is a
 * value that can be swapped, include/fub/HyperbolicMethod.hpp:155-159). */
int fvs_set_method(fvs_ctx* ctx, int base_method, int reconstruction, int limiter);
/* Select the arithmetic of the fused sweep (fvs_sweep, fvs_advance_hierarchy on
 * levels without flux registers / cut cells) and of the all-direction dt pass of
 * regular levels (fvs_compute_stable_dt_all); all other entry points always
 * compute in FVS_ARITH_IEEE.  No counterpart in the reference (its arithmetic
 * is fixed by the compiler flags of the build, CMakeLists.txt). */
int fvs_set_arithmetic(fvs_ctx* ctx, int arithmetic);
/* FluxMethod::GetStencilWidth(), README.md:882-892 */
int fvs_stencil_width(const fvs_ctx* ctx);

/* replaces: IntegratorContext::ResetHierarchyConfiguration (allocation of the
 * scratch and flux MultiFabs, src/AMReX/IntegratorContext.cpp:292-345).
 * `owner[b]` is the comm rank that holds box b (NULL: all on rank 0); only
 * owned boxes get device storage.  Requires scratch_gcw >= flux_gcw + stencil
 * width of the widest method that will be used (2). */
int fvs_level_define(fvs_ctx* ctx, int level, const int domain_lo[3],
                     const int domain_hi[3], const double dx[3],
                     const int periodic[3], int nboxes, const int (*box_lo)[3],
                     const int (*box_hi)[3], const int* owner, int scratch_gcw,
                     int flux_gcw);

/* replaces: BoundarySet of the GriddingAlgorithm
 * (include/fub/AMReX/boundary_condition/BoundarySet.hpp, src/.../BoundarySet.cpp:26-41).
 * Conditions are applied in the order they were added. */
int fvs_level_add_boundary(fvs_ctx* ctx, int level, int dir, int side, int kind);

/* Geometry of the scratch array of one box (cells incl. ghosts). */
int fvs_state_box(const fvs_ctx* ctx, int level, int box, int lo[3], int hi[3],
                  int* ncomp);
/* Geometry of the flux array of one box in direction dir (face centred in dir). */
int fvs_flux_box(const fvs_ctx* ctx, int level, int box, int dir, int lo[3],
                 int hi[3], int* ncomp);

/* replaces: host<->MultiFab views, IntegratorContext::GetScratch(level)
 * (src/AMReX/IntegratorContext.cpp:229-231) + MakeView (ViewFArrayBox.hpp:182-228).
 * host_fab holds (rank+4) components over fvs_state_box in FAB layout. */
int fvs_upload_state(fvs_ctx* ctx, int level, int box, const double* host_fab);
int fvs_download_state(fvs_ctx* ctx, int level, int box, double* host_fab);
/* A sub-box [lo, hi] (global cell indices, inside the state array of the box incl.
 * its ghost cells) of the current state into host[ncomp][hi2-lo2+1][hi1-lo1+1][hi0-lo0+1]:
 * what amrex::FArrayBox::copy of a sub-box does on the host side
 * (the reference reads sub-boxes of GetScratch(level) in its output and tagging code,
 * src/AMReX/IntegratorContext.cpp:303-331). */
int fvs_download_state_region(fvs_ctx* ctx, int level, int box, const int lo[3], const int hi[3],
                              double* host);
/* IntegratorContext::CopyDataToScratch / CopyScratchToData
 * (src/AMReX/IntegratorContext.cpp:629-651): the level data of the patch
 * hierarchy carries NO ghost cells; host_fab holds (rank+4) components over
 * the box's valid cells only (FAB layout).  Copying data to scratch leaves the
 * ghost cells to fvs_fill_ghost_layer (FillMultiFabFromLevel = the same data
 * plus the ghost fill).  Stream ordered like the *_async calls: page-locked
 * buffers, valid until fvs_synchronize. */
int fvs_copy_data_to_scratch_async(fvs_ctx* ctx, int level, int box, const double* host_fab);
int fvs_copy_scratch_to_data_async(fvs_ctx* ctx, int level, int box, double* host_fab);
/* Stream-ordered variants for callers that pipeline several contexts (one per
 * batch of boxes): the copy is enqueued on the context's stream and the call
 * returns; host_fab must be page-locked and stay valid until
 * fvs_synchronize(ctx) (or any entry point that returns a host value). */
int fvs_upload_state_async(fvs_ctx* ctx, int level, int box, const double* host_fab);
int fvs_download_state_async(fvs_ctx* ctx, int level, int box, double* host_fab);
int fvs_synchronize(fvs_ctx* ctx);
/* IntegratorContext::GetFluxes(level, dir), src/AMReX/IntegratorContext.cpp:233-236 */
int fvs_upload_fluxes(fvs_ctx* ctx, int level, int box, int dir, const double* host_fab);
int fvs_download_fluxes(fvs_ctx* ctx, int level, int box, int dir, double* host_fab);
/* Device pointer of the current state array of a box (for zero-copy callers). */
int fvs_state_device_ptr(fvs_ctx* ctx, int level, int box, double** dptr);

/* ---- the three policies of fub::HyperbolicMethod, unfused --------------- */
/* FluxMethodAdapter::ComputeStableDt (include/fub/AMReX/FluxMethodAdapter.hpp:69-114)
 * followed by MinAll over the ranks (src/ext/Mpi.cpp:54-59). */
int fvs_compute_stable_dt(fvs_ctx* ctx, int level, int dir, double* dt_out);
/* FluxMethodAdapter::ComputeNumericFluxes (FluxMethodAdapter.hpp:120-150) */
int fvs_compute_numeric_fluxes(fvs_ctx* ctx, int level, double dt, int dir);
/* ForwardIntegrator::UpdateConservatively (src/AMReX/TimeIntegrator.cpp:29-67) */
int fvs_update_conservatively(fvs_ctx* ctx, int level, double dt, int dir);
/* Reconstruction::CompleteFromCons (include/fub/AMReX/Reconstruction.hpp:44-60) */
int fvs_complete_from_cons(fvs_ctx* ctx, int level);

/* ---- fused fast path ---------------------------------------------------- */
/* One directional sweep of DimensionalSplitLevelIntegrator
 * (include/fub/solver/DimensionalSplitLevelIntegrator.hpp:235-256):
 * ComputeStableDt + MinAll + ComputeNumericFluxes + UpdateConservatively +
 * CompleteFromCons in ONE kernel pass.  *level_dt_out receives the stable dt
 * of the state BEFORE the sweep; returns FVS_TIME_STEP_TOO_LARGE (and leaves
 * the state untouched) if it is smaller than dt. */
int fvs_sweep(fvs_ctx* ctx, int level, double dt, int dir, double* level_dt_out);

/* IntegratorContext::FillGhostLayerSingleLevel (src/AMReX/IntegratorContext.cpp:422-449):
 * intra-level ghost copy (FillBoundary, incl. periodic images and boxes owned
 * by other ranks) followed by every physical boundary condition. */
int fvs_fill_ghost_layer(fvs_ctx* ctx, int level);
/* IntegratorContext::ApplyBoundaryCondition(level, dir) (IntegratorContext.cpp:347-359) */
int fvs_apply_boundary_condition(fvs_ctx* ctx, int level, int dir);

/* DimensionalSplitLevelIntegrator::ComputeStableDt(level): min over all
 * directions (DimensionalSplitLevelIntegrator.hpp:204-219), one fused pass. */
int fvs_compute_stable_dt_all(fvs_ctx* ctx, int level, double* dt_out);

/* DimensionalSplitLevelIntegrator::AdvanceLevelNonRecursively
 * (DimensionalSplitLevelIntegrator.hpp:221-276) for one level followed by the
 * single-level part of SubcycleFineFirstSolver::AdvanceHierarchy
 * (SubcycleFineFirstSolver.hpp:177-193: cycles += 1, time += dt, ghost refill
 * every regrid_frequency cycles).  Sweeps run speculatively without host
 * synchronisation; on a CFL violation the level is rolled back to its state at
 * entry, *failure_dt receives the offending level dt and the call returns
 * FVS_TIME_STEP_TOO_LARGE.  sweep_dts (may be NULL) receives the per-sweep
 * stable dts; *n_sweep_dts is its capacity on entry, the count on exit. */
int fvs_advance_hierarchy(fvs_ctx* ctx, double dt, int split_kind,
                          double* failure_dt, double* sweep_dts, int* n_sweep_dts);

/* fub::RunSimulation without output (include/fub/RunSimulation.hpp:83-191):
 * stable dt -> min(final_time - t, cfl * dt) -> advance -> retry on
 * TimeStepTooLarge.  dt_sequence (capacity dt_capacity) receives the accepted
 * step sizes; returns the number of accepted steps (>= 0) or FVS_ERROR. */
long fvs_run(fvs_ctx* ctx, double cfl, double final_time, long max_cycles,
             int split_kind, double* dt_sequence, long dt_capacity, int* retries);

int fvs_set_regrid_frequency(fvs_ctx* ctx, int frequency);
int fvs_set_cycles(fvs_ctx* ctx, long cycles, double time_point);
long fvs_cycles(const fvs_ctx* ctx);
double fvs_time_point(const fvs_ctx* ctx);

/* fub::MinAll (src/ext/Mpi.cpp:54-59): min over the ranks of the context. */
int fvs_min_all(fvs_ctx* ctx, double* inout);

/* ---- multi-GPU plumbing (one process per GPU) --------------------------- */
/* NCCL bootstrap: rank 0 calls fvs_comm_unique_id, the host distributes the
 * 128 bytes (e.g. torch.distributed broadcast), every rank calls
 * fvs_comm_init.  Replaces ParallelContext::CommunicatorAll
 * (src/AMReX/IntegratorContext.cpp:208-210). */
#define FVS_UNIQUE_ID_BYTES 128
int fvs_comm_unique_id(void* out_id);
int fvs_comm_init(fvs_ctx* ctx, const void* id);

/* Host-only description of the halo exchange (no CUDA needed; used by the
 * gloo tests).  A "copy" moves region [lo,hi] of the cell index space, read
 * from box src_box at index - shift, to the ghost cells of box dst_box.
 * Returns the number of copies; fills at most `capacity` records. */
typedef struct fvs_halo_copy {
  int dst_box, src_box;
  int dst_owner, src_owner;
  int lo[3], hi[3];
  int shift[3];
} fvs_halo_copy;
long fvs_halo_plan(int rank, const int domain_lo[3], const int domain_hi[3],
                   const int periodic[3], int nboxes, const int (*box_lo)[3],
                   const int (*box_hi)[3], const int* owner, int scratch_gcw,
                   fvs_halo_copy* out, long capacity);

/* Host-only: the faces that FluxMethodAdapter::ComputeStableDt evaluates in
 * direction dir for one box (tile size 1024000 x 8 x 8, growntilebox, stencil
 * 2*width; FluxMethodAdapter.hpp:82-98 + FluxMethod.hpp:347-353).  mask has
 * one byte per cell index of the grown box along dir: mask[i] != 0 iff the
 * face between cells i and i+1 takes part. Returns the mask length. */
int fvs_stable_dt_face_mask(int rank, const int box_lo[3], const int box_hi[3],
                            int scratch_gcw, int stencil_width, int dir,
                            unsigned char* mask, int capacity);

/* Host-only: split boxes of a level over `nranks` ranks (space-filling
 * order, equal cell counts); replaces amrex::DistributionMapping for the
 * uniform case. */
int fvs_distribute_boxes(int rank, int nboxes, const int (*box_lo)[3],
                         const int (*box_hi)[3], int nranks, int* owner_out);

/* Host-only: load balance of a cut-cell level,
 * cutcell::GriddingAlgorithm::LoadBalance
 * (src/AMReX/cutcell/GriddingAlgorithm.cpp:190-211): cells of regular fabs
 * weigh 1, cells of single-valued fabs `cutcell_weight`
 * (PatchHierarchyOptions::cutcell_load_balance_weight, 6.0 upstream), covered
 * cells 0.001; knapsack over the box weights (heaviest box first onto the
 * lightest rank).  fab_type[i]: FVS_FAB_*; covered_cells[i]: cells of box i
 * with alpha == 0 (may be NULL).  Covered boxes are dropped by the caller
 * beforehand (PostProcessBaseGrids, :310-330: keep boxes whose fab_type is not
 * FVS_FAB_COVERED). */
int fvs_distribute_boxes_cutcell(int rank, int nboxes, const int (*box_lo)[3],
                                 const int (*box_hi)[3], const int* fab_type,
                                 const long long* covered_cells, double cutcell_weight,
                                 int nranks, int* owner_out);

/* ---- regrid hand-off -------------------------------------------------------- */
/* Host-only: clusters the tags of level l into boxes of level l + 1 -- the
 * hand-off between the device tagging kernels below and fvs_regrid_level.
 * Upstream this is amrex::AmrCore::regrid behind
 * GriddingAlgorithm::RegridAllFinerlevels (src/AMReX/GriddingAlgorithm.cpp:86-110),
 * Berger-Rigoutsos clustering from an unpinned dependency; this stand-in keeps
 * its contract (disjoint boxes covering every refinable tagged cell, aligned to
 * `blocking_factor` fine cells, at most `max_grid_size` fine cells long,
 * properly nested with `nesting_buffer` coarse cells inside the boxes of level
 * l) with a deterministic greedy merge of flagged blocks.
 * tags: one byte per cell of the level-l DOMAIN (x fastest), != 0 = tagged.
 * box_lo/hi: the boxes of level l.  Returns the number of boxes (the first
 * `capacity` are written, in the index space of level l + 1), < 0 on error. */
long fvs_cluster_tags(int rank, const int domain_lo[3], const int domain_hi[3],
                      const int periodic[3], int nboxes, const int (*box_lo)[3],
                      const int (*box_hi)[3], const char* tags, const int ratio[3],
                      int blocking_factor, int max_grid_size, int nesting_buffer,
                      int (*out_lo)[3], int (*out_hi)[3], long capacity);
/* GriddingAlgorithm::RemakeLevel / MakeNewLevelFromCoarse + CopyDataToScratch
 * (src/AMReX/GriddingAlgorithm.cpp:217-260, src/AMReX/IntegratorContext.cpp:
 * 601-638): gives level >= 1 new boxes.  Every cell of the new scratch arrays
 * inside the domain takes the old level's valid data where an old box holds it
 * and the coarse parent (pc_interp) elsewhere; then the level's ghost fill.
 * Domain, cell size, ghost widths and boundary conditions are kept (taken from
 * the coarser level if the level did not exist).  Collective over the ranks of
 * the context.  EB geometry of the new boxes must be set again
 * (fvs_level_set_eb).  nboxes == 0 removes the level and every finer one
 * (GriddingAlgorithm::ClearLevel). */
int fvs_regrid_level(fvs_ctx* ctx, int level, int nboxes, const int (*box_lo)[3],
                     const int (*box_hi)[3], const int* owner);

/* ---- embedded boundary: Klein-2009 KbnCutCellMethod ------------------------ */
/* Hands the EB geometry of one box to the library and turns the level into a
 * cut-cell level: replaces cutcell::PatchHierarchy::GetCutCellData
 * (src/AMReX/cutcell/PatchHierarchy.cpp:507-547, the arrays come from
 * amrex::EBFArrayBoxFactory: getVolFrac, getAreaFrac, getBndryCent,
 * getBndryNormal) and runs FillCutCellData
 * (src/AMReX/cutcell/FillCutCellData.cpp:28-118) plus IsCutCell
 * (src/CutCellData.cpp:225-256) on the device.  alpha / bcent / bnorm live on
 * fvs_state_box (1 / rank / rank components), beta[d] on that box with one more
 * face along d; FAB layout.  Boxes that were not given keep fab_type regular.
 * Requires rank >= 2 and scratch_gcw == flux_gcw + stencil width (the cut-cell
 * context sizes its flux arrays that way, cutcell/IntegratorContext.cpp:523-525).
 * From then on fvs_compute_stable_dt / fvs_compute_numeric_fluxes /
 * fvs_update_conservatively / fvs_complete_from_cons / fvs_sweep /
 * fvs_advance_hierarchy follow the cut-cell adapters
 * (include/fub/AMReX/cutcell/FluxMethod.hpp:153-289,
 * src/AMReX/cutcell/TimeIntegrator.cpp:317-411,
 * include/fub/AMReX/cutcell/Reconstruction.hpp:169-193). */
int fvs_level_set_eb(fvs_ctx* ctx, int level, int box, const double* alpha,
                     const double* const beta[3], const double* bcent, const double* bnorm,
                     int fab_type);
int fvs_set_cutcell_riemann_solver(fvs_ctx* ctx, int kind);
/* cutcell::FluxMethod::PreAdvanceHierarchy (include/fub/AMReX/cutcell/FluxMethod.hpp:108-133):
 * reference states of all cut cells of all levels.  fvs_advance_hierarchy calls
 * it itself (cutcell::IntegratorContext::PreAdvanceLevel, IntegratorContext.cpp:802-806). */
int fvs_pre_advance_hierarchy(fvs_ctx* ctx);
/* Box / component count of an EB array (FVS_EB_*), and its contents. */
int fvs_eb_array_box(const fvs_ctx* ctx, int level, int box, int kind, int dir, int lo[3],
                     int hi[3], int* ncomp);
int fvs_download_eb_array(fvs_ctx* ctx, int level, int box, int kind, int dir, double* host_fab);
/* fub::IsCutCell for every cell of fvs_state_box (1 byte per cell); returns
 * the number of cut cells (the length of the compacted list) or FVS_ERROR. */
long fvs_cut_cell_flags(fvs_ctx* ctx, int level, int box, unsigned char* flags);

/* ---- coarse-fine operations between two levels ----------------------------- */
/* Levels are defined with fvs_level_define(level = 0, 1, ...); index spaces
 * are related by the refinement ratio (default 2 per space dimension,
 * PatchHierarchyOptions::refine_ratio).  Both levels must be owned by this
 * rank (cross-rank coarse-fine exchange is listed as next in DESIGN.md).
 * The flux-register entry points read the per-direction flux arrays that
 * fvs_compute_numeric_fluxes fills (IntegratorContext::GetFluxes). */
int fvs_set_refine_ratio(fvs_ctx* ctx, const int ratio[3]);
/* IntegratorContext::ResetCoarseFineFluxes (src/AMReX/IntegratorContext.cpp:499-512):
 * (re)builds the flux register of `fine` (one plane of coarse faces per side
 * of every fine box, internal borders cleared) and zeroes it. */
int fvs_reset_coarse_fine_fluxes(fvs_ctx* ctx, int fine, int coarse);
/* IntegratorContext::AccumulateCoarseFineFluxes (:472-497): FineAdd into the
 * register of `level` (if level > 0) and CrseInit(ADD) with the opposite sign
 * into the register of level + 1 (if it exists), scaled by vol * scale / dx. */
int fvs_accumulate_coarse_fine_fluxes(fvs_ctx* ctx, int level, double scale, int dir);
/* IntegratorContext::ApplyFluxCorrection (:514-531): Reflux in every direction. */
int fvs_apply_flux_correction(fvs_ctx* ctx, int fine, int coarse, double dt);
/* IntegratorContext::CoarsenConservatively (:451-470): amrex::average_down of
 * the conservative components. */
int fvs_coarsen_conservatively(fvs_ctx* ctx, int fine, int coarse);
/* IntegratorContext::FillGhostLayerTwoLevels (:380-420): FillPatchTwoLevels
 * with the piecewise-constant interpolater, then the fine level's own ghost
 * exchange and physical boundary conditions. */
int fvs_fill_ghost_layer_two_levels(fvs_ctx* ctx, int fine, int coarse);

/* SubcycleFineFirstSolver::ComputeStableDt (include/fub/solver/SubcycleFineFirstSolver.hpp:100-112):
 * min over the levels of (total refinement ratio) * level dt.  With more than
 * one level defined, fvs_advance_hierarchy runs AdvanceLevel (:114-175) fine
 * level first with `ratio` subcycles, flux registers, flux correction and
 * conservative coarsening, and refills all ghost layers afterwards (:177-192);
 * a CFL violation on any level rolls every level back. */
int fvs_compute_stable_dt_hierarchy(fvs_ctx* ctx, double* dt_out);

/* ---- tagging (patch level) ------------------------------------------------ */
/* ScalarGradientDetector_::TagCellsForRefinement
 * (include/fub/tagging_method/GradientDetector.hpp:156-232) on the valid cells
 * of a box, reading the device-resident state and its first ghost layer;
 * condition c tags where the relative jump of state component comps[c]
 * exceeds tols[c] (at most 8 conditions); use_alpha != 0 selects the EB
 * variant (all three volume fractions > 0).  tags (1 byte per valid cell,
 * x fastest) is OR-ed into. */
int fvs_tag_gradient(fvs_ctx* ctx, int level, int box, int ncond, const int* comps,
                     const double* tols, int use_alpha, char* tags);
/* The tags of a whole level for the regrid hand-off (what
 * GriddingAlgorithm::ErrorEst, src/AMReX/GriddingAlgorithm.cpp:177-183, collects
 * through AnyTaggingMethod): GradientDetector conditions (ncond may be 0) and /
 * or TagCutCells on every box this rank owns, then TagBuffer(buffer_width) on
 * the level's map.  tags: one byte per cell of the level's DOMAIN (x fastest),
 * OR-ed into.  One device synchronisation for the whole level. */
int fvs_tag_level(fvs_ctx* ctx, int level, int ncond, const int* comps, const double* tols,
                  int use_alpha, int tag_cutcells, int buffer_width, char* tags);
/* fub::TagCutCells (src/tagging_method/TagCutCells.cpp:27-35) */
int fvs_tag_cutcells(fvs_ctx* ctx, int level, int box, char* tags);
/* fub::TagBuffer (src/tagging_method/TagBuffer.cpp:27-86) on a dense view of
 * extents n[] (x fastest), in place. */
int fvs_tag_buffer(fvs_ctx* ctx, int rank, char* tags, const int n[3], int width);

/* ---- output -------------------------------------------------------------- */
/* fub::amrex::WriteCheckpointFile (src/AMReX/PatchHierarchy.cpp:320-371) for
 * the device-resident hierarchy: <directory>/Header ("Checkpoint File", finest
 * level, cycles and time points per level, the BoxArray of every level) and
 * per level the VisMF pair Level_<l>/data_H + Level_<l>/data_D_<rank> holding
 * the level data (valid cells, all rank+4 components, FAB order, native
 * doubles).  Collective over the ranks of the context; every rank writes the
 * FABs of its own boxes, rank 0 the headers.  The on-disk layout restates
 * AMReX's published VisMF format (version 1, "NFiles"); AMReX is absent here,
 * so it is checked by a reader written against the same description
 * (tests/test_gpu_checkpoint.py), not by AMReX itself. */
int fvs_write_checkpoint(fvs_ctx* ctx, const char* directory);

/* ---- statistics ---------------------------------------------------------- */
/* Number of kernels this context has launched so far. */
long fvs_kernel_launches(const fvs_ctx* ctx);
/* Lazy stable dt.  The fused sweep and fvs_compute_stable_dt_all evaluate the
 * signal speeds of FluxMethod<>::ComputeStableDt (include/fub/flux_method/
 * FluxMethod.hpp:341-382) exactly only on faces whose cheap upper bound comes
 * near the maximum; the bound is compared with a hint taken from the previous
 * evaluation of that direction and the result is verified against the hint
 * afterwards (a violated hint repeats the evaluation without one), so the
 * returned dt is bit-identical either way.  on = 0 disables the hints (every
 * evaluation starts from zero).  Uploading a state drops the hints of its level;
 * on = 2 keeps them (the tests use it to plant a stale hint and exercise the
 * repeat).  No reference counterpart.  Default: 1. */
int fvs_set_dt_hints(fvs_ctx* ctx, int on);
/* Number of evaluations that had to be repeated because a hint was violated. */
long fvs_dt_hint_redos(const fvs_ctx* ctx);
/* The CUDA stream (cudaStream_t) all work of the context is ordered on. */
void* fvs_stream(fvs_ctx* ctx);
/* Elapsed device time (ms) of the most recent fvs_sweep / sweep inside
 * fvs_advance_hierarchy, measured with CUDA events on the context's stream
 * when timing was enabled with fvs_enable_timing(ctx, 1). */
int fvs_enable_timing(fvs_ctx* ctx, int on);
int fvs_sweep_time_ms(fvs_ctx* ctx, double* total_ms, long* n_sweeps);

#ifdef __cplusplus
}
#endif
#endif /* FVS_B200_H */
