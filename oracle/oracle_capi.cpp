// TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.
//
// oracle/oracle_capi.cpp: C entry points of the CPU oracle (liboracle.so),
// loaded with ctypes by tests/, __graft_entry__.smoke() and the cpu_baseline
// leg of bench.py -- and by nothing else.
//
// Parity pinning: the reference cannot be compiled in this environment
// (Eigen, Boost, Vc, AMReX, MPI are absent), so the oracle is
// pinned against the reference's own unit tests, restated in
// tests/test_oracle_golden.py:
//   tests/test.HyperbolicSplitPatchIntegrator.cpp:37-102 (update known answers)
//   tests/test.simd_FluxMethod.cpp:54-186 (scalar == array, bitwise)
//   tests/test.PerfectGas.ExactRiemannProblem.cpp:55-118 (Toro p*, u*)
//   tests/test.State.cpp:40-66 (SoA component layout)
// Third-party arithmetic (Eigen reduction order, AMReX tiling / FillBoundary /
// boxDiff) is restated from the published algorithms and is "parity
// unpinned" against a real AMReX build.
#ifdef _OPENMP
#include <omp.h>
#endif
#include "scalar_hllem.hpp"
#include "level.hpp"
#include "amr.hpp"

#include <cstdio>
#include <cstring>
#include <string>

using namespace fvso;

namespace {
thread_local std::string g_error;
template <typename F> int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1;
  }
}
Box make_box(const int* lo, const int* hi) {
  Box b;
  for (int d = 0; d < 3; ++d) {
    b.lo[d] = lo[d];
    b.hi[d] = hi[d];
  }
  return b;
}
} // namespace

extern "C" {

const char* fvso_last_error() { return g_error.c_str(); }

// Threads of the OpenMP tile loops (the reference: OMP_NUM_THREADS).  bench.py sets
// it explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers.
int fvso_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

// ------------------------------------------------------------------ level
void* fvso_level_create(int rank, const int* dom_lo, const int* dom_hi,
                        const double* dx, const int* periodic, int nboxes,
                        const int* box_lo, const int* box_hi, int scratch_gcw,
                        int flux_gcw, double gamma) {
  Level* L = new Level;
  L->rank = rank;
  L->domain = make_box(dom_lo, dom_hi);
  for (int d = 0; d < 3; ++d) {
    L->dx[d] = dx[d];
    L->periodic[d] = periodic[d] != 0;
  }
  for (int b = 0; b < nboxes; ++b) {
    L->boxes.push_back(make_box(box_lo + 3 * b, box_hi + 3 * b));
  }
  L->scratch_gcw = scratch_gcw;
  L->flux_gcw = flux_gcw;
  L->eq = PerfectGasEq(gamma);
  L->allocate();
  return L;
}
void fvso_level_destroy(void* h) { delete static_cast<Level*>(h); }

void fvso_level_set_method(void* h, int base, int recon, int limiter) {
  Level* L = static_cast<Level*>(h);
  L->method.base = base;
  L->method.recon = recon;
  L->method.limiter = limiter;
}
void fvso_level_set_regrid_frequency(void* h, int f) {
  static_cast<Level*>(h)->regrid_frequency = f;
}
void fvso_level_add_bc(void* h, int dir, int side, int kind) {
  static_cast<Level*>(h)->bcs.push_back({dir, side, kind});
}
void fvso_level_set_cycles(void* h, long cycles, double time_point) {
  Level* L = static_cast<Level*>(h);
  L->cycles = cycles;
  L->time_point = time_point;
}
long fvso_level_cycles(void* h) { return static_cast<Level*>(h)->cycles; }
double fvso_level_time(void* h) { return static_cast<Level*>(h)->time_point; }

double* fvso_level_scratch(void* h, int box, int* lo, int* hi, int* ncomp) {
  Fab& f = static_cast<Level*>(h)->scratch[size_t(box)];
  for (int d = 0; d < 3; ++d) {
    lo[d] = f.box.lo[d];
    hi[d] = f.box.hi[d];
  }
  *ncomp = f.ncomp;
  return f.data.data();
}
double* fvso_level_fluxes(void* h, int box, int dir, int* lo, int* hi,
                          int* ncomp) {
  Fab& f = static_cast<Level*>(h)->fluxes[dir][size_t(box)];
  for (int d = 0; d < 3; ++d) {
    lo[d] = f.box.lo[d];
    hi[d] = f.box.hi[d];
  }
  *ncomp = f.ncomp;
  return f.data.data();
}

int fvso_fill_ghost_layer(void* h) {
  return guarded([&] { static_cast<Level*>(h)->FillGhostLayerSingleLevel(); });
}
int fvso_fill_boundary(void* h) {
  return guarded([&] { static_cast<Level*>(h)->FillBoundary(); });
}
int fvso_apply_boundary_condition(void* h, int dir) {
  return guarded([&] { static_cast<Level*>(h)->ApplyBoundaryCondition(dir); });
}
int fvso_compute_stable_dt(void* h, int dir, double* out) {
  return guarded([&] { *out = static_cast<Level*>(h)->ComputeStableDt(dir); });
}
int fvso_compute_numeric_fluxes(void* h, double dt, int dir) {
  return guarded(
      [&] { static_cast<Level*>(h)->ComputeNumericFluxes_(dt, dir); });
}
int fvso_update_conservatively(void* h, double dt, int dir) {
  return guarded(
      [&] { static_cast<Level*>(h)->UpdateConservatively_(dt, dir); });
}
int fvso_complete_from_cons(void* h) {
  return guarded([&] { static_cast<Level*>(h)->CompleteFromCons_(); });
}
// One directional sweep without the dt check (flux + update + complete).
int fvso_sweep(void* h, double dt, int dir) {
  return guarded([&] {
    Level* L = static_cast<Level*>(h);
    L->ComputeNumericFluxes_(dt, dir);
    L->UpdateConservatively_(dt, dir);
    L->CompleteFromCons_();
  });
}
// returns 0 ok, 1 TimeStepTooLarge (failure dt in *too_large_dt), -1 error.
// sweep_dts (may be null) receives the per-sweep level dt values.
int fvso_advance_hierarchy(void* h, double dt, int split_kind,
                           double* too_large_dt, double* sweep_dts,
                           int* n_sweep_dts) {
  int status = 0;
  int rc = guarded([&] {
    Level* L = static_cast<Level*>(h);
    std::vector<double> sd;
    auto r = L->AdvanceHierarchy(dt, split_kind, &sd);
    if (r) {
      status = 1;
      *too_large_dt = r->dt;
    }
    if (sweep_dts && n_sweep_dts) {
      const int cap = *n_sweep_dts;
      *n_sweep_dts = int(sd.size());
      for (int i = 0; i < cap && i < int(sd.size()); ++i) sweep_dts[i] = sd[size_t(i)];
    }
  });
  return rc < 0 ? rc : status;
}
// Runs the RunSimulation loop; writes the accepted dt sequence.  Returns the
// number of accepted steps (>= 0) or -1 on error; *retries counts CFL retries.
int fvso_run(void* h, double cfl, double final_time, long max_cycles,
             int split_kind, double* dt_sequence, int dt_capacity,
             int* retries) {
  int n = 0;
  int rc = guarded([&] {
    Level* L = static_cast<Level*>(h);
    RunOptions opt;
    opt.cfl = cfl;
    opt.final_time = final_time;
    opt.max_cycles = max_cycles;
    opt.split_kind = split_kind;
    std::vector<double> seq;
    const int r = RunSimulation(*L, opt, &seq);
    if (retries) *retries = r;
    n = int(seq.size());
    for (int i = 0; i < n && i < dt_capacity; ++i) dt_sequence[i] = seq[size_t(i)];
  });
  return rc < 0 ? rc : n;
}

// ------------------------------------------------------------ cut cells
int fvso_level_enable_eb(void* h, int kbn_riemann) {
  return guarded([&] {
    Level* L = static_cast<Level*>(h);
    L->kbn_riemann = kbn_riemann;
    L->allocate_eb();
  });
}
// kind: 0 alpha, 1 beta, 2 bcent, 3 bnorm, 4 us, 5 sl, 6 sr, 7 ds,
//       8 reference states, 9 boundary fluxes, 10 stabilised, 11 shielded
//       left, 12 shielded right fluxes
double* fvso_level_eb_array(void* h, int box, int kind, int dir, int* lo, int* hi, int* ncomp) {
  Level* L = static_cast<Level*>(h);
  if (!L->has_eb) return nullptr;
  const size_t b = size_t(box);
  Fab* f = nullptr;
  switch (kind) {
  case 0: f = &L->eb[b].alpha; break;
  case 1: f = &L->eb[b].beta[dir]; break;
  case 2: f = &L->eb[b].bcent; break;
  case 3: f = &L->eb[b].bnorm; break;
  case 4: f = &L->eb[b].us[dir]; break;
  case 5: f = &L->eb[b].sl[dir]; break;
  case 6: f = &L->eb[b].sr[dir]; break;
  case 7: f = &L->eb[b].ds[dir]; break;
  case 8: f = &L->references[b]; break;
  case 9: f = &L->boundary_fluxes[b]; break;
  case 10: f = &L->stabilized[dir][b]; break;
  case 11: f = &L->shielded_left[dir][b]; break;
  case 12: f = &L->shielded_right[dir][b]; break;
  default: return nullptr;
  }
  for (int d = 0; d < 3; ++d) {
    lo[d] = f->box.lo[d];
    hi[d] = f->box.hi[d];
  }
  *ncomp = f->ncomp;
  return f->data.data();
}
int fvso_level_eb_finalize(void* h, int box, int fab_type) {
  return guarded([&] {
    Level* L = static_cast<Level*>(h);
    L->eb[size_t(box)].fab_type = fab_type;
    FillCutCellData(L->eb[size_t(box)], L->rank);
  });
}
int fvso_level_is_cut_cell(void* h, int box, unsigned char* out) {
  return guarded([&] {
    Level* L = static_cast<Level*>(h);
    const CutCellData& g = L->eb[size_t(box)];
    const Box& bx = g.alpha.box;
    size_t n = 0;
    for (int k = bx.lo[2]; k <= bx.hi[2]; ++k)
      for (int j = bx.lo[1]; j <= bx.hi[1]; ++j)
        for (int i = bx.lo[0]; i <= bx.hi[0]; ++i) out[n++] = IsCutCell(g, L->rank, i, j, k) ? 1 : 0;
  });
}
int fvso_level_pre_advance(void* h) {
  return guarded([&] { static_cast<Level*>(h)->PreAdvanceHierarchy(); });
}
// ------------------------------------------------------------ tagging
int fvso_tag_gradient(void* h, int box, int ncond, const int* comps, const double* tols,
                      int use_alpha, char* tags) {
  return guarded([&] {
    Level* L = static_cast<Level*>(h);
    const size_t b = size_t(box);
    TagGradient(tags, L->boxes[b], L->scratch[b], use_alpha ? &L->eb[b].alpha : nullptr, L->rank,
                ncond, comps, tols);
  });
}
int fvso_tag_cutcells(void* h, int box, char* tags) {
  return guarded([&] {
    Level* L = static_cast<Level*>(h);
    TagCutCells(tags, L->boxes[size_t(box)], L->eb[size_t(box)].alpha);
  });
}
int fvso_tag_buffer(char* tags, int rank, const int* n, int width) {
  return guarded([&] { TagBuffer(tags, rank, n, width); });
}

// ------------------------------------------------------------ two levels
void* fvso_cf_create(void* fine, void* coarse, const int* ratio) {
  CoarseFine* cf = new CoarseFine;
  for (int d = 0; d < 3; ++d) cf->ratio[d] = ratio[d];
  cf->define(*static_cast<Level*>(fine), *static_cast<Level*>(coarse));
  return cf;
}
void fvso_cf_destroy(void* cf) { delete static_cast<CoarseFine*>(cf); }
int fvso_cf_reset(void* cf) {
  return guarded([&] { static_cast<CoarseFine*>(cf)->Reset(); });
}
// level_is_fine != 0: FineAdd with `level` as the fine level, else CrseInit
int fvso_cf_accumulate(void* cf, void* level, int level_is_fine, double scale, int dir) {
  return guarded([&] {
    Level* L = static_cast<Level*>(level);
    CoarseFine* c = static_cast<CoarseFine*>(cf);
    AccumulateCoarseFineFluxes(level_is_fine ? c : nullptr, *L, level_is_fine ? nullptr : c, scale,
                               dir);
  });
}
int fvso_cf_apply_flux_correction(void* cf, void* coarse) {
  return guarded([&] {
    Level* L = static_cast<Level*>(coarse);
    for (int d = 0; d < L->rank; ++d) static_cast<CoarseFine*>(cf)->Reflux(*L, d, 1.0);
  });
}
int fvso_cf_coarsen_conservatively(void* cf, void* fine, void* coarse) {
  return guarded([&] {
    CoarsenConservatively(*static_cast<Level*>(fine), *static_cast<Level*>(coarse),
                          static_cast<CoarseFine*>(cf)->ratio);
  });
}
int fvso_cf_fill_ghost_layer_two_levels(void* cf, void* fine, void* coarse) {
  return guarded([&] {
    FillGhostLayerTwoLevels(*static_cast<Level*>(fine), *static_cast<Level*>(coarse),
                            static_cast<CoarseFine*>(cf)->ratio);
  });
}

int fvso_regrid_fill(void* fresh, void* old, void* coarse, const int* ratio) {
  return guarded([&] {
    RegridFill(*static_cast<Level*>(fresh), static_cast<const Level*>(old),
               *static_cast<Level*>(coarse), ratio);
  });
}

// --------------------------------------------------------------- hierarchy
void* fvso_hierarchy_create(int nlevels, void** levels, const int* ratio) {
  Hierarchy* H = new Hierarchy;
  for (int l = 0; l < nlevels; ++l) H->levels.push_back(static_cast<Level*>(levels[l]));
  for (int d = 0; d < 3; ++d) H->ratio[d] = ratio[d];
  H->define();
  return H;
}
void fvso_hierarchy_destroy(void* h) { delete static_cast<Hierarchy*>(h); }
int fvso_hierarchy_stable_dt(void* h, double* out) {
  return guarded([&] { *out = static_cast<Hierarchy*>(h)->ComputeStableDt(); });
}
int fvso_hierarchy_advance(void* h, double dt, int split_kind, double* too_large_dt,
                           double* sweep_dts, int* n_sweep_dts) {
  int status = 0;
  int rc = guarded([&] {
    std::vector<double> sd;
    auto r = static_cast<Hierarchy*>(h)->AdvanceHierarchy(dt, split_kind, &sd);
    if (r) {
      status = 1;
      *too_large_dt = r->dt;
    }
    if (sweep_dts && n_sweep_dts) {
      const int cap = *n_sweep_dts;
      *n_sweep_dts = int(sd.size());
      for (int i = 0; i < cap && i < int(sd.size()); ++i) sweep_dts[i] = sd[size_t(i)];
    }
  });
  return rc < 0 ? rc : status;
}

// ----------------------------------------------------- unit-level helpers
// UpdateConservatively on a bare array pair: `cells` x-cells, 1 component,
// direction X (tests/test.HyperbolicSplitPatchIntegrator.cpp).
void fvso_unit_update_conservatively(double* next, const double* prev,
                                     const double* flux, int n, double dt,
                                     double dx) {
  Fab nf, pf, ff;
  Box cb;
  cb.hi[0] = n - 1;
  nf.define(cb, 1);
  pf.define(cb, 1);
  ff.define(surrounding_nodes(cb, 0), 1);
  std::memcpy(pf.data.data(), prev, sizeof(double) * size_t(n));
  std::memcpy(ff.data.data(), flux, sizeof(double) * size_t(n + 1));
  UpdateConservatively(nf, pf, ff, cb, 1, dt, dx, 0);
  std::memcpy(next, nf.data.data(), sizeof(double) * size_t(n));
}

// CompleteFromCons of single states: in = [rho, m.., E], out = [.., p, a]
void fvso_unit_complete_from_cons(int rank, double gamma, const double* cons,
                                  double* complete) {
  PerfectGasEq eq(gamma);
  auto run = [&](auto Dc) {
    constexpr int D = decltype(Dc)::value;
    Cons<double, D> c;
    c.rho = cons[0];
    for (int d = 0; d < D; ++d) c.mom[d] = cons[1 + d];
    c.E = cons[1 + D];
    Complete<double, D> q;
    CompleteFromCons<double, D>(eq, q, c);
    complete[0] = q.rho;
    for (int d = 0; d < D; ++d) complete[1 + d] = q.mom[d];
    complete[1 + D] = q.E;
    complete[2 + D] = q.p;
    complete[3 + D] = q.a;
  };
  if (rank == 1) run(std::integral_constant<int, 1>{});
  else if (rank == 2) run(std::integral_constant<int, 2>{});
  else run(std::integral_constant<int, 3>{});
}

} // extern "C"

namespace {
template <int D, typename T>
void unpack(Complete<T, D>* st, const double* in, int nstates, int lane) {
  for (int s = 0; s < nstates; ++s) {
    const double* q = in + s * (D + 4);
    if constexpr (std::is_same_v<T, double>) {
      (void)lane;
      st[s].rho = q[0];
      for (int d = 0; d < D; ++d) st[s].mom[d] = q[1 + d];
      st[s].E = q[1 + D];
      st[s].p = q[2 + D];
      st[s].a = q[3 + D];
    } else {
      st[s].rho = A8(q[0]);
      for (int d = 0; d < D; ++d) st[s].mom[d] = A8(q[1 + d]);
      st[s].E = A8(q[1 + D]);
      st[s].p = A8(q[2 + D]);
      st[s].a = A8(q[3 + D]);
    }
  }
}
} // namespace

extern "C" {

// Single-face flux of any method from 2w Complete states ([rho,m..,E,p,a] per
// state): the array path, lane 0 of a broadcast chunk.
int fvso_unit_numeric_flux(int rank, double gamma, int base, int recon,
                           int limiter, const double* states, double dt,
                           double dx, int dir, double* flux_out) {
  return guarded([&] {
    PerfectGasEq eq(gamma);
    MethodSpec ms{base, recon, limiter};
    auto run = [&](auto Dc) {
      constexpr int D = decltype(Dc)::value;
      with_method<D>(ms, eq, [&](const auto& m) {
        using M = std::decay_t<decltype(m)>;
        Complete<A8, D> st[2 * M::kWidth];
        unpack<D, A8>(st, states, 2 * M::kWidth, 0);
        Cons<A8, D> f;
        m.ComputeNumericFlux(f, st, dt, dx, dir);
        flux_out[0] = f.rho.v[0];
        for (int d = 0; d < D; ++d) flux_out[1 + d] = f.mom[d].v[0];
        flux_out[1 + D] = f.E.v[0];
      });
    };
    if (rank == 1) run(std::integral_constant<int, 1>{});
    else if (rank == 2) run(std::integral_constant<int, 2>{});
    else run(std::integral_constant<int, 3>{});
  });
}

// The reference's scalar HLLEM code path (scalar_hllem.hpp; Rank 1 and 2), restated
// independently of the array path: 2 Complete states -> Conservative flux.
int fvso_unit_hllem_scalar(int rank, double gamma, int larrouturou, const double* states, int dir,
                           double* flux_out) {
  return guarded([&] {
    auto run = [&](auto Dc) {
      constexpr int D = decltype(Dc)::value;
      fvso_scalar::State<D> st[2];
      for (int k = 0; k < 2; ++k) {
        const double* p = states + k * (D + 4);
        st[k].rho = p[0];
        for (int d = 0; d < D; ++d) st[k].mom[d] = p[1 + d];
        st[k].E = p[1 + D];
        st[k].p = p[2 + D];
        st[k].a = p[3 + D];
      }
      fvso_scalar::Flux<D> f;
      fvso_scalar::HllemScalar<D>(f, st[0], st[1], gamma, dir, larrouturou != 0);
      flux_out[0] = f.rho;
      for (int d = 0; d < D; ++d) flux_out[1 + d] = f.mom[d];
      flux_out[1 + D] = f.E;
    };
    if (rank == 1) run(std::integral_constant<int, 1>{});
    else if (rank == 2) run(std::integral_constant<int, 2>{});
    else throw std::runtime_error("the reference's scalar HLLEM path is complete for Rank 1 and 2 only");
  });
}

// Scalar code path of MusclHancock<HllEinfeldt, MinMod|VanLeer|NoLimiter> and
// of first-order HllEinfeldt (tests/test.simd_FluxMethod.cpp compares scalar
// against array bit for bit).
int fvso_unit_numeric_flux_scalar_hll(int rank, double gamma, int recon,
                                      int limiter, const double* states,
                                      double dt, double dx, int dir,
                                      double* flux_out) {
  return guarded([&] {
    PerfectGasEq eq(gamma);
    auto run = [&](auto Dc) {
      constexpr int D = decltype(Dc)::value;
      auto go = [&](const auto& m) {
        using M = std::decay_t<decltype(m)>;
        Complete<double, D> st[2 * M::kWidth];
        unpack<D, double>(st, states, 2 * M::kWidth, 0);
        Cons<double, D> f;
        m.ComputeNumericFlux(f, st, dt, dx, dir);
        flux_out[0] = f.rho;
        for (int d = 0; d < D; ++d) flux_out[1 + d] = f.mom[d];
        flux_out[1 + D] = f.E;
      };
      using Hll = HllEinfeldt<double, D>;
      if (recon == kFirstOrder) go(Hll(eq));
      else if (limiter == kNoLimiter) go(MusclHancock<Hll, NoLimiter>(eq));
      else if (limiter == kMinMod) go(MusclHancock<Hll, MinMod>(eq));
      else go(MusclHancock<Hll, VanLeer>(eq));
    };
    if (rank == 1) run(std::integral_constant<int, 1>{});
    else if (rank == 2) run(std::integral_constant<int, 2>{});
    else run(std::integral_constant<int, 3>{});
  });
}

// Exact Riemann solver: p*, u* (Toro) and the sampled state at x/t = 0.
int fvso_unit_exact_riemann(int rank, double gamma, const double* left,
                            const double* right, int dir, double* pu_star,
                            double* solution) {
  return guarded([&] {
    PerfectGasEq eq(gamma);
    auto run = [&](auto Dc) {
      constexpr int D = decltype(Dc)::value;
      Complete<double, D> st[2];
      double both[2 * (D + 4)];
      std::memcpy(both, left, sizeof(double) * (D + 4));
      std::memcpy(both + D + 4, right, sizeof(double) * (D + 4));
      unpack<D, double>(st, both, 2, 0);
      ExactRiemannSolver<D> rs(eq);
      auto pu = rs.ComputeMiddleState(st[0], st[1], dir);
      pu_star[0] = pu[0];
      pu_star[1] = pu[1];
      Complete<double, D> sol = zero_state<D>();
      rs.SolveRiemannProblem(sol, st[0], st[1], dir);
      solution[0] = sol.rho;
      for (int d = 0; d < D; ++d) solution[1 + d] = sol.mom[d];
      solution[1 + D] = sol.E;
      solution[2 + D] = sol.p;
      solution[3 + D] = sol.a;
    };
    if (rank == 1) run(std::integral_constant<int, 1>{});
    else if (rank == 2) run(std::integral_constant<int, 2>{});
    else run(std::integral_constant<int, 3>{});
  });
}

// AMReX tiling helper exposed for tests: number of tiles and their boxes.
int fvso_unit_tiles(int rank, const int* lo, const int* hi,
                    const int* tile_size, const int* ngrow, int nodal_dir,
                    int* out_boxes /* 12 ints per tile: tile lo,hi,grown lo,hi */,
                    int capacity) {
  std::vector<Tile> ts;
  build_tiles(ts, make_box(lo, hi), rank, tile_size, ngrow, nodal_dir);
  for (int t = 0; t < int(ts.size()) && t < capacity; ++t) {
    for (int d = 0; d < 3; ++d) {
      out_boxes[12 * t + d] = ts[size_t(t)].tile.lo[d];
      out_boxes[12 * t + 3 + d] = ts[size_t(t)].tile.hi[d];
      out_boxes[12 * t + 6 + d] = ts[size_t(t)].grown.lo[d];
      out_boxes[12 * t + 9 + d] = ts[size_t(t)].grown.hi[d];
    }
  }
  return int(ts.size());
}

} // extern "C"
