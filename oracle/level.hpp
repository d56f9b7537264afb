// TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.
//
// oracle/level.hpp: CPU restatement of the reference's L3/L4 orchestration
// for ONE refinement level made of several boxes:
//   IntegratorContext (scratch + per-direction flux arrays, ghost fill, BCs),
//   FluxMethodAdapter / ForwardIntegrator / Reconstruction (tile loops),
//   DimensionalSplitLevelIntegrator + split methods, and the RunSimulation
//   step loop with its CFL-retry.
// Citations are relative to /root/reference.
#pragma once
#include "patch.hpp"
#include "cutcell.hpp"

#include <cfloat>
#include <functional>
#include <optional>

namespace fvso {

enum BaseKind { kHllemLarrouturou = 0, kHllem = 1, kHllEinfeldt = 2, kGodunovExact = 3 };
enum ReconKind {
  kFirstOrder = 0,
  kMusclHancock = 1,
  kMusclHancockPrim = 2,
  kMusclHancockChar = 3,
  // MusclHancock2 (MusclHancockMethod2.hpp); `limiter` is then a Limiter2Kind
  kMh2ConservativeRecon = 4,
  kMh2PrimitiveRecon = 5,
  kMh2CharacteristicsRecon = 6
};
enum LimiterKind { kNoLimiter = 0, kMinMod = 1, kVanLeer = 2 };
enum BcKind { kTransmissive = 0, kReflective = 1 };
enum SplitKind { kGodunovSplitting = 0, kStrangSplitting = 1, kStrangSplittingLumped = 2 };

struct MethodSpec {
  int base{kHllemLarrouturou};
  int recon{kMusclHancock};
  int limiter{kVanLeer};
  int stencil_width() const { return recon == kFirstOrder ? 1 : 2; }
};

struct BoundaryCondition {
  int dir;
  int side;
  int kind;
};

// Calls f(method_object) with the statically typed flux method that
// corresponds to the runtime MethodSpec.
template <int D, typename F>
void with_method(const MethodSpec& ms, const PerfectGasEq& eq, F&& f) {
  auto with_base = [&](auto wrap) {
    switch (ms.base) {
    case kHllemLarrouturou: wrap(Hllem<D, true>(eq)); break;
    case kHllem: wrap(Hllem<D, false>(eq)); break;
    case kHllEinfeldt: wrap(HllEinfeldt<A8, D>(eq)); break;
    case kGodunovExact: wrap(GodunovExact<D>(eq)); break;
    default: throw std::runtime_error("unknown base method");
    }
  };
  if (ms.recon == kMusclHancockPrim) {
    // perfect_gas::MusclHancockPrim is a complete method: own limiter, HLLEM
    // with the Larrouturou modification (MusclHancockPrimMethod.hpp:175)
    f(MusclHancockPrim<D>(eq));
  } else if (ms.recon == kMusclHancockChar) {
    f(MusclHancockChar<D>(eq));
  } else if (ms.recon >= kMh2ConservativeRecon) {
    with_base([&](auto base) {
      using Base = decltype(base);
      switch (ms.recon) {
      case kMh2ConservativeRecon: f(MusclHancock2<Base, kMh2Conservative>(eq, ms.limiter)); break;
      case kMh2PrimitiveRecon: f(MusclHancock2<Base, kMh2Primitive>(eq, ms.limiter)); break;
      case kMh2CharacteristicsRecon:
        f(MusclHancock2<Base, kMh2Characteristics>(eq, ms.limiter));
        break;
      default: throw std::runtime_error("unknown reconstruction");
      }
    });
  } else if (ms.recon == kFirstOrder) {
    with_base([&](auto base) { f(base); });
  } else {
    with_base([&](auto base) {
      using Base = decltype(base);
      switch (ms.limiter) {
      case kNoLimiter: f(MusclHancock<Base, NoLimiter>(eq)); break;
      case kMinMod: f(MusclHancock<Base, MinMod>(eq)); break;
      case kVanLeer: f(MusclHancock<Base, VanLeer>(eq)); break;
      default: throw std::runtime_error("unknown limiter");
      }
    });
  }
}

struct TimeStepTooLarge {
  double dt;
};

struct Level {
  int rank{1};
  Box domain;
  double dx[3]{1, 1, 1};
  bool periodic[3]{false, false, false};
  std::vector<Box> boxes;
  // include/fub/AMReX/IntegratorContext.hpp:52-54
  int scratch_gcw{4};
  int flux_gcw{2};
  int regrid_frequency{1};
  std::vector<Fab> scratch;
  std::vector<Fab> fluxes[3];
  std::vector<BoundaryCondition> bcs;
  PerfectGasEq eq;
  MethodSpec method;
  long cycles{0};
  double time_point{0.0};
  // embedded boundary (src/AMReX/cutcell/IntegratorContext.cpp:470-533)
  bool has_eb{false};
  int kbn_riemann{kRiemannHllEinfeldt};
  std::vector<CutCellData> eb;
  std::vector<Fab> references, boundary_fluxes;
  std::vector<Fab> stabilized[3], shielded_left[3], shielded_right[3];

  int ncomp() const { return rank + 4; }
  int ncons() const { return rank + 2; }

  // src/AMReX/IntegratorContext.cpp:303-324
  void allocate() {
    scratch.resize(boxes.size());
    for (int d = 0; d < 3; ++d) fluxes[d].clear();
    for (int d = 0; d < rank; ++d) fluxes[d].resize(boxes.size());
    for (size_t b = 0; b < boxes.size(); ++b) {
      scratch[b].define(grow_all(boxes[b], rank, scratch_gcw), ncomp());
      for (int d = 0; d < rank; ++d) {
        Box fb = surrounding_nodes(boxes[b], d);
        for (int e = 0; e < rank; ++e) {
          fb = grow(fb, e, e == d ? flux_gcw : scratch_gcw);
        }
        fluxes[d][b].define(fb, ncons());
      }
    }
  }

  // cut-cell arrays of the context; geometry is filled in by the caller
  void allocate_eb() {
    if (scratch_gcw != flux_gcw + method.stencil_width()) {
      // the cut-cell context allocates fluxes with flux_gcw + stencil width
      // ghost faces across the sweep direction (IntegratorContext.cpp:523-525)
      throw std::runtime_error("EB levels need scratch_gcw == flux_gcw + stencil width");
    }
    has_eb = true;
    eb.resize(boxes.size());
    references.resize(boxes.size());
    boundary_fluxes.resize(boxes.size());
    for (int d = 0; d < rank; ++d) {
      stabilized[d].resize(boxes.size());
      shielded_left[d].resize(boxes.size());
      shielded_right[d].resize(boxes.size());
    }
    for (size_t b = 0; b < boxes.size(); ++b) {
      eb[b].define(scratch[b].box, rank);
      references[b].define(scratch[b].box, ncomp());
      boundary_fluxes[b].define(scratch[b].box, ncons());
      for (int d = 0; d < rank; ++d) {
        stabilized[d][b].define(fluxes[d][b].box, ncons());
        shielded_left[d][b].define(fluxes[d][b].box, ncons());
        shielded_right[d][b].define(fluxes[d][b].box, ncons());
      }
    }
  }

  // include/fub/AMReX/cutcell/FluxMethod.hpp:108-133
  void PreAdvanceHierarchy() {
    if (!has_eb) return;
    for (size_t b = 0; b < boxes.size(); ++b) {
      if (eb[b].fab_type != kFabSingleValued) continue;
      switch (rank) {
      case 2: Kbn<2>(eq, kbn_riemann).PreAdvanceHierarchy(references[b], scratch[b], eb[b], scratch[b].box); break;
      case 3: Kbn<3>(eq, kbn_riemann).PreAdvanceHierarchy(references[b], scratch[b], eb[b], scratch[b].box); break;
      default: throw std::runtime_error("cut cells need rank 2 or 3");
      }
    }
  }

  // include/fub/AMReX/cutcell/FluxMethod.hpp:258-289
  template <int D> double ComputeStableDtEbD(int dir) const {
    const int tile_size[3] = {1024000, 16, 32};
    const std::vector<TileRef> tiles = Tiles(tile_size, dir);
    const double dxd = dx[dir];
    const int w = method.stencil_width();
    double min_dt = std::numeric_limits<double>::infinity();
    if constexpr (D >= 2) {
      const Kbn<D> kbn(eq, kbn_riemann);
      with_method<D>(method, eq, [&](const auto& m) {
        for (size_t t = 0; t < tiles.size(); ++t) {
          const size_t b = tiles[t].fab;
          Box cells, faces;
          GetCellsAndFacesInStencilRange(cells, faces, scratch[b].box, tiles[t].tile.grown, w, dir);
          if (eb[b].fab_type == kFabSingleValued) {
            min_dt = std::min(min_dt, kbn.ComputeStableDt(m, scratch[b], eb[b], cells, dxd, dir));
          } else if (eb[b].fab_type == kFabRegular) {
            min_dt = std::min(min_dt, fvso::ComputeStableDt(m, scratch[b], cells, dxd, dir));
          }
        }
      });
    }
    return min_dt;
  }

  // include/fub/AMReX/cutcell/FluxMethod.hpp:153-256
  template <int D> void ComputeNumericFluxesEbD(double dt, int dir) {
    if constexpr (D >= 2) {
      const Kbn<D> kbn(eq, kbn_riemann);
      for (size_t b = 0; b < boxes.size(); ++b) {
        std::fill(boundary_fluxes[b].data.begin(), boundary_fluxes[b].data.end(), 0.0);
        if (eb[b].fab_type == kFabSingleValued) {
          kbn.ComputeBoundaryFluxes(boundary_fluxes[b], scratch[b], references[b], eb[b],
                                    scratch[b].box, dir);
        }
      }
      const int tile_size[3] = {1024000, 16, 32};
      const std::vector<TileRef> tiles = Tiles(tile_size, dir);
      const double dxd = dx[dir];
      const int w = method.stencil_width();
      with_method<D>(method, eq, [&](const auto& m) {
        for (size_t t = 0; t < tiles.size(); ++t) {
          const size_t b = tiles[t].fab;
          Box cells, faces;
          GetCellsAndFacesInStencilRange(cells, faces, scratch[b].box, tiles[t].tile.grown, w, dir);
          if (eb[b].fab_type == kFabSingleValued) {
            kbn.ComputeRegularFluxes(m, fluxes[dir][b], faces, scratch[b], eb[b], dt, dxd, dir);
            kbn.ComputeCutCellFluxes(stabilized[dir][b], shielded_left[dir][b],
                                     shielded_right[dir][b], boundary_fluxes[b], fluxes[dir][b],
                                     faces, eb[b], dir);
          } else if (eb[b].fab_type == kFabRegular) {
            fvso::ComputeNumericFluxes(m, fluxes[dir][b], faces, scratch[b], dt, dxd, dir);
          }
        }
      });
    }
  }

  // src/AMReX/cutcell/TimeIntegrator.cpp:317-411
  void UpdateConservativelyEb(double dt, int dir) {
    const int tile_size[3] = {1024000, 16, 32};
    const std::vector<TileRef> tiles = Tiles(tile_size, -1);
    const double dt_over_dx = dt / dx[dir];
    for (size_t t = 0; t < tiles.size(); ++t) {
      const size_t b = tiles[t].fab;
      const Box all_faces = surrounding_nodes(tiles[t].tile.grown, dir);
      const Box flux_box = intersect(all_faces, fluxes[dir][b].box);
      const Box cell_box = enclosed_cells(flux_box, dir);
      if (eb[b].fab_type == kFabSingleValued) {
        CutCellUpdateConservatively(scratch[b], fluxes[dir][b], stabilized[dir][b],
                                    shielded_left[dir][b], shielded_right[dir][b],
                                    boundary_fluxes[b], eb[b], cell_box, ncons(), dt_over_dx, dir);
      } else if (eb[b].fab_type == kFabRegular) {
        fvso::UpdateConservatively(scratch[b], scratch[b], fluxes[dir][b], cell_box, ncons(), dt,
                                   dx[dir], dir);
      }
    }
  }

  // include/fub/AMReX/cutcell/Reconstruction.hpp:169-193: the flag type is
  // taken per grown tile box
  int CompleteTileType(size_t b, const Box& grown) const {
    return CellFlagType(eb[b], rank, grown);
  }
  void CompleteFromConsEb() {
    const int tile_size[3] = {1024000, 16, 32};
    const std::vector<TileRef> tiles = Tiles(tile_size, -1);
    for (size_t t = 0; t < tiles.size(); ++t) {
      const size_t b = tiles[t].fab;
      const Box& box = tiles[t].tile.grown;
      const int type = CompleteTileType(b, box);
      if (type == kFabRegular) {
        switch (rank) {
        case 2: CompleteFromConsBox<2>(eq, scratch[b], box); break;
        default: CompleteFromConsBox<3>(eq, scratch[b], box); break;
        }
      } else if (type == kFabSingleValued) {
        switch (rank) {
        case 2: CompleteFromConsBoxMasked<2>(eq, scratch[b], eb[b].alpha, box); break;
        default: CompleteFromConsBoxMasked<3>(eq, scratch[b], eb[b].alpha, box); break;
        }
      }
    }
  }

  // -------------------------------------------------------------------------
  // AMReX: FabArray::FillBoundary(periodicity) -- every ghost cell that lies
  // inside the valid box of another fab (or of a periodic image) is copied.
  void FillBoundary() {
    std::vector<std::array<int, 3>> shifts;
    int r[3];
    for (int d = 0; d < 3; ++d) r[d] = (d < rank && periodic[d]) ? 1 : 0;
    for (int sz = -r[2]; sz <= r[2]; ++sz)
      for (int sy = -r[1]; sy <= r[1]; ++sy)
        for (int sx = -r[0]; sx <= r[0]; ++sx)
          shifts.push_back({sx * int(domain.len(0)), sy * int(domain.len(1)),
                            sz * int(domain.len(2))});
    // Copy from a snapshot-free source is fine: sources are valid cells,
    // destinations are ghost cells, the two sets are disjoint.
    for (size_t i = 0; i < boxes.size(); ++i) {
      Fab& dst = scratch[i];
      for (size_t j = 0; j < boxes.size(); ++j) {
        const Fab& src = scratch[j];
        for (const auto& s : shifts) {
          if (i == j && s[0] == 0 && s[1] == 0 && s[2] == 0) continue;
          Box image = boxes[j];
          for (int d = 0; d < 3; ++d) image = shift(image, d, s[d]);
          const Box region = intersect(dst.box, image);
          if (!region.ok()) continue;
          for (int c = 0; c < ncomp(); ++c)
            for (int k = region.lo[2]; k <= region.hi[2]; ++k)
              for (int j2 = region.lo[1]; j2 <= region.hi[1]; ++j2)
                for (int i2 = region.lo[0]; i2 <= region.hi[0]; ++i2)
                  dst(i2, j2, k, c) = src(i2 - s[0], j2 - s[1], k - s[2], c);
        }
      }
    }
  }

  // AMReX: Geometry::growNonPeriodicDomain
  Box GrowNonPeriodicDomain(int ngrow) const {
    Box b = domain;
    for (int d = 0; d < rank; ++d) {
      if (!periodic[d]) b = grow(b, d, ngrow);
    }
    return b;
  }

  // src/AMReX/boundary_condition/TransmissiveBoundary.cpp:60-93
  void FillTransmissive(const BoundaryCondition& bc) {
    const int idir = bc.dir;
    if (periodic[idir]) return;
    const int ngrow = scratch_gcw;
    Box slab = GrowNonPeriodicDomain(ngrow);
    if (bc.side == 0) {
      slab.hi[idir] = domain.lo[idir] - 1;
    } else {
      slab.lo[idir] = domain.hi[idir] + 1;
    }
    if (!slab.ok()) return;
    const int boundary = bc.side == 0 ? domain.lo[idir] : domain.hi[idir];
    for (Fab& fab : scratch) {
      const Box fill = intersect(fab.box, slab);
      if (!fill.ok()) continue;
      for (int c = 0; c < ncomp(); ++c)
        for (int k = fill.lo[2]; k <= fill.hi[2]; ++k)
          for (int j = fill.lo[1]; j <= fill.hi[1]; ++j)
            for (int i = fill.lo[0]; i <= fill.hi[0]; ++i) {
              int src[3] = {i, j, k};
              src[idir] = boundary; // MapToSrc :31-43
              fab(i, j, k, c) = fab(src[0], src[1], src[2], c);
            }
    }
  }

  // include/fub/AMReX/boundary_condition/ReflectiveBoundary.hpp:67-94,
  // include/fub/boundary_condition/ReflectiveBoundary.hpp:57-70,
  // src/boundary_condition/ReflectiveBoundary.cpp:24-36,
  // src/equations/PerfectGas.cpp:266-300.
  // AMReX: complementIn(grown, domain) = boxDiff, which peels slabs from the
  // highest dimension down: the slab of dimension d spans the full grown
  // extent in dimensions < d and only the domain extent in dimensions > d.
  void FillReflective(const BoundaryCondition& bc) {
    const int d = bc.dir;
    if (periodic[d]) return; // no ghost slab outside a periodic direction
    const int ngrow = scratch_gcw;
    Box slab = GrowNonPeriodicDomain(ngrow);
    for (int e = d + 1; e < rank; ++e) {
      slab.lo[e] = domain.lo[e];
      slab.hi[e] = domain.hi[e];
    }
    if (bc.side == 0) {
      slab.hi[d] = domain.lo[d] - 1;
    } else {
      slab.lo[d] = domain.hi[d] + 1;
    }
    if (!slab.ok()) return;
    for (Fab& fab : scratch) {
      const Box fill = intersect(fab.box, slab);
      if (!fill.ok()) continue;
      for (int k = fill.lo[2]; k <= fill.hi[2]; ++k)
        for (int j = fill.lo[1]; j <= fill.hi[1]; ++j)
          for (int i = fill.lo[0]; i <= fill.hi[0]; ++i) {
            int idx[3] = {i, j, k};
            int src[3] = {i, j, k};
            if (bc.side == 0) {
              const int distance = (fill.hi[d] + 1) - idx[d];
              src[d] = idx[d] + 2 * distance - 1;
            } else {
              const int distance = idx[d] - (fill.lo[d] - 1);
              src[d] = idx[d] - 2 * distance + 1;
            }
            for (int c = 0; c < ncomp(); ++c) {
              double v = fab(src[0], src[1], src[2], c);
              // m - 2 (m . e_d) e_d  ==  -m_d exactly
              if (c == 1 + d) v = v - 2 * (v * 1.0 * 1.0);
              fab(i, j, k, c) = v;
            }
          }
    }
  }

  void ApplyBc(const BoundaryCondition& bc) {
    if (bc.kind == kTransmissive) {
      FillTransmissive(bc);
    } else {
      FillReflective(bc);
    }
  }

  // src/AMReX/IntegratorContext.cpp:347-359 + BoundarySet.cpp:35-41
  void ApplyBoundaryCondition(int dir) {
    for (const BoundaryCondition& bc : bcs) {
      if (bc.dir == dir) ApplyBc(bc);
    }
  }

  // src/AMReX/IntegratorContext.cpp:422-444: FillPatchSingleLevel with the
  // scratch as its own source = FillBoundary + physical BC functor (all BCs
  // of the BoundarySet in order, BoundarySet.cpp:26-32).
  void FillGhostLayerSingleLevel() {
    FillBoundary();
    for (const BoundaryCondition& bc : bcs) ApplyBc(bc);
  }

  // -------------------------------------------------------------------------
  struct TileRef {
    size_t fab;
    Tile tile;
  };
  std::vector<TileRef> Tiles(const int tile_size[3], int nodal_dir) const {
    std::vector<TileRef> out;
    for (size_t b = 0; b < boxes.size(); ++b) {
      std::vector<Tile> ts;
      int ngrow[3];
      for (int d = 0; d < 3; ++d) {
        ngrow[d] = (nodal_dir >= 0 && d == nodal_dir) ? flux_gcw : scratch_gcw;
      }
      build_tiles(ts, boxes[b], rank, tile_size, ngrow, nodal_dir);
      for (const Tile& t : ts) out.push_back({b, t});
    }
    return out;
  }

  // include/fub/AMReX/FluxMethodAdapter.hpp:69-114 (tile size 1024000,8,8;
  // growntilebox; min starts at DBL_MAX)
  template <int D> double ComputeStableDtD(int dir) const {
    const int tile_size[3] = {1024000, 8, 8};
    const std::vector<TileRef> tiles = Tiles(tile_size, -1);
    const double dxd = dx[dir];
    double min_dt = std::numeric_limits<double>::max();
    with_method<D>(method, eq, [&](const auto& m) {
#pragma omp parallel for schedule(dynamic) reduction(min : min_dt)
      for (size_t t = 0; t < tiles.size(); ++t) {
        const double dt = fvso::ComputeStableDt(m, scratch[tiles[t].fab],
                                          tiles[t].tile.grown, dxd, dir);
        min_dt = std::min(min_dt, dt);
      }
    });
    return min_dt;
  }
  double ComputeStableDt(int dir) const {
    if (has_eb) {
      return rank == 2 ? ComputeStableDtEbD<2>(dir) : ComputeStableDtEbD<3>(dir);
    }
    switch (rank) {
    case 1: return ComputeStableDtD<1>(dir);
    case 2: return ComputeStableDtD<2>(dir);
    default: return ComputeStableDtD<3>(dir);
    }
  }

  // include/fub/AMReX/FluxMethodAdapter.hpp:120-150 (ForEachFab tiles
  // 1024000,16,32 over the face-centred flux MultiFab)
  template <int D> void ComputeNumericFluxesD(double dt, int dir) {
    const int tile_size[3] = {1024000, 16, 32};
    const std::vector<TileRef> tiles = Tiles(tile_size, dir);
    const double dxd = dx[dir];
    const int w = method.stencil_width();
    with_method<D>(method, eq, [&](const auto& m) {
#pragma omp parallel for schedule(dynamic)
      for (size_t t = 0; t < tiles.size(); ++t) {
        const size_t b = tiles[t].fab;
        Box cells, faces;
        GetCellsAndFacesInStencilRange(cells, faces, scratch[b].box,
                                       tiles[t].tile.grown, w, dir);
        fvso::ComputeNumericFluxes(m, fluxes[dir][b], faces, scratch[b], dt, dxd,
                             dir);
      }
    });
  }
  void ComputeNumericFluxes_(double dt, int dir) {
    if (has_eb) {
      if (rank == 2) {
        ComputeNumericFluxesEbD<2>(dt, dir);
      } else {
        ComputeNumericFluxesEbD<3>(dt, dir);
      }
      return;
    }
    switch (rank) {
    case 1: ComputeNumericFluxesD<1>(dt, dir); break;
    case 2: ComputeNumericFluxesD<2>(dt, dir); break;
    default: ComputeNumericFluxesD<3>(dt, dir); break;
    }
  }

  // src/AMReX/TimeIntegrator.cpp:29-67
  void UpdateConservatively_(double dt, int dir) {
    if (has_eb) return UpdateConservativelyEb(dt, dir);
    const int tile_size[3] = {1024000, 16, 32};
    const std::vector<TileRef> tiles = Tiles(tile_size, -1);
#pragma omp parallel for schedule(dynamic)
    for (size_t t = 0; t < tiles.size(); ++t) {
      const size_t b = tiles[t].fab;
      const Box all_faces = surrounding_nodes(tiles[t].tile.grown, dir);
      const Box flux_box = intersect(all_faces, fluxes[dir][b].box);
      const Box cell_box = enclosed_cells(flux_box, dir);
      fvso::UpdateConservatively(scratch[b], scratch[b], fluxes[dir][b], cell_box,
                           ncons(), dt, dx[dir], dir);
    }
  }

  // include/fub/AMReX/Reconstruction.hpp:44-60
  void CompleteFromCons_() {
    if (has_eb) return CompleteFromConsEb();
    const int tile_size[3] = {1024000, 16, 32};
    const std::vector<TileRef> tiles = Tiles(tile_size, -1);
#pragma omp parallel for schedule(dynamic)
    for (size_t t = 0; t < tiles.size(); ++t) {
      const size_t b = tiles[t].fab;
      switch (rank) {
      case 1: CompleteFromConsBox<1>(eq, scratch[b], tiles[t].tile.grown); break;
      case 2: CompleteFromConsBox<2>(eq, scratch[b], tiles[t].tile.grown); break;
      default: CompleteFromConsBox<3>(eq, scratch[b], tiles[t].tile.grown); break;
      }
    }
  }

  // -------------------------------------------------------------------------
  // include/fub/solver/DimensionalSplitLevelIntegrator.hpp:204-219
  double ComputeStableDtAllDirections() const {
    double min_dt = std::numeric_limits<double>::max();
    for (int d = 0; d < rank; ++d) {
      min_dt = std::min(min_dt, ComputeStableDt(d));
    }
    return min_dt;
  }

  using AdvanceFn = std::function<std::optional<TimeStepTooLarge>(double)>;

  // src/split_method/{Godunov,Strang,StrangLumped}Splitting.cpp and the
  // recursion in include/fub/split_method/SplittingMethod.hpp:53-75
  static std::optional<TimeStepTooLarge>
  Split2(int kind, double dt, const AdvanceFn& a1, const AdvanceFn& a2) {
    if (kind == kGodunovSplitting) {
      if (auto r = a1(dt)) return r;
      return a2(dt);
    } else if (kind == kStrangSplitting) {
      if (auto r = a1(0.5 * dt)) return r;
      if (auto r = a2(0.5 * dt)) return r;
      if (auto r = a2(0.5 * dt)) return r;
      return a1(0.5 * dt);
    } else {
      if (auto r = a1(0.5 * dt)) return r;
      if (auto r = a2(dt)) return r;
      return a1(0.5 * dt);
    }
  }
  static std::optional<TimeStepTooLarge>
  Split3(int kind, double dt, const AdvanceFn& a1, const AdvanceFn& a2,
         const AdvanceFn& a3) {
    if (kind == kStrangSplitting) {
      // StrangSplitting.cpp:44-69 overrides the three-operator case
      if (auto r = a1(0.5 * dt)) return r;
      if (auto r = a2(0.5 * dt)) return r;
      if (auto r = a3(0.5 * dt)) return r;
      if (auto r = a3(0.5 * dt)) return r;
      if (auto r = a2(0.5 * dt)) return r;
      return a1(0.5 * dt);
    }
    // SplittingMethod.hpp:53-60: Advance(dt, a1, [a2, a3 recursively])
    AdvanceFn inner = [&](double t) { return Split2(kind, t, a2, a3); };
    return Split2(kind, dt, a1, inner);
  }

  // include/fub/solver/DimensionalSplitLevelIntegrator.hpp:221-276
  // hook of IntegratorContext::AccumulateCoarseFineFluxes(level, scale, dir),
  // called after every sweep (DimensionalSplitLevelIntegrator.hpp:258-260)
  std::function<void(double, int)> accumulate_hook;

  std::optional<TimeStepTooLarge>
  AdvanceLevelNonRecursively(double dt, int split_kind,
                             std::vector<double>* sweep_dts = nullptr,
                             long cycles0 = -1, int subcycle_first = 0) {
    // MakeSplitDirections, include/fub/split_method/SplittingMethod.hpp:115-126:
    // the cycle counter of level 0 and the subcycle index
    const long cyc = cycles0 >= 0 ? cycles0 : cycles;
    const int is_odd = int((cyc % 2 + subcycle_first) % 2);
    int dirs[3];
    for (int i = 0; i < rank; ++i) dirs[i] = is_odd ? (rank - 1 - i) : i;
    int counts[3] = {0, 0, 0};
    auto make = [&](int slot) -> AdvanceFn {
      const int dir = dirs[slot];
      return [this, dir, slot, &counts,
              sweep_dts](double split_dt) -> std::optional<TimeStepTooLarge> {
        if (counts[slot] > 0) {
          ApplyBoundaryCondition(dir);
        }
        counts[slot] += 1;
        const double level_dt = ComputeStableDt(dir); // MinAll: one process
        if (sweep_dts) sweep_dts->push_back(level_dt);
        if (level_dt < split_dt) {
          return TimeStepTooLarge{level_dt};
        }
        ComputeNumericFluxes_(split_dt, dir);
        UpdateConservatively_(split_dt, dir);
        CompleteFromCons_();
        if (accumulate_hook) accumulate_hook(split_dt, dir);
        return std::nullopt;
      };
    };
    if (rank == 1) {
      return make(0)(dt);
    } else if (rank == 2) {
      return Split2(split_kind, dt, make(0), make(1));
    }
    return Split3(split_kind, dt, make(0), make(1), make(2));
  }

  // include/fub/solver/SubcycleFineFirstSolver.hpp:114-193 for one level:
  // AdvanceLevelNonRecursively, cycles += 1, time += dt
  // (src/AMReX/IntegratorContext.cpp:653-676), then the ghost refill every
  // `regrid_frequency` cycles (:177-192).
  std::optional<TimeStepTooLarge>
  AdvanceHierarchy(double dt, int split_kind,
                   std::vector<double>* sweep_dts = nullptr) {
    PreAdvanceHierarchy(); // cutcell::IntegratorContext::PreAdvanceLevel, :802-806
    if (auto r = AdvanceLevelNonRecursively(dt, split_kind, sweep_dts)) {
      return r;
    }
    cycles += 1;
    time_point = time_point + dt;
    if (cycles % regrid_frequency == 0) {
      FillGhostLayerSingleLevel();
    }
    return std::nullopt;
  }
};

struct RunOptions {
  double cfl{0.8};
  double final_time{1.0};
  long max_cycles{-1};
  double smallest_time_step_size{1e-12};
  int split_kind{kStrangSplitting};
};

// include/fub/RunSimulation.hpp:83-191 without output: stable dt -> limited
// dt -> AdvanceHierarchy -> on TimeStepTooLarge restore the backup and retry
// with the failure dt.
inline int RunSimulation(Level& level, const RunOptions& opt,
                         std::vector<double>* dt_sequence,
                         std::vector<double>* sweep_dts = nullptr) {
  const double eps = opt.smallest_time_step_size;
  Level backup = level;
  std::optional<double> failure_dt;
  int retries = 0;
  while (level.time_point + eps < opt.final_time &&
         (opt.max_cycles < 0 || level.cycles < opt.max_cycles)) {
    const double stable_dt =
        failure_dt ? *failure_dt : level.ComputeStableDtAllDirections();
    const double limited_dt =
        std::min(opt.final_time - level.time_point, opt.cfl * stable_dt);
    auto result = level.AdvanceHierarchy(limited_dt, opt.split_kind, sweep_dts);
    if (result) {
      failure_dt = result->dt;
      level = backup;
      ++retries;
      if (retries > 100) throw std::runtime_error("too many CFL retries");
    } else {
      if (dt_sequence) dt_sequence->push_back(limited_dt);
      failure_dt.reset();
      backup = level;
    }
  }
  return retries;
}

} // namespace fvso
