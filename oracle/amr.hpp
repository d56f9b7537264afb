// TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.
//
// oracle/amr.hpp: CPU restatement of the coarse-fine operations of the
// reference's IntegratorContext for two levels that live in one process:
//   CoarsenConservatively      src/AMReX/IntegratorContext.cpp:451-470  (amrex::average_down)
//   AccumulateCoarseFineFluxes :472-497 (FluxRegister::FineAdd / CrseInit(ADD))
//   ResetCoarseFineFluxes      :499-512 (ClearInternalBorders + setVal(0))
//   ApplyFluxCorrection        :514-531 (FluxRegister::Reflux)
//   FillGhostLayerTwoLevels    :380-420 (FillPatchTwoLevels with pc_interp)
//
// PARITY UNPINNED: the arithmetic lives in AMReX (`development`, unpinned,
// absent from /root/reference: conan/conan-amrex/conanfile.py:29) and the
// reference has no test with numeric assertions for it.  Restated from the
// published algorithms: average_down sums the fine cells of a coarse cell in
// (k, j, i) order and multiplies by 1 / ratio^D; a flux register holds, per
// fine box and per side, the coarse faces of the box boundary, CrseInit adds
// mult * coarse flux, FineAdd adds mult * fine flux for every fine face of the
// coarse face in (k, j, i) order, Reflux adds -+ scale * register / volume to
// the coarse cell outside the fine box; pc_interp copies the coarse parent.
#pragma once
#include "level.hpp"

namespace fvso {

inline int coarsen_index(int i, int r) { return i >= 0 ? i / r : -((-i + r - 1) / r); }
inline Box coarsen(const Box& b, const int ratio[3]) {
  Box c;
  for (int d = 0; d < 3; ++d) {
    c.lo[d] = coarsen_index(b.lo[d], ratio[d]);
    c.hi[d] = coarsen_index(b.hi[d], ratio[d]);
  }
  return c;
}
inline Box refine(const Box& b, const int ratio[3]) {
  Box f;
  for (int d = 0; d < 3; ++d) {
    f.lo[d] = b.lo[d] * ratio[d];
    f.hi[d] = (b.hi[d] + 1) * ratio[d] - 1;
  }
  return f;
}

// One register: the coarse faces of side `side` (0 lower, 1 upper) of fine box
// `fine_box` in direction `dir`; face index = index of the cell to its right.
struct FluxRegisterFab {
  int fine_box, dir, side;
  Fab reg;                         // ncons components on the face plane
  std::vector<unsigned char> open; // 1 where the outside coarse cell is not covered by a fine box
};

// amrex::FluxRegister::CrseInit / Reflux copy with geom.periodicity(): the outside
// cell of a register face on a periodic domain boundary is the cell on the other side
// of the domain.  Returns the shift (a multiple of the domain length along dir) that
// takes `cell` into the domain, 0 if the direction is not periodic.
inline int periodic_shift(const Level& coarse, int dir, int cell) {
  if (!coarse.periodic[dir]) return 0;
  const int len = coarse.domain.hi[dir] - coarse.domain.lo[dir] + 1;
  if (cell < coarse.domain.lo[dir]) return len;
  if (cell > coarse.domain.hi[dir]) return -len;
  return 0;
}

struct CoarseFine {
  int ratio[3]{2, 2, 2};
  std::vector<FluxRegisterFab> regs;

  void define(const Level& fine, const Level& coarse) {
    regs.clear();
    for (int d = fine.rank; d < 3; ++d) ratio[d] = 1;
    for (size_t f = 0; f < fine.boxes.size(); ++f) {
      const Box cb = coarsen(fine.boxes[f], ratio);
      for (int d = 0; d < fine.rank; ++d)
        for (int side = 0; side < 2; ++side) {
          FluxRegisterFab r;
          r.fine_box = int(f);
          r.dir = d;
          r.side = side;
          Box fb = cb;
          fb.lo[d] = fb.hi[d] = side == 0 ? cb.lo[d] : cb.hi[d] + 1;
          r.reg.define(fb, fine.ncons());
          r.open.assign(size_t(fb.size()), 1);
          // ClearInternalBorders: faces whose outside cell lies under another fine box
          size_t n = 0;
          for (int k = fb.lo[2]; k <= fb.hi[2]; ++k)
            for (int j = fb.lo[1]; j <= fb.hi[1]; ++j)
              for (int i = fb.lo[0]; i <= fb.hi[0]; ++i, ++n) {
                int out[3] = {i, j, k};
                if (side == 0) out[d] -= 1;
                out[d] += periodic_shift(coarse, d, out[d]);
                for (size_t g = 0; g < fine.boxes.size(); ++g) {
                  if (coarsen(fine.boxes[g], ratio).contains(out[0], out[1], out[2])) r.open[n] = 0;
                }
                if (!coarse.domain.contains(out[0], out[1], out[2])) r.open[n] = 0;
              }
          regs.push_back(std::move(r));
        }
    }
  }

  void Reset() {
    for (FluxRegisterFab& r : regs) std::fill(r.reg.data.begin(), r.reg.data.end(), 0.0);
  }

  // FineAdd: level == fine
  void FineAdd(const Level& fine, int dir, double mult) {
    for (FluxRegisterFab& r : regs) {
      if (r.dir != dir) continue;
      const Fab& flux = fine.fluxes[dir][size_t(r.fine_box)];
      const Box& fb = r.reg.box;
      for (int c = 0; c < r.reg.ncomp; ++c)
        for (int k = fb.lo[2]; k <= fb.hi[2]; ++k)
          for (int j = fb.lo[1]; j <= fb.hi[1]; ++j)
            for (int i = fb.lo[0]; i <= fb.hi[0]; ++i) {
              int lo[3] = {i * ratio[0], j * ratio[1], k * ratio[2]};
              int hi[3] = {lo[0] + ratio[0] - 1, lo[1] + ratio[1] - 1, lo[2] + ratio[2] - 1};
              hi[dir] = lo[dir];
              for (int kk = lo[2]; kk <= hi[2]; ++kk)
                for (int jj = lo[1]; jj <= hi[1]; ++jj)
                  for (int ii = lo[0]; ii <= hi[0]; ++ii) {
                    r.reg(i, j, k, c) += mult * flux(ii, jj, kk, c);
                  }
            }
    }
  }
  // CrseInit(..., ADD): level == coarse
  void CrseInit(const Level& coarse, int dir, double mult) {
    for (FluxRegisterFab& r : regs) {
      if (r.dir != dir) continue;
      const Box& fb = r.reg.box;
      for (size_t cbx = 0; cbx < coarse.boxes.size(); ++cbx) {
        const Fab& flux = coarse.fluxes[dir][cbx];
        for (int c = 0; c < r.reg.ncomp; ++c)
          for (int k = fb.lo[2]; k <= fb.hi[2]; ++k)
            for (int j = fb.lo[1]; j <= fb.hi[1]; ++j)
              for (int i = fb.lo[0]; i <= fb.hi[0]; ++i) {
                // a face shared by two coarse boxes is initialised by the box
                // that owns the OUTSIDE cell (its periodic image on a periodic
                // boundary; the face then is that box's face on the other side)
                int out[3] = {i, j, k};
                if (r.side == 0) out[dir] -= 1;
                const int shift = periodic_shift(coarse, dir, out[dir]);
                out[dir] += shift;
                if (!coarse.boxes[cbx].contains(out[0], out[1], out[2])) continue;
                int face[3] = {i, j, k};
                face[dir] += shift;
                r.reg(i, j, k, c) += mult * flux(face[0], face[1], face[2], c);
              }
      }
    }
  }
  // Reflux into the coarse scratch
  void Reflux(Level& coarse, int dir, double scale) {
    const double vol = coarse.dx[0] * (coarse.rank > 1 ? coarse.dx[1] : 1.0) *
                       (coarse.rank > 2 ? coarse.dx[2] : 1.0);
    for (FluxRegisterFab& r : regs) {
      if (r.dir != dir) continue;
      const Box& fb = r.reg.box;
      for (size_t cbx = 0; cbx < coarse.boxes.size(); ++cbx) {
        Fab& S = coarse.scratch[cbx];
        size_t n = 0;
        for (int k = fb.lo[2]; k <= fb.hi[2]; ++k)
          for (int j = fb.lo[1]; j <= fb.hi[1]; ++j)
            for (int i = fb.lo[0]; i <= fb.hi[0]; ++i, ++n) {
              if (!r.open[n]) continue;
              int out[3] = {i, j, k};
              if (r.side == 0) out[dir] -= 1;
              out[dir] += periodic_shift(coarse, dir, out[dir]);
              if (!coarse.boxes[cbx].contains(out[0], out[1], out[2])) continue;
              for (int c = 0; c < r.reg.ncomp; ++c) {
                // lower side: the register face is the right face of the cell
                const double sign = r.side == 0 ? -1.0 : 1.0;
                S(out[0], out[1], out[2], c) += sign * scale * r.reg(i, j, k, c) / vol;
              }
            }
      }
    }
  }
};

// src/AMReX/IntegratorContext.cpp:472-497
inline void AccumulateCoarseFineFluxes(CoarseFine* to_coarser, const Level& level,
                                       CoarseFine* to_finer, double scale, int dir) {
  const double vol = level.dx[0] * (level.rank > 1 ? level.dx[1] : 1.0) *
                     (level.rank > 2 ? level.dx[2] : 1.0);
  const double vol_dt_over_dx = vol * scale / level.dx[dir];
  if (to_coarser) to_coarser->FineAdd(level, dir, vol_dt_over_dx);
  if (to_finer) to_finer->CrseInit(level, dir, -vol_dt_over_dx);
}

// amrex::average_down on the conservative components.  Cut-cell levels
// (src/AMReX/cutcell/IntegratorContext.cpp:662-677) call amrex::EB_average_down
// (S_fine, S_crse, first, size, 2) instead: fabs whose flag type is regular or
// covered take the plain average, single-valued fabs weight every fine cell with
// its volume fraction, crse = sum(fine * vfrac) / sum(vfrac), and fall back to
// the lower-left fine cell where sum(vfrac) <= 1e-30 (a covered coarse cell).
// PARITY UNPINNED like the rest of this file (AMReX's amrex_eb_avgdown restated
// from the published algorithm; AMReX decides per tile box, not per fab -- the two
// formulas agree to the bit wherever every fine cell has vfrac == 1 and wherever
// all fine data are zero, so the granularity does not show).
inline void CoarsenConservatively(const Level& fine, Level& coarse, const int ratio[3]) {
  const double volfrac = 1.0 / double(ratio[0] * ratio[1] * ratio[2]);
  for (size_t f = 0; f < fine.boxes.size(); ++f) {
    const Box cf = coarsen(fine.boxes[f], ratio);
    for (size_t cbx = 0; cbx < coarse.boxes.size(); ++cbx) {
      const Box region = intersect(cf, coarse.boxes[cbx]);
      if (!region.ok()) continue;
      const Fab& F = fine.scratch[f];
      Fab& Cc = coarse.scratch[cbx];
      if (fine.has_eb && fine.eb[f].fab_type == kFabSingleValued) {
        const Fab& vf = fine.eb[f].alpha;
        for (int n = 0; n < fine.ncons(); ++n)
          for (int k = region.lo[2]; k <= region.hi[2]; ++k)
            for (int j = region.lo[1]; j <= region.hi[1]; ++j)
              for (int i = region.lo[0]; i <= region.hi[0]; ++i) {
                double c = 0.0;
                double cv = 0.0;
                for (int kr = 0; kr < ratio[2]; ++kr)
                  for (int jr = 0; jr < ratio[1]; ++jr)
                    for (int ir = 0; ir < ratio[0]; ++ir) {
                      const int fi = i * ratio[0] + ir, fj = j * ratio[1] + jr, fk = k * ratio[2] + kr;
                      const double tmp = vf(fi, fj, fk, 0);
                      cv += tmp;
                      c += F(fi, fj, fk, n) * tmp;
                    }
                if (cv > 1.e-30) {
                  c /= cv;
                } else {
                  c = F(i * ratio[0], j * ratio[1], k * ratio[2], n);
                }
                Cc(i, j, k, n) = c;
              }
        continue;
      }
      for (int n = 0; n < fine.ncons(); ++n)
        for (int k = region.lo[2]; k <= region.hi[2]; ++k)
          for (int j = region.lo[1]; j <= region.hi[1]; ++j)
            for (int i = region.lo[0]; i <= region.hi[0]; ++i) {
              double c = 0.0;
              for (int kr = 0; kr < ratio[2]; ++kr)
                for (int jr = 0; jr < ratio[1]; ++jr)
                  for (int ir = 0; ir < ratio[0]; ++ir) {
                    c += F(i * ratio[0] + ir, j * ratio[1] + jr, k * ratio[2] + kr, n);
                  }
              Cc(i, j, k, n) = volfrac * c;
            }
    }
  }
}

// FillPatchTwoLevels with pc_interp: ghost cells inside the domain take the
// coarse parent unless a fine box (or a periodic image) holds them, then the
// physical boundary conditions of the fine level.  On a periodic domain the
// ghost cells beyond the periodic boundary take the parent from the coarse
// level's periodic image (FillPatchTwoLevels works on the domain grown in the
// periodic directions and gathers the coarse data with geom.periodicity()).
inline void FillGhostLayerTwoLevels(Level& fine, const Level& coarse, const int ratio[3]) {
  int r[3];
  for (int d = 0; d < 3; ++d) r[d] = (d < coarse.rank && coarse.periodic[d]) ? 1 : 0;
  Box fill_domain = fine.domain; // GrowPeriodicDomain
  for (int d = 0; d < fine.rank; ++d) {
    if (fine.periodic[d]) fill_domain = grow(fill_domain, d, fine.scratch_gcw);
  }
  for (size_t f = 0; f < fine.boxes.size(); ++f) {
    Fab& F = fine.scratch[f];
    for (size_t cbx = 0; cbx < coarse.boxes.size(); ++cbx)
      for (int sz = -r[2]; sz <= r[2]; ++sz)
        for (int sy = -r[1]; sy <= r[1]; ++sy)
          for (int sx = -r[0]; sx <= r[0]; ++sx) {
            // shift of the image in coarse index space
            const int s[3] = {sx * int(coarse.domain.len(0)), sy * int(coarse.domain.len(1)),
                              sz * int(coarse.domain.len(2))};
            Box image = coarse.boxes[cbx];
            for (int d = 0; d < 3; ++d) image = shift(image, d, s[d]);
            const Box region = intersect(intersect(F.box, refine(image, ratio)), fill_domain);
            if (!region.ok()) continue;
            const Fab& Cc = coarse.scratch[cbx];
            for (int n = 0; n < fine.ncomp(); ++n)
              for (int k = region.lo[2]; k <= region.hi[2]; ++k)
                for (int j = region.lo[1]; j <= region.hi[1]; ++j)
                  for (int i = region.lo[0]; i <= region.hi[0]; ++i) {
                    if (fine.boxes[f].contains(i, j, k)) continue;
                    F(i, j, k, n) = Cc(coarsen_index(i, ratio[0]) - s[0], coarsen_index(j, ratio[1]) - s[1],
                                       coarsen_index(k, ratio[2]) - s[2], n);
                  }
          }
  }
  fine.FillGhostLayerSingleLevel();
}

// GriddingAlgorithm::RemakeLevel / MakeNewLevelFromCoarse followed by
// IntegratorContext::CopyDataToScratch (src/AMReX/GriddingAlgorithm.cpp:217-260,
// src/AMReX/IntegratorContext.cpp:601-638): `fresh` is the level on its new boxes.
// FillPatchTwoLevels onto the new layout: every cell of the new scratch arrays inside
// the (periodically grown) domain takes the old level's valid data where an old box
// holds it, the coarse parent (pc_interp) elsewhere; then the physical boundaries.
// `old` may be null (the level did not exist).
inline void RegridFill(Level& fresh, const Level* old, const Level& coarse, const int ratio[3]) {
  int r[3];
  for (int d = 0; d < 3; ++d) r[d] = (d < coarse.rank && coarse.periodic[d]) ? 1 : 0;
  Box fill_domain = fresh.domain;
  for (int d = 0; d < fresh.rank; ++d) {
    if (fresh.periodic[d]) fill_domain = grow(fill_domain, d, fresh.scratch_gcw);
  }
  for (size_t f = 0; f < fresh.boxes.size(); ++f) {
    Fab& F = fresh.scratch[f];
    for (size_t cbx = 0; cbx < coarse.boxes.size(); ++cbx)
      for (int sz = -r[2]; sz <= r[2]; ++sz)
        for (int sy = -r[1]; sy <= r[1]; ++sy)
          for (int sx = -r[0]; sx <= r[0]; ++sx) {
            const int s[3] = {sx * int(coarse.domain.len(0)), sy * int(coarse.domain.len(1)),
                              sz * int(coarse.domain.len(2))};
            Box image = coarse.boxes[cbx];
            for (int d = 0; d < 3; ++d) image = shift(image, d, s[d]);
            const Box region = intersect(intersect(F.box, refine(image, ratio)), fill_domain);
            if (!region.ok()) continue;
            const Fab& Cc = coarse.scratch[cbx];
            for (int n = 0; n < fresh.ncomp(); ++n)
              for (int k = region.lo[2]; k <= region.hi[2]; ++k)
                for (int j = region.lo[1]; j <= region.hi[1]; ++j)
                  for (int i = region.lo[0]; i <= region.hi[0]; ++i) {
                    F(i, j, k, n) = Cc(coarsen_index(i, ratio[0]) - s[0], coarsen_index(j, ratio[1]) - s[1],
                                       coarsen_index(k, ratio[2]) - s[2], n);
                  }
          }
    if (!old) continue;
    for (size_t o = 0; o < old->boxes.size(); ++o)
      for (int sz = -r[2]; sz <= r[2]; ++sz)
        for (int sy = -r[1]; sy <= r[1]; ++sy)
          for (int sx = -r[0]; sx <= r[0]; ++sx) {
            const int s[3] = {sx * int(old->domain.len(0)), sy * int(old->domain.len(1)),
                              sz * int(old->domain.len(2))};
            Box image = old->boxes[o];
            for (int d = 0; d < 3; ++d) image = shift(image, d, s[d]);
            const Box region = intersect(F.box, image);
            if (!region.ok()) continue;
            const Fab& O = old->scratch[o];
            for (int n = 0; n < fresh.ncomp(); ++n)
              for (int k = region.lo[2]; k <= region.hi[2]; ++k)
                for (int j = region.lo[1]; j <= region.hi[1]; ++j)
                  for (int i = region.lo[0]; i <= region.hi[0]; ++i) {
                    F(i, j, k, n) = O(i - s[0], j - s[1], k - s[2], n);
                  }
          }
  }
  if (old) {
    fresh.cycles = old->cycles;
    fresh.time_point = old->time_point;
  }
  fresh.FillGhostLayerSingleLevel();
}

// ---------------------------------------------------------------------------
// include/fub/solver/SubcycleFineFirstSolver.hpp:100-193 for a static
// hierarchy (no regridding) held by one process.
struct Hierarchy {
  std::vector<Level*> levels;
  std::vector<CoarseFine> cf; // cf[l] couples level l (fine) with l - 1
  int ratio[3]{2, 2, 2};

  void define() {
    cf.assign(levels.size(), CoarseFine());
    for (size_t l = 1; l < levels.size(); ++l) {
      for (int d = 0; d < 3; ++d) cf[l].ratio[d] = ratio[d];
      cf[l].define(*levels[l], *levels[l - 1]);
    }
  }
  int max_ratio() const {
    int m = 1;
    for (int d = 0; d < levels[0]->rank; ++d) m = std::max(m, ratio[d]);
    return m;
  }
  int total_ratio(int level) const {
    int r = 1;
    for (int l = 1; l <= level; ++l) r *= max_ratio();
    return r;
  }
  // :100-112
  double ComputeStableDt() const {
    double min_dt = std::numeric_limits<double>::max();
    for (size_t l = 0; l < levels.size(); ++l) {
      min_dt = std::min(min_dt, total_ratio(int(l)) * levels[l]->ComputeStableDtAllDirections());
    }
    return min_dt;
  }
  // :114-175
  std::optional<TimeStepTooLarge> AdvanceLevel(int this_level, double dt, int sub_first,
                                               int split_kind, std::vector<double>* sweep_dts) {
    Level& L = *levels[size_t(this_level)];
    L.PreAdvanceHierarchy();
    const int next = this_level + 1;
    const bool has_next = next < int(levels.size());
    if (has_next) {
      cf[size_t(next)].Reset();
      const int r = max_ratio();
      for (int i = 0; i < r; ++i) {
        if (auto e = AdvanceLevel(next, dt / r, i, split_kind, sweep_dts)) return e;
      }
    }
    L.accumulate_hook = [&, this_level, has_next, next](double scale, int dir) {
      AccumulateCoarseFineFluxes(this_level > 0 ? &cf[size_t(this_level)] : nullptr, L,
                                 has_next ? &cf[size_t(next)] : nullptr, scale, dir);
    };
    auto result = L.AdvanceLevelNonRecursively(dt, split_kind, sweep_dts, levels[0]->cycles,
                                               sub_first);
    L.accumulate_hook = nullptr;
    if (result) {
      result->dt *= total_ratio(this_level);
      return result;
    }
    if (has_next) {
      for (int d = 0; d < L.rank; ++d) cf[size_t(next)].Reflux(L, d, 1.0);
      CoarsenConservatively(*levels[size_t(next)], L, cf[size_t(next)].ratio);
      L.CompleteFromCons_();
    }
    if (this_level > 0) { // PostAdvanceLevel
      L.cycles += 1;
      L.time_point = L.time_point + dt;
    }
    return std::nullopt;
  }
  // :177-193
  std::optional<TimeStepTooLarge> AdvanceHierarchy(double dt, int split_kind,
                                                   std::vector<double>* sweep_dts) {
    std::vector<Level> backup;
    for (Level* l : levels) backup.push_back(*l);
    if (auto e = AdvanceLevel(0, dt, 0, split_kind, sweep_dts)) {
      for (size_t l = 0; l < levels.size(); ++l) *levels[l] = backup[l];
      return e;
    }
    Level& L0 = *levels[0];
    L0.cycles += 1;
    L0.time_point = L0.time_point + dt;
    if (L0.cycles % L0.regrid_frequency == 0) {
      L0.FillGhostLayerSingleLevel();
      for (size_t l = 1; l < levels.size(); ++l) {
        FillGhostLayerTwoLevels(*levels[l], *levels[l - 1], cf[l].ratio);
      }
    }
    return std::nullopt;
  }
};

} // namespace fvso
