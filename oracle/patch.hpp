// TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.
//
// oracle/patch.hpp: CPU restatement of the reference's view-lifting layer
// (fub::FluxMethod<Base>, HyperbolicPatchIntegrator, CompleteFromConsFn) and
// of the index-box arithmetic the AMReX adapters do around it.
//
// Memory layout of a fab: Fortran order, x fastest, then y, z, component
// slowest; the box includes ghost cells (include/fub/core/mdspan.hpp:731-732,
// include/fub/AMReX/ViewFArrayBox.hpp:74-92).
//
// AMReX itself is absent from /root/reference (unpinned `development` branch,
// conan/conan-amrex/conanfile.py:29).  The pieces of its published algorithm
// the hot path depends on are restated here and marked "AMReX:".
#pragma once
#include "numerics.hpp"

#include <algorithm>
#include <cstring>
#include <vector>

namespace fvso {

struct Box {
  int lo[3]{0, 0, 0};
  int hi[3]{0, 0, 0}; // inclusive
  bool ok() const {
    return lo[0] <= hi[0] && lo[1] <= hi[1] && lo[2] <= hi[2];
  }
  long len(int d) const { return long(hi[d]) - lo[d] + 1; }
  long size() const { return ok() ? len(0) * len(1) * len(2) : 0; }
  bool contains(int i, int j, int k) const {
    return lo[0] <= i && i <= hi[0] && lo[1] <= j && j <= hi[1] &&
           lo[2] <= k && k <= hi[2];
  }
};

inline Box intersect(const Box& a, const Box& b) {
  Box r;
  for (int d = 0; d < 3; ++d) {
    r.lo[d] = std::max(a.lo[d], b.lo[d]);
    r.hi[d] = std::min(a.hi[d], b.hi[d]);
  }
  return r;
}
inline Box grow(Box b, int d, int n) {
  b.lo[d] -= n;
  b.hi[d] += n;
  return b;
}
inline Box grow_all(Box b, int rank, int n) {
  for (int d = 0; d < rank; ++d) b = grow(b, d, n);
  return b;
}
inline Box shift(Box b, int d, int n) {
  b.lo[d] += n;
  b.hi[d] += n;
  return b;
}
// AMReX: surroundingNodes(cell box, d) -- faces lo..hi+1
inline Box surrounding_nodes(Box b, int d) {
  b.hi[d] += 1;
  return b;
}
// AMReX: enclosedCells(face box in d) -- cells lo..hi-1
inline Box enclosed_cells(Box b, int d) {
  b.hi[d] -= 1;
  return b;
}

struct Fab {
  std::vector<double> data;
  Box box;
  int ncomp{0};
  long sy{0}, sz{0}, sc{0};
  void define(const Box& b, int nc) {
    box = b;
    ncomp = nc;
    sy = b.len(0);
    sz = sy * b.len(1);
    sc = sz * b.len(2);
    data.assign(size_t(sc) * nc, 0.0);
  }
  long index(int i, int j, int k) const {
    return (i - box.lo[0]) + sy * (j - box.lo[1]) + sz * (k - box.lo[2]);
  }
  double& operator()(int i, int j, int k, int c) {
    return data[size_t(index(i, j, k) + sc * c)];
  }
  const double& operator()(int i, int j, int k, int c) const {
    return data[size_t(index(i, j, k) + sc * c)];
  }
  long stride(int d) const { return d == 0 ? 1 : (d == 1 ? sy : sz); }
};

// ---------------------------------------------------------------------------
// AMReX: FabArrayBase::buildTileArray + MFIter::tilebox / growntilebox.
// A valid box of length L in direction d is cut into nt = max(L / tile, 1)
// tiles (integer division) of size L / nt, the first L - nt * (L / nt) tiles
// one cell longer.  For a fab that is nodal in direction `nodal_dir` every
// tile but the last gives up its upper face.  growntilebox() extends a tile
// by the ghost width on each side where it touches the valid box boundary.
struct Tile {
  Box tile;  // in the index space of the fab (cells, or faces in nodal_dir)
  Box grown; // growntilebox()
};

inline void build_tiles(std::vector<Tile>& out, const Box& valid_cells,
                        int rank, const int tile_size[3], const int ngrow[3],
                        int nodal_dir /* -1: cell centred */) {
  int nt[3] = {1, 1, 1}, tsize[3] = {1, 1, 1}, nleft[3] = {0, 0, 0};
  for (int d = 0; d < rank; ++d) {
    const int ncells = int(valid_cells.len(d));
    nt[d] = std::max(ncells / tile_size[d], 1);
    tsize[d] = ncells / nt[d];
    nleft[d] = ncells - nt[d] * tsize[d];
  }
  Box valid = valid_cells;
  if (nodal_dir >= 0) valid = surrounding_nodes(valid, nodal_dir);
  for (int kz = 0; kz < nt[2]; ++kz)
    for (int jy = 0; jy < nt[1]; ++jy)
      for (int ix = 0; ix < nt[0]; ++ix) {
        const int ijk[3] = {ix, jy, kz};
        Tile t;
        for (int d = 0; d < 3; ++d) {
          if (d >= rank) {
            t.tile.lo[d] = valid_cells.lo[d];
            t.tile.hi[d] = valid_cells.hi[d];
            continue;
          }
          int small, big;
          if (ijk[d] < nleft[d]) {
            small = ijk[d] * (tsize[d] + 1);
            big = small + tsize[d];
          } else {
            small = ijk[d] * tsize[d] + nleft[d];
            big = small + tsize[d] - 1;
          }
          t.tile.lo[d] = small + valid_cells.lo[d];
          t.tile.hi[d] = big + valid_cells.lo[d];
        }
        if (nodal_dir >= 0) {
          // convert to nodal: hi+1, then drop the upper face unless last tile
          t.tile.hi[nodal_dir] += 1;
          if (t.tile.hi[nodal_dir] < valid.hi[nodal_dir]) {
            t.tile.hi[nodal_dir] -= 1;
          }
        }
        t.grown = t.tile;
        for (int d = 0; d < rank; ++d) {
          if (t.grown.lo[d] == valid.lo[d]) t.grown.lo[d] -= ngrow[d];
          if (t.grown.hi[d] == valid.hi[d]) t.grown.hi[d] += ngrow[d];
        }
        out.push_back(t);
      }
}

// ---------------------------------------------------------------------------
// Loading/storing chunks of 8 lanes along x.
// include/fub/StateArray.hpp (Load/LoadN/Store/StoreN): LoadN only touches the
// first n lanes, the others keep stale data.
template <int D>
inline void load_complete(Complete<A8, D>& q, const Fab& f, long off, int n) {
  const double* p = f.data.data() + off;
  for (int l = 0; l < n; ++l) q.rho.v[l] = p[l];
  for (int d = 0; d < D; ++d) {
    const double* pm = p + f.sc * (1 + d);
    for (int l = 0; l < n; ++l) q.mom[d].v[l] = pm[l];
  }
  const double* pe = p + f.sc * (1 + D);
  for (int l = 0; l < n; ++l) q.E.v[l] = pe[l];
  const double* pp = p + f.sc * (2 + D);
  for (int l = 0; l < n; ++l) q.p.v[l] = pp[l];
  const double* pa = p + f.sc * (3 + D);
  for (int l = 0; l < n; ++l) q.a.v[l] = pa[l];
}
template <int D>
inline void load_cons(Cons<A8, D>& q, const Fab& f, long off, int n) {
  const double* p = f.data.data() + off;
  for (int l = 0; l < n; ++l) q.rho.v[l] = p[l];
  for (int d = 0; d < D; ++d) {
    const double* pm = p + f.sc * (1 + d);
    for (int l = 0; l < n; ++l) q.mom[d].v[l] = pm[l];
  }
  const double* pe = p + f.sc * (1 + D);
  for (int l = 0; l < n; ++l) q.E.v[l] = pe[l];
}
template <int D>
inline void store_cons(Fab& f, long off, const Cons<A8, D>& q, int n) {
  double* p = f.data.data() + off;
  for (int l = 0; l < n; ++l) p[l] = q.rho.v[l];
  for (int d = 0; d < D; ++d) {
    double* pm = p + f.sc * (1 + d);
    for (int l = 0; l < n; ++l) pm[l] = q.mom[d].v[l];
  }
  double* pe = p + f.sc * (1 + D);
  for (int l = 0; l < n; ++l) pe[l] = q.E.v[l];
}
template <int D>
inline void store_complete(Fab& f, long off, const Complete<A8, D>& q, int n) {
  store_cons<D>(f, off, q, n);
  double* pp = f.data.data() + off + f.sc * (2 + D);
  for (int l = 0; l < n; ++l) pp[l] = q.p.v[l];
  double* pa = f.data.data() + off + f.sc * (3 + D);
  for (int l = 0; l < n; ++l) pa[l] = q.a.v[l];
}
inline void load_scalar(A8& a, const Fab& f, long off, int comp, int n) {
  const double* p = f.data.data() + off + f.sc * comp;
  for (int l = 0; l < n; ++l) a.v[l] = p[l];
}

// Tail lanes beyond n are never used for a result (StoreN / mask discard
// them); give them benign content instead of the reference's stale
// data so that no NaN/denormal noise is generated.
template <int D> inline void init_benign(Complete<A8, D>& q) {
  q.rho = A8(1.0);
  for (int d = 0; d < D; ++d) q.mom[d] = A8(0.0);
  q.E = A8(1.0);
  q.p = A8(1.0);
  q.a = A8(1.0);
}

// ---------------------------------------------------------------------------
// fub::FluxMethod<Base>::ComputeNumericFluxes(SimdTag)
// include/fub/flux_method/FluxMethod.hpp:246-293.
// `faces` is the face box (face f in `dir` lies between cells f-1 and f); the
// stencil of face f is cells f-w .. f+w-1.
template <typename Method>
void ComputeNumericFluxes(const Method& m, Fab& fluxes, const Box& faces,
                          const Fab& states, double dt, double dx, int dir) {
  constexpr int D = Method::kDim;
  constexpr int W = Method::kWidth;
  constexpr int S = 2 * W;
  if (!faces.ok()) return;
  Complete<A8, D> stencil[S];
  for (int s = 0; s < S; ++s) init_benign<D>(stencil[s]);
  Cons<A8, D> flux;
  const long sstride = states.stride(dir);
  for (int k = faces.lo[2]; k <= faces.hi[2]; ++k)
    for (int j = faces.lo[1]; j <= faces.hi[1]; ++j) {
      int i = faces.lo[0];
      int n = faces.hi[0] - i + 1;
      while (n >= kChunk) {
        const long c0 = states.index(i, j, k) - W * sstride;
        for (int s = 0; s < S; ++s) {
          load_complete<D>(stencil[s], states, c0 + s * sstride, kChunk);
        }
        m.ComputeNumericFlux(flux, stencil, dt, dx, dir);
        store_cons<D>(fluxes, fluxes.index(i, j, k), flux, kChunk);
        i += kChunk;
        n = faces.hi[0] - i + 1;
      }
      // tail (:286-290); executed even for n == 0 in the reference, a no-op
      if (n > 0) {
        const long c0 = states.index(i, j, k) - W * sstride;
        for (int s = 0; s < S; ++s) {
          load_complete<D>(stencil[s], states, c0 + s * sstride, n);
        }
        m.ComputeNumericFlux(flux, stencil, dt, dx, dir);
        store_cons<D>(fluxes, fluxes.index(i, j, k), flux, n);
      }
    }
}

// fub::FluxMethod<Base>::ComputeStableDt(SimdTag)
// include/fub/flux_method/FluxMethod.hpp:341-382.  `cells` is the (tile) cell
// box handed in by the adapter; stencil positions are all i with
// i .. i+2w-1 inside the box along `dir`.
template <typename Method>
double ComputeStableDt(const Method& m, const Fab& states, const Box& cells,
                       double dx, int dir) {
  constexpr int D = Method::kDim;
  constexpr int W = Method::kWidth;
  constexpr int S = 2 * W;
  A8 min_dt(std::numeric_limits<double>::infinity());
  Box pos = cells;
  pos.hi[dir] -= (S - 1);
  if (!pos.ok()) return min_coeff(min_dt);
  Complete<A8, D> stencil[S];
  for (int s = 0; s < S; ++s) init_benign<D>(stencil[s]);
  const long sstride = states.stride(dir);
  for (int k = pos.lo[2]; k <= pos.hi[2]; ++k)
    for (int j = pos.lo[1]; j <= pos.hi[1]; ++j) {
      int i = pos.lo[0];
      int n = pos.hi[0] - i + 1;
      while (n >= kChunk) {
        const long c0 = states.index(i, j, k);
        for (int s = 0; s < S; ++s) {
          load_complete<D>(stencil[s], states, c0 + s * sstride, kChunk);
        }
        min_dt = vmin(min_dt, m.ComputeStableDt(stencil, dx, dir));
        i += kChunk;
        n = pos.hi[0] - i + 1;
      }
      if (n > 0) {
        const long c0 = states.index(i, j, k);
        for (int s = 0; s < S; ++s) {
          load_complete<D>(stencil[s], states, c0 + s * sstride, n);
        }
        A8 lane;
        for (int l = 0; l < kChunk; ++l) lane.v[l] = double(l);
        const M8 mask = lane < double(n);
        min_dt = select(mask, vmin(min_dt, m.ComputeStableDt(stencil, dx, dir)),
                        min_dt);
      }
    }
  return min_coeff(min_dt);
}

// ---------------------------------------------------------------------------
// HyperbolicPatchIntegrator<SimdTag>::UpdateConservatively
// src/HyperbolicPatchIntegrator.cpp:35-85: per component and per x-row
//   next = prev + (dt/dx) * (fL - fR),  dt/dx one division (:78).
inline void UpdateConservatively(Fab& next, const Fab& prev, const Fab& fluxes,
                                 const Box& cells, int ncons, double dt,
                                 double dx, int dir) {
  if (!cells.ok()) return;
  const double dt_dx = dt / dx;
  const long fs = fluxes.stride(dir);
  for (int c = 0; c < ncons; ++c)
    for (int k = cells.lo[2]; k <= cells.hi[2]; ++k)
      for (int j = cells.lo[1]; j <= cells.hi[1]; ++j) {
        const int i0 = cells.lo[0];
        const int n = int(cells.len(0));
        double* out = &next(i0, j, k, c);
        const double* in = &prev(i0, j, k, c);
        const double* fL = &fluxes(i0, j, k, c);
        const double* fR = fL + fs;
        for (int i = 0; i < n; ++i) {
          out[i] = in[i] + dt_dx * (fL[i] - fR[i]);
        }
      }
}

// CompleteFromConsFn::CompleteFromCons(SimdTag)
// include/fub/CompleteFromCons.hpp:141-179
template <int D>
void CompleteFromConsBox(const PerfectGasEq& eq, Fab& fab, const Box& cells) {
  if (!cells.ok()) return;
  Cons<A8, D> cons;
  Complete<A8, D> complete;
  std::memset(static_cast<void*>(&cons), 0, sizeof(cons));
  for (int l = 0; l < kChunk; ++l) cons.rho.v[l] = 1.0;
  for (int k = cells.lo[2]; k <= cells.hi[2]; ++k)
    for (int j = cells.lo[1]; j <= cells.hi[1]; ++j) {
      int i = cells.lo[0];
      int n = cells.hi[0] - i + 1;
      while (n >= kChunk) {
        load_cons<D>(cons, fab, fab.index(i, j, k), kChunk);
        CompleteFromCons<A8, D>(eq, complete, cons);
        store_complete<D>(fab, fab.index(i, j, k), complete, kChunk);
        i += kChunk;
        n = cells.hi[0] - i + 1;
      }
      if (n > 0) {
        load_cons<D>(cons, fab, fab.index(i, j, k), n);
        CompleteFromCons<A8, D>(eq, complete, cons);
        store_complete<D>(fab, fab.index(i, j, k), complete, n);
      }
    }
}

// src/AMReX/ViewFArrayBox.cpp:66-79
inline void GetCellsAndFacesInStencilRange(Box& cells, Box& faces,
                                           const Box& cell_validbox,
                                           const Box& face_tilebox,
                                           int stencil_width, int dir) {
  cells = intersect(cell_validbox,
                    grow(enclosed_cells(face_tilebox, dir), dir, stencil_width));
  faces = intersect(face_tilebox,
                    surrounding_nodes(grow(cells, dir, -stencil_width), dir));
}

} // namespace fvso
