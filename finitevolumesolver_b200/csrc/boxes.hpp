// boxes.hpp -- host-side index-box arithmetic of the patch level: the pieces
// of AMReX's published box calculus the reference's adapters rely on
// (surroundingNodes / enclosedCells / grow, MFIter tiling with growntilebox,
// FillBoundary region enumeration, boxDiff order of the boundary slabs) and
// the iteration domains the adapters derive from them.  No CUDA in here.
//
// Reference: include/fub/AMReX/FluxMethodAdapter.hpp:69-150,
// src/AMReX/ViewFArrayBox.cpp:66-79, src/AMReX/TimeIntegrator.cpp:29-67,
// src/AMReX/IntegratorContext.cpp:303-324,422-449,
// src/AMReX/boundary_condition/TransmissiveBoundary.cpp:60-93,
// include/fub/AMReX/boundary_condition/ReflectiveBoundary.hpp:67-94.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <vector>

namespace fvs {

struct Box {
  int lo[3]{0, 0, 0};
  int hi[3]{0, 0, 0}; // inclusive
  bool ok() const { return lo[0] <= hi[0] && lo[1] <= hi[1] && lo[2] <= hi[2]; }
  int len(int d) const { return hi[d] - lo[d] + 1; }
  long long size() const { return ok() ? (long long)len(0) * len(1) * len(2) : 0; }
};

inline Box make_box(const int lo[3], const int hi[3]) {
  Box b;
  for (int d = 0; d < 3; ++d) {
    b.lo[d] = lo[d];
    b.hi[d] = hi[d];
  }
  return b;
}
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
inline Box grow_rank(Box b, int rank, int n) {
  for (int d = 0; d < rank; ++d) b = grow(b, d, n);
  return b;
}
inline Box shifted(Box b, const int s[3]) {
  for (int d = 0; d < 3; ++d) {
    b.lo[d] += s[d];
    b.hi[d] += s[d];
  }
  return b;
}

// Scratch array box: valid box grown by scratch_gcw in every space dimension
// (IntegratorContext.cpp:303-312).
inline Box state_box(const Box& valid, int rank, int scratch_gcw) {
  return grow_rank(valid, rank, scratch_gcw);
}
// Flux array box of direction dir: faces of the valid box, grown by flux_gcw
// along dir and by scratch_gcw across (IntegratorContext.cpp:313-324).
inline Box flux_box(const Box& valid, int rank, int dir, int scratch_gcw, int flux_gcw) {
  Box fb = valid;
  fb.hi[dir] += 1;
  for (int e = 0; e < rank; ++e) fb = grow(fb, e, e == dir ? flux_gcw : scratch_gcw);
  return fb;
}

// Faces (along dir) whose numeric flux the adapter computes: every face of the
// flux box whose 2w-cell stencil lies inside the scratch box
// (FluxMethodAdapter.hpp:136-139 with ViewFArrayBox.cpp:66-79).  Returned as an
// inclusive range of face indices; face f lies between cells f-1 and f.
inline void flux_face_range(const Box& valid, int dir, int scratch_gcw, int flux_gcw, int w,
                            int& flo, int& fhi) {
  flo = std::max(valid.lo[dir] - flux_gcw, valid.lo[dir] - scratch_gcw + w);
  fhi = std::min(valid.hi[dir] + 1 + flux_gcw, valid.hi[dir] + scratch_gcw - w + 1);
}
// Cells (along dir) that UpdateConservatively touches: both faces in the flux
// box (TimeIntegrator.cpp:41-46).
inline void update_cell_range(const Box& valid, int dir, int flux_gcw, int& ulo, int& uhi) {
  ulo = valid.lo[dir] - flux_gcw;
  uhi = valid.hi[dir] + flux_gcw;
}

// AMReX tiling of a cell-centred valid box (FabArrayBase::buildTileArray):
// nt = max(L / tile, 1) tiles of size L / nt, the first L - nt * (L / nt) one
// cell longer.  Returns the tile boundaries [lo_t, hi_t] along one dimension.
inline std::vector<std::array<int, 2>> tile_ranges(int lo, int hi, int tile_size) {
  const int ncells = hi - lo + 1;
  const int nt = std::max(ncells / tile_size, 1);
  const int tsize = ncells / nt;
  const int nleft = ncells - nt * tsize;
  std::vector<std::array<int, 2>> out;
  for (int t = 0; t < nt; ++t) {
    int small, big;
    if (t < nleft) {
      small = t * (tsize + 1);
      big = small + tsize;
    } else {
      small = t * tsize + nleft;
      big = small + tsize - 1;
    }
    out.push_back({small + lo, big + lo});
  }
  return out;
}

// Faces taking part in FluxMethodAdapter::ComputeStableDt along dir
// (FluxMethodAdapter.hpp:82-98: MFIter tiles of 1024000 x 8 x 8 with
// growntilebox; FluxMethod.hpp:347-353: every position whose 2w stencil fits
// in the tile box; MusclHancockMethod.hpp:208-218: the inner pair is used).
// mask is indexed by (cell index - state box lo); mask[i] != 0 iff the face
// between cells i and i+1 is evaluated.  The tiling across `dir` does not
// matter (all rows are covered); only the tiling ALONG dir does.
inline std::vector<unsigned char> stable_dt_face_mask(const Box& valid, int scratch_gcw, int w,
                                                      int dir) {
  const int tile_size = dir == 0 ? 1024000 : 8;
  const int lo = valid.lo[dir], hi = valid.hi[dir];
  const int n = hi - lo + 1 + 2 * scratch_gcw;
  std::vector<unsigned char> mask((size_t)n, 0);
  for (const auto& t : tile_ranges(lo, hi, tile_size)) {
    int tlo = t[0], thi = t[1];
    if (tlo == lo) tlo -= scratch_gcw; // growntilebox
    if (thi == hi) thi += scratch_gcw;
    // stencil positions s .. s+2w-1 inside [tlo, thi]; face between s+w-1 and s+w
    for (int s = tlo; s + 2 * w - 1 <= thi; ++s) {
      mask[(size_t)(s + w - 1 - (lo - scratch_gcw))] = 1;
    }
  }
  return mask;
}

// FabArray::FillBoundary (with periodic images): every ghost cell of box i
// that lies inside the valid box of box j (shifted by a period) is copied.
struct HaloCopy {
  int dst_box, src_box;
  Box region;   // in the destination's index space
  int shift[3]; // source index = destination index - shift
};
inline std::vector<HaloCopy> halo_plan(int rank, const Box& domain, const bool periodic[3],
                                       const std::vector<Box>& boxes, int scratch_gcw) {
  std::vector<std::array<int, 3>> shifts;
  int r[3];
  for (int d = 0; d < 3; ++d) r[d] = (d < rank && periodic[d]) ? 1 : 0;
  for (int sz = -r[2]; sz <= r[2]; ++sz)
    for (int sy = -r[1]; sy <= r[1]; ++sy)
      for (int sx = -r[0]; sx <= r[0]; ++sx)
        shifts.push_back({sx * domain.len(0), sy * domain.len(1), sz * domain.len(2)});
  std::vector<HaloCopy> out;
  for (size_t i = 0; i < boxes.size(); ++i) {
    const Box dst = state_box(boxes[i], rank, scratch_gcw);
    for (size_t j = 0; j < boxes.size(); ++j) {
      for (const auto& s : shifts) {
        if (i == j && s[0] == 0 && s[1] == 0 && s[2] == 0) continue;
        const Box image = shifted(boxes[j], s.data());
        const Box region = intersect(dst, image);
        if (!region.ok()) continue;
        HaloCopy c;
        c.dst_box = (int)i;
        c.src_box = (int)j;
        c.region = region;
        for (int d = 0; d < 3; ++d) c.shift[d] = s[d];
        out.push_back(c);
      }
    }
  }
  return out;
}

// Ghost slab of one physical boundary (dir, side) in the domain's index
// space.  Transmissive fills the slab over the full grown extent across dir
// (TransmissiveBoundary.cpp:60-93); reflective follows AMReX's boxDiff order:
// full grown extent in dimensions below dir, domain extent above
// (ReflectiveBoundary.hpp:67-94).
inline Box boundary_slab(int rank, const Box& domain, const bool periodic[3], int scratch_gcw,
                         int dir, int side, bool reflective) {
  Box slab = domain;
  for (int d = 0; d < rank; ++d) {
    if (!periodic[d]) slab = grow(slab, d, scratch_gcw);
  }
  if (reflective) {
    for (int e = dir + 1; e < rank; ++e) {
      slab.lo[e] = domain.lo[e];
      slab.hi[e] = domain.hi[e];
    }
  }
  if (side == 0) {
    slab.hi[dir] = domain.lo[dir] - 1;
  } else {
    slab.lo[dir] = domain.hi[dir] + 1;
  }
  return slab;
}

// Equal-count assignment of boxes to ranks in Morton order of their lower
// corners (uniform-case stand-in for amrex::DistributionMapping's SFC
// strategy).
inline std::vector<int> distribute_boxes(const std::vector<Box>& boxes, int nranks) {
  const size_t n = boxes.size();
  std::vector<size_t> order(n);
  for (size_t i = 0; i < n; ++i) order[i] = i;
  int mn[3] = {0, 0, 0};
  for (size_t i = 0; i < n; ++i)
    for (int d = 0; d < 3; ++d) mn[d] = i == 0 ? boxes[i].lo[d] : std::min(mn[d], boxes[i].lo[d]);
  auto morton = [&](const Box& b) {
    uint64_t code = 0;
    for (int bit = 0; bit < 21; ++bit)
      for (int d = 0; d < 3; ++d) {
        const uint64_t v = (uint64_t)(b.lo[d] - mn[d]);
        code |= ((v >> bit) & 1ull) << (3 * bit + d);
      }
    return code;
  };
  std::stable_sort(order.begin(), order.end(),
                   [&](size_t a, size_t b) { return morton(boxes[a]) < morton(boxes[b]); });
  long long total = 0;
  for (const Box& b : boxes) total += b.size();
  std::vector<int> owner(n, 0);
  long long acc = 0;
  for (size_t k = 0; k < n; ++k) {
    const size_t i = order[k];
    // rank of the box's centre of mass in the cumulative cell count
    const long long mid = acc + boxes[i].size() / 2;
    int rk = (int)((mid * nranks) / std::max(total, 1LL));
    owner[i] = std::min(std::max(rk, 0), nranks - 1);
    acc += boxes[i].size();
  }
  return owner;
}


// Load balance of a cut-cell level, cutcell::GriddingAlgorithm::LoadBalance
// (src/AMReX/cutcell/GriddingAlgorithm.cpp:190-211): every cell of a regular fab
// weighs 1, every cell of a single-valued (cut) fab `cutcell_weight`
// (PatchHierarchyOptions::cutcell_load_balance_weight, 6.0), every covered
// cell 0.001; DistributionMapping::makeKnapSack on the box weights.  The
// knapsack itself is AMReX's (unpinned): restated as its published greedy
// first step -- heaviest box first onto the lightest rank, ties to the lower
// rank -- which is what its swap refinement starts from.
// fab_type: -1 covered, 0 regular, 1 single valued; covered_cells[i]: number of
// covered cells of box i (alpha == 0).
inline std::vector<double> cutcell_box_weights(const std::vector<Box>& boxes, const int* fab_type,
                                               const long long* covered_cells,
                                               double cutcell_weight) {
  std::vector<double> w(boxes.size());
  for (size_t i = 0; i < boxes.size(); ++i) {
    const long long cells = boxes[i].size();
    const long long cov = fab_type[i] < 0 ? cells : (covered_cells ? covered_cells[i] : 0);
    const double per_cell = fab_type[i] == 1 ? cutcell_weight : 1.0;
    w[i] = per_cell * (double)(cells - cov) + 0.001 * (double)cov;
  }
  return w;
}
inline std::vector<int> knapsack(const std::vector<double>& weights, int nranks) {
  const size_t n = weights.size();
  std::vector<size_t> order(n);
  for (size_t i = 0; i < n; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(),
                   [&](size_t a, size_t b) { return weights[a] > weights[b]; });
  std::vector<double> load((size_t)nranks, 0.0);
  std::vector<int> owner(n, 0);
  for (size_t k = 0; k < n; ++k) {
    int best = 0;
    for (int r = 1; r < nranks; ++r)
      if (load[(size_t)r] < load[(size_t)best]) best = r;
    owner[order[k]] = best;
    load[(size_t)best] += weights[order[k]];
  }
  return owner;
}

// Regrid hand-off: clusters the tags of level l (dense byte map over the level's
// domain, x fastest; != 0 = tagged) into the boxes of level l + 1.  The
// reference leaves this to amrex::AmrCore::regrid (Berger-Rigoutsos,
// src/AMReX/GriddingAlgorithm.cpp:86-110 -- unpinned, absent here); this is a
// deterministic stand-in with the same contract: boxes are disjoint, cover every
// tagged cell that may be refined, are aligned to `blocking_factor` fine cells,
// at most `max_grid_size` fine cells long, and properly nested -- a block is
// refined only if all of its cells and their `nesting_buffer` neighbourhood lie
// inside the level's own boxes (or outside the domain).
// Blocks of blocking_factor / ratio coarse cells are flagged, then merged greedily
// into rectangles in x, y, z order.  Output boxes are in the index space of level
// l + 1.
inline std::vector<Box> cluster_tags(int rank, const Box& domain, const bool periodic[3],
                                     const std::vector<Box>& level_boxes, const char* tags,
                                     const int ratio[3], int blocking_factor, int max_grid_size,
                                     int nesting_buffer) {
  int bf[3], nb[3], n[3], maxb[3];
  for (int d = 0; d < 3; ++d) {
    n[d] = domain.len(d);
    bf[d] = d < rank ? std::max(blocking_factor / ratio[d], 1) : 1;
    nb[d] = (n[d] + bf[d] - 1) / bf[d];
    maxb[d] = d < rank ? std::max(max_grid_size / (bf[d] * ratio[d]), 1) : 1;
  }
  const long long ncells = (long long)n[0] * n[1] * n[2];
  // cells the next level may cover: inside this level's boxes, eroded by the buffer
  std::vector<unsigned char> covered((size_t)ncells, 0);
  for (const Box& b : level_boxes) {
    const Box c = intersect(b, domain);
    if (!c.ok()) continue;
    for (int k = c.lo[2]; k <= c.hi[2]; ++k)
      for (int j = c.lo[1]; j <= c.hi[1]; ++j)
        for (int i = c.lo[0]; i <= c.hi[0]; ++i)
          covered[(size_t)((i - domain.lo[0]) + (long long)n[0] * ((j - domain.lo[1]) + (long long)n[1] * (k - domain.lo[2])))] = 1;
  }
  auto is_covered = [&](int i, int j, int k) {
    int p[3] = {i, j, k};
    for (int d = 0; d < 3; ++d) {
      if (p[d] < 0 || p[d] >= n[d]) {
        if (d < rank && periodic[d]) {
          p[d] = (p[d] % n[d] + n[d]) % n[d];
        } else {
          return true; // beyond a physical boundary: no nesting constraint
        }
      }
    }
    return covered[(size_t)(p[0] + (long long)n[0] * (p[1] + (long long)n[1] * p[2]))] != 0;
  };
  const int nbuf[3] = {rank > 0 ? nesting_buffer : 0, rank > 1 ? nesting_buffer : 0,
                       rank > 2 ? nesting_buffer : 0};
  std::vector<unsigned char> flagged((size_t)nb[0] * nb[1] * nb[2], 0);
  for (int bk = 0; bk < nb[2]; ++bk)
    for (int bj = 0; bj < nb[1]; ++bj)
      for (int bi = 0; bi < nb[0]; ++bi) {
        const int lo[3] = {bi * bf[0], bj * bf[1], bk * bf[2]};
        const int hi[3] = {std::min(lo[0] + bf[0], n[0]) - 1, std::min(lo[1] + bf[1], n[1]) - 1,
                           std::min(lo[2] + bf[2], n[2]) - 1};
        bool any = false;
        for (int k = lo[2]; k <= hi[2] && !any; ++k)
          for (int j = lo[1]; j <= hi[1] && !any; ++j)
            for (int i = lo[0]; i <= hi[0]; ++i)
              if (tags[(size_t)(i + (long long)n[0] * (j + (long long)n[1] * k))]) {
                any = true;
                break;
              }
        if (!any) continue;
        bool nested = true;
        for (int k = lo[2] - nbuf[2]; k <= hi[2] + nbuf[2] && nested; ++k)
          for (int j = lo[1] - nbuf[1]; j <= hi[1] + nbuf[1] && nested; ++j)
            for (int i = lo[0] - nbuf[0]; i <= hi[0] + nbuf[0]; ++i)
              if (!is_covered(i, j, k)) {
                nested = false;
                break;
              }
        if (nested) flagged[(size_t)(bi + (long long)nb[0] * (bj + (long long)nb[1] * bk))] = 1;
      }
  auto F = [&](int bi, int bj, int bk) -> unsigned char& {
    return flagged[(size_t)(bi + (long long)nb[0] * (bj + (long long)nb[1] * bk))];
  };
  std::vector<Box> out;
  for (int bk = 0; bk < nb[2]; ++bk)
    for (int bj = 0; bj < nb[1]; ++bj)
      for (int bi = 0; bi < nb[0]; ++bi) {
        if (F(bi, bj, bk) != 1) continue;
        int ei = bi, ej = bj, ek = bk;
        while (ei + 1 < nb[0] && ei + 1 - bi < maxb[0] && F(ei + 1, bj, bk) == 1) ++ei;
        auto row_ok = [&](int j, int k) {
          for (int i = bi; i <= ei; ++i)
            if (F(i, j, k) != 1) return false;
          return true;
        };
        while (ej + 1 < nb[1] && ej + 1 - bj < maxb[1] && row_ok(ej + 1, bk)) ++ej;
        auto slab_ok = [&](int k) {
          for (int j = bj; j <= ej; ++j)
            if (!row_ok(j, k)) return false;
          return true;
        };
        while (ek + 1 < nb[2] && ek + 1 - bk < maxb[2] && slab_ok(ek + 1)) ++ek;
        for (int k = bk; k <= ek; ++k)
          for (int j = bj; j <= ej; ++j)
            for (int i = bi; i <= ei; ++i) F(i, j, k) = 2;
        Box b;
        const int blo[3] = {bi, bj, bk}, bhi[3] = {ei, ej, ek};
        for (int d = 0; d < 3; ++d) {
          const int clo = domain.lo[d] + blo[d] * bf[d];
          const int chi = std::min(domain.lo[d] + (bhi[d] + 1) * bf[d], domain.hi[d] + 1) - 1;
          b.lo[d] = d < rank ? clo * ratio[d] : 0;
          b.hi[d] = d < rank ? (chi + 1) * ratio[d] - 1 : 0;
        }
        out.push_back(b);
      }
  return out;
}

} // namespace fvs
