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

} // namespace fvs
