"""TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.

ctypes binding of the CPU oracle (oracle/liboracle.so).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# enums shared with include/fvs_b200.h (same numeric values)
BASE_HLLEM_LARROUTUROU, BASE_HLLEM, BASE_HLL_EINFELDT, BASE_GODUNOV_EXACT = 0, 1, 2, 3
RECON_FIRST_ORDER, RECON_MUSCL_HANCOCK, RECON_MUSCL_HANCOCK_PRIM, RECON_MUSCL_HANCOCK_CHAR = 0, 1, 2, 3
LIMITER_NONE, LIMITER_MINMOD, LIMITER_VANLEER = 0, 1, 2
# MusclHancock2<Gradient, Reconstruction, Base> (MusclHancockMethod2.hpp) and the limiters of
# CentralDifferenceGradient (Gradient.hpp:31-59), passed in place of LIMITER_*
RECON_MH2_CONSERVATIVE, RECON_MH2_PRIMITIVE, RECON_MH2_CHARACTERISTICS = 4, 5, 6
LIMITER2_NONE, LIMITER2_UPWIND, LIMITER2_MINMOD, LIMITER2_VANLEER = 0, 1, 2, 3
BC_TRANSMISSIVE, BC_REFLECTIVE = 0, 1
SPLIT_GODUNOV, SPLIT_STRANG, SPLIT_STRANG_LUMPED = 0, 1, 2
FAB_COVERED, FAB_REGULAR, FAB_SINGLEVALUED = -1, 0, 1
RIEMANN_HLL_EINFELDT, RIEMANN_EXACT = 0, 1
(EB_ALPHA, EB_BETA, EB_BCENT, EB_BNORM, EB_US, EB_SL, EB_SR, EB_DS, EB_REFERENCE,
 EB_BOUNDARY_FLUX, EB_STABILIZED, EB_SHIELDED_LEFT, EB_SHIELDED_RIGHT) = range(13)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def build_timing():
    """Timing-only builds (oracle/Makefile `timing`): {name: path} of the ones that exist."""
    try:
        subprocess.check_call(["make", "-s", "-j2", "-C", _HERE, "timing"])
    except (subprocess.CalledProcessError, OSError):
        pass
    out = {}
    for name in ("v3fma", "v4"):
        path = os.path.join(_HERE, "_native", f"liboracle_{name}.so")
        if os.path.exists(path):
            out[name] = path
    return out


def unload():
    """Forget the loaded build so that use_library() can bind another one (the old
    handle stays mapped; Level objects created from it keep working)."""
    global _LIB
    _LIB = None


def use_library(path):
    """Binds this module to another build of the oracle (bench.py: the native timing
    build).  Must be called before the first lib()."""
    global _LIB
    assert _LIB is None, "oracle library already loaded"
    _LIB = _load(path)
    return _LIB


def set_num_threads(n):
    return lib().fvso_set_num_threads(int(n))


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = _load(path)
    return _LIB


def _load(path):
    L = C.CDLL(path)
    L.fvso_last_error.restype = C.c_char_p
    L.fvso_level_create.restype = C.c_void_p
    L.fvso_level_scratch.restype = C.POINTER(C.c_double)
    L.fvso_level_fluxes.restype = C.POINTER(C.c_double)
    L.fvso_level_eb_array.restype = C.POINTER(C.c_double)
    L.fvso_cf_create.restype = C.c_void_p
    L.fvso_hierarchy_create.restype = C.c_void_p
    L.fvso_level_cycles.restype = C.c_long
    L.fvso_level_time.restype = C.c_double
    return L


def _ia(x):
    return (C.c_int * len(x))(*[int(v) for v in x])


def _da(x):
    return (C.c_double * len(x))(*[float(v) for v in x])


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _check(rc):
    if rc < 0:
        raise RuntimeError(lib().fvso_last_error().decode())
    return rc


def pad3(v, fill=0):
    v = list(v)
    return v + [fill] * (3 - len(v))


class Level:
    """One refinement level of several boxes on the CPU oracle."""

    def __init__(self, rank, domain_lo, domain_hi, dx, boxes, scratch_gcw=4,
                 flux_gcw=2, periodic=(0, 0, 0), gamma=1.4):
        self.rank = rank
        self.boxes = [(pad3(lo), pad3(hi)) for lo, hi in boxes]
        lo = [v for b in self.boxes for v in b[0]]
        hi = [v for b in self.boxes for v in b[1]]
        self.h = C.c_void_p(lib().fvso_level_create(
            rank, _ia(pad3(domain_lo)), _ia(pad3(domain_hi)), _da(pad3(dx, 1.0)),
            _ia(pad3(periodic)), len(self.boxes), _ia(lo), _ia(hi),
            scratch_gcw, flux_gcw, C.c_double(gamma)))
        self.ncomp = rank + 4
        self.ncons = rank + 2

    def __del__(self):
        try:
            lib().fvso_level_destroy(self.h)
        except Exception:
            pass

    def set_method(self, base, recon, limiter):
        lib().fvso_level_set_method(self.h, base, recon, limiter)

    def add_bc(self, direction, side, kind):
        lib().fvso_level_add_bc(self.h, direction, side, kind)

    def set_regrid_frequency(self, f):
        lib().fvso_level_set_regrid_frequency(self.h, f)

    def set_cycles(self, cycles, time_point=0.0):
        lib().fvso_level_set_cycles(self.h, C.c_long(cycles), C.c_double(time_point))

    @property
    def cycles(self):
        return lib().fvso_level_cycles(self.h)

    @property
    def time_point(self):
        return lib().fvso_level_time(self.h)

    def _view(self, ptr, lo, hi, nc):
        shape = (nc, hi[2] - lo[2] + 1, hi[1] - lo[1] + 1, hi[0] - lo[0] + 1)
        n = int(np.prod(shape))
        arr = np.ctypeslib.as_array(ptr, shape=(n,)).reshape(shape)
        return arr, list(lo), list(hi)

    def scratch(self, box):
        """numpy view [comp, z, y, x] of the scratch fab (with ghosts)."""
        lo, hi, nc = (C.c_int * 3)(), (C.c_int * 3)(), C.c_int()
        p = lib().fvso_level_scratch(self.h, box, lo, hi, C.byref(nc))
        return self._view(p, lo, hi, nc.value)

    def fluxes(self, box, direction):
        lo, hi, nc = (C.c_int * 3)(), (C.c_int * 3)(), C.c_int()
        p = lib().fvso_level_fluxes(self.h, box, direction, lo, hi, C.byref(nc))
        return self._view(p, lo, hi, nc.value)

    def fill_ghost_layer(self):
        _check(lib().fvso_fill_ghost_layer(self.h))

    def fill_boundary(self):
        _check(lib().fvso_fill_boundary(self.h))

    def apply_boundary_condition(self, direction):
        _check(lib().fvso_apply_boundary_condition(self.h, direction))

    def compute_stable_dt(self, direction):
        out = C.c_double()
        _check(lib().fvso_compute_stable_dt(self.h, direction, C.byref(out)))
        return out.value

    def compute_numeric_fluxes(self, dt, direction):
        _check(lib().fvso_compute_numeric_fluxes(self.h, C.c_double(dt), direction))

    def update_conservatively(self, dt, direction):
        _check(lib().fvso_update_conservatively(self.h, C.c_double(dt), direction))

    def complete_from_cons(self):
        _check(lib().fvso_complete_from_cons(self.h))

    def sweep(self, dt, direction):
        _check(lib().fvso_sweep(self.h, C.c_double(dt), direction))

    def advance_hierarchy(self, dt, split):
        """Returns (ok, failure_dt, sweep_dts)."""
        too = C.c_double()
        buf = (C.c_double * 16)()
        n = C.c_int(16)
        rc = _check(lib().fvso_advance_hierarchy(self.h, C.c_double(dt), split,
                                                 C.byref(too), buf, C.byref(n)))
        return rc == 0, too.value, [buf[i] for i in range(min(n.value, 16))]

    # ---- cut cells / tagging ------------------------------------------------
    def enable_eb(self, kbn_riemann=RIEMANN_HLL_EINFELDT):
        _check(lib().fvso_level_enable_eb(self.h, kbn_riemann))

    def eb_array(self, box, kind, direction=0):
        lo, hi, nc = (C.c_int * 3)(), (C.c_int * 3)(), C.c_int()
        p = lib().fvso_level_eb_array(self.h, box, kind, direction, lo, hi, C.byref(nc))
        if not p:
            raise RuntimeError("no such EB array")
        return self._view(p, lo, hi, nc.value)

    def set_eb(self, box, geom, fab_type):
        """geom: dict alpha [1,z,y,x], beta list of [1,z,y,x(+1 in d)], bcent, bnorm [rank,z,y,x]."""
        self.eb_array(box, EB_ALPHA)[0][...] = geom["alpha"]
        for d in range(self.rank):
            self.eb_array(box, EB_BETA, d)[0][...] = geom["beta"][d]
        self.eb_array(box, EB_BCENT)[0][...] = geom["bcent"]
        self.eb_array(box, EB_BNORM)[0][...] = geom["bnorm"]
        _check(lib().fvso_level_eb_finalize(self.h, box, fab_type))

    def is_cut_cell(self, box):
        arr, lo, hi = self.eb_array(box, EB_ALPHA)
        out = np.zeros(arr.shape[1:], dtype=np.uint8)
        _check(lib().fvso_level_is_cut_cell(self.h, box, out.ctypes.data_as(C.POINTER(C.c_ubyte))))
        return out

    def pre_advance_hierarchy(self):
        _check(lib().fvso_level_pre_advance(self.h))

    def _tag_shape(self, box):
        lo, hi = self.boxes[box]
        return (hi[2] - lo[2] + 1, hi[1] - lo[1] + 1, hi[0] - lo[0] + 1)

    def tag_gradient(self, box, conditions, use_alpha=False, tags=None):
        """conditions: [(component, tolerance)]; tags |= over the valid box."""
        if tags is None:
            tags = np.zeros(self._tag_shape(box), dtype=np.int8)
        comps = _ia([c for c, _ in conditions])
        tols = _da([t for _, t in conditions])
        _check(lib().fvso_tag_gradient(self.h, box, len(conditions), comps, tols,
                                       1 if use_alpha else 0, tags.ctypes.data_as(C.c_char_p)))
        return tags

    def tag_cutcells(self, box, tags=None):
        if tags is None:
            tags = np.zeros(self._tag_shape(box), dtype=np.int8)
        _check(lib().fvso_tag_cutcells(self.h, box, tags.ctypes.data_as(C.c_char_p)))
        return tags

    def run(self, cfl, final_time, max_cycles, split):
        cap = max(int(max_cycles), 1) if max_cycles >= 0 else 100000
        buf = np.zeros(cap)
        retries = C.c_int()
        n = _check(lib().fvso_run(self.h, C.c_double(cfl), C.c_double(final_time),
                                  C.c_long(max_cycles), split, _dp(buf), cap,
                                  C.byref(retries)))
        return buf[:n].copy(), retries.value


class CoarseFine:
    """Coarse-fine operations between two oracle levels (oracle/amr.hpp)."""

    def __init__(self, fine, coarse, ratio=(2, 2, 2)):
        self.fine, self.coarse = fine, coarse
        r = [ratio[d] if d < fine.rank else 1 for d in range(3)]
        self.h = C.c_void_p(lib().fvso_cf_create(fine.h, coarse.h, _ia(r)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().fvso_cf_destroy(self.h)
            self.h = None

    def reset(self):
        _check(lib().fvso_cf_reset(self.h))

    def accumulate(self, level, scale, direction):
        _check(lib().fvso_cf_accumulate(self.h, level.h, 1 if level is self.fine else 0,
                                        C.c_double(scale), direction))

    def apply_flux_correction(self):
        _check(lib().fvso_cf_apply_flux_correction(self.h, self.coarse.h))

    def coarsen_conservatively(self):
        _check(lib().fvso_cf_coarsen_conservatively(self.h, self.fine.h, self.coarse.h))

    def fill_ghost_layer_two_levels(self):
        _check(lib().fvso_cf_fill_ghost_layer_two_levels(self.h, self.fine.h, self.coarse.h))


def regrid_fill(fresh, old, coarse, ratio=(2, 2, 2)):
    """GriddingAlgorithm::RemakeLevel + CopyDataToScratch: fills the level `fresh` (already
    defined on its new boxes) from the old level (may be None) and the coarse level."""
    r = [ratio[d] if d < fresh.rank else 1 for d in range(3)]
    _check(lib().fvso_regrid_fill(fresh.h, old.h if old is not None else None, coarse.h, _ia(r)))


class Hierarchy:
    """Static multi-level hierarchy with fine-first subcycling (oracle/amr.hpp)."""

    def __init__(self, levels, ratio=(2, 2, 2)):
        self.levels = list(levels)
        r = [ratio[d] if d < levels[0].rank else 1 for d in range(3)]
        arr = (C.c_void_p * len(levels))(*[l.h for l in levels])
        self.h = C.c_void_p(lib().fvso_hierarchy_create(len(levels), arr, _ia(r)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().fvso_hierarchy_destroy(self.h)
            self.h = None

    def compute_stable_dt(self):
        out = C.c_double()
        _check(lib().fvso_hierarchy_stable_dt(self.h, C.byref(out)))
        return out.value

    def advance_hierarchy(self, dt, split):
        too = C.c_double()
        buf = (C.c_double * 64)()
        n = C.c_int(64)
        rc = _check(lib().fvso_hierarchy_advance(self.h, C.c_double(dt), split, C.byref(too), buf,
                                                 C.byref(n)))
        return rc == 0, too.value, [buf[i] for i in range(min(n.value, 64))]


def tag_buffer(tags, rank, width):
    """fub::TagBuffer on a dense [z, y, x] int8 view, in place."""
    n = _ia(list(tags.shape[::-1]))
    _check(lib().fvso_tag_buffer(tags.ctypes.data_as(C.c_char_p), rank, n, width))
    return tags


def update_conservatively(prev, flux, dt, dx):
    prev = np.ascontiguousarray(prev, dtype=np.float64)
    flux = np.ascontiguousarray(flux, dtype=np.float64)
    out = np.full_like(prev, np.nan)
    lib().fvso_unit_update_conservatively(_dp(out), _dp(prev), _dp(flux),
                                          len(prev), C.c_double(dt), C.c_double(dx))
    return out


def complete_from_cons(rank, cons, gamma=1.4):
    cons = np.ascontiguousarray(cons, dtype=np.float64)
    out = np.zeros(rank + 4)
    lib().fvso_unit_complete_from_cons(rank, C.c_double(gamma), _dp(cons), _dp(out))
    return out


def numeric_flux(rank, base, recon, limiter, states, dt, dx, direction, gamma=1.4):
    states = np.ascontiguousarray(states, dtype=np.float64)
    out = np.zeros(rank + 2)
    _check(lib().fvso_unit_numeric_flux(rank, C.c_double(gamma), base, recon, limiter,
                                        _dp(states), C.c_double(dt), C.c_double(dx),
                                        direction, _dp(out)))
    return out


def numeric_flux_scalar_hll(rank, recon, limiter, states, dt, dx, direction, gamma=1.4):
    states = np.ascontiguousarray(states, dtype=np.float64)
    out = np.zeros(rank + 2)
    _check(lib().fvso_unit_numeric_flux_scalar_hll(rank, C.c_double(gamma), recon,
                                                   limiter, _dp(states), C.c_double(dt),
                                                   C.c_double(dx), direction, _dp(out)))
    return out


def hllem_scalar(rank, larrouturou, states, direction, gamma=1.4):
    """The reference's scalar HLLEM code path (scalar_hllem.hpp), Rank 1 and 2."""
    states = np.ascontiguousarray(states, dtype=np.float64)
    out = np.zeros(rank + 2)
    _check(lib().fvso_unit_hllem_scalar(rank, C.c_double(gamma), int(larrouturou), _dp(states),
                                        direction, _dp(out)))
    return out


def exact_riemann(rank, left, right, direction=0, gamma=1.4):
    left = np.ascontiguousarray(left, dtype=np.float64)
    right = np.ascontiguousarray(right, dtype=np.float64)
    pu = np.zeros(2)
    sol = np.zeros(rank + 4)
    _check(lib().fvso_unit_exact_riemann(rank, C.c_double(gamma), _dp(left), _dp(right),
                                         direction, _dp(pu), _dp(sol)))
    return pu, sol


def tiles(rank, lo, hi, tile_size, ngrow, nodal_dir=-1):
    cap = 4096
    buf = (C.c_int * (12 * cap))()
    n = lib().fvso_unit_tiles(rank, _ia(pad3(lo)), _ia(pad3(hi)), _ia(pad3(tile_size, 1)),
                              _ia(pad3(ngrow)), nodal_dir, buf, cap)
    out = []
    for t in range(n):
        v = [buf[12 * t + i] for i in range(12)]
        out.append((v[0:3], v[3:6], v[6:9], v[9:12]))
    return out
