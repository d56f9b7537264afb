"""ctypes binding of libfvs_b200.so (include/fvs_b200.h).

This is harness plumbing for tests/ and bench.py: every call goes straight
through the C ABI, the same entry points the C++ policies in include/fub_b200/
use.  There is no CPU fallback: loading fails loudly if the library has not
been built, and creating a context fails without a CUDA device.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# FVS_B200_LIB: tuning builds (tools/build_variant.py) only; the product library is libfvs_b200.so
LIB_PATH = os.environ.get("FVS_B200_LIB") or os.path.join(_HERE, "libfvs_b200.so")

BASE_HLLEM_LARROUTUROU, BASE_HLLEM, BASE_HLL_EINFELDT, BASE_GODUNOV_EXACT = 0, 1, 2, 3
RECON_FIRST_ORDER, RECON_MUSCL_HANCOCK, RECON_MUSCL_HANCOCK_PRIM, RECON_MUSCL_HANCOCK_CHAR = 0, 1, 2, 3
LIMITER_NONE, LIMITER_MINMOD, LIMITER_VANLEER = 0, 1, 2
# MusclHancock2<Gradient, Reconstruction, Base> (MusclHancockMethod2.hpp) and the limiters of
# CentralDifferenceGradient (Gradient.hpp:31-59), passed in place of LIMITER_*
RECON_MH2_CONSERVATIVE, RECON_MH2_PRIMITIVE, RECON_MH2_CHARACTERISTICS = 4, 5, 6
LIMITER2_NONE, LIMITER2_UPWIND, LIMITER2_MINMOD, LIMITER2_VANLEER = 0, 1, 2, 3
BC_TRANSMISSIVE, BC_REFLECTIVE = 0, 1
SPLIT_GODUNOV, SPLIT_STRANG, SPLIT_STRANG_LUMPED = 0, 1, 2
OK, ERROR, TIME_STEP_TOO_LARGE = 0, -1, 1
FAB_COVERED, FAB_REGULAR, FAB_SINGLEVALUED = -1, 0, 1
RIEMANN_HLL_EINFELDT, RIEMANN_EXACT = 0, 1
ARITH_IEEE, ARITH_TOLERANCE = 0, 1
(EB_ALPHA, EB_BETA, EB_BCENT, EB_BNORM, EB_US, EB_SL, EB_SR, EB_DS, EB_REFERENCE,
 EB_BOUNDARY_FLUX, EB_STABILIZED, EB_SHIELDED_LEFT, EB_SHIELDED_RIGHT) = range(13)
UNIQUE_ID_BYTES = 128

# every symbol include/fvs_b200.h declares
SYMBOLS = [
    "fvs_last_error", "fvs_version", "fvs_device_count", "fvs_create", "fvs_destroy",
    "fvs_set_method", "fvs_set_arithmetic", "fvs_stencil_width", "fvs_level_define", "fvs_level_add_boundary",
    "fvs_state_box", "fvs_flux_box", "fvs_upload_state", "fvs_download_state",
    "fvs_upload_fluxes", "fvs_download_fluxes", "fvs_state_device_ptr",
    "fvs_compute_stable_dt", "fvs_compute_numeric_fluxes", "fvs_update_conservatively",
    "fvs_complete_from_cons", "fvs_sweep", "fvs_fill_ghost_layer",
    "fvs_apply_boundary_condition", "fvs_compute_stable_dt_all", "fvs_advance_hierarchy",
    "fvs_run", "fvs_set_regrid_frequency", "fvs_set_cycles", "fvs_cycles", "fvs_time_point",
    "fvs_min_all", "fvs_comm_unique_id", "fvs_comm_init", "fvs_halo_plan",
    "fvs_stable_dt_face_mask", "fvs_distribute_boxes", "fvs_kernel_launches", "fvs_stream",
    "fvs_distribute_boxes_cutcell", "fvs_cluster_tags", "fvs_regrid_level", "fvs_tag_level",
    "fvs_write_checkpoint",
    "fvs_set_dt_hints", "fvs_dt_hint_redos", "fvs_download_state_region",
    "fvs_enable_timing", "fvs_sweep_time_ms",
    "fvs_level_set_eb", "fvs_set_cutcell_riemann_solver", "fvs_pre_advance_hierarchy",
    "fvs_eb_array_box", "fvs_download_eb_array", "fvs_cut_cell_flags",
    "fvs_tag_gradient", "fvs_tag_cutcells", "fvs_tag_buffer",
    "fvs_set_refine_ratio", "fvs_reset_coarse_fine_fluxes", "fvs_accumulate_coarse_fine_fluxes",
    "fvs_apply_flux_correction", "fvs_coarsen_conservatively", "fvs_fill_ghost_layer_two_levels",
    "fvs_compute_stable_dt_hierarchy",
    "fvs_upload_state_async", "fvs_download_state_async", "fvs_synchronize",
    "fvs_copy_data_to_scratch_async", "fvs_copy_scratch_to_data_async",
]


class Config(C.Structure):
    _fields_ = [("rank", C.c_int), ("gamma", C.c_double), ("base_method", C.c_int),
                ("reconstruction", C.c_int), ("limiter", C.c_int), ("device", C.c_int),
                ("comm_rank", C.c_int), ("comm_size", C.c_int)]


class HaloCopy(C.Structure):
    _fields_ = [("dst_box", C.c_int), ("src_box", C.c_int), ("dst_owner", C.c_int),
                ("src_owner", C.c_int), ("lo", C.c_int * 3), ("hi", C.c_int * 3),
                ("shift", C.c_int * 3)]


_LIB = None


def lib():
    """Loads libfvs_b200.so; raises if it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m finitevolumesolver_b200.build` "
                "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.fvs_last_error.restype = C.c_char_p
        L.fvs_cycles.restype = C.c_long
        L.fvs_time_point.restype = C.c_double
        L.fvs_run.restype = C.c_long
        L.fvs_halo_plan.restype = C.c_long
        L.fvs_cluster_tags.restype = C.c_long
        L.fvs_kernel_launches.restype = C.c_long
        L.fvs_dt_hint_redos.restype = C.c_long
        L.fvs_cut_cell_flags.restype = C.c_long
        L.fvs_stream.restype = C.c_void_p
        _LIB = L
    return _LIB


class FvsError(RuntimeError):
    pass


def _check(rc):
    if rc < 0:
        raise FvsError(lib().fvs_last_error().decode())
    return rc


def _i3(v, fill=0):
    v = list(v) + [fill] * (3 - len(v))
    return (C.c_int * 3)(*[int(x) for x in v])


def _d3(v, fill=1.0):
    v = list(v) + [fill] * (3 - len(v))
    return (C.c_double * 3)(*[float(x) for x in v])


def _boxes(boxes):
    n = len(boxes)
    lo = ((C.c_int * 3) * n)()
    hi = ((C.c_int * 3) * n)()
    for b, (l, h) in enumerate(boxes):
        l = list(l) + [0] * (3 - len(l))
        h = list(h) + [0] * (3 - len(h))
        for d in range(3):
            lo[b][d] = int(l[d])
            hi[b][d] = int(h[d])
    return lo, hi


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def device_count():
    return lib().fvs_device_count()


def halo_plan(rank, domain_lo, domain_hi, periodic, boxes, owner, scratch_gcw):
    lo, hi = _boxes(boxes)
    own = (C.c_int * len(boxes))(*[int(o) for o in owner])
    n = _check(lib().fvs_halo_plan(rank, _i3(domain_lo), _i3(domain_hi), _i3(periodic),
                                   len(boxes), lo, hi, own, scratch_gcw, None, C.c_long(0)))
    buf = (HaloCopy * max(n, 1))()
    _check(lib().fvs_halo_plan(rank, _i3(domain_lo), _i3(domain_hi), _i3(periodic),
                               len(boxes), lo, hi, own, scratch_gcw, buf, C.c_long(n)))
    return [dict(dst_box=c.dst_box, src_box=c.src_box, dst_owner=c.dst_owner,
                 src_owner=c.src_owner, lo=list(c.lo), hi=list(c.hi), shift=list(c.shift))
            for c in buf[:n]]


def stable_dt_face_mask(rank, box_lo, box_hi, scratch_gcw, stencil_width, direction):
    n = _check(lib().fvs_stable_dt_face_mask(rank, _i3(box_lo), _i3(box_hi), scratch_gcw,
                                             stencil_width, direction, None, 0))
    buf = (C.c_ubyte * n)()
    _check(lib().fvs_stable_dt_face_mask(rank, _i3(box_lo), _i3(box_hi), scratch_gcw,
                                         stencil_width, direction, buf, n))
    return np.frombuffer(buf, dtype=np.uint8).copy()


def distribute_boxes(rank, boxes, nranks):
    lo, hi = _boxes(boxes)
    out = (C.c_int * len(boxes))()
    _check(lib().fvs_distribute_boxes(rank, len(boxes), lo, hi, nranks, out))
    return list(out)


def distribute_boxes_cutcell(rank, boxes, fab_types, covered_cells, nranks, cutcell_weight=6.0):
    """cutcell::GriddingAlgorithm::LoadBalance: knapsack over cut-cell weighted boxes."""
    lo, hi = _boxes(boxes)
    ft = (C.c_int * len(boxes))(*[int(t) for t in fab_types])
    cov = (C.c_longlong * len(boxes))(*[int(c) for c in covered_cells])
    out = (C.c_int * len(boxes))()
    _check(lib().fvs_distribute_boxes_cutcell(rank, len(boxes), lo, hi, ft, cov,
                                              C.c_double(cutcell_weight), nranks, out))
    return list(out)


def cluster_tags(rank, domain_lo, domain_hi, periodic, boxes, tags, ratio=(2, 2, 2),
                 blocking_factor=8, max_grid_size=128, nesting_buffer=1):
    """tags: int8 array [z, y, x] over the level's domain.  Returns the boxes of the next
    finer level as [(lo, hi), ...]."""
    lo, hi = _boxes(boxes)
    tags = np.ascontiguousarray(tags, dtype=np.int8)
    r = [ratio[d] if d < rank else 1 for d in range(3)]
    args = (rank, _i3(domain_lo), _i3(domain_hi), _i3(periodic), len(boxes), lo, hi,
            tags.ctypes.data_as(C.c_char_p), _i3(r, 1), blocking_factor, max_grid_size, nesting_buffer)
    n = lib().fvs_cluster_tags(*args, None, None, C.c_long(0))
    if n < 0:
        raise RuntimeError(lib().fvs_last_error().decode())
    out_lo = ((C.c_int * 3) * max(n, 1))()
    out_hi = ((C.c_int * 3) * max(n, 1))()
    lib().fvs_cluster_tags(*args, out_lo, out_hi, C.c_long(n))
    return [(list(out_lo[i]), list(out_hi[i])) for i in range(n)]


def comm_unique_id():
    buf = (C.c_char * UNIQUE_ID_BYTES)()
    _check(lib().fvs_comm_unique_id(buf))
    return bytes(buf)


class Context:
    """One fvs_ctx with a single level (level 0)."""

    def __init__(self, rank, domain_lo, domain_hi, dx, boxes, scratch_gcw=4, flux_gcw=2,
                 periodic=(0, 0, 0), gamma=1.4, base=BASE_HLLEM_LARROUTUROU,
                 recon=RECON_MUSCL_HANCOCK, limiter=LIMITER_VANLEER, device=-1,
                 comm_rank=0, comm_size=1, owner=None):
        self.rank = rank
        self.ncomp = rank + 4
        self.ncons = rank + 2
        self.boxes = [(list(lo) + [0] * (3 - len(lo)), list(hi) + [0] * (3 - len(hi)))
                      for lo, hi in boxes]
        self.owner = list(owner) if owner is not None else [0] * len(boxes)
        self.comm_rank = comm_rank
        self.level = 0
        self._owns = True
        cfg = Config(rank, gamma, base, recon, limiter, device, comm_rank, comm_size)
        self.h = C.c_void_p()
        _check(lib().fvs_create(C.byref(cfg), C.byref(self.h)))
        lo, hi = _boxes(self.boxes)
        own = (C.c_int * len(boxes))(*self.owner)
        _check(lib().fvs_level_define(self.h, self.level, _i3(domain_lo), _i3(domain_hi), _d3(dx),
                                      _i3(periodic), len(boxes), lo, hi, own, scratch_gcw,
                                      flux_gcw))

    def close(self):
        if self.h and self._owns:
            lib().fvs_destroy(self.h)
        self.h = C.c_void_p()

    def define_level(self, level, domain_lo, domain_hi, dx, boxes, scratch_gcw=4, flux_gcw=2,
                     periodic=(0, 0, 0), owner=None):
        """Defines a further (finer) level and returns a view of the context
        whose methods address that level."""
        import copy
        v = copy.copy(self)
        v._owns = False
        v.level = level
        v.boxes = [(list(lo) + [0] * (3 - len(lo)), list(hi) + [0] * (3 - len(hi)))
                   for lo, hi in boxes]
        v.owner = list(owner) if owner is not None else [self.comm_rank] * len(boxes)
        lo, hi = _boxes(v.boxes)
        own = (C.c_int * len(boxes))(*v.owner)
        _check(lib().fvs_level_define(self.h, level, _i3(domain_lo), _i3(domain_hi), _d3(dx),
                                      _i3(periodic), len(boxes), lo, hi, own, scratch_gcw,
                                      flux_gcw))
        return v

    def write_checkpoint(self, directory):
        _check(lib().fvs_write_checkpoint(self.h, str(directory).encode()))

    def regrid_level(self, level, boxes, owner=None):
        """fvs_regrid_level: new boxes for `level` (>= 1), data from the old level where it
        exists and from the coarse parent elsewhere.  Returns a view addressing the level."""
        import copy
        v = copy.copy(self)
        v._owns = False
        v.level = level
        v.boxes = [(list(lo) + [0] * (3 - len(lo)), list(hi) + [0] * (3 - len(hi)))
                   for lo, hi in boxes]
        v.owner = list(owner) if owner is not None else [self.comm_rank] * len(boxes)
        lo, hi = _boxes(v.boxes)
        own = (C.c_int * len(boxes))(*v.owner)
        _check(lib().fvs_regrid_level(self.h, level, len(boxes), lo, hi, own))
        return v

    # ---- coarse-fine (two levels) ---------------------------------------------
    def set_refine_ratio(self, ratio):
        _check(lib().fvs_set_refine_ratio(self.h, _i3(ratio, 1)))

    def reset_coarse_fine_fluxes(self, fine, coarse):
        _check(lib().fvs_reset_coarse_fine_fluxes(self.h, fine, coarse))

    def accumulate_coarse_fine_fluxes(self, level, scale, direction):
        _check(lib().fvs_accumulate_coarse_fine_fluxes(self.h, level, C.c_double(scale), direction))

    def apply_flux_correction(self, fine, coarse, dt=0.0):
        _check(lib().fvs_apply_flux_correction(self.h, fine, coarse, C.c_double(dt)))

    def coarsen_conservatively(self, fine, coarse):
        _check(lib().fvs_coarsen_conservatively(self.h, fine, coarse))

    def fill_ghost_layer_two_levels(self, fine, coarse):
        _check(lib().fvs_fill_ghost_layer_two_levels(self.h, fine, coarse))

    def compute_stable_dt_hierarchy(self):
        out = C.c_double()
        _check(lib().fvs_compute_stable_dt_hierarchy(self.h, C.byref(out)))
        return out.value

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def owned_boxes(self):
        return [b for b, o in enumerate(self.owner) if o == self.comm_rank]

    def comm_init(self, unique_id):
        _check(lib().fvs_comm_init(self.h, unique_id))

    def set_method(self, base, recon, limiter):
        _check(lib().fvs_set_method(self.h, base, recon, limiter))

    def set_arithmetic(self, arithmetic):
        _check(lib().fvs_set_arithmetic(self.h, arithmetic))

    def add_bc(self, direction, side, kind):
        _check(lib().fvs_level_add_boundary(self.h, self.level, direction, side, kind))

    def set_regrid_frequency(self, f):
        _check(lib().fvs_set_regrid_frequency(self.h, f))

    # ---- cut cells / tagging ------------------------------------------------
    def enable_eb(self, kbn_riemann=RIEMANN_HLL_EINFELDT):
        _check(lib().fvs_set_cutcell_riemann_solver(self.h, kbn_riemann))

    def set_eb(self, box, geom, fab_type):
        alpha = np.ascontiguousarray(geom["alpha"], dtype=np.float64)
        betas = [np.ascontiguousarray(b, dtype=np.float64) for b in geom["beta"]]
        bcent = np.ascontiguousarray(geom["bcent"], dtype=np.float64)
        bnorm = np.ascontiguousarray(geom["bnorm"], dtype=np.float64)
        ptrs = (C.POINTER(C.c_double) * 3)(*[_dp(b) for b in betas])
        _check(lib().fvs_level_set_eb(self.h, self.level, box, _dp(alpha), ptrs, _dp(bcent), _dp(bnorm),
                                      fab_type))

    def pre_advance_hierarchy(self):
        _check(lib().fvs_pre_advance_hierarchy(self.h))

    def eb_array(self, box, kind, direction=0):
        lo, hi, nc = (C.c_int * 3)(), (C.c_int * 3)(), C.c_int()
        _check(lib().fvs_eb_array_box(self.h, self.level, box, kind, direction, lo, hi, C.byref(nc)))
        out = np.empty(self._shape(list(lo), list(hi), nc.value))
        _check(lib().fvs_download_eb_array(self.h, self.level, box, kind, direction, _dp(out)))
        return out

    def is_cut_cell(self, box):
        out = np.zeros(self.state_shape(box)[1:], dtype=np.uint8)
        n = lib().fvs_cut_cell_flags(self.h, self.level, box, out.ctypes.data_as(C.POINTER(C.c_ubyte)))
        if n < 0:
            raise FvsError(lib().fvs_last_error().decode())
        return out, n

    def _tag_shape(self, box):
        lo, hi = self.boxes[box]
        return (hi[2] - lo[2] + 1, hi[1] - lo[1] + 1, hi[0] - lo[0] + 1)

    def tag_gradient(self, box, conditions, use_alpha=False, tags=None):
        if tags is None:
            tags = np.zeros(self._tag_shape(box), dtype=np.int8)
        comps = (C.c_int * len(conditions))(*[c for c, _ in conditions])
        tols = (C.c_double * len(conditions))(*[t for _, t in conditions])
        _check(lib().fvs_tag_gradient(self.h, self.level, box, len(conditions), comps, tols,
                                      1 if use_alpha else 0, tags.ctypes.data_as(C.c_char_p)))
        return tags

    def tag_level(self, domain_n, conditions=(), use_alpha=False, cutcells=False, buffer_width=0,
                  tags=None):
        """Dense [z, y, x] int8 tag map of this level (fvs_tag_level)."""
        if tags is None:
            tags = np.zeros((domain_n[2], domain_n[1], domain_n[0]), dtype=np.int8)
        comps = (C.c_int * max(len(conditions), 1))(*[c for c, _ in conditions])
        tols = (C.c_double * max(len(conditions), 1))(*[t for _, t in conditions])
        _check(lib().fvs_tag_level(self.h, self.level, len(conditions), comps, tols,
                                   1 if use_alpha else 0, 1 if cutcells else 0, buffer_width,
                                   tags.ctypes.data_as(C.c_char_p)))
        return tags

    def tag_cutcells(self, box, tags=None):
        if tags is None:
            tags = np.zeros(self._tag_shape(box), dtype=np.int8)
        _check(lib().fvs_tag_cutcells(self.h, self.level, box, tags.ctypes.data_as(C.c_char_p)))
        return tags

    def tag_buffer(self, tags, rank, width):
        n = (C.c_int * 3)(*(list(tags.shape[::-1]) + [1, 1])[:3])
        _check(lib().fvs_tag_buffer(self.h, rank, tags.ctypes.data_as(C.c_char_p), n, width))
        return tags

    def set_cycles(self, cycles, time_point=0.0):
        _check(lib().fvs_set_cycles(self.h, C.c_long(cycles), C.c_double(time_point)))

    @property
    def cycles(self):
        return lib().fvs_cycles(self.h)

    @property
    def time_point(self):
        return lib().fvs_time_point(self.h)

    def state_box(self, box):
        lo, hi, nc = (C.c_int * 3)(), (C.c_int * 3)(), C.c_int()
        _check(lib().fvs_state_box(self.h, self.level, box, lo, hi, C.byref(nc)))
        return list(lo), list(hi), nc.value

    def flux_box(self, box, direction):
        lo, hi, nc = (C.c_int * 3)(), (C.c_int * 3)(), C.c_int()
        _check(lib().fvs_flux_box(self.h, self.level, box, direction, lo, hi, C.byref(nc)))
        return list(lo), list(hi), nc.value

    @staticmethod
    def _shape(lo, hi, nc):
        return (nc, hi[2] - lo[2] + 1, hi[1] - lo[1] + 1, hi[0] - lo[0] + 1)

    def state_shape(self, box):
        return self._shape(*self.state_box(box))

    def upload_state(self, box, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        assert arr.shape == self.state_shape(box), (arr.shape, self.state_shape(box))
        _check(lib().fvs_upload_state(self.h, self.level, box, _dp(arr)))

    def download_state(self, box, out=None):
        if out is None:
            out = np.empty(self.state_shape(box))
        _check(lib().fvs_download_state(self.h, self.level, box, _dp(out)))
        return out

    def download_state_region(self, box, lo, hi):
        """[ncomp, z, y, x] copy of the cells lo..hi (global indices) of the box's state array."""
        lo3, hi3 = list(lo) + [0] * (3 - len(lo)), list(hi) + [0] * (3 - len(hi))
        out = np.empty((self.ncomp, hi3[2] - lo3[2] + 1, hi3[1] - lo3[1] + 1, hi3[0] - lo3[0] + 1))
        _check(lib().fvs_download_state_region(self.h, self.level, box, _i3(lo3), _i3(hi3), _dp(out)))
        return out

    def upload_state_async(self, box, arr):
        """arr must be a C-contiguous float64 array in page-locked memory."""
        assert arr.dtype == np.float64 and arr.flags["C_CONTIGUOUS"]
        assert arr.shape == self.state_shape(box), (arr.shape, self.state_shape(box))
        _check(lib().fvs_upload_state_async(self.h, self.level, box, _dp(arr)))

    def download_state_async(self, box, out):
        assert out.dtype == np.float64 and out.flags["C_CONTIGUOUS"]
        _check(lib().fvs_download_state_async(self.h, self.level, box, _dp(out)))

    def data_shape(self, box):
        """[ncomp, z, y, x] of the level data of a box: valid cells, no ghosts."""
        lo, hi = self.boxes[box]
        return (self.ncomp, hi[2] - lo[2] + 1, hi[1] - lo[1] + 1, hi[0] - lo[0] + 1)

    def copy_data_to_scratch_async(self, box, arr):
        assert arr.dtype == np.float64 and arr.flags["C_CONTIGUOUS"]
        assert arr.shape == self.data_shape(box), (arr.shape, self.data_shape(box))
        _check(lib().fvs_copy_data_to_scratch_async(self.h, self.level, box, _dp(arr)))

    def copy_scratch_to_data_async(self, box, out):
        assert out.dtype == np.float64 and out.flags["C_CONTIGUOUS"]
        assert out.shape == self.data_shape(box), (out.shape, self.data_shape(box))
        _check(lib().fvs_copy_scratch_to_data_async(self.h, self.level, box, _dp(out)))

    def synchronize(self):
        _check(lib().fvs_synchronize(self.h))

    def download_fluxes(self, box, direction):
        out = np.empty(self._shape(*self.flux_box(box, direction)))
        _check(lib().fvs_download_fluxes(self.h, self.level, box, direction, _dp(out)))
        return out

    def upload_fluxes(self, box, direction, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        assert arr.shape == self._shape(*self.flux_box(box, direction))
        _check(lib().fvs_upload_fluxes(self.h, self.level, box, direction, _dp(arr)))

    def compute_stable_dt(self, direction):
        out = C.c_double()
        _check(lib().fvs_compute_stable_dt(self.h, self.level, direction, C.byref(out)))
        return out.value

    def compute_stable_dt_all(self):
        out = C.c_double()
        _check(lib().fvs_compute_stable_dt_all(self.h, self.level, C.byref(out)))
        return out.value

    def compute_numeric_fluxes(self, dt, direction):
        _check(lib().fvs_compute_numeric_fluxes(self.h, self.level, C.c_double(dt), direction))

    def update_conservatively(self, dt, direction):
        _check(lib().fvs_update_conservatively(self.h, self.level, C.c_double(dt), direction))

    def complete_from_cons(self):
        _check(lib().fvs_complete_from_cons(self.h, self.level))

    def sweep(self, dt, direction):
        """Fused sweep. Returns (status, level_dt)."""
        out = C.c_double()
        rc = _check(lib().fvs_sweep(self.h, self.level, C.c_double(dt), direction, C.byref(out)))
        return rc, out.value

    def fill_ghost_layer(self):
        _check(lib().fvs_fill_ghost_layer(self.h, self.level))

    def apply_boundary_condition(self, direction):
        _check(lib().fvs_apply_boundary_condition(self.h, self.level, direction))

    def advance_hierarchy(self, dt, split):
        """Returns (ok, failure_dt, sweep_dts)."""
        too = C.c_double()
        buf = (C.c_double * 64)()
        n = C.c_int(64)
        rc = _check(lib().fvs_advance_hierarchy(self.h, C.c_double(dt), split, C.byref(too),
                                                buf, C.byref(n)))
        return rc == 0, too.value, [buf[i] for i in range(min(n.value, 64))]

    def run(self, cfl, final_time, max_cycles, split):
        cap = max(int(max_cycles), 1) if max_cycles >= 0 else 100000
        buf = np.zeros(cap)
        retries = C.c_int()
        n = lib().fvs_run(self.h, C.c_double(cfl), C.c_double(final_time), C.c_long(max_cycles),
                          split, _dp(buf), C.c_long(cap), C.byref(retries))
        _check(n)
        return buf[:n].copy(), retries.value

    def min_all(self, value):
        v = C.c_double(value)
        _check(lib().fvs_min_all(self.h, C.byref(v)))
        return v.value

    @property
    def kernel_launches(self):
        return lib().fvs_kernel_launches(self.h)

    def set_dt_hints(self, on):
        _check(lib().fvs_set_dt_hints(self.h, int(on)))

    @property
    def dt_hint_redos(self):
        return lib().fvs_dt_hint_redos(self.h)

    @property
    def stream(self):
        return lib().fvs_stream(self.h)

    def enable_timing(self, on=True):
        _check(lib().fvs_enable_timing(self.h, 1 if on else 0))

    def sweep_time_ms(self):
        ms, n = C.c_double(), C.c_long()
        _check(lib().fvs_sweep_time_ms(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value
