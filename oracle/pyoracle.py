"""ctypes bindings for the CPU oracles.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl
reference` legs may import this module; the product package (gloc3d_b200)
never does.  See oracle/gloc_oracle.h for what each function restates
(reference file:line) and for the parity status of each stage.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libgloc_oracle.so")
_REF = os.path.join(_HERE, "_ref", "libnanoflann_ref.so")

_u64p = C.POINTER(C.c_uint64)
_f32p = C.POINTER(C.c_float)


def build(force: bool = False) -> None:
    """Compile the C restatement (always) and oracle/_ref (when /root/reference exists)."""
    if force or not os.path.exists(_LIB) or not os.path.exists(_REF):
        subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)


def _ptr(a: np.ndarray, ty):
    return a.ctypes.data_as(ty)


class MatchResult(C.Structure):
    _fields_ = [
        ("found", C.c_int),
        ("score", C.c_float),
        ("scan_index", C.c_int),
        ("x_offset", C.c_int),
        ("y_offset", C.c_int),
        ("pose_x", C.c_double),
        ("pose_y", C.c_double),
        ("pose_yaw", C.c_double),
        ("n_scored", C.c_longlong),
    ]

    def as_tuple(self):
        return (self.found, self.score, self.scan_index, self.x_offset, self.y_offset,
                self.pose_x, self.pose_y, self.pose_yaw)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.gloc_oracle_l2.restype = C.c_float
        L.gloc_oracle_l2.argtypes = [_f32p, _f32p, C.c_size_t]
        L.gloc_oracle_knn.argtypes = [_f32p, C.c_size_t, C.c_size_t, _f32p, C.c_size_t,
                                      C.c_size_t, _u64p, _f32p]
        L.gloc_oracle_knn_mt.argtypes = L.gloc_oracle_knn.argtypes + [C.c_int]
        L.gloc_oracle_topk_merge.argtypes = [_u64p, _f32p, C.c_size_t, C.c_size_t,
                                             C.c_size_t, _u64p, _f32p]
        L.gloc_oracle_min_cost.restype = C.c_float
        L.gloc_oracle_max_cost.restype = C.c_float
        L.gloc_oracle_value_to_cost.restype = C.c_float
        L.gloc_oracle_value_to_cost.argtypes = [C.c_uint16]
        L.gloc_oracle_cost_to_value.restype = C.c_uint16
        L.gloc_oracle_cost_to_value.argtypes = [C.c_float]
        L.gloc_oracle_cell_value.restype = C.c_uint8
        L.gloc_oracle_cell_value.argtypes = [C.c_float]
        L.gloc_oracle_level1_from_cells.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.gloc_oracle_precomp_from_cells.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int,
                                                     C.c_void_p]
        L.gloc_oracle_precomp_from_level1.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int,
                                                      C.c_void_p]
        L.gloc_oracle_search_params.argtypes = [C.c_double, C.c_double, _f32p, C.c_int,
                                                C.c_double, C.POINTER(C.c_int),
                                                C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.gloc_oracle_grid_to_points.restype = C.c_int
        L.gloc_oracle_grid_to_points.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double,
                                                 C.c_double, C.c_double, C.c_void_p, C.c_int]
        L.gloc_oracle_discretize.argtypes = [_f32p, C.c_int, C.c_double, C.c_double,
                                             C.c_double, C.c_int, C.c_double, C.c_double,
                                             C.c_double, C.c_double, C.c_void_p]
        L.gloc_oracle_csm_match.restype = C.c_int
        L.gloc_oracle_csm_match.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double,
                                            C.c_double, C.c_double, C.c_int, _f32p, C.c_int,
                                            C.c_double, C.c_double, C.c_double, C.c_int,
                                            C.c_int, C.c_double, C.c_float, C.c_int,
                                            C.POINTER(MatchResult)]
        L.gloc_oracle_csm_match_full_submap.restype = C.c_int
        L.gloc_oracle_csm_match_full_submap.argtypes = [C.c_void_p, C.c_int, C.c_int,
                                                        C.c_double, C.c_double, C.c_double,
                                                        C.c_int, _f32p, C.c_int, C.c_float,
                                                        C.c_int, C.POINTER(MatchResult)]
        L.gloc_oracle_csm_match_batch_mt.argtypes = [
            C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
            C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_int,
            C.c_int, C.c_int, C.c_double, C.c_float, C.c_int, C.c_int, C.POINTER(MatchResult)]
        L.gloc_oracle_bev_project.restype = C.c_int
        L.gloc_oracle_bev_project.argtypes = [
            C.c_void_p, C.c_size_t, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_size_t,
            C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
            C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_size_t),
            C.POINTER(C.c_size_t)]
        L.gloc_oracle_crop_pad.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def have_ref() -> bool:
    if not os.path.exists(_REF):
        try:
            build()
        except Exception:
            return False
    return os.path.exists(_REF)


def ref():
    """The reference's own nanoflann, compiled from /root/reference (oracle/_ref)."""
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libnanoflann_ref.so is missing")
        R = C.CDLL(_REF)
        R.gloc_ref_knn_build.restype = C.c_void_p
        R.gloc_ref_knn_build.argtypes = [_f32p, C.c_size_t, C.c_size_t, C.c_int]
        R.gloc_ref_knn_query.argtypes = [C.c_void_p, _f32p, C.c_size_t, _u64p, _f32p]
        R.gloc_ref_knn_query_batch.argtypes = [C.c_void_p, _f32p, C.c_size_t, C.c_size_t,
                                               _u64p, _f32p, C.c_int]
        R.gloc_ref_knn_free.argtypes = [C.c_void_p]
        _ref = R
    return _ref


# ----------------------------------------------------------------- stage 1

def l2(q: np.ndarray, x: np.ndarray) -> np.float32:
    q = np.ascontiguousarray(q, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    return np.float32(lib().gloc_oracle_l2(_ptr(q, _f32p), _ptr(x, _f32p), q.size))


def knn(db: np.ndarray, q: np.ndarray, k: int, nthreads: int = 1):
    db = np.ascontiguousarray(db, np.float32)
    q = np.ascontiguousarray(q, np.float32)
    n, dim = db.shape if db.ndim == 2 else (0, q.shape[1])
    nq = q.shape[0]
    idx = np.empty((nq, k), np.uint64)
    d2 = np.empty((nq, k), np.float32)
    lib().gloc_oracle_knn_mt(_ptr(db, _f32p), n, dim, _ptr(q, _f32p), nq, k,
                             _ptr(idx, _u64p), _ptr(d2, _f32p), nthreads)
    return idx, d2


def topk_merge(idx: np.ndarray, d2: np.ndarray):
    """idx/d2: [g, nq, k] per-shard lists with GLOBAL indices."""
    idx = np.ascontiguousarray(idx, np.uint64)
    d2 = np.ascontiguousarray(d2, np.float32)
    g, nq, k = idx.shape
    oi = np.empty((nq, k), np.uint64)
    od = np.empty((nq, k), np.float32)
    lib().gloc_oracle_topk_merge(_ptr(idx, _u64p), _ptr(d2, _f32p), g, nq, k,
                                 _ptr(oi, _u64p), _ptr(od, _f32p))
    return oi, od


class RefTree:
    """The reference's InvKeyTree (nanoflann KD-tree, leaf 10) over a copy of db."""

    def __init__(self, db: np.ndarray, leaf_max_size: int = 10):
        db = np.ascontiguousarray(db, np.float32)
        self.n, self.dim = db.shape
        self._h = ref().gloc_ref_knn_build(_ptr(db, _f32p), self.n, self.dim, leaf_max_size)

    def query(self, q: np.ndarray, k: int, nthreads: int = 1):
        q = np.ascontiguousarray(q, np.float32).reshape(-1, self.dim)
        nq = q.shape[0]
        idx = np.empty((nq, k), np.uint64)
        d2 = np.empty((nq, k), np.float32)
        ref().gloc_ref_knn_query_batch(self._h, _ptr(q, _f32p), nq, k, _ptr(idx, _u64p),
                                       _ptr(d2, _f32p), nthreads)
        return idx, d2

    def close(self):
        if self._h:
            ref().gloc_ref_knn_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ----------------------------------------------------------------- stage 2

def value_to_cost(v: int) -> np.float32:
    return np.float32(lib().gloc_oracle_value_to_cost(int(v)))


def cost_to_value(c: float) -> int:
    return int(lib().gloc_oracle_cost_to_value(float(np.float32(c))))


def cell_value(p: float) -> int:
    return int(lib().gloc_oracle_cell_value(float(np.float32(p))))


def level1_from_cells(cells: np.ndarray) -> np.ndarray:
    """cells: uint16 [ny, nx] (flat index nx*y + x).  Returns uint8 [ny, nx]."""
    cells = np.ascontiguousarray(cells, np.uint16)
    ny, nx = cells.shape
    out = np.empty((ny, nx), np.uint8)
    lib().gloc_oracle_level1_from_cells(cells.ctypes.data, nx, ny, out.ctypes.data)
    return out


def precomp_from_cells(cells: np.ndarray, width: int) -> np.ndarray:
    cells = np.ascontiguousarray(cells, np.uint16)
    ny, nx = cells.shape
    out = np.empty((ny + width - 1, nx + width - 1), np.uint8)
    lib().gloc_oracle_precomp_from_cells(cells.ctypes.data, nx, ny, width, out.ctypes.data)
    return out


def precomp_from_level1(level1: np.ndarray, width: int) -> np.ndarray:
    level1 = np.ascontiguousarray(level1, np.uint8)
    ny, nx = level1.shape
    out = np.empty((ny + width - 1, nx + width - 1), np.uint8)
    lib().gloc_oracle_precomp_from_level1(level1.ctypes.data, nx, ny, width, out.ctypes.data)
    return out


def search_params(lin_window: float, ang_window: float, pts: np.ndarray, resolution: float):
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
    nl, na, st = C.c_int(), C.c_int(), C.c_double()
    lib().gloc_oracle_search_params(lin_window, ang_window, _ptr(pts, _f32p), pts.shape[0],
                                    resolution, C.byref(nl), C.byref(na), C.byref(st))
    return nl.value, na.value, st.value


def grid_to_points(cells: np.ndarray, resolution: float, ox: float, oy: float) -> np.ndarray:
    cells = np.ascontiguousarray(cells, np.uint16)
    ny, nx = cells.shape
    n = lib().gloc_oracle_grid_to_points(cells.ctypes.data, nx, ny, resolution, ox, oy, None, 0)
    pts = np.zeros((n, 3), np.float32)
    lib().gloc_oracle_grid_to_points(cells.ctypes.data, nx, ny, resolution, ox, oy,
                                     pts.ctypes.data, n)
    return pts


def discretize(pts, init, n_ang, ang_step, resolution, max_x, max_y) -> np.ndarray:
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
    S = 2 * n_ang + 1
    out = np.empty((S, pts.shape[0], 2), np.int32)
    lib().gloc_oracle_discretize(_ptr(pts, _f32p), pts.shape[0], init[0], init[1], init[2],
                                 n_ang, ang_step, resolution, max_x, max_y, out.ctypes.data)
    return out


def csm_match(level1, resolution, max_x, max_y, depth, pts, init, n_lin, n_ang, ang_step,
              min_score, mode=0) -> MatchResult:
    level1 = np.ascontiguousarray(level1, np.uint8)
    ny, nx = level1.shape
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
    r = MatchResult()
    lib().gloc_oracle_csm_match(level1.ctypes.data, nx, ny, resolution, max_x, max_y, depth,
                                _ptr(pts, _f32p), pts.shape[0], init[0], init[1], init[2],
                                n_lin, n_ang, ang_step, min_score, mode, C.byref(r))
    return r


def csm_match_full_submap(level1, resolution, max_x, max_y, depth, pts, min_score,
                          mode=0) -> MatchResult:
    level1 = np.ascontiguousarray(level1, np.uint8)
    ny, nx = level1.shape
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
    r = MatchResult()
    lib().gloc_oracle_csm_match_full_submap(level1.ctypes.data, nx, ny, resolution, max_x,
                                            max_y, depth, _ptr(pts, _f32p), pts.shape[0],
                                            min_score, mode, C.byref(r))
    return r


def csm_match_batch(grids, resolution, max_x, max_y, depth, pts_list, inits, n_lin, n_ang,
                    ang_step, min_score, mode=0, nthreads=1):
    """grids: list of uint8 [ny,nx] (same shape); pts_list: list of [P_i,3] float32."""
    grids = [np.ascontiguousarray(g, np.uint8) for g in grids]
    pts_list = [np.ascontiguousarray(p, np.float32).reshape(-1, 3) for p in pts_list]
    n = len(grids)
    ny, nx = grids[0].shape
    gp = (C.c_void_p * n)(*[g.ctypes.data for g in grids])
    pp = (C.c_void_p * n)(*[p.ctypes.data for p in pts_list])
    npts = (C.c_int * n)(*[p.shape[0] for p in pts_list])
    init = np.ascontiguousarray(inits, np.float64).reshape(n, 3)
    out = (MatchResult * n)()
    lib().gloc_oracle_csm_match_batch_mt(gp, nx, ny, resolution, max_x, max_y, depth, pp, npts,
                                         init.ctypes.data_as(C.POINTER(C.c_double)), n, n_lin,
                                         n_ang, ang_step, min_score, mode, nthreads, out)
    return list(out)


def bev_project(pts: np.ndarray, resolution: float = 0.2, max_range: float = 100.0):
    """pts: [n, 3 or 4] float32.  Returns (img uint8 [h, w] with 0 = occupied / 255 = free,
    (ox, oy, resolution), (min_ix, min_iy), n_hit_voxels, n_occupied) as the reference's
    get_projected_grid (loop_detector.cpp:122-135) does for one scan."""
    pts = np.ascontiguousarray(pts, np.float32)
    n, stride = pts.shape
    w, h, mx, my = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    ox, oy = C.c_double(), C.c_double()
    nv, no = C.c_size_t(), C.c_size_t()
    args = (C.byref(w), C.byref(h), C.byref(mx), C.byref(my), C.byref(ox), C.byref(oy),
            C.byref(nv), C.byref(no))
    lib().gloc_oracle_bev_project(pts.ctypes.data, n, stride, resolution, max_range, None, 0, *args)
    img = np.empty((h.value, w.value), np.uint8)
    lib().gloc_oracle_bev_project(pts.ctypes.data, n, stride, resolution, max_range,
                                  img.ctypes.data, img.size, *args)
    res = float(np.float32(resolution))
    return img, (ox.value, oy.value, res), (mx.value, my.value), nv.value, no.value


def crop_pad(img: np.ndarray, width: int = 768, height: int = 768) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty((height, width), np.uint8)
    lib().gloc_oracle_crop_pad(img.ctypes.data, img.shape[1], img.shape[0], width, height,
                               out.ctypes.data)
    return out


# ------------------------------------------------- stage 2: the REFERENCE's own matcher
# oracle/_ref/libcsm_ref.so = registration/2d/*.cpp (+ 3d/probability_values.cpp, point_cloud.cpp)
# compiled unmodified from /root/reference against oracle/shim/ (see oracle/csm_ref.cpp).

_CSM_REF = os.path.join(_HERE, "_ref", "libcsm_ref.so")
_csm_ref = None


class RefMatchResult(C.Structure):
    _fields_ = [("found", C.c_int), ("score", C.c_float), ("pose_x", C.c_double), ("pose_y", C.c_double),
                ("pose_yaw", C.c_double), ("scan_index", C.c_int), ("x_offset", C.c_int),
                ("y_offset", C.c_int), ("cand_score", C.c_float)]


def have_csm_ref() -> bool:
    if not os.path.exists(_CSM_REF):
        try:
            build()
        except Exception:
            return False
    return os.path.exists(_CSM_REF)


def csm_ref():
    global _csm_ref
    if _csm_ref is None:
        if not have_csm_ref():
            raise RuntimeError("oracle/_ref/libcsm_ref.so is missing")
        R = C.CDLL(_CSM_REF)
        d, i, f, vp = C.c_double, C.c_int, C.c_float, C.c_void_p
        R.gloc_ref_value_to_cost.restype = f
        R.gloc_ref_value_to_cost.argtypes = [C.c_uint16]
        R.gloc_ref_cost_to_value.restype = C.c_uint16
        R.gloc_ref_cost_to_value.argtypes = [f]
        R.gloc_ref_csm_precomp.argtypes = [vp, i, i, d, d, d, i, i, vp]
        R.gloc_ref_csm_search_params.argtypes = [d, d, vp, i, d, C.POINTER(i), C.POINTER(i), C.POINTER(d)]
        R.gloc_ref_csm_discretize.argtypes = [vp, i, d, d, d, i, d, d, d, d, vp]
        R.gloc_ref_csm_match.restype = i
        R.gloc_ref_csm_match.argtypes = [vp, i, i, d, d, d, i, vp, i, d, d, d, i, i, d, f, C.POINTER(RefMatchResult)]
        R.gloc_ref_csm_match_grid.restype = i
        R.gloc_ref_csm_match_grid.argtypes = [vp, i, i, d, d, d, i, d, d, vp, i, i, d, d, d, d, d, d, d, f,
                                              C.POINTER(RefMatchResult)]
        R.gloc_ref_csm_match_full_submap.restype = i
        R.gloc_ref_csm_match_full_submap.argtypes = [vp, i, i, d, d, d, i, vp, i, f, C.POINTER(RefMatchResult)]
        R.gloc_ref_csm_match_batch_mt.argtypes = [C.POINTER(vp), i, i, d, d, d, i, C.POINTER(vp), C.POINTER(i),
                                                  C.POINTER(d), i, i, i, d, f, i, C.POINTER(RefMatchResult)]
        _csm_ref = R
    return _csm_ref


def ref_precomp(cells: np.ndarray, depth: int, index: int) -> np.ndarray:
    cells = np.ascontiguousarray(cells, np.uint16)
    ny, nx = cells.shape
    w = 1 << index
    out = np.empty((ny + w - 1, nx + w - 1), np.uint8)
    csm_ref().gloc_ref_csm_precomp(cells.ctypes.data, nx, ny, 0.2, 10.0, 10.0, depth, index, out.ctypes.data)
    return out


def ref_search_params(lin_window, ang_window, pts, resolution):
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
    nl, na, st = C.c_int(), C.c_int(), C.c_double()
    csm_ref().gloc_ref_csm_search_params(lin_window, ang_window, pts.ctypes.data, pts.shape[0], resolution,
                                         C.byref(nl), C.byref(na), C.byref(st))
    return nl.value, na.value, st.value


def ref_discretize(pts, init, n_ang, ang_step, resolution, max_x, max_y) -> np.ndarray:
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
    out = np.empty((2 * n_ang + 1, pts.shape[0], 2), np.int32)
    csm_ref().gloc_ref_csm_discretize(pts.ctypes.data, pts.shape[0], init[0], init[1], init[2], n_ang, ang_step,
                                      resolution, max_x, max_y, out.ctypes.data)
    return out


def ref_csm_match(cells, resolution, max_x, max_y, depth, pts, init, n_lin, n_ang, ang_step,
                  min_score) -> RefMatchResult:
    """cells: Grid2D's uint16 correspondence-cost cells [ny, nx]."""
    cells = np.ascontiguousarray(cells, np.uint16)
    ny, nx = cells.shape
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
    r = RefMatchResult()
    csm_ref().gloc_ref_csm_match(cells.ctypes.data, nx, ny, resolution, max_x, max_y, depth, pts.ctypes.data,
                                 pts.shape[0], init[0], init[1], init[2], n_lin, n_ang, ang_step, min_score,
                                 C.byref(r))
    return r


def ref_csm_match_full_submap(cells, resolution, max_x, max_y, depth, pts, min_score) -> RefMatchResult:
    cells = np.ascontiguousarray(cells, np.uint16)
    ny, nx = cells.shape
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
    r = RefMatchResult()
    csm_ref().gloc_ref_csm_match_full_submap(cells.ctypes.data, nx, ny, resolution, max_x, max_y, depth,
                                             pts.ctypes.data, pts.shape[0], min_score, C.byref(r))
    return r


def ref_csm_match_batch(cells_list, resolution, max_x, max_y, depth, pts_list, inits, n_lin, n_ang, ang_step,
                        min_score, nthreads=1):
    """The reference matcher over independent (grid, scan) pairs, one matcher per pair."""
    cells_list = [np.ascontiguousarray(c, np.uint16) for c in cells_list]
    pts_list = [np.ascontiguousarray(p, np.float32).reshape(-1, 3) for p in pts_list]
    n = len(cells_list)
    ny, nx = cells_list[0].shape
    gp = (C.c_void_p * n)(*[c.ctypes.data for c in cells_list])
    pp = (C.c_void_p * n)(*[p.ctypes.data for p in pts_list])
    npts = (C.c_int * n)(*[p.shape[0] for p in pts_list])
    init = np.ascontiguousarray(inits, np.float64).reshape(n, 3)
    out = (RefMatchResult * n)()
    csm_ref().gloc_ref_csm_match_batch_mt(gp, nx, ny, resolution, max_x, max_y, depth, pp, npts,
                                          init.ctypes.data_as(C.POINTER(C.c_double)), n, n_lin, n_ang, ang_step,
                                          min_score, nthreads, out)
    return list(out)


# ------------------------------------------------- BEV projection: the REFERENCE's own code
# oracle/_ref/libbev_ref.so = 3d/submap_3d.cpp (Submap3D, ProjectToCvMat), 3d/range_data_inserter_3d.cpp,
# 3d/range_data.cpp, 3d/hybrid_grid.h ... compiled unmodified from /root/reference against oracle/shim/
# (see oracle/bev_ref.cpp).
_BEV_REF = os.path.join(_HERE, "_ref", "libbev_ref.so")
_bev_ref = None


def have_bev_ref() -> bool:
    if not os.path.exists(_BEV_REF):
        try:
            build(force=os.path.isdir("/root/reference/registration"))
        except Exception:
            return False
    return os.path.exists(_BEV_REF)


def ref_bev_project(pts: np.ndarray):
    """The reference's get_projected_grid for one scan ([n, 3 or 4] float32): (img uint8 [h, w] with
    0 = occupied / 255 = free, (ox, oy, resolution))."""
    global _bev_ref
    if _bev_ref is None:
        if not have_bev_ref():
            raise RuntimeError("oracle/_ref/libbev_ref.so is missing")
        R = C.CDLL(_BEV_REF)
        R.gloc_ref_bev_project.restype = C.c_int
        R.gloc_ref_bev_project.argtypes = ([C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t] +
                                           [C.POINTER(C.c_int)] * 2 + [C.POINTER(C.c_double)] * 3)
        _bev_ref = R
    pts = np.ascontiguousarray(pts, np.float32)
    n, stride = pts.shape
    w, h = C.c_int(), C.c_int()
    ox, oy, res = C.c_double(), C.c_double(), C.c_double()
    args = (C.byref(w), C.byref(h), C.byref(ox), C.byref(oy), C.byref(res))
    _bev_ref.gloc_ref_bev_project(pts.ctypes.data, n, stride, None, 0, *args)
    img = np.empty((h.value, w.value), np.uint8)
    if _bev_ref.gloc_ref_bev_project(pts.ctypes.data, n, stride, img.ctypes.data, img.size, *args) != 0:
        raise RuntimeError("gloc_ref_bev_project failed")
    return img, (ox.value, oy.value, res.value)
